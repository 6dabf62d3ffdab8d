"""Minimal stand-in for orbkit/display.py:27-42: print unless options.quiet."""
from . import options


def display(string):
    if not options.quiet:
        print(string)
