#!/usr/bin/env python
"""config-2-shaped timing (222 AOs, 222 MOs, 150^3 points): which MO-tile width the host picks and what it costs"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'scripts')); sys.path.insert(0, REPO)
import perf_matrix as pm
pm.run('c2 n_mo=222', 222, N=150, heavy=6, light=3)
