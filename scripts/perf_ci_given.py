#!/usr/bin/env python
"""The reference's own detCI call pattern: molist = rho_compute(qc, calc_mo=True) once on the host, then
ci_core.rho / jab(zero, sing, molist[, molistdrv]) per state pair.  300 MOs, 30 active orbitals, 20000 terms, 64^3 points,
NumPy arrays in and out; options.ci_fast None (reference order, bit-identical) and True (re-ordered device sums)."""
import os, sys, time
import numpy
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import orbkit_b200 as ok
from orbkit_b200.detci import ci_core
from orbkit_b200.engine import get_engine
ok.options.quiet = True
rng = numpy.random.default_rng(7)
n_mo, npts = 300, 64 ** 3
mo = rng.normal(size=(n_mo, npts))
dmo = rng.normal(size=(3, n_mo, npts))
act = numpy.sort(rng.choice(n_mo, size=30, replace=False))
pairs = act[rng.integers(0, 30, size=(20000, 2))]
zero = [[], []]
sing = [list(rng.normal(size=20000)), [list(map(int, p)) for p in pairs]]
res = {}
for fast in (None, True):
    ok.options.ci_fast = fast
    for name, f in (('rho', lambda: ci_core.rho(zero, sing, mo, slice_length=npts)),
                    ('jab', lambda: ci_core.jab(zero, sing, mo, dmo, slice_length=npts))):
        f()
        t0 = time.perf_counter()
        for _ in range(3):
            out = f()
        dt = (time.perf_counter() - t0) / 3
        h2d = get_engine().traffic()[0]
        f()
        h2d = get_engine().traffic()[0] - h2d
        res[(name, fast)] = out.copy()
        print('ci_fast=%-5s %-4s %9.2f ms   H2D %.1f MB per call   last kernel %s' % (fast, name, dt * 1e3, h2d / 1e6,
                                                                                  get_engine().last_kernel()), flush=True)
for name in ('rho', 'jab'):
    a, b = res[(name, None)], res[(name, True)]
    print('%-4s max |fast - exact| = %.3e, max |exact| = %.3e' % (name, numpy.abs(a - b).max(), numpy.abs(a).max()))
