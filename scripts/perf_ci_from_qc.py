#!/usr/bin/env python
"""detci.ci_core.*_from_qc end to end (host arrays out) for a CI-shaped case: 300 MOs, 30 active orbitals, 20000 terms,
96^3 points -- default (active orbitals only + re-ordered device sums) against options.ci_fast = False (reference order)."""
import os, sys, time
import numpy
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import orbkit_b200 as ok
from orbkit_b200 import synth
from orbkit_b200.detci import ci_core
from orbkit_b200.engine import get_engine
ok.options.quiet = True
qc = synth.to_qcinfo(synth.make_molecule(n_heavy=12, n_light=10, n_mo=300, seed=5, spherical=True))
rng = numpy.random.default_rng(7)
act = numpy.sort(rng.choice(300, size=30, replace=False))
pairs = act[rng.integers(0, 30, size=(20000, 2))]
zero = [[], []]
sing = [list(rng.normal(size=20000)), [list(map(int, p)) for p in pairs]]
ax = numpy.linspace(-10, 10, 96)
ok.grid.set_grid(ax, ax, ax, is_vector=False)
res = {}
for fast in (None, False):
    ok.options.ci_fast = fast
    for name, f in (('rho', ci_core.rho_from_qc), ('jab', ci_core.jab_from_qc), ('a_nabla_b', ci_core.a_nabla_b_from_qc)):
        f(qc, zero, sing)
        t0 = time.perf_counter()
        for _ in range(3):
            out = f(qc, zero, sing)
        dt = (time.perf_counter() - t0) / 3
        res[(name, fast)] = out.copy()
        print('ci_fast=%-5s %-10s %9.2f ms   last kernel %s' % (fast, name, dt * 1e3, get_engine().last_kernel()), flush=True)
for name in ('rho', 'jab', 'a_nabla_b'):
    a, b = res[(name, None)], res[(name, False)]
    print('%-10s max |fast - exact| = %.3e, max |exact| = %.3e' % (name, numpy.abs(a - b).max(), numpy.abs(b).max()))
