#!/usr/bin/env python
"""rho_compute on vector grids of a few thousand points (benchmark molecule): ms per call."""
import os, sys, time
import numpy
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import orbkit_b200 as ok
from orbkit_b200 import synth, grid, options
options.quiet = True
qc = synth.to_qcinfo(synth.make_molecule(n_heavy=24, n_light=20, n_mo=82, seed=0, spherical=True))
rng = numpy.random.default_rng(3)
for n in (1000, 4000, 6000, 10000, 14000, 18000, 24000):
    p = rng.uniform(-8, 8, size=(3, n))
    grid.set_grid(p[0], p[1], p[2], is_vector=True)
    for _ in range(5):
        ok.rho_compute(qc)
    t0 = time.perf_counter()
    for _ in range(50):
        ok.rho_compute(qc)
    print('n %6d  %8.1f us per call' % (n, (time.perf_counter() - t0) / 50 * 1e6), flush=True)
