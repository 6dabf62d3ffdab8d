#!/usr/bin/env python
"""cProfile of the small-call path: rho_compute(qc) on a NEW 1000-point vector grid per call (bench.py 'latency' block)."""
import cProfile, os, pstats, sys, time
import numpy
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import orbkit_b200 as ok
from orbkit_b200 import synth, grid, options
options.quiet = True
qc = synth.to_qcinfo(synth.make_molecule(n_heavy=24, n_light=20, n_mo=82, seed=0, spherical=True))
rng = numpy.random.default_rng(3)
pts = [rng.uniform(-8, 8, size=(3, 1000)) for _ in range(300)]
def call(p):
    grid.set_grid(p[0], p[1], p[2], is_vector=True)
    return ok.rho_compute(qc)
for p in pts[:50]:
    call(p)
t0 = time.perf_counter()
for p in pts[50:250]:
    call(p)
print('us per call: %.1f' % ((time.perf_counter() - t0) / 200 * 1e6))
pr = cProfile.Profile(); pr.enable()
for p in pts[50:250]:
    call(p)
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(18)
