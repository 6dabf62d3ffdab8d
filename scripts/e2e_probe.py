#!/usr/bin/env python
"""Where does the end-to-end time of rho_compute go when the handle caches are dropped every step?"""
import os, sys, time
import numpy, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import orbkit_b200 as ok
from orbkit_b200 import synth
from orbkit_b200.engine import get_engine
ok.options.quiet = True
eng = get_engine()
qc = synth.to_qcinfo(synth.make_molecule(n_heavy=24, n_light=20, n_mo=82, seed=0, spherical=True))
ax = numpy.linspace(-12, 12, 200)
ok.grid.set_grid(ax, ax, ax, is_vector=False)
for i in range(8):
    t = [time.perf_counter()]
    eng.clear_caches(); t.append(time.perf_counter())
    basis = eng.basis(qc.geo_spec, qc.ao_spec); t.append(time.perf_counter())
    mo = eng.mos_of(basis, qc.mo_spec); t.append(time.perf_counter())
    g = eng.grid_regular(ax, ax, ax); t.append(time.perf_counter())
    rho = eng.host_array((8000000,)); d = eng.host_array((3, 8000000)); t.append(time.perf_counter())
    eng.eval_rho(mo, g, [1, 2, 3], rho=rho, delta=d); t.append(time.perf_counter())
    del rho, d; t.append(time.perf_counter())
    print('step %d [ms]: clear %.1f basis %.1f mos %.1f grid %.1f host_alloc %.1f eval %.1f del %.1f' %
          ((i,) + tuple(1e3 * (b - a) for a, b in zip(t[:-1], t[1:]))))
for i in range(6):
    eng.clear_caches()
    t0 = time.perf_counter(); r = ok.rho_compute(qc, drv=['x', 'y', 'z']); t1 = time.perf_counter()
    print('rho_compute, caches dropped: %.1f ms' % (1e3 * (t1 - t0)))
