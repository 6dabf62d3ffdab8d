#!/usr/bin/env python
"""bench.py's order of events with per-step end-to-end timings (why is e2e slower inside bench.py?)"""
import os, sys, time
import numpy, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import orbkit_b200 as ok
from orbkit_b200 import synth
from orbkit_b200._lib import OKB_FLAG_OUT_DEVICE
from orbkit_b200.engine import get_engine
ok.options.quiet = True
eng = get_engine()
dev = torch.device('cuda', 0)
qc = synth.to_qcinfo(synth.make_molecule(n_heavy=24, n_light=20, n_mo=82, seed=0, spherical=True))
ax = numpy.linspace(-12, 12, 200)
basis = eng.basis(qc.geo_spec, qc.ao_spec)
mo = eng.mos_of(basis, qc.mo_spec)
g = eng.grid_regular(ax, ax, ax)
out = torch.zeros((4, 8000000), dtype=torch.float64, device=dev)
for _ in range(4):
    eng.eval_rho(mo, g, [1, 2, 3], 0, 8000000, rho=out[0].data_ptr(), delta=out[1:].data_ptr(), flags=OKB_FLAG_OUT_DEVICE)
eng.sync()
if os.environ.get('PROBE_PEAKS', '1') == '1':
    print('fp64', eng.measure_fp64(0, 1.0)[0], eng.measure_fp64(1, 1.0)[0])
ok.grid.set_grid(ax, ax, ax, is_vector=False)
for i in range(10):
    t0 = time.perf_counter(); eng.clear_caches(); t1 = time.perf_counter()
    r = ok.rho_compute(qc, drv=['x', 'y', 'z']); t2 = time.perf_counter()
    print('e2e step %d: clear %.1f ms, rho_compute %.1f ms' % (i, 1e3 * (t1 - t0), 1e3 * (t2 - t1)))
