#!/usr/bin/env python
"""rho + laplacian on the benchmark molecule, device resident (for ncu captures of the two passes)."""
import os, sys
import numpy, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from orbkit_b200 import synth
from orbkit_b200._lib import OKB_FLAG_OUT_DEVICE
from orbkit_b200.engine import get_engine
eng = get_engine()
dev = torch.device('cuda', eng.device)
qc = synth.to_qcinfo(synth.make_molecule(n_heavy=24, n_light=20, n_mo=82, seed=0, spherical=True))
ax = numpy.linspace(-12, 12, 200)
basis = eng.basis(qc.geo_spec, qc.ao_spec)
mo = eng.mos_of(basis, qc.mo_spec)
g = eng.grid_regular(ax, ax, ax)
n = 200 ** 3
out = torch.zeros((4, n), dtype=torch.float64, device=dev)
for _ in range(2):
    eng.eval_rho(mo, g, [4, 5, 6], 0, n, rho=out[0].data_ptr(), delta=out[1:].data_ptr(), flags=OKB_FLAG_OUT_DEVICE)
eng.sync()
print('done', eng.last_kernel())
