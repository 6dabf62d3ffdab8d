#!/usr/bin/env python
"""detCI pair contraction on device-resident MOs of the Config-5 molecule (for ncu): args = mode (rho|jab)"""
import os, sys
import numpy, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from orbkit_b200 import synth, _lib
from orbkit_b200._lib import OKB_FLAG_OUT_DEVICE, OKB_FLAG_IN_DEVICE
from orbkit_b200.engine import get_engine
eng = get_engine()
dev = torch.device('cuda', eng.device)
qc = synth.to_qcinfo(synth.make_molecule(n_heavy=12, n_light=10, n_mo=500, seed=5, spherical=True))
rng = numpy.random.default_rng(5)
pairs = rng.integers(0, 500, size=(1000, 2))
terms = (rng.normal(size=1000), pairs[:, 0].astype(numpy.intc), pairs[:, 1].astype(numpy.intc))
ax = numpy.linspace(-10, 10, 96)
basis = eng.basis(qc.geo_spec, qc.ao_spec)
mo = eng.mos_of(basis, qc.mo_spec)
g = eng.grid_regular(ax, ax, ax)
n = 96 ** 3
buf = torch.empty((4, 500, n), dtype=torch.float64, device=dev)
eng.eval_mo(mo, g, [0, 1, 2, 3], 0, n, out=buf.data_ptr(), flags=OKB_FLAG_OUT_DEVICE)
out = torch.zeros((3, n), dtype=torch.float64, device=dev)
mode = _lib.OKB_CI_JAB if (len(sys.argv) > 1 and sys.argv[1] == 'jab') else _lib.OKB_CI_RHO
for _ in range(3):
    eng.ci_contract(mode, terms, buf[0].data_ptr(), buf[1:].data_ptr(), n_mo=500, npts=n, ld=n, out=out.data_ptr(),
                    flags=OKB_FLAG_OUT_DEVICE | OKB_FLAG_IN_DEVICE)
eng.sync()
print('done', eng.last_kernel())
