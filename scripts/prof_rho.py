#!/usr/bin/env python
"""rho (+ derivative codes given on the command line) on the benchmark molecule, device-resident."""
import os, sys
import numpy, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import orbkit_b200 as ok
from orbkit_b200 import synth
from orbkit_b200._lib import OKB_FLAG_OUT_DEVICE
from orbkit_b200.engine import get_engine
eng = get_engine()
dev = torch.device('cuda', eng.device)
# PROF_CASE=c2: the Config-2 shape with the 21 occupied orbitals (222 AOs, 150^3); default: the benchmark molecule (200^3)
if os.environ.get('PROF_CASE') == 'c2':
    qc = synth.to_qcinfo(synth.make_molecule(n_heavy=6, n_light=3, n_mo=21, seed=0, spherical=True))
    ax = numpy.linspace(-12, 12, 150)
else:
    qc = synth.to_qcinfo(synth.make_molecule(n_heavy=24, n_light=20, n_mo=82, seed=0, spherical=True))
    ax = numpy.linspace(-12, 12, 200)
basis = eng.basis(qc.geo_spec, qc.ao_spec)
mo = eng.mos_of(basis, qc.mo_spec)
g = eng.grid_regular(ax, ax, ax)
codes = [int(a) for a in sys.argv[1:]]
out = torch.zeros((8, len(ax) ** 3), dtype=torch.float64, device=dev)
for _ in range(3):
    eng.eval_rho(mo, g, codes, rho=out[0].data_ptr(), delta=out[1:].data_ptr() if codes else None, flags=OKB_FLAG_OUT_DEVICE)
eng.sync()
print('done', eng.last_kernel())
