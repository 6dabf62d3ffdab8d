#!/usr/bin/env python
"""calc_ao on the benchmark molecule, device-resident (for ncu captures of the SINK_AO kernel)."""
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
qc = synth.to_qcinfo(synth.make_molecule(n_heavy=24, n_light=20, n_mo=82, seed=0, spherical=True))
ax = numpy.linspace(-12, 12, 200)
basis = eng.basis(qc.geo_spec, qc.ao_spec)
g = eng.grid_regular(ax, ax, ax)
nsub = 1000000
buf = torch.empty((1, 1000, nsub), dtype=torch.float64, device=dev)
codes = [int(sys.argv[1])] if len(sys.argv) > 1 else [0]
for _ in range(3):
    eng.eval_ao(basis, g, codes, 4000000, 4000000 + nsub, out=buf.data_ptr(), flags=OKB_FLAG_OUT_DEVICE)
eng.sync()
print('done', eng.last_kernel())
