#!/usr/bin/env python
"""Config-2 shape with the 21 occupied orbitals (222 AOs, 150^3 points): rho / rho+grad / rho+lap, device resident.
OKB_VARIANT=<substring> forces a kernel variant for the sets it matches."""
import os, sys
import numpy, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from orbkit_b200 import synth
from orbkit_b200._lib import OKB_FLAG_OUT_DEVICE
from orbkit_b200.engine import get_engine
eng = get_engine(); dev = torch.device('cuda', eng.device)
stream = torch.cuda.ExternalStream(eng.stream_ptr(), device=dev)
n_mo = int(os.environ.get('NMO', '21'))
qc = synth.to_qcinfo(synth.make_molecule(n_heavy=6, n_light=3, n_mo=n_mo, seed=0, spherical=True))
N = 150
ax = numpy.linspace(-12, 12, N)
basis = eng.basis(qc.geo_spec, qc.ao_spec); mo = eng.mos_of(basis, qc.mo_spec); g = eng.grid_regular(ax, ax, ax)
n_ao = qc.ao_spec.get_ao_num()
out = torch.zeros((8, N ** 3), dtype=torch.float64, device=dev)
for label, codes, D in (('rho', [], 1), ('rho+grad', [1, 2, 3], 4), ('rho+lap', [4, 5, 6], 7)):
    f = lambda: eng.eval_rho(mo, g, codes, rho=out[0].data_ptr(), delta=out[1:].data_ptr() if codes else None, flags=OKB_FLAG_OUT_DEVICE)
    f(); f(); eng.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream); [f() for _ in range(5)]; e1.record(stream)
    eng.sync()
    ms = e0.elapsed_time(e1) / 5
    print('c2 n_mo=%d %-9s %8.3f ms  %6.2f TFLOP/s alg  checksum %.12e  %s' % (
        n_mo, label, ms, 2.0 * n_mo * n_ao * D * N ** 3 / ms / 1e9, float(out[:1 + len(codes)].sum()), eng.last_kernel()), flush=True)
