#!/usr/bin/env python
"""detCI pair contraction (Config 5: 500 MOs, 1000 pairs, 96^3 points, device-resident MOs): staged kernel against the
gather kernel (OKB_CI_GATHER=1), timing + bit-identity of the two."""
import os, sys, json
import numpy, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from orbkit_b200 import synth, _lib
from orbkit_b200._lib import OKB_FLAG_OUT_DEVICE, OKB_FLAG_IN_DEVICE
from orbkit_b200.engine import get_engine
eng = get_engine(); dev = torch.device('cuda', eng.device)
stream = torch.cuda.ExternalStream(eng.stream_ptr(), device=dev)
qc = synth.to_qcinfo(synth.make_molecule(n_heavy=12, n_light=10, n_mo=500, seed=5, spherical=True))
rng = numpy.random.default_rng(5)
pairs = rng.integers(0, 500, size=(1000, 2))
terms = (rng.normal(size=1000), pairs[:, 0].astype(numpy.intc), pairs[:, 1].astype(numpy.intc))
ax = numpy.linspace(-10, 10, 96)
basis = eng.basis(qc.geo_spec, qc.ao_spec); mo = eng.mos_of(basis, qc.mo_spec); g = eng.grid_regular(ax, ax, ax)
n = 96 ** 3
buf = torch.empty((4, 500, n), dtype=torch.float64, device=dev)
eng.eval_mo(mo, g, [0, 1, 2, 3], 0, n, out=buf.data_ptr(), flags=OKB_FLAG_OUT_DEVICE)
out = torch.zeros((3, n), dtype=torch.float64, device=dev)
hbm = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))['hbm_gbs']
for name, mode, sets in (('rho', _lib.OKB_CI_RHO, 1), ('jab', _lib.OKB_CI_JAB, 4)):
    f = lambda: eng.ci_contract(mode, terms, buf[0].data_ptr(), buf[1:].data_ptr(), n_mo=500, npts=n, ld=n, out=out.data_ptr(),
                                flags=OKB_FLAG_OUT_DEVICE | OKB_FLAG_IN_DEVICE)
    f(); eng.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream); [f() for _ in range(5)]; e1.record(stream)
    eng.sync()
    ms = e0.elapsed_time(e1) / 5
    by = 8.0 * 500 * sets * n + 8.0 * (1 if sets == 1 else 3) * n
    print('%s %-22s %8.3f ms  %7.1f GB/s algorithmic = %.2f of the HBM peak  checksum %.12e' % (
        name, eng.last_kernel(), ms, by / ms / 1e6, by / ms / 1e6 / hbm, float(out[:1 if sets == 1 else 3].double().sum())))
    numpy.save(os.path.join(REPO, 'gpurun_out', 'ci_%s_%s.npy' % (name, 'gather' if os.environ.get('OKB_CI_GATHER') else 'staged')),
               out[:1 if sets == 1 else 3, ::97].cpu().numpy())
