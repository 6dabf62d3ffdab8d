#!/usr/bin/env python
"""calc_ao (SINK_AO) device-resident timing on the benchmark molecule and a Cartesian / small-basis variant; one line
per case: ms, points/s, GB/s stored, kernel variant.  OKB_AO_VARIANT=<substring> selects a kernel variant (A/B)."""
import os, sys
import numpy, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import orbkit_b200 as ok
from orbkit_b200 import synth
from orbkit_b200._lib import OKB_FLAG_OUT_DEVICE
from orbkit_b200.engine import get_engine
ok.options.quiet = True
eng = get_engine()
dev = torch.device('cuda', eng.device)
stream = torch.cuda.ExternalStream(eng.stream_ptr(), device=dev)

def timed(fn, reps=5):
    fn(); fn(); eng.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(reps): fn()
        e1.record(stream)
    eng.sync()
    return e0.elapsed_time(e1) / reps

for name, heavy, light, sph, codes in (('c3 sph', 24, 20, True, [0]), ('c3 cart', 24, 20, False, [0]), ('c2-like', 6, 6, True, [0]),
                                       ('c3 sph d/dx', 24, 20, True, [1]), ('c3 sph d/dz', 24, 20, True, [3]), ('c3 sph d2/dx2', 24, 20, True, [4]),
                                       ('c3 sph d2/dz2', 24, 20, True, [6]), ('c3 sph grad', 24, 20, True, [1, 2, 3]), ('c3 sph 7 sets', 24, 20, True, [0, 1, 2, 3, 4, 5, 6]),
                                       ('c3 sph 10 sets', 24, 20, True, list(range(10)))):
    qc = synth.to_qcinfo(synth.make_molecule(n_heavy=heavy, n_light=light, n_mo=8, seed=0, spherical=sph))
    n_ao = qc.ao_spec.get_ao_num()
    ax = numpy.linspace(-12, 12, 200)
    basis = eng.basis(qc.geo_spec, qc.ao_spec)
    g = eng.grid_regular(ax, ax, ax)
    nsub = min(8000000, int(8e9 // (8 * n_ao * len(codes)))) // 1024 * 1024
    buf = torch.empty((len(codes), n_ao, nsub), dtype=torch.float64, device=dev)
    ms = timed(lambda: eng.eval_ao(basis, g, codes, 0, nsub, out=buf.data_ptr(), flags=OKB_FLAG_OUT_DEVICE))
    print('%-12s n_ao %-5d pts %-8d %8.3f ms  %.3e pts/s  %6.0f GB/s out  %s' % (
        name, n_ao, nsub, ms, nsub / ms * 1e3, 8.0 * n_ao * len(codes) * nsub / (ms * 1e-3) / 1e9, eng.last_kernel()))
    del buf
