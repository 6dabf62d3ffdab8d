#!/usr/bin/env python
"""Device-resident timing of the time-dependent detCI contraction okb_ci_td (cy_ci.get_rho_full / get_j_full):
out[t][x] = sum_k w[t][k] in[k][x] for several (nt, nstate) at 128^3 points, with the HBM / FP64 rooflines.
usage: perf_td.py [one]    ('one': a single launch of the largest case, for ncu)"""
import os, sys, json
import numpy, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from orbkit_b200._lib import OKB_FLAG_OUT_DEVICE, OKB_FLAG_IN_DEVICE
from orbkit_b200.engine import get_engine
eng = get_engine()
dev = torch.device('cuda', eng.device)
stream = torch.cuda.ExternalStream(eng.stream_ptr(), device=dev)
peaks = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))
hbm = peaks['hbm_gbs']
fp64 = 37.0                                   # TFLOP/s, DMMA issue bound measured by bench.py (okb_measure_fp64)
npts = 128 ** 3
rng = numpy.random.default_rng(0)
cases = [(64, 2), (64, 4), (256, 4), (256, 7), (512, 10), (128, 16), (512, 13), (512, 17), (64, 15), (82, 44)]
if len(sys.argv) > 1 and sys.argv[1] == 'one':
    cases = [tuple(int(v) for v in os.environ.get('TD_CASE', '512:10').split(':'))]
for nt, ns in cases:
    nk = ns * (ns + 1) // 2
    n = npts if nt * npts * 8 < 40e9 else npts // 4
    w = rng.normal(size=(nt, nk))
    data = torch.randn((nk, n), dtype=torch.float64, device=dev)
    out = torch.empty((nt, n), dtype=torch.float64, device=dev)
    f = lambda: eng.ci_td(w, data.data_ptr(), out=out.data_ptr(), nk=nk, n=n, ld_in=n, ld_out=n,
                          flags=OKB_FLAG_IN_DEVICE | OKB_FLAG_OUT_DEVICE)
    f(); eng.sync()
    if len(sys.argv) > 1 and sys.argv[1] == 'one':
        break
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream); [f() for _ in range(3)]; e1.record(stream)
    eng.sync()
    ms = e0.elapsed_time(e1) / 3
    by = 8.0 * (nk + nt) * n
    fl = 2.0 * nt * nk * n
    t_min = max(by / (hbm * 1e9), fl / (fp64 * 1e12)) * 1e3
    ref = (torch.from_numpy(w).to(dev) @ data[:, :1000])
    err = float((out[:, :1000] - ref).abs().max() / ref.abs().max())
    print('nt %4d nstate %2d (nk %3d) n %8d  %8.3f ms  %7.1f GB/s  %6.2f TFLOP/s  roofline %s frac %.2f  relerr %.1e' % (
        nt, ns, nk, n, ms, by / ms / 1e6, fl / ms / 1e9, 'hbm' if by / (hbm * 1e9) > fl / (fp64 * 1e12) else 'fp64',
        t_min / ms, err))
    del data, out
