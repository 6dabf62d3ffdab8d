#!/usr/bin/env python
"""detCI pair contraction (Config 5: 500 MOs, 1000 pairs, 96^3 points, device-resident MOs): the bit-identical gather kernel
against the OKB_FLAG_CI_FAST kernel (terms split over warps, rows staged once): timing and agreement; also a CI-like
case with few active MOs and many terms (40 MOs, 20000 terms)."""
import os, sys, json
import numpy, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from orbkit_b200 import synth, _lib
from orbkit_b200._lib import OKB_FLAG_OUT_DEVICE, OKB_FLAG_IN_DEVICE, OKB_FLAG_CI_FAST
from orbkit_b200.engine import get_engine
eng = get_engine(); dev = torch.device('cuda', eng.device)
stream = torch.cuda.ExternalStream(eng.stream_ptr(), device=dev)
hbm = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))['hbm_gbs']
n = 96 ** 3
REPS = int(os.environ.get("REPS", "5"))
ax = numpy.linspace(-10, 10, 96)
CASES = [tuple(int(v) for v in c.split(':')) for c in os.environ.get('CASES', '500:1000,40:20000,120:5000').split(',')]
for n_mo, n_terms in CASES:
    qc = synth.to_qcinfo(synth.make_molecule(n_heavy=12, n_light=10, n_mo=n_mo, seed=5, spherical=True))
    rng = numpy.random.default_rng(5)
    pairs = rng.integers(0, n_mo, size=(n_terms, 2))
    terms = (rng.normal(size=n_terms), pairs[:, 0].astype(numpy.intc), pairs[:, 1].astype(numpy.intc))
    basis = eng.basis(qc.geo_spec, qc.ao_spec); mo = eng.mos_of(basis, qc.mo_spec); g = eng.grid_regular(ax, ax, ax)
    buf = torch.empty((4, n_mo, n), dtype=torch.float64, device=dev)
    eng.eval_mo(mo, g, [0, 1, 2, 3], 0, n, out=buf.data_ptr(), flags=OKB_FLAG_OUT_DEVICE)
    res = {}
    for name, mode, sets in (('rho', _lib.OKB_CI_RHO, 1), ('jab', _lib.OKB_CI_JAB, 4), ('anb', _lib.OKB_CI_A_NABLA_B, 4)):
        for fast in (0, OKB_FLAG_CI_FAST):
            out = torch.zeros((3, n), dtype=torch.float64, device=dev)
            f = lambda: eng.ci_contract(mode, terms, buf[0].data_ptr(), buf[1:].data_ptr(), n_mo=n_mo, npts=n, ld=n,
                                        out=out.data_ptr(), flags=OKB_FLAG_OUT_DEVICE | OKB_FLAG_IN_DEVICE | fast)
            f(); eng.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record(stream); [f() for _ in range(REPS)]; e1.record(stream)
            eng.sync()
            ms = e0.elapsed_time(e1) / REPS
            nc = 1 if sets == 1 else 3
            by = 8.0 * n_mo * sets * n + 8.0 * nc * n
            res[(name, fast)] = out[:nc].clone()
            print('n_mo %4d terms %6d %s %-22s %8.3f ms  %7.1f GB/s algorithmic = %.2f of the HBM peak' % (
                n_mo, n_terms, name, eng.last_kernel(), ms, by / ms / 1e6, by / ms / 1e6 / hbm), flush=True)
        a, b = res[(name, 0)], res[(name, OKB_FLAG_CI_FAST)]
        print('    max |fast - exact| = %.3e   max |exact| = %.3e' % (float((a - b).abs().max()), float(a.abs().max())), flush=True)
    del buf
