#!/usr/bin/env python
"""Device-resident timing of the main request types on the benchmark molecule (run on a B200).
Prints one line per workload: ms, points/s, algorithmic TFLOP/s or GB/s, kernel variant."""
import json, os, sys, time
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

def timed(fn, reps=3):
    fn(); eng.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(reps): fn()
        e1.record(stream)
    eng.sync()
    return e0.elapsed_time(e1) / reps

def run(name, n_mo, spherical=True, N=200, heavy=24, light=20):
    spec = synth.make_molecule(n_heavy=heavy, n_light=light, n_mo=n_mo, seed=0, spherical=spherical)
    qc = synth.to_qcinfo(spec)
    n_ao = qc.ao_spec.get_ao_num()
    ax = numpy.linspace(-12, 12, N)
    basis = eng.basis(qc.geo_spec, qc.ao_spec)
    mo = eng.mos_of(basis, qc.mo_spec)
    g = eng.grid_regular(ax, ax, ax)
    npts = N ** 3
    out = torch.zeros((8, npts), dtype=torch.float64, device=dev)
    rows = []
    for label, codes, D in (('rho', [], 1), ('rho+grad', [1, 2, 3], 4), ('rho+lap', [4, 5, 6], 7)):
        ms = timed(lambda: eng.eval_rho(mo, g, codes, rho=out[0].data_ptr(), delta=out[1:].data_ptr() if codes else None, flags=OKB_FLAG_OUT_DEVICE))
        tf = 2.0 * n_mo * n_ao * D * npts / (ms * 1e-3) / 1e12
        rows.append((name, label, npts, ms, npts / ms * 1e3, '%.2f TFLOP/s alg' % tf, eng.last_kernel()))
    # MOs (value) for a point range that fits: n_mo rows
    nsub = min(npts, int(6e9 // (8 * max(n_mo, 1))))
    mobuf = torch.empty((1, n_mo, nsub), dtype=torch.float64, device=dev)
    ms = timed(lambda: eng.eval_mo(mo, g, [0], 0, nsub, out=mobuf.data_ptr(), flags=OKB_FLAG_OUT_DEVICE))
    rows.append((name, 'calc_mo', nsub, ms, nsub / ms * 1e3, '%.2f TFLOP/s alg, %.0f GB/s out' % (2.0 * n_mo * n_ao * nsub / (ms * 1e-3) / 1e12, 8.0 * n_mo * nsub / (ms * 1e-3) / 1e9), eng.last_kernel()))
    del mobuf
    nsub = min(npts, int(8e9 // (8 * n_ao)))
    aobuf = torch.empty((1, n_ao, nsub), dtype=torch.float64, device=dev)
    for label, codes in (('calc_ao', [0]),):
        ms = timed(lambda: eng.eval_ao(basis, g, codes, 0, nsub, out=aobuf.data_ptr(), flags=OKB_FLAG_OUT_DEVICE))
        rows.append((name, label, nsub, ms, nsub / ms * 1e3, '%.0f GB/s out' % (8.0 * n_ao * nsub / (ms * 1e-3) / 1e9), eng.last_kernel()))
    del aobuf
    for r in rows:
        print('%-14s %-9s pts %-9d %9.2f ms  %.3e pts/s  %-32s %s' % r)
    return rows

def run_ci(name, n_heavy, n_light, n_pairs, N):
    """Config 5: detCI-style batch -- n_pairs random MO pairs of an all-MO calculation.
    (a) the pair contraction alone on device-resident MOs (HBM bound: 8*n_mo B/pt in, 8 B/pt out),
    (b) the fused path okb_eval_ci (MOs evaluated slab by slab on the device, then contracted)."""
    from orbkit_b200 import _lib
    spec = synth.make_molecule(n_heavy=n_heavy, n_light=n_light, n_mo=n_heavy * 30 + n_light * 14, seed=5, spherical=True)
    qc = synth.to_qcinfo(spec)
    n_mo = len(qc.mo_spec)
    n_ao = qc.ao_spec.get_ao_num()
    rng = numpy.random.default_rng(5)
    pairs = rng.integers(0, n_mo, size=(n_pairs, 2))
    terms = (rng.normal(size=n_pairs), pairs[:, 0].astype(numpy.intc), pairs[:, 1].astype(numpy.intc))
    ax = numpy.linspace(-10, 10, N)
    basis = eng.basis(qc.geo_spec, qc.ao_spec)
    mo = eng.mos_of(basis, qc.mo_spec)
    g = eng.grid_regular(ax, ax, ax)
    npts = N ** 3
    rows = []
    out = torch.zeros((3, npts), dtype=torch.float64, device=dev)
    # (a) MOs (+ gradient) resident in HBM
    nsub = min(npts, int(24e9 // (8 * 4 * n_mo)))
    mobuf = torch.empty((4, n_mo, nsub), dtype=torch.float64, device=dev)
    eng.eval_mo(mo, g, [0, 1, 2, 3], 0, nsub, out=mobuf.data_ptr(), flags=OKB_FLAG_OUT_DEVICE)
    fl = OKB_FLAG_OUT_DEVICE | _lib.OKB_FLAG_IN_DEVICE
    for label, mode, nset, ncomp in (('ci rho', _lib.OKB_CI_RHO, 1, 1), ('ci jab', _lib.OKB_CI_JAB, 4, 3)):
        ms = timed(lambda: eng.ci_contract(mode, terms, mobuf[0].data_ptr(), mobuf[1:].data_ptr(), n_mo=n_mo, npts=nsub,
                                           ld=nsub, out=out.data_ptr(), flags=fl))
        gbs = 8.0 * (n_mo * nset + ncomp) * nsub / (ms * 1e-3) / 1e9
        rows.append((name, label, nsub, ms, nsub / ms * 1e3, '%.0f GB/s alg (%d pairs)' % (gbs, n_pairs), eng.last_kernel()))
    del mobuf
    # (b) fused
    for label, mode, D in (('rho_from_qc', _lib.OKB_CI_RHO, 1), ('jab_from_qc', _lib.OKB_CI_JAB, 4)):
        ms = timed(lambda: eng.eval_ci(mode, terms, mo, g, out=out.data_ptr(), flags=OKB_FLAG_OUT_DEVICE), reps=2)
        tf = 2.0 * n_mo * n_ao * D * npts / (ms * 1e-3) / 1e12
        rows.append((name, label, npts, ms, npts / ms * 1e3, '%.2f TFLOP/s alg (MO part)' % tf, eng.last_kernel()))
    for r in rows:
        print('%-14s %-11s pts %-9d %9.2f ms  %.3e pts/s  %-32s %s' % r)
    return rows


if __name__ == '__main__':
    allrows = []
    allrows += run('c3 n_mo=82', 82)
    allrows += run('c3 n_mo=1000', 1000, N=100)
    allrows += run('c3 cart n_mo=82', 82, spherical=False, N=128)
    if os.environ.get('PERF_BIG', '1') == '1':
        allrows += run_ci('c5 500 MOs', 12, 10, 1000, 128)
        # Config 2 shape (222 spherical AOs: 12 heavy x 14... here 6 C-like + 3 H-like = 222 AOs), 21 occupied MOs / all MOs, 150^3
        allrows += run('c2 n_mo=21', 21, N=150, heavy=6, light=3)
        allrows += run('c2 n_mo=222', 222, N=150, heavy=6, light=3)
        # Config 4 shape: ~3000 AOs, 246 MOs, 256^3 (one GPU's share of the 8-GPU run is 1/8 of this)
        allrows += run('c4 n_mo=246', 246, N=160, heavy=72, light=60)
    # end-to-end pieces of rho_compute on the bench workload
    spec = synth.make_molecule(n_heavy=24, n_light=20, n_mo=82, seed=0, spherical=True)
    qc = synth.to_qcinfo(spec)
    ax = numpy.linspace(-12, 12, 200)
    ok.grid.set_grid(ax, ax, ax, is_vector=False)
    for i in range(5):
        eng.clear_caches()
        t0 = time.perf_counter(); basis = eng.basis(qc.geo_spec, qc.ao_spec); t1 = time.perf_counter()
        mo = eng.mos_of(basis, qc.mo_spec); t2 = time.perf_counter()
        g = eng.grid_regular(ax, ax, ax); t3 = time.perf_counter()
        rho = eng.host_array((8000000,)); d = eng.host_array((3, 8000000)); t4 = time.perf_counter()
        eng.eval_rho(mo, g, [1, 2, 3], rho=rho, delta=d); t5 = time.perf_counter()
        r = ok.rho_compute(qc, drv=['x', 'y', 'z']); t6 = time.perf_counter()
        print('e2e pieces [ms]: basis %.1f mos %.1f grid %.1f host_alloc %.1f eval(host out) %.1f | rho_compute (cached handles) %.1f'
              % tuple(1e3 * v for v in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5)))
    json.dump([list(r) for r in allrows], open(os.path.join(REPO, 'gpurun_out', 'perf_matrix.json'), 'w'))
