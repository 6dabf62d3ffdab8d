#!/bin/bash
# A/B of a rho request under different environments: args = "codes" then one env string per run
# (e.g. "OKB_VARIANT=WM2xWN4" or "OKB_PLAIN_CHUNKS=1"; use "X=0" for the default)
codes="$1"; shift
for v in "$@"; do
  echo "== [$v] codes=[$codes]"
  env $v python - <<PY
import os, sys, numpy, torch
sys.path.insert(0, '.')
from orbkit_b200 import synth
from orbkit_b200._lib import OKB_FLAG_OUT_DEVICE
from orbkit_b200.engine import get_engine
eng = get_engine(); dev = torch.device('cuda', eng.device)
stream = torch.cuda.ExternalStream(eng.stream_ptr(), device=dev)
n_mo = int(os.environ.get('AB_NMO', '82')); N = int(os.environ.get('AB_N', '200'))
qc = synth.to_qcinfo(synth.make_molecule(n_heavy=24, n_light=20, n_mo=n_mo, seed=0, spherical=True))
ax = numpy.linspace(-12, 12, N)
basis = eng.basis(qc.geo_spec, qc.ao_spec); mo = eng.mos_of(basis, qc.mo_spec); g = eng.grid_regular(ax, ax, ax)
codes = [int(c) for c in "$codes".split()] if "$codes".strip() else []
out = torch.zeros((8, N ** 3), dtype=torch.float64, device=dev)
f = lambda: eng.eval_rho(mo, g, codes, rho=out[0].data_ptr(), delta=out[1:].data_ptr() if codes else None, flags=OKB_FLAG_OUT_DEVICE)
f(); f(); eng.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(stream):
    e0.record(stream); [f() for _ in range(3)]; e1.record(stream)
eng.sync()
ms = e0.elapsed_time(e1) / 3
import ctypes; nch = ctypes.c_int(0); eng.lib.okb_basis_info(basis[0].ptr, None, None, None, ctypes.byref(nch))
D = 1 if not codes else (4 if max(codes) <= 3 else 7 if max(codes) <= 6 else 10)
print('%.2f ms  %.2f TFLOP/s alg  %s  chunks %d  sum %.9f' % (ms, 2.0 * n_mo * 1000 * D * N ** 3 / ms / 1e9, eng.last_kernel(), nch.value, float(out[0].sum())))
PY
done
