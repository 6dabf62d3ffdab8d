#!/bin/bash
# rho only (SET_VAL) on the benchmark molecule under forced variants
for v in ${VARIANTS:-"" MB11xBN4xWM1xWN4xNPW8 MB11xBN2xWM1xWN4xNPW8}; do
  echo "== OKB_VARIANT=[$v]"
  OKB_VARIANT="$v" python - <<'PY'
import os, sys, numpy, torch
sys.path.insert(0, '.')
from orbkit_b200 import synth
from orbkit_b200._lib import OKB_FLAG_OUT_DEVICE
from orbkit_b200.engine import get_engine
eng = get_engine(); dev = torch.device('cuda', eng.device)
stream = torch.cuda.ExternalStream(eng.stream_ptr(), device=dev)
qc = synth.to_qcinfo(synth.make_molecule(n_heavy=24, n_light=20, n_mo=82, seed=0, spherical=True))
ax = numpy.linspace(-12, 12, 200)
basis = eng.basis(qc.geo_spec, qc.ao_spec); mo = eng.mos_of(basis, qc.mo_spec); g = eng.grid_regular(ax, ax, ax)
out = torch.zeros((8, 8000000), dtype=torch.float64, device=dev)
f = lambda: eng.eval_rho(mo, g, [], rho=out[0].data_ptr(), flags=OKB_FLAG_OUT_DEVICE)
f(); f(); eng.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(stream):
    e0.record(stream); [f() for _ in range(5)]; e1.record(stream)
eng.sync()
print('rho %.2f ms  %s  sum %.9f' % (e0.elapsed_time(e1) / 5, eng.last_kernel(), float(out[0].sum())))
PY
done
