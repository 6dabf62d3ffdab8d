#!/bin/bash
# rho + laplacian (two passes) under forced tile variants
for v in ${VARIANTS:-"" MB11x MB10x MB10R2}; do
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
for codes in ([1, 2, 3], [4, 5, 6]):
    f = lambda: eng.eval_rho(mo, g, codes, rho=out[0].data_ptr(), delta=out[1:].data_ptr(), flags=OKB_FLAG_OUT_DEVICE)
    f(); f(); eng.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream); [f() for _ in range(3)]; e1.record(stream)
    eng.sync()
    print('codes %-10s %.2f ms  %s' % (codes, e0.elapsed_time(e1) / 3, eng.last_kernel()))
PY
done
