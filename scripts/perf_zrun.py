#!/usr/bin/env python
"""A/B timing of the z-run calc_ao kernel variants (OKB_ZRUN=<J><PF>, OKB_ZRUN_RG=<row groups per CTA>) on the benchmark
molecule, device resident: one subprocess per variant (the variant is read once per process)."""
import os, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = r'''
import os, sys, numpy, torch
sys.path.insert(0, %r)
import orbkit_b200 as ok
from orbkit_b200 import synth
from orbkit_b200._lib import OKB_FLAG_OUT_DEVICE
from orbkit_b200.engine import get_engine
ok.options.quiet = True
eng = get_engine(); dev = torch.device('cuda', eng.device)
stream = torch.cuda.ExternalStream(eng.stream_ptr(), device=dev)
qc = synth.to_qcinfo(synth.make_molecule(n_heavy=24, n_light=20, n_mo=8, seed=0, spherical=True))
ax = numpy.linspace(-12, 12, 200)
basis = eng.basis(qc.geo_spec, qc.ao_spec); g = eng.grid_regular(ax, ax, ax)
for nsub in (1000000, 240000):
    buf = torch.empty((1, 1000, nsub), dtype=torch.float64, device=dev)
    fn = lambda: eng.eval_ao(basis, g, [0], 4000000, 4000000 + nsub, out=buf.data_ptr(), flags=OKB_FLAG_OUT_DEVICE)
    fn(); fn(); eng.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(10): fn()
        e1.record(stream)
    eng.sync()
    ms = e0.elapsed_time(e1) / 10
    print('%%-34s pts %%-8d %%7.3f ms  %%6.0f GB/s stored' %% (eng.last_kernel() + ' RG=' + os.environ.get('OKB_ZRUN_RG', 'auto'), nsub, ms, 8e3 * nsub / (ms * 1e-3) / 1e9), flush=True)
    del buf
''' % REPO
for var, rg in (('44', ''), ('40', ''), ('48', ''), ('84', ''), ('24', ''), ('44', '2'), ('44', '4'), ('40', '4'), ('84', '2')):
    env = dict(os.environ, OKB_ZRUN=var)
    if rg:
        env['OKB_ZRUN_RG'] = rg
    subprocess.run([sys.executable, '-c', WORKER], env=env)
