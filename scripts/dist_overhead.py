#!/usr/bin/env python
"""Where the time of a sharded rho_compute call goes when a rank has ~1e6 points (the 8-GPU strong-scaling case of the
200^3 benchmark, reproduced on 2 ranks with a 100 x 100 x 200 grid): device time of the shard, the evaluation into the
node-shared host array, the collectives of the assembly, the whole public call.
run: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dist_overhead.py"""
import os, sys, time
import numpy, torch
import torch.distributed as tdist
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
rank = int(os.environ.get('RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1))
local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
if world > 1:
    tdist.init_process_group('nccl', device_id=torch.device('cuda', local))
import orbkit_b200 as ok
from orbkit_b200 import synth, dist as okdist
from orbkit_b200._lib import OKB_FLAG_OUT_DEVICE
from orbkit_b200.engine import get_engine
ok.options.quiet = True
eng = get_engine(); dev = torch.device('cuda', eng.device)
qc = synth.to_qcinfo(synth.make_molecule(n_heavy=24, n_light=20, n_mo=82, seed=0, spherical=True))
nx = 50 * world
ax, ay, az = numpy.linspace(-12, 12, nx), numpy.linspace(-12, 12, 100), numpy.linspace(-12, 12, 200)
npts = nx * 100 * 200
ok.grid.set_grid(ax, ay, az, is_vector=False)
basis = eng.basis(qc.geo_spec, qc.ao_spec); mo = eng.mos_of(basis, qc.mo_spec); g = eng.grid_regular(ax, ay, az)
p0, p1 = okdist.shard_range(npts, rank, world)
buf = torch.zeros((4, p1 - p0), dtype=torch.float64, device=dev)
def timed(f, n=10, warm=3):
    for _ in range(warm): f()
    torch.cuda.synchronize()
    if world > 1: tdist.barrier()
    t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
t_dev = timed(lambda: (eng.eval_rho(mo, g, [1, 2, 3], p0, p1, rho=buf[0].data_ptr(), delta=buf[1:].data_ptr(), flags=OKB_FLAG_OUT_DEVICE), eng.sync()))
host = eng.host_array((4, p1 - p0))
t_host = timed(lambda: eng.eval_rho(mo, g, [1, 2, 3], p0, p1, rho=host[0], delta=host[1:]))
t_coll = timed(lambda: (okdist.all_reduce_min([1, 1, 1]), tdist.barrier())) if world > 1 else 0.0
keep = []
def call():
    keep.append(ok.rho_compute(qc, drv=['x', 'y', 'z']))
    del keep[:-1]
t_call = timed(call)
print('rank %d: %d points of %d | device %.2f ms | eval to pinned host %.2f ms | all_reduce_min + barrier %.2f ms | rho_compute %.2f ms'
      % (rank, p1 - p0, npts, t_dev, t_host, t_coll, t_call), flush=True)
if world > 1:
    tdist.barrier(); tdist.destroy_process_group()
