"""GPU: the multi-rank driver (one process per rank, point shards written side by side into one
node-shared host array) returns, on EVERY rank, exactly what a single process returns.
Two ranks share cuda:0 here (process group on gloo, so no NCCL duplicate-GPU restriction); on a
multi-GPU box the same code runs with one GPU per rank on nccl (bench.py --gpus N)."""
import os
import socket
import subprocess
import sys

import numpy
import pytest

from conftest import REPO, golden_qc

pytestmark = pytest.mark.gpu

_WORKER = r'''
import os, sys, numpy, torch
import torch.distributed as dist
sys.path.insert(0, %(repo)r)
sys.path.insert(0, os.path.join(%(repo)r, 'tests'))
rank = int(sys.argv[1])
os.environ['LOCAL_RANK'] = '0'
dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%(port)d', rank=rank, world_size=2)
import orbkit_b200 as ok
from conftest import golden_qc
ok.options.quiet = True
qc, a = golden_qc('synth_small_sph')
ax, ay, az = numpy.linspace(-5, 5, 23), numpy.linspace(-4, 4, 17), numpy.linspace(-3, 3, 11)
ok.grid.set_grid(ax, ay, az, is_vector=False)
out = {}
out['rho'], out['drho'] = ok.rho_compute(qc, drv=['x', 'y', 'z'])
r, d2, lap = ok.rho_compute(qc, laplacian=True)
out['d2'], out['lap'] = d2, lap
out['mo'] = ok.rho_compute(qc, calc_mo=True, drv=[None, 'z'])
out['ao'] = ok.rho_compute(qc, calc_ao=True)
ok.grid.set_grid(a['vx'], a['vy'], a['vz'], is_vector=True)      # fewer points than one shard alignment
out['vec_rho'] = ok.rho_compute(qc)
out['again'] = ok.rho_compute(qc)                                # second segment generation
# ADVICE r01: results of consecutive same-shape calls that are all KEPT must not alias (extras.mo_set keeps the
# densities of every MO set in a list and converts after the loop)
sets = []
for lo in range(0, 8, 2):
    q = qc.copy()
    q.mo_spec = qc.mo_spec[lo:lo + 2]
    sets.append(ok.rho_compute(q))
for i, r in enumerate(sets):
    out['set%%d' %% i] = r
numpy.savez(%(out)r + '_%%d.npz' %% rank, **{k: numpy.array(v) for k, v in out.items()})
dist.barrier()
dist.destroy_process_group()
print('rank', rank, 'ok')
'''


def test_two_ranks_equal_one_process(tmp_path):
    import orbkit_b200 as ok
    ok.options.quiet = True
    qc, a = golden_qc('synth_small_sph')
    ax, ay, az = numpy.linspace(-5, 5, 23), numpy.linspace(-4, 4, 17), numpy.linspace(-3, 3, 11)
    ok.grid.set_grid(ax, ay, az, is_vector=False)
    ref = {}
    ref['rho'], ref['drho'] = ok.rho_compute(qc, drv=['x', 'y', 'z'])
    _, ref['d2'], ref['lap'] = ok.rho_compute(qc, laplacian=True)
    ref['mo'] = ok.rho_compute(qc, calc_mo=True, drv=[None, 'z'])
    ref['ao'] = ok.rho_compute(qc, calc_ao=True)
    ok.grid.set_grid(a['vx'], a['vy'], a['vz'], is_vector=True)
    ref['vec_rho'] = ok.rho_compute(qc)
    ref['again'] = ref['vec_rho']
    for i, lo in enumerate(range(0, 8, 2)):
        q = qc.copy()
        q.mo_spec = qc.mo_spec[lo:lo + 2]
        ref['set%d' % i] = ok.rho_compute(q)
    assert not numpy.array_equal(ref['set0'], ref['set2'])
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / 'dist_worker.py'
    script.write_text(_WORKER % {'repo': REPO, 'port': port, 'out': str(tmp_path / 'res')})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=600)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    for r in range(2):
        got = numpy.load(str(tmp_path / 'res') + '_%d.npz' % r)
        for k, v in ref.items():
            assert got[k].shape == numpy.asarray(v).shape, k
            assert numpy.array_equal(got[k], v), 'rank %d: %s differs from the single-process result' % (r, k)


_NCCL_WORKER = r'''
import os, sys, json, numpy, torch
import torch.distributed as dist
sys.path.insert(0, %(repo)r)
sys.path.insert(0, os.path.join(%(repo)r, 'tests'))
rank = int(sys.argv[1])
os.environ['LOCAL_RANK'] = str(rank)
torch.cuda.set_device(rank)
dist.init_process_group('nccl', init_method='tcp://127.0.0.1:%(port)d', rank=rank, world_size=2,
                        device_id=torch.device('cuda', rank))
import orbkit_b200 as ok
from orbkit_b200 import dist as okdist
from orbkit_b200._lib import OKB_FLAG_OUT_DEVICE
from orbkit_b200.engine import get_engine
from conftest import golden_qc
ok.options.quiet = True
qc, a = golden_qc('synth_small_sph')
ax, ay, az = numpy.linspace(-5, 5, 61), numpy.linspace(-4, 4, 47), numpy.linspace(-3, 3, 33)
ok.grid.set_grid(ax, ay, az, is_vector=False)
out = {}
# (1) the public API: shards streamed into the node-shared host array, N_el / norms all-reduced on NCCL
out['rho'], out['drho'] = ok.rho_compute(qc, drv=['x', 'y', 'z'])
# (2) the device-side assembly: this rank's shard evaluated into a device tensor, all-gathered over NVLink
eng = get_engine()
dev = torch.device('cuda', eng.device)
npts = len(ax) * len(ay) * len(az)
p0, p1 = okdist.shard_range(npts, rank, 2)
basis = eng.basis(qc.geo_spec, qc.ao_spec)
mo = eng.mos_of(basis, qc.mo_spec)
g = eng.grid_regular(ax, ay, az)
shard = torch.zeros((4, p1 - p0), dtype=torch.float64, device=dev)
eng.eval_rho(mo, g, [1, 2, 3], p0, p1, rho=shard[0].data_ptr(), delta=shard[1:].data_ptr(), flags=OKB_FLAG_OUT_DEVICE)
eng.sync()
full = okdist.gather_points(shard, npts)
out['gathered'] = full.cpu().numpy()
# bandwidth of the gather at a size where NVLink matters (32 B per point, 8e6 points per rank)
big = torch.zeros((4, 8000000), dtype=torch.float64, device=dev)
okdist.gather_points(big, 16000000)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    okdist.gather_points(big, 16000000)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
out['gather_ms'] = ms
out['gather_gbs_received'] = 32.0 * 8000000 / (ms * 1e-3) / 1e9
numpy.savez(%(out)r + '_%%d.npz' %% rank, **{k: numpy.array(v) for k, v in out.items()})
dist.barrier()
dist.destroy_process_group()
print('rank', rank, 'ok', 'gather %%.2f ms = %%.0f GB/s received' %% (ms, out['gather_gbs_received']))
'''


def test_nccl_two_gpus_shared_array_and_device_gather(tmp_path):
    """one GPU per rank on NCCL (skipped on a single-GPU box): the shared-host-array assembly of rho_compute and the
    device-side dist.gather_points over NVLink both reproduce the single-process result bit for bit; the gather's
    bandwidth is printed (profiles/r02_nccl_gather.txt keeps a measured value)"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    import orbkit_b200 as ok
    ok.options.quiet = True
    qc, a = golden_qc('synth_small_sph')
    ax, ay, az = numpy.linspace(-5, 5, 61), numpy.linspace(-4, 4, 47), numpy.linspace(-3, 3, 33)
    ok.grid.set_grid(ax, ay, az, is_vector=False)
    rho, drho = ok.rho_compute(qc, drv=['x', 'y', 'z'])
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / 'nccl_worker.py'
    script.write_text(_NCCL_WORKER % {'repo': REPO, 'port': port, 'out': str(tmp_path / 'res')})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=900)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    print(outs[0].strip().splitlines()[-1])
    for r in range(2):
        got = numpy.load(str(tmp_path / 'res') + '_%d.npz' % r)
        assert numpy.array_equal(got['rho'], rho) and numpy.array_equal(got['drho'], drho)
        assert numpy.array_equal(got['gathered'][0], rho.reshape(-1))
        assert numpy.array_equal(got['gathered'][1:], drho.reshape(3, -1))
        assert got['gather_gbs_received'] > 50.0              # far above PCIe: the shards travelled over NVLink
