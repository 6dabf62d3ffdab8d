#!/usr/bin/env python
"""BASELINE configs[3]: synthetic ~3000-basis-function system (72 heavy + 60 light atoms, 246 doubly occupied MOs), rho +
grad rho on a 256^3 grid, sharded over the ranks of one box (torchrun) through the public API: NumPy arrays back on
every rank.  Checks: sub-sample parity against the CPU oracle (reference objects), electron count against the sum of
the shards.  One JSON line from rank 0.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/run_c4.py"""
import json, os, sys, time
import numpy, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'oracle'))
import orbkit_b200 as ok
from orbkit_b200 import synth

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
ok.options.quiet = True
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
qc = synth.to_qcinfo(synth.make_molecule(n_heavy=72, n_light=60, n_mo=246, seed=0, spherical=True, box=12.0))
ax = numpy.linspace(-16.0, 16.0, N)
ok.grid.set_grid(ax, ax, ax, is_vector=False)
times = []
for it in range(3):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rho, drho = ok.rho_compute(qc, drv=['x', 'y', 'z'])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    times.append(time.perf_counter() - t0)
n_ao = qc.ao_spec.get_ao_num()
d3r = (ax[1] - ax[0]) ** 3
if rank == 0:
    import oracle
    rng = numpy.random.default_rng(1)
    idx = numpy.sort(rng.choice(N ** 3, size=192, replace=False))
    ix, iy, iz = numpy.unravel_index(idx, (N, N, N))
    kind = 'ref' if oracle.have_ref() else 'port'
    r_ref, d_ref = oracle.rho_compute(qc, ax[ix], ax[iy], ax[iz], is_vector=True, drv=['x', 'y', 'z'], kind=kind)
    got_r, got_d = rho.reshape(-1)[idx], drho.reshape(3, -1)[:, idx]
    tol = lambda ref: 1e-10 * numpy.abs(ref) + 1e-14 * numpy.abs(ref).max()
    ok_r = bool((numpy.abs(got_r - r_ref) <= tol(r_ref)).all())
    ok_d = bool((numpy.abs(got_d - d_ref) <= tol(d_ref)).all())
    t = min(times[1:])
    print(json.dumps({'config': 'C4: %d AOs, 246 MOs, rho + grad rho, %d^3 points' % (n_ao, N), 'n_gpus': world,
                      'e2e_s': [round(v, 4) for v in times], 'points_per_s': N ** 3 / t,
                      'tflops_alg': 2.0 * 246 * n_ao * 4 * N ** 3 / t / 1e12,
                      'parity_192_points_vs_oracle': {'rho': ok_r, 'grad': ok_d, 'kind': kind},
                      'electrons': float(rho.sum() * d3r), 'shape': list(rho.shape)}), flush=True)
if world > 1:
    dist.destroy_process_group()
