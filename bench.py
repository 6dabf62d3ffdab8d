#!/usr/bin/env python
"""bench.py -- grid points/s of the ORBKIT grid path (rho + grad rho, FP64) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json north_star target, configs[2] molecule): synthetic C24H20-like molecule
with cc-pVTZ-shaped shells (360 contractions, 784 primitives, 1140 Cartesian -> 1000 spherical AOs,
82 doubly occupied MOs, numpy default_rng(0)); rho and (d/dx,d/dy,d/dz) rho on a regular grid of
200 x 200 x 200 points over [-12,12]^3 bohr PER GPU.  A step is one pass of the hot path over that
grid.  N > 1: weak scaling -- the x axis carries 200*N points over the same box and every rank
owns a contiguous x-slab of 8e6 points; there is no data-path collective.

One JSON line is printed by rank 0:
  value         whole-job points/s, outputs left in HBM, CUDA-event timed on the launch stream
  e2e           the same through orbkit_b200.rho_compute (QCinfo in, NumPy out): per step the
                basis tables, MO coefficients and grid axes are re-uploaded (handle caches dropped)
                and the result comes back to host memory inside the timed region
  roofline      the fused kernel against the FP64 peak MEASURED in this run: the larger of the sustained
                DFMA and DMMA microbenchmarks (MEASURED_PEAKS.json has no FP64 entry); algorithmic
                flops = 2*n_mo*n_ao*4 per point
  cpu_baseline  the reference's CPU path (oracle/_ref objects) on the box's host cores, bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

GRID_N = 200
BOX = 12.0
DRV = ['x', 'y', 'z']
N_AO, N_MO, D_SETS = 1000, 82, 4
ALG_FLOPS_PER_POINT = 2.0 * N_MO * N_AO * D_SETS          # 656 kFLOP
ALG_BYTES_PER_POINT = 8.0 * (1 + 3)                        # rho + 3 derivatives out, 0 in
WORKLOAD = ('synthetic 1000-AO/82-MO molecule (BASELINE configs[2] generator, seed 0), rho+grad rho, '
            '200^3 regular grid on [-12,12]^3 per GPU')


def molecule():
    from orbkit_b200 import synth
    return synth.make_molecule(n_heavy=24, n_light=20, n_mo=N_MO, seed=0, spherical=True)


def axes(n_gpus):
    ax = numpy.linspace(-BOX, BOX, GRID_N)
    return numpy.linspace(-BOX, BOX, GRID_N * n_gpus), ax, ax


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region"""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, mx, pw = [], set(), None, []
        try:
            for line in open(self.path):
                f = [t.strip() for t in line.split(',')]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx = float(f[2]); pw.append(float(f[3]))
                except ValueError:
                    continue
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(numpy.median(sm)), sm_max_mhz=mx, reasons=sorted(reasons),
                       samples=len(sm), power_w_max=max(pw) if pw else None)
        return out


def dist_setup(n_gpus):
    import torch
    rank, world, local = 0, 1, 0
    if 'RANK' in os.environ and int(os.environ.get('WORLD_SIZE', '1')) > 1:
        import torch.distributed as dist
        rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    return rank, world, local


def barrier(world):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x, world, device):
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the B200 arm has no CPU fallback')
    rank, world, local = dist_setup(args.gpus)
    if world != args.gpus and world > 1:
        raise SystemExit('--gpus %d but WORLD_SIZE=%d' % (args.gpus, world))
    import orbkit_b200 as ok
    from orbkit_b200 import synth, dist as okdist
    from orbkit_b200._lib import OKB_FLAG_OUT_DEVICE
    from orbkit_b200.engine import get_engine
    ok.options.quiet = True
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    spec = molecule()
    qc = synth.to_qcinfo(spec)
    gx, gy, gz = axes(world)
    npts_total = len(gx) * len(gy) * len(gz)
    eng = get_engine()
    stream = torch.cuda.ExternalStream(eng.stream_ptr(), device=dev)

    # ---- resident inputs: tables, coefficients, axes on the device; outputs stay in HBM -------------
    basis = eng.basis(qc.geo_spec, qc.ao_spec)
    mo = eng.mos_of(basis, qc.mo_spec)
    g = eng.grid_regular(gx, gy, gz)
    p0, p1 = okdist.shard_range(npts_total, rank, world)
    n_loc = p1 - p0
    out = torch.zeros((4, n_loc), dtype=torch.float64, device=dev)
    codes = [1, 2, 3]

    def step():
        eng.eval_rho(mo, g, codes, p0, p1, rho=out[0].data_ptr(), delta=out[1:].data_ptr(), flags=OKB_FLAG_OUT_DEVICE)

    # FP64 roofline denominator, measured here (burst + sustained), rank 0 only prints it
    if args.no_peaks:       # profiler runs: keep the launch list short (numbers of profiles/r01_fp64_micro.txt)
        dfma_burst, dmma_burst, dfma_sust, dmma_sust = 34.0, 37.0, 34.0, 36.9
    else:
        dfma_burst, _ = eng.measure_fp64(0, 0.0)
        dmma_burst, _ = eng.measure_fp64(1, 0.0)
        dfma_sust, _ = eng.measure_fp64(0, 1.0)
        dmma_sust, _ = eng.measure_fp64(1, 1.0)
    fp64_peak = max(dfma_sust, dmma_sust)

    for _ in range(args.warmup):
        step()
    eng.sync()
    sampler = ClockSampler(local)
    barrier(world)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with torch.cuda.stream(stream):
        evs[0].record(stream)
        for i in range(args.steps):
            step()
            evs[i + 1].record(stream)
    eng.sync()
    barrier(world)
    launches = eng.launch_count() - launches0
    total_ms = evs[0].elapsed_time(evs[-1])
    per_step = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    total_ms = max_over_ranks(total_ms, world, dev)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = npts_total / (ms_per_step * 1e-3)
    kernel_name = eng.last_kernel()
    electrons = float(out[0].sum().item()) * (gx[1] - gx[0]) * (gy[1] - gy[0]) * (gz[1] - gz[0])
    if world > 1:
        electrons = float(okdist.all_reduce_sum([electrons], local)[0])

    # ---- the other request types of the path on the same molecule and shard, device resident (reported, not the
    # headline): rho only, rho + laplacian (BASELINE configs[2]), all-AO store (the HBM-bound request) -----------------
    also = None
    if not args.no_also:
        def timed(fn, reps=2):
            fn()
            eng.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record(stream)
                for _ in range(reps):
                    fn()
                e1.record(stream)
            eng.sync()
            return e0.elapsed_time(e1) / reps
        ms_rho = timed(lambda: eng.eval_rho(mo, g, [], p0, p1, rho=out[0].data_ptr(), flags=OKB_FLAG_OUT_DEVICE))
        ms_lap = timed(lambda: eng.eval_rho(mo, g, [4, 5, 6], p0, p1, rho=out[0].data_ptr(), delta=out[1:].data_ptr(),
                                            flags=OKB_FLAG_OUT_DEVICE))
        n_sub = min(n_loc, 1000000) // 1024 * 1024
        aobuf = torch.empty((1, N_AO, n_sub), dtype=torch.float64, device=dev)
        ms_ao = timed(lambda: eng.eval_ao(basis, g, [0], p0, p0 + n_sub, out=aobuf.data_ptr(), flags=OKB_FLAG_OUT_DEVICE))
        ao_kernel = eng.last_kernel()
        ms_ao_dx = timed(lambda: eng.eval_ao(basis, g, [1], p0, p0 + n_sub, out=aobuf.data_ptr(), flags=OKB_FLAG_OUT_DEVICE))
        ao_dx_kernel = eng.last_kernel()
        del aobuf
        also = {'rho_ms': round(ms_rho, 3), 'rho_tflops_alg': round(2.0 * N_MO * N_AO * n_loc / ms_rho / 1e9, 2),
                'rho_laplacian_ms': round(ms_lap, 3),
                'rho_laplacian_tflops_alg': round(2.0 * N_MO * N_AO * 7 * n_loc / ms_lap / 1e9, 2),
                'calc_ao_ms_per_1e6_points': round(ms_ao * 1e6 / n_sub, 3),
                'calc_ao_gbs_stored': round(8.0 * N_AO * n_sub / ms_ao / 1e6, 1), 'calc_ao_kernel': ao_kernel,
                'calc_ao_ddx_gbs_stored': round(8.0 * N_AO * n_sub / ms_ao_dx / 1e6, 1), 'calc_ao_ddx_kernel': ao_dx_kernel,
                'note': 'per GPU, device resident, same molecule; algorithmic flops 2*n_mo*n_ao*D per point (D = 1, 7)'}

    # ---- end to end through the public API: QCinfo + grid in, NumPy out, every step ------------------
    ok.grid.set_grid(gx, gy, gz, is_vector=False)
    e2e_steps = args.steps

    def e2e_step():
        eng.clear_caches()                  # tables / coefficients / axes are uploaded again
        return ok.rho_compute(qc, drv=DRV)

    def time_e2e(fn, steps, sync_ranks=True):
        """wall time per step of `fn` (max over ranks when `sync_ranks`), 3 untimed calls first"""
        r = None
        for _ in range(3):
            r = fn()
        if sync_ranks:
            barrier(world)
        t0 = time.perf_counter()
        for _ in range(steps):
            r = fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if sync_ranks:
            barrier(world)
            dt = max_over_ranks(dt, world, dev)
        return dt / steps, r

    if args.no_e2e:
        e2e_steps = 0
    h0, d0 = eng.traffic()
    r = [numpy.zeros((len(gx), len(gy), len(gz)))]
    t_e2e = float('nan')
    if e2e_steps:
        for _ in range(3):                  # page-locked result buffers come from a caching allocator
            r = e2e_step()
        barrier(world)
        h0, d0 = eng.traffic()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            r = e2e_step()
        torch.cuda.synchronize()
        t_e2e = time.perf_counter() - t0
        barrier(world)
    h1, d1 = eng.traffic()
    t_e2e = max_over_ranks(t_e2e, world, dev)
    e2e_value = npts_total * e2e_steps / t_e2e if e2e_steps else None
    d2h = (d1 - d0) / max(e2e_steps, 1)
    h2d = (h1 - h0) / max(e2e_steps, 1)
    if world > 1:                           # whole job: every rank copies its own shard into the shared host array
        d2h, h2d = [float(v) for v in okdist.all_reduce_sum([d2h, h2d], local)]
    e2e_ok = bool(numpy.isfinite(r[0]).all() and r[0].shape == (len(gx), len(gy), len(gz)))
    del r

    # ---- STRONG scaling in the same run (north_star: "sharding grid points across the 8 GPUs"): a FIXED grid over the
    # N ranks, device-timed and through rho_compute, against the same grid on one GPU (rank 0 alone) -----------------
    strong = None
    if not args.no_strong:
        strong = {'efficiency_def': 't_1 / (N * t_N); t_1 measured on rank 0 alone in this run'}
        cases = [('c3_200cube', spec, qc, numpy.linspace(-BOX, BOX, GRID_N), ALG_FLOPS_PER_POINT, args.steps)]
        if not args.no_c4:
            spec4 = synth.make_molecule(n_heavy=72, n_light=60, n_mo=246, seed=0, spherical=True)
            cases.append(('c4_256cube', spec4, synth.to_qcinfo(spec4), numpy.linspace(-BOX, BOX, 256),
                          2.0 * 246 * 3000 * 4, 2))
        for name, sp_, qc_, ax_, flops_pt, ksteps in cases:
            n_all = len(ax_) ** 3
            b_ = eng.basis(qc_.geo_spec, qc_.ao_spec)
            m_ = eng.mos_of(b_, qc_.mo_spec)
            g_ = eng.grid_regular(ax_, ax_, ax_)
            q0, q1 = okdist.shard_range(n_all, rank, world)
            buf = torch.zeros((4, n_all if rank == 0 else max(q1 - q0, 1)), dtype=torch.float64, device=dev)

            def dev_ms(a, b_end, reps):
                ld = buf.shape[1]
                f = lambda: eng.eval_rho(m_, g_, codes, a, b_end, rho=buf[0].data_ptr(), delta=buf[1:].data_ptr(),
                                         flags=OKB_FLAG_OUT_DEVICE, ld=ld)
                f()
                eng.sync()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(stream):
                    e0.record(stream)
                    for _ in range(reps):
                        f()
                    e1.record(stream)
                eng.sync()
                return e0.elapsed_time(e1) / reps
            barrier(world)
            t_n = max_over_ranks(dev_ms(q0, q1, ksteps) if q1 > q0 else 0.0, world, dev)
            t_1 = t_n
            if world > 1:
                t_1 = dev_ms(0, n_all, max(1, min(ksteps, 3))) if rank == 0 else 0.0
                t_1 = max_over_ranks(t_1, world, dev)
            del buf
            ok.grid.set_grid(ax_, ax_, ax_, is_vector=False)
            fn = lambda: ok.rho_compute(qc_, drv=DRV)
            e_n, res = time_e2e(fn, ksteps)
            e_n *= 1e3
            e_ok = bool(res[0].shape == (len(ax_),) * 3 and numpy.isfinite(res[1]).all())
            del res
            e_1 = e_n
            if world > 1:
                e_1 = 0.0
                if rank == 0:
                    with okdist.local_only():
                        e_1, res = time_e2e(fn, max(1, min(ksteps, 5)), sync_ranks=False)
                    e_1 *= 1e3
                    del res
                e_1 = max_over_ranks(e_1, world, dev)
            strong[name] = {'points': n_all, 'steps': ksteps,
                            'device_ms_1gpu': round(t_1, 3), 'device_ms': round(t_n, 3),
                            'device_efficiency': round(t_1 / (world * t_n), 4),
                            'device_points_per_s': n_all / (t_n * 1e-3),
                            'device_tflops_alg': round(flops_pt * n_all / (t_n * 1e-3) / 1e12, 2),
                            'e2e_ms_1gpu': round(e_1, 3), 'e2e_ms': round(e_n, 3),
                            'e2e_efficiency': round(e_1 / (world * e_n), 4),
                            'e2e_points_per_s': n_all / (e_n * 1e-3), 'e2e_ok': e_ok}
        ok.grid.set_grid(gx, gy, gz, is_vector=False)

    # ---- device-side gather of the shards over NVLink (dist.gather_points on NCCL): the alternative assembly for
    # callers that want the full result as a device tensor ---------------------------------------------
    gather = None
    if world > 1:
        shard = torch.zeros((4, n_loc), dtype=torch.float64, device=dev)
        full = okdist.gather_points(shard, npts_total)
        torch.cuda.synchronize()
        barrier(world)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            full = okdist.gather_points(shard, npts_total)
        e1.record()
        torch.cuda.synchronize()
        g_ms = max_over_ranks(e0.elapsed_time(e1) / 5, world, dev)
        recv = 32.0 * (npts_total - n_loc)
        gather = {'api': 'orbkit_b200.dist.gather_points (NCCL all-gather + unpadding copy)', 'ms': round(g_ms, 3),
                  'bytes_received_per_rank': recv, 'gbs_per_rank': round(recv / (g_ms * 1e-3) / 1e9, 1),
                  'nvlink_peak_gbs_per_direction': 900.0}
        del full, shard

    # ---- small-call latency (the reference's cubature example calls rho_compute thousands of times on small vector
    # grids with new points each time, examples/basic_examples/orbkit_interface_to_cubature.py:76-102) ---------------
    latency = None
    if rank == 0 and not args.no_latency:
        with okdist.local_only():
            rng = numpy.random.default_rng(1)
            pts = rng.uniform(-BOX, BOX, size=(64, 3, 1000))

            def call(i):
                ok.grid.x, ok.grid.y, ok.grid.z = pts[i % 64]
                ok.grid.is_initialized, ok.grid.is_vector = True, True
                return ok.rho_compute(qc, drv=None, numproc=1)
            ok.grid.set_grid(pts[0][0], pts[0][1], pts[0][2], is_vector=True)
            for i in range(10):
                call(i)
            t0 = time.perf_counter()
            ncall = 200
            for i in range(ncall):
                out1 = call(i)
            us = (time.perf_counter() - t0) / ncall * 1e6
            latency = {'api': 'orbkit_b200.rho_compute(qc) on a NEW 1000-point vector grid per call, handles cached',
                       'us_per_call': round(us, 1), 'points_per_call': 1000, 'calls': ncall,
                       'ok': bool(out1.shape == (1000,) and numpy.isfinite(out1).all())}
        ok.grid.set_grid(gx, gy, gz, is_vector=False)
    barrier(world)

    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant (only) kernel of the step -----------------------------------------
    kern_ms = float(numpy.mean(per_step))
    achieved = ALG_FLOPS_PER_POINT * n_loc / (kern_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(REPO, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(kernel_name)
        except Exception:
            traffic = None
    roofline = {'bound': 'fp64', 'achieved': round(achieved, 3), 'peak': round(fp64_peak, 3), 'unit': 'TFLOP/s',
                'frac': round(achieved / fp64_peak, 4), 'traffic': traffic,
                'kernel': kernel_name, 'kernel_ms': round(kern_ms, 3),
                'peak_source': 'larger of the FP64 DFMA and DMMA (mma.sync.m8n8k4.f64) issue-bound microbenchmarks '
                               'measured in this run, each sustained for 1 s (MEASURED_PEAKS.json has no FP64 entry; '
                               'tcgen05 has no FP64 kind)',
                'peak_dfma_sustained': round(dfma_sust, 3), 'peak_dmma_m8n8k4_sustained': round(dmma_sust, 3),
                'peak_dfma_burst': round(dfma_burst, 3), 'peak_dmma_m8n8k4_burst': round(dmma_burst, 3),
                'alg_flops_per_point': ALG_FLOPS_PER_POINT,
                'hbm': {'alg_bytes_per_point': ALG_BYTES_PER_POINT,
                        'achieved_gbs': round(ALG_BYTES_PER_POINT * n_loc / (kern_ms * 1e-3) / 1e9, 2),
                        'peak_gbs': _peaks().get('hbm_gbs')}}
    cpu = None if args.no_cpu else cpu_baseline(spec, args)
    line = {'metric': 'grid points/s, rho + grad rho (FP64)', 'value': value, 'unit': 'points/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic (seeded random molecule and MO coefficients)',
            'config': {'workload': WORKLOAD, 'points_per_gpu': n_loc, 'points_total': npts_total,
                       'n_ao': N_AO, 'n_cart': 1140, 'n_mo': N_MO, 'derivative_sets': D_SETS,
                       'parallelism': 'points sharded over %d rank(s), no data-path collective' % world,
                       'l2': 'outputs 256 MB/step/GPU exceed the 126 MB L2; inputs are 1 MB of tables by design'},
            'e2e': {'value': e2e_value, 'unit': 'points/s', 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h, 'steps': e2e_steps, 'api': 'orbkit_b200.rho_compute(qc, drv=["x","y","z"])',
                    'ok': e2e_ok},
            'strong': strong, 'gather': gather, 'latency': latency,
            'gpu_launches': launches, 'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu,
            'electrons': electrons, 'also': also}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _peaks():
    try:
        return json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))
    except Exception:
        return {}


def cpu_sample(n_gpus, npts):
    """contiguous run of grid points from the middle of the benchmark grid (vector form)"""
    gx, gy, gz = axes(1)
    start = (GRID_N // 2) * GRID_N * GRID_N
    idx = numpy.arange(start, start + npts)
    i, rem = numpy.divmod(idx, GRID_N * GRID_N)
    j, k = numpy.divmod(rem, GRID_N)
    return gx[i], gy[j], gz[k]


def cpu_baseline(spec, args, steps=1):
    sys.path.insert(0, os.path.join(REPO, 'oracle'))
    import cpu_bench
    cores = os.cpu_count() or 1
    workers = min(cores, 64)
    npts = workers * 20000 if args.cpu_points <= 0 else args.cpu_points
    x, y, z = cpu_sample(1, npts)
    res = cpu_bench.time_cpu(spec, x, y, z, DRV, nproc=workers, slice_length=2000, repeats=steps)
    # the reference's default slice_length (core.py:314: 1e4) as well: two slices per worker
    res4 = cpu_bench.time_cpu(spec, x, y, z, DRV, nproc=workers, slice_length=10000, repeats=steps)
    # one small call (1000 new points, density only) on one core: what the cubature example pays per call
    import oracle
    from orbkit_b200 import synth
    qc = synth.to_qcinfo(spec)
    rng = numpy.random.default_rng(1)
    px, py, pz = rng.uniform(-BOX, BOX, size=(3, 1000))
    kind = cpu_bench.default_kind()
    oracle.rho_compute(qc, px, py, pz, is_vector=True, kind=kind)
    t0 = time.perf_counter()
    for _ in range(2):
        oracle.rho_compute(qc, px, py, pz, is_vector=True, kind=kind)
    small_us = (time.perf_counter() - t0) / 2 * 1e6
    return {'value': res['points_per_s'], 'unit': 'points/s', 'cores': res['cores'], 'kind': res['kind'],
            'sample': '%d contiguous points from the middle x-plane of the 200^3 grid, %d worker processes, '
                      'slices of 2000 points (reference Pool driver, core.py:503-536); %.1f s'
                      % (res['npts'], res['cores'], res['seconds']), 'host_cores': cores,
            'value_slice_1e4': res4['points_per_s'], 'seconds_slice_1e4': round(res4['seconds'], 2),
            'us_per_1000_point_call_1core': round(small_us, 1)}


def run_reference(args):
    """the reference's own CPU implementation of the path on the host cores (oracle/_ref)"""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    spec = molecule()
    sys.path.insert(0, os.path.join(REPO, 'oracle'))
    import cpu_bench
    cores = os.cpu_count() or 1
    workers = min(cores, 64)
    npts = workers * 20000 if args.cpu_points <= 0 else args.cpu_points
    x, y, z = cpu_sample(1, npts)
    kind = cpu_bench.default_kind()
    res4 = cpu_bench.time_cpu(spec, x, y, z, DRV, nproc=workers, slice_length=10000, kind=kind)   # reported beside
    times = []
    for s in range(args.warmup + args.steps):
        res = cpu_bench.time_cpu(spec, x, y, z, DRV, nproc=workers, slice_length=2000, kind=kind)
        if s >= args.warmup:
            times.append(res['seconds'])
    sec = float(numpy.mean(times))
    value = npts / sec
    sample = ('%d contiguous points of the 200^3 grid per step, %d worker processes, slices of 2000 points'
              % (npts, res['cores']))
    line = {'impl': 'reference', 'metric': 'grid points/s, rho + grad rho (FP64)', 'value': value, 'unit': 'points/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic (seeded random molecule and MO coefficients)',
            'config': {'workload': WORKLOAD, 'sample': sample},
            'cpu_baseline': {'value': value, 'unit': 'points/s', 'cores': res['cores'], 'kind': res['kind'],
                             'sample': sample, 'host_cores': cores, 'value_slice_1e4': res4['points_per_s']},
            'e2e': {'value': value, 'unit': 'points/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--cpu-points', type=int, default=0, help='size of the CPU sample (0: 20000 per worker, about 12 s per slice length)')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg (profiling runs)')
    ap.add_argument('--no-e2e', action='store_true', help='skip the end-to-end leg (profiling runs)')
    ap.add_argument('--no-peaks', action='store_true', help='skip the FP64 peak microbenchmarks (profiling runs)')
    ap.add_argument('--no-also', action='store_true', help='skip the secondary request types (profiling runs)')
    ap.add_argument('--no-strong', action='store_true', help='skip the strong-scaling block')
    ap.add_argument('--no-c4', action='store_true', help='strong-scaling block without the 3000-AO / 256^3 case')
    ap.add_argument('--no-latency', action='store_true', help='skip the small-call latency block')
    args = ap.parse_args()
    # stdout carries the ONE JSON line only: libraries that write to file descriptor 1 (NCCL's version banner at
    # communicator creation) are sent to stderr; the JSON line goes to the original descriptor
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, 'w')
    if args.impl == 'reference':
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_b200(args)


if __name__ == '__main__':
    main()
