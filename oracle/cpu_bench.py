"""Timed CPU baseline of the grid path (TEST/BENCH INFRASTRUCTURE ONLY -- see oracle/oracle.py).

Runs the reference's own CPU implementation of rho + derivatives -- oracle/_ref (the reference's
c_lcreator + Cython mocreator compiled unmodified) when it is present, else the C port -- over
contiguous point slices in a pool of worker PROCESSES, mirroring the reference's
multiprocessing.Pool driver (orbkit/core.py:503-536, slice_length=1e4 by default).  Workers are
spawned (not forked) so that a CUDA context in the parent is never inherited.
"""
import multiprocessing
import os
import sys
import time

import numpy

_HERE = os.path.dirname(os.path.abspath(__file__))
_STATE = {}


def _init(repo, spec, spherical_p, kind):
    for p in (repo, _HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oracle
    from orbkit_b200 import synth
    _STATE['oracle'] = oracle
    _STATE['qc'] = synth.to_qcinfo(spec)
    _STATE['kind'] = kind


def _work(args):
    x, y, z, drv = args
    o = _STATE['oracle']
    res = o.slice_rho(_STATE['qc'], x, y, z, drv=drv, kind=_STATE['kind'])
    return float(res[0].sum())


def default_kind():
    sys.path.insert(0, _HERE)
    import oracle
    return 'ref' if oracle.have_ref() else 'port'


def time_cpu(spec, x, y, z, drv, nproc=None, slice_length=2000, kind=None, repeats=1):
    """Evaluate rho (+ delta_rho for `drv`) on the vector-grid sample (x,y,z) with `nproc` worker
    processes; returns dict(points_per_s, seconds, cores, kind, npts)."""
    repo = os.path.dirname(_HERE)
    nproc = nproc or os.cpu_count() or 1
    kind = kind or default_kind()
    npts = len(x)
    jobs = [(x[i:i + slice_length], y[i:i + slice_length], z[i:i + slice_length], drv)
            for i in range(0, npts, slice_length)]
    nproc = max(1, min(nproc, len(jobs)))
    ctx = multiprocessing.get_context('spawn')
    with ctx.Pool(nproc, initializer=_init, initargs=(repo, spec, None, kind)) as pool:
        pool.map(_work, jobs[:nproc])            # warm-up: imports, page-in
        best = None
        for _ in range(repeats):
            t0 = time.perf_counter()
            pool.map(_work, jobs)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return {'points_per_s': npts / best, 'seconds': best, 'cores': nproc,
            'kind': 'reference' if kind == 'ref' else 'port', 'npts': npts}
