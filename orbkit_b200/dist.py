"""Multi-GPU sharding of the grid path: one process per GPU (torch.distributed, NCCL over NVLink).

The reference parallelises by handing contiguous point ranges to forked workers
(orbkit/core.py:437-447, 503-536; omp_functions.py:31-95).  Here each rank owns one contiguous
point range; every grid point is independent given the (replicated) basis tables and MO
coefficients, so there is NO data-path collective.  NVLink is used only for
  * the final gather of the output shards (all_gather of equally padded shards), and
  * the all-reduce of integrated quantities (MO norms, electron count).
With the gloo backend the same code runs on CPU tensors (used by the world_size-2 tests).
"""
import numpy

ALIGN = 128   # shard boundaries fall on CTA tile boundaries


def is_distributed():
    try:
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    except Exception:
        return False


def rank_world():
    if not is_distributed():
        return 0, 1
    import torch.distributed as dist
    return dist.get_rank(), dist.get_world_size()


def shard_range(npts, rank, world, align=ALIGN):
    """Contiguous [p0, p1) of rank `rank`: npts split into `world` ranges balanced to +-1 tile."""
    ntiles = (npts + align - 1) // align
    base, extra = divmod(ntiles, world)
    t0 = rank * base + min(rank, extra)
    t1 = t0 + base + (1 if rank < extra else 0)
    return min(t0 * align, npts), min(t1 * align, npts)


def shard_sizes(npts, world, align=ALIGN):
    return [shard_range(npts, r, world, align) for r in range(world)]


def gather_points(local, npts, device_tensor=True):
    """All-gather shards along the LAST axis.  `local` is a torch tensor (..., n_local) living on the
    backend's device (CUDA for nccl, CPU for gloo); returns the full (..., npts) tensor on every rank."""
    import torch
    import torch.distributed as dist
    rank, world = rank_world()
    if world == 1:
        return local
    ranges = shard_sizes(npts, world)
    nmax = max(b - a for a, b in ranges)
    lead = tuple(local.shape[:-1])
    padded = torch.zeros(lead + (nmax,), dtype=local.dtype, device=local.device)
    padded[..., :local.shape[-1]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    full = torch.empty(lead + (npts,), dtype=local.dtype, device=local.device)
    for (a, b), part in zip(ranges, parts):
        full[..., a:b] = part[..., :b - a]
    return full


def all_reduce_sum(values, device=None):
    """Sum a small float64 vector (MO norms, electron count) over all ranks; returns numpy."""
    values = numpy.asarray(values, dtype=numpy.float64)
    if not is_distributed():
        return values
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(values.copy())
    if dist.get_backend() == 'nccl':
        t = t.cuda(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
