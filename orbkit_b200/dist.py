"""Multi-GPU sharding of the grid path: one process per GPU (torch.distributed, NCCL over NVLink).

The reference parallelises by handing contiguous point ranges to forked workers
(orbkit/core.py:437-447, 503-536; omp_functions.py:31-95).  Here each rank owns one contiguous
point range; every grid point is independent given the (replicated) basis tables and MO
coefficients, so there is NO data-path collective.

Assembly of host results (the reference copies slice results into one preallocated array,
core.py:527-536): the ranks of one node share ONE host array in POSIX shared memory
(`shared_host_array`); every rank's kernel output streams device -> host straight into its own
point range of that array (page-locked by cudaHostRegister, so the copies overlap the compute), and
a barrier completes the result on every rank.  No NVLink or PCIe traffic beyond each shard's own
bytes.  NVLink/NCCL is used only for
  * `gather_points` when the caller wants the full result as a DEVICE tensor, and
  * the all-reduce of integrated quantities (MO norms, electron count).
With the gloo backend the same code runs on CPU tensors (used by the world_size-2 tests).
"""
import atexit
import mmap
import os
import uuid

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


# ---- one host array shared by the ranks of a node ------------------------------------------------------
class _Segment:
    """an mmap of a /dev/shm file that every rank of the job has opened (the file is unlinked once all
    ranks hold it: the memory lives as long as a mapping does)"""

    def __init__(self, nbytes, name, create):
        path = os.path.join('/dev/shm', name)
        fd = os.open(path, os.O_RDWR | (os.O_CREAT | os.O_EXCL if create else 0), 0o600)
        try:
            if create:
                os.ftruncate(fd, nbytes)
            self.map = mmap.mmap(fd, nbytes)
        finally:
            os.close(fd)
        self.nbytes = nbytes
        self.path = path
        self.registered = False

    def register_cuda(self, ranges=None):
        """page-lock the mapping for this process's CUDA context (async D2H into it then overlaps compute).
        `ranges`: byte ranges [(offset, nbytes), ...] this rank writes -- only those pages are locked (page-locking
        is the expensive part of a new segment: every rank locking all of a 0.5 GB array took 1.7 s at 8 ranks)."""
        if self.registered:
            return True
        try:
            import ctypes
            import torch
            base = ctypes.addressof(ctypes.c_char.from_buffer(self.map))
            rt = torch.cuda.cudart()
            if not ranges:
                ranges = [(0, self.nbytes)]
            page, ok, last_end = mmap.PAGESIZE, True, 0
            for off, n in sorted(ranges):
                a = max(off // page * page, last_end)                # page aligned, never twice the same page
                b = min(-(-(off + n) // page) * page, -(-self.nbytes // page) * page)
                if b > a:
                    ok = ok and int(rt.cudaHostRegister(base + a, b - a, 0)) == 0
                    last_end = b
            self.registered = ok
        except Exception:
            self.registered = False
        return self.registered


_segments = {}      # (nbytes, generation) -> _Segment
_generation = {}    # nbytes -> how many arrays of this size were handed out


def shared_host_array(shape, pin=True, own=None):
    """float64 array of `shape` in node-shared memory, the SAME memory on every rank (collective call:
    every rank must call it with the same shape).  Two segments per size are used in turn, so a result
    stays valid until the second-next call with the same shape.  `own = (p0, p1)`: this rank writes the
    entries [p0, p1) of the last axis of every row; with at most 64 rows only those pages are page-locked."""
    import torch.distributed as dist
    rank, world = rank_world()
    nbytes = max(int(numpy.prod(shape)) * 8, 8)
    gen = _generation.get(nbytes, 0)
    _generation[nbytes] = gen + 1
    key = (nbytes, gen & 1)
    seg = _segments.get(key)
    if seg is None:
        name = ['okb200_%s' % uuid.uuid4().hex if rank == 0 else None]
        if rank == 0:
            seg = _Segment(nbytes, name[0], create=True)
        dist.broadcast_object_list(name, src=0)
        if rank != 0:
            seg = _Segment(nbytes, name[0], create=False)
        dist.barrier()
        if rank == 0:
            os.unlink(seg.path)
        if pin:
            ranges = None
            rows = int(numpy.prod(shape[:-1])) if len(shape) > 1 else 1
            if own is not None and rows <= 64:
                n_last = int(shape[-1])
                ranges = [((r * n_last + own[0]) * 8, (own[1] - own[0]) * 8) for r in range(rows) if own[1] > own[0]]
                ranges = ranges or [(0, 8)]
            seg.register_cuda(ranges)
        _segments[key] = seg
    return numpy.frombuffer(seg.map, dtype=numpy.float64, count=int(numpy.prod(shape))).reshape(shape)


@atexit.register
def _drop_segments():
    _segments.clear()
