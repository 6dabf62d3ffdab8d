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


_local_only = False


class local_only:
    """context manager: inside it this process works alone (is_distributed() is False) although a process group
    exists -- bench.py times the one-GPU run of a strong-scaling pair on rank 0 with it"""

    def __enter__(self):
        global _local_only
        self.prev, _local_only = _local_only, True

    def __exit__(self, *exc):
        global _local_only
        _local_only = self.prev


def is_distributed():
    if _local_only:
        return False
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
def single_node():
    """True when every rank of the job runs on this host (the shared-memory assembly needs that); decided once by
    an all-gather of (hostname, boot id)"""
    global _single_node
    if _single_node is None:
        import socket
        import torch.distributed as dist
        if os.environ.get('OKB_DIST_FORCE_GATHER'):
            _single_node = False
            return _single_node
        try:
            boot = open('/proc/sys/kernel/random/boot_id').read().strip()
        except OSError:
            boot = ''
        mine = (socket.gethostname(), boot)
        everyone = [None] * dist.get_world_size()
        dist.all_gather_object(everyone, mine)
        _single_node = all(e == everyone[0] for e in everyone)
    return _single_node


_single_node = None


def all_reduce_min(values, device=None):
    """element-wise minimum of a small integer vector over all ranks (agreement on which cached segments are free)"""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.int64)
    if dist.get_backend() == 'nccl':
        t = t.cuda(device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return [int(v) for v in t.cpu().tolist()]


class _Segment:
    """an mmap of a /dev/shm file that every rank of the node has opened (the file is unlinked once all
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
        self.locked = []            # page-locked (address, nbytes) ranges of this process
        self.key = None
        self.user = None            # weak reference to the array handed out last

    def busy(self):
        """does the caller still hold the array (or a view of it) that was handed out last?"""
        return self.user is not None and self.user() is not None

    def register_cuda(self, ranges=None):
        """page-lock the mapping for this process's CUDA context (async D2H into it then overlaps compute).
        `ranges`: byte ranges [(offset, nbytes), ...] this rank writes -- only those pages are locked (page-locking
        is the expensive part of a new segment: every rank locking all of a 0.5 GB array took 1.7 s at 8 ranks)."""
        if self.locked:
            return True
        try:
            import ctypes
            import torch
            base = ctypes.addressof(ctypes.c_char.from_buffer(self.map))
            rt = torch.cuda.cudart()
            if not ranges:
                ranges = [(0, self.nbytes)]
            page, last_end = mmap.PAGESIZE, 0
            for off, n in sorted(ranges):
                a = max(off // page * page, last_end)                # page aligned, never twice the same page
                b = min(-(-(off + n) // page) * page, -(-self.nbytes // page) * page)
                if b > a:
                    if int(rt.cudaHostRegister(base + a, b - a, 0)) != 0:
                        self.release_cuda()
                        return False
                    self.locked.append((base + a, b - a))
                    last_end = b
        except Exception:
            self.release_cuda()
        return bool(self.locked)

    def release_cuda(self):
        if not self.locked:
            return
        try:
            import torch
            rt = torch.cuda.cudart()
            for addr, _ in self.locked:
                rt.cudaHostUnregister(addr)
        except Exception:
            pass
        self.locked = []


# Cached segments in creation order -- the SAME list on every rank, because segments are created, reused and evicted
# collectively on state all ranks agreed on (the all-reduced free flags below).
_segments = []
CACHE_BYTES = 8 << 30       # free segments beyond this total are unmapped, oldest first


def _drop(seg):
    seg.release_cuda()
    seg.user = None            # the mapping itself goes away with its last view


def shared_host_array(shape, pin=True, own=None):
    """float64 array of `shape` in node-shared memory, the SAME memory on every rank (collective call: every rank
    must call it with the same shape).  `own = (p0, p1)`: this rank writes the entries [p0, p1) of the last axis of
    every row; with at most 64 rows only those pages are page-locked.

    Life time: the array (and every view of it) stays valid for as long as ANY rank holds a reference to it -- a
    cached segment is handed out again only when every rank reported it free (one all-reduce of a few integers per
    call, which is also the barrier that orders the previous readers before the next writers).  Segments are keyed by
    (shape, own range), so the page-locked ranges always match the rows a rank writes; free segments beyond
    CACHE_BYTES are unmapped."""
    import weakref
    import torch.distributed as dist
    rank, world = rank_world()
    shape = tuple(int(v) for v in shape)
    count = int(numpy.prod(shape))
    nbytes = max(count * 8, 8)
    key = (shape, None if own is None else (int(own[0]), int(own[1])), bool(pin))
    # agreement: a segment is free only if no rank still holds its array.  (`key` differs between ranks only in the
    # own range, which is a function of the rank: equal shapes <=> equal positions in the list.)
    free = all_reduce_min([0 if s.busy() else 1 for s in _segments] + [1]) if _segments else [1]
    seg = None
    for s, f in zip(_segments, free):
        if f and s.key == key:
            seg = s
            break
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
        seg.key = key
        if pin:
            ranges = None
            rows = int(numpy.prod(shape[:-1])) if len(shape) > 1 else 1
            if own is not None and rows <= 64:
                n_last = int(shape[-1])
                ranges = [((r * n_last + own[0]) * 8, (own[1] - own[0]) * 8) for r in range(rows) if own[1] > own[0]]
                ranges = ranges or [(0, 8)]
            seg.register_cuda(ranges)
        # evict agreed-free segments, oldest first, while the cache is over its budget (same decision on every rank)
        total = nbytes + sum(s.nbytes for s in _segments)
        keep = []
        for s, f in zip(_segments, free):
            if f and total > CACHE_BYTES:
                total -= s.nbytes
                _drop(s)
            else:
                keep.append(s)
        _segments[:] = keep + [seg]
    root = numpy.frombuffer(seg.map, dtype=numpy.float64, count=count)
    seg.user = weakref.ref(root)     # every view (the reshaped array, its rows, slices of those) keeps `root` alive as its .base
    return root.reshape(shape)


def gather_rows(local, npts, p0, p1):
    """Multi-node fallback of the shared-memory assembly: every rank contributes the columns [p0, p1) of a
    (rows, npts) host result; all ranks return the full array (all-gather over the process group)."""
    import torch
    import torch.distributed as dist
    local = numpy.ascontiguousarray(local, dtype=numpy.float64).reshape((-1, p1 - p0))
    t = torch.from_numpy(local)
    if dist.get_backend() == 'nccl':
        t = t.cuda()
    return gather_points(t, npts).cpu().numpy()


@atexit.register
def _drop_segments():
    for s in _segments:
        _drop(s)
    _segments.clear()
