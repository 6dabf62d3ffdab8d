"""Streaming result sinks for `save_hdf5=` (orbkit/core.py:478-501, 584-603; orbkit/tools.py:303-314) and the
npz / HDF5 containers of `output.main_output` (orbkit/output/hdf5.py:10-100).

The reference creates chunked HDF5 datasets (`tools.zeros(..., hdf5_file=...)`), fills them slice by slice and reads
them back in full at the end.  Here the results never exist as one host array: point slabs come off the device
into a ring of two page-locked buffers and a writer thread puts every slab at its place in the file with
positional writes (`os.pwrite`, one per output row), overlapped with the evaluation of the next slab.  Peak host
memory is two slabs, whatever the size of the result.

Two containers with the SAME dataset names as the reference ('rho', 'delta_rho', 'mo_list' / 'ao_list',
'grid/x', 'grid/y', 'grid/z', 'grid/is_vector', 'grid/is_regular'):

  * HDF5 through h5py when it is importable and the file name ends in .h5 / .hdf5 (datasets of shape (..., npts)
    with the attribute `shape`, as the reference writes them);
  * otherwise an uncompressed `.npz` (a zip of .npy members, the layout of the reference's `npz_write`: member
    `<group>/<name>.npy`), written as plain .npy part files first and zipped member by member at the end (disk to
    disk, no host copy).  `numpy.load(path)` reads it; `ResultStore.arrays()` memory-maps the members in place.
"""
import os
import shutil
import struct
import zipfile
from concurrent.futures import ThreadPoolExecutor

import numpy

SLAB_BYTES = 64 << 20          # bytes of one staging slab (two are in flight)


def have_h5py():
    try:
        import h5py  # noqa: F401
        return True
    except Exception:
        return False


def wants_hdf5(path):
    return str(path).lower().endswith(('.h5', '.hdf5'))


def npz_name(path):
    path = str(path)
    return path if path.lower().endswith('.npz') else path + '.npz'


def _lead_dims(shape, n_points):
    """the leading axes of `shape` in front of the trailing grid axes that flatten to n_points points"""
    shape = tuple(shape)
    tail, k = 1, len(shape)
    while k > 0 and tail != n_points:
        k -= 1
        tail *= shape[k]
    if tail != n_points:
        raise ValueError('shape %s does not end in %d grid points' % (shape, n_points))
    return shape[:k]


# ---------------------------------------------------------------------------------------------------------------------
class _NpyDataset:
    """a .npy file of known shape that is filled by column ranges of its last (flattened point) axis"""

    def __init__(self, path, shape, n_points, create):
        self.path, self.shape, self.npts = path, tuple(int(v) for v in shape), int(n_points)
        self.rows = int(numpy.prod(self.shape)) // max(self.npts, 1) if self.npts else 0
        if create:
            with open(path, 'wb') as f:
                numpy.lib.format.write_array_header_2_0(
                    f, {'descr': '<f8', 'fortran_order': False, 'shape': self.shape})
                self.offset = f.tell()
                f.truncate(self.offset + 8 * int(numpy.prod(self.shape)))
        else:
            with open(path, 'rb') as f:
                numpy.lib.format.read_magic(f)
                numpy.lib.format.read_array_header_2_0(f)
                self.offset = f.tell()
        self.fd = os.open(path, os.O_RDWR)

    def write_cols(self, p0, p1, block):
        """block: C-contiguous float64 (rows, p1 - p0) -> the columns [p0, p1) of every row"""
        block = numpy.ascontiguousarray(block, dtype=numpy.float64).reshape((self.rows, p1 - p0))
        for r in range(self.rows):
            os.pwrite(self.fd, block[r].data, self.offset + 8 * (r * self.npts + p0))

    def close(self):
        if self.fd is not None:
            os.close(self.fd)
            self.fd = None


class _H5Dataset:
    def __init__(self, dset, shape, n_points):
        self.dset, self.shape, self.npts = dset, tuple(shape), int(n_points)

    def write_cols(self, p0, p1, block):
        self.dset[..., p0:p1] = numpy.asarray(block).reshape(self.dset.shape[:-1] + (p1 - p0,))

    def close(self):
        pass


class ResultStore:
    """One output file of a `save_hdf5=` request.  Under torch.distributed every rank opens the same part files
    (rank 0 creates them) and writes its own point range; rank 0 finalises."""

    def __init__(self, path, rank=0, world=1, barrier=None):
        self.rank, self.world, self.barrier = rank, world, barrier or (lambda: None)
        self.h5 = None
        self.sets = {}
        self.small = {}
        self._writer = ThreadPoolExecutor(1)
        self._pending = []
        if wants_hdf5(path) and have_h5py():
            if world > 1:
                raise NotImplementedError('save_hdf5 into an HDF5 file from several ranks needs parallel HDF5; '
                                          'use a .npz name')
            import h5py
            self.path = str(path)
            self.h5 = h5py.File(self.path, 'w')
        else:
            self.path = npz_name(path)
            self.parts = self.path + '.parts'
            if rank == 0:
                if os.path.isdir(self.parts):
                    shutil.rmtree(self.parts)
                os.makedirs(self.parts)
            self.barrier()

    # -- small members (grid axes, flags) ------------------------------------------------------------------------------
    def put(self, name, value):
        if self.h5 is not None:
            self.h5[name] = value
        else:
            self.small[name] = numpy.asarray(value)

    # -- big members ------------------------------------------------------------------------------------------------------
    def create(self, name, shape, n_points, chunk_points=None):
        """dataset `name` of final shape `shape` whose trailing axes flatten to `n_points` points"""
        shape = tuple(int(v) for v in shape)
        if self.h5 is not None:
            rows = _lead_dims(shape, n_points)
            vshape = rows + (n_points,)
            chunks = rows + (min(n_points, int(chunk_points or 10000)),) if n_points else None
            d = self.h5.create_dataset(name, vshape, dtype=numpy.float64, chunks=chunks)
            d.attrs['shape'] = shape
            ds = _H5Dataset(d, shape, n_points)
        else:
            part = os.path.join(self.parts, name.replace('/', '__') + '.npy')
            if self.rank == 0:
                ds = _NpyDataset(part, shape, n_points, create=True)
            self.barrier()
            if self.rank != 0:
                ds = _NpyDataset(part, shape, n_points, create=False)
        self.sets[name] = ds
        return ds

    def write_async(self, ds, p0, p1, block):
        """queue `block` for writing; returns a future (the caller must not reuse `block` before it is done)"""
        fut = self._writer.submit(ds.write_cols, p0, p1, block)
        self._pending.append(fut)
        return fut

    def drain(self):
        for f in self._pending:
            f.result()
        self._pending = []

    # -- finish ----------------------------------------------------------------------------------------------------------
    def close(self):
        self.drain()
        self._writer.shutdown()
        for ds in self.sets.values():
            ds.close()
        if self.h5 is not None:
            self.h5.close()
            return
        self.barrier()                           # every rank's slabs are in the part files
        if self.rank == 0:
            with zipfile.ZipFile(self.path, 'w', zipfile.ZIP_STORED, allowZip64=True) as zf:
                for name, val in self.small.items():
                    with zf.open(name + '.npy', 'w') as f:
                        numpy.lib.format.write_array(f, val, allow_pickle=False)
                for name, ds in self.sets.items():
                    zf.write(ds.path, arcname=name + '.npy')      # streamed from disk, stored uncompressed
            shutil.rmtree(self.parts)
        self.barrier()

    def arrays(self):
        """{name: read-only array} of the big members: memory maps into the finished .npz (no host copy until the
        caller touches the data); for HDF5 the datasets are read back in full like the reference does"""
        out = {}
        if self.h5 is not None:
            import h5py
            with h5py.File(self.path, 'r') as f:
                for name, ds in self.sets.items():
                    out[name] = f[name][...].reshape(ds.shape)
            return out
        with zipfile.ZipFile(self.path) as zf, open(self.path, 'rb') as raw:
            for name, ds in self.sets.items():
                info = zf.getinfo(name + '.npy')
                raw.seek(info.header_offset + 26)
                n_name, n_extra = struct.unpack('<HH', raw.read(4))
                raw.seek(info.header_offset + 30 + n_name + n_extra)
                numpy.lib.format.read_magic(raw)
                numpy.lib.format.read_array_header_2_0(raw)
                if int(numpy.prod(ds.shape)) == 0:
                    out[name] = numpy.zeros(ds.shape)
                else:
                    out[name] = numpy.memmap(self.path, dtype='<f8', mode='r', offset=raw.tell(), shape=ds.shape)
        return out


# ---------------------------------------------------------------------------------------------------------------------
# npz / hdf5 containers of main_output (orbkit/output/hdf5.py)
# ---------------------------------------------------------------------------------------------------------------------
def npz_write(filename, gname='', mode='w', compress=True, **namedict):
    """every keyword becomes the member `<gname>/<key>.npy` of the zip archive `filename` (None values are
    skipped; dictionaries become sub-groups) -- the container layout of the reference's npz_write"""
    filename = npz_name(filename)
    comp = zipfile.ZIP_DEFLATED if compress else zipfile.ZIP_STORED
    with zipfile.ZipFile(filename, mode=mode, compression=comp, allowZip64=True) as zf:
        def put(group, key, val):
            if val is None:
                return
            if isinstance(val, dict):
                for k, v in val.items():
                    put(os.path.join(group, key), k, v)
                return
            arr = numpy.asanyarray(val)
            if arr.dtype == object:
                arr = numpy.asarray(arr, dtype=str)
            with zf.open(os.path.join(group, key + '.npy'), 'w', force_zip64=True) as f:
                numpy.lib.format.write_array(f, arr, allow_pickle=False)
        for key, val in namedict.items():
            put(gname, key, val)
    return filename


def hdf5_write(filename, gname='', mode='w', **namedict):
    """the same through h5py (raises ImportError when h5py is not installed)"""
    import h5py
    if not str(filename).lower().endswith(('.h5', '.hdf5')):
        filename = str(filename) + '.h5'
    with h5py.File(filename, mode) as f:
        group = f.require_group(gname) if gname else f

        def put(g, key, val):
            if isinstance(val, dict):
                sub = g.require_group(key)
                for k, v in val.items():
                    put(sub, k, v)
            elif isinstance(val, (list, numpy.ndarray)):
                arr = numpy.array(val)
                if arr.dtype.kind in 'UO':
                    arr = numpy.asarray(arr, dtype='S')
                g[key] = arr
            else:
                g.attrs[key] = str(val) if val is None or isinstance(val, bool) else val
        for key, val in namedict.items():
            put(group, key, val)
    return filename
