"""Output sink of the grid path: Gaussian cube files (orbkit/output/cube.py:5-101, output/high_level.py:59-342).

    cube_creator(data, filename, geo_info, geo_spec, comments='', labels=None)     cube.py:5-101
    main_output(data, qc, outputname, otype, drv, datalabels, dataindices)          high_level.py:59-342 (cube types only)

The reference formats every value with a Python string operation (about 1e6 values per second); here the data block is
formatted on the device (csrc/okb_text.cuh, C ABI okb_format_cube): correctly rounded '%.5E' in 13 columns, the
reference's line structure, byte-identical text.  The few header lines are built on the host exactly as the reference
builds them.  `data` may be a NumPy array (staged in slabs) or a device tensor (torch, float64, C-contiguous) -- then the
values never visit the host.  Other output types (HDF5, npz, Amira, VMD, obj, mayavi) are outside the hot path and raise
NotImplementedError.
"""
import gzip
import os

import numpy

from . import _lib, grid
from .display import display
from .engine import get_engine

CUBE_SYNONYMS = ('cube', 'cb')


def cube_header(n_sets, geo_info, geo_spec, comments='', labels=None):
    """the lines in front of the data block (cube.py:47-84)"""
    lines = 'orbkit calculation\n'
    lines += ' %s\n' % comments
    natoms = len(geo_info)
    lines += ('%d' % (-natoms if labels is not None else natoms)).rjust(5)
    for ii in range(3):
        lines += ('%0.6f' % grid.min_[ii]).rjust(12)
    if n_sets > 1:
        lines += ('%d' % n_sets).rjust(12)
    for ii in range(3):
        lines += '\n' + ('%d' % grid.N_[ii]).rjust(5)
        for jj in range(3):
            lines += ('%0.6f' % (grid.delta_[ii] if jj == ii else 0)).rjust(12)
    lines += '\n'
    for ii in range(natoms):
        lines += ('%d' % round(float(geo_info[ii][2]))).rjust(5)
        lines += ('%0.6f' % float(geo_info[ii][1])).rjust(12)
        for jj in range(3):
            lines += ('%0.6f' % geo_spec[ii][jj]).rjust(12)
        lines += '\n'
    if labels is not None:
        lines += ('%d' % n_sets).rjust(5)
        c = 0
        for j in labels:
            c += 1
            lines += str(j).rjust(5)
            if c % 9 == 8:
                lines += '\n'
        lines += '\n'
    return lines


def _is_device_tensor(data):
    return type(data).__module__.startswith('torch') and hasattr(data, 'is_cuda') and data.is_cuda


def cube_body(data):
    """the data block of a cube file as bytes: data (n_sets, Nx, Ny, Nz) NumPy array or CUDA tensor (cube.py:86-96)"""
    eng = get_engine()
    shape = tuple(int(s) for s in data.shape)
    n_sets, nx, ny, nz = shape
    nbytes = eng.lib.okb_cube_body_bytes(n_sets, nx, ny, nz)
    if nbytes == 0:
        return numpy.empty(0, dtype=numpy.uint8)
    try:                                  # page-locked (torch's caching host allocator): the text comes back at PCIe speed
        import torch
        text = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True).numpy()
    except Exception:
        text = numpy.empty(nbytes, dtype=numpy.uint8)
    if _is_device_tensor(data):
        import torch
        if data.dtype != torch.float64 or not data.is_contiguous():
            raise ValueError('device data must be a C-contiguous float64 tensor')
        if data.device.index != eng.device:
            raise ValueError('device data lives on cuda:%s, the engine on cuda:%d' % (data.device.index, eng.device))
        torch.cuda.current_stream(data.device).synchronize()    # the formatter runs on the context's own stream
        ptr, flags = data.data_ptr(), _lib.OKB_FLAG_IN_DEVICE
    else:
        data = _lib.f64(data)
        ptr, flags = data.ctypes.data, 0
    _lib.check(eng.lib.okb_format_cube(eng.ctx, ptr, n_sets, nx, ny, nz, text.ctypes.data, nbytes, flags))
    return text


def cube_creator(data, filename, geo_info, geo_spec, comments='', labels=None, **kwargs):
    """Creates a plain text Gaussian cube file (cube.py:5-101): data of shape (Nx,Ny,Nz) or (Ndata,Nx,Ny,Nz);
    `filename` gets '.cube' appended unless it ends with cube, cb, cube.gz or cb.gz (gz: gzip-compressed)."""
    if not _is_device_tensor(data):
        data = numpy.asarray(data)
    if data.ndim < 3:
        raise AssertionError('data.ndim < ndim of grid')
    elif data.ndim == 3:
        data = data[None]
    elif data.ndim > 4:
        raise AssertionError('data.ndim > (ndim of grid) +2')
    if labels is not None:
        if labels is True or labels == 'auto':
            labels = list(range(len(data)))
        assert len(labels) == len(data)
        try:
            labels = [int(j) for j in labels]
        except ValueError:
            raise AssertionError('labels has to be list of integers.')
    assert tuple(data.shape[1:]) == tuple(grid.N_), 'The grid does not fit the data.'
    if not any(filename.endswith(ext) for ext in ['cube', 'cb', 'cube.gz', 'cb.gz']):
        filename += '.cube'
    head = cube_header(len(data), geo_info, geo_spec, comments=comments, labels=labels).encode('utf-8')
    body = cube_body(data)
    with (gzip.open(filename, 'wb') if filename.endswith('gz') else open(filename, 'wb')) as f:
        f.write(head)
        f.write(memoryview(body))
    return filename


H5_SYNONYMS = ('h5', 'hdf5')
NPZ_SYNONYMS = ('npz', 'numpy')


def _to_host(data):
    return data.detach().cpu().numpy() if _is_device_tensor(data) else numpy.asarray(data)


def hdf5_creator(data, filename, qcinfo=None, gname='', ftype='hdf5', mode='w', attrs={}, **kwargs):
    """HDF5 or .npz container with the reference's groups (orbkit/output/hdf5.py:10-54): `data`, `grid/{x,y,z,
    is_vector,is_regular}`, optional attributes and `qcinfo/...` (QCinfo.todict() with date and time).
    ftype 'hdf5'/'h5' needs h5py (ImportError otherwise); 'numpy'/'npz' writes a zip of .npy members."""
    import time
    from . import store
    if ftype.lower() in H5_SYNONYMS:
        write = store.hdf5_write
    elif ftype.lower() in NPZ_SYNONYMS:
        write = store.npz_write
    else:
        raise NotImplementedError('File format {0} not implemented for writing.'.format(ftype.lower()))
    fn = write(filename, mode=mode, gname=gname, data=numpy.asarray(data), **kwargs)
    write(fn, mode='a', gname=os.path.join(gname, 'grid'), x=grid.x, y=grid.y, z=grid.z,
          is_vector=bool(grid.is_vector), is_regular=bool(grid.is_regular))
    if attrs:
        write(fn, mode='a', gname=gname + '_attrs', **attrs)
    if qcinfo is not None:
        d = dict(qcinfo.todict())
        d['date'], d['time'] = time.strftime('%Y-%m-%d'), time.strftime('%H:%M:%S')
        write(fn, mode='a', gname=os.path.join(gname, 'qcinfo'), **d)
    return fn


def main_output(data, qc=None, outputname='data', otype='auto', gname='', drv=None, omit=[], datalabels='',
                dataindices=None, mode='w', **kwargs):
    """Writes `data` as cube file(s) with the reference's naming scheme (high_level.py:59-342, cube branch):
    data of shape N, (NDRV,)+N with `drv`, (Ndata,)+N, or (NDRV, Ndata)+N  ->  <outputname>[_<index>][_d<drv>].<ext>.
    Returns the list of files written."""
    if otype is None or otype == []:
        return []
    if isinstance(outputname, str) and '@' in outputname:
        outputname, gname = outputname.split('@')
    if isinstance(otype, str):
        otype = [otype]
    otype = list(otype)
    for i in range(len(otype)):
        if otype[i] == 'auto':
            outputname, ext = os.path.splitext(outputname)
            if ext != '' or len(otype) == 1:
                otype[i] = ext[1:]
    otype = [i for i in otype if i not in omit]
    other = [i for i in otype if i not in CUBE_SYNONYMS + H5_SYNONYMS + NPZ_SYNONYMS]
    if other:
        raise NotImplementedError('orbkit_b200 writes cube files ("cb"/"cube"), HDF5 ("h5") and "npz" containers; %r '
                                  'belong to the reference\'s output module' % (other,))
    if not otype:
        return []
    written = []
    base = outputname if isinstance(outputname, str) else outputname[-1]
    for t in otype:                     # containers take the data as it is, on any grid (high_level.py:298-311)
        if t in H5_SYNONYMS:
            display('\nSaving to Hierarchical Data Format file (HDF5)...\n\t' + base + '.' + t)
            written.append(hdf5_creator(_to_host(data), base + '.' + t, qcinfo=qc, gname=gname, ftype='hdf5', mode=mode))
        elif t in NPZ_SYNONYMS:
            display('\nSaving to a compressed .npz archive...\n\t' + base + '.npz')
            written.append(hdf5_creator(_to_host(data), base, qcinfo=qc, gname=gname, ftype='numpy', mode=mode))
    otype = [i for i in otype if i in CUBE_SYNONYMS]
    if not otype:
        return written
    ext = otype[0]
    if grid.is_vector and not grid.is_regular:
        display('For a non-regular vector grid (`if grid.is_vector and not grid.is_regular`)')
        display('only HDF5 is available as output format...')
        display('Skipping all other formats...')
        return written
    if qc is None:
        display('\nFor cube file output `qc` is a required keyword parameter in `main_output`.')
        return written
    dev = _is_device_tensor(data)
    if not dev:
        data = numpy.asarray(data)
    is_regular_vector = grid.is_vector and grid.is_regular
    dims = 1 if grid.is_vector else 3
    if drv is not None and isinstance(drv, str):
        drv = [drv]
    if data.ndim < dims:
        display('data.ndim < ndim of grid')
        return written
    elif data.ndim == dims:
        data = data[None, None]
    elif data.ndim == dims + 1:
        data = data[:, None] if drv is not None else data[None]
    elif data.ndim == dims + 2:
        if drv is None or len(drv) != data.shape[0]:
            drv = list(range(data.shape[0]))
    else:
        display('data.ndim > (ndim of grid) +2')
        return written
    if is_regular_vector:
        # a regular grid stored as a vector: the point index is x-major, z fastest (cy_grid.pyx:22-29)
        display('\nConverting the regular 1d vector grid to a 3d regular grid.')
        data = data.reshape(tuple(data.shape[:2]) + tuple(grid.N_))
    isstr = isinstance(outputname, str)
    if isinstance(datalabels, str):
        if data.shape[1] > 1:
            datalabels = [str(i) + ',' + datalabels for i in range(data.shape[1])]
        else:
            datalabels = [datalabels]
    if drv is not None:
        fid, label_id, it = '%(f)s_d%(d)s.', 'd/d%(d)s %(f)s', list(enumerate(drv))
    elif data.shape[0] > 1:
        fid, label_id, it = '%(f)s_%(d)s.', '%(d)s %(f)s', [(i, i) for i in range(data.shape[0])]
    else:
        fid, label_id, it = '%(f)s.', '%(f)s', [(0, None)]
    for idrv, jdrv in it:
        for idata in range(data.shape[1]):
            if isstr:
                index = str(idata) if dataindices is None else str(dataindices[idata])
                f = {'f': outputname + '_' + index if data.shape[1] > 1 else outputname, 'd': jdrv}
            else:
                f = {'f': outputname[idata], 'd': jdrv}
            label = label_id % {'f': datalabels[idata], 'd': jdrv}
            filename = fid % f + ext
            display('\nSaving to cube file...\n\t' + filename)
            cube_creator(data[idrv, idata], filename, qc.geo_info, qc.geo_spec, comments=label, **kwargs)
            written.append(filename)
    return written
