"""Drop-in for the reference's compiled module `orbkit.cy_core` (cy_core.pyx:21-101).

Same function names, argument order, dtypes and return values; the work is done by the sm_100a
kernels in libokb200.so through the C ABI (include/okb200.h).  Like the Cython buffers
(`mode="c"`, typed, `not None`) the array arguments must be C-contiguous NumPy arrays of the exact
dtype (float64 / intc): anything else raises ValueError (TypeError for None), as the reference does.
"""
import numpy

from . import _lib
from .engine import get_engine


def _buf(name, a, dtype, ndim):
    if a is None:
        raise TypeError("Argument '%s' must not be None" % name)
    if not isinstance(a, numpy.ndarray):
        raise TypeError("Argument '%s' has incorrect type (expected numpy.ndarray, got %s)"
                        % (name, type(a).__name__))
    if a.dtype != dtype:
        raise ValueError("Buffer dtype mismatch for '%s', expected '%s' but got '%s'"
                         % (name, numpy.dtype(dtype).name, a.dtype.name))
    if a.ndim != ndim:
        raise ValueError("Buffer has wrong number of dimensions for '%s' (expected %d, got %d)"
                         % (name, ndim, a.ndim))
    if not a.flags['C_CONTIGUOUS']:
        raise ValueError("ndarray '%s' is not C-contiguous" % name)
    return a


def aonorm(lx, ly, lz, alpha, is_normalized):
    """Primitive normalisation (cy_core.pyx:21 -> ao_norm, c_support.c:177-188)."""
    return _lib.load().okb_aonorm(int(lx), int(ly), int(lz), float(alpha), int(is_normalized))


def aoxyz(x, y, z, lx, ly, lz, alpha, drv):
    """Derivative prefactor of x^lx y^ly z^lz exp(-alpha r^2) (cy_core.pyx:24 -> get_ao_xyz)."""
    return _lib.load().okb_aoxyz(float(x), float(y), float(z), int(lx), int(ly), int(lz), float(alpha),
                                 int(drv))


def aocreator(lxlylz, assign, ao_coeffs, pnum_list, geo_spec, atom_indices, x, y, z, drv, is_normalized,
              exact_mixed=False):
    """All contracted Cartesian AOs (or one derivative `drv` in 0..9) on a vector grid
    (cy_core.pyx:51-78).  Returns float64 [n_cart, npts]."""
    lxlylz = _buf('lxlylz', lxlylz, numpy.intc, 2)
    assign = _buf('assign', assign, numpy.intc, 1)
    ao_coeffs = _buf('ao_coeffs', ao_coeffs, numpy.float64, 2)
    pnum_list = _buf('pnum_list', pnum_list, numpy.intc, 1)
    geo_spec = _buf('geo_spec', geo_spec, numpy.float64, 2)
    atom_indices = _buf('atom_indices', atom_indices, numpy.intc, 1)
    x = _buf('x', x, numpy.float64, 1)
    y = _buf('y', y, numpy.float64, 1)
    z = _buf('z', z, numpy.float64, 1)
    if not (len(x) == len(y) == len(z)):
        raise ValueError('Dimensions of x-, y-, and z- coordinate differ!')
    if len(assign) != len(pnum_list) or len(assign) != len(atom_indices):
        raise ValueError('assign, pnum_list and atom_indices differ in length')
    eng = get_engine()
    out = numpy.zeros((lxlylz.shape[0], x.shape[0]), dtype=numpy.float64)
    if x.shape[0] == 0 or lxlylz.shape[0] == 0:
        return out
    _lib.check(eng.lib.okb_aocreator(
        eng.ctx, _lib.iptr(lxlylz), _lib.iptr(assign), _lib.dptr(ao_coeffs), _lib.iptr(pnum_list),
        _lib.dptr(geo_spec), _lib.iptr(atom_indices), len(assign), lxlylz.shape[0], ao_coeffs.shape[0],
        geo_spec.shape[0], _lib.dptr(x), _lib.dptr(y), _lib.dptr(z), x.shape[0], int(drv),
        int(is_normalized), _lib.OKB_FLAG_EXACT_MIXED if exact_mixed else 0, _lib.dptr(out)))
    return out


def lcreator(ao_list, lxlylz, coeff_list, at_pos, x, y, z, ao_num, pnum, drv, is_normalized):
    """One contraction, written in place into the first `ao_num` rows of `ao_list`
    (cy_core.pyx:29-47 -> c_lcreator, c_grid-based.c:9-79)."""
    ao_list = _buf('ao_list', ao_list, numpy.float64, 2)
    lxlylz = _buf('lxlylz', lxlylz, numpy.intc, 2)
    coeff_list = _buf('coeff_list', coeff_list, numpy.float64, 2)
    at_pos = _buf('at_pos', at_pos, numpy.float64, 1)
    x = _buf('x', x, numpy.float64, 1)
    y = _buf('y', y, numpy.float64, 1)
    z = _buf('z', z, numpy.float64, 1)
    if ao_list.shape[0] < ao_num or lxlylz.shape[0] < ao_num or coeff_list.shape[0] < pnum:
        raise ValueError('lcreator: arrays shorter than ao_num / pnum')
    if x.shape[0] == 0:
        return None
    eng = get_engine()
    _lib.check(eng.lib.okb_lcreator(
        eng.ctx, _lib.dptr(ao_list), ao_list.shape[1], _lib.iptr(lxlylz), _lib.dptr(coeff_list),
        _lib.dptr(at_pos), _lib.dptr(x), _lib.dptr(y), _lib.dptr(z), x.shape[0], int(ao_num), int(pnum),
        int(drv), int(is_normalized), 0))
    return None


def mocreator(ao_list, mo_coeffs):
    """mo[i,j] = sum_k mo_coeffs[i,k] * ao_list[k,j] (cy_core.pyx:82-101) as an FP64 tile GEMM."""
    ao_list = _buf('ao_list', ao_list, numpy.float64, 2)
    mo_coeffs = _buf('mo_coeffs', mo_coeffs, numpy.float64, 2)
    if mo_coeffs.shape[1] != ao_list.shape[0]:
        raise ValueError('mocreator: mo_coeffs has %d columns but ao_list has %d rows'
                         % (mo_coeffs.shape[1], ao_list.shape[0]))
    out = numpy.zeros((mo_coeffs.shape[0], ao_list.shape[1]), dtype=numpy.float64)
    if out.size == 0 or ao_list.shape[0] == 0:
        return out
    eng = get_engine()
    _lib.check(eng.lib.okb_mocreator(eng.ctx, _lib.dptr(ao_list), _lib.dptr(mo_coeffs), ao_list.shape[0],
                                     ao_list.shape[1], mo_coeffs.shape[0], _lib.dptr(out)))
    return out
