"""Drop-in for the reference's compiled module `orbkit.cy_overlap` (cy_overlap.pyx:24-206): same names, argument order
and typed-buffer requirements.

    aooverlap(...)            analytic overlap matrix of contracted Cartesian Gaussians -> CUDA (okb_aooverlap,
                              csrc/okb_overlap.cuh), same primitive-pair order as the reference
    mooverlap / mooverlapmatrix   <mo_a| S |mo_b>: two dense contractions on the FP64 tensor cores (okb_ci_td)
    ommited_cca_norm / tmol_aomix_norm   per-function factors of double factorials (host arithmetic on a few integers)
"""
import numpy

from . import _lib
from .engine import get_engine


def _typed(a, dtype, ndim, name):
    if a is None:
        raise TypeError("Argument '%s' must not be None" % name)
    if not isinstance(a, numpy.ndarray):
        raise TypeError("Argument '%s' has incorrect type (expected numpy.ndarray, got %s)" % (name, type(a).__name__))
    if a.dtype != dtype:
        raise ValueError("Buffer dtype mismatch, expected '%s' but got '%s'" % ('double' if dtype == numpy.float64 else 'int',
                                                                                 a.dtype.name))
    if a.ndim != ndim:
        raise ValueError('Buffer has wrong number of dimensions (expected %d, got %d)' % (ndim, a.ndim))
    if not a.flags['C_CONTIGUOUS']:
        raise ValueError('ndarray is not C-contiguous')
    return a


def _dfact(n):
    """n!! with (<= 0)!! = 1 (c_support.c:190-195)"""
    r = numpy.ones_like(n, dtype=numpy.float64)
    n = numpy.array(n, dtype=numpy.int64)
    while (n > 1).any():
        r = numpy.where(n > 1, r * n, r)
        n = n - 2
    return r


def ommited_cca_norm(lxlylz):
    """sqrt((2lx-1)!! (2ly-1)!! (2lz-1)!! / (2l-1)!!): the Cartesian normalisation omitted in the CCA standard
    (cy_overlap.pyx:24-52)"""
    l = _typed(lxlylz, numpy.intc, 2, 'lxlylz').astype(numpy.int64)
    num = _dfact(2 * l[:, 0] - 1) * _dfact(2 * l[:, 1] - 1) * _dfact(2 * l[:, 2] - 1)
    return numpy.sqrt(num / _dfact(2 * l.sum(axis=1) - 1))


def tmol_aomix_norm(lxlylz):
    """sqrt((2lx-1)!! (2ly-1)!! (2lz-1)!!): Turbomole / ORCA (cy_overlap.pyx:54-72)"""
    l = _typed(lxlylz, numpy.intc, 2, 'lxlylz').astype(numpy.int64)
    return numpy.sqrt(_dfact(2 * l[:, 0] - 1) * _dfact(2 * l[:, 1] - 1) * _dfact(2 * l[:, 2] - 1))


def aooverlap(geo_spec_a, geo_spec_b, lxlylz_a, lxlylz_b, assign, ao_coeffs, pnum_list, atom_indices, drv, is_normalized):
    f, i = numpy.float64, numpy.intc
    geo_spec_a, geo_spec_b = _typed(geo_spec_a, f, 2, 'geo_spec_a'), _typed(geo_spec_b, f, 2, 'geo_spec_b')
    lxlylz_a, lxlylz_b = _typed(lxlylz_a, i, 2, 'lxlylz_a'), _typed(lxlylz_b, i, 2, 'lxlylz_b')
    assign, pnum_list = _typed(assign, i, 1, 'assign'), _typed(pnum_list, i, 1, 'pnum_list')
    ao_coeffs, atom_indices = _typed(ao_coeffs, f, 2, 'ao_coeffs'), _typed(atom_indices, i, 1, 'atom_indices')
    n = lxlylz_a.shape[0]
    if geo_spec_a.shape != geo_spec_b.shape or geo_spec_a.shape[1] != 3 or lxlylz_b.shape != lxlylz_a.shape:
        raise ValueError('geometries / exponent arrays of bra and ket differ in shape')
    out = numpy.zeros((n, n))
    if n == 0:
        return out
    eng = get_engine()
    _lib.check(eng.lib.okb_aooverlap(eng.ctx, _lib.dptr(geo_spec_a), _lib.dptr(geo_spec_b), geo_spec_a.shape[0],
                                     _lib.iptr(lxlylz_a), _lib.iptr(lxlylz_b), n, _lib.iptr(assign), _lib.dptr(ao_coeffs),
                                     _lib.iptr(pnum_list), _lib.iptr(atom_indices), len(assign), int(max(drv, 0)),
                                     int(is_normalized), _lib.dptr(out)))
    return out


def mooverlapmatrix(mo_a, mo_b, aoom, i_start, i_end):
    """moom[i - i_start, j] = sum_kl mo_a[i, k] mo_b[j, l] aoom[k, l] (cy_overlap.pyx:177-205) as (mo_a S) mo_b^T"""
    f = numpy.float64
    mo_a, mo_b, aoom = _typed(mo_a, f, 2, 'mo_a'), _typed(mo_b, f, 2, 'mo_b'), _typed(aoom, f, 2, 'aoom')
    a = numpy.ascontiguousarray(mo_a[i_start:i_end])
    if a.shape[0] == 0 or mo_b.shape[0] == 0:
        return numpy.zeros((a.shape[0], mo_b.shape[0]))
    eng = get_engine()
    half = eng.ci_td(a, aoom, out=numpy.empty((a.shape[0], aoom.shape[1])))                  # mo_a S
    return eng.ci_td(half, numpy.ascontiguousarray(mo_b.T), out=numpy.empty((a.shape[0], mo_b.shape[0])))


def mooverlap(mo_a, mo_b, aoom):
    f = numpy.float64
    mo_a, mo_b = _typed(mo_a, f, 1, 'mo_a'), _typed(mo_b, f, 1, 'mo_b')
    return float(mooverlapmatrix(mo_a[numpy.newaxis], mo_b[numpy.newaxis], aoom, 0, 1)[0, 0])
