"""The overlap part of `orbkit.analytical_integrals` (analytical_integrals.py:37-326) that the readers need:
get_ao_overlap (+ the ORCA renormalisation 'N' and the spherical transformation), get_mo_overlap(_matrix) and
check_mo_norm.  Dipole moments, nuclear-attraction and kinetic-energy integrals are out of scope."""
import numpy

from . import cy_overlap
from .engine import cart2sph_dense
from .orbitals import AOClass
from .tools import require, validate_drv


def get_ao_overlap(coord_a, coord_b, ao_spec, lxlylz_b=None, drv=None):
    """overlap matrix <a|b> of the basis `ao_spec` placed at the geometries coord_a (bra) and coord_b (ket)
    (analytical_integrals.py:37-116)"""
    if not isinstance(ao_spec, AOClass):
        raise TypeError('ao_spec must be an instance of the AOClass')
    if isinstance(drv, list) or (isinstance(drv, str) and len(drv) > 1):
        return [get_ao_overlap(coord_a, coord_b, ao_spec, lxlylz_b=lxlylz_b, drv=d) for d in drv]
    lxlylz_a = ao_spec.get_lxlylz()
    if lxlylz_b is None:
        lxlylz_b = numpy.array(lxlylz_a, copy=True)
    else:
        try:
            lxlylz_b = numpy.array(lxlylz_b, dtype=numpy.intc)
        except ValueError:
            raise ValueError('The keyword argument `lxlylz` has to be convertable into a numpy integer array.')
        if lxlylz_a.shape != lxlylz_b.shape:
            raise ValueError('The exponents lxlylz for basis set a and basis set b have to have the same shape.')
    drv = validate_drv(drv)
    if drv > 3:
        raise ValueError('Only first derivatives are currently supported for analytical integrals.')
    aoom = cy_overlap.aooverlap(require(coord_a, dtype='f'), require(coord_b, dtype='f'),
                                require(lxlylz_a, dtype='i'), require(lxlylz_b, dtype='i'),
                                require(ao_spec.get_nlxlylz_per_cont(), dtype='i'),
                                require(ao_spec.get_prim_coeffs(), dtype='f'),
                                require(ao_spec.get_nprim_per_cont(), dtype='i'),
                                require(ao_spec.get_assign_cont_to_atoms(), dtype='i'), drv,
                                int(bool(ao_spec.get_normalized())))
    if 'N' in ao_spec[0]:
        n = numpy.asarray(ao_spec[0]['N'], dtype=float)
        for i in range(len(aoom)):
            aoom[i, :] *= n[i] * n[:, 0]
    if ao_spec.spherical:
        aoom = cartesian2spherical_aoom(aoom, ao_spec)
    return aoom


def cartesian2spherical_aoom(ao_overlap_matrix, ao_spec):
    """T S T^T with T the Cartesian -> real-spherical table of core.cartesian2spherical
    (analytical_integrals.py:118-183, a quadruple Python loop there)"""
    t = cart2sph_dense(ao_spec, ao_overlap_matrix.shape[0])
    return t.dot(ao_overlap_matrix).dot(t.T)


def get_mo_overlap(mo_a, mo_b, ao_overlap_matrix):
    shape = numpy.shape(ao_overlap_matrix)
    if isinstance(mo_a, dict):
        mo_a = numpy.array(mo_a['coeffs'])
    if isinstance(mo_b, dict):
        mo_b = numpy.array(mo_b['coeffs'])
    if numpy.ndim(mo_a) != 1 or numpy.ndim(mo_b) != 1:
        raise ValueError('The coefficients of mo_a and mo_b have to be one-dimensional vectors.')
    if len(mo_a) != shape[0] or len(mo_b) != shape[1]:
        raise ValueError('The atomic orbital overlap matrix has to have the shape (len(mo_a), len(mo_b)).')
    return cy_overlap.mooverlap(require(mo_a, dtype='f'), require(mo_b, dtype='f'), require(ao_overlap_matrix, dtype='f'))


def _coeffs(mo):
    if hasattr(mo, 'get_coeffs'):
        return mo.get_coeffs()
    if isinstance(mo, (list, tuple)) and len(mo) and isinstance(mo[0], dict):
        return numpy.array([m['coeffs'] for m in mo])
    return numpy.asarray(mo, dtype=float)


def get_mo_overlap_matrix(mo_a, mo_b, ao_overlap_matrix, numproc=1):
    """<mo_a[i]| S |mo_b[j]> for all pairs (analytical_integrals.py:225-288; `numproc` is accepted and ignored: the
    contraction runs on the device)"""
    a, b = require(_coeffs(mo_a), dtype='f'), require(_coeffs(mo_b), dtype='f')
    s = require(ao_overlap_matrix, dtype='f')
    if a.ndim != 2 or b.ndim != 2 or a.shape[1] != s.shape[0] or b.shape[1] != s.shape[1]:
        raise ValueError('The atomic orbital overlap matrix has to have the shape (NAO_a, NAO_b).')
    return cy_overlap.mooverlapmatrix(a, b, s, 0, len(a))


def check_mo_norm(qc):
    """|| C S C^T - 1 ||_F (analytical_integrals.py:318-326)"""
    aoom = get_ao_overlap(qc.geo_spec, qc.geo_spec, qc.ao_spec)
    moom = get_mo_overlap_matrix(qc.mo_spec, qc.mo_spec, aoom)
    return numpy.linalg.norm(moom - numpy.eye(len(moom)))
