"""Drop-in for the time-dependent kernels of the reference's compiled module `orbkit.detci.cy_ci`
(cy_ci.pyx:101-122 get_rho_full, 126-151 get_j_full, 186-202 get_jab_full -- the reference's only OpenMP code): same
names, argument order, dtype / contiguity requirements (typed `mode="c"` buffers: anything else raises ValueError) and
freshly allocated float64 results.

    get_rho_full(ReS[nt, ns, ns], rho[npair, npts])            -> tdrho[nt, npts]
    get_j_full(ImS[nt, ns, ns], j[npair, 3, npts])             -> tdj[nt, 3, npts]
    get_jab_full(ImS[nb, nb], chi_n[nb, npts], nabla_chi_n[nc, nb, npts], mu) -> j[nc, npts]

`npair = ns (ns + 1) / 2` state pairs `count = (n, m >= n)`, n outer.  The first two are dense
(nt x npair) . (npair x npts) products and run on the FP64 tensor cores (okb_ci_td, csrc/okb_td.cuh) with the state
pairs in the reference's order: equal to the reference to rounding.  get_jab_full walks the pairs n > m with the pair
kernel of ci_core (okb_ci_jab_full): bit-identical to the reference.  (get_rho / get_jab / get_a_nabla_b, the slice
kernels of the same module, are reached through `ci_core.rho / jab / a_nabla_b`.)
"""
import numpy

from .. import options
from .._lib import OKB_FLAG_CI_FAST
from ..engine import get_engine


def _typed(a, ndim, name):
    """the acceptance rule of a Cython `np.ndarray[double, ndim=N, mode="c"] x not None` argument"""
    if a is None:
        raise TypeError("Argument '%s' must not be None" % name)
    if not isinstance(a, numpy.ndarray):
        raise TypeError("Argument '%s' has incorrect type (expected numpy.ndarray, got %s)" % (name, type(a).__name__))
    if a.dtype != numpy.float64:
        raise ValueError("Buffer dtype mismatch, expected 'double' but got '%s'" % a.dtype.name)
    if a.ndim != ndim:
        raise ValueError('Buffer has wrong number of dimensions (expected %d, got %d)' % (ndim, a.ndim))
    if not a.flags['C_CONTIGUOUS']:
        raise ValueError('ndarray is not C-contiguous')
    return a


def pair_weights_rho(ReS):
    """w[t, count] of get_rho_full: ReS[t, m, m] for m == n, 2 ReS[t, m, n] for m > n (cy_ci.pyx:115-119; the factor 2
    is exact, so w * rho is the reference's product)"""
    ns = ReS.shape[1]
    n, m = numpy.triu_indices(ns)                 # n outer, m >= n inner: the reference's count order
    w = ReS[:, m, n].copy()
    w[:, m != n] *= 2.0
    return numpy.ascontiguousarray(w)


def pair_weights_j(ImS):
    """w[t, count] of get_j_full: -2 ImS[t, n, m] for m > n, 0 for the diagonal pairs (cy_ci.pyx:141-146)"""
    ns = ImS.shape[1]
    n, m = numpy.triu_indices(ns)
    w = -2.0 * ImS[:, n, m]
    w[:, m == n] = 0.0
    return numpy.ascontiguousarray(w)


def get_rho_full(ReS, rho):
    ReS, rho = _typed(ReS, 3, 'ReS'), _typed(rho, 2, 'rho')
    nt, npts = ReS.shape[0], rho.shape[1]
    if nt == 0 or npts == 0:
        return numpy.zeros((nt, npts))
    w = pair_weights_rho(ReS)
    out = numpy.empty((nt, npts))
    get_engine().ci_td(w, rho[:w.shape[1]], out=out)
    return out


def get_j_full(ImS, j):
    ImS, j = _typed(ImS, 3, 'ImS'), _typed(j, 3, 'j')
    nt, npts = ImS.shape[0], j.shape[2]
    if nt == 0 or npts == 0:
        return numpy.zeros((nt, 3, npts))
    w = pair_weights_j(ImS)
    out = numpy.empty((nt, 3 * npts))
    get_engine().ci_td(w, j[:w.shape[1], :3].reshape(w.shape[1], 3 * npts), out=out)
    return out.reshape(nt, 3, npts)


def get_jab_full(ImS, chi_n, nabla_chi_n, mu):
    ImS, chi_n, nabla_chi_n = _typed(ImS, 2, 'ImS'), _typed(chi_n, 2, 'chi_n'), _typed(nabla_chi_n, 3, 'nabla_chi_n')
    nc, npts = nabla_chi_n.shape[0], chi_n.shape[1]
    if nc == 0 or npts == 0:
        return numpy.zeros((nc, npts))
    out = numpy.empty((nc, npts))
    # okb_ci_jab_full takes up to three components per call
    for c0 in range(0, nc, 3):
        get_engine().ci_jab_full(ImS, chi_n, nabla_chi_n[c0:c0 + 3], mu, out=out[c0:c0 + 3],
                                 flags=OKB_FLAG_CI_FAST if getattr(options, 'ci_fast', None) else 0)
    return out
