"""Electron (transition) densities and flux densities of CI wavefunctions on a grid.

Same call signatures and return shapes as the reference (orbkit/detci/ci_core.py):

    rho(zero, sing, molist, slice_length=1e4, numproc=1)                 -> (N...)      ci_core.py:90-136
    jab(zero, sing, molist, molistdrv, slice_length=1e4, numproc=1)      -> (3, N...)   ci_core.py:145-203
    a_nabla_b(zero, sing, molist, molistdrv, slice_length=1e4, numproc=1)-> (3, N...)   ci_core.py:212-267

`zero = [[prefactors per determinant], [orbital indices per determinant]]` and
`sing = [[products of CI coefficients], [[a, b], ...]]` are what detci.occ_check.compare returns.
The per-slice Cython loops (cy_ci.get_rho / get_jab / get_a_nabla_b) and the multiprocessing
slice driver are replaced by one CUDA kernel (csrc/okb_ci.cuh) that walks the term list in list order
with the reference's expression order, so for identical `molist` the results are bit-identical.
`numproc` is ignored.  `slice_length` is honoured only for its one visible effect in the reference: the
slice bounds are `arange(0, N+1, int(min(N, slice_length)))` (ci_core.py:123-125), so the points behind
the last FULL slice are never visited and stay 0 (the reference's golden refdata_h3+.npz pins this:
2601 points, slice_length=1e2 -> last point 0).  The default 1e4 covers all points whenever N <= 1e4 or
N is a multiple of it; pass `slice_length=N` (or use the `*_from_qc` functions) to evaluate everything.

B200-native additions: `rho_from_qc`, `jab_from_qc`, `a_nabla_b_from_qc` evaluate the MOs on the device
(fused AO -> MO kernel) and contract them there; `pair_products` returns mo[a]*mo[b] per pair.
With `options.ci_merge_terms = True` duplicate orbital pairs are merged on the host before the launch
(fewer terms, summation order differs from the reference at the 1e-16 level).
"""
import numpy

from .. import grid, options
from .._lib import OKB_CI_RHO, OKB_CI_JAB, OKB_CI_A_NABLA_B, OKB_CI_PAIRS, OKB_FLAG_CI_FAST
from ..engine import get_engine
from ..tools import require, validate_drv


def flatten_terms(zero, sing, with_zero=True):
    """(zero, sing) -> flat (coef, ia, ib) in the order the reference loops visit them."""
    if (not with_zero or len(zero[0]) == 0) and isinstance(sing[0], numpy.ndarray) and isinstance(sing[1], numpy.ndarray):
        # arrays in: no per-element list traffic (the Python lists of a 20 000-term expansion cost ~3 ms per call)
        pairs = numpy.asarray(sing[1], dtype=numpy.intc).reshape((-1, 2))
        coef = numpy.asarray(sing[0], dtype=numpy.float64).reshape(-1)
        if len(coef) != len(pairs):
            raise ValueError('sing: coefficient and index lists differ in length')
        return coef, numpy.ascontiguousarray(pairs[:, 0]), numpy.ascontiguousarray(pairs[:, 1])
    coef, ia, ib = [], [], []
    if with_zero:
        if len(zero[0]) != len(zero[1]):
            raise ValueError('zero: coefficient and index lists differ in length')
        for cs, idx in zip(zero[0], zero[1]):
            if len(cs) != len(idx):
                raise ValueError('zero: coefficient and index lists differ in length')
            coef.extend(cs)
            ia.extend(idx)
            ib.extend(idx)
    if len(sing[0]) != len(sing[1]):
        raise ValueError('sing: coefficient and index lists differ in length')
    coef.extend(sing[0])
    ia.extend(p[0] for p in sing[1])
    ib.extend(p[1] for p in sing[1])
    return (numpy.asarray(coef, dtype=numpy.float64), numpy.asarray(ia, dtype=numpy.intc),
            numpy.asarray(ib, dtype=numpy.intc))


def merge_terms(terms, n_mo, symmetric):
    """sum the coefficients of identical orbital pairs ((a,b) == (b,a) when `symmetric`)."""
    coef, ia, ib = terms
    if len(coef) == 0:
        return terms
    a, b = ia.astype(numpy.int64), ib.astype(numpy.int64)
    if symmetric:
        a, b = numpy.minimum(a, b), numpy.maximum(a, b)
    key = a * n_mo + b
    uniq, inv = numpy.unique(key, return_inverse=True)
    csum = numpy.zeros(len(uniq))
    numpy.add.at(csum, inv, coef)
    keep = csum != 0.0
    return (csum[keep], (uniq[keep] // n_mo).astype(numpy.intc), (uniq[keep] % n_mo).astype(numpy.intc))


def _terms(zero, sing, n_mo, mode):
    terms = flatten_terms(zero, sing, with_zero=(mode == OKB_CI_RHO))
    if len(terms[0]) and (terms[1].min() < 0 or max(terms[1].max(), terms[2].max()) >= n_mo or terms[2].min() < 0):
        raise ValueError('orbital index outside 0..%d' % (n_mo - 1))
    if getattr(options, 'ci_merge_terms', False) and mode in (OKB_CI_RHO,):
        terms = merge_terms(terms, n_mo, symmetric=True)
    elif getattr(options, 'ci_merge_terms', False) and mode == OKB_CI_A_NABLA_B:
        terms = merge_terms(terms, n_mo, symmetric=False)
    return terms


def _n_visited(n, slice_length):
    """points the reference's slice driver visits (ci_core.py:123-125)"""
    sl = abs(int(min(n, slice_length)))
    if sl == 0:
        raise ValueError('slice_length must not be zero')
    return (n // sl) * sl


def _given(mode, zero, sing, molist, molistdrv=None, slice_length=1e4):
    molist = require(molist, dtype='f')
    shape = molist.shape
    mo2 = molist.reshape((shape[0], -1))
    drv3 = None
    if molistdrv is not None:
        molistdrv = require(molistdrv, dtype='f')
        drv3 = molistdrv.reshape((3, shape[0], -1))
    ncomp = 1 if mode == OKB_CI_RHO else 3
    if mo2.shape[1] == 0:
        return numpy.zeros(((ncomp,) if ncomp > 1 else ()) + shape[1:])
    n_eval = _n_visited(mo2.shape[1], slice_length)
    if n_eval < mo2.shape[1]:
        # parity with the reference's slice driver (ci_core.py:123-125: arange(0, N+1, slice_length) never reaches the
        # points behind the last full slice); say so, the fused *_from_qc variants evaluate every point
        from ..display import display
        display('detci.ci_core: like the reference, the last %d of %d points (behind the last full slice of %d) are not '
                'evaluated and stay 0; pass a slice_length that divides the number of points' %
                (mo2.shape[1] - n_eval, mo2.shape[1], abs(int(min(mo2.shape[1], slice_length)))))
    out = get_engine().ci_contract(mode, _terms(zero, sing, shape[0], mode), mo2, drv3, n_eval=n_eval,
                                   flags=OKB_FLAG_CI_FAST if getattr(options, 'ci_fast', None) else 0)
    return out.reshape(shape[1:]) if ncomp == 1 else out.reshape((3,) + shape[1:])


def rho(zero, sing, molist, slice_length=1e4, numproc=1):
    """Electron (transition) density: sum_k c_k mo[a_k] mo[b_k]  (ci_core.py:90-136)."""
    return _given(OKB_CI_RHO, zero, sing, molist, None, slice_length)


def jab(zero, sing, molist, molistdrv, slice_length=1e4, numproc=1):
    """Imaginary part of the electronic (transition) flux density (ci_core.py:145-203):
    -1/2 sum_k c_k (mo[a_k] grad mo[b_k] - mo[b_k] grad mo[a_k])."""
    return _given(OKB_CI_JAB, zero, sing, molist, molistdrv, slice_length)


def a_nabla_b(zero, sing, molist, molistdrv, slice_length=1e4, numproc=1):
    """sum_k c_k mo[a_k] grad mo[b_k]  (ci_core.py:212-267)."""
    return _given(OKB_CI_A_NABLA_B, zero, sing, molist, molistdrv, slice_length)


def pair_products(pairs, molist):
    """mo[a]*mo[b] for every (a, b) in `pairs`: shape (len(pairs),) + N (the products the reference forms
    one by one in core.calc_mo_matrix / extras.calc_jmo, core.py:925-941)."""
    molist = require(molist, dtype='f')
    shape = molist.shape
    pairs = numpy.asarray(pairs, dtype=numpy.intc).reshape((-1, 2))
    terms = (numpy.zeros(len(pairs)), numpy.ascontiguousarray(pairs[:, 0]), numpy.ascontiguousarray(pairs[:, 1]))
    if len(pairs) and (pairs.min() < 0 or pairs.max() >= shape[0]):
        raise ValueError('orbital index outside 0..%d' % (shape[0] - 1))
    if len(pairs) == 0 or molist.size == 0:
        return numpy.zeros((len(pairs),) + shape[1:])
    out = get_engine().ci_contract(OKB_CI_PAIRS, terms, molist.reshape((shape[0], -1)))
    return out.reshape((len(pairs),) + shape[1:])


# ---- fused: MOs evaluated and contracted on the device -----------------------------------------------------
def _from_qc(mode, qc, zero, sing, drv, x, y, z, is_vector):
    from ..core import _resolve_grid, _grid_handle
    x, y, z, is_vector, N = _resolve_grid(x, y, z, is_vector)
    eng = get_engine()
    basis = eng.basis(require(qc.geo_spec, dtype='f'), qc.ao_spec)
    terms = _terms(zero, sing, len(qc.mo_spec), mode)
    # only the orbitals the term lists refer to are evaluated (a CI expansion over a few dozen active orbitals of a
    # few hundred MOs: the AO -> MO contraction, the dominant cost, shrinks by that factor); the order of the terms
    # and hence of the sums is untouched
    act = numpy.unique(numpy.concatenate((terms[1], terms[2])))
    subset = 0 < len(act) < len(qc.mo_spec)
    if subset:
        terms = (terms[0], numpy.searchsorted(act, terms[1]).astype(numpy.intc),
                 numpy.searchsorted(act, terms[2]).astype(numpy.intc))
    ncomp = 1 if mode == OKB_CI_RHO else 3
    lead = () if ncomp == 1 else (3,)
    if int(numpy.prod(N)) == 0:
        return numpy.zeros(lead + N)
    codes = [validate_drv(d) for d in drv]
    if len(codes) != 3 or any(c == 0 for c in codes):
        raise ValueError('`drv` must name three derivatives, e.g. ["x","y","z"] or ["xx","yy","zz"]')
    g = _grid_handle(eng, x, y, z, is_vector)
    fast = getattr(options, 'ci_fast', None)
    fast = fast is None or bool(fast)
    if mode == OKB_CI_RHO and fast and 0 < len(act) <= NATURAL_MAX:
        return _rho_natural(eng, basis, qc, act, terms, g).reshape(N)
    if subset:
        mo = eng.mos(basis, numpy.ascontiguousarray(require(qc.mo_spec.get_coeffs(), dtype='f')[act]),
                     numpy.ascontiguousarray(require(qc.mo_spec.get_occ(), dtype='f')[act]))
    else:                                  # every orbital is referred to (act = 0..n_mo-1) or there are no terms
        mo = eng.mos_of(basis, qc.mo_spec)
    out = eng.eval_ci(mode, terms, mo, g, drv_codes=codes, flags=OKB_FLAG_CI_FAST if fast else 0)
    return out.reshape(lead + N)


NATURAL_MAX = 256    #: largest active space whose pair matrix is diagonalised on the host (eigh of 256^2: ~10 ms)


def _rho_natural(eng, basis, qc, act, terms, g):
    """rho = sum_ab D_ab phi_a phi_b depends on the symmetric part of D only: with D_s = U diag(lam) U^T it is
    sum_k lam_k psi_k^2 over the natural (transition) orbitals psi = U^T phi -- ONE launch of the fused
    AO -> MO -> rho kernel with n_act orbitals and signed weights, no MO slab and no second kernel.  `terms` index
    into `act`.  Eigenvalues below 1e-15 of the largest are dropped (transition matrices are of low rank)."""
    lam, u = natural_orbitals(terms, len(act))
    if len(lam) == 0:
        return numpy.zeros(g.npts)
    coeffs = numpy.ascontiguousarray(u.T @ require(qc.mo_spec.get_coeffs(), dtype='f')[act])
    mo = eng.mos(basis, coeffs, numpy.ascontiguousarray(lam))
    return eng.eval_rho(mo, g, [])[0]


def natural_orbitals(terms, n):
    """(lam, U) with sum_t c_t phi_{a_t} phi_{b_t} = sum_k lam_k (sum_a U[a, k] phi_a)^2 for orbital indices below n:
    eigen-decomposition of the symmetric part of the pair matrix D[a, b] = sum of the c_t with (a_t, b_t) = (a, b);
    eigenvalues below 1e-15 of the largest are dropped (possibly all of them: an antisymmetric D gives rho = 0)."""
    d = numpy.bincount(numpy.asarray(terms[1]).astype(numpy.int64) * n + numpy.asarray(terms[2]),
                       weights=numpy.asarray(terms[0], dtype=float), minlength=n * n).reshape((n, n))
    lam, u = numpy.linalg.eigh(0.5 * (d + d.T))
    keep = numpy.abs(lam) > 1e-15 * max(numpy.abs(lam).max() if len(lam) else 0.0, 1e-300)
    return lam[keep], u[:, keep]


def rho_from_qc(qc, zero, sing, x=None, y=None, z=None, is_vector=None):
    """rho(zero, sing, rho_compute(qc, calc_mo=True)) without materialising the MOs on the host."""
    return _from_qc(OKB_CI_RHO, qc, zero, sing, ['x', 'y', 'z'], x, y, z, is_vector)


def jab_from_qc(qc, zero, sing, drv=('x', 'y', 'z'), x=None, y=None, z=None, is_vector=None):
    """jab(zero, sing, mo, d mo) with the MOs and their derivatives `drv` evaluated on the device."""
    return _from_qc(OKB_CI_JAB, qc, zero, sing, list(drv), x, y, z, is_vector)


def a_nabla_b_from_qc(qc, zero, sing, drv=('x', 'y', 'z'), x=None, y=None, z=None, is_vector=None):
    return _from_qc(OKB_CI_A_NABLA_B, qc, zero, sing, list(drv), x, y, z, is_vector)
