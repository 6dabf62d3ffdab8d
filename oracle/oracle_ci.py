"""CPU ORACLE for the detCI grid contractions (orbkit/detci/ci_core.py:85-267, cy_ci.pyx:70-240).

TEST INFRASTRUCTURE ONLY (see oracle.py): imported by tests/ and bench legs as the checker,
never by orbkit_b200.

    rho(zero, sing, molist)                  sum_k c_k mo[a_k] mo[b_k]              (ci_core.rho)
    jab(zero, sing, molist, molistdrv)       -1/2 sum_k c_k (mo[a] d mo[b] - mo[b] d mo[a])   (ci_core.jab)
    a_nabla_b(zero, sing, molist, molistdrv) sum_k c_k mo[a] d mo[b]                (ci_core.a_nabla_b)

`zero = [[coeffs per determinant], [orbital indices per determinant]]`,
`sing = [[coeffs], [[a, b], ...]]` exactly as detci.occ_check.compare returns them.

Backends: "port" = libokoracle.so (okor_ci_*, our C restatement in the reference's operation
order), "ref" = the reference's own cy_ci module compiled into oracle/_ref (pinned bit for bit
against each other and against the golden refdata_h3+.npz in tests/test_oracle_ci.py).
"""
import ctypes
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
_ip = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))
_lib = None


def _port():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(os.path.join(_HERE, 'libokoracle.so'))
        for name in ('okor_ci_rho', 'okor_ci_jab', 'okor_ci_a_nabla_b', 'okor_ci_rho_full', 'okor_ci_j_full',
                     'okor_ci_jab_full'):
            getattr(_lib, name).restype = None
    return _lib


def have_ref():
    import glob
    return bool(glob.glob(os.path.join(_HERE, '_ref', 'cy_ci*.so')))


def _ref():
    d = os.path.join(_HERE, '_ref')
    if d not in sys.path:
        sys.path.insert(0, d)
    import cy_ci
    return cy_ci


def flatten(zero, sing):
    """(zero, sing) lists -> flat arrays (zc, zi, sc, sa, sb) in list order."""
    zc = [c for cs in zero[0] for c in cs]
    zi = [i for idx in zero[1] for i in idx]
    if len(zc) != len(zi):
        raise ValueError('zero: coefficient and index lists differ in length')
    sc = list(sing[0])
    sa = [p[0] for p in sing[1]]
    sb = [p[1] for p in sing[1]]
    if not (len(sc) == len(sa) == len(sb)):
        raise ValueError('sing: coefficient and index lists differ in length')
    f = lambda v: np.ascontiguousarray(v, dtype=np.float64)
    i = lambda v: np.ascontiguousarray(v, dtype=np.intc)
    return f(zc), i(zi), f(sc), i(sa), i(sb)


def _prep(molist, molistdrv=None):
    molist = np.require(molist, dtype=np.float64, requirements='CA')
    shape = molist.shape
    mo = molist.reshape(shape[0], -1)
    if molistdrv is None:
        return mo, None, shape
    drv = np.require(molistdrv, dtype=np.float64, requirements='CA').reshape(3, shape[0], -1)
    return mo, drv, shape


def slices(n, slice_length):
    """the slice driver of ci_core.rho / jab / a_nabla_b (ci_core.py:123-125, 184-186, 248-250):
        slice_length = min(N, slice_length); ij = arange(0, N+1, abs(int(slice_length)))
    NOTE (reference quirk, pinned by refdata_h3+.npz): the points behind the last full slice,
    N - (N // slice_length) * slice_length of them, are never visited and stay 0."""
    sl = abs(int(min(n, slice_length)))
    ij = np.arange(0, n + 1, sl, dtype=np.intc)
    return list(zip(ij[:-1], ij[1:]))


def rho(zero, sing, molist, slice_length=1e4, kind='port'):
    mo, _, shape = _prep(molist)
    n = mo.shape[1]
    data = np.zeros(n)
    if kind != 'ref':
        zc, zi, sc, sa, sb = flatten(zero, sing)
    for i, j in slices(n, slice_length):
        if kind == 'ref':
            data[i:j] = _ref().get_rho(int(i), int(j), zero, sing, mo)
        else:
            out = np.zeros(j - i)
            _port().okor_ci_rho(_dp(out), ctypes.c_long(i), ctypes.c_long(j), ctypes.c_long(n), _dp(mo),
                                ctypes.c_long(len(zc)), _dp(zc), _ip(zi), ctypes.c_long(len(sc)), _dp(sc), _ip(sa),
                                _ip(sb))
            data[i:j] = out
    return data.reshape(shape[1:])


def _vec(fn_ref, fn_port, zero, sing, molist, molistdrv, slice_length, kind):
    mo, drv, shape = _prep(molist, molistdrv)
    n = mo.shape[1]
    data = np.zeros((3, n))
    if kind != 'ref':
        _, _, sc, sa, sb = flatten(zero, sing)
    for i, j in slices(n, slice_length):
        if kind == 'ref':
            data[:, i:j] = getattr(_ref(), fn_ref)(int(i), int(j), zero, sing, mo, drv)
        else:
            out = np.zeros((3, j - i))
            getattr(_port(), fn_port)(_dp(out), ctypes.c_long(i), ctypes.c_long(j), ctypes.c_long(n),
                                      ctypes.c_long(mo.shape[0]), _dp(mo), _dp(drv), ctypes.c_long(len(sc)),
                                      _dp(sc), _ip(sa), _ip(sb))
            data[:, i:j] = out
    return data.reshape((3,) + shape[1:])


def jab(zero, sing, molist, molistdrv, slice_length=1e4, kind='port'):
    return _vec('get_jab', 'okor_ci_jab', zero, sing, molist, molistdrv, slice_length, kind)


def a_nabla_b(zero, sing, molist, molistdrv, slice_length=1e4, kind='port'):
    return _vec('get_a_nabla_b', 'okor_ci_a_nabla_b', zero, sing, molist, molistdrv, slice_length, kind)


# ---- time-dependent contractions (cy_ci.pyx:101-151, 186-202) -----------------------------------------------
def get_rho_full(ReS, rho, kind='port'):
    """tdrho[t, r] = sum_{n <= m} (1 or 2) ReS[t, m, n] rho[count, r]   (cy_ci.get_rho_full)"""
    ReS = np.require(ReS, dtype=np.float64, requirements='CA')
    rho = np.require(rho, dtype=np.float64, requirements='CA')
    if kind == 'ref':
        return _ref().get_rho_full(ReS, rho)
    nt, ns, npts = ReS.shape[0], ReS.shape[1], rho.shape[1]
    out = np.zeros((nt, npts))
    _port().okor_ci_rho_full(_dp(out), _dp(ReS), _dp(rho), ctypes.c_long(nt), ctypes.c_long(ns), ctypes.c_long(npts))
    return out


def get_j_full(ImS, j, kind='port'):
    """tdj[t, d, r] = - sum_{n < m} 2 ImS[t, n, m] j[count, d, r]   (cy_ci.get_j_full)"""
    ImS = np.require(ImS, dtype=np.float64, requirements='CA')
    j = np.require(j, dtype=np.float64, requirements='CA')
    if kind == 'ref':
        return _ref().get_j_full(ImS, j)
    nt, ns, npts = ImS.shape[0], ImS.shape[1], j.shape[2]
    out = np.zeros((nt, 3, npts))
    _port().okor_ci_j_full(_dp(out), _dp(ImS), _dp(j), ctypes.c_long(nt), ctypes.c_long(ns), ctypes.c_long(npts))
    return out


def get_jab_full(ImS, chi_n, nabla_chi_n, mu, kind='port'):
    """j[c, r] = sum_n sum_{m < n} ImS[n, m] / mu (chi[n] d_c chi[m] - chi[m] d_c chi[n])   (cy_ci.get_jab_full)"""
    ImS = np.require(ImS, dtype=np.float64, requirements='CA')
    chi = np.require(chi_n, dtype=np.float64, requirements='CA')
    dchi = np.require(nabla_chi_n, dtype=np.float64, requirements='CA')
    if kind == 'ref':
        return _ref().get_jab_full(ImS, chi, dchi, float(mu))
    nb, npts, nc = ImS.shape[0], chi.shape[1], dchi.shape[0]
    out = np.zeros((nc, npts))
    _port().okor_ci_jab_full(_dp(out), _dp(ImS), _dp(chi), _dp(dchi), ctypes.c_double(mu), ctypes.c_long(nb),
                             ctypes.c_long(nc), ctypes.c_long(npts))
    return out


# ---- core.calc_mo_matrix / extras.calc_jmo ------------------------------------------------------------------
def calc_mo_matrix(qc_a, x, y, z, is_vector=False, qc_b=None, drv=None, kind='port'):
    """mo_matrix[d, n, m] = mo_bra[n] * d_drv[d] mo_ket[m] on the grid (core.py:841-941).

    drv None -> ket = the MO values (one set); a string -> that one derivative; a list -> one set per entry
    (core.py:881-890).  MOs come from the oracle's rho_compute(calc_mo=True) as in core.py:894-918.
    NOTE: the reference's two-QCinfo branch raises TypeError for every input (`drv[ibra]` with a list index,
    core.py:906); for qc_b the evident intent is restated here: bra = MO values of qc_a, ket = the requested sets
    of qc_b.  That branch is therefore NOT pinned by any reference output."""
    import oracle
    if drv is None:
        dl, iket = [None], [0]
    elif not isinstance(drv, list):
        dl, iket = [None, drv], [1]
    else:
        dl, iket = [None] + drv, list(range(1, len(drv) + 1))
    if qc_b is None or qc_b is qc_a:
        mo = oracle.rho_compute(qc_a, x, y, z, is_vector=is_vector, calc_mo=True, drv=dl, kind=kind)
        mo_bra, mo_ket = mo[[0]], mo[iket]
    else:
        mo_bra = oracle.rho_compute(qc_a, x, y, z, is_vector=is_vector, calc_mo=True, drv=[None], kind=kind)
        mo_ket = oracle.rho_compute(qc_b, x, y, z, is_vector=is_vector, calc_mo=True, drv=[dl[i] for i in iket],
                                    kind=kind)
    nmo_a, nmo_b = mo_bra.shape[1], mo_ket.shape[1]
    out = np.zeros((mo_ket.shape[0], nmo_a) + mo_ket.shape[1:])
    for n in range(nmo_a):                      # core.py:939-941
        for m in range(nmo_b):
            out[:, n, m] = mo_bra[:, n] * mo_ket[:, m]
    return out


def jmo_from_matrix(mo_matrix, indices):
    """jmo[:, n] = -0.5 * (mo_matrix[:, i, j] - mo_matrix[:, j, i])   (extras.py:480-483)"""
    jmo = np.zeros((mo_matrix.shape[0], len(indices)) + mo_matrix.shape[3:])
    for n, (i, j) in enumerate(indices):
        jmo[:, n] = - 0.5 * (mo_matrix[:, i, j] - mo_matrix[:, j, i])
    return jmo


def calc_jmo(qc, ij, x, y, z, is_vector=False, drv=['x', 'y', 'z'], kind='port', select=None):
    """extras.calc_jmo (extras.py:441-493): the MOs named in `ij` are selected (numpy.unique), their
    mo_matrix is formed and antisymmetrised per pair.  `select(qc, u)` returns the QCinfo restricted to the
    MO indices u (defaults to qc.copy() + mo_spec[u], as the reference does)."""
    ij = np.asarray(ij)
    if ij.ndim == 1 and len(ij) == 2:
        ij = ij.reshape((1, 2))
    assert ij.ndim == 2 and ij.shape[1] == 2
    u, indices = np.unique(ij, return_inverse=True)
    indices = indices.reshape((-1, 2))
    if select is None:
        qs = qc.copy()
        qs.mo_spec = qc.mo_spec[u]
    else:
        qs = select(qc, u)
    return jmo_from_matrix(calc_mo_matrix(qs, x, y, z, is_vector=is_vector, drv=drv, kind=kind), indices)
