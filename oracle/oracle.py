"""CPU ORACLE for the ORBKIT grid-based hot path -- numpy/ctypes host side.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this module,
and only as the checker or the timed CPU baseline.  ``orbkit_b200`` never imports it.

Parity is PINNED: the arithmetic runs either in ``oracle/_ref`` (the reference's own
C/Cython objects compiled from /root/reference, backend ``"ref"``) or in
``libokoracle.so`` (our C restatement, backend ``"port"``, bit-identical to ``_ref`` --
tests/test_oracle.py), and both reproduce the reference's golden
``refdata_rho_compute.npz`` and the Gaussian cubegen cube files (tests/golden/).

The functions restate, in the reference's operation order,
    core.ao_creator              orbkit/core.py:38-105
    core.cartesian2spherical     orbkit/core.py:135-176   (table: orbkit/tools.py:155-191)
    core.mo_creator              orbkit/core.py:107-132
    core.slice_rho               orbkit/core.py:179-308
    core.rho_compute (assembly)  orbkit/core.py:314-605
on top of the getter interface of AOClass/MOClass (orbkit/orbitals.py:336-409,
760-832); any object exposing those getters works (the reference's classes, the
product's mirror classes, or tests.fixtures.FlatQC).
"""
import ctypes
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_int_p = ctypes.POINTER(ctypes.c_int)


def _dp(a):
    return a.ctypes.data_as(_c_double_p)


def _ip(a):
    return a.ctypes.data_as(_c_int_p)


def _f64(a):
    return np.require(a, dtype=np.float64, requirements='CA')


def _i32(a):
    return np.require(a, dtype=np.intc, requirements='CA')


# --------------------------------------------------------------------------
# backends
# --------------------------------------------------------------------------
class _PortBackend:
    """libokoracle.so -- our C restatement (oracle/okoracle.c)."""
    kind = 'port'

    def __init__(self):
        path = os.path.join(_HERE, 'libokoracle.so')
        if not os.path.exists(path):
            raise OSError('oracle/libokoracle.so missing: run `make -C oracle`')
        L = ctypes.CDLL(path)
        L.okor_aocreator.restype = None
        L.okor_aocreator.argtypes = [_c_double_p, _c_int_p, _c_int_p, _c_double_p, _c_int_p,
                                     _c_double_p, _c_int_p, ctypes.c_int,
                                     _c_double_p, _c_double_p, _c_double_p, ctypes.c_long,
                                     ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.okor_mocreator.restype = None
        L.okor_mocreator.argtypes = [_c_double_p, _c_double_p, _c_double_p,
                                     ctypes.c_int, ctypes.c_int, ctypes.c_long]
        L.okor_grid2vector.restype = None
        L.okor_grid2vector.argtypes = [_c_double_p] * 4 + [ctypes.c_long] * 3
        L.okor_ao_norm.restype = ctypes.c_double
        L.okor_ao_norm.argtypes = [ctypes.c_int] * 3 + [ctypes.c_double, ctypes.c_int]
        L.okor_poly.restype = ctypes.c_double
        L.okor_poly.argtypes = [ctypes.c_double] * 3 + [ctypes.c_int] * 3 + \
                               [ctypes.c_double, ctypes.c_int, ctypes.c_int]
        self.L = L

    def aocreator(self, lxlylz, assign, coeffs, pnum, geo, atom_idx, x, y, z, drv,
                  normalized, exact_mixed=0):
        out = np.zeros((lxlylz.shape[0], x.shape[0]))
        self.L.okor_aocreator(_dp(out), _ip(lxlylz), _ip(assign), _dp(coeffs), _ip(pnum),
                              _dp(geo), _ip(atom_idx), len(assign), _dp(x), _dp(y), _dp(z),
                              x.shape[0], int(drv), int(normalized), int(exact_mixed))
        return out

    def mocreator(self, ao, C):
        mo = np.zeros((C.shape[0], ao.shape[1]))
        self.L.okor_mocreator(_dp(mo), _dp(ao), _dp(C), C.shape[0], C.shape[1], ao.shape[1])
        return mo

    def grid2vector(self, x, y, z):
        out = np.zeros((3, len(x) * len(y) * len(z)))
        self.L.okor_grid2vector(_dp(out), _dp(x), _dp(y), _dp(z), len(x), len(y), len(z))
        return out

    def aonorm(self, lx, ly, lz, alpha, normalized):
        return self.L.okor_ao_norm(lx, ly, lz, alpha, normalized)

    def aoxyz(self, X, Y, Z, lx, ly, lz, alpha, drv, exact_mixed=0):
        return self.L.okor_poly(X, Y, Z, lx, ly, lz, alpha, drv, exact_mixed)


class _RefBackend:
    """oracle/_ref -- the reference's own cy_core/cy_grid objects (built by `make -C oracle ref`)."""
    kind = 'reference'

    def __init__(self):
        d = os.path.join(_HERE, '_ref')
        if d not in sys.path:
            sys.path.insert(0, d)
        import cy_core  # noqa: the reference's extension module, compiled unmodified
        import cy_grid
        self.cy_core, self.cy_grid = cy_core, cy_grid

    def aocreator(self, lxlylz, assign, coeffs, pnum, geo, atom_idx, x, y, z, drv,
                  normalized, exact_mixed=0):
        if exact_mixed:
            raise ValueError('the reference has no exact mixed-derivative mode')
        return self.cy_core.aocreator(lxlylz, assign, coeffs, pnum, geo, atom_idx, x, y, z,
                                      int(drv), int(normalized))

    def mocreator(self, ao, C):
        return self.cy_core.mocreator(ao, C)

    def grid2vector(self, x, y, z):
        return self.cy_grid.grid2vector(x, y, z)

    def aonorm(self, lx, ly, lz, alpha, normalized):
        return self.cy_core.aonorm(lx, ly, lz, alpha, normalized)

    def aoxyz(self, X, Y, Z, lx, ly, lz, alpha, drv, exact_mixed=0):
        return self.cy_core.aoxyz(X, Y, Z, lx, ly, lz, alpha, drv)


_BACKENDS = {}


def backend(kind='port'):
    if kind not in _BACKENDS:
        _BACKENDS[kind] = _PortBackend() if kind == 'port' else _RefBackend()
    return _BACKENDS[kind]


def have_ref():
    try:
        backend('ref')
        return True
    except Exception:
        return False


# --------------------------------------------------------------------------
# cart -> real-spherical table (restates orbkit/tools.py:155-191, INCLUDING the
# reference's two wrong g rows: (4,-1) lists (0,3,1) twice, (4,0) is mis-scaled)
# entry: (l, m) -> ([(lx,ly,lz), ...], [coef, ...], factor)
# --------------------------------------------------------------------------
_s = np.sqrt
CART2SPH = {
    (0, 0): ([(0, 0, 0)], [1.], 1.),
    (1, -1): ([(0, 1, 0)], [1.], 1.),
    (1, 0): ([(0, 0, 1)], [1.], 1.),
    (1, 1): ([(1, 0, 0)], [1.], 1.),
    (2, -2): ([(1, 1, 0)], [1.], 1.),
    (2, -1): ([(0, 1, 1)], [1.], 1.),
    (2, 0): ([(0, 0, 2), (2, 0, 0), (0, 2, 0)], [1., -1 / 2., -1 / 2.], 1.),
    (2, 1): ([(1, 0, 1)], [1.], 1.),
    (2, 2): ([(2, 0, 0), (0, 2, 0)], [1., -1.], _s(3) / 2.),
    (3, -3): ([(0, 3, 0), (2, 1, 0)], [-_s(5), 3.], 1 / (2. * _s(2))),
    (3, -2): ([(1, 1, 1)], [1.], 1.),
    (3, -1): ([(0, 1, 2), (0, 3, 0), (2, 1, 0)],
              [_s(3 / 5.), -_s(3) / 4., -_s(3) / (4. * _s(5))], _s(2)),
    (3, 0): ([(0, 0, 3), (2, 0, 1), (0, 2, 1)],
             [1., -3 / (2 * _s(5)), -3 / (2 * _s(5))], 1.),
    (3, 1): ([(1, 0, 2), (3, 0, 0), (1, 2, 0)],
             [_s(3 / 5.), -_s(3) / 4., -_s(3) / (4. * _s(5))], _s(2)),
    (3, 2): ([(2, 0, 1), (0, 2, 1)], [1., -1.], _s(3) / 2.),
    (3, 3): ([(3, 0, 0), (1, 2, 0)], [_s(5), -3.], 1 / (2. * _s(2))),
    (4, -4): ([(3, 1, 0), (1, 3, 0)], [1., -1.], _s(2) * _s(5 / 8.)),
    (4, -3): ([(0, 3, 1), (2, 1, 1)], [-_s(5) / 4., 3 / 4.], _s(2)),
    (4, -2): ([(1, 1, 2), (3, 1, 0), (1, 3, 0)],
              [3 / _s(14), -_s(5) / (2 * _s(14)), -_s(5) / (2 * _s(14))], _s(2)),
    (4, -1): ([(0, 3, 1), (0, 3, 1), (2, 1, 1)],
              [_s(5 / 7.), -3 * _s(5) / (4. * _s(7)), -3 / (4. * _s(7))], _s(2)),
    (4, 0): ([(0, 0, 4), (4, 0, 0), (0, 4, 0), (2, 0, 2), (0, 2, 2), (2, 2, 0)],
             [1., 3 / 8., 3 / 8., -3 * _s(3) / _s(35), -3 * _s(3) / _s(35), -1 / 4.], _s(2)),
    (4, 1): ([(1, 0, 3), (3, 0, 1), (1, 2, 1)],
             [_s(5 / 7.), -3 * _s(5) / (4. * _s(7)), -3 / (4. * _s(7))], _s(2)),
    (4, 2): ([(2, 0, 2), (0, 2, 2), (4, 0, 0), (0, 4, 0)],
             [3 * _s(3) / (2. * _s(14)), -3 * _s(3) / (2. * _s(14)),
              -_s(5) / (4. * _s(2)), _s(5) / (4. * _s(2))], _s(2)),
    (4, 3): ([(3, 0, 1), (1, 2, 1)], [_s(5) / 4., -3 / 4.], _s(2)),
    (4, 4): ([(4, 0, 0), (0, 4, 0), (2, 2, 0)],
             [_s(35) / (8. * _s(2)), _s(35) / (8. * _s(2)), -3 * _s(3) / (4. * _s(2))], _s(2)),
}

_DRV_CODES = {None: 0, 'None': 0, '': 0, 'x': 1, 'y': 2, 'z': 3,
              'xx': 4, 'x2': 4, 'yy': 5, 'y2': 5, 'zz': 6, 'z2': 6,
              'xy': 7, 'yx': 7, 'xz': 8, 'zx': 8, 'yz': 9, 'zy': 9}


def validate_drv(drv):
    """orbkit/tools.py:225-239 (ints 0..9 pass through)."""
    if isinstance(drv, str) or drv is None:
        if drv not in _DRV_CODES:
            raise ValueError("The selection `drv=%s` is not valid!" % drv)
        return _DRV_CODES[drv]
    if isinstance(drv, (int, np.integer)) and 0 <= drv <= 9:
        return int(drv)
    raise ValueError("The selection `drv=%s` is not valid!" % drv)


# --------------------------------------------------------------------------
# operators
# --------------------------------------------------------------------------
def cartesian2spherical(ao_cart, ao_spec):
    """orbkit/core.py:135-176: row axpy in (spherical AO, term) order; the Cartesian row is
    looked up by exponent triple inside the owning contraction (stale `index0` kept on a miss)."""
    lxlylz = ao_spec.get_lxlylz()
    assign = ao_spec.get_assign_lxlylz_to_cont()
    rows_of = {}
    for i, j in enumerate(assign):
        rows_of.setdefault(int(j), []).append(i)
    lm = list(zip(ao_spec.get_assign_lm_to_cont(), ao_spec.get_lm()))
    out = np.zeros((len(lm),) + ao_cart.shape[1:])
    index0 = None
    for i0, (j0, k0) in enumerate(lm):
        exps, coefs, factor = CART2SPH[(int(k0[0]), int(k0[1]))]
        for c0 in range(len(exps)):
            for i, j in enumerate(rows_of[int(j0)]):
                if tuple(int(v) for v in lxlylz[j]) == exps[c0]:
                    index0 = i + rows_of[int(j0)][0]
            out[i0, :] += coefs[c0] * factor * ao_cart[index0, :]
    return out


def ao_creator(geo_spec, ao_spec, drv=None, x=None, y=None, z=None, is_vector=True,
               kind='port', exact_mixed=0):
    """orbkit/core.py:38-105 with explicit coordinates (no module-global grid)."""
    be = backend(kind)
    x, y, z = _f64(x), _f64(y), _f64(z)
    if not is_vector:
        shape = (len(x), len(y), len(z))
        x, y, z = be.grid2vector(x.copy(), y.copy(), z.copy())
    else:
        if len(x) != len(y) or len(x) != len(z):
            raise ValueError('Dimensions of x-, y-, and z- coordinate differ!')
        shape = (len(x),)
    ao = be.aocreator(_i32(ao_spec.get_lxlylz()), _i32(ao_spec.get_nlxlylz_per_cont()),
                      _f64(ao_spec.get_prim_coeffs()), _i32(ao_spec.get_nprim_per_cont()),
                      _f64(geo_spec), _i32(ao_spec.get_assign_cont_to_atoms()),
                      _f64(x), _f64(y), _f64(z), validate_drv(drv),
                      ao_spec.get_normalized(), exact_mixed)
    renorm = getattr(ao_spec, 'get_renorm', lambda: None)()
    if renorm is None and len(ao_spec) and isinstance(ao_spec[0], dict) and 'N' in ao_spec[0]:
        renorm = ao_spec[0]['N']
    if renorm is not None:
        ao *= renorm
    if ao_spec.spherical:
        ao = cartesian2spherical(ao, ao_spec)
    return ao.reshape((len(ao),) + shape)


def mo_creator(ao_list, mo_spec, kind='port'):
    """orbkit/core.py:107-132."""
    ao = _f64(ao_list)
    C = _f64(mo_spec.get_coeffs())
    mo = backend(kind).mocreator(ao.reshape(ao.shape[0], -1), C)
    return mo.reshape((C.shape[0],) + ao.shape[1:])


def _letters(d):
    """first-derivative codes of the product term of a two-letter code (core.py:287-296)."""
    if '2' in d or d[0] == d[1]:
        return d[0], d[0]
    return d[0], d[1]


def slice_rho(qc, x, y, z, drv=None, calc_mo=False, calc_ao=False, kind='port', exact_mixed=0):
    """orbkit/core.py:179-308 for one vector-grid slice."""
    geo, ao_spec, mo_spec = qc.geo_spec, qc.ao_spec, qc.mo_spec
    aoc = lambda d: ao_creator(geo, ao_spec, drv=d, x=x, y=y, z=z, is_vector=True,
                               kind=kind, exact_mixed=exact_mixed)
    moc = (lambda a: a) if calc_ao else (lambda a: mo_creator(a, mo_spec, kind=kind))
    if drv is not None and calc_mo:
        return np.array([moc(aoc(d)) for d in drv])
    mo = moc(aoc(None))
    if calc_mo:
        return np.array(mo)
    occ = mo_spec.get_occ()
    rho = np.zeros(len(x))
    mo_norm = np.zeros(len(mo))
    for i in range(len(mo)):
        mo_norm[i] = np.sum(np.square(mo[i]))
        rho += occ[i] * np.square(np.abs(mo[i]))
    if drv is None:
        return rho, mo_norm
    delta_rho = np.zeros((len(drv), len(x)))
    for n, d in enumerate(drv):
        dmo = mo_creator(aoc(d), mo_spec, kind=kind)
        if len(d) == 2:
            a, b = _letters(d)
            if a == b:
                d2 = np.array(mo_creator(aoc(a), mo_spec, kind=kind)) ** 2
            else:
                d2 = (np.array(mo_creator(aoc(a), mo_spec, kind=kind)) *
                      np.array(mo_creator(aoc(b), mo_spec, kind=kind)))
        for i in range(len(mo)):
            delta_rho[n] += occ[i] * 2 * dmo[i] * mo[i]
            if len(d) == 2:
                delta_rho[n] += occ[i] * 2 * d2[i]
    return rho, mo_norm, delta_rho


def rho_compute(qc, x, y, z, is_vector=False, calc_ao=False, calc_mo=False, drv=None,
                laplacian=False, numproc=1, slice_length=1e4, kind='port', exact_mixed=0,
                return_norm=False):
    """orbkit/core.py:314-605 with explicit coordinates; `numproc` host THREADS replace the
    reference's fork pool (the C kernels release the GIL; results do not depend on slicing)."""
    if calc_ao and calc_mo:
        raise ValueError('Choose either calc_ao=True or calc_mo=True')
    if calc_ao:
        calc_mo = True
    if laplacian:
        drv = ['xx', 'yy', 'zz']
    if drv is not None:
        try:
            drv = list(drv)
        except TypeError:
            drv = [drv]
    x, y, z = _f64(x), _f64(y), _f64(z)
    if not is_vector:
        shape = (len(x), len(y), len(z))
        x, y, z = backend(kind).grid2vector(x, y, z)
    else:
        shape = (len(x),)
    npts = len(x)
    numproc = max(1, int(numproc))
    if slice_length <= 0:
        slice_length = np.ceil(npts / float(numproc)) + 1
    slice_length = int(min(slice_length, npts)) or 1
    bounds = [(i, min(i + slice_length, npts)) for i in range(0, npts, slice_length)]

    def work(b):
        return slice_rho(qc, x[b[0]:b[1]], y[b[0]:b[1]], z[b[0]:b[1]], drv=drv, calc_mo=calc_mo,
                         calc_ao=calc_ao, kind=kind, exact_mixed=exact_mixed)

    if numproc > 1 and len(bounds) > 1:
        with ThreadPoolExecutor(numproc) as ex:
            results = list(ex.map(work, bounds))
    else:
        results = [work(b) for b in bounds]

    if calc_mo:
        out = np.concatenate(results, axis=-1)
        return out.reshape(out.shape[:-1] + shape)
    rho = np.concatenate([r[0] for r in results]).reshape(shape)
    mo_norm = sum(r[1] for r in results)
    if drv is None:
        return (rho, mo_norm) if return_norm else rho
    delta = np.concatenate([r[2] for r in results], axis=-1).reshape((len(drv),) + shape)
    ret = (rho, delta, delta.sum(axis=0)) if laplacian else (rho, delta)
    return ret + (mo_norm,) if return_norm else ret


def gross_atomic_density(atom_indices, qc, x, y, z, is_vector=True, drv=None, kind='port'):
    """orbkit/extras.py:306-385 for the atoms `atom_indices` (counting from ZERO, i.e. the `index` array
    of extras.atom2index), operation by operation:
        mo_info = sum_jj coeffs[ao_index[jj]] * ao[jj]            (plain += in AO order, :372-374)
        rho_atom += occ_num * mo_list[ii_mo] * mo_info            (:375)
    Returns (rho_atom, mo_atom) as lists per atom.

    AO -> atom assignment: the reference enumerates `core.l_deg(l=type)` CARTESIAN functions per
    shell (:365), which is only right for Cartesian bases; for spherical bases that enumeration runs
    past / mislabels the AO rows (reference bug, unpinned by any test).  This restatement follows the
    reference for Cartesian bases and assigns every spherical AO to the atom of its shell."""
    ao_list = ao_creator(qc.geo_spec, qc.ao_spec, drv=drv, x=x, y=y, z=z, is_vector=is_vector, kind=kind)
    mo_list = mo_creator(ao_list, qc.mo_spec, kind=kind)
    N = mo_list.shape[1:]
    cont_atom = np.asarray(qc.ao_spec.get_assign_cont_to_atoms(), dtype=int)
    if qc.ao_spec.spherical:
        ao_atom = cont_atom[np.asarray(qc.ao_spec.get_assign_lm_to_cont(), dtype=int)]
    else:
        ao_atom = np.repeat(cont_atom, np.asarray(qc.ao_spec.get_nlxlylz_per_cont(), dtype=int))
    coeffs = _f64(qc.mo_spec.get_coeffs())
    occ = _f64(qc.mo_spec.get_occ())
    rho_atom, mo_atom = [], []
    for a in atom_indices:
        ao_index = [ll for ll in range(len(ao_atom)) if ao_atom[ll] == a]
        rho = np.zeros(N)
        mos = []
        for ii_mo in range(coeffs.shape[0]):
            mo_info = np.zeros(N)
            for jj in ao_index:
                mo_info += coeffs[ii_mo, jj] * ao_list[jj]
            rho += occ[ii_mo] * mo_list[ii_mo] * mo_info
            mos.append(mo_info)
        rho_atom.append(rho)
        mo_atom.append(mos)
    return rho_atom, mo_atom


def sph2cart(r, theta, phi, kind='port'):
    """cy_grid.sph2cart (orbkit/cy_grid.pyx:58-75): (3, Nr*Ntheta*Nphi) Cartesian coordinates."""
    r, theta, phi = _f64(r), _f64(theta), _f64(phi)
    if kind == 'ref':
        return backend('ref').cy_grid.sph2cart(r, theta, phi)
    out = np.zeros((3, len(r) * len(theta) * len(phi)))
    lib = backend('port').L
    lib.okor_sph2cart.restype = None
    lib.okor_sph2cart(_dp(out), _dp(r), ctypes.c_long(len(r)), _dp(theta), ctypes.c_long(len(theta)), _dp(phi),
                      ctypes.c_long(len(phi)))
    return out


def cyl2cart(r, phi, zed, kind='port'):
    """cy_grid.cyl2cart (orbkit/cy_grid.pyx:79-97)."""
    r, phi, zed = _f64(r), _f64(phi), _f64(zed)
    if kind == 'ref':
        return backend('ref').cy_grid.cyl2cart(r, phi, zed)
    out = np.zeros((3, len(r) * len(phi) * len(zed)))
    lib = backend('port').L
    lib.okor_cyl2cart.restype = None
    lib.okor_cyl2cart(_dp(out), _dp(r), ctypes.c_long(len(r)), _dp(phi), ctypes.c_long(len(phi)), _dp(zed),
                      ctypes.c_long(len(zed)))
    return out
