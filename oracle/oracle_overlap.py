"""CPU ORACLE for the analytic overlap integrals (orbkit/cy_overlap.pyx:24-206, c_non-grid-based.c:9-52,
analytical_integrals.py:37-116, 318-326).

TEST INFRASTRUCTURE ONLY (see oracle.py): imported by tests/ as the checker, never by orbkit_b200.

Backends: "port" = libokoracle.so (okor_aooverlap, okor_cca_norm, okor_mooverlapmatrix: our C restatement in the
reference's operation order), "ref" = the reference's own cy_overlap module compiled into oracle/_ref (pinned against
each other in tests/test_oracle_overlap.py; bit for bit, both call the same libm).
"""
import ctypes
import glob
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
        for name in ('okor_aooverlap', 'okor_cca_norm', 'okor_mooverlapmatrix'):
            getattr(_lib, name).restype = None
    return _lib


def have_ref():
    return bool(glob.glob(os.path.join(_HERE, '_ref', 'cy_overlap*.so')))


def _ref():
    d = os.path.join(_HERE, '_ref')
    if d not in sys.path:
        sys.path.insert(0, d)
    import cy_overlap
    return cy_overlap


def aooverlap(geo_a, geo_b, lxlylz_a, lxlylz_b, assign, ao_coeffs, pnum_list, atom_indices, drv, is_normalized,
              kind='port'):
    f = lambda v: np.require(v, dtype=np.float64, requirements='CA')
    i = lambda v: np.require(v, dtype=np.intc, requirements='CA')
    geo_a, geo_b, ao_coeffs = f(geo_a), f(geo_b), f(ao_coeffs)
    lxlylz_a, lxlylz_b, assign, pnum_list, atom_indices = i(lxlylz_a), i(lxlylz_b), i(assign), i(pnum_list), i(atom_indices)
    if kind == 'ref':
        return _ref().aooverlap(geo_a, geo_b, lxlylz_a, lxlylz_b, assign, ao_coeffs, pnum_list, atom_indices, int(drv),
                                int(is_normalized))
    n = lxlylz_a.shape[0]
    out = np.zeros((n, n))
    _port().okor_aooverlap(_dp(out), _dp(geo_a), _dp(geo_b), _ip(lxlylz_a), _ip(lxlylz_b), ctypes.c_long(n), _ip(assign),
                           _dp(ao_coeffs), _ip(pnum_list), _ip(atom_indices), ctypes.c_long(len(assign)), ctypes.c_int(drv),
                           ctypes.c_int(is_normalized))
    return out


def ommited_cca_norm(lxlylz, kind='port', with_divisor=True):
    lxlylz = np.require(lxlylz, dtype=np.intc, requirements='CA')
    if kind == 'ref':
        return _ref().ommited_cca_norm(lxlylz) if with_divisor else _ref().tmol_aomix_norm(lxlylz)
    out = np.zeros(lxlylz.shape[0])
    _port().okor_cca_norm(_dp(out), _ip(lxlylz), ctypes.c_long(lxlylz.shape[0]), ctypes.c_int(1 if with_divisor else 0))
    return out


def mooverlapmatrix(mo_a, mo_b, aoom, kind='port'):
    f = lambda v: np.require(v, dtype=np.float64, requirements='CA')
    mo_a, mo_b, aoom = f(mo_a), f(mo_b), f(aoom)
    if kind == 'ref':
        return _ref().mooverlapmatrix(mo_a, mo_b, aoom, 0, mo_a.shape[0])
    out = np.zeros((mo_a.shape[0], mo_b.shape[0]))
    _port().okor_mooverlapmatrix(_dp(out), _dp(mo_a), _dp(mo_b), _dp(aoom), ctypes.c_long(mo_a.shape[0]),
                                 ctypes.c_long(mo_b.shape[0]), ctypes.c_long(aoom.shape[0]))
    return out


def ao_overlap_of(qc_arrays_or_aospec, geo_a, geo_b=None, drv=0, kind='port', lxlylz_b=None):
    """get_ao_overlap (analytical_integrals.py:37-116) WITHOUT the spherical transformation: the Cartesian overlap matrix
    of an orbkit_b200 / reference AOClass-like object (getters get_lxlylz, get_nlxlylz_per_cont, get_prim_coeffs,
    get_nprim_per_cont, get_assign_cont_to_atoms, get_normalized)."""
    ao = qc_arrays_or_aospec
    la = np.array(ao.get_lxlylz(), dtype=np.intc)
    lb = la.copy() if lxlylz_b is None else np.array(lxlylz_b, dtype=np.intc)
    geo_b = geo_a if geo_b is None else geo_b
    return aooverlap(geo_a, geo_b, la, lb, ao.get_nlxlylz_per_cont(), ao.get_prim_coeffs(), ao.get_nprim_per_cont(),
                     ao.get_assign_cont_to_atoms(), drv, int(bool(ao.get_normalized())), kind=kind)
