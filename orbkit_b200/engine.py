"""Host runtime above the C ABI: one context per process/device, cached device handles for the
basis tables, MO coefficients and grids, and pinned host outputs.

PyTorch is used for plumbing only (device selection, pinned host buffers, device output tensors,
streams for timing); every number is produced by libokb200.so.
"""
import ctypes
import hashlib
import os
from collections import OrderedDict

import numpy

from . import _lib
from .tools import get_cart2sph

SINK_AO, SINK_MO, SINK_RHO = 0, 1, 2


def _digest(*arrays):
    h = hashlib.blake2b(digest_size=16)
    for a in arrays:
        if a is None:
            h.update(b'\x00none')
            continue
        a = numpy.ascontiguousarray(a)
        h.update(str(a.dtype).encode() + str(a.shape).encode())
        h.update(a.view(numpy.uint8).reshape(-1).data)
    return h.digest()


def build_cart2sph_csr(ao_spec):
    """CSR form of core.cartesian2spherical (orbkit/core.py:157-174): for every spherical function
    (contraction j0, (l,m)) the Cartesian rows of contraction j0 whose exponent triple matches each
    table term, with value coef*factor.  A table term without a matching Cartesian function raises
    (the reference would silently reuse the previous row index)."""
    lxlylz = [tuple(t) for t in numpy.asarray(ao_spec.get_lxlylz()).tolist()]
    assign = numpy.asarray(ao_spec.get_assign_lxlylz_to_cont()).tolist()
    row_of = {}                      # (contraction, exponent triple) -> Cartesian row (the last match, like the reference)
    for i, (j, e) in enumerate(zip(assign, lxlylz)):
        row_of[(j, e)] = i
    ptr, col, val = [0], [], []
    for j0, lm in ao_spec.get_old_ao_spherical():
        exps, coefs, factor = get_cart2sph(int(lm[0]), int(lm[1]))
        j0 = int(j0)
        for e, c in zip(exps, coefs):
            hit = row_of.get((j0, tuple(e)))
            if hit is None:
                raise ValueError('cartesian2spherical: contraction %d has no Cartesian function %s '
                                 'needed by (l,m)=%s' % (j0, e, tuple(lm)))
            col.append(hit)
            val.append(c * factor)
        ptr.append(len(col))
    return (numpy.asarray(ptr, dtype=numpy.intc), numpy.asarray(col, dtype=numpy.intc),
            numpy.asarray(val, dtype=numpy.float64))


def cart2sph_dense(ao_spec, n_cart=None):
    """dense T[n_sph, n_cart] of the CSR table above (terms that hit the same Cartesian row add up)"""
    ptr, col, val = build_cart2sph_csr(ao_spec)
    n_sph = len(ptr) - 1
    n_cart = int(col.max()) + 1 if n_cart is None and len(col) else (n_cart or 0)
    t = numpy.zeros((n_sph, n_cart))
    numpy.add.at(t, (numpy.repeat(numpy.arange(n_sph), numpy.diff(ptr)), col), val)
    return t


def _stamped_key(obj, slot, compute):
    """`compute()` (a digest of the object's flat arrays), cached on the object for as long as the arrays it was taken
    from are the ones the getters hand out: AOClass / MOClass rebuild them in update() and stamp every rebuild
    (orbitals.py `_stamp`) -- the reference's own caching rule (`_up_to_date`, orbitals.py:336-409, 760-832).
    Objects without a stamp (e.g. the reference's classes) are hashed on every call."""
    if getattr(obj, '_up_to_date', False) and getattr(obj, '_stamp', None) is not None:
        hit = obj.__dict__.get(slot)
        if hit is not None and hit[0] == obj._stamp:
            return hit[1]
    key = compute()
    if getattr(obj, '_up_to_date', False) and getattr(obj, '_stamp', None) is not None:
        try:
            obj.__dict__[slot] = (obj._stamp, key)
        except Exception:
            pass
    return key


class _Handle:
    def __init__(self, ptr, destroy):
        self.ptr, self._destroy = ptr, destroy

    def close(self):
        if self.ptr:
            self._destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """Owns the okb_ctx of one CUDA device and small LRU caches of device handles."""
    CACHE = 8
    VECTOR_CACHE_POINTS = 0           # vector grids are uploaded per call: finding a cached copy means hashing 24 B/point
                                      # (blake2b, ~1 GB/s), which costs several times the upload it would save (1000 points:
                                      # 37 us of hashing in a 300 us call, scripts/prof_latency.py), and cached copies would
                                      # pile up in HBM.  (> 0: cache grids of up to that many points by content.)

    def __init__(self, device=None):
        self.lib = _lib.load()
        if device is None:
            device = int(os.environ.get('LOCAL_RANK', '0'))
            try:
                import torch
                if torch.cuda.is_available():
                    n = torch.cuda.device_count()
                    device = device % max(n, 1)
            except Exception:
                pass
        self.device = device
        ctx = ctypes.c_void_p()
        _lib.check(self.lib.okb_ctx_create(device, ctypes.byref(ctx)))
        self.ctx = ctx
        self._basis = OrderedDict()
        self._mo = OrderedDict()
        self._grid = OrderedDict()

    # ---- caches ---------------------------------------------------------------------------------
    def _put(self, cache, key, handle):
        cache[key] = handle
        while len(cache) > self.CACHE:
            cache.popitem(last=False)     # the C handle is destroyed when its last Python ref dies
        return handle

    @staticmethod
    def _ao_arrays(ao_spec):
        lxlylz = _lib.i32(ao_spec.get_lxlylz())
        assign = _lib.i32(ao_spec.get_nlxlylz_per_cont())
        coeffs = _lib.f64(ao_spec.get_prim_coeffs())
        pnum = _lib.i32(ao_spec.get_nprim_per_cont())
        atoms = _lib.i32(ao_spec.get_assign_cont_to_atoms())
        normalized = int(ao_spec.get_normalized())
        renorm = getattr(ao_spec, 'get_renorm', lambda: None)()
        if renorm is None and len(ao_spec) and isinstance(ao_spec[0], dict) and 'N' in ao_spec[0]:
            renorm = ao_spec[0]['N']
        if renorm is not None:
            renorm = _lib.f64(numpy.asarray(renorm, dtype=float).reshape(-1))
            if renorm.shape[0] != lxlylz.shape[0]:
                raise ValueError("ao_spec[0]['N'] must hold one factor per Cartesian function")
        lm = None
        if bool(ao_spec.spherical):
            lm = numpy.array([[j, l, m] for j, (l, m) in ao_spec.get_old_ao_spherical()], dtype=numpy.intc)
        return lxlylz, assign, coeffs, pnum, atoms, normalized, renorm, lm

    def basis(self, geo_spec, ao_spec):
        """Device basis tables for (geo_spec, ao_spec); returns (handle, n_cart, n_ao, key)."""
        geo = _lib.f64(geo_spec)
        arrays = []

        def ao_digest():
            arrays.append(self._ao_arrays(ao_spec))
            lxlylz, assign, coeffs, pnum, atoms, normalized, renorm, lm = arrays[0]
            return _digest(lxlylz, assign, coeffs, pnum, atoms, numpy.array([normalized]), renorm, lm)
        key = _stamped_key(ao_spec, '_okb_ao_key', ao_digest) + _digest(geo)
        if key in self._basis:
            self._basis.move_to_end(key)
            return self._basis[key]
        lxlylz, assign, coeffs, pnum, atoms, normalized, renorm, lm = arrays[0] if arrays else self._ao_arrays(ao_spec)
        if geo.ndim != 2 or geo.shape[1] != 3:
            raise ValueError('geo_spec must have shape (n_atoms, 3)')
        h = ctypes.c_void_p()
        _lib.check(self.lib.okb_basis_create(
            self.ctx, _lib.iptr(lxlylz), _lib.iptr(assign), _lib.dptr(coeffs), _lib.iptr(pnum),
            _lib.dptr(geo), _lib.iptr(atoms), len(assign), lxlylz.shape[0], coeffs.shape[0], geo.shape[0],
            normalized, _lib.dptr(renorm) if renorm is not None else None, ctypes.byref(h)))
        handle = _Handle(h, self.lib.okb_basis_destroy)
        n_cart = n_ao = lxlylz.shape[0]
        if bool(ao_spec.spherical):
            ptr, col, val = build_cart2sph_csr(ao_spec)
            n_ao = len(ptr) - 1
            _lib.check(self.lib.okb_basis_set_cart2sph(h, n_ao, _lib.iptr(ptr), _lib.iptr(col), _lib.dptr(val)))
        return self._put(self._basis, key, (handle, n_cart, n_ao, key))

    def mos_of(self, basis_entry, mo_spec):
        """MO coefficient handle of an MOClass-like object (get_coeffs / get_occ); a stamped object (see
        _stamped_key) is not copied or hashed again while its flat arrays are unchanged"""
        got = []

        def mo_digest():
            got.append((_lib.f64(mo_spec.get_coeffs()), _lib.f64(mo_spec.get_occ())))
            return _digest(*got[0])
        key = basis_entry[3] + _stamped_key(mo_spec, '_okb_mo_key', mo_digest)
        if key in self._mo:
            self._mo.move_to_end(key)
            return self._mo[key]
        coeffs, occ = got[0] if got else (_lib.f64(mo_spec.get_coeffs()), _lib.f64(mo_spec.get_occ()))
        return self.mos(basis_entry, coeffs, occ, key=key)

    def mos(self, basis_entry, coeffs, occ, key=None):
        handle_b, _, n_ao, bkey = basis_entry
        coeffs = _lib.f64(coeffs)
        occ = _lib.f64(occ)
        if coeffs.ndim != 2 or coeffs.shape[1] != n_ao:
            raise ValueError('MO coefficients have shape %s but the basis has %d AOs' % (coeffs.shape, n_ao))
        if occ.shape != (coeffs.shape[0],):
            raise ValueError('occupation numbers and MO coefficients differ in length')
        if key is None:
            key = bkey + _digest(coeffs, occ)
        if key in self._mo:
            self._mo.move_to_end(key)
            return self._mo[key]
        h = ctypes.c_void_p()
        _lib.check(self.lib.okb_mo_create(self.ctx, handle_b.ptr, coeffs.shape[0], _lib.dptr(coeffs),
                                          _lib.dptr(occ), ctypes.byref(h)))
        handle = _Handle(h, self.lib.okb_mo_destroy)
        handle.keepalive = handle_b           # the MO handle borrows the basis
        handle.n_mo = coeffs.shape[0]
        return self._put(self._mo, key, handle)

    def grid_regular(self, x, y, z):
        x, y, z = _lib.f64(x), _lib.f64(y), _lib.f64(z)
        key = b'r' + _digest(x, y, z)
        if key in self._grid:
            self._grid.move_to_end(key)
            return self._grid[key]
        h = ctypes.c_void_p()
        _lib.check(self.lib.okb_grid_regular(self.ctx, _lib.dptr(x), len(x), _lib.dptr(y), len(y),
                                             _lib.dptr(z), len(z), ctypes.byref(h)))
        handle = _Handle(h, self.lib.okb_grid_destroy)
        handle.npts = len(x) * len(y) * len(z)
        return self._put(self._grid, key, handle)

    def grid_product(self, kind, a0, a1, a2, affine=None):
        """spherical (kind 2: r, theta, phi) or cylindrical (kind 3: r, phi, zed) product grid whose Cartesian
        coordinates are generated on the device; affine = (3x3 matrix, translation) or None"""
        a0, a1, a2 = _lib.f64(a0), _lib.f64(a1), _lib.f64(a2)
        aff = None
        if affine is not None:
            aff = _lib.f64(numpy.concatenate([numpy.asarray(affine[0], dtype=float).reshape(9),
                                              numpy.asarray(affine[1], dtype=float).reshape(3)]))
        key = b'p%d' % kind + _digest(a0, a1, a2, aff)
        if key in self._grid:
            self._grid.move_to_end(key)
            return self._grid[key]
        h = ctypes.c_void_p()
        _lib.check(self.lib.okb_grid_product(self.ctx, kind, _lib.dptr(a0), len(a0), _lib.dptr(a1), len(a1),
                                             _lib.dptr(a2), len(a2), _lib.dptr(aff) if aff is not None else None,
                                             ctypes.byref(h)))
        handle = _Handle(h, self.lib.okb_grid_destroy)
        handle.npts = len(a0) * len(a1) * len(a2)
        return self._put(self._grid, key, handle)

    def grid_vector(self, x, y, z, cache=True):
        x, y, z = _lib.f64(x), _lib.f64(y), _lib.f64(z)
        if not (len(x) == len(y) == len(z)):
            raise ValueError('Dimensions of x-, y-, and z- coordinate differ!')
        cache = cache and len(x) <= self.VECTOR_CACHE_POINTS
        key = b'v' + _digest(x, y, z) if cache else None
        if cache and key in self._grid:
            self._grid.move_to_end(key)
            return self._grid[key]
        h = ctypes.c_void_p()
        _lib.check(self.lib.okb_grid_vector(self.ctx, x.ctypes.data, y.ctypes.data, z.ctypes.data, len(x), 0,
                                            ctypes.byref(h)))
        handle = _Handle(h, self.lib.okb_grid_destroy)
        handle.npts = len(x)
        return self._put(self._grid, key, handle) if cache else handle

    # ---- host output buffers -------------------------------------------------------------------------
    @staticmethod
    def host_array(shape):
        """float64 host array for results; page-locked (via torch's caching host allocator) so that
        device->host copies overlap compute.  Falls back to pageable memory if pinning fails."""
        try:
            import torch
            t = torch.empty(tuple(int(s) for s in shape), dtype=torch.float64, pin_memory=True)
            return t.numpy()
        except Exception:
            return numpy.empty(shape, dtype=numpy.float64)

    # ---- evaluation ---------------------------------------------------------------------------------------
    @staticmethod
    def _ptr(a):
        """address of a NumPy array, or the integer address itself (device pointers, slices of larger arrays)"""
        return a if isinstance(a, int) else a.ctypes.data

    def eval_ao(self, basis_entry, grid, codes, p0=0, p1=None, out=None, flags=0, ld=0):
        """`ld`: row stride (points) of `out` when it addresses a range inside a larger array (0 = dense)"""
        handle_b, _, n_ao, _ = basis_entry
        p1 = grid.npts if p1 is None else p1
        codes = _lib.i32(codes)
        if out is None:
            out = self.host_array((len(codes), n_ao, p1 - p0))
        _lib.check(self.lib.okb_eval_ao_ld(self.ctx, handle_b.ptr, grid.ptr, p0, p1, _lib.iptr(codes), len(codes),
                                           self._ptr(out), ld, flags))
        return out

    def eval_mo(self, mo, grid, codes, p0=0, p1=None, out=None, flags=0, ld=0):
        p1 = grid.npts if p1 is None else p1
        codes = _lib.i32(codes)
        if out is None:
            out = self.host_array((len(codes), mo.n_mo, p1 - p0))
        _lib.check(self.lib.okb_eval_mo_ld(self.ctx, mo.ptr, grid.ptr, p0, p1, _lib.iptr(codes), len(codes),
                                           self._ptr(out), ld, flags))
        return out

    def eval_rho(self, mo, grid, codes, p0=0, p1=None, rho=None, delta=None, want_norm=False, flags=0, ld=0):
        """returns (rho, delta_rho or None, mo_norm or None)"""
        p1 = grid.npts if p1 is None else p1
        codes = _lib.i32(codes)
        n = p1 - p0
        if rho is None:
            rho = self.host_array((n,))
        if delta is None and len(codes):
            delta = self.host_array((len(codes), n))
        norm = numpy.zeros(mo.n_mo) if want_norm else None
        _lib.check(self.lib.okb_eval_rho_ld(
            self.ctx, mo.ptr, grid.ptr, p0, p1, _lib.iptr(codes) if len(codes) else None, len(codes),
            self._ptr(rho), self._ptr(delta) if len(codes) else None, ld,
            norm.ctypes.data if want_norm else None, flags))
        return rho, (delta if len(codes) else None), norm

    # ---- detCI grid contractions ---------------------------------------------------------------------------
    @staticmethod
    def _ci_ncomp(mode, n_terms):
        if mode == _lib.OKB_CI_RHO:
            return 1
        return n_terms if mode == _lib.OKB_CI_PAIRS else (3 * n_terms if mode == _lib.OKB_CI_JAB_PAIRS else 3)

    def ci_contract(self, mode, terms, molist, molistdrv=None, n_mo=None, npts=None, ld=None, out=None, n_eval=None,
                    flags=0):
        """sum over MO pairs of given MO arrays (okb_ci_contract).  terms = (coef, ia, ib) flat arrays;
        molist (n_mo, npts) / molistdrv (3, n_mo, npts) NumPy arrays, or device pointers (ints) with
        OKB_FLAG_IN_DEVICE and explicit n_mo / npts / ld.  Only the first `n_eval` points (default: all)
        are evaluated; the rest of a host output stays 0."""
        coef, ia, ib = _lib.f64(terms[0]), _lib.i32(terms[1]), _lib.i32(terms[2])
        in_dev = bool(flags & _lib.OKB_FLAG_IN_DEVICE)
        out_dev = bool(flags & _lib.OKB_FLAG_OUT_DEVICE)
        if not in_dev:
            molist = _lib.f64(molist)
            n_mo, npts = molist.shape
            ld = npts
            if molistdrv is not None:
                molistdrv = _lib.f64(molistdrv)
                want = (n_mo, npts) if mode == _lib.OKB_CI_PAIRS else (3, n_mo, npts)
                if molistdrv.shape != want:
                    raise ValueError('molistdrv must have shape %s' % (want,))
        ld = npts if ld is None else ld
        n_eval = npts if n_eval is None else n_eval
        ncomp = self._ci_ncomp(mode, len(coef))
        if out is None:
            out = self.host_array((ncomp, npts))
            if n_eval < npts:
                out[:, n_eval:] = 0.0
        _lib.check(self.lib.okb_ci_contract(
            self.ctx, mode, n_mo, n_eval, ld, molist if in_dev else molist.ctypes.data,
            None if molistdrv is None else (molistdrv if in_dev else molistdrv.ctypes.data),
            len(coef), _lib.dptr(coef), _lib.iptr(ia), _lib.iptr(ib), out if out_dev else out.ctypes.data,
            npts, flags))
        return out

    def ci_td(self, w, data, out=None, nk=None, n=None, ld_in=None, ld_out=None, flags=0):
        """out[t, x] = sum_k w[t, k] data[k, x] (okb_ci_td: the time-dependent detCI contractions).  w: NumPy (nt, nk);
        data (nk, n) / out (nt, n): NumPy arrays, or device pointers (ints) with OKB_FLAG_IN_DEVICE / OKB_FLAG_OUT_DEVICE
        and explicit extents."""
        w = _lib.f64(w)
        nt = w.shape[0]
        in_dev = bool(flags & _lib.OKB_FLAG_IN_DEVICE)
        out_dev = bool(flags & _lib.OKB_FLAG_OUT_DEVICE)
        if not in_dev:
            data = _lib.f64(data)
            nk, n = data.shape
            ld_in = n
        if w.shape[1] != nk:
            raise ValueError('weights (%d, %d) do not match %d rows' % (w.shape + (nk,)))
        if out is None:
            out = self.host_array((nt, n))
        ld_out = n if ld_out is None else ld_out
        _lib.check(self.lib.okb_ci_td(self.ctx, nt, nk, n, _lib.dptr(w), data if in_dev else data.ctypes.data,
                                      ld_in if ld_in is not None else n, out if out_dev else out.ctypes.data, ld_out, flags))
        return out

    def ci_jab_full(self, ImS, chi, dchi, mu, out=None, flags=0):
        """cy_ci.get_jab_full on NumPy arrays (okb_ci_jab_full)"""
        ImS, chi, dchi = _lib.f64(ImS), _lib.f64(chi), _lib.f64(dchi)
        nb, npts = chi.shape
        ncomp = dchi.shape[0]
        if out is None:
            out = self.host_array((ncomp, npts))
        _lib.check(self.lib.okb_ci_jab_full(self.ctx, nb, ncomp, npts, npts, _lib.dptr(ImS), chi.ctypes.data,
                                            dchi.ctypes.data, float(mu), out.ctypes.data, npts, flags))
        return out

    def eval_ci(self, mode, terms, mo, grid, drv_codes=(1, 2, 3), p0=0, p1=None, out=None, flags=0):
        """fused: MOs of `mo` evaluated on the device slab by slab and contracted there (okb_eval_ci)"""
        coef, ia, ib = _lib.f64(terms[0]), _lib.i32(terms[1]), _lib.i32(terms[2])
        p1 = grid.npts if p1 is None else p1
        codes = _lib.i32(list(drv_codes))
        out_dev = bool(flags & _lib.OKB_FLAG_OUT_DEVICE)
        if out is None:
            out = self.host_array((self._ci_ncomp(mode, len(coef)), p1 - p0))
        _lib.check(self.lib.okb_eval_ci(self.ctx, mo.ptr, grid.ptr, p0, p1, mode, _lib.iptr(codes), len(coef),
                                        _lib.dptr(coef), _lib.iptr(ia), _lib.iptr(ib),
                                        out if out_dev else out.ctypes.data, flags))
        return out

    def measure_fp64(self, kind=0, min_seconds=0.0):
        """dense FP64 TFLOP/s of this device: kind 0 DFMA, 1 DMMA m8n8k4, 2 mixed; burst when
        min_seconds <= 0, else sustained over back-to-back launches"""
        tf, ms = ctypes.c_double(), ctypes.c_double()
        _lib.check(self.lib.okb_measure_fp64(self.ctx, kind, float(min_seconds), ctypes.byref(tf), ctypes.byref(ms)))
        return tf.value, ms.value

    def sync(self):
        _lib.check(self.lib.okb_ctx_sync(self.ctx))

    def stream_ptr(self):
        return self.lib.okb_ctx_stream(self.ctx)

    def launch_count(self):
        n = _lib.ll()
        _lib.check(self.lib.okb_ctx_launch_count(self.ctx, ctypes.byref(n)))
        return n.value

    def traffic(self):
        """(host->device, device->host) bytes moved by this context so far"""
        a, b = _lib.ll(), _lib.ll()
        _lib.check(self.lib.okb_ctx_traffic(self.ctx, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def clear_caches(self):
        """drop the cached device handles (basis tables, MO coefficients, grids)"""
        self._mo.clear()
        self._basis.clear()
        self._grid.clear()

    def last_kernel(self):
        buf = ctypes.create_string_buffer(256)
        _lib.check(self.lib.okb_ctx_last_kernel(self.ctx, buf, 256))
        return buf.value.decode()


_engine = None


def get_engine():
    """Process-wide engine (device = LOCAL_RANK when launched by torchrun, else 0)."""
    global _engine
    if _engine is None:
        _engine = Engine()
    return _engine


def reset_engine():
    global _engine
    _engine = None
