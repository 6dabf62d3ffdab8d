"""Grid-path operators and drivers with the reference's signatures (orbkit/core.py).

    ao_creator            core.py:38-105     contracted (Cartesian / real spherical) AOs or one derivative
    mo_creator            core.py:107-132    MO contraction of a given AO array
    cartesian2spherical   core.py:135-176    Cartesian -> spherical rows
    rho_compute           core.py:314-605    density / MOs / AOs and derivatives on the module grid
    rho_compute_no_slice  core.py:607-839    same with explicit coordinates and optional components
    calc_mo_matrix        core.py:841-941    MO products on the grid

Differences in mechanism (never in results): the reference expands a regular grid to a vector grid
in place, slices it, forks `numproc` workers and runs one full AO pass + one naive triple-loop MO
contraction per derivative code (slice_rho, core.py:179-308).  Here one fused kernel launch
evaluates every needed derivative set of the AOs tile by tile in shared memory, contracts them in
registers and reduces to rho / delta_rho on chip; regular-grid coordinates are generated in-kernel;
`numproc` and `slice_length` are accepted and ignored; when torch.distributed is initialised the
point range is sharded over the ranks (one GPU each) and every rank's results stream into its range of
one node-shared host array (dist.shared_host_array), which all ranks return.
"""
import numpy

from . import cy_core, cy_grid, grid, options
from . import dist as okdist
from ._lib import OKB_FLAG_EXACT_MIXED
from .display import display
from .engine import get_engine, build_cart2sph_csr, cart2sph_dense
from .qcinfo import QCinfo
from .tools import require, validate_drv, convert, zeros, reshape


def _flags():
    return OKB_FLAG_EXACT_MIXED if options.exact_mixed_derivatives else 0


def _drv_list(drv):
    """normalise `drv` to a list or None (core.py:401-408): 'xyz' -> ['x','y','z']."""
    if drv is None:
        return None
    if isinstance(drv, (int, numpy.integer)):
        return [int(drv)]
    try:
        return list(drv)
    except TypeError:
        return [drv]


def _grid_handle(eng, x, y, z, is_vector):
    if isinstance(x, grid.ProductGrid):          # spherical / cylindrical product grid generated on the device
        return eng.grid_product(x.kind, x.axes[0], x.axes[1], x.axes[2], affine=x.affine)
    return eng.grid_vector(x, y, z) if is_vector else eng.grid_regular(x, y, z)


def _resolve_grid(x, y, z, is_vector, init_vector=False):
    """the explicit coordinates or, by default, the module-global grid (core.py:64-71,702-709)"""
    if all(v is None for v in [x, y, z, is_vector]) and not grid.is_initialized:
        display('\nSetting up the grid...')
        grid.grid_init(is_vector=init_vector)
        display(grid.get_grid())
    pg = grid.product_grid()
    if pg is not None and x is None and y is None and z is None and (is_vector is None or is_vector):
        return pg, None, None, True, (pg.npts,)  # the coordinates exist only as a recipe (grid.sph2cart_vector, ...)
    x = grid.x if x is None else x
    y = grid.y if y is None else y
    z = grid.z if z is None else z
    is_vector = grid.is_vector if is_vector is None else is_vector
    x, y, z = require(x, dtype='f'), require(y, dtype='f'), require(z, dtype='f')
    if is_vector:
        if len(x) != len(y) or len(x) != len(z):
            raise ValueError('Dimensions of x-, y-, and z- coordinate differ!')
        N = (len(x),)
    else:
        N = (len(x), len(y), len(z))
    return x, y, z, bool(is_vector), N


# ---------------------------------------------------------------------------------------------------
# operators
# ---------------------------------------------------------------------------------------------------
def ao_creator(geo_spec, ao_spec, drv=None, x=None, y=None, z=None, is_vector=None):
    """All contracted atomic orbitals, or their derivative w.r.t. `drv`, on the grid.

    Returns ao_list with shape ((NAO,) + N), N = (Nx,Ny,Nz) for a regular grid, (Npts,) for a vector
    grid (core.py:38-105).  Spherical bases are transformed in-kernel."""
    x, y, z, is_vector, N = _resolve_grid(x, y, z, is_vector, init_vector=True)
    code = validate_drv(drv)
    geo_spec = require(geo_spec, dtype='f')
    eng = get_engine()
    basis = eng.basis(geo_spec, ao_spec)
    if int(numpy.prod(N)) == 0:
        return numpy.zeros((basis[2],) + N)
    g = _grid_handle(eng, x, y, z, is_vector)
    out = eng.eval_ao(basis, g, [code], flags=_flags())
    return out[0].reshape((basis[2],) + N)


def mo_creator(ao_list, mo_spec):
    """Molecular orbitals from a given AO array: ((NMO,) + N) (core.py:107-132)."""
    ao_list = require(ao_list, dtype='f')
    shape = ao_list.shape
    mo_coeffs = require(mo_spec.get_coeffs(), dtype='f')
    mo = cy_core.mocreator(ao_list.reshape((shape[0], -1)), mo_coeffs)
    return mo.reshape((len(mo_coeffs),) + shape[1:])


def cartesian2spherical(ao_list, ao_spec):
    """Cartesian -> real spherical AOs for a given AO array (core.py:135-176; table tools.cart2sph).
    Runs as T[n_sph,n_cart] x ao[n_cart,N] on the device (the DMMA contraction behind cy_core.mocreator)."""
    ao_list = require(ao_list, dtype='f')
    T = cart2sph_dense(ao_spec, ao_list.shape[0])
    out = cy_core.mocreator(ao_list.reshape((ao_list.shape[0], -1)), T)
    return out.reshape((T.shape[0],) + ao_list.shape[1:])


# ---------------------------------------------------------------------------------------------------
# the driver
# ---------------------------------------------------------------------------------------------------
def _compute(qc, x, y, z, is_vector, N, calc_ao, calc_mo, drv, want_norm):
    """Shared back end of rho_compute / rho_compute_no_slice.  Returns
         calc_ao/calc_mo:  array (n_sets, n_rows, npts)
         else:             (rho (npts,), delta (len(drv), npts) or None, mo_norm or None)"""
    eng = get_engine()
    geo_spec = require(qc.geo_spec, dtype='f')
    basis = eng.basis(geo_spec, qc.ao_spec)
    npts = int(numpy.prod(N))
    codes = [0] if drv is None else [validate_drv(d) for d in drv]
    if not calc_ao:
        mo = eng.mos_of(basis, qc.mo_spec)
    n_rows = basis[2] if calc_ao else (mo.n_mo if calc_mo else 1)
    if npts == 0:
        if calc_mo or calc_ao:
            return numpy.zeros((len(codes), n_rows, 0))
        return numpy.zeros(0), (numpy.zeros((len(codes), 0)) if drv is not None else None), \
            (numpy.zeros(mo.n_mo) if want_norm else None)
    g = _grid_handle(eng, x, y, z, is_vector)
    flags = _flags()

    if not okdist.is_distributed():
        if calc_ao:
            return eng.eval_ao(basis, g, codes, flags=flags)
        if calc_mo:
            return eng.eval_mo(mo, g, codes, flags=flags)
        ucodes = [] if drv is None else sorted(set(codes))
        rho, delta, norm = eng.eval_rho(mo, g, ucodes, want_norm=want_norm, flags=flags)
        if drv is not None and ucodes != codes:
            delta = delta[[ucodes.index(c) for c in codes]]
        return rho, delta, norm

    # ---- one rank per GPU: contiguous point shards written side by side into ONE node-shared host array ---
    # (each rank's device -> host copies go straight into its own point range and overlap its compute; a
    # barrier completes the result on every rank; no gather, no collective on the data path)
    import torch.distributed as tdist
    rank, world = okdist.rank_world()
    p0, p1 = okdist.shard_range(npts, rank, world)
    n_loc = p1 - p0
    shared = okdist.single_node()         # else: every rank evaluates into a local array and the shards are all-gathered
    if calc_ao or calc_mo:
        if shared:
            full = okdist.shared_host_array((len(codes), n_rows, npts), own=(p0, p1))
            if n_loc:
                dst = full.ctypes.data + 8 * p0
                if calc_ao:
                    eng.eval_ao(basis, g, codes, p0, p1, out=dst, flags=flags, ld=npts)
                else:
                    eng.eval_mo(mo, g, codes, p0, p1, out=dst, flags=flags, ld=npts)
            tdist.barrier()
            return full
        loc = numpy.zeros((len(codes), n_rows, 0))
        if n_loc:
            loc = eng.eval_ao(basis, g, codes, p0, p1, flags=flags) if calc_ao else eng.eval_mo(mo, g, codes, p0, p1, flags=flags)
        return okdist.gather_rows(loc, npts, p0, p1).reshape((len(codes), n_rows, npts))
    ucodes = [] if drv is None else sorted(set(codes))
    norm = None
    if shared:
        full = okdist.shared_host_array((1 + len(ucodes), npts), own=(p0, p1))
        if n_loc:
            _, _, norm = eng.eval_rho(mo, g, ucodes, p0, p1, rho=full.ctypes.data + 8 * p0,
                                      delta=(full.ctypes.data + 8 * (npts + p0)) if ucodes else None,
                                      want_norm=want_norm, flags=flags, ld=npts)
        tdist.barrier()
    else:
        loc = numpy.zeros((1 + len(ucodes), n_loc))
        if n_loc:
            r_loc, d_loc, norm = eng.eval_rho(mo, g, ucodes, p0, p1, want_norm=want_norm, flags=flags)
            loc[0] = r_loc
            if ucodes:
                loc[1:] = d_loc
        full = okdist.gather_rows(loc, npts, p0, p1)
    if want_norm:
        norm = okdist.all_reduce_sum(norm if norm is not None else numpy.zeros(mo.n_mo), eng.device)
    delta = None
    if drv is not None:
        delta = full[1:] if ucodes == codes else full[1:][[ucodes.index(c) for c in codes]]
    return full[0], delta, norm


def _compute_to_store(path, qc, x, y, z, is_vector, N, calc_ao, calc_mo, drv, want_norm, members=()):
    """`_compute` with the results streamed into the file `path` (save_hdf5=, core.py:478-501): point slabs come
    off the device into two page-locked buffers and a writer thread puts them at their place in the file while the
    next slab is evaluated (orbkit_b200/store.py) -- the result never exists as one host array.  Returns
    (arrays, mo_norm): `arrays` maps the dataset names 'rho', 'delta_rho', 'mo_list' / 'ao_list' to read-only arrays
    of the final shapes (memory maps of the finished file)."""
    from . import store as okstore
    eng = get_engine()
    geo_spec = require(qc.geo_spec, dtype='f')
    basis = eng.basis(geo_spec, qc.ao_spec)
    npts = int(numpy.prod(N))
    codes = [0] if drv is None else [validate_drv(d) for d in drv]
    mo = None if calc_ao else eng.mos_of(basis, qc.mo_spec)
    n_rows = basis[2] if calc_ao else (mo.n_mo if calc_mo else 1)
    rank, world = okdist.rank_world()
    barrier = None
    if world > 1:
        import torch.distributed as tdist
        barrier = tdist.barrier
    st = okstore.ResultStore(path, rank, world, barrier)
    for name, value in members:
        st.put(name, value)
    ucodes = [] if drv is None else sorted(set(codes))
    if calc_ao or calc_mo:
        lead = (n_rows,) if drv is None else (len(codes), n_rows)
        sets = [(('ao_list' if calc_ao else 'mo_list'), st.create('ao_list' if calc_ao else 'mo_list', lead + N, npts))]
        rows_total = len(codes) * n_rows
    else:
        sets = [('rho', st.create('rho', N, npts))]
        if drv is not None:
            sets.append(('delta_rho', st.create('delta_rho', (len(codes),) + N, npts)))
        rows_total = 1 + len(ucodes)
    norm = numpy.zeros(mo.n_mo) if (want_norm and mo is not None and not calc_mo) else None
    p_lo, p_hi = okdist.shard_range(npts, rank, world) if world > 1 else (0, npts)
    if npts:
        g = _grid_handle(eng, x, y, z, is_vector)
        flags = _flags()
        slab = max(1024, okstore.SLAB_BYTES // (8 * rows_total) // 1024 * 1024)
        bufs = [eng.host_array((rows_total * slab,)) for _ in range(2)]
        busy = [[], []]
        for i, p0 in enumerate(range(p_lo, p_hi, slab)):
            p1 = min(p0 + slab, p_hi)
            for f in busy[i % 2]:                    # the buffer's previous slab has reached the file
                f.result()
            view = bufs[i % 2][:rows_total * (p1 - p0)].reshape((rows_total, p1 - p0))
            if calc_ao:
                eng.eval_ao(basis, g, codes, p0, p1, out=view, flags=flags)
                busy[i % 2] = [st.write_async(sets[0][1], p0, p1, view)]
            elif calc_mo:
                eng.eval_mo(mo, g, codes, p0, p1, out=view, flags=flags)
                busy[i % 2] = [st.write_async(sets[0][1], p0, p1, view)]
            else:
                _, _, nrm = eng.eval_rho(mo, g, ucodes, p0, p1, rho=view[0], delta=view[1:] if ucodes else None,
                                         want_norm=norm is not None, flags=flags)
                if norm is not None:
                    norm += nrm
                busy[i % 2] = [st.write_async(sets[0][1], p0, p1, view[0])]
                if drv is not None:
                    order = [1 + ucodes.index(c) for c in codes]
                    block = view[order] if order != list(range(1, 1 + len(codes))) else view[1:]
                    busy[i % 2].append(st.write_async(sets[1][1], p0, p1, block))
    st.close()
    if norm is not None and world > 1:
        norm = okdist.all_reduce_sum(norm, eng.device)
    return st.arrays(), norm


def rho_compute(qc, calc_ao=False, calc_mo=False, drv=None, laplacian=False, numproc=1,
                slice_length=1e4, vector=None, save_hdf5=False, **kwargs):
    r"""Density, molecular orbitals, atomic orbitals or derivatives thereof on `orbkit_b200.grid`.

    Same arguments and return values as the reference (core.py:314-605):
      calc_mo (or calc_ao), drv is None      -> mo_list        ((NMO,)+N)
      calc_mo (or calc_ao), drv not None     -> delta_mo_list  ((NDRV,NMO)+N)
      else, drv is None                      -> rho            (N)
      else, drv not None                     -> rho, delta_rho ((NDRV,)+N)
      else, laplacian                        -> rho, delta_rho, laplacian_rho
    `numproc`, `slice_length`, `vector` are accepted for compatibility and ignored, except that
    numproc <= 0 routes to rho_compute_no_slice like the reference does.
    """
    if calc_ao and calc_mo:
        raise ValueError('Choose either calc_ao=True or calc_mo=True')
    elif calc_ao:
        calc_mo = True
    if numproc <= 0:
        return rho_compute_no_slice(qc, calc_ao=calc_ao, calc_mo=calc_mo and not calc_ao, drv=drv,
                                    laplacian=laplacian, **kwargs)
    if laplacian:
        if not (drv is None or drv == ['xx', 'yy', 'zz'] or drv == ['x2', 'y2', 'z2']):
            display('Note: You have set the option `laplacian` and specified values\nfor `drv`. '
                    'Both options are not compatible.\nThe option `drv` has been changed to '
                    '`drv=["xx","yy","zz"]`.')
        drv = ['xx', 'yy', 'zz']
    drv = _drv_list(drv)
    is_drv = drv is not None
    if isinstance(qc, dict):
        qc = QCinfo(qc)
    if not grid.is_initialized:
        display('\nSetting up the grid...')
        grid.grid_init()
        display(grid.get_grid())
    was_vector = grid.is_vector
    x, y, z, _, N = _resolve_grid(None, None, None, was_vector)
    mo_num = qc.ao_spec.get_ao_num() if calc_ao else len(qc.mo_spec)
    display('\nStarting the calculation of the %s...' % ('molecular orbitals' if calc_mo else 'density'))
    display('\nThere are %d contracted %s AOs' % (qc.ao_spec.get_ao_num(),
            'Cartesian' if not qc.ao_spec.spherical else 'spherical') +
            ('' if calc_ao else ' and %d MOs to be calculated.' % mo_num))

    show_norm = (not was_vector) and drv is None and not options.quiet
    if save_hdf5:
        # the reference's datasets (core.py:478-501) streamed slab by slab into the file; the returned arrays are
        # read-only maps of the file instead of a full host copy (`numpy.array(x)` materialises one)
        members = (('grid/x', grid.x), ('grid/y', grid.y), ('grid/z', grid.z), ('grid/is_vector', False),
                   ('grid/is_regular', was_vector))
        arrays, mo_norm = _compute_to_store(save_hdf5, qc, x, y, z, was_vector, N, calc_ao, calc_mo, drv,
                                            show_norm and not calc_mo, members)
        if calc_mo:
            res = arrays['ao_list' if calc_ao else 'mo_list']
            res = res.reshape((1, mo_num, -1)) if drv is None else res.reshape((len(drv), mo_num, -1))
        else:
            res = (arrays['rho'].reshape(-1), arrays['delta_rho'].reshape((len(drv), -1)) if is_drv else None, mo_norm)
    else:
        res = _compute(qc, x, y, z, was_vector, N, calc_ao, calc_mo, drv, want_norm=show_norm and not calc_mo)

    if calc_mo:
        mo_list = res[0] if drv is None else res
        if show_norm:
            display('\nNorm of the MOs:')
            labels = qc.ao_spec.get_labels() if calc_ao else qc.mo_spec.get_labels(format='print')
            for i in range(mo_num):
                display('\t%.6f\t%s %s' % (numpy.sum(numpy.square(mo_list[i])) * grid.d3r,
                                             'AO' if calc_ao else 'MO', labels[i]))
        return mo_list.reshape(((mo_num,) if drv is None else (len(drv), mo_num,)) + N)

    rho, delta_rho, mo_norm = res
    if show_norm:
        display('\nNorm of the MOs:')
        labels = qc.mo_spec.get_labels(format='print')
        for i in range(mo_num):
            display('\t%.6f\tMO %s' % (mo_norm[i] * grid.d3r, labels[i]))
    if not was_vector and not options.quiet:     # (a reduction over the whole grid: skipped when nothing is printed)
        display('We have ' + str(numpy.sum(rho) * grid.d3r) + ' electrons.')
    rho = rho.reshape(N)
    if not is_drv:
        return rho
    delta_rho = delta_rho.reshape((len(drv),) + N)
    if laplacian:
        return rho, delta_rho, delta_rho.sum(axis=0)
    return rho, delta_rho


def rho_compute_no_slice(qc, calc_ao=False, calc_mo=False, drv=None, laplacian=False,
                         return_components=False, x=None, y=None, z=None, is_vector=None, **kwargs):
    r"""As rho_compute but with explicit coordinates and, if `return_components`, the AO and MO
    arrays as well (core.py:607-839):
      calc_mo, drv None       -> ao_list, mo_list
      calc_mo, drv            -> delta_ao_list, delta_mo_list
      else, drv None          -> ao_list, mo_list, rho
      else, drv               -> ao_list, mo_list, rho, delta_ao_list, delta_mo_list, delta_rho[, laplacian_rho]
    """
    if calc_ao and calc_mo:
        raise ValueError('calc_ao and calc_mo are mutually exclusive arguments.'
                         'Use calc_mo and return_components instead!')
    x, y, z, is_vector, N = _resolve_grid(x, y, z, is_vector)
    was_vector = is_vector
    d3r = 1.0
    if not is_vector:
        for i in (x, y, z):
            if len(i) > 1:
                d3r *= (i[1] - i[0])
    if laplacian:
        drv = ['xx', 'yy', 'zz']
    if isinstance(qc, dict):
        qc = QCinfo(qc)
    drv = _drv_list(drv)
    n_ao = qc.ao_spec.get_ao_num()

    def shaped(a, lead):
        return a.reshape(lead + N)

    delta_ao_list = delta_mo_list = None
    if drv is not None:
        if calc_ao or return_components:
            delta_ao_list = shaped(_compute(qc, x, y, z, is_vector, N, True, True, drv, False), (len(drv), n_ao))
        if calc_ao:
            return delta_ao_list
        if calc_mo or return_components:
            delta_mo_list = shaped(_compute(qc, x, y, z, is_vector, N, False, True, drv, False),
                                   (len(drv), len(qc.mo_spec)))
        if calc_mo:
            return (delta_ao_list, delta_mo_list) if return_components else delta_mo_list
    ao_list = mo_list = None
    if calc_ao or return_components:
        ao_list = shaped(_compute(qc, x, y, z, is_vector, N, True, True, None, False)[0], (n_ao,))
    if calc_ao:
        return ao_list
    if calc_mo or return_components:
        mo_list = shaped(_compute(qc, x, y, z, is_vector, N, False, True, None, False)[0], (len(qc.mo_spec),))
        if not was_vector and not options.quiet:
            display('\nNorm of the MOs:')
            for i in range(len(mo_list)):
                display('\t%.6f\tMO %s' % (numpy.sum(mo_list[i] ** 2) * d3r, qc.mo_spec[i].get('sym', '')))
    if calc_mo:
        return (ao_list, mo_list) if return_components else mo_list

    rho, delta_rho, _ = _compute(qc, x, y, z, is_vector, N, False, False, drv, False)
    rho = rho.reshape(N)
    if not was_vector and not options.quiet:
        display('We have ' + str(numpy.sum(rho) * d3r) + ' electrons.')
    if drv is None:
        return (ao_list, mo_list, rho) if return_components else rho
    delta_rho = delta_rho.reshape((len(drv),) + N)
    delta = (delta_rho, delta_rho.sum(axis=0)) if laplacian else (delta_rho,)
    return ((ao_list, mo_list, rho, delta_ao_list, delta_mo_list,) + delta
            if return_components else (rho,) + delta)


def _ket_sets(drv):
    """the ket derivative sets of calc_mo_matrix (core.py:881-890): None -> the MO values, a string -> that one
    derivative, a list -> one set per entry"""
    if drv is None:
        return [None]
    return list(drv) if isinstance(drv, list) else [drv]


def calc_mo_matrix(qc_a, qc_b=None, drv=None, numproc=1, slice_length=1e4, save_hdf5=False, **kwargs):
    """mo_matrix[d, n, m] = mo_bra[n] * d_drv[d] mo_ket[m] on the module grid: ((NDRV, NMO_a, NMO_b) + N)
    (core.py:841-941).

    One QCinfo (the case extras.calc_jmo uses): per ket set the MOs and their derivative are evaluated slab by
    slab on the device and the NMO^2 products are formed there (okb_eval_ci, OKB_CI_PAIRS); only the products
    cross PCIe.  Two QCinfos: the reference's branch raises TypeError for every input (`drv[ibra]` with a list
    index, core.py:906); its evident intent -- bra = MO values of qc_a, ket = the requested sets of qc_b -- is
    implemented: both MO arrays are evaluated on the device and their products formed by okb_ci_contract."""
    from ._lib import OKB_CI_PAIRS
    sets = _ket_sets(drv)
    codes = [validate_drv(d) for d in sets]
    x, y, z, is_vector, N = _resolve_grid(None, None, None, None, init_vector=False)
    npts = int(numpy.prod(N))
    eng = get_engine()
    same = qc_b is None or qc_b is qc_a or qc_a == qc_b
    nmo_a = len(qc_a.mo_spec)
    nmo_b = nmo_a if same else len(qc_b.mo_spec)
    if npts == 0 or nmo_a == 0 or nmo_b == 0:
        return numpy.zeros((len(codes), nmo_a, nmo_b) + N)
    ia = numpy.repeat(numpy.arange(nmo_a, dtype=numpy.intc), nmo_b)
    ib = numpy.tile(numpy.arange(nmo_b, dtype=numpy.intc), nmo_a)
    out = eng.host_array((len(codes), nmo_a * nmo_b, npts))
    if same:
        basis = eng.basis(require(qc_a.geo_spec, dtype='f'), qc_a.ao_spec)
        mo = eng.mos_of(basis, qc_a.mo_spec)
        g = _grid_handle(eng, x, y, z, is_vector)
        terms = (numpy.zeros(len(ia)), ia, ib)
        for d, code in enumerate(codes):
            eng.eval_ci(OKB_CI_PAIRS, terms, mo, g, drv_codes=[code], out=out[d], flags=_flags())
    else:
        bra = _compute(qc_a, x, y, z, is_vector, N, False, True, None, False)[0]
        ket = _compute(qc_b, x, y, z, is_vector, N, False, True, sets, False)
        terms = (numpy.zeros(len(ia)), ia, (ib + nmo_a).astype(numpy.intc))
        for d in range(len(codes)):
            eng.ci_contract(OKB_CI_PAIRS, terms, numpy.concatenate([bra, ket[d]]), out=out[d])
    out = out.reshape((len(codes), nmo_a, nmo_b) + N)
    if save_hdf5:
        # the reference's file (core.py:919-938): the grid and the dataset 'mo_matrix'
        from . import store as okstore
        members = dict(x=grid.x, y=grid.y, z=grid.z, is_vector=bool(grid.is_vector))
        if okstore.wants_hdf5(save_hdf5) and okstore.have_h5py():
            okstore.hdf5_write(save_hdf5, mode='w', mo_matrix=out, grid=members)
        else:
            okstore.npz_write(save_hdf5, mode='w', compress=False, mo_matrix=out, grid=members)
    return out
