"""CPU-only checks of the host side: data-model mirror, grid state, drv parsing, the cart->sph CSR,
the C ABI surface (library loads and exports every symbol of include/okb200.h; no compute without
a GPU, and it says so loudly), and the multi-rank sharding logic on gloo (world_size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy
import pytest

from conftest import FIXTURES, REPO, golden_qc, load_golden


def test_library_exports_every_declared_symbol():
    from orbkit_b200 import _lib
    _lib.build()
    lib = _lib.load()
    header = open(os.path.join(REPO, 'include', 'okb200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = set(re.findall(r'\b(okb_[a-z0-9_]+)\s*\(', header))
    assert len(declared) >= 25
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.okb_version() >= 100
    # scalar helpers are host code and work without a GPU
    assert lib.okb_aonorm(0, 0, 0, 1.0, 1) == 1.0
    assert abs(lib.okb_aonorm(0, 0, 0, 1.0, 0) - (2 / numpy.pi) ** 0.75) < 1e-15


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from orbkit_b200 import _lib
    lib = _lib.load()
    ctx = ctypes.c_void_p()
    rc = lib.okb_ctx_create(0, ctypes.byref(ctx))
    assert rc != 0 and not ctx.value
    assert b'no CPU fallback' in lib.okb_last_error() or b'CUDA' in lib.okb_last_error()
    import orbkit_b200 as ok
    from orbkit_b200 import engine
    engine.reset_engine()
    qc, a = golden_qc('nh3_molpro')
    ok.grid.set_grid(a['vx'], a['vy'], a['vz'], is_vector=True)
    with pytest.raises(RuntimeError):
        ok.rho_compute(qc)
    with pytest.raises(RuntimeError):
        ok.cy_core.mocreator(numpy.zeros((3, 4)), numpy.zeros((2, 3)))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(REPO, 'orbkit_b200')
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(root, f)).read()
                assert 'import oracle' not in src and 'okoracle' not in src and 'oracle/' not in src, f


@pytest.mark.parametrize('name', FIXTURES)
def test_data_model_roundtrip(name):
    """AOClass / MOClass getters reproduce the flat arrays the reference's own classes produced"""
    qc, a = golden_qc(name)
    ao = qc.ao_spec
    for key, getter in [('_assign_cont_to_atoms', ao.get_assign_cont_to_atoms), ('_nprim_per_cont', ao.get_nprim_per_cont),
                        ('_prim_coeffs', ao.get_prim_coeffs), ('_assign_prim_to_cont', ao.get_assign_prim_to_cont),
                        ('_lxlylz', ao.get_lxlylz), ('_assign_lxlylz_to_cont', ao.get_assign_lxlylz_to_cont),
                        ('_nlxlylz_per_cont', ao.get_nlxlylz_per_cont)]:
        got = getter()
        assert numpy.array_equal(got, a['ao.' + key]), key
        assert got.flags['C_CONTIGUOUS'] and got.dtype in (numpy.intc, numpy.float64)
    assert ao.get_normalized() == int(a['ao.normalized'])
    assert ao.spherical == bool(a['ao.spherical'])
    if ao.spherical:
        assert numpy.array_equal(numpy.array(ao.get_lm()), a['ao._lm'])
        assert numpy.array_equal(ao.get_assign_lm_to_cont(), a['ao._assign_lm_to_cont'])
        assert ao.get_ao_num() == len(a['ao._lm'])
    assert numpy.array_equal(qc.mo_spec.get_coeffs(), a['mo.coeffs'])
    assert numpy.array_equal(qc.mo_spec.get_occ(), a['mo.occ'])
    assert qc.mo_spec.get_coeffs().shape[1] == ao.get_ao_num()
    # list-of-dict <-> flat round trip, copy semantics, equality
    qc2 = qc.copy()
    assert qc2 == qc
    qc2.mo_spec[0]['occ_num'] = 7.5
    qc2.mo_spec.update()
    assert qc2.mo_spec.get_occ()[0] == 7.5 and qc.mo_spec.get_occ()[0] != 7.5
    ao2 = type(ao)(restart=ao.todict())
    assert ao2 == ao


def test_set_lm_dict_order_and_mo_selection():
    from orbkit_b200 import synth
    qc = synth.to_qcinfo(synth.make_molecule(n_heavy=1, n_light=0, n_mo=6, seed=1))
    lm = qc.ao_spec.get_lm()
    # s, s, s, s, p(1,1),(1,-1),(1,0) ... d: 0,+1,-1,+2,-2  (orbitals.py:303-316)
    assert lm[:4] == [(0, 0)] * 4 and lm[4:7] == [(1, 1), (1, -1), (1, 0)]
    d0 = [i for i, t in enumerate(lm) if t[0] == 2][0]
    assert lm[d0:d0 + 5] == [(2, 0), (2, 1), (2, -1), (2, 2), (2, -2)]
    mo = qc.mo_spec
    for i in range(3, 6):
        mo[i]['occ_num'] = 0.0
    mo.update()
    assert mo.get_homo() == 2 and mo.get_lumo() == 3
    assert list(mo.select('homo-1:lumo+1').get_indices()) == [1, 2, 3]
    assert len(mo[[0, 5]]) == 2 and len(mo[1:]) == 5 and len(mo.select('all_mo')) == 6
    assert list(mo['lumo'].get_indices()) == [3]
    with pytest.raises(ValueError):
        bad = qc.ao_spec[:]
        bad[0]['pnum'] = -abs(bad[0]['pnum'])
        bad.update()


def test_homo_lumo_on_an_unrestricted_mo_list():
    """orbitals.py:566-588: get_homo / get_lumo first sort the MO list by energy IN PLACE (with a warning).  For an
    unrestricted set stored as alpha block + beta block (fchk / molden with all_mo=True) the unsorted list would put
    the first alpha virtual in front of the occupied beta orbitals."""
    import warnings
    from orbkit_b200 import synth
    qc = synth.to_qcinfo(synth.make_molecule(n_heavy=1, n_light=0, n_mo=8, seed=2))
    mo = qc.mo_spec
    #            alpha: 2 occupied, 2 virtual          beta: 1 occupied, 3 virtual
    energies = [-1.0, -0.5, 0.2, 0.9, -0.9, 0.1, 0.3, 1.1]
    occ = [1.0, 1.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0]
    for i in range(8):
        mo[i]['energy'], mo[i]['occ_num'], mo[i]['sym'] = energies[i], occ[i], '%d.%s' % (i % 4 + 1, 'ab'[i // 4])
    mo.update()
    assert not mo.is_energy_sorted
    assert mo.get_homo(sort=False) == 4 and mo.get_lumo(sort=False) == 2       # straddling the two blocks
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        homo = mo.get_homo()
    assert any('not sorted by energy' in str(x.message) for x in w)
    # sorted energies: -1.0 -0.9 -0.5 | 0.1 0.2 0.3 0.9 1.1
    assert mo.is_energy_sorted and homo == 2 and mo.get_lumo() == 3
    assert [m['sym'] for m in mo] == ['1.a', '1.b', '2.a', '2.b', '3.a', '3.b', '4.a', '4.b']
    assert list(mo.select('homo-1:lumo+1').get_indices()) == [1, 2, 3]
    assert numpy.allclose(mo.get_eig(), sorted(energies)) and numpy.allclose(mo.get_occ(), [1, 1, 1, 0, 0, 0, 0, 0])


def test_validate_drv_and_tables():
    from orbkit_b200.tools import validate_drv, exp, cart2sph, get_cart2sph, l_deg
    expect = {None: 0, 'None': 0, '': 0, 'x': 1, 'y': 2, 'z': 3, 'xx': 4, 'x2': 4, 'yy': 5, 'y2': 5, 'zz': 6,
              'z2': 6, 'xy': 7, 'yx': 7, 'xz': 8, 'zx': 8, 'yz': 9, 'zy': 9}
    for k, v in expect.items():
        assert validate_drv(k) == v
    for i in range(10):
        assert validate_drv(i) == i        # ints pass through: drv=0 means "no derivative"
    for bad in ('q', 'xyz', 10, -1, 1.5):
        with pytest.raises(ValueError):
            validate_drv(bad)
    for l in range(5):
        assert len(exp[l]) == l_deg(l) == (l + 1) * (l + 2) // 2
        assert len(cart2sph[l]) == l_deg(l, cartesian_basis=False) == 2 * l + 1
        assert all(sum(t) == l for t in exp[l])
    assert get_cart2sph(2, 0)[1] == [1., -0.5, -0.5]


def test_cart2sph_csr_equals_reference_loop(oracle_mod):
    """the CSR handed to the device reproduces core.cartesian2spherical on random data"""
    from orbkit_b200.engine import build_cart2sph_csr
    rng = numpy.random.default_rng(3)
    for name in ('h2o_gaussian_sph', 'lih_psi4_sph_f', 'synth_small_sph'):
        qc, _ = golden_qc(name)
        ptr, col, val = build_cart2sph_csr(qc.ao_spec)
        cart = rng.standard_normal((len(qc.ao_spec.get_lxlylz()), 13))
        ref = oracle_mod.cartesian2spherical(cart, qc.ao_spec)
        got = numpy.zeros_like(ref)
        for j in range(len(ptr) - 1):
            for t in range(ptr[j], ptr[j + 1]):
                got[j] += val[t] * cart[col[t]]
        assert numpy.array_equal(got, ref)


def test_grid_module_state():
    from orbkit_b200 import grid, cy_grid
    grid.reset_grid()
    grid.min_, grid.max_, grid.N_ = [-1., -2., 0.], [1., 2., 0.], [3, 5, 1]
    grid.delta_ = numpy.zeros((3, 1))
    grid.grid_init(force=True)
    assert grid.get_shape() == (3, 5, 1) and not grid.is_vector and grid.is_regular
    assert numpy.allclose(grid.x, [-1, 0, 1]) and numpy.allclose(grid.y, [-2, -1, 0, 1, 2]) and grid.z.tolist() == [0.]
    assert abs(grid.d3r - 1.0) < 1e-15
    x0, y0, z0 = grid.tolist()
    grid.grid2vector()
    assert grid.is_vector and len(grid.x) == 15
    # x slowest, z fastest (cy_grid.pyx:22-29)
    assert numpy.array_equal(grid.x, numpy.repeat(x0, 5)) and numpy.array_equal(grid.y, numpy.tile(y0, 3))
    v = numpy.arange(15.0)
    assert numpy.array_equal(grid.mv2g(d=v), v.reshape(3, 5, 1))
    grid.vector2grid(3, 5, 1)
    assert numpy.array_equal(grid.x, x0) and numpy.array_equal(grid.y, y0) and numpy.array_equal(grid.z, z0)
    with pytest.raises(ValueError):
        grid.grid2vector(); grid.vector2grid(2, 5, 1)
    qc, _ = golden_qc('h2o_molpro_cart')
    grid.adjust_to_geo(qc, extend=2.0, step=1)
    grid.grid_init(is_vector=False, force=True)
    g = load_golden('ref_rho_compute')
    assert numpy.allclose(grid.x, g['x']) and numpy.allclose(grid.y, g['y']) and numpy.allclose(grid.z, g['z'])
    out = cy_grid.grid2vector(g['x'], g['y'], g['z'])
    assert out.shape == (3, 315)
    back = cy_grid.vector2grid(out[0], out[1], out[2], 5, 9, 7)
    assert all(numpy.array_equal(p, q) for p, q in zip(back, (g['x'], g['y'], g['z'])))
    grid.set_grid([0., 1.], [0., 1.], [0., 1.], is_vector=True)
    assert grid.is_vector and not grid.is_regular and grid.get_shape() == (2,)


def test_oracle_grid2vector_matches(oracle_mod):
    from orbkit_b200 import cy_grid
    x, y, z = numpy.linspace(0, 1, 3), numpy.linspace(2, 3, 4), numpy.linspace(-1, 0, 5)
    assert numpy.array_equal(cy_grid.grid2vector(x, y, z), oracle_mod.backend('port').grid2vector(x, y, z))


def test_shard_ranges_cover_and_balance():
    from orbkit_b200.dist import shard_range, shard_sizes, ALIGN
    for npts in (0, 1, 127, 128, 129, 1000, 200 ** 3, 256 ** 3 + 17):
        for world in (1, 2, 3, 4, 8):
            r = shard_sizes(npts, world)
            assert r[0][0] == 0 and r[-1][1] == npts
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) < 2 * ALIGN
            assert all(a % ALIGN == 0 for a, _ in r if a < npts)


_GLOO_SCRIPT = r'''
import os, sys, numpy, torch
import torch.distributed as dist
import os
sys.path.insert(0, %(repo)r)
from orbkit_b200 import dist as okdist
dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%(port)d', rank=int(sys.argv[1]), world_size=2)
rank, world = okdist.rank_world()
assert okdist.is_distributed() and world == 2
npts = 1000
full = numpy.arange(4 * npts, dtype=numpy.float64).reshape(4, npts)
p0, p1 = okdist.shard_range(npts, rank, world)
local = torch.from_numpy(full[:, p0:p1].copy())
got = okdist.gather_points(local, npts).numpy()
assert numpy.array_equal(got, full), 'gather'
norm = okdist.all_reduce_sum(numpy.array([1.0 + rank, 2.0]))
assert numpy.allclose(norm, [3.0, 4.0]), norm
# one shard empty
p0, p1 = okdist.shard_range(100, rank, world)
local = torch.from_numpy(numpy.arange(100.0)[None, p0:p1].copy())
assert numpy.array_equal(okdist.gather_points(local, 100).numpy()[0], numpy.arange(100.0))
# node-shared host array: every rank writes its own point range, a barrier completes it on both.
# Life time (ADVICE r01): a result stays valid for as long as ANY rank holds it -- five same-shape calls whose results
# are all kept (rank 1 keeps only views of them) never alias; a segment is reused once every rank dropped its array.
assert okdist.single_node()
kept = []
for rep in range(5):
    p0, p1 = okdist.shard_range(npts, rank, world)
    shared = okdist.shared_host_array((4, npts), pin=False, own=(p0, p1))
    shared[:, p0:p1] = full[:, p0:p1] + rep
    dist.barrier()
    assert numpy.array_equal(shared, full + rep), 'shared host array'
    kept.append(shared if rank == 0 else shared[1:, 5:])
    del shared
for rep, arr in enumerate(kept):
    assert numpy.array_equal(arr, (full + rep) if rank == 0 else (full + rep)[1:, 5:]), 'result %%d was overwritten' %% rep
assert len(okdist._segments) == 5
dist.barrier()
# rank 0 drops everything, rank 1 still holds result 3: only segments free on BOTH ranks are handed out again
if rank == 0:
    kept = []
else:
    kept = [kept[3]]
again = okdist.shared_host_array((4, npts), pin=False, own=okdist.shard_range(npts, rank, world))
assert len(okdist._segments) == 5, 'a free segment is reused'
again[:] = -1.0
dist.barrier()
if rank == 1:
    assert numpy.array_equal(kept[0], (full + 3)[1:, 5:]), 'a held result was handed out again'
del again
# another shape gets its own segment; the budget evicts free segments (same decision on both ranks)
okdist.CACHE_BYTES = 4 * npts * 8 * 2
other = okdist.shared_host_array((2, npts), pin=False)
other[:] = rank
dist.barrier()
assert len(okdist._segments) <= 3, len(okdist._segments)
if rank == 1:
    assert numpy.array_equal(kept[0], (full + 3)[1:, 5:])
del other, kept
# multi-node fallback: the shards are all-gathered instead
p0, p1 = okdist.shard_range(npts, rank, world)
assert numpy.array_equal(okdist.gather_rows(full[:, p0:p1], npts, p0, p1), full)
dist.barrier()
assert not [f for f in os.listdir('/dev/shm') if f.startswith('okb200_')], 'segments are unlinked once attached'
dist.barrier()
dist.destroy_process_group()
print('rank', rank, 'ok')
'''


def test_world_size_2_gloo_gather_and_allreduce(tmp_path):
    import socket
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / 'gloo_worker.py'
    script.write_text(_GLOO_SCRIPT % {'repo': REPO, 'port': port})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_product_grid_state_is_lazy_and_materialises_like_the_reference():
    """grid.sph2cart_vector / cyl2cart_vector / grid_sym_op / grid_translate keep a device recipe; grid.x/y/z
    appear on first access with the reference's values (grid.py:293-321, 373-419)"""
    from orbkit_b200 import grid, cy_grid
    r, th, ph = numpy.linspace(0.1, 3, 4), numpy.linspace(0, numpy.pi, 5), numpy.linspace(0, 2 * numpy.pi, 6)
    grid.sph2cart_vector(r, th, ph)
    pg = grid.product_grid()
    assert pg is not None and pg.kind == 2 and pg.npts == 120 and pg.affine is None
    assert grid.is_initialized and grid.is_vector and not grid.is_regular
    S = grid.rot(0.3, 2)
    grid.grid_sym_op(S)
    grid.grid_translate(0.5, -1.0, 2.0)
    grid.grid_sym_op(grid.reflect(numpy.array([0, 1])))
    assert grid.product_grid() is pg and pg.affine is not None
    ref = numpy.dot(grid.reflect(numpy.array([0, 1])), numpy.dot(S, cy_grid.sph2cart(r, th, ph)) + numpy.array([[0.5], [-1.0], [2.0]]))
    x = grid.x                                  # first access materialises and drops the recipe
    assert grid.product_grid() is None
    assert numpy.allclose(numpy.array([x, grid.y, grid.z]), ref, rtol=0, atol=1e-14)
    assert grid.get_shape() == (120,)
    # cylindrical, untouched: exactly the reference's expression values
    grid.cyl2cart_vector(r, ph, [0.0, 1.0])
    assert grid.product_grid().kind == 3
    assert numpy.array_equal(numpy.array(grid.tolist()), cy_grid.cyl2cart(r, ph, [0.0, 1.0]))
    assert grid.product_grid() is None
    # set_grid / grid_init replace a pending recipe
    grid.sph2cart_vector(r, th, ph)
    grid.set_grid(numpy.arange(3.0), numpy.arange(2.0), numpy.arange(4.0), is_vector=False)
    assert grid.product_grid() is None and len(grid.x) == 3 and not grid.is_vector
    assert numpy.allclose(grid.inversion(), -numpy.eye(3)) and grid.rot(0.0, 1).shape == (3, 3)
    with pytest.raises(ValueError):
        grid.grid_sym_op(numpy.eye(2))


# ---- readers (orbkit_b200.read) ------------------------------------------------------------------------------------
def _flat_qc(qc):
    ao, mo = qc.ao_spec.todict(), qc.mo_spec.todict()
    out = {'geo_spec': numpy.asarray(qc.geo_spec), 'geo_info': numpy.asarray(qc.geo_info, dtype=str)}
    for k in ['normalized', 'spherical', '_assign_cont_to_atoms', '_nprim_per_cont', '_prim_coeffs',
              '_assign_prim_to_cont', '_lxlylz', '_assign_lxlylz_to_cont', '_nlxlylz_per_cont']:
        out['ao.' + k] = numpy.asarray(ao[k])
    out['ao._cont_types'] = numpy.asarray(ao['_cont_types'], dtype=str)
    if ao['spherical']:
        out['ao._lm'] = numpy.asarray(ao['_lm'], dtype=numpy.intc)
        out['ao._assign_lm_to_cont'] = numpy.asarray(ao['_assign_lm_to_cont'])
    for k in ('coeffs', 'occ', 'eig'):
        out['mo.' + k] = numpy.asarray(mo[k], dtype=float)
    out['mo.sym'] = numpy.asarray(mo['sym'], dtype=str)
    out['mo.spin'] = numpy.asarray(mo['spin'], dtype=str)
    return out


def test_fchk_reader_equals_reference_reader(tmp_path):
    """read_gaussian_fchk == the reference's reader on the reference's own Gaussian test outputs: every flat QCinfo array
    identical (goldens written by running the reference: make_golden.py, make_golden_read.py)"""
    import io
    import os
    from conftest import load_golden, reader_input
    from orbkit_b200 import read, options
    options.quiet = True
    inputs = str(tmp_path)
    for name in ('h2o_rhf_sph.fchk', 'h2o_uhf_sph.fchk', 'h2o_rhf_cart.fchk'):
        reader_input(name, inputs)
    extra = load_golden('read_fchk')
    cases = [(load_golden('h2o_gaussian_sph'), '', 'h2o_rhf_sph.fchk', dict(all_mo=True)),
             (load_golden('h2o_gaussian_sph_occ'), '', 'h2o_rhf_sph.fchk', dict(all_mo=False)),
             (load_golden('h2o_gaussian_uhf'), '', 'h2o_uhf_sph.fchk', dict(all_mo=True)),
             (extra, 'cart.', 'h2o_rhf_cart.fchk', dict(all_mo=True)),
             (extra, 'uhf_beta.', 'h2o_uhf_sph.fchk', dict(all_mo=True, spin='beta')),
             (extra, 'uhf_occ.', 'h2o_uhf_sph.fchk', dict(all_mo=False))]
    for g, prefix, fn, kw in cases:
        qc = read.main_read(os.path.join(inputs, fn), **kw)
        for k, v in _flat_qc(qc).items():
            ref = g[prefix + k]
            assert v.shape == ref.shape and (v == ref).all(), (fn, kw, k)
        if prefix:
            assert qc.etot == float(g[prefix + 'etot'])
    # file objects (text and binary), explicit itype, error conventions
    with open(os.path.join(inputs, 'h2o_rhf_sph.fchk'), 'rb') as f:
        qb = read.main_read(io.BytesIO(f.read()), itype='fchk', all_mo=True)
    assert qb == read.read_gaussian_fchk(os.path.join(inputs, 'h2o_rhf_sph.fchk'), all_mo=True)
    with pytest.raises(IOError):
        read.main_read(os.path.join(inputs, 'h2o_rhf_sph.fchk'), spin='alpha')        # restricted file
    with pytest.raises(IOError):
        read.main_read(os.path.join(inputs, 'h2o_uhf_sph.fchk'), spin='gamma')
    with pytest.raises(NotImplementedError):
        read.main_read(os.path.join(inputs, 'h2o_rhf_sph.fchk'), itype='native')    # the reference's npz / hdf5 containers


def test_gaussian_log_reader_equals_reference_reader(tmp_path):
    """read_gaussian_log == the reference's reader on its four Gaussian .log test outputs (GFINPUT + POP=FULL): restricted /
    unrestricted, pure-spherical (the (l, m) labels come from the coefficient rows) / Cartesian, all / occupied orbitals"""
    import os
    from conftest import load_golden, reader_input
    from orbkit_b200 import read, options
    options.quiet = True
    inputs = str(tmp_path)
    g = load_golden('read_glog')
    for name, fn, kw in [('rhf_sph', 'h2o_rhf_sph.inp.log', dict(all_mo=True)),
                         ('uhf_cart', 'h2o_uhf_cart.inp.log', dict(all_mo=True)),
                         ('uhf_sph_occ', 'h2o_uhf_sph.inp.log', dict(all_mo=False)),
                         ('rhf_cart_occ', 'h2o_rhf_cart.inp.log', dict(all_mo=False))]:
        path = reader_input(fn, inputs)
        assert read.find_itype(path) == 'gaussian_log'
        qc = read.main_read(path, **kw)
        for k, v in _flat_qc(qc).items():
            ref = g[name + '.' + k]
            assert v.shape == ref.shape and (v == ref).all(), (fn, kw, k)
        assert qc.etot == float(g[name + '.etot'])
    # the fchk file of the same calculation holds the same basis and (to the printed digits) the same orbitals
    fq = read.main_read(reader_input('h2o_rhf_sph.fchk', inputs), all_mo=True)
    lq = read.main_read(os.path.join(inputs, 'h2o_rhf_sph.inp.log'), all_mo=True)
    assert numpy.abs(numpy.abs(fq.mo_spec.get_coeffs()) - numpy.abs(lq.mo_spec.get_coeffs())).max() < 1e-5
    with pytest.raises(IOError):                               # like the reference: no per-orbital spin with symmetries
        read.main_read(os.path.join(inputs, 'h2o_uhf_sph.inp.log'), spin='beta')
    with pytest.raises(IOError):
        read.main_read(os.path.join(inputs, 'h2o_uhf_sph.inp.log'), spin='gamma')
    with pytest.raises(IOError):
        read.read_gaussian_log(reader_input('nh3.mold', inputs))                  # no `Entering Link 1`


def test_natural_orbitals_of_the_pair_matrix():
    """detci.ci_core.natural_orbitals: sum_t c_t phi_a phi_b == sum_k lam_k psi_k^2 (the identity behind rho_from_qc's single
    fused launch), rank reduction, the antisymmetric case"""
    from orbkit_b200.detci.ci_core import natural_orbitals, flatten_terms
    rng = numpy.random.default_rng(31)
    n, nt = 9, 400
    terms = (rng.normal(size=nt), rng.integers(0, n, size=nt).astype(numpy.intc), rng.integers(0, n, size=nt).astype(numpy.intc))
    phi = rng.normal(size=(n, 57))
    ref = numpy.einsum('t,tx,tx->x', terms[0], phi[terms[1]], phi[terms[2]])
    lam, u = natural_orbitals(terms, n)
    got = numpy.einsum('k,kx->x', lam, (u.T @ phi) ** 2)
    assert lam.shape == (n,) and u.shape == (n, n)
    assert numpy.abs(got - ref).max() <= 1e-12 * numpy.abs(ref).max()
    # a transition between two determinants differing in one orbital: rank 2 (eigenvalues +c/2, -c/2)
    lam, u = natural_orbitals((numpy.array([0.7]), numpy.array([2]), numpy.array([5])), 8)
    assert lam.shape == (2,) and numpy.allclose(sorted(lam), [-0.35, 0.35])
    # antisymmetric pair matrix: no density at all
    lam, u = natural_orbitals((numpy.array([1.0, -1.0]), numpy.array([1, 3]), numpy.array([3, 1])), 4)
    assert lam.shape == (0,) and u.shape == (4, 0)
    # the host side of rho_from_qc on a stand-in engine (coefficients of the natural orbitals, signed weights)
    import types
    from orbkit_b200.detci import ci_core
    n_mo, n_ao, npts = 12, 7, 33
    C, chi = rng.normal(size=(n_mo, n_ao)), rng.normal(size=(n_ao, npts))
    qc = types.SimpleNamespace(mo_spec=types.SimpleNamespace(get_coeffs=lambda: C))

    class Eng:
        def mos(self, basis, coeffs, occ):
            return coeffs, occ

        def eval_rho(self, mo, g, codes):
            return numpy.einsum('k,kx->x', mo[1], (mo[0] @ chi) ** 2), None, None
    act = numpy.array([1, 4, 5, 9])
    ia, ib, c = rng.integers(0, 4, size=50).astype(numpy.intc), rng.integers(0, 4, size=50).astype(numpy.intc), rng.normal(size=50)
    got = ci_core._rho_natural(Eng(), None, qc, act, (c, ia, ib), types.SimpleNamespace(npts=npts))
    mo_values = C @ chi
    ref = numpy.einsum('t,tx,tx->x', c, mo_values[act[ia]], mo_values[act[ib]])
    assert numpy.abs(got - ref).max() <= 1e-12 * numpy.abs(ref).max()
    # arrays in, no list traffic: same flat terms as the list form
    zero, sing = [[], []], [[0.5, -1.5], [[0, 1], [2, 2]]]
    a = flatten_terms(zero, sing)
    b = flatten_terms(zero, [numpy.array(sing[0]), numpy.array(sing[1])])
    assert all((x == y).all() and x.dtype == y.dtype for x, y in zip(a, b))


def test_cclib_bridge_equals_reference_conversion():
    """read_cclib.convert_cclib == the reference's convert_cclib (read/cclib_parser.py:56-219) on cclib-shaped inputs
    (tests/cclib_cases.py; golden written by running the reference, make_golden_cclib.py): Cartesian / spherical AO labels,
    restricted / unrestricted / restricted open shell, natural orbitals, data without `aonames`"""
    import cclib_cases as cases
    from conftest import load_golden
    from orbkit_b200 import read, options
    from orbkit_b200.read_cclib import convert_cclib
    options.quiet = True
    g = load_golden('read_cclib')
    for name, kw in cases.CASES:
        k = cases.key(name, kw)
        qc = convert_cclib(cases.case(name), **kw)
        flat = _flat_qc(qc)
        for kk, v in flat.items():
            ref = g[k + '.' + kk]
            if v.dtype.kind == 'f':
                assert v.shape == ref.shape and (v == ref).all(), (k, kk)
            else:
                assert v.shape == ref.shape and (v == ref).all(), (k, kk)
    # the reference raises IOError for `spin` on restricted data ...
    assert str(g['rhf_cart_all_alpha.error']) == 'OSError'
    with pytest.raises(IOError):
        convert_cclib(cases.case('rhf_cart'), all_mo=True, spin='alpha')
    with pytest.raises(IOError):
        convert_cclib(cases.case('uhf_sph'), spin='gamma')
    # ... and crashes on a typo for unrestricted data (AttributeError, cclib_parser.py:157); here the selection it was
    # about to make is carried out: the orbitals of that spin of the all-orbital conversion, in order
    assert str(g['uhf_sph_all_beta.error']) == 'AttributeError'
    both = convert_cclib(cases.case('uhf_sph'), all_mo=True)
    beta = convert_cclib(cases.case('uhf_sph'), all_mo=True, spin='beta')
    pick = [i for i, s in enumerate(both.mo_spec.get_spin()) if s == 'beta'] if hasattr(both.mo_spec, 'get_spin') else \
        [i for i, mo in enumerate(both.mo_spec) if mo['spin'] == 'beta']
    assert len(beta.mo_spec) == len(pick) == 23
    assert (beta.mo_spec.get_coeffs() == both.mo_spec.get_coeffs()[pick]).all()
    assert [mo['sym'] for mo in beta.mo_spec] == [both.mo_spec[i]['sym'] for i in pick]
    # natural orbitals need their occupation numbers; main_read routes itype='cclib' to the parser front end
    cc = cases.case('natorb')
    del cc.nooccnos
    with pytest.raises(IOError):
        convert_cclib(cc)
    with pytest.raises(IOError):
        read.main_read('some.log', itype='cclib')                       # no parser named
    try:
        import cclib  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError):
            read.main_read('some.log', itype='cclib', cclib_parser='Gaussian')


def test_wfn_and_wfx_readers_equal_reference_readers(tmp_path):
    """read_wfn / read_wfx == the reference's readers on its GAMESS .wfn and ORCA .wfx test outputs: every flat QCinfo array
    identical (primitive-based inputs: pnum = -1, explicit lxlylz rows; goldens written by running the reference)"""
    import io
    import os
    from conftest import load_golden, reader_input
    from orbkit_b200 import read, options
    options.quiet = True
    inputs = str(tmp_path)
    wfn, wfx = reader_input('water_gamess-us.wfn', inputs), reader_input('1.wfx', inputs)
    cases = [(load_golden('water_gamess_wfn'), '', wfn, dict(all_mo=True)),
             (load_golden('h2o_orca_wfx'), '', wfx, dict(all_mo=True)),
             (load_golden('read_wf'), 'wfx_beta.', wfx, dict(all_mo=True, spin='beta'))]
    for g, prefix, path, kw in cases:
        qc = read.main_read(path, **kw)
        for k, v in _flat_qc(qc).items():
            ref = g[prefix + k]
            assert v.shape == ref.shape and (v == ref).all(), (path, kw, k)
        assert qc.ao_spec.get_normalized() and not qc.ao_spec.spherical
    with open(wfx, 'rb') as f:
        assert read.main_read(io.BytesIO(f.read()), itype='wfx', all_mo=True) == read.main_read(wfx)
    with pytest.raises(IOError):
        read.main_read(wfn, spin='alpha')                    # not supported by the .wfn reader
    with pytest.raises(IOError):
        read.main_read(wfx, spin='gamma')
    with pytest.raises(IOError):
        read.read_wfx(io.StringIO('<Keywords>\n STO\n</Keywords>\n'))


def test_gamess_and_aomix_readers_equal_reference_readers(tmp_path):
    """read_gamess / read_aomix == the reference's readers on its GAMESS-US (formaldehyde, f functions, explicit lxlylz) and
    Turbomole tm2aomix test outputs; the file type is found from the content like the reference does (magic strings)"""
    import os
    from conftest import load_golden, reader_input
    from orbkit_b200 import read, options
    options.quiet = True
    inputs = str(tmp_path)
    log, aomix = reader_input('formaldehyde.log', inputs), reader_input('aomix.in', inputs)
    assert read.find_itype(log) == 'gamess' and read.find_itype(aomix) == 'aomix'
    assert read.find_itype(reader_input('nh3.mold', inputs)) == 'molden'
    for g, path in ((load_golden('formaldehyde_gamess'), log), (load_golden('h2o_turbomole_aomix'), aomix)):
        qc = read.main_read(path, all_mo=True)
        for k, v in _flat_qc(qc).items():
            assert v.shape == g[k].shape and (v == g[k]).all(), (path, k)
    occ = read.main_read(log)                                  # all_mo=False: occupied orbitals only
    assert len(occ.mo_spec) == 8 and (occ.mo_spec.get_occ() == 2.0).all()
    with pytest.raises(IOError):
        read.main_read(log, spin='alpha')                      # restricted calculation
    with pytest.raises(NotImplementedError):
        read.read_gamess(log, read_properties=True)
    with pytest.raises(IOError):
        read.read_aomix(log)                                   # no [AOMix Format] keyword


def test_molden_reader_equals_reference_reader(tmp_path):
    """read_molden == the reference's reader on its Molpro and Psi4 test outputs.  Molpro files: every flat QCinfo array
    identical.  The Psi4 file triggers the renormalisation of the contractions (molden.py:375-398), whose self-overlaps
    are taken from a closed form here instead of the reference's recursion: primitive coefficients agree to 1 ulp."""
    import os
    from conftest import load_golden, reader_input
    from orbkit_b200 import read, options
    options.quiet = True
    inputs = str(tmp_path)
    for name in ('h2o_rhf_sph.molden', 'nh3.mold', 'lih_cis_aug-cc-pVTZ.out.default.molden', 'h2o_rhf_sph.fchk'):
        reader_input(name, inputs)
    for fix, fn, exact in [('h2o_molpro_cart', 'h2o_rhf_sph.molden', True), ('nh3_molpro', 'nh3.mold', True),
                           ('lih_psi4_sph_f', 'lih_cis_aug-cc-pVTZ.out.default.molden', False)]:
        g = load_golden(fix)
        qc = read.main_read(os.path.join(inputs, fn), all_mo=True)
        for k, v in _flat_qc(qc).items():
            ref = g[k]
            assert v.shape == ref.shape, (fn, k)
            if exact or v.dtype.kind != 'f':
                assert (v == ref).all(), (fn, k)
            else:
                assert numpy.allclose(v, ref, rtol=1e-14, atol=0.0), (fn, k)
    occ = read.main_read(os.path.join(inputs, 'h2o_rhf_sph.molden'))
    assert len(occ.mo_spec) == 5 and (occ.mo_spec.get_occ() == 2.0).all()
    norms = read.cartesian_self_overlap(occ.ao_spec)
    assert numpy.abs(norms - 1.0).max() < 1e-5            # Molpro writes normalised contractions
    with pytest.raises(IOError):
        read.main_read(os.path.join(inputs, 'h2o_rhf_sph.molden'), spin='alpha')
    with pytest.raises(IOError):
        read.read_molden(os.path.join(inputs, 'h2o_rhf_sph.fchk'))


# ---- streaming result store (orbkit_b200.store) ---------------------------------------------------------------------
def test_result_store_streams_slabs_into_an_npz(tmp_path):
    """save_hdf5 sink (core.py:478-501): column slabs written out of order land at their place; the finished file is
    a plain .npz (numpy.load reads it) with the reference's dataset names, and arrays() maps it without a host copy"""
    from orbkit_b200 import store
    rng = numpy.random.default_rng(3)
    N = (7, 5, 11)
    npts = int(numpy.prod(N))
    mo = rng.normal(size=(2, 3, npts))
    rho = rng.normal(size=npts)
    st = store.ResultStore(str(tmp_path / 'out'))            # no .h5 / .npz suffix -> out.npz
    st.put('grid/x', numpy.arange(7.0))
    st.put('grid/is_vector', False)
    d_mo = st.create('mo_list', (2, 3) + N, npts)
    d_rho = st.create('rho', N, npts)
    bounds = [0, 40, 41, 200, npts]
    for a, b in reversed(list(zip(bounds[:-1], bounds[1:]))):
        st.write_async(d_mo, a, b, numpy.ascontiguousarray(mo[:, :, a:b]))
        st.write_async(d_rho, a, b, rho[a:b].copy())
    st.close()
    assert st.path.endswith('out.npz') and not os.path.exists(st.path + '.parts')
    with numpy.load(st.path) as f:
        assert sorted(f.files) == ['grid/is_vector', 'grid/x', 'mo_list', 'rho']
        assert numpy.array_equal(f['mo_list'], mo.reshape((2, 3) + N)) and numpy.array_equal(f['rho'], rho.reshape(N))
        assert numpy.array_equal(f['grid/x'], numpy.arange(7.0)) and not f['grid/is_vector']
    arr = st.arrays()
    assert isinstance(arr['mo_list'], numpy.memmap) and arr['mo_list'].shape == (2, 3) + N
    assert numpy.array_equal(arr['mo_list'], mo.reshape((2, 3) + N)) and numpy.array_equal(arr['rho'], rho.reshape(N))
    # containers of main_output: groups become member prefixes, None is skipped, dictionaries nest
    fn = store.npz_write(str(tmp_path / 'c'), gname='g', data=mo, skip=None, grid={'x': numpy.arange(3.0), 'is_vector': True})
    with numpy.load(fn) as f:
        assert sorted(f.files) == ['g/data', 'g/grid/is_vector', 'g/grid/x'] and numpy.array_equal(f['g/data'], mo)
    assert store._lead_dims((3, 5, 4, 4, 4), 64) == (3, 5) and store._lead_dims((64,), 64) == ()
    if not store.have_h5py():
        with pytest.raises(ImportError):
            store.hdf5_write(str(tmp_path / 'x.h5'), data=mo)


def test_main_output_npz_container(tmp_path):
    """main_output(otype='npz') (output/high_level.py:306-311, output/hdf5.py:10-54): data, grid and qcinfo groups"""
    import orbkit_b200 as ok
    from conftest import golden_qc
    ok.options.quiet = True
    qc, a = golden_qc('h2o_gaussian_sph')
    ok.grid.set_grid(numpy.arange(3.0), numpy.arange(4.0), numpy.arange(5.0), is_vector=False)
    data = numpy.random.default_rng(0).normal(size=(2, 3, 4, 5))
    out = ok.main_output(data, qc, outputname=str(tmp_path / 'res'), otype='npz', gname='run1')
    assert out == [str(tmp_path / 'res') + '.npz']
    with numpy.load(out[0]) as f:
        assert numpy.array_equal(f['run1/data'], data)
        assert numpy.array_equal(f['run1/grid/y'], numpy.arange(4.0)) and not f['run1/grid/is_vector']
        assert numpy.allclose(f['run1/qcinfo/geo_spec'], qc.geo_spec)
        assert numpy.array_equal(f['run1/qcinfo/mo_spec/coeffs'], qc.mo_spec.get_coeffs())
        assert numpy.array_equal(f['run1/qcinfo/ao_spec/_lxlylz'], qc.ao_spec.get_lxlylz())
        assert 'run1/qcinfo/date' in f.files
    # 'auto' takes the type from the file name; cube and npz together
    both = ok.main_output(data[0, 0], qc, outputname=str(tmp_path / 'b'), otype=['npz'])
    assert both == [str(tmp_path / 'b') + '.npz']
    from orbkit_b200 import store
    if not store.have_h5py():
        with pytest.raises(ImportError):
            ok.main_output(data, qc, outputname=str(tmp_path / 'h'), otype='h5')
    with pytest.raises(NotImplementedError):
        ok.main_output(data, qc, outputname=str(tmp_path / 'h'), otype='am')
