"""Parity of the CUDA path (through the public API and the C ABI) against
   (1) the outputs the reference itself wrote into tests/golden/ (make_golden.py),
   (2) the reference's own golden refdata_rho_compute.npz and the Gaussian cubegen KATs,
   (3) the pinned CPU oracle on seeded random inputs and edge-case sizes,
   (4) size-independent properties at large sizes.
Tolerance (FP64 path): |d| <= 1e-10*|ref| + 1e-14*max|ref|; electron counts to 1e-8."""
import numpy
import pytest

from conftest import DRV10, FIXTURES, assert_close, golden_qc, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ok():
    import orbkit_b200
    orbkit_b200.options.quiet = True
    return orbkit_b200


def set_regular(ok, x, y, z):
    ok.grid.set_grid(numpy.array(x), numpy.array(y), numpy.array(z), is_vector=False)


def set_vector(ok, x, y, z):
    ok.grid.set_grid(numpy.array(x), numpy.array(y), numpy.array(z), is_vector=True)


def test_reference_golden_refdata(ok):
    """orbkit/test/grid_based/rho_compute.py: rho, laplacian, 24 MOs x 10 derivative codes"""
    qc, _ = golden_qc('h2o_molpro_cart')
    g = load_golden('ref_rho_compute')
    for vector in (False, True):
        set_regular(ok, g['x'], g['y'], g['z'])
        if vector:
            ok.grid.grid2vector()
        shp = (-1,) if vector else (5, 9, 7)
        assert_close(ok.rho_compute(qc, slice_length=0), g['zero'].reshape(shp), 'rho')
        assert_close(ok.rho_compute(qc, numproc=2), g['zero'].reshape(shp), 'rho numproc')
        assert_close(ok.rho_compute(qc, laplacian=True)[-1], g['two'].reshape(shp), 'laplacian')
        mo = ok.rho_compute(qc, calc_mo=True, drv=DRV10)
        assert_close(mo, g['four'].reshape((10, 24) + ((315,) if vector else (5, 9, 7))), 'mo10')
    # the grid module state is restored like the reference does (core.py:433,569)
    assert ok.grid.is_vector


def test_gaussian_cubegen_kat(ok):
    """orbkit/test/grid_based/cube_files.py with the reference's own tolerances"""
    k = load_golden('cube_kat')
    qc, _ = golden_qc('h2o_gaussian_sph_occ')
    eq = lambda a, b, tol=1e-5: numpy.allclose(a, b, rtol=tol * 1e2, atol=tol)
    set_regular(ok, k['x'], k['y'], k['z'])
    rho, drho = ok.rho_compute(qc, drv='xyz')
    rho2, d2, lap = ok.rho_compute(qc, laplacian=True)
    assert eq(rho, rho2) and eq(rho, k['rho']) and eq(drho, k['drho'], 1e-3) and eq(lap, k['laplacian'])
    qc.mo_spec = qc.mo_spec[1:]
    assert eq(ok.rho_compute(qc), k['rho_valence'])
    qc.mo_spec = qc.mo_spec['homo']
    assert eq(ok.rho_compute(qc, calc_mo=True), k['homo'])


@pytest.mark.parametrize('name', FIXTURES)
def test_fixture_molecules_vs_reference_outputs(ok, name):
    qc, a = golden_qc(name)
    set_regular(ok, a['rx'], a['ry'], a['rz'])
    assert_close(ok.rho_compute(qc), a['reg.rho'], name + ' reg.rho')
    r, d = ok.rho_compute(qc, drv=['x', 'y', 'z'])
    assert_close(r, a['reg.rho'], name + ' reg.rho(drv)')
    assert_close(d, a['reg.drho'], name + ' reg.drho')
    r, d, l = ok.rho_compute(qc, laplacian=True)
    assert_close(d, a['reg.d2rho'], name + ' reg.d2rho')
    assert_close(l, a['reg.lap'], name + ' reg.lap', afloor=3e-14)
    set_vector(ok, a['vx'], a['vy'], a['vz'])
    assert_close(ok.rho_compute(qc), a['vec.rho'], name + ' vec.rho')
    r, d = ok.rho_compute(qc, drv=['x', 'y', 'z'])
    assert_close(d, a['vec.drho'], name + ' vec.drho')
    r, d, l = ok.rho_compute(qc, laplacian=True)
    assert_close(d, a['vec.d2rho'], name + ' vec.d2rho')
    r, d = ok.rho_compute(qc, drv=['xy', 'z', 'yz', 'x2'])
    assert_close(d, a['vec.dmixed'], name + ' vec.dmixed')
    if 'vec.mo10' in a.files:
        assert_close(ok.rho_compute(qc, calc_mo=True, drv=DRV10), a['vec.mo10'], name + ' mo10')
        assert_close(ok.rho_compute(qc, calc_ao=True, drv=DRV10), a['vec.ao10'], name + ' ao10')
        assert_close(ok.extras.calc_ao(qc, drv=['z', 'xy']), a['vec.ao10'][[3, 5]], name + ' calc_ao')
    else:
        assert_close(ok.rho_compute(qc, calc_mo=True), a['vec.mo'], name + ' mo')
        assert_close(ok.rho_compute(qc, calc_ao=True), a['vec.ao'], name + ' ao')
        assert_close(ok.rho_compute(qc, calc_mo=True, drv=['z'])[0], a['vec.mo_z'], name + ' mo_z')
        assert_close(ok.rho_compute(qc, calc_ao=True, drv=['yy'])[0], a['vec.ao_yy'], name + ' ao_yy')


@pytest.mark.parametrize('name', ['h2o_gaussian_sph', 'lih_psi4_sph_f', 'water_gamess_wfn'])
def test_operators_and_components(ok, oracle_mod, name):
    """core.ao_creator / mo_creator / cartesian2spherical / rho_compute_no_slice(return_components)"""
    qc, a = golden_qc(name)
    x, y, z = a['vx'], a['vy'], a['vz']
    for drv in (None, 'x', 'zz', 'xz', 2, 9):
        got = ok.core.ao_creator(qc.geo_spec, qc.ao_spec, drv=drv, x=x, y=y, z=z, is_vector=True)
        ref = oracle_mod.ao_creator(qc.geo_spec, qc.ao_spec, drv=drv, x=x, y=y, z=z, is_vector=True)
        assert_close(got, ref, '%s ao_creator drv=%s' % (name, drv))
    ao = oracle_mod.ao_creator(qc.geo_spec, qc.ao_spec, x=a['rx'], y=a['ry'], z=a['rz'], is_vector=False)
    got = ok.core.ao_creator(qc.geo_spec, qc.ao_spec, x=a['rx'], y=a['ry'], z=a['rz'], is_vector=False)
    assert got.shape == ao.shape == (qc.ao_spec.get_ao_num(), 4, 5, 3)
    assert_close(got, ao, name + ' ao_creator regular')
    assert_close(ok.core.mo_creator(ao, qc.mo_spec), oracle_mod.mo_creator(ao, qc.mo_spec), name + ' mo_creator')
    if qc.ao_spec.spherical:
        import copy
        cart_spec = copy.deepcopy(qc.ao_spec)
        cart_spec.spherical = False
        cart = oracle_mod.ao_creator(qc.geo_spec, cart_spec, x=x, y=y, z=z, is_vector=True)
        assert_close(ok.core.cartesian2spherical(cart, qc.ao_spec),
                     oracle_mod.cartesian2spherical(cart, qc.ao_spec), name + ' cart2sph')
    res = ok.rho_compute_no_slice(qc, drv=['x', 'y', 'z'], return_components=True, x=x, y=y, z=z, is_vector=True)
    ao_l, mo_l, rho, dao, dmo, drho = res
    assert_close(rho, a['vec.rho'], name + ' no_slice rho')
    assert_close(drho, a['vec.drho'], name + ' no_slice drho')
    assert_close(mo_l, a['vec.mo10'][0], name + ' no_slice mo')
    assert_close(dao, a['vec.ao10'][1:4], name + ' no_slice dao')
    assert_close(dmo, a['vec.mo10'][1:4], name + ' no_slice dmo')
    r, d, l = ok.rho_compute(qc, laplacian=True, numproc=0, x=x, y=y, z=z, is_vector=True)
    assert_close(d, a['vec.d2rho'], name + ' numproc=0 route')


def test_cy_core_dropins_random_shells(ok, oracle_mod):
    rng = numpy.random.default_rng(11)
    from orbkit_b200.tools import exp, exp_wfn
    port = oracle_mod.backend('port')
    for trial in range(6):
        L = list(rng.integers(0, 5, size=5))
        table = exp_wfn if trial % 2 else exp
        lxlylz = numpy.array(sum([table[l] for l in L], []), dtype=numpy.intc)
        assign = numpy.array([len(table[l]) for l in L], dtype=numpy.intc)
        pnum = numpy.array(rng.integers(1, 7, size=5), dtype=numpy.intc)
        coeffs = numpy.stack([10 ** rng.uniform(-1, 3.5, pnum.sum()), rng.uniform(-1, 1, pnum.sum())], axis=1)
        geo = rng.uniform(-2, 2, size=(3, 3))
        atoms = numpy.array(rng.integers(0, 3, size=5), dtype=numpy.intc)
        npts = [1, 31, 33, 127, 129, 1000][trial]
        x, y, z = rng.uniform(-3, 3, size=(3, npts))
        for normalized in (0, 1):
            for drv in range(10):
                got = ok.cy_core.aocreator(lxlylz, assign, coeffs, pnum, geo, atoms, x, y, z, drv, normalized)
                ref = port.aocreator(lxlylz, assign, coeffs, pnum, geo, atoms, x, y, z, drv, normalized)
                assert_close(got, ref, 'aocreator trial %d norm %d drv %d' % (trial, normalized, drv))
        for drv in (7, 8, 9):
            got = ok.cy_core.aocreator(lxlylz, assign, coeffs, pnum, geo, atoms, x, y, z, drv, 0, exact_mixed=True)
            ref = port.aocreator(lxlylz, assign, coeffs, pnum, geo, atoms, x, y, z, drv, 0, exact_mixed=1)
            assert_close(got, ref, 'aocreator exact mixed drv %d' % drv)
        C = rng.standard_normal((int(rng.integers(1, 70)), lxlylz.shape[0]))
        assert_close(ok.cy_core.mocreator(ref, C), port.mocreator(ref, C), 'mocreator')
        # lcreator writes the first ao_num rows in place at the array's own row stride
        big = numpy.full((assign[0] + 2, npts), -7.0)
        ok.cy_core.lcreator(big, lxlylz, coeffs, geo[atoms[0]].copy(), x, y, z, int(assign[0]), int(pnum[0]), 3, 0)
        one = port.aocreator(lxlylz[:assign[0]], assign[:1], coeffs[:pnum[0]], pnum[:1], geo, atoms[:1], x, y, z, 3, 0)
        assert_close(big[:assign[0]], one, 'lcreator')
        assert (big[assign[0]:] == -7.0).all()
    for (lx, ly, lz) in [(0, 0, 0), (1, 0, 0), (2, 1, 0), (1, 1, 1), (0, 0, 4)]:
        assert abs(ok.cy_core.aonorm(lx, ly, lz, 0.8, 0) / port.aonorm(lx, ly, lz, 0.8, 0) - 1) < 1e-15
        assert ok.cy_core.aonorm(lx, ly, lz, 0.8, 1) == 1.0
        for drv in range(10):
            r = port.aoxyz(0.3, -1.1, 0.6, lx, ly, lz, 1.3, drv)
            assert abs(ok.cy_core.aoxyz(0.3, -1.1, 0.6, lx, ly, lz, 1.3, drv) - r) <= 1e-13 * max(1, abs(r))


def test_edge_sizes_and_point_ranges(ok, oracle_mod):
    """empty grid, single point, ragged tails, regular == vector, results independent of ranges"""
    from orbkit_b200.engine import get_engine
    qc, a = golden_qc('synth_small_sph')
    rng = numpy.random.default_rng(5)
    set_vector(ok, numpy.zeros(0), numpy.zeros(0), numpy.zeros(0))
    assert ok.rho_compute(qc).shape == (0,)
    assert ok.rho_compute(qc, calc_mo=True).shape == (9, 0)
    for n in (1, 2, 63, 64, 65, 257, 4099):
        x, y, z = rng.uniform(-6, 6, size=(3, n))
        set_vector(ok, x, y, z)
        r, d = ok.rho_compute(qc, drv='xyz')
        rr, dr = oracle_mod.rho_compute(qc, x, y, z, is_vector=True, drv='xyz')
        assert_close(r, rr, 'rho n=%d' % n)
        assert_close(d, dr, 'drho n=%d' % n)
    ax, ay, az = numpy.linspace(-5, 5, 7), numpy.linspace(-4, 4, 6), numpy.linspace(-3, 3, 11)
    set_regular(ok, ax, ay, az)
    reg = ok.rho_compute(qc, laplacian=True)
    ok.grid.grid2vector()
    vec = ok.rho_compute(qc, laplacian=True)
    # regular grids take their exponentials from separable per-axis tables, vector grids evaluate
    # exp(-a r^2) per point: same values within the stated tolerance, and both match the oracle
    ref = oracle_mod.rho_compute(qc, ax, ay, az, is_vector=False, laplacian=True)
    for p, q, r in zip(reg, vec, ref):
        assert_close(p.reshape(q.shape), q, 'regular vs vector')
        assert_close(p, r.reshape(p.shape), 'regular vs oracle')
    # explicit sub-ranges through the engine == the full evaluation
    eng = get_engine()
    basis = eng.basis(qc.geo_spec, qc.ao_spec)
    mo = eng.mos(basis, qc.mo_spec.get_coeffs(), qc.mo_spec.get_occ())
    g = eng.grid_regular(ax, ay, az)
    full, dfull, nrm = eng.eval_rho(mo, g, [1, 2, 3], want_norm=True)
    parts = [eng.eval_rho(mo, g, [1, 2, 3], p0, p1, want_norm=True) for p0, p1 in ((0, 100), (100, 101), (101, 462))]
    assert numpy.array_equal(numpy.concatenate([p[0] for p in parts]), full)
    assert numpy.array_equal(numpy.concatenate([p[1] for p in parts], axis=1), dfull)
    assert_close(sum(p[2] for p in parts), nrm, 'mo_norm parts')
    mo_ref = oracle_mod.rho_compute(qc, ax, ay, az, is_vector=False, calc_mo=True)
    assert_close(nrm, (mo_ref.reshape(9, -1) ** 2).sum(axis=1), 'mo_norm', rtol=1e-12)


def test_documented_limits_of_the_chunk_tables(ok, oracle_mod):
    """DESIGN 7 'known limits': a contraction of up to 96 primitives and shells of up to 28 functions (L = 6, i shells --
    beyond every published basis set) are evaluated; 97 primitives or an L = 7 shell (36 functions) are refused with a
    message, never silently truncated."""
    from orbkit_b200._lib import OkbError
    rng = numpy.random.default_rng(12)
    port = oracle_mod.backend('port')
    geo = numpy.zeros((1, 3)); atoms = numpy.zeros(1, dtype=numpy.intc)
    x, y, z = rng.uniform(-2, 2, size=(3, 77))

    def shell(L):
        return numpy.array([(a, b, L - a - b) for a in range(L, -1, -1) for b in range(L - a, -1, -1)], dtype=numpy.intc)
    s_fn = shell(0)
    for npr, fine in ((96, True), (97, False)):
        coeffs = numpy.stack([10 ** rng.uniform(-1, 3, npr), rng.uniform(-1, 1, npr)], axis=1)
        args = (s_fn, numpy.array([1], dtype=numpy.intc), coeffs, numpy.array([npr], dtype=numpy.intc), geo, atoms, x, y, z)
        if fine:
            for drv in (0, 1, 4):
                assert_close(ok.cy_core.aocreator(*args, drv, 0), port.aocreator(*args, drv, 0), '96 primitives drv %d' % drv)
        else:
            with pytest.raises(OkbError, match='primitives'):
                ok.cy_core.aocreator(*args, 0, 0)
    coeffs = numpy.array([[0.7, 1.0], [0.2, 0.5]])
    for L, fine in ((6, True), (7, False)):
        fn = shell(L)
        args = (fn, numpy.array([len(fn)], dtype=numpy.intc), coeffs, numpy.array([2], dtype=numpy.intc), geo, atoms, x, y, z)
        if fine:
            for drv in (0, 2, 6):
                assert_close(ok.cy_core.aocreator(*args, drv, 0), port.aocreator(*args, drv, 0), 'L = 6 drv %d' % drv)
        else:
            with pytest.raises(OkbError, match='functions'):
                ok.cy_core.aocreator(*args, 0, 0)


def test_many_mos_multiple_mo_tiles(ok, oracle_mod):
    """n_mo larger than one MO tile (MC<=96): AO tiles are regenerated per MO tile"""
    from orbkit_b200 import synth
    spec = synth.make_molecule(n_heavy=2, n_light=3, n_mo=230, seed=21, spherical=True)
    qc = synth.to_qcinfo(spec)
    rng = numpy.random.default_rng(2)
    x, y, z = rng.uniform(-9, 9, size=(3, 300))
    set_vector(ok, x, y, z)
    r, d, l = ok.rho_compute(qc, laplacian=True)
    rr, dr, lr = oracle_mod.rho_compute(qc, x, y, z, is_vector=True, laplacian=True)
    assert_close(r, rr, 'rho 230 MOs')
    assert_close(d, dr, 'd2rho 230 MOs')
    mo = ok.rho_compute(qc, calc_mo=True, drv=['y'])
    assert_close(mo, oracle_mod.rho_compute(qc, x, y, z, is_vector=True, calc_mo=True, drv=['y']), 'mo_y 230')


def test_errors_match_reference_conventions(ok):
    qc, a = golden_qc('nh3_molpro')
    set_vector(ok, a['vx'], a['vy'], a['vz'])
    with pytest.raises(ValueError):
        ok.rho_compute(qc, calc_ao=True, calc_mo=True)
    with pytest.raises(ValueError):
        ok.rho_compute(qc, drv=['q'])
    with pytest.raises(ValueError):
        ok.core.ao_creator(qc.geo_spec, qc.ao_spec, x=a['vx'], y=a['vy'][:-1], z=a['vz'], is_vector=True)
    with pytest.raises(ValueError):
        ok.cy_core.mocreator(numpy.zeros((3, 4)), numpy.zeros((2, 5)))
    with pytest.raises(ValueError):
        ok.cy_core.mocreator(numpy.zeros((3, 4), dtype=numpy.float32), numpy.zeros((2, 3)))
    with pytest.raises(TypeError):
        ok.cy_core.mocreator(None, numpy.zeros((2, 3)))


def test_full_size_properties_config3(ok, oracle_mod):
    """BASELINE config 3 molecule (1000 AOs, 82 MOs) on a 48^3 cut of the benchmark box:
       subsample parity vs the oracle, additivity over MO subsets, gradient vs central differences,
       electron count consistency between independent evaluations."""
    qc, a = golden_qc('synth_c3')
    ax = numpy.linspace(-12, 12, 48)
    set_regular(ok, ax, ax, ax)
    rho, drho = ok.rho_compute(qc, drv=['x', 'y', 'z'])
    assert rho.shape == (48, 48, 48) and (rho >= 0).all()
    # parity on a strided subsample (the oracle needs seconds per 1e3 points at this size)
    idx = numpy.arange(0, 48 ** 3, 48 ** 3 // 96)[:96]
    X, Y, Z = numpy.meshgrid(ax, ax, ax, indexing='ij')
    xs, ys, zs = X.ravel()[idx], Y.ravel()[idx], Z.ravel()[idx]
    rr, dr = oracle_mod.rho_compute(qc, xs, ys, zs, is_vector=True, drv=['x', 'y', 'z'],
                                    kind='ref' if oracle_mod.have_ref() else 'port')
    assert_close(rho.ravel()[idx], rr, 'c3 rho subsample')
    assert_close(drho.reshape(3, -1)[:, idx], dr, 'c3 drho subsample')
    # additivity: rho(all MOs) = rho(first 40) + rho(rest)
    qa, qb = qc.copy(), qc.copy()
    qa.mo_spec = qc.mo_spec[:40]
    qb.mo_spec = qc.mo_spec[40:]
    ra, rb = ok.rho_compute(qa), ok.rho_compute(qb)
    assert_close(ra + rb, rho, 'additivity', rtol=1e-12)
    d3r = (ax[1] - ax[0]) ** 3
    assert abs((ra.sum() + rb.sum() - rho.sum()) * d3r) < 1e-8
    # gradient vs central differences of rho on the vector grid around a few points
    h = 1e-4
    p = numpy.stack([xs[:12], ys[:12], zs[:12]])
    for ax_i in range(3):
        e = numpy.zeros((3, 1)); e[ax_i] = h
        set_vector(ok, *(p + e))
        rp = ok.rho_compute(qc)
        set_vector(ok, *(p - e))
        rm = ok.rho_compute(qc)
        fd = (rp - rm) / (2 * h)
        assert numpy.allclose(fd, dr[ax_i, :12], rtol=1e-5, atol=1e-7 * numpy.abs(dr).max())


def test_config1_h2o_80cube_full_parity(ok, oracle_mod):
    """BASELINE configs[0] at full size: H2O RHF (Gaussian fchk, spherical d), rho and grad rho on the
    regular 80^3 grid over [-6,6]^3, every point against the reference's own objects."""
    qc, _ = golden_qc('h2o_gaussian_sph_occ')
    ax = numpy.linspace(-6.0, 6.0, 80)
    set_regular(ok, ax, ax, ax)
    rho, drho = ok.rho_compute(qc, drv=['x', 'y', 'z'])
    kind = 'ref' if oracle_mod.have_ref() else 'port'
    r_ref, d_ref = oracle_mod.rho_compute(qc, ax, ax, ax, is_vector=False, drv=['x', 'y', 'z'], numproc=4,
                                          slice_length=20000, kind=kind)
    assert rho.shape == (80, 80, 80)
    assert_close(rho, r_ref, 'C1 rho 80^3')
    assert_close(drho, d_ref, 'C1 grad rho 80^3')
    d3r = (ax[1] - ax[0]) ** 3
    assert abs(rho.sum() * d3r - r_ref.sum() * d3r) < 1e-8          # electron count
    assert abs(rho.sum() * d3r - 10.0) < 0.15                       # 10 electrons; the O 1s cusp is under-resolved at 0.15 bohr


def test_config1_end_to_end_fchk_to_cube(ok, oracle_mod, tmp_path):
    """BASELINE configs[0] from the file to the file: main_read(h2o_rhf_sph.fchk) -> grid_init 80^3 on [-6,6]^3 ->
    rho_compute -> main_output(otype='cb'); the cube file equals the reference's text of the oracle's density wherever
    the two densities round to the same six digits (everywhere but at a handful of exact rounding boundaries)"""
    import os
    import oracle_out
    from conftest import reader_input
    qc = ok.main_read(reader_input('h2o_rhf_sph.fchk', tmp_path), all_mo=False)
    ok.grid.min_, ok.grid.max_, ok.grid.N_ = [-6.0] * 3, [6.0] * 3, [80] * 3
    ok.grid.delta_ = [0, 0, 0]
    ok.grid.is_initialized = False
    ok.grid.grid_init()
    rho = ok.rho_compute(qc)
    fn = ok.main_output(rho, qc, outputname=str(tmp_path / 'h2o_rho'), otype='cb', datalabels='rho')[0]
    ax = numpy.linspace(-6.0, 6.0, 80)
    kind = 'ref' if oracle_mod.have_ref() else 'port'
    r_ref = oracle_mod.rho_compute(qc, ax, ax, ax, is_vector=False, numproc=4, slice_length=20000, kind=kind)
    assert_close(rho, r_ref, 'C1 rho from the fchk file')
    got = open(fn, 'rb').read()
    # the file is exactly the reference's text of the density that was computed here ...
    head = ok.output.cube_header(1, qc.geo_info, qc.geo_spec, comments='rho').encode()
    assert got[:len(head)] == head
    sample = rho[::9, ::7]                                     # rows (x, y) of the file: 13*80 + 13 + 1 bytes each
    rb = 13 * 80 + 80 // 6 + 1
    for ix, x in enumerate(range(0, 80, 9)):
        for iy, y in enumerate(range(0, 80, 7)):
            off = len(head) + (x * 80 + y) * rb
            assert got[off:off + rb] == oracle_out.cube_body(sample[ix:ix + 1, iy:iy + 1]), (x, y)
    # ... and its numbers are the oracle's density to the six printed digits
    vals = numpy.array(got[len(head):].split(), dtype=float).reshape(80, 80, 80)
    assert numpy.abs(vals - r_ref).max() <= 5.001e-6 * numpy.abs(r_ref).max()
    assert numpy.all(numpy.abs(vals - r_ref) <= 5.001e-6 * numpy.maximum(numpy.abs(r_ref), 1e-300) + 1e-99)


def test_benchmark_size_properties_200cube(ok, oracle_mod):
    """The benchmark workload itself (1000 AOs, 82 MOs, 200^3 points): parity on a random sub-sample
    against the reference objects, and size-independent properties of the full result."""
    from orbkit_b200.engine import get_engine
    qc, _ = golden_qc('synth_c3')
    ax = numpy.linspace(-12.0, 12.0, 200)
    set_regular(ok, ax, ax, ax)
    rho, drho = ok.rho_compute(qc, drv=['x', 'y', 'z'])
    assert rho.shape == (200, 200, 200) and numpy.isfinite(rho).all() and (rho >= 0).all()
    rng = numpy.random.default_rng(9)
    idx = rng.choice(200 ** 3, size=64, replace=False)
    i, rem = numpy.divmod(idx, 200 * 200)
    j, k = numpy.divmod(rem, 200)
    kind = 'ref' if oracle_mod.have_ref() else 'port'
    rr, dr = oracle_mod.rho_compute(qc, ax[i], ax[j], ax[k], is_vector=True, drv=['x', 'y', 'z'], kind=kind)
    assert_close(rho.ravel()[idx], rr, 'C3 rho sample')
    assert_close(drho.reshape(3, -1)[:, idx], dr, 'C3 grad rho sample')
    # (1) checksum of checksums: eight disjoint point ranges reproduce the full evaluation bit for bit
    eng = get_engine()
    basis = eng.basis(qc.geo_spec, qc.ao_spec)
    mo = eng.mos(basis, qc.mo_spec.get_coeffs(), qc.mo_spec.get_occ())
    g = eng.grid_regular(ax, ax, ax)
    bounds = numpy.linspace(0, 200 ** 3, 9).astype(int)
    bounds[1:-1] += [3, -5, 17, 0, 31, -1, 64]                      # ragged, not tile aligned
    norm_sum = numpy.zeros(82)
    for a, b in zip(bounds[:-1], bounds[1:]):
        r, d, n = eng.eval_rho(mo, g, [1, 2, 3], int(a), int(b), want_norm=True)
        assert numpy.array_equal(r, rho.ravel()[a:b])
        assert numpy.array_equal(d, drho.reshape(3, -1)[:, a:b])
        norm_sum += n
    # (2) two independent accumulation paths: sum_i occ_i * (sum_p phi_i^2)  ==  sum_p rho
    occ = qc.mo_spec.get_occ()
    assert abs((occ * norm_sum).sum() / rho.sum() - 1.0) < 1e-11


def test_product_grids_generated_on_the_device(ok, oracle_mod):
    """grid.sph2cart_vector / cyl2cart_vector (+ grid_sym_op / grid_translate): the kernels generate the Cartesian
    coordinates from the three axis vectors.  Untransformed grids reproduce the reference's coordinates bit for
    bit, so the results must be IDENTICAL to the vector-grid evaluation of those coordinates; transformed grids and
    the oracle comparison use the stated tolerance."""
    qc, _ = golden_qc('lih_psi4_sph_f')
    rng = numpy.random.default_rng(12)
    r = numpy.sort(rng.uniform(0.05, 4.0, 9))
    th = numpy.linspace(0.0, numpy.pi, 7)
    ph = numpy.linspace(0.0, 2 * numpy.pi, 11)
    zed = numpy.linspace(-2.0, 2.5, 5)
    for kind in ('sph', 'cyl'):
        if kind == 'sph':
            ok.grid.sph2cart_vector(r, th, ph)
            xyz = oracle_mod.sph2cart(r, th, ph)
        else:
            ok.grid.cyl2cart_vector(r, ph, zed)
            xyz = oracle_mod.cyl2cart(r, ph, zed)
        assert ok.grid.product_grid() is not None
        rho, drho = ok.rho_compute(qc, drv='xyz')
        mo = ok.rho_compute(qc, calc_mo=True)
        ao = ok.rho_compute(qc, calc_ao=True, drv=['zz'])
        assert ok.grid.product_grid() is not None          # still a recipe: nobody asked for grid.x
        assert rho.shape == (xyz.shape[1],) and mo.shape == (len(qc.mo_spec), xyz.shape[1])
        ok.grid.set_grid(xyz[0], xyz[1], xyz[2], is_vector=True)
        r2, d2 = ok.rho_compute(qc, drv='xyz')
        assert numpy.array_equal(rho, r2) and numpy.array_equal(drho, d2)
        assert numpy.array_equal(mo, ok.rho_compute(qc, calc_mo=True))
        assert numpy.array_equal(ao, ok.rho_compute(qc, calc_ao=True, drv=['zz']))
        rr, dr = oracle_mod.rho_compute(qc, xyz[0], xyz[1], xyz[2], is_vector=True, drv='xyz')
        assert_close(rho, rr, kind + ' rho')
        assert_close(drho, dr, kind + ' drho')
    # symmetry operation + translation on the recipe == the reference's host transformation of the coordinates
    ok.grid.sph2cart_vector(r, th, ph)
    S = ok.grid.rot(0.7, 1)
    ok.grid.grid_sym_op(S)
    ok.grid.grid_translate(0.25, -0.5, 1.0)
    rho = ok.rho_compute(qc)
    assert ok.grid.product_grid() is not None
    xyz = numpy.dot(S, oracle_mod.sph2cart(r, th, ph)) + numpy.array([[0.25], [-0.5], [1.0]])
    assert_close(rho, oracle_mod.rho_compute(qc, xyz[0], xyz[1], xyz[2], is_vector=True), 'sym op', rtol=1e-9)
    assert numpy.allclose(numpy.array([ok.grid.x, ok.grid.y, ok.grid.z]), xyz, rtol=0, atol=1e-14)
    assert ok.grid.product_grid() is None


def test_laplacian_phi_cache_subranges(ok, oracle_mod, monkeypatch):
    """rho + laplacian runs as SET_GRAD (MO values left in HBM) + SET_D2P (three second-derivative sets): the same
    results with the scratch cut into several sub-ranges, with ragged point counts, and as the uncached two-pass form"""
    from orbkit_b200 import synth
    spec = synth.make_molecule(n_heavy=3, n_light=3, n_mo=101, seed=4, spherical=True)
    qc = synth.to_qcinfo(spec)
    rng = numpy.random.default_rng(8)
    x, y, z = rng.uniform(-8, 8, size=(3, 5003))
    set_vector(ok, x, y, z)
    rr, dr, lr = oracle_mod.rho_compute(qc, x, y, z, is_vector=True, laplacian=True)
    r, d, l = ok.rho_compute(qc, laplacian=True)
    assert ok.engine.get_engine().last_kernel().startswith('ws-dmma/SET_D2P/')
    assert_close(r, rr, 'rho')
    assert_close(d, dr, 'd2rho')
    assert_close(l, lr, 'laplacian', afloor=3e-14)
    monkeypatch.setenv('OKB_PHI_CACHE_PTS', '2048')          # 3 sub-ranges: 2048 + 2048 + 907
    r2, d2, l2 = ok.rho_compute(qc, laplacian=True)
    monkeypatch.delenv('OKB_PHI_CACHE_PTS')
    assert numpy.array_equal(r, r2) and numpy.array_equal(d, d2) and numpy.array_equal(l, l2)
    # the MO norms of the first pass are not disturbed by the second one
    eng = ok.engine.get_engine()
    basis = eng.basis(qc.geo_spec, qc.ao_spec)
    mo = eng.mos(basis, qc.mo_spec.get_coeffs(), qc.mo_spec.get_occ())
    g = eng.grid_vector(x, y, z)
    _, _, nrm = eng.eval_rho(mo, g, [4, 5, 6], want_norm=True)
    mo_ref = oracle_mod.rho_compute(qc, x, y, z, is_vector=True, calc_mo=True)
    assert_close(nrm, (mo_ref ** 2).sum(axis=1), 'mo_norm', rtol=1e-12)


@pytest.mark.parametrize('name', ['h2o_gaussian_sph', 'lih_psi4_sph_f', 'water_gamess_wfn', 'synth_small_cart_g',
                                  'synth_small_sph', 'h2o_orca_wfx'])
def test_calc_ao_zrun_kernel_regular_grids(ok, oracle_mod, name):
    """calc_ao values on REGULAR grids run the z-run kernel (csrc/okb_ao_zrun.cuh: separable exponentials, one z point
    per thread, polynomial in Z per output row): against the oracle, against the exponential-per-point kernel on the
    same points as a vector grid, ragged axis lengths (z runs shorter than a warp, longer than a CTA), and point
    sub-ranges that start and end inside a z run"""
    from orbkit_b200.engine import get_engine
    qc, a = golden_qc(name)
    eng = get_engine()
    rng = numpy.random.default_rng(17)
    for shape in ((3, 4, 5), (2, 3, 37), (3, 2, 200), (1, 2, 300), (5, 1, 1)):
        ax = [numpy.sort(rng.uniform(-3.5, 3.5, n)) for n in shape]
        set_regular(ok, *ax)
        got = ok.rho_compute(qc, calc_ao=True)
        assert eng.last_kernel().startswith('zrun/'), eng.last_kernel()
        assert got.shape == (qc.ao_spec.get_ao_num(),) + shape
        ref = oracle_mod.ao_creator(qc.geo_spec, qc.ao_spec, x=ax[0], y=ax[1], z=ax[2], is_vector=False)
        assert_close(got, ref, '%s zrun %s' % (name, shape))
        # the same points as a vector grid (exponential per point)
        X, Y, Z = numpy.meshgrid(*ax, indexing='ij')
        set_vector(ok, X.ravel(), Y.ravel(), Z.ravel())
        vec = ok.rho_compute(qc, calc_ao=True)
        assert not eng.last_kernel().startswith('zrun/')
        assert_close(got.reshape(vec.shape), vec, '%s zrun vs vector %s' % (name, shape))
    # point ranges cut inside z runs reproduce the full evaluation bit for bit
    ax = [numpy.linspace(-3, 3, 4), numpy.linspace(-2, 2, 5), numpy.linspace(-3, 3, 41)]
    basis = eng.basis(qc.geo_spec, qc.ao_spec)
    g = eng.grid_regular(*ax)
    full = eng.eval_ao(basis, g, [0])
    assert eng.last_kernel().startswith('zrun/')
    for p0, p1 in ((0, 1), (7, 300), (41, 82), (163, 164), (500, 820), (819, 820)):
        part = eng.eval_ao(basis, g, [0], p0, p1)
        assert numpy.array_equal(part, full[:, :, p0:p1]), (p0, p1)
    # ao_creator (core.py:38-105) takes the same route
    got = ok.core.ao_creator(qc.geo_spec, qc.ao_spec, x=ax[0], y=ax[1], z=ax[2], is_vector=False)
    assert numpy.array_equal(got.reshape(full[0].shape), full[0])
    # first and pure second derivatives (codes 1..6): one launch per code, A(Z) R0 + B(Z) R1 [+ C(Z) R2]
    for shape in ((3, 4, 5), (2, 3, 37), (2, 2, 130)):
        ax = [numpy.sort(rng.uniform(-3.5, 3.5, n)) for n in shape]
        set_regular(ok, *ax)
        drv = ['x', 'y', 'z', 'xx', 'yy', 'zz']
        got = ok.rho_compute(qc, calc_ao=True, drv=drv)
        assert eng.last_kernel().startswith('zrun/ONE6'), eng.last_kernel()
        for i, d in enumerate(drv):
            ref = oracle_mod.ao_creator(qc.geo_spec, qc.ao_spec, drv=d, x=ax[0], y=ax[1], z=ax[2], is_vector=False)
            assert_close(got[i], ref.reshape(got[i].shape), '%s zrun d/d%s %s' % (name, d, shape), afloor=1e-13)
        one = ok.core.ao_creator(qc.geo_spec, qc.ao_spec, drv='y', x=ax[0], y=ax[1], z=ax[2], is_vector=False)
        assert eng.last_kernel().startswith('zrun/ONE2')
        assert numpy.array_equal(one.reshape(got[1].shape), got[1])
        mixed = ok.rho_compute(qc, calc_ao=True, drv=[None, 'z', 'xy'])          # a mixed code: not the z-run kernel
        assert not eng.last_kernel().startswith('zrun/')
        assert_close(mixed[1], got[2], '%s zrun vs tile kernel d/dz' % name, afloor=1e-13)


_REM_WORKER = r'''
import os, sys, numpy
sys.path.insert(0, %(repo)r)
sys.path.insert(0, os.path.join(%(repo)r, 'tests'))
import orbkit_b200 as ok
from orbkit_b200.engine import get_engine
from conftest import golden_qc, assert_close
ok.options.quiet = True
seen = set()
for name in %(names)r:
    qc, a = golden_qc(name)
    ok.grid.set_grid(a['rx'], a['ry'], a['rz'], is_vector=False)
    r, d = ok.rho_compute(qc, drv=['x', 'y', 'z'])
    seen.add(get_engine().last_kernel())
    assert_close(r, a['reg.rho'], name + ' reg.rho')
    assert_close(d, a['reg.drho'], name + ' reg.drho')
    r, d, l = ok.rho_compute(qc, laplacian=True)
    seen.add(get_engine().last_kernel())
    assert_close(d, a['reg.d2rho'], name + ' reg.d2rho')
    assert_close(l, a['reg.lap'], name + ' reg.lap', afloor=3e-14)
    ok.grid.set_grid(a['vx'], a['vy'], a['vz'], is_vector=True)
    r, d = ok.rho_compute(qc, drv=['x', 'y', 'z'])
    assert_close(d, a['vec.drho'], name + ' vec.drho')
    m = ok.rho_compute(qc, calc_mo=True, drv=['z'])[0]
    assert_close(m, a['vec.mo10'][3] if 'vec.mo10' in a.files else a['vec.mo_z'], name + ' mo_z')
    # mo_norm path (want_norm): regular grid, no derivative -> the displayed norms are accumulated
    # two rho_compute calls with the same inputs are bit-identical (the remainder partial sums are added in a fixed order)
    r2, d2 = ok.rho_compute(qc, drv=['x', 'y', 'z'])
    assert numpy.array_equal(d, d2) and numpy.array_equal(r, r2)
# results do not depend on how the points are cut into launches (slabs of a big request, shards of a multi-GPU job): the
# remainder orbitals' partial sums are grouped by producer warp, and which warp takes which shell is fixed per tile
qc, a = golden_qc('lih_psi4_sph_f')
eng = get_engine()
ax = numpy.linspace(-4.0, 4.0, 40)
basis = eng.basis(qc.geo_spec, qc.ao_spec)
mo = eng.mos(basis, qc.mo_spec.get_coeffs(), qc.mo_spec.get_occ())
g = eng.grid_regular(ax, ax, ax)
n = 40 ** 3
whole = eng.eval_rho(mo, g, [1, 2, 3])
for cut in (128, 4096, 33920):
    lo = eng.eval_rho(mo, g, [1, 2, 3], 0, cut)
    hi = eng.eval_rho(mo, g, [1, 2, 3], cut, n)
    assert numpy.array_equal(numpy.concatenate([lo[0], hi[0]]), whole[0]), cut
    assert numpy.array_equal(numpy.concatenate([lo[1], hi[1]], axis=1), whole[1]), cut
print('kernels', sorted(seen))
assert any('MB10R2' in k for k in seen), seen
'''


def test_remainder_orbital_tiles_forced_on_small_molecules(tmp_path):
    """The 8 MB + 2 tiles (two remainder orbitals contracted by the producer warps, okb_ws.cuh REM) are chosen by the
    cost model only for MO counts such as 82 or 246; here they are FORCED (OKB_VARIANT, read once per process) on the
    small fixture molecules: one chunk per pass (the producers run several tiles ahead of the consumers, so the hand-over
    of the partial sums is exercised across tiles), fewer MOs than one tile, ragged point counts, both laplacian passes,
    the SINK_MO epilogue, vector and regular grids."""
    import os
    import subprocess
    import sys
    from conftest import REPO
    script = tmp_path / 'rem_worker.py'
    script.write_text(_REM_WORKER % {'repo': REPO, 'names': ['h2o_gaussian_sph', 'lih_psi4_sph_f', 'synth_small_cart_g']})
    env = dict(os.environ, OKB_VARIANT='MB10R2')
    p = subprocess.run([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900)
    assert p.returncode == 0, p.stdout.decode()[-3000:]
