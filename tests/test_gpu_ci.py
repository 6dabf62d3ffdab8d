"""GPU parity of the detCI grid contractions (orbkit_b200.detci.ci_core, C ABI okb_ci_contract /
okb_eval_ci) against
   (1) the reference's own outputs for its test orbkit/test/detci/h3+.py (tests/golden/h3p_detci.npz)
       and its committed golden refdata_h3+.npz,
   (2) the pinned CPU oracle (oracle/oracle_ci.py) on seeded random term lists, bit for bit,
   (3) properties at Config-5 scale (500 MOs, 1000 pairs): linearity in the coefficients, symmetry,
       per-pair products summing to the contracted density.
The contraction is bit-exact for identical MO arrays; MOs evaluated on the device carry the FP64
tolerance of the grid path (|d| <= 1e-10*|ref| + 1e-14*max|ref|)."""
import numpy
import pytest

from conftest import assert_close, load_golden
from test_oracle_ci import lists_from_golden, random_lists

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ok():
    import orbkit_b200
    orbkit_b200.options.quiet = True
    orbkit_b200.options.ci_merge_terms = False
    return orbkit_b200


@pytest.fixture(scope='module')
def oci(oracle_mod):
    import oracle_ci
    return oracle_ci


def test_h3p_reference_test_reproduced(ok, oci):
    """the reference's detCI test, end to end on the device"""
    from orbkit_b200.detci import ci_core
    g = load_golden('h3p_detci')
    qc = ok.QCinfo.from_arrays(g)
    ok.grid.set_grid(g['x'], g['y'], g['z'], is_vector=False)
    molist = ok.rho_compute(qc, calc_mo=True, slice_length=1e2, drv=[None, 'x', 'y', 'z', 'xx', 'yy', 'zz'])
    mo, d1, d2 = molist[0], molist[1:4], molist[-3:]
    for pair in range(int(g['n_pairs'])):
        zero, sing = lists_from_golden(g, pair)
        rho = ci_core.rho(zero, sing, mo, slice_length=1e2)
        jab = ci_core.jab(zero, sing, mo, d1, slice_length=1e2)
        nj = -numpy.sum(ci_core.jab(zero, sing, mo, d2, slice_length=1e2), axis=0)
        anb = ci_core.a_nabla_b(zero, sing, mo, d1, slice_length=1e2)
        assert rho.shape == (51, 51, 1) and jab.shape == (3, 51, 51, 1)
        for got, key in ((rho, 'rho_01'), (jab, 'j_01'), (nj, 'nabla_j_01'), (anb, 'a_nabla_b_01')):
            assert_close(got, g[key][pair], '%s pair %d' % (key, pair))
            if 'published.' + key in g.files:            # the reference's own tolerance (test/tools.py:11-21)
                assert numpy.allclose(got, g['published.' + key][pair], rtol=1e-3, atol=1e-5)
        # the reference's slice driver never visits the point behind the last full slice
        assert rho.reshape(-1)[-1] == 0.0 and (jab.reshape(3, -1)[:, -1] == 0.0).all()
        # fused variants (MOs never leave the device) evaluate every point
        full = ci_core.rho(zero, sing, mo, slice_length=mo[0].size)
        assert_close(ci_core.rho_from_qc(qc, zero, sing), full, 'rho_from_qc')
        assert_close(ci_core.jab_from_qc(qc, zero, sing), ci_core.jab(zero, sing, mo, d1, slice_length=2601), 'jab_from_qc')
        assert_close(ci_core.jab_from_qc(qc, zero, sing, drv=['xx', 'yy', 'zz']),
                     ci_core.jab(zero, sing, mo, d2, slice_length=2601), 'jab_from_qc d2')
        assert_close(ci_core.a_nabla_b_from_qc(qc, zero, sing),
                     ci_core.a_nabla_b(zero, sing, mo, d1, slice_length=2601), 'a_nabla_b_from_qc')


def test_random_terms_bitwise_vs_oracle(ok, oci):
    """same MO arrays in -> bit-identical results (reference expression order, no FMA contraction)"""
    from orbkit_b200.detci import ci_core
    rng = numpy.random.default_rng(21)
    for n_mo, shape, n_det, n_sing in ((3, (1,), 0, 1), (7, (33,), 4, 19), (20, (5, 13, 4), 30, 400),
                                       (64, (4099,), 100, 2500), (5, (130,), 3, 0)):
        zero, sing = random_lists(rng, n_mo, n_det, n_sing)
        mo = rng.normal(size=(n_mo,) + shape)
        dmo = rng.normal(size=(3, n_mo) + shape)
        n = mo[0].size
        for sl in (1e4, 100, n):
            if sl == 100 and n < 100:
                continue
            assert numpy.array_equal(ci_core.rho(zero, sing, mo, slice_length=sl), oci.rho(zero, sing, mo, sl))
            assert numpy.array_equal(ci_core.jab(zero, sing, mo, dmo, slice_length=sl), oci.jab(zero, sing, mo, dmo, sl))
            assert numpy.array_equal(ci_core.a_nabla_b(zero, sing, mo, dmo, slice_length=sl),
                                     oci.a_nabla_b(zero, sing, mo, dmo, sl))
        # the caller's arrays keep their shape (the reference reshapes them in place and restores, ci_core.py:112-135)
        assert mo.shape == (n_mo,) + shape and dmo.shape == (3, n_mo) + shape
    # few active orbitals among many MOs: only the rows the lists refer to cross PCIe, same bits out
    act = numpy.array([3, 17, 18, 40, 41, 59])
    zero, sing = random_lists(rng, len(act), 5, 300)
    zero = [zero[0], [[int(act[i]) for i in idx] for idx in zero[1]]]
    sing = [sing[0], [[int(act[a]), int(act[b])] for a, b in sing[1]]]
    mo = rng.normal(size=(60, 1501)); dmo = rng.normal(size=(3, 60, 1501))
    from orbkit_b200.engine import get_engine
    h0 = get_engine().traffic()[0]
    got = ci_core.rho(zero, sing, mo, slice_length=1501)
    assert get_engine().traffic()[0] - h0 < 8 * 1501 * (len(act) + 1) + 16 * 400          # 6 of the 60 rows + the terms
    assert numpy.array_equal(got, oci.rho(zero, sing, mo, 1501))
    assert numpy.array_equal(ci_core.jab(zero, sing, mo, dmo, slice_length=1501), oci.jab(zero, sing, mo, dmo, 1501))
    assert numpy.array_equal(ci_core.a_nabla_b(zero, sing, mo, dmo, slice_length=1501), oci.a_nabla_b(zero, sing, mo, dmo, 1501))
    prod = ci_core.pair_products(numpy.array([[3, 41], [59, 59]]), mo)
    assert numpy.array_equal(prod, numpy.stack([mo[3] * mo[41], mo[59] * mo[59]]))
    # empty grid, error conventions
    assert ci_core.rho([[], []], [[], []], numpy.zeros((4, 0))).shape == (0,)
    assert (ci_core.rho([[], []], [[], []], numpy.ones((4, 9))) == 0.0).all()
    with pytest.raises(ValueError):
        ci_core.rho([[], []], [[1.0], [[0, 4]]], numpy.ones((4, 9)))
    with pytest.raises(ValueError):
        ci_core.jab([[], []], [[1.0], [[0, 1]]], numpy.ones((4, 9)), numpy.ones((2, 4, 9)))


def test_merged_terms_and_pair_products(ok, oci):
    from orbkit_b200.detci import ci_core
    rng = numpy.random.default_rng(22)
    zero, sing = random_lists(rng, 12, 40, 900)
    mo = rng.normal(size=(12, 777))
    dmo = rng.normal(size=(3, 12, 777))
    exact = ci_core.rho(zero, sing, mo)
    exact_anb = ci_core.a_nabla_b(zero, sing, mo, dmo)
    ok.options.ci_merge_terms = True
    try:
        assert_close(ci_core.rho(zero, sing, mo), exact, 'merged rho', rtol=1e-12, afloor=1e-13)
        assert_close(ci_core.a_nabla_b(zero, sing, mo, dmo), exact_anb, 'merged a_nabla_b', rtol=1e-12, afloor=1e-13)
        assert numpy.array_equal(ci_core.jab(zero, sing, mo, dmo), oci.jab(zero, sing, mo, dmo))   # never merged
    finally:
        ok.options.ci_merge_terms = False
    pairs = numpy.array(sing[1][:50])
    prod = ci_core.pair_products(pairs, mo)
    assert prod.shape == (50, 777)
    assert numpy.array_equal(prod, mo[pairs[:, 0]] * mo[pairs[:, 1]])


def test_fast_sums_dense_and_split(ok, oci):
    """options.ci_fast = True (OKB_FLAG_CI_FAST): the re-ordered device sums -- dense orbital-pair matrix on the FP64
    tensor path when there are many terms per orbital pair, terms split over the warps of a CTA otherwise -- against the
    bit-exact path.  Tolerance: 1e-12 of the largest result (sums of up to 2500 terms of O(1) products, re-associated);
    repeated calls are bit-identical (fixed grouping)."""
    import os
    from orbkit_b200.detci import ci_core, cy_ci
    from orbkit_b200.engine import get_engine
    rng = numpy.random.default_rng(23)
    cases = ((12, (777,), 40, 900, 'ci-dense'),          # many terms per pair, ragged point count
             (64, (4099,), 100, 2500, 'ci-dense'),
             (7, (2, 65), 4, 19, 'ci-'),                  # dense for rho and jab, split for a_nabla_b
             (150, (1500,), 0, 300, 'ci-fast'),           # few terms over many orbitals: split over warps + ragged tail
             (600, (257,), 10, 500, 'ci/'),               # row sets too wide for the split kernel: gather kernel
             (3, (1,), 0, 1, None))                        # a single point
    for n_mo, shape, n_det, n_sing, kernel in cases:
        zero, sing = random_lists(rng, n_mo, n_det, n_sing)
        mo = rng.normal(size=(n_mo,) + shape)
        dmo = rng.normal(size=(3, n_mo) + shape)
        n = mo[0].size
        exact = (ci_core.rho(zero, sing, mo, slice_length=n), ci_core.jab(zero, sing, mo, dmo, slice_length=n),
                 ci_core.a_nabla_b(zero, sing, mo, dmo, slice_length=n))
        ok.options.ci_fast = True
        try:
            fast = (ci_core.rho(zero, sing, mo, slice_length=n), ci_core.jab(zero, sing, mo, dmo, slice_length=n),
                    ci_core.a_nabla_b(zero, sing, mo, dmo, slice_length=n))
            if kernel:
                assert get_engine().last_kernel().startswith(kernel), (n_mo, get_engine().last_kernel())
            again = ci_core.jab(zero, sing, mo, dmo, slice_length=n)
        finally:
            ok.options.ci_fast = None
        for f, e, name in zip(fast, exact, ('rho', 'jab', 'a_nabla_b')):
            assert f.shape == e.shape
            assert numpy.abs(f - e).max() <= 1e-12 * max(numpy.abs(e).max(), 1e-300), (n_mo, name)
        assert numpy.array_equal(again, fast[1])
    # device-resident rows with an odd row stride (unaligned: scalar loads in the dense kernel, gather kernel instead of
    # the split kernel) and 16-byte aligned ones, results left on the device
    import torch
    from orbkit_b200 import _lib
    eng = get_engine()
    dev = torch.device('cuda', eng.device)
    fl = _lib.OKB_FLAG_IN_DEVICE | _lib.OKB_FLAG_OUT_DEVICE
    for n_mo, n_terms, ld, npts in ((9, 400, 1001, 999), (9, 400, 1002, 1000), (90, 60, 1001, 1001), (90, 60, 1000, 1000)):
        pairs = rng.integers(0, n_mo, size=(n_terms, 2))
        terms = (rng.normal(size=n_terms), pairs[:, 0].astype(numpy.intc), pairs[:, 1].astype(numpy.intc))
        buf = torch.from_numpy(rng.normal(size=(4, n_mo, ld))).to(dev)
        for mode, nc in ((_lib.OKB_CI_RHO, 1), (_lib.OKB_CI_JAB, 3), (_lib.OKB_CI_A_NABLA_B, 3)):
            outs = []
            for fast in (0, _lib.OKB_FLAG_CI_FAST):
                out = torch.full((3 * npts + 8,), 7.0, dtype=torch.float64, device=dev)     # rows of stride npts + guard
                eng.ci_contract(mode, terms, buf[0].data_ptr(), buf[1:].data_ptr(), n_mo=n_mo, npts=npts, ld=ld,
                                out=out.data_ptr(), flags=fl | fast)
                eng.sync()
                outs.append(out.cpu().numpy())
            e, f = outs
            assert (f[nc * npts:] == 7.0).all() and (e[nc * npts:] == 7.0).all()    # nothing written outside the result
            e, f = e[:nc * npts], f[:nc * npts]
            assert numpy.abs(f - e).max() <= 1e-12 * numpy.abs(e).max(), (n_mo, ld, mode)
    # cy_ci.get_jab_full: the state-pair sum is a dense antisymmetric matrix by construction
    nb, npts = 9, 1111
    ImS = rng.normal(size=(nb, nb)); chi = rng.normal(size=(nb, npts)); dchi = rng.normal(size=(2, nb, npts))
    exact = cy_ci.get_jab_full(ImS, chi, dchi, 1.7)
    assert numpy.array_equal(exact, oci.get_jab_full(ImS, chi, dchi, 1.7))
    ok.options.ci_fast = True
    try:
        fast = cy_ci.get_jab_full(ImS, chi, dchi, 1.7)
        assert get_engine().last_kernel() == 'ci-dense/jab_full'
    finally:
        ok.options.ci_fast = None
    assert numpy.abs(fast - exact).max() <= 1e-12 * numpy.abs(exact).max()
    # ci_fast = False: the fused call keeps the reference's order as well
    g = load_golden('h3p_detci')
    qc = ok.QCinfo.from_arrays(g)
    ok.grid.set_grid(g['x'], g['y'], g['z'], is_vector=False)
    zero, sing = lists_from_golden(g, 0)
    a = ci_core.rho_from_qc(qc, zero, sing)          # natural orbitals of the pair matrix: one fused rho launch
    assert not get_engine().last_kernel().startswith('ci')
    j = ci_core.jab_from_qc(qc, zero, sing)
    assert get_engine().last_kernel().startswith('ci-')
    ok.options.ci_fast = False
    try:
        b = ci_core.rho_from_qc(qc, zero, sing)
        assert get_engine().last_kernel() == 'ci/rho'
        jb = ci_core.jab_from_qc(qc, zero, sing)
        assert get_engine().last_kernel() in ('ci/jab', 'ci-seq/jab')          # reference order, either kernel
    finally:
        ok.options.ci_fast = None
    assert numpy.abs(a - b).max() <= 1e-12 * max(numpy.abs(b).max(), 1e-300)
    assert numpy.abs(j - jb).max() <= 1e-12 * max(numpy.abs(jb).max(), 1e-300)


def test_config5_scale_properties(ok):
    """500 AOs / 500 MOs, 1000 random pairs on 32^3 points of the Config-5 generator: the fused device
    path against (a) the two-step path through host MOs, (b) linearity, (c) sum of per-pair products"""
    from orbkit_b200 import synth
    from orbkit_b200.detci import ci_core
    spec = synth.make_molecule(n_heavy=12, n_light=10, n_mo=500, seed=5, spherical=True)
    qc = synth.to_qcinfo(spec)
    n_mo = len(qc.mo_spec)
    rng = numpy.random.default_rng(5)
    pairs = rng.integers(0, n_mo, size=(1000, 2))
    coef = rng.normal(size=1000)
    zero = [[], []]
    sing = [list(coef), [list(p) for p in pairs]]
    ax = numpy.linspace(-9, 9, 32)
    ok.grid.set_grid(ax, ax, ax, is_vector=False)
    rho_dev = ci_core.rho_from_qc(qc, zero, sing)
    mos = ok.rho_compute(qc, calc_mo=True, drv=[None, 'x', 'y', 'z'])
    rho_host = ci_core.rho(zero, sing, mos[0], slice_length=mos[0][0].size)
    assert_close(rho_dev, rho_host, 'fused vs two-step')
    ref = numpy.einsum('k,kxyz,kxyz->xyz', coef, mos[0][pairs[:, 0]], mos[0][pairs[:, 1]])
    assert_close(rho_dev, ref, 'fused vs einsum', rtol=1e-10, afloor=1e-13)
    j_dev = ci_core.jab_from_qc(qc, zero, sing)
    jref = -0.5 * (numpy.einsum('k,kxyz,dkxyz->dxyz', coef, mos[0][pairs[:, 0]], mos[1:4][:, pairs[:, 1]]) -
                   numpy.einsum('k,kxyz,dkxyz->dxyz', coef, mos[0][pairs[:, 1]], mos[1:4][:, pairs[:, 0]]))
    assert_close(j_dev, jref, 'jab fused vs einsum', rtol=1e-10, afloor=1e-13)
    # linearity in the CI coefficients and antisymmetry of the flux density under a <-> b
    sing2 = [list(2.5 * coef), sing[1]]
    assert_close(ci_core.rho_from_qc(qc, zero, sing2), 2.5 * rho_dev, 'linearity', rtol=1e-13, afloor=1e-15)
    swapped = [sing[0], [[b, a] for a, b in sing[1]]]
    assert_close(ci_core.jab_from_qc(qc, zero, swapped), -j_dev, 'antisymmetry', rtol=1e-13, afloor=1e-15)
    sub = mos[0][:, ::6, ::6, ::6]
    prods = ci_core.pair_products(pairs, sub)
    assert_close(numpy.tensordot(coef, prods, axes=1), ci_core.rho(zero, sing, sub, slice_length=sub[0].size),
                 'sum of pair products', rtol=1e-12, afloor=1e-13)


def test_gross_atomic_density(ok, oracle_mod):
    """extras.gross_atomic_density (extras.py:306-385): fused device path, the two-step path with
    caller-supplied components, derivatives and the gross atomic MOs against the oracle"""
    from conftest import golden_qc
    for name in ('h2o_molpro_cart', 'h2o_gaussian_sph'):
        qc, a = golden_qc(name)
        ok.grid.set_grid(a['vx'], a['vy'], a['vz'], is_vector=True)
        ref_rho, ref_mo = oracle_mod.gross_atomic_density([0, 1, 2], qc, a['vx'], a['vy'], a['vz'], is_vector=True)
        got = ok.extras.gross_atomic_density('all', qc)
        assert len(got) == 3
        for g, r in zip(got, ref_rho):
            assert_close(g, r, name + ' gross atomic density')
        # the atoms' densities add up to the density
        assert_close(sum(got), ok.rho_compute(qc), name + ' sum over atoms', rtol=1e-10, afloor=1e-13)
        # single atom by its number (counting from one), with the gross atomic MOs
        rho2, mo2 = ok.extras.gross_atomic_density(2, qc, bReturnmo=True)
        assert_close(rho2[0], ref_rho[1], name + ' atom 2')
        assert len(mo2) == 1 and len(mo2[0]) == len(qc.mo_spec)
        assert_close(numpy.array(mo2[0]), numpy.array(ref_mo[1]), name + ' gross atomic MOs')
        # caller-supplied components, and a derivative
        ao = ok.ao_creator(qc.geo_spec, qc.ao_spec)
        mo = ok.mo_creator(ao, qc.mo_spec)
        assert_close(ok.extras.gross_atomic_density([1, 3], qc, ao_list=ao, mo_list=mo)[1], ref_rho[2], name + ' given')
        dref, _ = oracle_mod.gross_atomic_density([0], qc, a['vx'], a['vy'], a['vz'], is_vector=True, drv='x')
        assert_close(ok.extras.gross_atomic_density(1, qc, drv='x')[0], dref[0], name + ' drv=x')
    # regular grid shape
    ax = numpy.linspace(-3, 3, 6)
    ok.grid.set_grid(ax, ax, ax, is_vector=False)
    assert ok.extras.gross_atomic_density(1, qc)[0].shape == (6, 6, 6)
    with pytest.raises(ValueError):
        ok.extras.gross_atomic_density(7, qc)


def test_calc_mo_matrix_and_calc_jmo(ok, oci):
    """core.calc_mo_matrix (core.py:841-941) and extras.calc_jmo (extras.py:441-493) on the device against the
    reference's own outputs (tests/golden/h2o_mo_matrix.npz) and the pinned oracle"""
    from conftest import golden_qc
    qc, g = golden_qc('h2o_mo_matrix')
    x, y, z = g['grid.x'], g['grid.y'], g['grid.z']
    ok.grid.set_grid(x, y, z, is_vector=False)
    eng = ok.engine.get_engine()
    assert_close(ok.core.calc_mo_matrix(qc, drv=['x', 'y', 'z']), g['mm_xyz'], 'mo_matrix xyz')
    assert eng.last_kernel() == 'ci/pairs'
    assert_close(ok.core.calc_mo_matrix(qc), g['mm_none'], 'mo_matrix')
    assert_close(ok.core.calc_mo_matrix(qc, drv='xx'), g['mm_xx'], 'mo_matrix xx')
    assert_close(ok.extras.calc_jmo(qc, g['ij']), g['jmo'], 'jmo')
    assert eng.last_kernel() == 'ci/jpairs'
    assert_close(ok.extras.calc_jmo(qc, g['ij'], drv=['z', 'x']), g['jmo_zx'], 'jmo zx')
    assert_close(ok.extras.calc_jmo(qc, [4, 1]), g['jmo_one'], 'jmo one pair')
    # two QCinfos (the reference's branch raises TypeError, core.py:906): the restated intent, against the oracle
    qa, qb = qc.copy(), qc.copy()
    qa.mo_spec = qc.mo_spec[numpy.array([0, 1, 2])]
    qb.mo_spec = qc.mo_spec[numpy.array([1, 2, 3, 4])]
    got = ok.core.calc_mo_matrix(qa, qb, drv=['z', 'y'])
    assert got.shape == (2, 3, 4, 5, 6, 7)
    assert_close(got, oci.calc_mo_matrix(qa, x, y, z, qc_b=qb, drv=['z', 'y']), 'mo_matrix a,b')
    # vector grid, ragged point count, more pairs than a term batch, a larger molecule
    qc, a = golden_qc('synth_small_sph')
    ok.grid.set_grid(a['vx'], a['vy'], a['vz'], is_vector=True)
    rng = numpy.random.default_rng(5)
    ij = rng.integers(0, len(qc.mo_spec), size=(300, 2))
    got = ok.extras.calc_jmo(qc, ij)
    ref = oci.calc_jmo(qc, ij, a['vx'], a['vy'], a['vz'], is_vector=True)
    assert got.shape == (3, 300, len(a['vx']))
    assert_close(got, ref, 'jmo synth')
    # antisymmetry is exact: j(i, j) == -j(j, i)
    assert numpy.array_equal(ok.extras.calc_jmo(qc, ij[:, ::-1]), -got)
    with pytest.raises(ValueError):
        ok.extras.calc_jmo(qc, ij, drv=[None, 'x', 'y'])
    with pytest.raises(NotImplementedError):
        ok.extras.calc_jmo(qc, ij, otype='am')


def test_time_dependent_contractions(ok, oci):
    """cy_ci.get_rho_full / get_j_full on the FP64 tensor cores against the oracle (the reference's own module when
    oracle/_ref is built).  The DMMA k-steps run in the reference's order but round once per fused step, so the stated
    tolerance is |d| <= 1e-10 |ref| + 1e-13 max|ref| (measured: a few 1e-16 max|ref|).  get_jab_full walks the pairs with
    the sequential pair kernel: bit for bit."""
    from orbkit_b200.detci import cy_ci
    kind = 'ref' if oci.have_ref() else 'port'
    rng = numpy.random.default_rng(17)
    # (nt, nstate, npts): one pair; ragged everything; more pairs than one k chunk (nstate 12 -> 78 pairs); many time steps
    # (70, 12), (130, 17), (40, 13): 78 / 153 / 91 pairs, several k chunks
    for nt, ns, npts in ((1, 1, 1), (5, 2, 127), (33, 4, 1000), (70, 12, 257), (300, 3, 4099), (130, 17, 515), (40, 13, 300)):
        npair = ns * (ns + 1) // 2
        ReS, ImS = rng.normal(size=(nt, ns, ns)), rng.normal(size=(nt, ns, ns))
        rho, j = rng.normal(size=(npair, npts)), rng.normal(size=(npair, 3, npts))
        for got, ref in ((cy_ci.get_rho_full(ReS, rho), oci.get_rho_full(ReS, rho, kind=kind)),
                         (cy_ci.get_j_full(ImS, j), oci.get_j_full(ImS, j, kind=kind))):
            assert got.shape == ref.shape and got.dtype == numpy.float64
            tol = 1e-10 * numpy.abs(ref) + 1e-13 * numpy.abs(ref).max()
            assert (numpy.abs(got - ref) <= tol).all(), (nt, ns, npts, numpy.abs(got - ref).max())
    for nb, nc, npts in ((1, 3, 9), (2, 1, 33), (7, 3, 1025), (25, 2, 300)):
        S, chi, dchi = rng.normal(size=(nb, nb)), rng.normal(size=(nb, npts)), rng.normal(size=(nc, nb, npts))
        assert numpy.array_equal(cy_ci.get_jab_full(S, chi, dchi, 1836.15), oci.get_jab_full(S, chi, dchi, 1836.15, kind=kind))
    # linearity in the weights at a larger size (size-independent property): tdrho(a S1 + S2) = a tdrho(S1) + tdrho(S2)
    nt, ns, npts = 64, 6, 200000
    npair = ns * (ns + 1) // 2
    S1, S2, rho = rng.normal(size=(nt, ns, ns)), rng.normal(size=(nt, ns, ns)), rng.normal(size=(npair, npts))
    lhs = cy_ci.get_rho_full(2.0 * S1 + S2, rho)
    rhs = 2.0 * cy_ci.get_rho_full(S1, rho) + cy_ci.get_rho_full(S2, rho)
    assert numpy.abs(lhs - rhs).max() <= 1e-12 * numpy.abs(rhs).max()
    sub = rng.integers(0, npts, size=64)
    ref = oci.get_rho_full(S1, numpy.ascontiguousarray(rho[:, sub]), kind=kind)
    assert numpy.abs(cy_ci.get_rho_full(S1, rho)[:, sub] - ref).max() <= 1e-13 * numpy.abs(ref).max()
