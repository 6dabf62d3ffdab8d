"""BASELINE.json configs[1..4] at their STATED shapes through the public API on the device, each checked against the
reference's own objects (oracle/_ref; the bit-identical C restatement where _ref is absent) on a random sub-sample
of the full grid, plus the size-independent properties the configuration offers.

    C2  benzene (geometry of the reference's example input), def2-TZVP-shaped spherical basis: 222 AOs (252 Cartesian),
        all 222 MOs on 150^3 points; density + gradient of the 21 doubly occupied MOs; N_el consistency
    C3  1000 AOs / 82 MOs, rho + Laplacian (laplacian=True) on 200^3 points
    C4  3000 AOs / 246 MOs, rho + gradient on one rank's x-slab (1/8) of the 256^3 grid
    C5  500 AOs / 500 MOs, 1000 MO pairs: transition density and flux density on 128^3 points, against the oracle's
        detCI loops fed with the REFERENCE's MOs

Tolerance (FP64 path): |d| <= 1e-10*|ref| + 1e-14*max|ref|; the detCI sums of 1000 signed products use the
absolute floor 1e-13*max|ref| (cancellation between the terms); electron counts agree to 1e-8.
"""
import numpy
import pytest

from conftest import assert_close, load_golden

pytestmark = pytest.mark.gpu

NSAMPLE = 64


@pytest.fixture(scope='module')
def ok():
    import orbkit_b200
    orbkit_b200.options.quiet = True
    return orbkit_b200


def sample_points(shape, axes, seed):
    """NSAMPLE random points of a regular grid: (flat indices, x, y, z)"""
    rng = numpy.random.default_rng(seed)
    n = int(numpy.prod(shape))
    idx = numpy.sort(rng.choice(n, size=NSAMPLE, replace=False))
    i, rem = numpy.divmod(idx, shape[1] * shape[2])
    j, k = numpy.divmod(rem, shape[2])
    return idx, axes[0][i], axes[1][j], axes[2][k]


def ref_kind(oracle_mod):
    return 'ref' if oracle_mod.have_ref() else 'port'


def test_config2_benzene_tzvp_shaped_150cube(ok, oracle_mod):
    """BASELINE configs[1]: "Benzene def2-TZVP spherical basis: all MOs + density + gradient on a 150^3 grid".
    def2-TZVP-shaped: the contraction pattern C [5s3p2d1f] / H [3s1p] (222 spherical AOs); exponents from memory,
    MO coefficients seeded random (orbkit_b200.synth.make_benzene_tzvp)."""
    from orbkit_b200 import synth
    g = load_golden('benzene_geometry')
    spec = synth.make_benzene_tzvp(g['geo_spec'], g['geo_info'], n_occ=21, seed=2, all_mo=True)
    assert synth.counts(spec) == {'n_cont': 90, 'n_prim': 156, 'n_cart': 252, 'n_ao': 222, 'n_mo': 222}
    qc = synth.to_qcinfo(spec)
    # adjust_to_geo-style box: the molecule's extent + 5 bohr (grid.adjust_to_geo, grid.py:135-160)
    lo, hi = qc.geo_spec.min(axis=0) - 5.0, qc.geo_spec.max(axis=0) + 5.0
    axes = [numpy.linspace(lo[a], hi[a], 150) for a in range(3)]
    ok.grid.set_grid(axes[0], axes[1], axes[2], is_vector=False)
    d3r = numpy.prod([a[1] - a[0] for a in axes])
    idx, sx, sy, sz = sample_points((150, 150, 150), axes, seed=31)
    kind = ref_kind(oracle_mod)

    # ---- all 222 MOs on 150^3 points (6 GB of results) -------------------------------------------------------
    mo = ok.rho_compute(qc, calc_mo=True)
    assert mo.shape == (222, 150, 150, 150)
    mo_ref = oracle_mod.rho_compute(qc, sx, sy, sz, is_vector=True, calc_mo=True, kind=kind)
    assert_close(mo.reshape(222, -1)[:, idx], mo_ref, 'C2 all MOs')
    norm_occ = numpy.array([numpy.square(mo[i]).sum() for i in range(21)])
    del mo

    # ---- density + gradient of the 21 occupied MOs ------------------------------------------------------------
    qo = qc.copy()
    qo.mo_spec = qc.mo_spec[:21]
    rho, drho = ok.rho_compute(qo, drv=['x', 'y', 'z'])
    assert rho.shape == (150, 150, 150) and drho.shape == (3, 150, 150, 150)
    r_ref, d_ref = oracle_mod.rho_compute(qo, sx, sy, sz, is_vector=True, drv=['x', 'y', 'z'], kind=kind)
    assert_close(rho.ravel()[idx], r_ref, 'C2 rho')
    assert_close(drho.reshape(3, -1)[:, idx], d_ref, 'C2 grad rho')
    # the virtual MOs carry occupation 0: the density of all 222 MOs is the density of the 21 occupied ones
    rho_all = ok.rho_compute(qc)
    assert_close(rho_all, rho, 'C2 rho with the virtual MOs', rtol=1e-12, afloor=1e-15)
    # N_el two ways: sum_p rho d3r  ==  sum_i occ_i sum_p phi_i^2 d3r   (to 1e-8, north_star)
    n_el = rho.sum() * d3r
    assert abs(n_el - 2.0 * norm_occ.sum() * d3r) < 1e-8 * max(1.0, abs(n_el))


def test_config3_laplacian_200cube(ok, oracle_mod):
    """BASELINE configs[2]: "density + Laplacian on a 200^3 grid" of the ~1000-basis-function molecule: the two-pass
    request (SET_GRAD + SET_D2P) at full size against the reference objects on a sub-sample"""
    from conftest import golden_qc
    qc, _ = golden_qc('synth_c3')
    ax = numpy.linspace(-12.0, 12.0, 200)
    ok.grid.set_grid(ax, ax, ax, is_vector=False)
    rho, d2rho, lap = ok.rho_compute(qc, laplacian=True)
    assert rho.shape == (200, 200, 200) and d2rho.shape == (3, 200, 200, 200) and lap.shape == (200, 200, 200)
    assert ok.engine.get_engine().last_kernel().startswith('ws-dmma/SET_D2P/')
    idx, sx, sy, sz = sample_points((200, 200, 200), [ax, ax, ax], seed=32)
    r_ref, d_ref, l_ref = oracle_mod.rho_compute(qc, sx, sy, sz, is_vector=True, laplacian=True, kind=ref_kind(oracle_mod))
    assert_close(rho.ravel()[idx], r_ref, 'C3 rho')
    assert_close(d2rho.reshape(3, -1)[:, idx], d_ref, 'C3 d2 rho')
    assert_close(lap.ravel()[idx], l_ref, 'C3 laplacian', afloor=3e-14)
    assert numpy.array_equal(lap, d2rho.sum(axis=0))
    # the density of the two-pass request is the density of the plain request
    assert_close(rho, ok.rho_compute(qc), 'C3 rho of the plain request', rtol=1e-12, afloor=1e-15)


def test_config4_3000ao_slab_of_256cube(ok, oracle_mod):
    """BASELINE configs[3]: "~3000-basis-function system on a 256^3 grid sharded across 8xB200": the molecule of the
    scaling run (config-3 generator x 3 atoms: 3000 AOs, 246 occupied MOs) on ONE rank's share of the grid -- the
    x-slab dist.shard_range hands rank 3 of 8 -- evaluated exactly as that rank does (point range of the full grid)"""
    from orbkit_b200 import synth, dist as okdist
    from orbkit_b200.engine import get_engine
    spec = synth.make_molecule(n_heavy=72, n_light=60, n_mo=246, seed=0, spherical=True)
    assert synth.counts(spec)['n_ao'] == 3000 and synth.counts(spec)['n_cart'] == 3420
    qc = synth.to_qcinfo(spec)
    ax = numpy.linspace(-12.0, 12.0, 256)
    npts = 256 ** 3
    p0, p1 = okdist.shard_range(npts, 3, 8)
    assert p1 - p0 == npts // 8
    eng = get_engine()
    basis = eng.basis(qc.geo_spec, qc.ao_spec)
    mo = eng.mos_of(basis, qc.mo_spec)
    g = eng.grid_regular(ax, ax, ax)
    rho, drho, norm = eng.eval_rho(mo, g, [1, 2, 3], p0, p1, want_norm=True)
    assert rho.shape == (p1 - p0,) and drho.shape == (3, p1 - p0) and numpy.isfinite(drho).all() and (rho >= 0).all()
    rng = numpy.random.default_rng(33)
    idx = numpy.sort(rng.choice(p1 - p0, size=NSAMPLE, replace=False))
    i, rem = numpy.divmod(idx + p0, 256 * 256)
    j, k = numpy.divmod(rem, 256)
    r_ref, d_ref = oracle_mod.rho_compute(qc, ax[i], ax[j], ax[k], is_vector=True, drv=['x', 'y', 'z'],
                                          kind=ref_kind(oracle_mod))
    assert_close(rho[idx], r_ref, 'C4 rho')
    assert_close(drho[:, idx], d_ref, 'C4 grad rho')
    # two accumulation paths over the slab: sum_i occ_i sum_p phi_i^2 == sum_p rho
    assert abs((qc.mo_spec.get_occ() * norm).sum() / rho.sum() - 1.0) < 1e-11
    # the same points through the public API on the slab as a grid of its own (x planes 96..127)
    assert p0 == 96 * 256 * 256
    ok.grid.set_grid(ax[96:128], ax, ax, is_vector=False)
    r2, d2 = ok.rho_compute(qc, drv=['x', 'y', 'z'])
    assert_close(r2.ravel(), rho, 'C4 public API', rtol=1e-12, afloor=1e-15)
    assert_close(d2.reshape(3, -1), drho, 'C4 public API grad', rtol=1e-12, afloor=1e-15)


def test_config5_500mo_1000pairs_128cube(ok, oracle_mod):
    """BASELINE configs[4]: "detCI-style batch: 1000 MO pairs / transition densities for a 500-basis-function molecule
    on a 128^3 grid": rho_from_qc / jab_from_qc (MOs evaluated and contracted on the device) on the full grid against
    the oracle's detCI loops (cy_ci.get_rho / get_jab) fed with the REFERENCE's MOs at the sample points"""
    import oracle_ci
    from orbkit_b200 import synth
    from orbkit_b200.detci import ci_core
    ok.options.ci_merge_terms = False
    spec = synth.make_molecule(n_heavy=12, n_light=10, n_mo=500, seed=5, spherical=True)
    assert synth.counts(spec)['n_ao'] == 500
    qc = synth.to_qcinfo(spec)
    rng = numpy.random.default_rng(5)
    pairs = rng.integers(0, 500, size=(1000, 2))
    coef = rng.normal(size=1000)
    # a few determinants' worth of diagonal terms, as detci.occ_check.compare returns them, and the 1000 pairs
    zero = [[list(rng.normal(size=4)), list(rng.normal(size=3))],
            [[int(v) for v in rng.integers(0, 500, size=4)], [int(v) for v in rng.integers(0, 500, size=3)]]]
    sing = [list(coef), [[int(a), int(b)] for a, b in pairs]]
    ax = numpy.linspace(-10.0, 10.0, 128)
    ok.grid.set_grid(ax, ax, ax, is_vector=False)
    rho = ci_core.rho_from_qc(qc, zero, sing)
    jab = ci_core.jab_from_qc(qc, zero, sing)
    assert rho.shape == (128, 128, 128) and jab.shape == (3, 128, 128, 128)
    idx, sx, sy, sz = sample_points((128, 128, 128), [ax, ax, ax], seed=35)
    kind = ref_kind(oracle_mod)
    mos = oracle_mod.rho_compute(qc, sx, sy, sz, is_vector=True, calc_mo=True, drv=[None, 'x', 'y', 'z'], kind=kind)
    ck = 'ref' if oracle_ci.have_ref() else 'port'
    rho_ref = oracle_ci.rho(zero, sing, mos[0], slice_length=NSAMPLE, kind=ck)
    jab_ref = oracle_ci.jab(zero, sing, mos[0], mos[1:4], slice_length=NSAMPLE, kind=ck)
    assert_close(rho.ravel()[idx], rho_ref, 'C5 transition density', afloor=1e-13)
    assert_close(jab.reshape(3, -1)[:, idx], jab_ref, 'C5 flux density', afloor=1e-13)
    # linearity in the CI coefficients at full size (exact scaling by a power of two)
    sing2 = [list(4.0 * coef), sing[1]]
    zero2 = [[[4.0 * c for c in cs] for cs in zero[0]], zero[1]]
    assert numpy.array_equal(ci_core.rho_from_qc(qc, zero2, sing2), 4.0 * rho)
