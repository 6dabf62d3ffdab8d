"""CPU: the detCI oracle (oracle/oracle_ci.py) is pinned to the reference.

  * port (okor_ci_* in libokoracle.so) == the reference's own cy_ci module (oracle/_ref) bit for bit on
    random term lists and on the h3+ FCI fixture;
  * both reproduce tests/golden/h3p_detci.npz, written by running the reference's test
    orbkit/test/detci/h3+.py (tests/golden/make_golden_detci.py), and agree with the reference's
    committed golden refdata_h3+.npz (`published.*`) to its own tolerance;
  * host logic of orbkit_b200.detci.ci_core: flattening, merging.
"""
import numpy
import pytest

from conftest import load_golden


def lists_from_golden(g, pair):
    """rebuild the (zero, sing) Python lists of state pair `pair`"""
    zc, zi, zn = g['p%d.zc' % pair], g['p%d.zi' % pair], g['p%d.zn' % pair]
    zero = [[], []]
    o = 0
    for n in zn:
        zero[0].append([float(c) for c in zc[o:o + n]])
        zero[1].append([int(i) for i in zi[o:o + n]])
        o += n
    sing = [[float(c) for c in g['p%d.sc' % pair]],
            [[int(a), int(b)] for a, b in zip(g['p%d.sa' % pair], g['p%d.sb' % pair])]]
    return zero, sing


def random_lists(rng, n_mo, n_det, n_sing):
    zero = [[], []]
    for _ in range(n_det):
        k = int(rng.integers(1, 5))
        zero[0].append([float(v) for v in rng.normal(size=k)])
        zero[1].append([int(v) for v in rng.integers(0, n_mo, size=k)])
    sing = [[float(v) for v in rng.normal(size=n_sing)],
            [[int(a), int(b)] for a, b in rng.integers(0, n_mo, size=(n_sing, 2))]]
    return zero, sing


@pytest.fixture(scope='module')
def oci(oracle_mod):
    import oracle_ci
    return oracle_ci


def test_port_equals_reference_cy_ci_bitwise(oci):
    if not oci.have_ref():
        pytest.skip('oracle/_ref/cy_ci not built')
    rng = numpy.random.default_rng(11)
    for n_mo, npts, n_det, n_sing in ((3, 1, 0, 1), (7, 33, 4, 19), (20, 257, 30, 400)):
        zero, sing = random_lists(rng, n_mo, n_det, n_sing)
        mo = rng.normal(size=(n_mo, npts))
        dmo = rng.normal(size=(3, n_mo, npts))
        assert numpy.array_equal(oci.rho(zero, sing, mo, kind='port'), oci.rho(zero, sing, mo, kind='ref'))
        assert numpy.array_equal(oci.jab(zero, sing, mo, dmo, kind='port'), oci.jab(zero, sing, mo, dmo, kind='ref'))
        assert numpy.array_equal(oci.a_nabla_b(zero, sing, mo, dmo, kind='port'),
                                 oci.a_nabla_b(zero, sing, mo, dmo, kind='ref'))


def test_h3p_golden_reproduced(oci, oracle_mod):
    """MOs of the H3+ FCI fixture from the (pinned) grid oracle -> CI contractions == what the reference
    produced for its own test, bit for bit; and == the reference's committed golden to 1e-15."""
    from orbkit_b200 import QCinfo
    g = load_golden('h3p_detci')
    qc = QCinfo.from_arrays(g)
    mos = oracle_mod.rho_compute(qc, g['x'], g['y'], g['z'], is_vector=False, calc_mo=True,
                                 drv=[None, 'x', 'y', 'z', 'xx', 'yy', 'zz'])
    mo, d1, d2 = mos[0], mos[1:4], mos[4:7]
    kinds = ['port'] + (['ref'] if oci.have_ref() else [])
    for pair in range(int(g['n_pairs'])):
        zero, sing = lists_from_golden(g, pair)
        for kind in kinds:
            # slice_length=1e2 as in the reference's test: 2601 points -> the last one is never visited (0.0)
            assert numpy.array_equal(oci.rho(zero, sing, mo, 1e2, kind=kind), g['rho_01'][pair])
            assert numpy.array_equal(oci.jab(zero, sing, mo, d1, 1e2, kind=kind), g['j_01'][pair])
            assert numpy.array_equal(-oci.jab(zero, sing, mo, d2, 1e2, kind=kind).sum(axis=0), g['nabla_j_01'][pair])
            assert numpy.array_equal(oci.a_nabla_b(zero, sing, mo, d1, 1e2, kind=kind), g['a_nabla_b_01'][pair])
            assert g['rho_01'][pair].reshape(-1)[-1] == 0.0 and oci.rho(zero, sing, mo, kind=kind).reshape(-1)[-1] != 0.0
    for key in ('rho_01', 'j_01', 'nabla_j_01'):
        pub = g['published.' + key]
        assert numpy.abs(g[key] - pub).max() <= 1e-15 * numpy.abs(pub).max()


def test_flatten_and_merge_host_logic():
    from orbkit_b200.detci import ci_core
    zero = [[[1.0, 2.0], [0.5]], [[0, 1], [1]]]
    sing = [[0.25, -0.75, 0.5], [[0, 2], [2, 0], [1, 1]]]
    c, a, b = ci_core.flatten_terms(zero, sing)
    assert c.tolist() == [1.0, 2.0, 0.5, 0.25, -0.75, 0.5]
    assert a.tolist() == [0, 1, 1, 0, 2, 1] and b.tolist() == [0, 1, 1, 2, 0, 1]
    c2, a2, b2 = ci_core.flatten_terms(zero, sing, with_zero=False)
    assert c2.tolist() == [0.25, -0.75, 0.5] and a2.tolist() == [0, 2, 1]
    mc, ma, mb = ci_core.merge_terms((c, a, b), 3, symmetric=True)
    dense = numpy.zeros((3, 3))
    for cc, aa, bb in zip(mc, ma, mb):
        dense[aa, bb] += cc
    want = numpy.zeros((3, 3))
    for cc, aa, bb in zip(c, a, b):
        want[min(aa, bb), max(aa, bb)] += cc
    assert numpy.allclose(dense, want) and len(mc) == 3
    mc, ma, mb = ci_core.merge_terms((c, a, b), 3, symmetric=False)
    assert len(mc) == 4
    with pytest.raises(ValueError):
        ci_core.flatten_terms([[[1.0]], [[0, 1]]], sing)
    with pytest.raises(ValueError):
        ci_core.flatten_terms(zero, [[1.0], []])


def test_mo_matrix_and_jmo_golden_reproduced(oci, oracle_mod):
    """core.calc_mo_matrix / extras.calc_jmo restated in oracle_ci == the reference's own outputs
    (tests/golden/h2o_mo_matrix.npz, written by running the reference: make_golden_jmo.py), bit for bit."""
    from conftest import golden_qc
    qc, g = golden_qc('h2o_mo_matrix')
    x, y, z = g['grid.x'], g['grid.y'], g['grid.z']
    for kind in ['port'] + (['ref'] if oracle_mod.have_ref() else []):
        assert numpy.array_equal(oci.calc_mo_matrix(qc, x, y, z, drv=['x', 'y', 'z'], kind=kind), g['mm_xyz'])
        assert numpy.array_equal(oci.calc_mo_matrix(qc, x, y, z, kind=kind), g['mm_none'])
        assert numpy.array_equal(oci.calc_mo_matrix(qc, x, y, z, drv='xx', kind=kind), g['mm_xx'])
        assert numpy.array_equal(oci.calc_jmo(qc, g['ij'], x, y, z, kind=kind), g['jmo'])
        assert numpy.array_equal(oci.calc_jmo(qc, g['ij'], x, y, z, drv=['z', 'x'], kind=kind), g['jmo_zx'])
        assert numpy.array_equal(oci.calc_jmo(qc, [4, 1], x, y, z, kind=kind), g['jmo_one'])
    assert g['jmo'].shape == (3, 5, 5, 6, 7) and (g['jmo'][:, 3] == 0.0).all()     # the pair (3, 3) has no flux


def _td_case(rng, nt, ns, npts):
    npair = ns * (ns + 1) // 2
    return (rng.normal(size=(nt, ns, ns)), rng.normal(size=(nt, ns, ns)), rng.normal(size=(npair, npts)),
            rng.normal(size=(npair, 3, npts)))


def test_time_dependent_port_equals_reference_bitwise(oci):
    """get_rho_full / get_j_full / get_jab_full (cy_ci.pyx:101-151, 186-202): the C restatement against the reference's
    own compiled module"""
    if not oci.have_ref():
        pytest.skip('oracle/_ref/cy_ci not built')
    rng = numpy.random.default_rng(5)
    for nt, ns, npts in ((1, 1, 1), (3, 2, 17), (9, 5, 130), (40, 3, 64)):
        ReS, ImS, rho, j = _td_case(rng, nt, ns, npts)
        assert numpy.array_equal(oci.get_rho_full(ReS, rho), oci.get_rho_full(ReS, rho, kind='ref'))
        assert numpy.array_equal(oci.get_j_full(ImS, j), oci.get_j_full(ImS, j, kind='ref'))
    for nb, nc, npts in ((1, 3, 5), (2, 1, 33), (6, 3, 129), (11, 2, 40)):
        S, chi, dchi = rng.normal(size=(nb, nb)), rng.normal(size=(nb, npts)), rng.normal(size=(nc, nb, npts))
        assert numpy.array_equal(oci.get_jab_full(S, chi, dchi, 1836.15), oci.get_jab_full(S, chi, dchi, 1836.15, kind='ref'))


def test_time_dependent_host_logic(oci):
    """orbkit_b200.detci.cy_ci: the packed pair weights reproduce the reference's loops as a plain matrix product
    (to rounding: the product sums in another order), and the typed-buffer argument checks of the compiled module"""
    from orbkit_b200.detci import cy_ci
    rng = numpy.random.default_rng(6)
    ReS, ImS, rho, j = _td_case(rng, 7, 4, 50)
    ref = oci.get_rho_full(ReS, rho)
    got = cy_ci.pair_weights_rho(ReS) @ rho
    assert numpy.abs(got - ref).max() <= 1e-13 * numpy.abs(ref).max()
    ref = oci.get_j_full(ImS, j)
    got = (cy_ci.pair_weights_j(ImS) @ j.reshape(j.shape[0], -1)).reshape(ref.shape)
    assert numpy.abs(got - ref).max() <= 1e-13 * numpy.abs(ref).max()
    with pytest.raises(ValueError):
        cy_ci.get_rho_full(ReS.astype(numpy.float32), rho)
    with pytest.raises(ValueError):
        cy_ci.get_rho_full(ReS, rho.T)                       # not C-contiguous
    with pytest.raises(ValueError):
        cy_ci.get_j_full(ImS, j[:, 0])                       # wrong number of dimensions
    with pytest.raises(TypeError):
        cy_ci.get_jab_full(None, rho, j, 1.0)
    assert cy_ci.get_rho_full(ReS[:0], rho).shape == (0, 50)
