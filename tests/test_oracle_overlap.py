"""CPU: the oracle of the analytic overlap integrals (oracle/oracle_overlap.py) is pinned to the reference, and the host
logic around it (orbkit_b200.cy_overlap norm factors, the Molden CCA renormalisation of orbkit_b200.read).

  * port (okor_aooverlap / okor_cca_norm / okor_mooverlapmatrix in libokoracle.so) == the reference's own cy_overlap
    module (oracle/_ref) bit for bit on random contractions, all derivative codes;
  * port reproduces tests/golden/overlap.npz, written by running the reference's analytical_integrals.get_ao_overlap on
    its Gaussian test output (make_golden_overlap.py), bit for bit;
  * read_molden on a CCA-normalised Cartesian [6D] file == the reference reader's QCinfo (same fixture).
"""
import os

import numpy
import pytest

from conftest import load_golden


@pytest.fixture(scope='module')
def oo(oracle_mod):
    import oracle_overlap
    return oracle_overlap


def random_basis(rng, n_atom=3, n_cont=6, lmax=3):
    """contractions in the standard Cartesian order of tools.exp"""
    from orbkit_b200.tools import exp
    assign, pnum, atoms, lx, co = [], [], [], [], []
    for _ in range(n_cont):
        l = int(rng.integers(0, lmax + 1))
        fns = exp[l]
        assign.append(len(fns))
        lx += [list(f) for f in fns]
        k = int(rng.integers(1, 4))
        pnum.append(k)
        atoms.append(int(rng.integers(0, n_atom)))
        co += [[float(10 ** rng.uniform(-1, 1.5)), float(rng.uniform(0.1, 1.0))] for _ in range(k)]
    i = lambda v: numpy.array(v, dtype=numpy.intc)
    return (rng.normal(size=(n_atom, 3)) * 1.5, i(lx), i(assign), numpy.array(co), i(pnum), i(atoms))


def test_port_equals_reference_cy_overlap_bitwise(oo):
    if not oo.have_ref():
        pytest.skip('oracle/_ref/cy_overlap not built')
    rng = numpy.random.default_rng(21)
    for trial in range(4):
        geo, lx, assign, co, pnum, atoms = random_basis(rng, n_cont=3 + 2 * trial)
        geo_b = geo + 0.2 * rng.normal(size=geo.shape)
        for drv in (0, 1, 2, 3):
            for isn in (0, 1):
                a = oo.aooverlap(geo, geo_b, lx, lx, assign, co, pnum, atoms, drv, isn)
                b = oo.aooverlap(geo, geo_b, lx, lx, assign, co, pnum, atoms, drv, isn, kind='ref')
                assert numpy.array_equal(a, b), (trial, drv, isn)
        assert numpy.array_equal(oo.ommited_cca_norm(lx), oo.ommited_cca_norm(lx, kind='ref'))
        assert numpy.array_equal(oo.ommited_cca_norm(lx, with_divisor=False), oo.ommited_cca_norm(lx, kind='ref', with_divisor=False))
        s = oo.aooverlap(geo, geo, lx, lx, assign, co, pnum, atoms, 0, 0)
        ma, mb = rng.normal(size=(4, len(lx))), rng.normal(size=(3, len(lx)))
        assert numpy.array_equal(oo.mooverlapmatrix(ma, mb, s), oo.mooverlapmatrix(ma, mb, s, kind='ref'))
    # symmetry and unit diagonal of a normalised single-primitive basis
    geo, lx, assign, co, pnum, atoms = random_basis(rng, n_cont=5)
    s = oo.aooverlap(geo, geo, lx, lx, assign, co, pnum, atoms, 0, 0)
    assert numpy.allclose(s, s.T, atol=1e-14)


def _qc_overlap(oo, g, prefix='', kind='port', drv=0):
    return oo.aooverlap(g[prefix + 'geo_spec'], g[prefix + 'geo_spec'], g[prefix + 'ao._lxlylz'], g[prefix + 'ao._lxlylz'],
                        g[prefix + 'ao._nlxlylz_per_cont'], g[prefix + 'ao._prim_coeffs'], g[prefix + 'ao._nprim_per_cont'],
                        g[prefix + 'ao._assign_cont_to_atoms'], drv, int(bool(g[prefix + 'ao.normalized'])), kind=kind)


def test_oracle_reproduces_the_reference_overlap_golden(oo):
    gold, qc = load_golden('overlap'), load_golden('read_fchk')
    for key, drv in (('h2o_cart.S', 0), ('h2o_cart.Sx', 1), ('h2o_cart.Sz', 3)):
        assert numpy.array_equal(_qc_overlap(oo, qc, 'cart.', drv=drv), gold[key]), key
    moom = oo.mooverlapmatrix(qc['cart.mo.coeffs'], qc['cart.mo.coeffs'], gold['h2o_cart.S'])
    assert numpy.array_equal(moom, gold['h2o_cart.moom'])
    assert abs(numpy.linalg.norm(moom - numpy.eye(len(moom))) - float(gold['h2o_cart.dev'])) < 1e-15


def test_norm_factors_and_cca_molden_reader(oo, tmp_path):
    from orbkit_b200 import cy_overlap, read, options
    from orbkit_b200.tools import exp
    options.quiet = True
    lx = numpy.array([list(f) for l in range(5) for f in exp[l]], dtype=numpy.intc)
    assert numpy.array_equal(cy_overlap.ommited_cca_norm(lx), oo.ommited_cca_norm(lx))
    assert numpy.array_equal(cy_overlap.tmol_aomix_norm(lx), oo.ommited_cca_norm(lx, with_divisor=False))
    with pytest.raises(ValueError):
        cy_overlap.ommited_cca_norm(lx.astype(numpy.int64))
    gold = load_golden('overlap')
    path = os.path.join(str(tmp_path), 'cca_test.molden')
    with open(path, 'wb') as f:
        f.write(gold['cca.file'].tobytes())
    qc = read.main_read(path, all_mo=True)
    from test_host import _flat_qc
    for k, v in _flat_qc(qc).items():
        ref = gold['cca.' + k]
        assert v.shape == ref.shape, k
        if v.dtype.kind == 'f':
            # the contractions' self overlaps come from a closed form here, from the recursion in the reference
            assert numpy.abs(v - ref).max() <= 4e-16 * max(1.0, numpy.abs(ref).max()), k
        else:
            assert (v == ref).all(), k
    assert abs(qc.mo_spec[0]['coeffs'][4] / -1.95103519 - 1) < 1e-8          # the CCA factor reached the d coefficients
