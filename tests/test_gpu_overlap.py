"""GPU parity of the analytic overlap integrals (orbkit_b200.cy_overlap.aooverlap -> okb_aooverlap, csrc/okb_overlap.cuh;
orbkit_b200.analytical_integrals) against the pinned CPU oracle and the goldens written by running the reference.
The kernel walks the primitive pairs in the reference's order with the reference's recursion; only exp / pow are the
device's: tolerance |d| <= 1e-12 max|ref| (measured ~1e-15)."""
import os

import numpy
import pytest

from conftest import load_golden, reader_input
from test_oracle_overlap import random_basis, _qc_overlap

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def oo(oracle_mod):
    import oracle_overlap
    return oracle_overlap


def test_aooverlap_vs_oracle(oo):
    from orbkit_b200 import cy_overlap
    kind = 'ref' if oo.have_ref() else 'port'
    rng = numpy.random.default_rng(31)
    for trial, (n_cont, lmax) in enumerate(((1, 0), (4, 2), (9, 3), (14, 4))):
        geo, lx, assign, co, pnum, atoms = random_basis(rng, n_cont=n_cont, lmax=lmax)
        geo_b = geo + 0.3 * rng.normal(size=geo.shape)
        for drv in (0, 1, 2, 3):
            for isn in (0, 1):
                ref = oo.aooverlap(geo, geo_b, lx, lx, assign, co, pnum, atoms, drv, isn, kind=kind)
                got = cy_overlap.aooverlap(geo, geo_b, lx, lx, assign, co, pnum, atoms, drv, isn)
                assert got.shape == ref.shape
                assert numpy.abs(got - ref).max() <= 1e-12 * max(numpy.abs(ref).max(), 1e-300), (trial, drv, isn)
        s = oo.aooverlap(geo, geo, lx, lx, assign, co, pnum, atoms, 0, 0, kind=kind)
        ma, mb = rng.normal(size=(5, len(lx))), rng.normal(size=(3, len(lx)))
        ref = oo.mooverlapmatrix(ma, mb, s, kind=kind)
        got = cy_overlap.mooverlapmatrix(ma, mb, s, 0, 5)
        assert numpy.abs(got - ref).max() <= 1e-12 * numpy.abs(ref).max()
        assert abs(cy_overlap.mooverlap(ma[1], mb[2], s) - ref[1, 2]) <= 1e-12 * numpy.abs(ref).max()
        assert cy_overlap.mooverlapmatrix(ma, mb, s, 2, 4).shape == (2, 3)
    with pytest.raises(ValueError):
        cy_overlap.aooverlap(geo.astype(numpy.float32), geo, lx, lx, assign, co, pnum, atoms, 0, 0)


def test_get_ao_overlap_and_check_norm_vs_reference_goldens(tmp_path):
    """analytical_integrals.get_ao_overlap on the reference's QCinfo of its Gaussian (Cartesian d) and Psi4 (spherical d, f)
    test outputs == the matrices the reference computed (tests/golden/overlap.npz); main_read(check_norm=True)"""
    import orbkit_b200 as ok
    from orbkit_b200 import analytical_integrals as ai
    ok.options.quiet = True
    gold = load_golden('overlap')
    g = load_golden('read_fchk')
    qc = ok.QCinfo.from_arrays({k[5:]: g[k] for k in g.keys() if k.startswith('cart.')})
    for key, drv in (('h2o_cart.S', None), ('h2o_cart.Sx', 'x'), ('h2o_cart.Sz', 'z')):
        got = ai.get_ao_overlap(qc.geo_spec, qc.geo_spec, qc.ao_spec, drv=drv)
        assert numpy.abs(got - gold[key]).max() <= 1e-12, key
    both = ai.get_ao_overlap(qc.geo_spec, qc.geo_spec, qc.ao_spec, drv=['x', 'z'])
    assert len(both) == 2 and numpy.abs(both[1] - gold['h2o_cart.Sz']).max() <= 1e-12
    moom = ai.get_mo_overlap_matrix(qc.mo_spec, qc.mo_spec, gold['h2o_cart.S'])
    assert numpy.abs(moom - gold['h2o_cart.moom']).max() <= 1e-12
    assert abs(ai.check_mo_norm(qc) - float(gold['h2o_cart.dev'])) <= 1e-11
    assert abs(ai.get_mo_overlap(qc.mo_spec[0], qc.mo_spec[1], gold['h2o_cart.S']) - gold['h2o_cart.moom'][0, 1]) <= 1e-12
    with pytest.raises(ValueError):
        ai.get_ao_overlap(qc.geo_spec, qc.geo_spec, qc.ao_spec, drv=4)      # only first derivatives (code > 3)
    assert len(ai.get_ao_overlap(qc.geo_spec, qc.geo_spec, qc.ao_spec, drv='xz')) == 2   # a string of letters is a list
    with pytest.raises(TypeError):
        ai.get_ao_overlap(qc.geo_spec, qc.geo_spec, list(qc.ao_spec))
    # real-spherical d and f shells: T S T^T
    qs = ok.QCinfo.from_arrays(load_golden('lih_psi4_sph_f'))
    s = ai.get_ao_overlap(qs.geo_spec, qs.geo_spec, qs.ao_spec)
    assert s.shape == gold['lih_sph.S'].shape and numpy.abs(s - gold['lih_sph.S']).max() <= 1e-12
    # the high-level reader with the norm check (read/high_level.py:74-77)
    path = reader_input('h2o_rhf_cart.fchk', tmp_path)
    q2 = ok.main_read(path, all_mo=True, check_norm=True)
    assert len(q2.mo_spec) == len(qc.mo_spec)
    q2.mo_spec[0]['coeffs'][0] += 0.5
    q2.mo_spec.update()
    assert ai.check_mo_norm(q2) > 1e-2
