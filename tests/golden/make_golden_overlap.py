"""Golden fixtures for the analytic overlap integrals and the Molden CCA renormalisation, written by RUNNING THE
REFERENCE (orbkit/analytical_integrals.py, orbkit/cy_overlap.pyx, orbkit/read/molden.py).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_overlap.py

Writes tests/golden/overlap.npz:
    h2o_cart.S / .Sx / .Sz   get_ao_overlap of h2o_rhf_cart.fchk (Cartesian d), plain and drv='x', 'z'
    h2o_cart.moom            get_mo_overlap_matrix(mo_spec, mo_spec, S)
    h2o_cart.dev             check_mo_norm-style deviation ||moom - 1||
    lih_sph.S / .dev         the same for the Psi4 LiH aug-cc-pVTZ Molden file (real-spherical d and f shells)
    cca.<flat QCinfo>        the reference reader's QCinfo of the synthetic Molden file below
    cca.file                 that file's bytes: water in a Cartesian [6D] basis whose d functions are normalised as the
                             CCA standard prescribes, i.e. the Cartesian factors are omitted
(the QCinfo inputs come from read_fchk.npz / lih_psi4_sph_f.npz / reader_inputs.npz, i.e. the same files)
"""
import os
import sys

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg     # noqa: E402

CCA_MOLDEN = """[Molden Format]
[Title]
 synthetic CCA-normalised Cartesian d test (orbkit_b200 tests)
[Atoms] AU
 O     1    8     0.000000     0.000000     0.221000
 H     2    1     0.000000     1.431000    -0.884000
 H     3    1     0.000000    -1.431000    -0.884000
[GTO]
  1 0
 s   3 1.00
      130.7093200        0.15432897
       23.8088610        0.53532814
        6.4436083        0.44463454
 p   2 1.00
        5.0331513        0.15591627
        1.1695961        0.60768372
 d   1 1.00
        1.2000000        1.00000000

  2 0
 s   2 1.00
        3.4252509        0.15432897
        0.6239137        0.53532814

  3 0
 s   2 1.00
        3.4252509        0.15432897
        0.6239137        0.53532814

[6D]
[MO]
 Sym= 1a
 Ene= -20.25
 Spin= Alpha
 Occup= 2.0
%s
 Sym= 2a
 Ene= -1.26
 Spin= Alpha
 Occup= 2.0
%s
 Sym= 3a
 Ene= 0.30
 Spin= Alpha
 Occup= 0.0
%s
"""


def mo_block(rng, n):
    return '\n'.join(' %3d  %14.8f' % (i + 1, v) for i, v in enumerate(rng.normal(size=n)))


def main():
    scratch = mg.build_reference()
    sys.path.insert(0, scratch)
    mg.shim()
    from orbkit import options, read
    from orbkit.analytical_integrals import get_ao_overlap, get_mo_overlap_matrix
    options.quiet = True
    options.no_log = True
    odir = os.path.join(scratch, 'orbkit', 'test', 'outputs_for_testing')
    out = {}
    for name, rel, kw in [('h2o_cart', 'gaussian/h2o_rhf_cart.fchk', dict(all_mo=True)),
                          ('lih_sph', 'psi4/lih_cis_aug-cc-pVTZ.out.default.molden', dict(all_mo=True))]:
        qc = read.main_read(os.path.join(odir, rel), **kw)
        s = get_ao_overlap(qc.geo_spec, qc.geo_spec, qc.ao_spec)
        out[name + '.S'] = s
        if name == 'h2o_cart':
            out[name + '.Sx'] = get_ao_overlap(qc.geo_spec, qc.geo_spec, qc.ao_spec, drv='x')
            out[name + '.Sz'] = get_ao_overlap(qc.geo_spec, qc.geo_spec, qc.ao_spec, drv='z')
        moom = get_mo_overlap_matrix(qc.mo_spec, qc.mo_spec, s)
        out[name + '.moom'] = moom
        out[name + '.dev'] = numpy.array(numpy.linalg.norm(moom - numpy.eye(len(moom))))
        print(name, s.shape, 'deviation from orthonormality %.3e' % out[name + '.dev'])
    rng = numpy.random.default_rng(42)
    n_basis = 1 + 3 + 6 + 1 + 1
    text = CCA_MOLDEN % (mo_block(rng, n_basis), mo_block(rng, n_basis), mo_block(rng, n_basis))
    path = os.path.join(scratch, 'cca_test.molden')
    with open(path, 'w') as f:
        f.write(text)
    qc = read.main_read(path, all_mo=True)
    for k, v in mg.qc_arrays(qc).items():
        out['cca.' + k] = v
    out['cca.file'] = numpy.frombuffer(text.encode(), dtype=numpy.uint8)
    print('cca:', len(qc.mo_spec), 'MOs', qc.ao_spec.get_ao_num(), 'AOs; first d coefficients', qc.mo_spec[0]['coeffs'][4:10])
    numpy.savez_compressed(os.path.join(HERE, 'overlap.npz'), **out)


if __name__ == '__main__':
    main()
