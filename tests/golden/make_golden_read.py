"""Golden fixture for the fchk reader, written by RUNNING THE REFERENCE's reader (orbkit/read/gaussian_fchk.py).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_read.py

Writes tests/golden/read_fchk.npz: per case `<c>.<flat QCinfo arrays>` (make_golden.qc_arrays) of
    cart      h2o_rhf_cart.fchk, all_mo=True          (6D 10F basis)
    uhf_beta  h2o_uhf_sph.fchk, all_mo=True, spin='beta'
    uhf_occ   h2o_uhf_sph.fchk, all_mo=False
and tests/golden/reader_inputs.npz: the bytes of the quantum-chemistry program outputs the readers are tested on
(`file.<name>`; Gaussian fchk, Molpro / Psi4 Molden files from the reference's test data orbkit/test/outputs_for_testing) --
the tests write them to a scratch directory and read them with orbkit_b200.read.
The spherical restricted / unrestricted cases are pinned by h2o_gaussian_sph*.npz / h2o_gaussian_uhf.npz (make_golden.py).
tests/golden/read_wf.npz: `wfx_beta.<flat QCinfo arrays>` of orca/1.wfx read with spin='beta'.
"""
import os
import sys

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg     # noqa: E402


def main():
    scratch = mg.build_reference()
    sys.path.insert(0, scratch)
    mg.shim()
    from orbkit import options, read
    options.quiet = True
    options.no_log = True
    gdir = os.path.join(scratch, 'orbkit', 'test', 'outputs_for_testing', 'gaussian')
    out = {}
    for name, fn, kw in [('cart', 'h2o_rhf_cart.fchk', dict(all_mo=True)),
                         ('uhf_beta', 'h2o_uhf_sph.fchk', dict(all_mo=True, spin='beta')),
                         ('uhf_occ', 'h2o_uhf_sph.fchk', dict(all_mo=False))]:
        qc = read.main_read(os.path.join(gdir, fn), **kw)
        for k, v in mg.qc_arrays(qc).items():
            out[name + '.' + k] = v
        out[name + '.etot'] = numpy.array(qc.etot)
        print(name, len(qc.mo_spec), 'MOs', qc.ao_spec.get_ao_num(), 'AOs')
    numpy.savez_compressed(os.path.join(HERE, 'read_fchk.npz'), **out)
    odir = os.path.join(scratch, 'orbkit', 'test', 'outputs_for_testing')
    # primitive-based wave-function files: the unrestricted ORCA .wfx file restricted to one spin (the full reads are
    # pinned by water_gamess_wfn.npz / h2o_orca_wfx.npz, make_golden.py)
    out = {}
    qc = read.main_read(os.path.join(odir, 'orca', '1.wfx'), all_mo=True, spin='beta')
    for k, v in mg.qc_arrays(qc).items():
        out['wfx_beta.' + k] = v
    print('wfx_beta', len(qc.mo_spec), 'MOs', qc.ao_spec.get_ao_num(), 'AOs')
    numpy.savez_compressed(os.path.join(HERE, 'read_wf.npz'), **out)
    # Gaussian .log files (GFINPUT + POP=FULL): restricted spherical, unrestricted Cartesian, the occupied orbitals of the unrestricted
    # spherical file, occupied orbitals only
    out = {}
    for name, fn, kw in [('rhf_sph', 'h2o_rhf_sph.inp.log', dict(all_mo=True)),
                         ('uhf_cart', 'h2o_uhf_cart.inp.log', dict(all_mo=True)),
                         ('uhf_sph_occ', 'h2o_uhf_sph.inp.log', dict(all_mo=False)),
                         ('rhf_cart_occ', 'h2o_rhf_cart.inp.log', dict(all_mo=False))]:
        qc = read.main_read(os.path.join(gdir, fn), interactive=False, **kw)
        for k, v in mg.qc_arrays(qc).items():
            out[name + '.' + k] = v
        out[name + '.etot'] = numpy.array(qc.etot)
        print('glog', name, len(qc.mo_spec), 'MOs', qc.ao_spec.get_ao_num(), 'AOs')
    numpy.savez_compressed(os.path.join(HERE, 'read_glog.npz'), **out)
    files = {}
    for rel in ['gaussian/h2o_rhf_sph.fchk', 'gaussian/h2o_uhf_sph.fchk', 'gaussian/h2o_rhf_cart.fchk',
                'molpro/h2o_rhf_sph.molden', 'molpro/nh3.mold', 'psi4/lih_cis_aug-cc-pVTZ.out.default.molden',
                'gaussian/h2o_rhf_sph.inp.log', 'gaussian/h2o_uhf_cart.inp.log', 'gaussian/h2o_uhf_sph.inp.log',
                'gaussian/h2o_rhf_cart.inp.log', 'gamess/water_gamess-us.wfn', 'orca/1.wfx', 'gamess/formaldehyde.log',
                'turbomole/h2o_rhf_sph/aomix.in']:
        with open(os.path.join(odir, rel), 'rb') as f:
            files['file.' + os.path.basename(rel)] = numpy.frombuffer(f.read(), dtype=numpy.uint8)
    numpy.savez_compressed(os.path.join(HERE, 'reader_inputs.npz'), **files)
    print('reader_inputs.npz:', sorted(files))


if __name__ == '__main__':
    main()
