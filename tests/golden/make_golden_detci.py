"""Golden fixture for the detCI grid contractions, written by RUNNING THE REFERENCE.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_detci.py

Re-runs the reference's own test orbkit/test/detci/h3+.py up to the grid quantities (psi4 FCI of
H3+, spherical cc-pVTZ, three states -> three state pairs) inside a scratch copy of the reference
with its extensions built (cy_grid, cy_core, cy_overlap, detci/cy_occ_check, detci/cy_ci), checks
the results against the reference's golden refdata_h3+.npz (agreement to 1e-15 relative; the committed
file was produced on another compiler) and writes
tests/golden/h3p_detci.npz:
    QCinfo as flat arrays + the grid,
    per state pair p: flattened (zero, sing) lists  p<i>.zc/zi/zn (zn = entries per determinant), p<i>.sc/sa/sb,
    the reference's outputs rho_01, j_01, nabla_j_01 as computed here, and `published.*` = refdata_h3+.npz,
    a_nabla_b of every pair (not in the reference golden; computed by the reference here).
"""
import os
import subprocess
import sys

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg     # noqa: E402

SETUP_CI = r'''
from setuptools import setup, Extension
from Cython.Build import cythonize
import numpy
I = [numpy.get_include(), 'orbkit']
exts = [Extension('orbkit.detci.cy_occ_check', ['orbkit/detci/cy_occ_check.pyx'], include_dirs=I),
        Extension('orbkit.detci.cy_ci', ['orbkit/detci/cy_ci.pyx'], include_dirs=I,
                  extra_compile_args=['-fopenmp'], extra_link_args=['-fopenmp'])]
setup(name='okci', ext_modules=cythonize(exts, language_level=3))
'''


def main():
    scratch = mg.build_reference()
    import glob
    if not glob.glob(os.path.join(scratch, 'orbkit', 'detci', 'cy_ci*.so')):
        with open(os.path.join(scratch, 'setup_ci.py'), 'w') as f:
            f.write(SETUP_CI)
        env = dict(os.environ, CC='/usr/bin/gcc', LDSHARED='/usr/bin/gcc -shared')
        subprocess.check_call([sys.executable, 'setup_ci.py', 'build_ext', '--inplace'], cwd=scratch, env=env,
                              stdout=subprocess.DEVNULL)
    sys.path.insert(0, scratch)
    mg.shim()
    from orbkit import grid, main_read, core, options, detci
    options.quiet = True
    options.no_log = True
    tdir = os.path.join(scratch, 'orbkit', 'test', 'detci')
    file_ci = os.path.join(tdir, 'read', 'outputs_for_testing', 'h3+_fci_cc-pVTZ.out')
    if not os.path.exists(file_ci):
        file_ci = glob.glob(os.path.join(scratch, 'orbkit', 'test', '**', 'h3+_fci_cc-pVTZ.out'), recursive=True)[0]
    qc = main_read(file_ci + '.default.molden', all_mo=True)
    qc, ci = detci.ci_read.main_ci_read(qc, file_ci, itype='psi4_detci', threshold=0.0)

    grid.min_ = [-2.5, -2.5, 0.0]
    grid.max_ = [2.5, 2.5, 0.0]
    grid.delta_ = [0.1, 0.1, 0.1]
    grid.grid_init()
    gx, gy, gz = grid.x.copy(), grid.y.copy(), grid.z.copy()
    molist = core.rho_compute(qc, calc_mo=True, slice_length=1e2, drv=[None, 'x', 'y', 'z', 'xx', 'yy', 'zz'],
                              numproc=1)
    molistdrv, molistdrv2, mo = molist[1:4], molist[-3:], molist[0]

    out = mg.qc_arrays(qc)
    out.update(x=gx, y=gy, z=gz)
    rho_01, j_01, nabla_j_01, anb = [], [], [], []
    pair = 0
    for a in range(len(ci)):
        for b in range(a + 1, len(ci)):
            zero, sing = detci.occ_check.compare(ci[a], ci[b], numproc=1)
            rho_01.append(detci.ci_core.rho(zero, sing, mo, slice_length=1e2, numproc=1))
            j_01.append(detci.ci_core.jab(zero, sing, mo, molistdrv, slice_length=1e2, numproc=1))
            nabla_j_01.append(-numpy.sum(detci.ci_core.jab(zero, sing, mo, molistdrv2, slice_length=1e2, numproc=1),
                                         axis=0))
            anb.append(detci.ci_core.a_nabla_b(zero, sing, mo, molistdrv, slice_length=1e2, numproc=1))
            out['p%d.zc' % pair] = numpy.array([c for cs in zero[0] for c in cs], dtype=float)
            out['p%d.zi' % pair] = numpy.array([i for idx in zero[1] for i in idx], dtype=numpy.intc)
            out['p%d.zn' % pair] = numpy.array([len(cs) for cs in zero[0]], dtype=numpy.intc)
            out['p%d.sc' % pair] = numpy.array(sing[0], dtype=float)
            out['p%d.sa' % pair] = numpy.array([p[0] for p in sing[1]], dtype=numpy.intc)
            out['p%d.sb' % pair] = numpy.array([p[1] for p in sing[1]], dtype=numpy.intc)
            print('pair %d (%s -> %s): %d zero entries in %d determinants, %d singles' % (
                pair, ci[a].info['state'], ci[b].info['state'], len(out['p%d.zc' % pair]), len(zero[0]), len(sing[0])))
            pair += 1
    ref = numpy.load(os.path.join(tdir, 'refdata_h3+.npz'))
    for key, mine in (('rho_01', rho_01), ('j_01', j_01), ('nabla_j_01', nabla_j_01)):
        d = numpy.abs(numpy.array(mine) - ref[key]).max()
        print('reference golden %-11s max abs diff %.3e (max |ref| %.3e)' % (key, d, numpy.abs(ref[key]).max()))
        # the reference's refdata_h3+.npz was written on another machine/compiler: last-bit differences
        # (1e-17 absolute); the reference's own test accepts rtol 1e-3 / atol 1e-5 (test/tools.py:11-21)
        assert d <= 1e-15 * numpy.abs(ref[key]).max()
        out[key] = numpy.array(mine)          # what the reference computes HERE (bit-exact oracle target)
        out['published.' + key] = ref[key]    # the reference's committed golden
    out['a_nabla_b_01'] = numpy.array(anb)
    out['n_pairs'] = numpy.array(pair)
    numpy.savez_compressed(os.path.join(HERE, 'h3p_detci.npz'), **out)
    print('wrote h3p_detci.npz: n_ao=%d n_mo=%d grid %s' % (qc.ao_spec.get_ao_num(), len(qc.mo_spec), mo.shape[1:]))


if __name__ == '__main__':
    main()
