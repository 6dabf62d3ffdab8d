"""Golden fixture for core.calc_mo_matrix / extras.calc_jmo, written by RUNNING THE REFERENCE.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_jmo.py

Writes tests/golden/h2o_mo_matrix.npz:
    QCinfo of outputs_for_testing/gaussian/h2o_rhf_sph.fchk (occupied MOs) as flat arrays + a ragged 5x6x7 grid,
    mm_xyz   = core.calc_mo_matrix(qc, drv=['x','y','z'])       (3, NMO, NMO, 5, 6, 7)   core.py:841-941
    mm_none  = core.calc_mo_matrix(qc)                           (1, NMO, NMO, ...)
    mm_xx    = core.calc_mo_matrix(qc, drv='xx')                 (1, NMO, NMO, ...)
    ij, jmo  = extras.calc_jmo(qc, ij)                           (3, len(ij), ...)        extras.py:441-493
    jmo_zx   = extras.calc_jmo(qc, ij, drv=['z','x'])
    jmo_one  = extras.calc_jmo(qc, [4, 1])                       a single pair given as a flat list

The two-QCinfo branch of calc_mo_matrix (qc_b is not qc_a, core.py:903-918) cannot be run: it indexes a list with a
list (`drv[ibra]`, core.py:906) and raises TypeError for every input, so it has no reference output to pin.
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
    from orbkit import grid, options, read, core, extras
    options.quiet = True
    options.no_log = True
    options.no_output = True
    odir = os.path.join(scratch, 'orbkit', 'test', 'outputs_for_testing')
    qc = read.main_read(os.path.join(odir, 'gaussian', 'h2o_rhf_sph.fchk'), all_mo=False)
    x, y, z = numpy.linspace(-2, 2, 5), numpy.linspace(-2.1, 2, 6), numpy.linspace(-1.5, 2.2, 7)
    grid.x, grid.y, grid.z = x.copy(), y.copy(), z.copy()
    grid.N_ = [5, 6, 7]
    grid.is_initialized, grid.is_regular, grid.is_vector = True, True, False
    out = mg.qc_arrays(qc)
    out.update({'grid.x': x, 'grid.y': y, 'grid.z': z})
    out['mm_xyz'] = core.calc_mo_matrix(qc, drv=['x', 'y', 'z'])
    out['mm_none'] = core.calc_mo_matrix(qc)
    out['mm_xx'] = core.calc_mo_matrix(qc, drv='xx')
    ij = numpy.array([[0, 1], [2, 4], [4, 2], [3, 3], [1, 0]])
    out['ij'] = ij
    out['jmo'] = extras.calc_jmo(qc, ij.copy())
    out['jmo_zx'] = extras.calc_jmo(qc, ij.copy(), drv=['z', 'x'])
    out['jmo_one'] = extras.calc_jmo(qc, [4, 1])
    for k in ('mm_xyz', 'mm_none', 'mm_xx', 'jmo', 'jmo_zx', 'jmo_one'):
        print(k, out[k].shape, float(numpy.abs(out[k]).max()))
    numpy.savez_compressed(os.path.join(HERE, 'h2o_mo_matrix.npz'), **out)


if __name__ == '__main__':
    main()
