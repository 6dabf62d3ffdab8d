"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

What it does
  1. copies /root/reference to a scratch dir, builds the reference's own extensions there
     (cy_grid, cy_core, cy_overlap) with a minimal setup script -- nothing is written to
     /root/reference or into this repo except the fixture files below;
  2. installs the tiny import shim the reference needs on modern numpy / without h5py
     (SURVEY.md Appendix B);
  3. self-checks the build against the reference's own golden refdata_rho_compute.npz
     (max abs diff must be 0.0) and the Gaussian cubegen cube files;
  4. writes, per fixture molecule,  tests/golden/<name>.npz  holding
        - the QCinfo as flat arrays (AOClass.todict()/MOClass.todict() content),
        - the grid used,
        - the reference's outputs: rho, delta_rho(x,y,z), laplacian set, all MOs x 10 derivative
          codes, all AOs x 10 derivative codes (small grids only);
  5. copies the reference's golden data / KAT cube values:
        ref_rho_compute.npz (zero..four), cube_kat.npz.

The synthetic benchmark molecules are produced by orbkit_b200.synth (pure numpy, seeded) and fed
through the reference's AOClass/MOClass, so the reference computes on identical inputs.
"""
import glob
import os
import shutil
import subprocess
import sys
import tempfile
import types

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
DRV10 = [None, 'x', 'y', 'z', 'xx', 'xy', 'xz', 'yy', 'yz', 'zz']

SETUP_MIN = r'''
from setuptools import setup, Extension
from Cython.Build import cythonize
import numpy
I = [numpy.get_include(), 'orbkit']
exts = [Extension('orbkit.cy_grid',    ['orbkit/cy_grid.pyx'], include_dirs=I),
        Extension('orbkit.cy_core',    ['orbkit/cy_core.pyx', 'orbkit/c_grid-based.c', 'orbkit/c_support.c'], include_dirs=I),
        Extension('orbkit.cy_overlap', ['orbkit/cy_overlap.pyx', 'orbkit/c_non-grid-based.c', 'orbkit/c_support.c'], include_dirs=I)]
setup(name='okmin', ext_modules=cythonize(exts, language_level=3))
'''


def build_reference():
    scratch = os.environ.get('OKREF_SCRATCH', os.path.join(tempfile.gettempdir(), 'okref'))
    if not glob.glob(os.path.join(scratch, 'orbkit', 'cy_core*.so')):
        if os.path.exists(scratch):
            shutil.rmtree(scratch)
        shutil.copytree(REF, scratch)
        subprocess.check_call(['chmod', '-R', 'u+w', scratch])
        with open(os.path.join(scratch, 'setup_min.py'), 'w') as f:
            f.write(SETUP_MIN)
        env = dict(os.environ, CC='/usr/bin/gcc', LDSHARED='/usr/bin/gcc -shared')
        subprocess.check_call([sys.executable, 'setup_min.py', 'build_ext', '--inplace'],
                              cwd=scratch, env=env, stdout=subprocess.DEVNULL)
    return scratch


def shim():
    for m in ['h5py', 'skimage', 'skimage.measure', 'matplotlib', 'matplotlib.pyplot']:
        if m not in sys.modules:
            try:
                __import__(m)
            except Exception:
                sys.modules[m] = types.ModuleType(m)
    if not hasattr(sys.modules['skimage'], 'measure'):
        sys.modules['skimage'].measure = sys.modules['skimage.measure']
    for name, val in [('product', numpy.prod), ('int', int), ('float', float), ('bool', bool)]:
        if not hasattr(numpy, name):
            setattr(numpy, name, val)


def qc_arrays(qc):
    """flat, pickle-free representation of a reference QCinfo"""
    ao = qc.ao_spec.todict()
    mo = qc.mo_spec.todict()
    out = {'geo_spec': numpy.asarray(qc.geo_spec, dtype=float),
           'geo_info': numpy.asarray(qc.geo_info, dtype=str)}
    for k in ['normalized', 'spherical', '_assign_cont_to_atoms', '_nprim_per_cont', '_prim_coeffs',
              '_assign_prim_to_cont', '_lxlylz', '_assign_lxlylz_to_cont', '_nlxlylz_per_cont']:
        out['ao.' + k] = numpy.asarray(ao[k])
    out['ao._cont_types'] = numpy.asarray(ao['_cont_types'], dtype=str)
    if ao['spherical']:
        out['ao._lm'] = numpy.asarray(ao['_lm'], dtype=numpy.intc)
        out['ao._assign_lm_to_cont'] = numpy.asarray(ao['_assign_lm_to_cont'])
    if 'N' in qc.ao_spec[0]:
        out['ao.N'] = numpy.asarray(qc.ao_spec[0]['N'])
    out['mo.coeffs'] = numpy.asarray(mo['coeffs'], dtype=float)
    out['mo.occ'] = numpy.asarray(mo['occ'], dtype=float)
    out['mo.eig'] = numpy.asarray(mo['eig'], dtype=float)
    out['mo.sym'] = numpy.asarray(mo['sym'], dtype=str)
    out['mo.spin'] = numpy.asarray(mo['spin'], dtype=str)
    return out


def main():
    scratch = build_reference()
    sys.path.insert(0, scratch)
    sys.path.insert(0, REPO)
    shim()
    import orbkit
    from orbkit import grid, options, read, core
    from orbkit.qcinfo import QCinfo
    from orbkit.orbitals import AOClass, MOClass
    options.quiet = True
    options.no_log = True
    tdir = os.path.join(scratch, 'orbkit', 'test')
    odir = os.path.join(tdir, 'outputs_for_testing')

    def set_regular(x, y, z):
        grid.x, grid.y, grid.z = [numpy.array(v, dtype=float) for v in (x, y, z)]
        grid.is_initialized, grid.is_regular, grid.is_vector = True, True, False
        grid.N_ = [len(grid.x), len(grid.y), len(grid.z)]

    def set_vector(x, y, z):
        grid.x, grid.y, grid.z = [numpy.array(v, dtype=float) for v in (x, y, z)]
        grid.is_initialized, grid.is_regular, grid.is_vector = True, False, True

    # ---- (3) self-check: reference golden + copy --------------------------------------
    qc = read.main_read(os.path.join(odir, 'molpro', 'h2o_rhf_sph.molden'), all_mo=True)
    grid.adjust_to_geo(qc, extend=2.0, step=1)
    grid.grid_init(is_vector=False, force=True)
    gx, gy, gz = grid.x.copy(), grid.y.copy(), grid.z.copy()
    ref = numpy.load(os.path.join(tdir, 'grid_based', 'refdata_rho_compute.npz'))
    mine = [core.rho_compute(qc, slice_length=0),
            core.rho_compute(qc, numproc=2),
            core.rho_compute(qc, laplacian=True, slice_length=0)[-1],
            core.rho_compute(qc, laplacian=True, numproc=2)[-1],
            core.rho_compute(qc, calc_mo=True, drv=DRV10, slice_length=0)]
    for key, arr in zip(['zero', 'one', 'two', 'three', 'four'], mine):
        d = numpy.abs(arr - ref[key]).max()
        print('reference golden %-5s max abs diff %.3e' % (key, d))
        assert d == 0.0
    # 'zero'=='one' and 'two'=='three' in the reference file; keep one of each (size)
    numpy.savez_compressed(os.path.join(HERE, 'ref_rho_compute.npz'),
                           zero=ref['zero'], two=ref['two'], four=ref['four'], x=gx, y=gy, z=gz)

    # ---- Gaussian cubegen known-answer values (test/grid_based/cube_files.py) ----------
    gdir = os.path.join(odir, 'gaussian')
    dx = 4.970736
    cx = numpy.arange(3) * dx - 4.970736
    cz = numpy.arange(3) * dx - 4.732975
    grad = numpy.genfromtxt(os.path.join(gdir, 'h2o_rhf_sph_grad.cube'), skip_header=9).reshape((-1,))
    krho = numpy.zeros((3, 3, 3))
    kdrho = numpy.zeros((3, 3, 3, 3))
    c = 0
    for i in range(3):
        for j in range(3):
            for k in range(3):
                krho[i, j, k] = grad[c]; c += 1
                for l in range(3):
                    kdrho[l, i, j, k] = grad[c]; c += 1
    klap = numpy.genfromtxt(os.path.join(gdir, 'h2o_rhf_sph_laplace.cube'), skip_header=9).reshape((3, 3, 3))
    kval = numpy.genfromtxt(os.path.join(gdir, 'h2o_rhf_sph.cube'), skip_header=9).reshape((3, 3, 3))
    kmo = numpy.genfromtxt(os.path.join(gdir, 'h2o_rhf_sph_mo.cube'), skip_header=10).reshape((1, 3, 3, 3))
    numpy.savez_compressed(os.path.join(HERE, 'cube_kat.npz'), x=cx, y=cx, z=cz, rho=krho, drho=kdrho,
                           laplacian=klap, rho_valence=kval, homo=kmo)

    # ---- (4) fixture molecules ---------------------------------------------------------
    rng = numpy.random.default_rng(1234)

    def dump(name, qc, full=True, npts_vec=48, box=3.0, extra=None):
        """reference outputs for `qc` on a small regular grid and a random vector grid"""
        arr = qc_arrays(qc)
        geo = numpy.asarray(qc.geo_spec, dtype=float)
        lo, hi = geo.min(axis=0) - box, geo.max(axis=0) + box
        rx, ry, rz = (numpy.linspace(lo[0], hi[0], 4), numpy.linspace(lo[1], hi[1], 5),
                      numpy.linspace(lo[2], hi[2], 3))
        vx, vy, vz = (rng.uniform(lo[i], hi[i], npts_vec) for i in range(3))
        arr.update(rx=rx, ry=ry, rz=rz, vx=vx, vy=vy, vz=vz)
        # regular grid, sliced driver
        set_regular(rx, ry, rz)
        arr['reg.rho'] = core.rho_compute(qc, numproc=1)
        r, d = core.rho_compute(qc, drv=['x', 'y', 'z'], numproc=1)
        arr['reg.drho'] = d
        r, d, l = core.rho_compute(qc, laplacian=True, numproc=1)
        arr['reg.d2rho'] = d
        arr['reg.lap'] = l
        # vector grid
        set_vector(vx, vy, vz)
        arr['vec.rho'] = core.rho_compute(qc, numproc=1)
        r, d = core.rho_compute(qc, drv=['x', 'y', 'z'], numproc=1)
        arr['vec.drho'] = d
        r, d, l = core.rho_compute(qc, laplacian=True, numproc=1)
        arr['vec.d2rho'] = d
        r, d = core.rho_compute(qc, drv=['xy', 'z', 'yz', 'x2'], numproc=1)
        arr['vec.dmixed'] = d
        if full:
            arr['vec.mo10'] = core.rho_compute(qc, calc_mo=True, drv=DRV10, numproc=1)
            arr['vec.ao10'] = core.rho_compute(qc, calc_ao=True, drv=DRV10, numproc=1)
        else:
            arr['vec.mo'] = core.rho_compute(qc, calc_mo=True, numproc=1)
            arr['vec.ao'] = core.rho_compute(qc, calc_ao=True, numproc=1)
            arr['vec.mo_z'] = core.rho_compute(qc, calc_mo=True, drv=['z'], numproc=1)[0]
            arr['vec.ao_yy'] = core.rho_compute(qc, calc_ao=True, drv=['yy'], numproc=1)[0]
        if extra:
            arr.update(extra)
        numpy.savez_compressed(os.path.join(HERE, name + '.npz'), **arr)
        print('wrote %-28s n_cart=%d n_ao=%d n_mo=%d sph=%s' % (
            name, len(qc.ao_spec.get_lxlylz()), qc.ao_spec.get_ao_num(), len(qc.mo_spec),
            qc.ao_spec.spherical))

    files = [
        ('h2o_molpro_cart', 'molpro/h2o_rhf_sph.molden', dict(all_mo=True)),
        ('h2o_gaussian_sph', 'gaussian/h2o_rhf_sph.fchk', dict(all_mo=True)),
        ('h2o_gaussian_uhf', 'gaussian/h2o_uhf_sph.fchk', dict(all_mo=True)),
        ('lih_psi4_sph_f', 'psi4/lih_cis_aug-cc-pVTZ.out.default.molden', dict(all_mo=True)),
        ('water_gamess_wfn', 'gamess/water_gamess-us.wfn', dict(all_mo=True)),
        ('h2o_orca_wfx', 'orca/1.wfx', dict(all_mo=True)),
        ('h2o_turbomole_aomix', 'turbomole/h2o_rhf_sph/aomix.in', dict(all_mo=True)),
        ('nh3_molpro', 'molpro/nh3.mold', dict(all_mo=True)),
    ]
    for name, rel, kw in files:
        path = os.path.join(odir, rel)
        if not os.path.exists(path):
            cands = glob.glob(os.path.join(odir, os.path.dirname(rel), '*' + os.path.splitext(rel)[1]))
            print('missing', rel, 'candidates', cands)
            continue
        try:
            qc = read.main_read(path, **kw)
        except Exception as e:  # a reader needing an absent dependency is not on the hot path
            print('skip %s: %r' % (rel, e))
            continue
        dump(name, qc)

    # h2o fchk occupied-only = config 1 input; also record the cube KAT inputs
    qc = read.main_read(os.path.join(gdir, 'h2o_rhf_sph.fchk'), all_mo=False)
    dump('h2o_gaussian_sph_occ', qc)

    # formaldehyde (160 AOs, explicit lxlylz, f functions): reduced output set
    try:
        qc = read.main_read(os.path.join(odir, 'gamess', 'formaldehyde.log'), all_mo=True)
        dump('formaldehyde_gamess', qc, full=False, npts_vec=24)
    except Exception as e:
        print('skip formaldehyde: %r' % (e,))

    # ---- synthetic molecules through the reference's own classes ------------------------
    from orbkit_b200 import synth

    def to_ref_qc(spec):
        qc = QCinfo()
        qc.geo_spec = numpy.array(spec['geo_spec'])
        qc.geo_info = numpy.array(spec['geo_info'])
        qc.ao_spec = AOClass([dict(d) for d in spec['ao_spec']])
        if spec['spherical']:
            qc.ao_spec.set_lm_dict(p=[1, 0])   # p order (1,1),(1,-1),(1,0)  (orbitals.py:303-316)
        qc.mo_spec = MOClass([dict(d) for d in spec['mo_spec']])
        qc.ao_spec.update()
        qc.mo_spec.update()
        return qc

    for name, kw, npv in [('synth_small_sph', dict(n_heavy=2, n_light=2, n_mo=9, seed=3, spherical=True), 40),
                          ('synth_small_cart_g', dict(n_heavy=1, n_light=1, n_mo=6, seed=5, spherical=False,
                                                      with_g=True), 40),
                          ('synth_c3', dict(n_heavy=24, n_light=20, n_mo=82, seed=0, spherical=True), 32)]:
        spec = synth.make_molecule(**kw)
        qc = to_ref_qc(spec)
        dump(name, qc, full=(name != 'synth_c3'), npts_vec=npv, box=2.0)


if __name__ == '__main__':
    main()
