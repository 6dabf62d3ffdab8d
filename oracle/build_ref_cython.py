"""Cythonize the reference's own cy_core.pyx / cy_grid.pyx / cy_overlap.pyx / detci/cy_ci.pyx into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  The .pyx/.c sources are read where they lie under the
reference tree (argv[1], default /root/reference); generated C and objects go to a
scratch directory; only the resulting extension modules (`cy_core`, `cy_grid`, `cy_overlap`, `cy_ci`) are
written to oracle/_ref/ (git-ignored).  The reference's own setup.py is not used.
"""
import os, sys, shutil, tempfile, glob

def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
    here = os.path.dirname(os.path.abspath(__file__))
    dest = os.path.join(here, '_ref')
    os.makedirs(dest, exist_ok=True)
    import numpy
    from setuptools import setup, Extension
    from Cython.Build import cythonize
    src = os.path.join(ref, 'orbkit')
    scratch = tempfile.mkdtemp(prefix='okref_build_')
    # cythonize writes the generated .c next to the .pyx unless build_dir is given
    exts = [
        Extension('cy_grid', [os.path.join(src, 'cy_grid.pyx')],
                  include_dirs=[numpy.get_include(), src]),
        Extension('cy_core', [os.path.join(src, 'cy_core.pyx'),
                              os.path.join(src, 'c_grid-based.c'),
                              os.path.join(src, 'c_support.c')],
                  include_dirs=[numpy.get_include(), src]),
        # analytic overlap integrals (read/molden.py:378-398 renormalisation, check_mo_norm)
        Extension('cy_overlap', [os.path.join(src, 'cy_overlap.pyx'),
                                 os.path.join(src, 'c_non-grid-based.c'),
                                 os.path.join(src, 'c_support.c')],
                  include_dirs=[numpy.get_include(), src]),
        # detCI grid contractions (prange loops: OpenMP, as in the reference's setup.py)
        Extension('cy_ci', [os.path.join(src, 'detci', 'cy_ci.pyx')],
                  include_dirs=[numpy.get_include(), src],
                  extra_compile_args=['-fopenmp'], extra_link_args=['-fopenmp']),
    ]
    cwd = os.getcwd()
    os.chdir(scratch)
    try:
        setup(name='okref', script_args=['build_ext', '--build-lib', dest,
                                         '--build-temp', os.path.join(scratch, 'tmp')],
              ext_modules=cythonize(exts, language_level=3,
                                    build_dir=os.path.join(scratch, 'cy')))
    finally:
        os.chdir(cwd)
        shutil.rmtree(scratch, ignore_errors=True)
    print('built:', sorted(os.path.basename(p) for p in glob.glob(os.path.join(dest, '*.so'))))

if __name__ == '__main__':
    main()
