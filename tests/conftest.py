import os
import sys

import numpy
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, 'tests', 'golden')
for p in (REPO, os.path.join(REPO, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)

FIXTURES = ['h2o_molpro_cart', 'h2o_gaussian_sph', 'h2o_gaussian_uhf', 'h2o_gaussian_sph_occ',
            'lih_psi4_sph_f', 'water_gamess_wfn', 'h2o_orca_wfx', 'h2o_turbomole_aomix', 'nh3_molpro',
            'formaldehyde_gamess', 'synth_small_sph', 'synth_small_cart_g', 'synth_c3']
DRV10 = [None, 'x', 'y', 'z', 'xx', 'xy', 'xz', 'yy', 'yz', 'zz']


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu')


def load_golden(name):
    return numpy.load(os.path.join(GOLDEN, name + '.npz'))


def golden_qc(name):
    from orbkit_b200 import QCinfo
    arr = load_golden(name)
    return QCinfo.from_arrays(arr), arr


def assert_close(got, ref, what='', rtol=1e-10, afloor=1e-14):
    """the stated tolerance: |d| <= 1e-10*|ref| + 1e-14*max|ref|   (FP64 path)"""
    got, ref = numpy.asarray(got), numpy.asarray(ref)
    assert got.shape == ref.shape, '%s: shape %s != %s' % (what, got.shape, ref.shape)
    if ref.size == 0:
        return
    tol = rtol * numpy.abs(ref) + afloor * numpy.abs(ref).max()
    err = numpy.abs(got - ref)
    bad = err > tol
    assert not bad.any(), '%s: %d/%d outside tolerance, worst |d|=%.3e at ref=%.3e (max|ref|=%.3e)' % (
        what, bad.sum(), bad.size, err[bad].max(), ref[bad][numpy.argmax(err[bad])], numpy.abs(ref).max())


@pytest.fixture(scope='session')
def oracle_mod():
    import subprocess
    d = os.path.join(REPO, 'oracle')
    if not os.path.exists(os.path.join(d, 'libokoracle.so')):
        subprocess.check_call(['make', '-C', d, 'CC=gcc'])
    import oracle
    return oracle


def reader_input(name, directory):
    """one of the program outputs of tests/golden/reader_inputs.npz written to `directory`; returns its path"""
    data = load_golden('reader_inputs')['file.' + name]
    path = os.path.join(str(directory), name)
    with open(path, 'wb') as f:
        f.write(data.tobytes())
    return path
