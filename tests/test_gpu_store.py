"""GPU: `save_hdf5=` streams the results slab by slab into the file (orbkit_b200/store.py; reference behaviour
core.py:478-501, 584-603, 919-938): contents equal the in-memory results, the file is a plain .npz with the
reference's dataset names, and the host never holds more than the two staging slabs (peak RSS of a fresh process)."""
import os
import subprocess
import sys

import numpy
import pytest

from conftest import REPO, assert_close, golden_qc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ok():
    import orbkit_b200
    orbkit_b200.options.quiet = True
    return orbkit_b200


def test_save_hdf5_streams_to_npz_equal_to_memory(ok, tmp_path, monkeypatch):
    from orbkit_b200 import store
    qc, a = golden_qc('synth_small_sph')
    monkeypatch.setattr(store, 'SLAB_BYTES', 1 << 16)            # many slabs
    ax = [numpy.linspace(-5, 5, 23), numpy.linspace(-4, 4, 17), numpy.linspace(-3, 3, 31)]
    ok.grid.set_grid(*ax, is_vector=False)
    ref_rho, ref_d = ok.rho_compute(qc, drv=['z', 'x', 'z2'])
    ref_mo = ok.rho_compute(qc, calc_mo=True, drv=[None, 'y'])
    ref_ao = ok.rho_compute(qc, calc_ao=True)
    fn = str(tmp_path / 'rho.npz')
    rho, d = ok.rho_compute(qc, drv=['z', 'x', 'z2'], save_hdf5=fn)
    assert isinstance(rho, numpy.memmap) and rho.shape == ref_rho.shape and d.shape == ref_d.shape
    assert numpy.array_equal(rho, ref_rho) and numpy.array_equal(d, ref_d)
    with numpy.load(fn) as f:
        assert set(f.files) == {'rho', 'delta_rho', 'grid/x', 'grid/y', 'grid/z', 'grid/is_vector', 'grid/is_regular'}
        assert numpy.array_equal(f['delta_rho'], ref_d) and numpy.array_equal(f['grid/z'], ax[2])
    mo = ok.rho_compute(qc, calc_mo=True, drv=[None, 'y'], save_hdf5=str(tmp_path / 'mo'))   # name without suffix
    assert os.path.exists(str(tmp_path / 'mo.npz')) and numpy.array_equal(mo, ref_mo)
    ao = ok.rho_compute(qc, calc_ao=True, save_hdf5=str(tmp_path / 'ao.npz'))
    assert numpy.array_equal(ao, ref_ao)
    with numpy.load(str(tmp_path / 'ao.npz')) as f:
        assert numpy.array_equal(f['ao_list'], ref_ao)
    r3, d3, lap = ok.rho_compute(qc, laplacian=True, save_hdf5=str(tmp_path / 'lap.npz'))
    rr, dd, ll = ok.rho_compute(qc, laplacian=True)
    assert numpy.array_equal(r3, rr) and numpy.array_equal(d3, dd) and numpy.array_equal(lap, ll)
    # vector grid, density only
    ok.grid.set_grid(a['vx'], a['vy'], a['vz'], is_vector=True)
    assert numpy.array_equal(ok.rho_compute(qc, save_hdf5=str(tmp_path / 'v.npz')), ok.rho_compute(qc))
    # calc_mo_matrix writes the reference's 'mo_matrix' dataset (core.py:919-938)
    ok.grid.set_grid(ax[0][:4], ax[1][:3], ax[2][:5], is_vector=False)
    mm = ok.core.calc_mo_matrix(qc, drv=['x'], save_hdf5=str(tmp_path / 'mm'))
    with numpy.load(str(tmp_path / 'mm.npz')) as f:
        assert numpy.array_equal(f['mo_matrix'], mm) and numpy.array_equal(f['grid/x'], ax[0][:4])


_RSS_WORKER = r'''
import os, resource, sys, numpy
sys.path.insert(0, %(repo)r)
sys.path.insert(0, os.path.join(%(repo)r, 'tests'))
import orbkit_b200 as ok
from conftest import golden_qc
ok.options.quiet = True
qc, a = golden_qc('synth_small_sph')
n = 160
ax = numpy.linspace(-6, 6, n)
ok.grid.set_grid(ax[:8], ax[:8], ax[:8], is_vector=False)
ok.rho_compute(qc, calc_mo=True, save_hdf5=%(out)r + '_warm.npz')      # CUDA context, pinned slabs, writer thread
rss0 = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss * 1024
ok.grid.set_grid(ax, ax, ax, is_vector=False)
mo = ok.rho_compute(qc, calc_mo=True, save_hdf5=%(out)r + '.npz')
rss1 = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss * 1024
nbytes = mo.size * 8
idx = numpy.random.default_rng(0).choice(n ** 3, 2000, replace=False)
sample = numpy.asarray(mo.reshape(mo.shape[0], -1)[:, idx])
numpy.save(%(out)r + '_sample.npy', sample)
numpy.save(%(out)r + '_idx.npy', idx)
print('RESULT', nbytes, rss0, rss1)
'''


def test_save_hdf5_peak_host_memory_is_two_slabs(ok, tmp_path):
    """all MOs of a 160^3 grid (9 MOs: 295 MB) streamed to disk in a fresh process: the peak resident set grows by
    far less than the result (two 64 MB slabs + page cache noise), and a sample of the file equals the in-memory path"""
    out = str(tmp_path / 'big')
    script = tmp_path / 'rss_worker.py'
    script.write_text(_RSS_WORKER % {'repo': REPO, 'out': out})
    p = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    nbytes, rss0, rss1 = [int(v) for v in p.stdout.split('RESULT')[1].split()]
    assert nbytes > 250e6
    assert rss1 - rss0 < 0.6 * nbytes, 'peak RSS grew by %.0f MB for a %.0f MB result' % ((rss1 - rss0) / 1e6, nbytes / 1e6)
    qc, a = golden_qc('synth_small_sph')
    idx = numpy.load(out + '_idx.npy')
    ax = numpy.linspace(-6, 6, 160)
    i, rem = numpy.divmod(idx, 160 * 160)
    j, k = numpy.divmod(rem, 160)
    ok.grid.set_grid(ax[i], ax[j], ax[k], is_vector=True)
    assert_close(numpy.load(out + '_sample.npy'), ok.rho_compute(qc, calc_mo=True), 'streamed sample')
