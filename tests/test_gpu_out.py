"""GPU parity of the cube output sink (orbkit_b200.output, C ABI okb_format_cube) -- byte for byte against
   (1) files written by the reference's own cube_creator (tests/golden/cube_text.npz),
   (2) the pinned oracle (oracle/oracle_out.py = Python's correctly rounded '%.5E') on random values over the whole
       double range, values next to rounding ties, ragged shapes and several data sets per file,
   (3) end to end: extras.calc_mo(..., otype='cb') / main_output file names and contents."""
import gzip
import os

import numpy
import pytest

from conftest import load_golden, golden_qc
from test_oracle_out import CASES, case_args

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ok():
    import orbkit_b200
    orbkit_b200.options.quiet = True
    return orbkit_b200


@pytest.fixture(scope='module')
def oout(oracle_mod):
    import oracle_out
    return oracle_out


def set_case_grid(ok, a):
    ok.grid.min_, ok.grid.N_, ok.grid.delta_ = list(a['min_']), list(a['N_']), list(a['delta_'])
    ok.grid.is_initialized, ok.grid.is_vector, ok.grid.is_regular = True, False, True


def test_reference_files_reproduced(ok, tmp_path):
    g = load_golden('cube_text')
    for name in CASES:
        a = case_args(g, name)
        set_case_grid(ok, a)
        fn = ok.output.cube_creator(g[name + '.data'], str(tmp_path / name), g['geo_info'], g['geo_spec'],
                                    comments=a['comments'], labels=a['labels'])
        assert fn.endswith(name + '.cube')
        assert open(fn, 'rb').read() == g[name + '.text'].tobytes(), name
    assert ok.engine.get_engine().last_kernel() == 'cube/format'
    # gzip variant, and data that lives on the device
    import torch
    a = case_args(g, 'special')
    set_case_grid(ok, a)
    dev = torch.from_numpy(g['special.data']).to('cuda')
    fn = ok.output.cube_creator(dev, str(tmp_path / 'dev.cb.gz'), g['geo_info'], g['geo_spec'], comments=a['comments'])
    assert gzip.open(fn, 'rb').read() == g['special.text'].tobytes()


def py_body(data):
    """Python's '%.5E' per value, vectorised over rows (same text as oracle_out.cube_body, faster for 1e6 values)"""
    n_sets, nx, ny, nz = data.shape
    vals = numpy.moveaxis(data, 0, -1).reshape(nx * ny, nz * n_sets)
    rows = []
    for r in vals:
        s = ''.join(('%.5E' % v).rjust(13) + ('\n' if c % 6 == 5 else '') for c, v in enumerate(r))
        rows.append(s + '\n')
    return ''.join(rows).encode()


def test_random_values_whole_double_range(ok, oout):
    rng = numpy.random.default_rng(0)
    # random bit patterns: every exponent incl. subnormals, inf and nan patterns
    bits = rng.integers(0, 2 ** 64, size=3 * 17 * 23 * 31, dtype=numpy.uint64)
    data = bits.view(numpy.float64).reshape(3, 17, 23, 31)
    got = ok.output.cube_body(data).tobytes()
    assert got == py_body(data)
    assert py_body(data[:, :2, :3]) == oout.cube_body(data[:, :2, :3])          # the fast checker == the oracle
    # ordinary magnitudes, one set, several shapes (row lengths around multiples of 6, single values, long rows)
    for shape in [(1, 1, 1, 1), (1, 2, 3, 5), (1, 3, 2, 6), (2, 2, 2, 3), (1, 1, 2, 1025), (4, 3, 1, 7), (1, 64, 64, 70)]:
        data = rng.normal(size=shape) * 10.0 ** rng.integers(-20, 20, size=shape)
        assert ok.output.cube_body(data).tobytes() == py_body(data), shape
    # empty
    assert len(ok.output.cube_body(numpy.zeros((1, 0, 4, 4)))) == 0


def test_values_next_to_ties(ok):
    """(2N+1)/2 * 10^p is an exact tie of '%.5E' whenever it is representable; its neighbours are the hardest non-ties"""
    rng = numpy.random.default_rng(1)
    vals = []
    for p in range(-12, 22):
        for N in rng.integers(100000, 1000000, size=40):
            num = (2 * int(N) + 1)
            v = num * 10.0 ** (p - 5) / 2.0 if p >= 5 else num / (2.0 * 10.0 ** (5 - p))
            vals += [v, numpy.nextafter(v, numpy.inf), numpy.nextafter(v, -numpy.inf), -v]
    # ties that are exactly representable below 1e5: (2q+1) / (2 10^j) with 5^j | 2q+1, i.e. odd o / 2^(j+1)
    n_ties = 0
    for j in range(0, 9):
        lo, hi = 200001 // 5 ** j + 1, 1999999 // 5 ** j
        for o in rng.integers(lo, hi + 1, size=30):
            o = int(o) | 1
            if lo <= o <= hi:
                v = o / 2.0 ** (j + 1)
                assert ('%.6E' % v).endswith('5E' + ('%.6E' % v)[-3:]) and float('%.6E' % v) == v     # an exact tie
                vals += [v, -v, numpy.nextafter(v, numpy.inf), numpy.nextafter(v, -numpy.inf)]
                n_ties += 1
    assert n_ties > 100
    vals = numpy.array(vals, dtype=numpy.float64)
    pad = (-len(vals)) % 7
    data = numpy.concatenate([vals, numpy.ones(pad)]).reshape(1, 1, -1, 7)
    assert ok.output.cube_body(data).tobytes() == py_body(data)


def test_calc_mo_writes_cube_files(ok, tmp_path):
    """extras.calc_mo(..., otype='cb'): one file per MO named <ofid>_<index>.cb, contents == oracle text of the MO values"""
    qc, g = golden_qc('h2o_mo_matrix')
    ok.grid.min_, ok.grid.max_, ok.grid.N_ = [-2.0, -2.1, -1.5], [2.0, 2.0, 2.2], [5, 6, 7]
    ok.grid.delta_ = [0, 0, 0]
    ok.grid.is_initialized = False
    ok.grid.grid_init()
    import oracle_out
    base = str(tmp_path / 'h2o')
    mo = ok.extras.calc_mo(qc, 'all_mo', otype='cb', ofid=base)
    labels = qc.mo_spec.get_labels()
    for i in range(len(qc.mo_spec)):
        fn = '%s_%d.cb' % (base, i)
        assert os.path.exists(fn), fn
        want = oracle_out.cube_text(mo[i], qc.geo_info, qc.geo_spec, ok.grid.min_, ok.grid.N_,
                                    [float(numpy.ravel(d)[0]) for d in ok.grid.delta_], comments=labels[i])
        assert open(fn, 'rb').read() == want.encode()
    # derivatives: <ofid>_<index>_d<drv>.cube ; rho through main_output
    ok.extras.calc_mo(qc, [0, 1], drv=['x', 'z'], otype='cube', ofid=base)
    assert os.path.exists(base + '_0_dx.cube') and os.path.exists(base + '_1_dz.cube')
    rho = ok.rho_compute(qc)
    assert ok.main_output(rho, qc, outputname=base + '_rho', otype='cb') == [base + '_rho.cb']
    want = oracle_out.cube_text(rho, qc.geo_info, qc.geo_spec, ok.grid.min_, ok.grid.N_,
                                [float(numpy.ravel(d)[0]) for d in ok.grid.delta_], comments='')
    assert open(base + '_rho.cb', 'rb').read() == want.encode()
    ok.options.no_output = True
    assert ok.extras.calc_ao(qc, otype='cb', ofid=str(tmp_path / 'none')).shape[0] == qc.ao_spec.get_ao_num()
    ok.options.no_output = False
    assert not os.path.exists(str(tmp_path / 'none_0.cb'))
    with pytest.raises(NotImplementedError):
        ok.main_output(rho, qc, outputname=base, otype='am')
