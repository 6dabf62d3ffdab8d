"""Golden fixture for the cube output sink, written by RUNNING THE REFERENCE's cube_creator (output/cube.py:5-101).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_cube.py

Writes tests/golden/cube_text.npz: the QCinfo of the H2O fchk fixture (geo_info / geo_spec for the header) and, per
case <c>, the input `<c>.data`, the grid attributes `<c>.min_ / N_ / delta_`, optional `<c>.labels`, the comment line and
`<c>.text` = the bytes of the file the reference wrote:
    rho      the H2O density on a 5 x 6 x 7 grid (nz = 7: one full line of six values + one value per row)
    sets3    three data sets in one file with labels (values of the sets interleaved per point), nz = 4
    six      nz = 6: the row ends right behind a line break (the reference then writes an empty line)
    special  '%.5E' corner cases: exact ties (round-half-even), values next to ties, carries 9.999995 -> 1.00000E+01,
             three-digit exponents, subnormals, signed zeros, inf, nan
"""
import os
import sys
import tempfile

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg     # noqa: E402


def special_values():
    v = [0.0, -0.0, 1.0, -1.0, 100000.5, 100001.5, 100002.5, -100000.5, 999999.5, 9.999995, 9.9999949999999, 9.99999500000001,
         1.000005, 1.0000050000000001, 0.5, 0.25, 1.5e-5, 123456.5, 1234565.0, 12345650.0, 1.234565e10, 1.234575e10,
         1e100, -1e100, 1e-100, -1e-100, 9.999996e99, 9.999995e-101, 1.7976931348623157e308, -1.7976931348623157e308,
         2.2250738585072014e-308, 4.9406564584124654e-324, -4.9406564584124654e-324, 1.2345e-310, 9.99999e-320,
         float('inf'), float('-inf'), float('nan'), 3.141592653589793, -2.718281828459045e-7, 6.02214076e23,
         1e5, 1e6 - 0.5, 1e6 - 0.49999, 1e-5, 1.00000049999e-5, 0.1, 0.3, 2.0 ** 60, 2.0 ** -60, 2.0 ** 1000, 2.0 ** -1000,
         5e-324 * 3, 1e22, 1e23, 8.5, 7.5e-3]
    # exact ties at other magnitudes: (2N+1)/2 * 10^p with few bits
    for N, p in [(123456, 0), (123457, 0), (500000, 1), (100000, 3), (999998, 2), (131072, 5), (131073, 5),
                 (262144, -1), (262145, -1), (390625, -6), (390626, -6), (195312, -7), (195313, -7)]:
        v.append((2 * N + 1) / 2.0 * 10.0 ** (p - 5) if p - 5 >= 0 else (2 * N + 1) / 2.0 / 10.0 ** (5 - p))
    rng = numpy.random.default_rng(7)
    v += list(rng.normal(size=20) * 10.0 ** rng.integers(-30, 30, size=20))
    return numpy.array(v, dtype=numpy.float64)


def main():
    scratch = mg.build_reference()
    sys.path.insert(0, scratch)
    mg.shim()
    from orbkit import grid, options, read, core
    from orbkit.output.cube import cube_creator
    options.quiet = True
    options.no_log = True
    odir = os.path.join(scratch, 'orbkit', 'test', 'outputs_for_testing')
    qc = read.main_read(os.path.join(odir, 'gaussian', 'h2o_rhf_sph.fchk'), all_mo=False)
    out = mg.qc_arrays(qc)
    tmp = tempfile.mkdtemp()

    def case(name, data, min_, N_, delta_, comments='', labels=None):
        grid.min_, grid.N_, grid.delta_ = list(min_), list(N_), list(delta_)
        grid.x, grid.y, grid.z = [numpy.arange(n, dtype=float) for n in N_]     # cube_creator only takes their lengths
        grid.is_initialized = True
        fn = os.path.join(tmp, name + '.cube')
        cube_creator(data, fn, qc.geo_info, qc.geo_spec, comments=comments, labels=labels)
        out[name + '.data'] = numpy.array(data)
        out[name + '.min_'] = numpy.array(min_, dtype=float)
        out[name + '.N_'] = numpy.array(N_, dtype=int)
        out[name + '.delta_'] = numpy.array(delta_, dtype=float)
        out[name + '.comments'] = numpy.array(comments)
        if labels is not None:
            out[name + '.labels'] = numpy.array(labels, dtype=int)
        out[name + '.text'] = numpy.frombuffer(open(fn, 'rb').read(), dtype=numpy.uint8)
        print(name, out[name + '.data'].shape, len(out[name + '.text']), 'bytes')

    grid.min_, grid.max_, grid.N_ = [-2.0, -2.1, -1.5], [2.0, 2.0, 2.2], [5, 6, 7]
    grid.is_initialized = False
    grid.grid_init()
    rho = core.rho_compute(qc)
    case('rho', rho, grid.min_, grid.N_, [float(numpy.ravel(d)[0]) for d in grid.delta_], comments='rho of h2o')
    rng = numpy.random.default_rng(3)
    case('sets3', rng.normal(size=(3, 3, 2, 4)) * 10.0 ** rng.integers(-12, 12, size=(3, 3, 2, 4)),
         [-1.0, 0.0, 1.0], [3, 2, 4], [0.5, 0.25, 0.125], comments='three sets', labels=[1, 2, 30])
    case('six', rng.normal(size=(2, 2, 6)), [0.0, 0.0, 0.0], [2, 2, 6], [1.0, 1.0, 1.0])
    sv = special_values()
    n = len(sv)
    nz = 9
    pad = (-n) % (2 * nz)
    sv = numpy.concatenate([sv, rng.normal(size=pad)]).reshape((2, -1, nz))
    case('special', sv, [0.0, 0.0, 0.0], list(sv.shape), [1.0, 1.0, 1.0], comments='corner cases of %.5E')
    numpy.savez_compressed(os.path.join(HERE, 'cube_text.npz'), **out)


if __name__ == '__main__':
    main()
