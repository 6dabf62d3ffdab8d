"""Golden fixture for the cclib bridge, written by RUNNING THE REFERENCE's orbkit/read/cclib_parser.py:convert_cclib on the
cclib-shaped namespaces of tests/cclib_cases.py (cclib itself is not needed: the function only reads attributes).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_cclib.py
Writes tests/golden/read_cclib.npz: `<case key>.<flat QCinfo arrays>` (make_golden.qc_arrays); `<key>.error` holds the
exception class name where the reference raises."""
import os
import sys

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import make_golden as mg     # noqa: E402
import cclib_cases as cases  # noqa: E402


def main():
    scratch = mg.build_reference()
    sys.path.insert(0, scratch)
    mg.shim()
    from orbkit import options
    from orbkit.read.cclib_parser import convert_cclib
    options.quiet = True
    options.no_log = True
    out = {}
    for name, kw in cases.CASES + [('uhf_sph', dict(all_mo=True, spin='beta')), ('rhf_cart', dict(all_mo=True, spin='alpha'))]:
        k = cases.key(name, kw)
        try:
            qc = convert_cclib(cases.case(name), **kw)
        except Exception as e:          # noqa: BLE001
            out[k + '.error'] = numpy.array(type(e).__name__)
            print(k, 'raises', type(e).__name__, e)
            continue
        for kk, v in mg.qc_arrays(qc).items():
            out[k + '.' + kk] = v
        print(k, len(qc.mo_spec), 'MOs', qc.ao_spec.get_ao_num(), 'AOs', 'spherical' if qc.ao_spec.spherical else 'cartesian')
    numpy.savez_compressed(os.path.join(HERE, 'read_cclib.npz'), **out)


if __name__ == '__main__':
    main()
