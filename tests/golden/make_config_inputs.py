"""Write tests/golden/benzene_geometry.npz: the atoms of BASELINE configs[1] (SURVEY.md 8d "C2").

Run in the build container only (needs /root/reference):
    python tests/golden/make_config_inputs.py

The geometry is READ from the reference's example input
    /root/reference/examples/orbkit_applications/Benzene_via_terminal/benzene.molden   ([Atoms] Angs block)
with the product's Molden reader (identical QCinfo arrays to the reference's reader, tests/test_host.py), i.e. in
bohr as `QCinfo.geo_spec` holds it.  Only the 12 atom records travel; the def2-TZVP-shaped basis is built by
orbkit_b200.synth.make_benzene_tzvp (the basis set itself is not part of the reference and there is no network).
"""
import os
import sys

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
SRC = '/root/reference/examples/orbkit_applications/Benzene_via_terminal/benzene.molden'

if __name__ == '__main__':
    sys.path.insert(0, REPO)
    import orbkit_b200 as ok
    ok.options.quiet = True
    qc = ok.read.read_molden(SRC, all_mo=False)
    geo_spec = numpy.array(qc.geo_spec, dtype=float)
    geo_info = numpy.array(qc.geo_info)
    assert geo_spec.shape == (12, 3) and list(geo_info[:, 0]) == ['C'] * 6 + ['H'] * 6
    # a regular hexagon in the xy plane: C-C 1.3886 A, C-H 1.0822 A
    cc = numpy.linalg.norm(geo_spec[0] - geo_spec[1]) * 0.52917720859
    assert abs(cc - 1.3886) < 1e-3, cc
    out = os.path.join(HERE, 'benzene_geometry.npz')
    numpy.savez(out, geo_spec=geo_spec, geo_info=geo_info, source=numpy.array(SRC))
    print('wrote', out, geo_spec.round(6).tolist())
