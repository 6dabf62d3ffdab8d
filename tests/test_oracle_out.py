"""CPU: the cube-text oracle (oracle/oracle_out.py) reproduces, byte for byte, the files the reference's own
cube_creator wrote (tests/golden/cube_text.npz, generator make_golden_cube.py); host logic of orbkit_b200.output."""
import numpy
import pytest

from conftest import load_golden

CASES = ['rho', 'sets3', 'six', 'special']


@pytest.fixture(scope='module')
def oout(oracle_mod):
    import oracle_out
    return oracle_out


def case_args(g, name):
    labels = [int(v) for v in g[name + '.labels']] if name + '.labels' in g.files else None
    return dict(min_=g[name + '.min_'], N_=[int(v) for v in g[name + '.N_']], delta_=g[name + '.delta_'],
                comments=str(g[name + '.comments']), labels=labels)


def test_cube_oracle_reproduces_reference_files(oout):
    g = load_golden('cube_text')
    for name in CASES:
        text = oout.cube_text(g[name + '.data'], g['geo_info'], g['geo_spec'], **case_args(g, name))
        assert text.encode('utf-8') == g[name + '.text'].tobytes(), name
        body = oout.cube_body(g[name + '.data'])
        assert g[name + '.text'].tobytes().endswith(body), name
    # the corner cases are really in the file: ties to even, carry into the exponent, three-digit exponents, inf / nan
    special = g['special.text'].tobytes().decode()
    for token in ['  1.00000E+05', '  1.00002E+05', ' -1.00000E+05', '  1.00000E+06', '  1.00000E+01', ' 1.00000E+100',
                  '-1.00000E+100', ' 1.00000E-100', ' 4.94066E-324', '-4.94066E-324', '          INF', '         -INF',
                  '          NAN', ' -0.00000E+00', ' 1.79769E+308', '-1.79769E+308']:
        assert token in special, token


def test_cube_header_and_sizes_host_logic():
    """header lines and the closed-form size of the data block (no device needed)"""
    import orbkit_b200 as ok
    from orbkit_b200 import output, _lib
    g = load_golden('cube_text')
    lib = _lib.load()
    for name in CASES:
        a = case_args(g, name)
        ok.grid.min_, ok.grid.N_, ok.grid.delta_ = list(a['min_']), list(a['N_']), list(a['delta_'])
        data = g[name + '.data']
        n_sets = 1 if data.ndim == 3 else data.shape[0]
        head = output.cube_header(n_sets, g['geo_info'], g['geo_spec'], comments=a['comments'], labels=a['labels'])
        text = g[name + '.text'].tobytes()
        assert text.startswith(head.encode()), name
        nx, ny, nz = data.shape[-3:]
        assert lib.okb_cube_body_bytes(n_sets, nx, ny, nz) == len(text) - len(head), name
    assert lib.okb_cube_body_bytes(0, 1, 1, 1) == -1
    assert lib.okb_cube_body_bytes(1, 0, 5, 5) == 0
