"""Seeded synthetic molecules for the benchmark configs (BASELINE.json configs 2-5).

Pure numpy, no device code.  The generator follows SURVEY.md section 8(d): "heavy" atoms carry
the shell pattern  s8 s8 s1 s1 p3 p1 p1 d1 d1 f1  (30 spherical / 35 Cartesian functions,
26 primitives), "light" atoms  s3 s1 s1 p1 p1 d1  (14 / 15 functions, 8 primitives) --
cc-pVTZ-shaped.  24 heavy + 20 light atoms give S=360 contractions, P=784 primitives,
n_cart=1140, n_ao=1000 (config 3).

The result is a plain dict in the reference's list-of-dict data model
(orbkit/orbitals.py:25-57, qcinfo.py:35-64) so that it can be fed both to
orbkit_b200.QCinfo and -- on the oracle side -- to the reference's own classes.
"""
import numpy

_HEAVY = [('s', 8), ('s', 8), ('s', 1), ('s', 1), ('p', 3), ('p', 1), ('p', 1),
          ('d', 1), ('d', 1), ('f', 1)]
_LIGHT = [('s', 3), ('s', 1), ('s', 1), ('p', 1), ('p', 1), ('d', 1)]
_NSPH = {'s': 1, 'p': 3, 'd': 5, 'f': 7, 'g': 9}
_NCART = {'s': 1, 'p': 3, 'd': 6, 'f': 10, 'g': 15}


def make_molecule(n_heavy=24, n_light=20, n_mo=82, seed=0, spherical=True, box=8.0,
                  with_g=False, occ=2.0):
    """Return {'geo_spec','geo_info','ao_spec','mo_spec','spherical'} (list-of-dict model)."""
    rng = numpy.random.default_rng(seed)
    n_at = n_heavy + n_light
    geo_spec = rng.uniform(-box, box, size=(n_at, 3))
    geo_info = [['C' if i < n_heavy else 'H', str(i + 1), '6.0' if i < n_heavy else '1.0']
                for i in range(n_at)]
    ao_spec = []
    for iat in range(n_at):
        pattern = list(_HEAVY if iat < n_heavy else _LIGHT)
        if with_g and iat < n_heavy:
            pattern.append(('g', 2))
        for typ, pnum in pattern:
            if pnum > 1:
                alpha = numpy.sort(10.0 ** rng.uniform(-1.0, 3.0, size=pnum))[::-1]
            else:
                alpha = 10.0 ** rng.uniform(-1.0, 0.5, size=1)
            coef = rng.uniform(0.1, 1.0, size=pnum)
            ao_spec.append({'atom': iat, 'type': typ, 'pnum': pnum,
                            'coeffs': numpy.stack([alpha, coef], axis=1)})
    deg = _NSPH if spherical else _NCART
    n_ao = sum(deg[a['type']] for a in ao_spec)
    coeffs = rng.standard_normal((n_mo, n_ao)) / numpy.sqrt(n_ao)
    mo_spec = [{'coeffs': coeffs[i].copy(), 'energy': -10.0 + 0.1 * i, 'occ_num': float(occ),
                'sym': '%d.1' % (i + 1)} for i in range(n_mo)]
    return {'geo_spec': geo_spec, 'geo_info': geo_info, 'ao_spec': ao_spec,
            'mo_spec': mo_spec, 'spherical': bool(spherical)}


# def2-TZVP-SHAPED shells for benzene (BASELINE configs[1], SURVEY.md 8d "C2"): the contraction pattern of def2-TZVP,
# C {62111/411/11/1} -> [5s3p2d1f], H {311/1} -> [3s1p].  The basis set is not part of the reference and there is no
# network: the exponents / contraction coefficients below were written down from memory of the published table and are
# NOT verified against it -- every run that uses them is labelled "def2-TZVP-shaped".  What the configuration is about
# (222 spherical / 252 Cartesian functions, 11+4 shells per atom, realistic exponent ranges) does not depend on them.
_TZVP_C = [
    ('s', [(13575.349682, 0.00022245814352), (2035.2333680, 0.0017232738252), (463.22562359, 0.0089255715314),
           (131.20019598, 0.035727984502), (42.853015891, 0.11076259931), (15.584185766, 0.24295627626)]),
    ('s', [(6.2067138508, 0.41440263448), (2.5764896527, 0.23744968655)]),
    ('s', [(0.57696339419, 1.0)]), ('s', [(0.22972831358, 1.0)]), ('s', [(0.095164440028, 1.0)]),
    ('p', [(34.697232244, 0.0053333657805), (7.9582622826, 0.035864109092), (2.3780826883, 0.14215873329),
           (0.81433208183, 0.34270471845)]),
    ('p', [(0.28887547253, 1.0)]), ('p', [(0.10056823671, 1.0)]),
    ('d', [(1.097, 1.0)]), ('d', [(0.318, 1.0)]),
    ('f', [(0.761, 1.0)]),
]
_TZVP_H = [
    ('s', [(34.0613410, 0.0060251978), (5.1235746, 0.045021094), (1.1646626, 0.20189726)]),
    ('s', [(0.32723041, 1.0)]), ('s', [(0.10307241, 1.0)]),
    ('p', [(0.8, 1.0)]),
]


def make_benzene_tzvp(geo_spec, geo_info, n_occ=21, seed=2, all_mo=True):
    """BASELINE configs[1]: benzene at the geometry of the reference's example (tests/golden/benzene_geometry.npz),
    def2-TZVP-shaped spherical basis (222 AOs), seeded random MO coefficients for all 222 MOs (or the n_occ occupied
    ones), occupation 2 on the first n_occ."""
    geo_spec = numpy.array(geo_spec, dtype=float)
    ao_spec = []
    for iat, info in enumerate(geo_info):
        for typ, prims in (_TZVP_C if str(info[0]).upper() == 'C' else _TZVP_H):
            ao_spec.append({'atom': iat, 'type': typ, 'pnum': len(prims), 'coeffs': numpy.array(prims, dtype=float)})
    n_ao = sum(_NSPH[a['type']] for a in ao_spec)
    rng = numpy.random.default_rng(seed)
    coeffs = rng.standard_normal((n_ao, n_ao)) / numpy.sqrt(n_ao)
    n_mo = n_ao if all_mo else n_occ
    mo_spec = [{'coeffs': coeffs[i].copy(), 'energy': -11.0 + 0.1 * i, 'occ_num': 2.0 if i < n_occ else 0.0,
                'sym': '%d.1' % (i + 1)} for i in range(n_mo)]
    return {'geo_spec': geo_spec, 'geo_info': [list(map(str, g)) for g in geo_info], 'ao_spec': ao_spec,
            'mo_spec': mo_spec, 'spherical': True}


def counts(spec):
    ao = spec['ao_spec']
    return {'n_cont': len(ao), 'n_prim': sum(len(a['coeffs']) for a in ao),
            'n_cart': sum(_NCART[a['type']] for a in ao),
            'n_ao': sum((_NSPH if spec['spherical'] else _NCART)[a['type']] for a in ao),
            'n_mo': len(spec['mo_spec'])}


def to_qcinfo(spec):
    """Build an orbkit_b200.QCinfo from the list-of-dict spec."""
    from .qcinfo import QCinfo
    from .orbitals import AOClass, MOClass
    qc = QCinfo()
    qc.geo_spec = numpy.array(spec['geo_spec'], dtype=float)
    qc.geo_info = numpy.array(spec['geo_info'])
    qc.ao_spec = AOClass([dict(d) for d in spec['ao_spec']])
    if spec['spherical']:
        qc.ao_spec.set_lm_dict(p=[1, 0])
    qc.mo_spec = MOClass([dict(d) for d in spec['mo_spec']])
    qc.ao_spec.update()
    qc.mo_spec.update()
    return qc
