"""cclib-shaped inputs (plain namespaces with the ccData attributes orbkit's bridge reads: read/cclib_parser.py:56-219)
shared by tests/golden/make_golden_cclib.py (runs the REFERENCE's convert_cclib on them) and tests/test_host.py (runs
orbkit_b200.read_cclib.convert_cclib on the same objects).  cclib itself is not needed for either."""
import types

import numpy


def _gbasis(rng, shells_per_atom):
    out = []
    for shells in shells_per_atom:
        atom = []
        for typ, pnum in shells:
            atom.append((typ, [(float(10 ** rng.uniform(-1, 2)), float(rng.uniform(0.1, 1.0))) for _ in range(pnum)]))
        out.append(atom)
    return out


CART = {'S': ['S'], 'P': ['PX', 'PY', 'PZ'], 'D': ['XX', 'YY', 'ZZ', 'XY', 'XZ', 'YZ'],
        'F': ['XXX', 'YYY', 'ZZZ', 'XYY', 'XXY', 'XXZ', 'XZZ', 'YZZ', 'YYZ', 'XYZ']}
SPH = {'S': ['S'], 'P': ['PX', 'PY', 'PZ'], 'D': ['D 0', 'D+1', 'D-1', 'D+2', 'D-2'],
       'F': ['F 0', 'F+1', 'F-1', 'F+2', 'F-2', 'F+3', 'F-3']}


def _aonames(shells_per_atom, table, symbols):
    names = []
    for ia, shells in enumerate(shells_per_atom):
        n = {}
        for typ, _ in shells:
            n[typ] = n.get(typ, 0) + 1
            for lab in table[typ]:
                names.append('%s%d_%d%s' % (symbols[ia], ia + 1, n[typ] + 'SPDF'.index(typ), lab))
    return names


def case(name):
    """name: rhf_cart (aonames, Cartesian d), uhf_sph (aonames with +/- labels, unrestricted), natorb (natural orbitals),
    noao_sph / noao_cart (no aonames: default order, spherical detected from the coefficient count), triplet (restricted
    open shell: two singly occupied orbitals)"""
    rng = numpy.random.default_rng(sum(map(ord, name)))
    symbols = ['O', 'H', 'H']
    atomnos = numpy.array([8, 1, 1])
    shells = [[('S', 3), ('S', 1), ('P', 2), ('D', 1)], [('S', 2), ('P', 1)], [('S', 2), ('P', 1)]]
    if name in ('uhf_sph', 'noao_sph'):
        shells[0].append(('F', 1))
    sph = name in ('uhf_sph', 'noao_sph')
    table = SPH if sph else CART
    nao = sum(len(table[t]) for sh in shells for t, _ in sh)
    cc = types.SimpleNamespace()
    cc.natom = 3
    cc.atomnos = atomnos
    cc.atomcoords = numpy.array([[[0.0, 0.0, 0.1173], [0.0, 0.7572, -0.4692], [0.0, -0.7572, -0.4692]]])
    cc.gbasis = _gbasis(rng, shells)
    cc.coreelectrons = numpy.zeros(3, dtype=int)
    cc.charge = 0
    cc.mult = 3 if name == 'triplet' else (2 if name == 'uhf_sph' else 1)
    if name == 'uhf_sph':
        cc.charge = 1
    nmo = nao - 2
    nspin = 2 if name == 'uhf_sph' else 1
    cc.nmo = nmo
    cc.mocoeffs = [rng.normal(size=(nmo, nao)) for _ in range(nspin)]
    cc.moenergies = [numpy.sort(rng.uniform(-500, 30, size=nmo)) for _ in range(nspin)]
    irreps = ['A1', 'B2', 'A1', 'B1', 'A2']
    cc.mosyms = [[irreps[int(v)] for v in rng.integers(0, 5, size=nmo)] for _ in range(nspin)]
    cc.homos = numpy.array([4, 3]) if nspin == 2 else numpy.array([4])
    if not name.startswith('noao'):
        cc.aonames = _aonames(shells, table, symbols)
    if name == 'natorb':
        cc.nocoeffs = rng.normal(size=(nmo, nao))
        cc.nooccnos = numpy.concatenate((numpy.sort(rng.uniform(0, 2, size=8))[::-1], numpy.zeros(nmo - 8)))
    return cc


CASES = [('rhf_cart', dict(all_mo=True)), ('rhf_cart', dict(all_mo=False)), ('uhf_sph', dict(all_mo=True)),
         ('uhf_sph', dict(all_mo=False)), ('natorb', dict(all_mo=False)), ('noao_sph', dict(all_mo=True)),
         ('noao_cart', dict(all_mo=True)), ('triplet', dict(all_mo=True))]


def key(name, kw):
    return name + ('_all' if kw.get('all_mo') else '_occ') + ('_' + kw['spin'] if kw.get('spin') else '')
