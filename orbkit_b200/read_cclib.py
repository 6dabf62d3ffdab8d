"""cclib bridge (orbkit/read/cclib_parser.py:10-219): `convert_cclib` turns the ccData object of a cclib parser into a
QCinfo, `read_with_cclib` runs the parser first.  cclib is imported only by `read_with_cclib` (it is an optional
dependency of the reference as well); `convert_cclib` reads attributes only and accepts any object that has them
(atomcoords, atomnos, natom, gbasis, mocoeffs, moenergies, mosyms, homos, coreelectrons, charge, mult; optional aonames,
nmo, nocoeffs + nooccnos).

Pinned to the reference's convert_cclib on the cclib-shaped inputs of tests/cclib_cases.py (tests/golden/read_cclib.npz).
One deliberate difference: for unrestricted data the reference's `spin='alpha'|'beta'` raises AttributeError before it
selects anything (`qc.mo_spec.spinpola`, cclib_parser.py:157, a typo); here the selection its following lines implement
is carried out."""
from importlib import import_module

import numpy

from .display import display
from .orbitals import AOClass, MOClass
from .qcinfo import QCinfo
from .read import AA_TO_A0, _E, _HBAR, _ME, _E0, get_atom_symbol
from .read_wf import _format_geo
from .tools import l_deg, lquant

# eV -> Hartree exactly as orbkit/units.py:20-33 derives it
_A0 = 4 * numpy.pi * _E0 * _HBAR ** 2 / (_ME * _E ** 2)
EV_TO_HA = _E / (_HBAR ** 2 / (_ME * _A0 ** 2))


def read_with_cclib(filename, cclib_parser=None, all_mo=False, spin=None, **kwargs):
    """parse `filename` with cclib's `cclib_parser` ('Gaussian', 'Gamess', 'Orca') and convert the result"""
    if not isinstance(filename, str):
        raise AssertionError('cclib reads from file names')
    if not isinstance(cclib_parser, str):
        raise IOError('cclib requires the specification of parser, e.g., cclib_parser="Gaussian".')
    if cclib_parser == 'Molpro':
        display('\nThe Molpro basis set is not properly read by the cclib parser.')
        display('Please create a molden file with Molpro, i.e., \n\tput,molden,output.molden,NEW;\n')
    parsedic = {'Gaussian': 'gaussianparser', 'Gamess': 'gamessparser', 'Orca': 'orcaparser'}
    module = import_module('cclib.parser.{}'.format(parsedic[cclib_parser]))
    if cclib_parser != 'Gaussian':
        cclib_parser = cclib_parser.upper()
    cc = getattr(module, cclib_parser)(filename).parse()
    return convert_cclib(cc, all_mo=all_mo, spin=spin)


def _m_of_label(label):
    """magnetic quantum number from a cclib AO name such as 'C1_3D+2', 'O1_2PX', 'H2_1S' (cclib_parser.py:122-130)"""
    m = label.lower().split('_')[-1]
    m = m.replace('+', ' +').replace('-', ' -').replace('s', 's 0').split(' ')
    p = 'yzx'.find(m[0][-1])
    return p - 1 if p != -1 else int(m[-1])


def _restricted_occupations(n_el, unpaired, nmo):
    """occupation pattern of a restricted (open-shell) determinant as the reference assigns it orbital by orbital
    (cclib_parser.py:177-186): doubly occupied while more electrons are left than unpaired ones, then singly occupied"""
    occ = numpy.zeros(nmo)
    for n in range(nmo):
        if n_el > unpaired:
            occ[n], n_el = 2.0, n_el - 2.0
        elif 0.0 < n_el <= unpaired:
            occ[n], n_el, unpaired = 1.0, n_el - 1.0, unpaired - 1.0
    return occ


def convert_cclib(ccData, all_mo=False, spin=None):
    qc = QCinfo()
    qc.geo_spec = numpy.asarray(ccData.atomcoords[0]) * AA_TO_A0
    qc.geo_info = [[get_atom_symbol(ccData.atomnos[i]), str(i + 1), str(ccData.atomnos[i])] for i in range(ccData.natom)]
    _format_geo(qc)
    aos = []
    for i in range(ccData.natom):
        for typ, prims in ccData.gbasis[i]:
            aos.append({'atom': i, 'type': str(typ).lower(), 'pnum': len(prims),
                        'coeffs': numpy.array([[p[0], p[1]] for p in prims], dtype=float).reshape((len(prims), 2))})
    spherical = False
    if hasattr(ccData, 'aonames'):
        cartesian = not any('+' in n or '-' in n for n in ccData.aonames)
        spherical = not cartesian
        count = 0
        for ao in aos:
            labels = ccData.aonames[count:count + l_deg(lquant[ao['type']], cartesian_basis=cartesian)]
            count += len(labels)
            if cartesian:
                ao['lxlylz'] = [(n.lower().count('x'), n.lower().count('y'), n.lower().count('z')) for n in labels]
            else:
                ao['lm'] = [(lquant[ao['type']], _m_of_label(n)) for n in labels]
    natural = hasattr(ccData, 'nocoeffs')
    if natural and not hasattr(ccData, 'nooccnos'):
        raise IOError('There are natural orbital coefficients (`nocoeffs`) in the cclib ccData, but no natural '
                      'occupation numbers (`nooccnos`)!')
    nspin = len(ccData.mosyms)                       # 1: restricted (one set of orbitals), 2: alpha and beta sets
    if spin is not None:
        if spin not in ('alpha', 'beta'):
            raise IOError('`spin=%s` is not a valid option' % spin)
        if nspin == 1:
            raise IOError('The keyword `spin` is only supported for unrestricted calculations.')
        display('Converting only molecular orbitals of spin %s.' % spin)
    nmo = ccData.nmo if hasattr(ccData, 'nmo') else len(ccData.mocoeffs[0])
    tags = [('', None)] if nspin == 1 else [('_a', 'alpha'), ('_b', 'beta')]
    aufbau = _restricted_occupations(numpy.sum(ccData.atomnos) - numpy.sum(ccData.coreelectrons) - ccData.charge,
                                     ccData.mult - 1, nmo)
    seen, mos = {}, []
    for n in range(nmo):                              # orbital n of every spin set, alpha before beta
        for s, (suffix, label) in enumerate(tags):
            irrep = '%s%s' % (ccData.mosyms[s][n], suffix)
            seen[irrep] = seen.get(irrep, 0) + 1      # counted for both spins even when one of them is dropped
            if label is not None and spin is not None and spin != label:
                continue
            if natural:
                occ = ccData.nooccnos[n]
            elif nspin == 2:
                occ = 1.0 if n <= ccData.homos[s] else 0.0
            else:
                occ = aufbau[n]
            mo = {'coeffs': numpy.array((ccData.nocoeffs if natural else ccData.mocoeffs[s])[n], dtype=float),
                  'energy': 0.0 if natural else ccData.moenergies[s][n] * EV_TO_HA,
                  'occ_num': occ, 'sym': '%d.%s' % (seen[irrep], irrep)}
            if label is not None:
                mo['spin'] = label
            mos.append(mo)
    qc.ao_spec = AOClass(aos)
    qc.ao_spec.spherical = spherical
    qc.mo_spec = MOClass(mos)
    if not hasattr(ccData, 'aonames'):
        display('The attribute `aonames` is not present in the parsed data.')
        display('Using the default order of basis functions.')
        c_cart = sum(l_deg(l=ao['type'], cartesian_basis=True) for ao in qc.ao_spec)
        c_sph = sum(l_deg(l=ao['type'], cartesian_basis=False) for ao in qc.ao_spec)
        c = qc.mo_spec.get_coeffs().shape[-1]
        if c != c_cart and c == c_sph:
            qc.ao_spec.set_lm_dict(p=[0, 1])
        elif c != c_cart:
            display('Warning: The basis set type does not match with pure spherical or pure Cartesian basis!')
            display('Please specify qc.ao_spec["lxlylz"] and/or qc.ao_spec["lm"] by your self.')
    if not all_mo:
        for i in range(len(qc.mo_spec))[::-1]:
            if qc.mo_spec[i]['occ_num'] < 0.0000001:
                del qc.mo_spec[i]
    qc.mo_spec.update()
    qc.ao_spec.update()
    return qc
