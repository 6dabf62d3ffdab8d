"""Readers for the primitive-based wave-function formats: AIM `.wfn` (orbkit/read/wfn.py:8-123) and `.wfx`
(orbkit/read/wfx.py:8-162).  Every primitive is its own "contraction" with `pnum = -1` (pre-normalised, c = 1) and an
explicit `lxlylz` row in the wfn order of the Cartesian functions (tools.exp_wfn) -- the input kind the AO kernels take
through their generic shell path.

Mechanism: the keyword lines of a .wfn file / the tagged sections of a .wfx file are located once and their number
columns converted by NumPy, instead of the reference's token-by-token state machine.  The resulting QCinfo is identical
(tests/test_host.py compares every flat array with the reference readers' outputs, tests/golden/*.npz) -- including
the reference's quirk that the `type` letter of a .wfn primitive is looked up with the POSITION of its type number in
the line (wfn.py:80: `orbit[sum(lxlylz[int(i)-1])]` with the loop index `i`).
"""
import re

import numpy

from .display import display
from .orbitals import AOClass, MOClass
from .qcinfo import QCinfo
from .tools import orbit, exp_wfn

_LXLYLZ_WFN = numpy.array([f for l in exp_wfn for f in l], dtype=numpy.int64)


def _text_of(fname):
    if isinstance(fname, str):
        with open(fname, 'rb') as f:
            return f.read().decode('iso-8859-1'), fname
    data = fname.read()
    if isinstance(data, bytes):
        data = data.decode('iso-8859-1')
    return data, getattr(fname, 'name', '<stream>')


def _format_geo(qc):
    """qcinfo.format_geo (qcinfo.py:96-109) after the readers' 'remove numbers from atom names'"""
    from .read import get_atom_symbol
    info = []
    for name, idx, charge in qc.geo_info:
        name = ''.join(k for k in name if not k.isdigit())
        info.append([get_atom_symbol(name), idx, float(charge)])
    qc.geo_info = numpy.array(info)
    qc.geo_spec = numpy.array(qc.geo_spec, dtype=float)


def spin_check(spin, restricted, has_alpha, has_beta):
    """read/tools.py:36-54"""
    if spin is None:
        return
    if restricted:
        raise IOError('The keyword `spin` is only supported for unrestricted calculations.')
    if spin not in ('alpha', 'beta'):
        raise IOError('`spin=%s` is not a valid option' % spin)
    if not has_alpha and not has_beta:
        raise IOError('Molecular orbitals in the input file do not contain `Spin=` keyword')
    if (spin == 'alpha' and not has_alpha) or (spin == 'beta' and not has_beta):
        raise IOError('You requested `%s` orbitals, but None of them are present.' % spin)
    display('Reading only molecular orbitals of spin %s.' % spin)


def select_spin(mos, restricted, spin=None):
    """qcinfo.select_spin (qcinfo.py:164-194) on a list of MO records"""
    if spin is not None:
        mos = [mo for mo in mos if mo['spin'] == spin]
    for mo in mos:
        if restricted:
            del mo['spin']
        else:
            mo['sym'] += '_%s' % mo['spin'][0]
    return mos


# ---- .wfn -----------------------------------------------------------------------------------------------------------
def read_wfn(fname, all_mo=False, spin=None, **kwargs):
    """QCinfo of an AIM .wfn file (wfn.py:8-123); `all_mo` has no effect (the file lists what it lists)"""
    if spin is not None:
        raise IOError('The option `spin` is not supported for the `.wfn` reader.')
    text, name = _text_of(fname)
    qc = QCinfo()
    qc.geo_info, qc.geo_spec = [], []
    aos, mos = [], []
    ao_num = at_left = 0
    c_type = c_exp = c_mo = 0
    section = None
    for line in text.splitlines():
        tok = line.split()
        if 'GAUSSIAN' in line or 'GTO' in line:
            if len(tok) == 8:
                ao_num, at_left = int(tok[4]), int(tok[6])
                section = 'geo'
        elif 'CENTRE ASSIGNMENTS' in line:
            aos += [{'atom': int(t) - 1, 'pnum': -1, 'coeffs': None, 'lxlylz': None} for t in line[20:].split()]
        elif 'TYPE ASSIGNMENTS' in line:
            for i, t in enumerate(line[18:].split()):
                aos[c_type]['lxlylz'] = _LXLYLZ_WFN[int(t) - 1][numpy.newaxis]
                aos[c_type]['type'] = orbit[int(_LXLYLZ_WFN[i - 1].sum())]        # position index: the reference's quirk
                c_type += 1
        elif 'EXPONENTS' in line:
            for t in line.replace('EXPONENTS', '').replace('D', 'E').split():
                aos[c_exp]['coeffs'] = numpy.array([[float(t), 1.0]])
                c_exp += 1
        elif 'MO' in line and 'OCC NO =' in line and 'ORB. ENERGY =' in line:
            rest = line[25:].split()
            mos.append({'coeffs': numpy.zeros(ao_num), 'energy': float(rest[7]), 'occ_num': float(rest[3]),
                        'sym': '%s.1' % tok[1]})
            section, c_mo = 'mo', 0
        elif section == 'geo':
            if not at_left:
                section = None
            else:
                qc.geo_info.append([tok[0], tok[-7][:-1], tok[-1]])
                qc.geo_spec.append([float(v) for v in tok[-6:-3]])
                at_left -= 1
        elif section == 'mo':
            for t in tok:
                if c_mo < ao_num:
                    mos[-1]['coeffs'][c_mo] = float(t.replace('D', 'E'))
                    c_mo += 1
                if c_mo == ao_num:
                    section = None
    _format_geo(qc)
    qc.ao_spec = AOClass(aos)
    qc.mo_spec = MOClass(mos)
    qc.mo_spec.update()
    qc.ao_spec.update()
    return qc


# ---- .wfx -----------------------------------------------------------------------------------------------------------
_RE_SECTION = re.compile(r'<(?P<tag>[^/<>][^<>]*)>(?P<body>.*?)</(?P=tag)>', re.S)


def read_wfx(fname, all_mo=False, spin=None, **kwargs):
    """QCinfo of an AIM .wfx file (wfx.py:8-162): tagged sections `<Name> ... </Name>`"""
    text, name = _text_of(fname)
    sections, mo_blocks = {}, []
    for m in _RE_SECTION.finditer(text):
        tag, body = m.group('tag').strip(), m.group('body')
        if tag == 'Molecular Orbital Primitive Coefficients':
            # <MO Number> n </MO Number> followed by the coefficients of that orbital
            parts = re.split(r'<MO Number>\s*(\d+)\s*</MO Number>', body)
            mo_blocks = [(int(parts[i]), parts[i + 1]) for i in range(1, len(parts) - 1, 2)]
        else:
            sections.setdefault(tag, body)
    if 'GTO' not in sections.get('Keywords', ''):
        raise IOError('No valid .wfx file!\nMissing:\n<Keywords>\n  GTO\n</Keywords>')

    def need(tag, before):
        if before not in sections:
            raise IOError('`<%s>` has to be found before `<%s>`.' % (before, tag))
        return sections[tag]

    at_num = int(sections['Number of Nuclei'])
    names = [l.replace(' ', '') for l in need('Nuclear Names', 'Number of Nuclei').strip().splitlines()][:at_num]
    charges = [l.replace(' ', '') for l in need('Atomic Numbers', 'Number of Nuclei').strip().splitlines()][:at_num]
    coords = numpy.array(need('Nuclear Cartesian Coordinates', 'Number of Nuclei').split(), dtype=float).reshape((-1, 3))[:at_num]
    qc = QCinfo()
    qc.geo_info = [[names[i], i + 1, charges[i]] for i in range(at_num)]
    qc.geo_spec = [list(c) for c in coords]
    ao_num = int(sections['Number of Primitives'])
    centers = numpy.array(sections['Primitive Centers'].split(), dtype=int)
    types = numpy.array(sections['Primitive Types'].split(), dtype=int)
    expo = numpy.array(sections['Primitive Exponents'].replace('D', 'E').split(), dtype=float)
    aos = [{'atom': int(centers[i]) - 1, 'pnum': -1, 'coeffs': numpy.array([[expo[i], 1.0]]),
            'lxlylz': _LXLYLZ_WFN[types[i] - 1][numpy.newaxis], 'type': orbit[int(_LXLYLZ_WFN[types[i] - 1].sum())]}
           for i in range(ao_num)]
    mo_num = int(sections['Number of Occupied Molecular Orbitals'])
    occ = numpy.array(sections['Molecular Orbital Occupation Numbers'].split(), dtype=float)
    ene = numpy.array(sections['Molecular Orbital Energies'].split(), dtype=float)
    spins = [l.replace(' ', '').replace('and', '_').lower()
             for l in sections['Molecular Orbital Spin Types'].strip().splitlines()][:mo_num]
    restricted = all('_' in s for s in spins)
    mos = [{'coeffs': numpy.zeros(ao_num), 'energy': float(ene[i]), 'occ_num': float(occ[i]), 'spin': spins[i],
            'sym': '%s.1' % (i + 1)} for i in range(mo_num)]
    for number, body in mo_blocks:
        mos[number - 1]['coeffs'][:] = numpy.array(body.replace('D', 'E').split(), dtype=float)[:ao_num]
    spin_check(spin, restricted, any(s == 'alpha' for s in spins), any(s == 'beta' for s in spins))
    mos = select_spin(mos, restricted, spin=spin)
    _format_geo(qc)
    qc.ao_spec = AOClass(aos)
    qc.mo_spec = MOClass(mos)
    qc.mo_spec.update()
    qc.ao_spec.update()
    return qc
