"""Readers -> QCinfo for the grid path (orbkit/read/high_level.py:33-79).

Only the formats BASELINE configs[0] names are built: Gaussian formatted checkpoint files
(read_gaussian_fchk, orbkit/read/gaussian_fchk.py:11-324).  The other nine readers of the reference stay with the
reference; `main_read` raises NotImplementedError for them.

Mechanism: an fchk file is a sequence of named sections (`<name, 40 columns> <type I/R/C> [N=] <value or count>` followed,
for arrays, by the values).  The file is cut into sections once, the arrays are converted by NumPy, and the QCinfo is
assembled from the named arrays -- instead of the reference's line-by-line state machine.  The resulting QCinfo is
identical (tests/test_host.py compares every flat array with the reference reader's output, tests/golden/*.npz).
"""
import re

import numpy

from .display import display
from .orbitals import MOClass
from .qcinfo import QCinfo
from .tools import orbit, lquant

SYMBOLS = ('H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr Rb Sr Y Zr '
           'Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu Hf Ta W Re Os Ir '
           'Pt Au Hg Tl Pb Bi Po At Rn Fr Ra Ac Th Pa U Np Pu Am Cm Bk Cf Es Fm Md No Lr Rf Db Sg Bh Hs Mt Ds Rg Cn Nh '
           'Fl Mc Lv Ts Og').split()

_HEAD = re.compile(r'^(?P<name>.{40}) {3}(?P<type>[IRCLH]) {3}(?P<rest>.*)$')


def get_atom_symbol(atom):
    """atomic number (int or numeric string) -> symbol; anything else is title-cased (tools.py:67-86)"""
    try:
        return SYMBOLS[int(atom) - 1]
    except ValueError:
        return str(atom).title()


def fchk_sections(lines):
    """{name: scalar | 1-d array of int / float | str} of a formatted checkpoint file"""
    sec, i, n = {}, 0, len(lines)
    while i < n:
        m = _HEAD.match(lines[i].rstrip('\n'))
        i += 1
        if not m:
            continue
        name, typ, rest = m.group('name').strip(), m.group('type'), m.group('rest').strip()
        if rest.startswith('N='):
            count = int(rest[2:])
            per_line = {'I': 6, 'R': 5, 'C': 5, 'H': 9, 'L': 72}[typ]
            nl = (count + per_line - 1) // per_line
            body = ' '.join(l.rstrip('\n') for l in lines[i:i + nl])
            i += nl
            if typ == 'I':
                sec[name] = numpy.array(body.split(), dtype=numpy.int64)
            elif typ == 'R':
                sec[name] = numpy.array(body.split(), dtype=numpy.float64)
            else:
                sec[name] = body
        else:
            sec[name] = int(rest) if typ == 'I' else float(rest) if typ == 'R' else rest
    return sec


def read_gaussian_fchk(fname, all_mo=False, spin=None, **kwargs):
    """QCinfo of a Gaussian FChk file (gaussian_fchk.py:11-324): geometry, shells (SP shells split into s + p),
    alpha (and beta) MOs with occupations from the electron counts; `all_mo=False` keeps the occupied MOs only."""
    if isinstance(fname, str):
        with open(fname, 'r', encoding='iso-8859-1') as f:
            lines = f.readlines()
    else:
        raw = fname.read()
        lines = (raw.decode('iso-8859-1') if isinstance(raw, bytes) else raw).splitlines(True)
    s = fchk_sections(lines)
    has_beta = any('beta mo coefficients' in k.lower() for k in s)
    is_6d = int(s.get('Pure/Cartesian d shells', 0)) == 1
    is_10f = int(s.get('Pure/Cartesian f shells', 0)) == 1
    if is_6d != is_10f:
        raise IOError('Please apply a Spherical Harmonics (5D, 7F) or a Cartesian Gaussian Basis Set (6D, 10F)!')
    cartesian = is_6d and is_10f
    if spin is not None:
        if spin not in ('alpha', 'beta'):
            raise IOError('`spin=%s` is not a valid option' % spin)
        if not has_beta:
            raise IOError('The keyword `spin` is only supported for unrestricted calculations.')
        display('Reading only molecular orbitals of spin %s.' % spin)
    restricted = not has_beta
    qc = QCinfo()
    qc.etot = float(s['Total Energy']) if 'Total Energy' in s else 0.0
    # geometry: [symbol, running index from 1, nuclear charge]
    numbers, charges = s['Atomic numbers'], s['Nuclear charges']
    qc.geo_info = numpy.array([[get_atom_symbol(int(z)), str(i + 1), str(float(c))]
                               for i, (z, c) in enumerate(zip(numbers, charges))])
    qc.geo_spec = numpy.array(s['Current cartesian coordinates'], dtype=float).reshape((-1, 3))
    # shells
    types, pnum, atoms = s['Shell types'], s['Number of primitives per shell'], s['Shell to atom map']
    expo, coef = s['Primitive exponents'], s['Contraction coefficients']
    sp = s.get('P(S=P) Contraction coefficients', None)
    if not cartesian:
        qc.ao_spec.spherical = True
    off = 0
    for t, n, a in zip(types, pnum, atoms):
        n = int(n)
        letter = orbit[abs(int(t))]
        l = lquant[letter]
        rec = {'type': letter, 'pnum': n, 'atom': int(a) - 1,
               'coeffs': numpy.stack([expo[off:off + n], coef[off:off + n]], axis=1)}
        if not cartesian:
            rec['lm'] = []
            for m in (range(0, l + 1) if l != 1 else [1, 0]):
                rec['lm'].append((l, m))
                if m != 0:
                    rec['lm'].append((l, -m))
        if sp is not None and letter == 'p' and numpy.abs(sp[off:off + n]).sum() > 0:
            # an SP shell: the s part keeps the contraction coefficients, the p part takes the P(S=P) ones.
            # (The reference copies the p record, lm list included, and renames it 's': gaussian_fchk.py:271-282.)
            srec = {k: (v.copy() if hasattr(v, 'copy') else v) for k, v in rec.items()}
            srec['type'] = 's'
            qc.ao_spec.append(srec)
            rec['coeffs'][:, 1] = sp[off:off + n]
        qc.ao_spec.append(rec)
        off += n
    # molecular orbitals: occupations from the electron counts (gaussian_fchk.py:176-196)
    n_el = (int(s['Number of alpha electrons']), int(s['Number of beta electrons']))
    n_bas = int(s['Number of basis functions'])
    mos = []
    for which, label in ((0, 'Alpha'), (1, 'Beta')):
        if label + ' Orbital Energies' not in s:
            continue
        eig = s[label + ' Orbital Energies']
        cf = s[label + ' MO coefficients'].reshape((-1, n_bas))
        if restricted and n_el[0] == n_el[1]:
            n_occ, occ = n_el[0], 2
        else:
            n_occ, occ = n_el[which], 1
        for i in range(len(eig)):
            mos.append({'coeffs': numpy.array(cf[i]) if i < len(cf) else numpy.zeros(n_bas), 'energy': float(eig[i]),
                        'occ_num': float(occ if i < n_occ else 0), 'sym': '%i.1' % (i + 1), 'spin': label.lower()})
    if not all_mo:
        mos = [mo for mo in mos if mo['occ_num'] >= 0.0000001]
    if spin is not None:
        mos = [mo for mo in mos if mo['spin'] == spin]
    for mo in mos:
        if restricted:
            del mo['spin']
        else:
            mo['sym'] += '_%s' % mo['spin'][0]
    if sum(abs(mo['energy']) for mo in mos) < 0.0000001:
        display('Attention!\n\tThis FChk file contains natural orbitals. (There are no energy eigenvalues.)\n\t'
                'In this case, Gaussian does not print the respective natural occupation numbers!')
    qc.mo_spec = MOClass(mos)
    qc.mo_spec.update()
    qc.ao_spec.update()
    return qc


readers = {'gaussian.fchk': read_gaussian_fchk, 'fchk': read_gaussian_fchk}
_OTHER = ('molden', 'aomix', 'gamess', 'gaussian.log', 'gaussian_log', 'wfn', 'wfx', 'cclib', 'native')


def find_itype(fname):
    """file type from the name or the content (read/tools.py:find_itype, reduced to what is told apart here)"""
    name = fname if isinstance(fname, str) else getattr(fname, 'name', '')
    low = name.lower()
    if low.endswith(('.fchk', '.fch')):
        return 'fchk'
    for ext, t in (('.molden', 'molden'), ('.mold', 'molden'), ('.wfn', 'wfn'), ('.wfx', 'wfx'), ('.log', 'gaussian.log'),
                   ('.in', 'aomix'), ('.npz', 'native'), ('.hdf5', 'native'), ('.h5', 'native')):
        if low.endswith(ext):
            return t
    if isinstance(fname, str):
        with open(fname, 'r', encoding='iso-8859-1') as f:
            head = f.read(4096)
        if 'Number of atoms' in head and re.search(r'^.{40} {3}[IR] ', head, re.M):
            return 'fchk'
    raise NotImplementedError('cannot determine the type of %r; pass itype=' % (name,))


def main_read(fname, all_mo=False, spin=None, itype='auto', check_norm=False, **kwargs):
    """High-level reading interface (read/high_level.py:33-79) for the formats built here."""
    if itype == 'auto':
        itype = find_itype(fname)
    if itype not in readers:
        if itype in _OTHER:
            raise NotImplementedError('orbkit_b200 reads Gaussian .fchk files; use the reference\'s reader for %r and pass '
                                      'its QCinfo (or QCinfo(qc.todict())) to orbkit_b200' % itype)
        raise KeyError(itype)
    if check_norm:
        raise NotImplementedError('check_norm needs the analytical overlap integrals (out of scope)')
    display('Loading data from {0} type file {1}\n'.format(itype, fname if isinstance(fname, str)
                                                           else getattr(fname, 'name', '<stream>')))
    return readers[itype](fname, all_mo=all_mo, spin=spin, **kwargs)
