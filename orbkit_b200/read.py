"""Readers -> QCinfo for the grid path (orbkit/read/high_level.py:33-79).

This module: Gaussian formatted checkpoint files (read_gaussian_fchk, orbkit/read/gaussian_fchk.py:11-324), Molden files
(read_molden, orbkit/read/molden.py:47-402), `find_itype` and `main_read`.  The other formats live in read_wf.py (.wfn, .wfx),
read_gamess.py, read_aomix.py, read_glog.py (Gaussian .log) and read_cclib.py (cclib bridge); the `native` containers are not built
(`main_read` raises NotImplementedError for them).

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
from .tools import orbit, lquant, exp as cart_exponents

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


# ---- Molden ---------------------------------------------------------------------------------------------------------
_FLOAT = r'([-+]?\d+\.?\d*[ed]?[+-]?\d+|NaN)'
_RE_MOLDEN = re.compile(r'\[\s*molden\s+format\s*\]', re.I)
_RE_ATOMS = re.compile(r'\[atoms\]\s*\(?(angs|au)\)?', re.I)
_RE_ATOM = re.compile(r'\s*([a-z]+)\s+(\d+)\s+(\d+)\s+' + r'\s+'.join((_FLOAT,) * 3), re.I)
_RE_BASIS = re.compile(r'\s*(\d+)\s+(\d+)$', re.I)
_RE_CONTRACTION = re.compile(r'\s*([a-z]+)\s+(\d+)\s+(\d+(\.\d+)?)\s*($)', re.I)
_RE_PRIMITIVE = re.compile(r'\s*' + r'\s+'.join((_FLOAT,) * 2), re.I)
_SPH_FLAGS, _CART_FLAGS = ['5d', '7f', '9g'], ['6d', '10f', '15g']
_RE_FLAGLINE = re.compile(r'\[((' + '|'.join(_SPH_FLAGS + _CART_FLAGS) + r')+)\]', re.I)
_RE_FLAG = re.compile(r'(\d+[dfg])', re.I)
_RE_SYM = re.compile(r'\s*sym\s*=\s*(\S+)', re.I)
_RE_ENERGY = re.compile(r'\s*ene(?:rgy)?\s*=\s*' + _FLOAT, re.I)
_RE_SPIN = re.compile(r'\s*spin\s*=\s*(alpha|beta)', re.I)
_RE_OCC = re.compile(r'\s*occup\s*=\s*' + _FLOAT, re.I)
_RE_COEFF = re.compile(r'\s*(\d+)\s+' + _FLOAT, re.I)
# Angstrom -> Bohr exactly as orbkit/units.py:5-28 derives it (CODATA 2014 constants)
_H, _E, _ME, _E0 = 6.626070040 * 1e-34, 1.6021766208 * 1e-19, 9.10938356 * 1e-31, 8.854187817 * 1e-12
_HBAR = _H / (2 * numpy.pi)
AA_TO_A0 = 1e-10 / (4 * numpy.pi * _E0 * _HBAR ** 2 / (_ME * _E ** 2))


def _dfact(n):
    r = 1.0
    while n > 1:
        r *= n
        n -= 2
    return r


def cartesian_self_overlap(ao_spec):
    """<chi|chi> of every contracted Cartesian function: the diagonal the reference takes from
    analytical_integrals.get_ao_overlap (molden.py:378-381), here in closed form -- all primitives of a contraction share
    their centre, so  S = sum_pq c_p c_q N_p N_q  prod_axis (2l-1)!! / (2(a_p+a_q))^l  (pi/(a_p+a_q))^(3/2)  with the
    primitive norm N = ao_norm of c_support.c:177-188."""
    out = []
    for rec in ao_spec:
        c = numpy.asarray(rec['coeffs'], dtype=float).reshape((-1, 2))
        a, w = c[:, 0], c[:, 1]
        s = a[:, None] + a[None, :]
        lxlylz = rec['lxlylz'] if 'lxlylz' in rec else cart_exponents[lquant[rec['type']]]
        for lx, ly, lz in lxlylz:
            L = lx + ly + lz
            if rec['pnum'] < 0:
                n = numpy.ones_like(a)
            else:
                n = (2.0 / numpy.pi) ** 0.75 * 2.0 ** L * a ** ((2.0 * L + 3.0) / 4.0) / numpy.sqrt(
                    _dfact(2 * lx - 1) * _dfact(2 * ly - 1) * _dfact(2 * lz - 1))
            ang = _dfact(2 * lx - 1) * _dfact(2 * ly - 1) * _dfact(2 * lz - 1) / (2.0 * s) ** L
            out.append(float(((w * n)[:, None] * (w * n)[None, :] * ang * (numpy.pi / s) ** 1.5).sum()))
    return numpy.array(out)


def read_molden(fname, all_mo=False, spin=None, i_md=-1, interactive=False, **kwargs):
    """QCinfo of a Molden file (molden.py:47-402): [Atoms], [GTO], the [5D]/[7F]/[9G] flags, [MO]; `i_md` selects the
    [Molden Format] section of files that hold several (never asked for interactively here)."""
    if isinstance(fname, str):
        with open(fname, 'r') as f:
            text = f.read()
        name = fname
    else:
        text = fname.read()
        name = getattr(fname, 'name', '<stream>')
        if isinstance(text, bytes):
            text = text.decode()
    entries = [m.start() for m in _RE_MOLDEN.finditer(text)]
    if not entries:
        raise IOError('The input file {:s} is no valid molden file!\n\nIt does not contain the keyword: '
                      '[Molden Format]\n'.format(name))
    if len(entries) > 1:
        i_md = list(range(len(entries)))[i_md]
        display('\tFound {:d} [Molden Format] keywords; selecting the element with index {:d}.'.format(len(entries), i_md))
        text = text[entries[i_md]:(entries + [None])[i_md + 1]]
    lines = text.splitlines()
    qc = QCinfo()
    qc.geo_info, qc.geo_spec = [], []
    sph_flags, cart_flags, angular = [], [], []
    by_orca, angstrom = False, False
    at_num, ao_type, row = 0, '', 0
    iline = 0
    for iline, line in enumerate(lines):
        low = line.lower()
        if 'orca' in low:
            by_orca = True
            continue
        if '_ENERGY=' in line:
            try:
                qc.etot = float(line.split()[1])
            except IndexError:
                pass
            continue
        m = _RE_ATOMS.match(line)
        if m:
            angstrom = m.group(1).lower() == 'angs'
            continue
        m = _RE_ATOM.match(line)
        if m:
            qc.geo_info.append(list(m.groups()[:3]))
            qc.geo_spec.append([float(f) for f in m.groups()[3:]])
            continue
        if '[sto]' in low:
            raise IOError('orbkit does not work for STOs!\nEXIT\n')
        m = _RE_BASIS.match(line)
        if m:
            at_num = int(m.group(1)) - 1
            continue
        m = _RE_FLAGLINE.match(low)
        if m:
            for flag in _RE_FLAG.findall(m.group(1)):
                (sph_flags if flag in _SPH_FLAGS else cart_flags).append(flag)
        m = _RE_CONTRACTION.match(line)
        if m:
            row, ao_type, pnum = 0, m.group(1).lower(), int(m.group(2))
            for l in ao_type:                       # "sp" shells become two contractions
                qc.ao_spec.append({'atom': at_num, 'type': l, 'pnum': -pnum if by_orca else pnum,
                                   'coeffs': numpy.zeros((pnum, 2))})
                if l not in angular:
                    angular.append(l)
            continue
        m = _RE_PRIMITIVE.match(line)
        if m:
            vals = numpy.array(low.replace('d', 'e').split(), dtype=numpy.float64)
            for i in range(len(ao_type)):
                qc.ao_spec[-len(ao_type) + i]['coeffs'][row, :] = [vals[0], vals[1 + i]]
            row += 1
            continue
        if '[mo]' in low:
            break
    # Cartesian or spherical functions?
    max_l = max(lquant[l] for l in angular)
    cartesian = True
    if max_l >= 2:
        used = orbit[2:max_l + 1]
        sph = [f for f in sph_flags if f[-1] in used]
        cart = [f for f in cart_flags if f[-1] in used]
        if sph and cart:
            raise IOError('The input file {} contains mixed spherical and Cartesian function ({}).'.format(
                name, ', '.join(sph + cart)))
        cartesian = not bool(sph)
    n_basis = sum(((lquant[ao['type']] + 1) * (lquant[ao['type']] + 2) // 2) if cartesian else (2 * lquant[ao['type']] + 1)
                  for ao in qc.ao_spec)
    # [MO]
    mos = []
    pending = {'sym': None, 'energy': None, 'occ_num': None, 'spin': None}
    new_mo, has_alpha, has_beta, restricted = False, False, False, False
    counters = {}
    for line in lines[iline:]:
        m = _RE_COEFF.match(line)
        if m:
            if new_mo:
                sym = pending['sym']
                digits = re.search(r'\d+', sym) if sym else None
                if digits:
                    a = digits.group()
                    if sym == a:
                        sym = '{:s}.1'.format(a)
                    elif not sym.startswith(a):
                        counters[a] = counters.get(a, 0) + 1
                        sym = '{:d}.{:s}'.format(counters[a], sym)
                sym = sym or '%d.1' % (len(mos) + 1)
                mos.append({'coeffs': numpy.zeros(n_basis), 'sym': sym, 'energy': pending['energy'],
                            'occ_num': pending['occ_num'], 'spin': pending['spin'] or 'alpha'})
                pending = {'sym': None, 'energy': None, 'occ_num': None, 'spin': None}
                new_mo = False
            value = float(m.group(2))
            if numpy.isnan(value):
                display('Warning: coefficient {:d} of MO {:s} is NaN! Using zero instead'.format(int(m.group(1)) - 1,
                                                                                               mos[-1]['sym']))
            else:
                mos[-1]['coeffs'][int(m.group(1)) - 1] = value
            continue
        new_mo = True
        m = _RE_SYM.match(line)
        if m:
            pending['sym'] = m.group(1)
            continue
        m = _RE_ENERGY.match(line)
        if m:
            pending['energy'] = m.group(1)
            continue
        m = _RE_SPIN.match(line)
        if m:
            pending['spin'] = m.group(1).lower()
            has_alpha |= pending['spin'] == 'alpha'
            has_beta |= pending['spin'] == 'beta'
            continue
        m = _RE_OCC.match(line)
        if m:
            pending['occ_num'] = float(m.group(1))
            restricted |= pending['occ_num'] > 1.0001
            continue
    if spin is not None:
        if restricted:
            raise IOError('The keyword `spin` is only supported for unrestricted calculations.')
        if spin not in ('alpha', 'beta'):
            raise IOError('`spin=%s` is not a valid option' % spin)
        if not has_alpha and not has_beta:
            raise IOError('Molecular orbitals in `molden` file do not contain `Spin=` keyword')
        if (spin == 'alpha' and not has_alpha) or (spin == 'beta' and not has_beta):
            raise IOError('You requested `%s` orbitals, but None of them are present.' % spin)
        display('Reading only molecular orbitals of spin %s.' % spin)
    if sph_flags:
        qc.ao_spec.set_lm_dict(p=[1, 0])
    if not all_mo:
        mos = [mo for mo in mos if mo['occ_num'] >= 0.0000001]
    if spin is not None:
        mos = [mo for mo in mos if mo['spin'] == spin]
    for mo in mos:
        if restricted:
            del mo['spin']
        else:
            mo['sym'] += '_%s' % mo['spin'][0]
    syms = [mo['sym'] for mo in mos]
    if syms[1:] == syms[:-1]:                      # ORCA gives every orbital the same name
        tail = syms[0].split('.')[-1]
        for i, mo in enumerate(mos):
            mo['sym'] = '%d.%s' % (i + 1, tail)
    # geometry: symbol, index, charge; Angstrom -> Bohr
    qc.geo_info = numpy.array([[get_atom_symbol(a[0]), a[1], str(float(a[2]))] for a in qc.geo_info])
    qc.geo_spec = numpy.array(qc.geo_spec, dtype=float)
    if angstrom:
        qc.geo_spec *= AA_TO_A0
    # normalisation of the contracted functions (molden.py:375-398)
    norm = cartesian_self_overlap(qc.ao_spec)
    if numpy.abs(norm - 1.0).max() > 1e-5:
        display('The atomic orbitals are not normalized correctly, renormalizing...\n')
        if not by_orca:
            j = 0
            for ao in qc.ao_spec:
                ao['coeffs'][:, 1] /= numpy.sqrt(norm[j])
                l = lquant[ao['type']]
                j += (l + 1) * (l + 2) // 2
        else:
            qc.ao_spec[0]['N'] = 1 / numpy.sqrt(norm[:, numpy.newaxis])
        if cart_flags:
            # explicit Cartesian flags ([6D], [10F], ...): the CCA standard omits the Cartesian normalisation
            # sqrt((2lx-1)!!(2ly-1)!!(2lz-1)!!/(2l-1)!!) of every function (molden.py:394-398, cy_overlap.pyx:24-52)
            from .cy_overlap import ommited_cca_norm
            cca = ommited_cca_norm(numpy.require(qc.ao_spec.get_lxlylz(), dtype=numpy.intc, requirements='CA'))
            for mo in mos:
                mo['coeffs'] *= cca
    qc.mo_spec = MOClass(mos)
    qc.mo_spec.update()
    qc.ao_spec.update()
    return qc


from .read_wf import read_wfn, read_wfx          # noqa: E402  (primitive-based wave-function files)
from .read_aomix import read_aomix               # noqa: E402
from .read_gamess import read_gamess             # noqa: E402
from .read_glog import read_gaussian_log         # noqa: E402

from .read_cclib import read_with_cclib, convert_cclib   # noqa: E402  (cclib itself is imported on use only)

readers = {'gaussian.fchk': read_gaussian_fchk, 'fchk': read_gaussian_fchk, 'molden': read_molden,
           'wfn': read_wfn, 'wfx': read_wfx, 'aomix': read_aomix, 'gamess': read_gamess,
           'gaussian.log': read_gaussian_log, 'gaussian_log': read_gaussian_log, 'cclib': read_with_cclib}
_OTHER = ('native',)


_MAGIC = (('molden', re.compile(r'\[[ ]{,}[Mm]olden[ ]+[Ff]ormat[ ]{,}\]')),
          ('gamess', re.compile(r'[Gg][Aa][Mm][Ee][Ss][Ss]')),
          ('gaussian_log', re.compile(r'[Cc]opyright[,\s\(\)c0-9]+[Gg]aussian\s{,},\s+Inc.')),
          ('aomix', re.compile(r'\[[ ]{,}[Aa][Oo][Mm]ix[ ]+[Ff]ormat[ ]{,}\]')))


def find_itype(fname, extension=None):
    """file type from the extension (fchk, wfx, wfn, native containers) or from magic strings in the content, tried in
    the reference's order: Molden, GAMESS-US, Gaussian log, AOMix (read/tools.py:71-145)"""
    name = fname if isinstance(fname, str) else getattr(fname, 'name', '')
    ext = (extension or name.split('.')[-1]).lower()
    if ext in ('fchk', 'wfx', 'wfn'):
        return ext
    if ext in ('numpy', 'npz', 'hdf5', 'h5'):
        return 'native'
    if isinstance(fname, str):
        with open(fname, 'rb') as f:
            text = f.read().decode('iso-8859-1')
    else:
        text = fname.read()
        fname.seek(0)
        if isinstance(text, bytes):
            text = text.decode('iso-8859-1')
    for itype, rex in _MAGIC:
        if rex.search(text):
            return itype
    raise NotImplementedError('File format not reccognized or reader not implemented!')


def main_read(fname, all_mo=False, spin=None, itype='auto', check_norm=False, **kwargs):
    """High-level reading interface (read/high_level.py:33-79) for the formats built here."""
    if itype == 'auto':
        itype = find_itype(fname)
    if itype not in readers:
        if itype in _OTHER:
            raise NotImplementedError('orbkit_b200 reads Gaussian .fchk / .log, Molden, .wfn, .wfx, GAMESS-US and AOMix files and cclib data; use the reference\'s reader for %r and pass '
                                      'its QCinfo (or QCinfo(qc.todict())) to orbkit_b200' % itype)
        raise KeyError(itype)
    display('Loading data from {0} type file {1}\n'.format(itype, fname if isinstance(fname, str)
                                                           else getattr(fname, 'name', '<stream>')))
    qc = readers[itype](fname, all_mo=all_mo, spin=spin, **kwargs)
    if check_norm:                               # read/high_level.py:74-77; the overlap integrals run on the device
        from .analytical_integrals import check_mo_norm
        deviation = check_mo_norm(qc)
        if deviation >= 1e-5:
            raise ValueError('Bad molecular orbital norm: {0:.4e}'.format(deviation))
    return qc
