"""Reader for Gaussian .log files written with GFINPUT and POP=FULL (orbkit/read/gaussian_log.py:9-405): the geometry
(standard, else input orientation), the basis set "in the form of general basis input", the orbital symmetries and the
"Molecular Orbital Coefficients" blocks (five orbitals per block; for pure-spherical bases the (l, m) of every function is
taken from the row labels of the first block).

Mechanism: the selected Link-1 job is scanned ONCE for the positions of its geometry / basis / symmetry / coefficient
sections; the sections the caller selects (`i_geo`, `i_ao`, `i_mo`; never asked for interactively here) are then parsed
as tables.  The QCinfo is identical to the reference reader's on the reference's own Gaussian test outputs
(tests/test_host.py, tests/golden/read_glog.npz) -- including its behaviour that `spin=` only works for files WITHOUT an
"Orbital symmetries" section (with one, no per-orbital spin is recorded and the request raises, gaussian_log.py:392-395).
"""
import re

import numpy

from .display import display
from .orbitals import AOClass, MOClass
from .qcinfo import QCinfo
from .read import AA_TO_A0, get_atom_symbol
from .read_wf import _text_of
from .tools import lquant


def _pick(count, i, what):
    if count == 0:
        raise IndexError(what)
    if count == 1:
        return 0
    try:
        i = list(range(count))[i]
    except (IndexError, TypeError):
        raise IOError('\tPlease give an integer from 0 to {0} (default: {0})! '.format(count - 1))
    display('\tSelecting the %s' % ('last element.' if i == count - 1 else 'element %d.' % i))
    return i


def _ndeg(l, cartesian):
    return (l + 1) * (l + 2) // 2 if cartesian else 2 * l + 1


def read_gaussian_log(fname, all_mo=False, spin=None, orientation='standard', i_link=-1, i_geo=-1, i_ao=-1, i_mo=-1,
                      interactive=True, **kwargs):
    text, name = _text_of(fname)
    lines = text.split('\n')
    links = [i for i, l in enumerate(lines) if ' Entering Link 1' in l]
    try:
        display('\tFound %d linked GAUSSIAN files.' % len(links))
        i_link = _pick(len(links), i_link, 'link')
    except IndexError:
        raise IOError('Found no `Entering Link 1` keyword!')
    lo, hi = links[i_link], (links + [len(lines)])[i_link + 1]
    # ---- one scan: where the sections of this job are ---------------------------------------------------------------
    geo = {'standard': [], 'input': []}
    basis_flags, ao_secs, sym_secs, mo_secs, states = [], [], [], [], []
    etot = None
    for i in range(lo, hi):
        l = lines[i]
        if ' orientation:' in l:
            for key in geo:
                if '%s orientation:' % key in l.lower():
                    geo[key].append(i)
            if orientation not in geo and '%s orientation:' % orientation in l.lower():
                geo.setdefault(orientation, []).append(i)
        elif 'Standard basis:' in l or 'General basis read from cards:' in l:
            if '(5D, 7F)' in l:
                basis_flags.append((i, False))
            elif '(6D, 10F)' in l:
                basis_flags.append((i, True))
            else:
                raise IOError('Please apply a Spherical Harmonics (5D, 7F) or a Cartesian Gaussian Basis Set (6D, 10F)!')
        elif 'AO basis set in the form of general basis input' in l:
            ao_secs.append(i)
        elif 'The electronic state is ' in l:
            states.append(l.split()[-1][:-1])
        elif 'Orbital symmetries:' in l:
            sym_secs.append(i)
        elif 'Orbital Coefficients:' in l:
            kind = l.split()[0]
            if kind != 'Beta':
                mo_secs.append([kind, i, None])
            else:
                mo_secs[-1][0], mo_secs[-1][2] = 'Alpha&Beta', i
        elif 'E(' in l:
            try:
                etot = float(l.split('=')[1].split()[0])
            except (IndexError, ValueError):
                pass
    display('\nContent of the GAUSSIAN .log file:')
    try:
        i_geo = _pick(len(geo.get(orientation, [])), i_geo, 'geometry')
    except IndexError:
        orientation = 'input'
        try:
            i_geo = _pick(len(geo['input']), i_geo, 'geometry')
        except IndexError:
            raise IOError('Found no geometry section! Are you sure this is a GAUSSIAN .log file?')
    try:
        i_ao = _pick(len(ao_secs), i_ao, 'atomic orbitals')
    except IndexError:
        raise IOError('Write GFINPUT in your GAUSSIAN route section to print the basis set information!')
    try:
        i_mo = _pick(len(mo_secs), i_mo, 'molecular orbitals')
    except IndexError:
        raise IOError('Write IOP(6/7=3) in your GAUSSIAN route section to print\n all molecular orbitals!')
    if spin is not None:
        if spin not in ('alpha', 'beta'):
            raise IOError('`spin=%s` is not a valid option' % spin)
        display('Reading only molecular orbitals of spin %s.' % spin)
    qc = QCinfo()
    qc.etot = etot if etot is not None else qc.etot
    # ---- geometry: four header lines, rows up to the closing dashes ----------------------------------------------------
    info, spec, i = [], [], geo[orientation][i_geo] + 5
    while i < hi:
        t = lines[i].split()
        info.append([get_atom_symbol(t[1]), t[0], float(t[1])])
        spec.append([float(v) for v in t[3:]])
        if '-----------' in lines[i + 1]:
            break
        i += 1
    qc.geo_info = numpy.array(info)
    qc.geo_spec = numpy.array(spec, dtype=float) * AA_TO_A0
    # ---- basis set ---------------------------------------------------------------------------------------------------
    a0 = ao_secs[i_ao]
    before = [c for pos, c in basis_flags if pos < a0]
    cartesian = (before or [c for _, c in basis_flags] or [True])[-1]
    aos, new_atom, at_num, shell, row, basis_count = [], True, 0, '', 0, 0
    i = a0 + 1
    while i < hi:
        l, t = lines[i], lines[i].split()
        if ' ****' in l:
            new_atom = True
            if i + 1 >= hi or not lines[i + 1].split():
                break
        elif new_atom:
            new_atom, at_num = False, int(t[0]) - 1
        elif len(t) == 4:
            shell, pnum, row = t[0].lower(), int(t[1]), 0
            for ch in shell:
                basis_count += _ndeg(lquant[ch], cartesian)
                aos.append({'atom': at_num, 'type': ch, 'pnum': pnum, 'coeffs': numpy.zeros((pnum, 2))})
                if not cartesian:
                    aos[-1]['lm'] = []
        else:
            vals = numpy.array(l.replace('D', 'e').split(), dtype=numpy.float64)
            for k in range(len(shell)):
                aos[-len(shell) + k]['coeffs'][row, :] = [vals[0], vals[1 + k]]
            row += 1
        i += 1
    # ---- orbital symmetries: the section in front of the selected coefficient section ------------------------------------
    mo_type, m0, m1 = mo_secs[i_mo]
    orb_sym, orb_spin = [], []
    prev_end = mo_secs[i_mo - 1][1] if i_mo > 0 else lo
    sym_here = [s for s in sym_secs if prev_end <= s < m0]
    if sym_here:
        add, i = '', sym_here[-1] + 1
        while i < m0 and 'electronic state' not in lines[i]:
            l = lines[i]
            if 'Alpha' in l:
                add = '_a'
            elif 'Beta' in l:
                add = '_b'
            orb_sym += [s + add for s in l[18:].replace('(', '').replace(')', '').split()]
            i += 1
    else:
        add = ''
        if 'Alpha' in mo_type:
            add, orb_spin = '_a', ['alpha'] * basis_count
        orb_sym = ['A1' + add] * basis_count
        if 'Beta' in mo_type:
            orb_spin += ['beta'] * basis_count
            orb_sym += ['A1_b'] * basis_count
    mos, seen = [], {}
    for k, s in enumerate(orb_sym):
        seen[s] = seen.get(s, 0) + 1
        mos.append({'coeffs': numpy.zeros(basis_count), 'energy': 0., 'sym': '%d.%s' % (seen[s], s)})
        if orb_spin:
            mos[-1]['spin'] = orb_spin[k]
    # ---- coefficient blocks ----------------------------------------------------------------------------------------------
    offset, index, fresh, c_sao, old_ao = 0, [], True, 0, -1
    i = m0 + 1
    while i < hi:
        l = lines[i]
        if 'Orbital Coefficients:' in l:                     # the Beta set of an unrestricted calculation follows
            fresh = True
            i += 1
            continue
        head = l[:21].split()
        if not head:
            cols = l[21:].split()
            if fresh:
                index, fresh = [offset + k for k in range(len(cols))], False
            else:
                for k, j in enumerate(index):
                    mos[j]['occ_num'] = int('O' in cols[k]) * (1 if mo_type in 'Alpha&Beta' else 2)
        elif 'Eigenvalues' in head:
            cols = l[21:].replace('-', ' -').split()
            for k, j in enumerate(index):
                mos[j]['occ_num' if mo_type == 'Natural' else 'energy'] = float(cols[k])
        else:
            if not re.fullmatch(r'[+-]?\d+', head[0]):
                del mos[index[-1] + 1:]
                break
            cols = l[21:].replace('-', ' -').split()
            if not cartesian and offset == 0:
                lab = l[:14].split()
                if old_ao != lab[-1] or len(lab) == 4:
                    old_ao = lab[-1]
                    c_sao += 1
                m = l[14:21].replace(' ', '').lower()
                p = 'yzx'.find(m) if len(m) == 1 else -1
                m = p - 1 if p != -1 else (0 if m == '' else int(m))
                aos[c_sao - 1]['lm'].append((lquant[l[13].lower()], m))
            for k, j in enumerate(index):
                mos[j]['coeffs'][int(head[0]) - 1] = float(cols[k])
            if int(head[0]) == basis_count:
                fresh, offset = True, index[-1] + 1
                if index[-1] + 1 == len(orb_sym):
                    break
        i += 1
    if not all_mo:
        mos = [mo for mo in mos if mo['occ_num'] >= 0.0000001]
    if spin is not None:
        if not orb_spin:
            raise IOError('You requested `%s` orbitals, but None of them are present.' % spin)
        mos = [mo for mo in mos if mo['spin'] == spin]
    qc.ao_spec = AOClass(aos)
    if not cartesian:
        qc.ao_spec.spherical = True
    qc.mo_spec = MOClass(mos)
    qc.mo_spec.update()
    qc.ao_spec.update()
    return qc
