"""Reader for GAMESS-US output files (orbkit/read/gamess.py:8-420): geometry, the `ATOMIC BASIS SET` table, the
`EIGENVECTORS` blocks (five orbitals per block: numbers, energies, symmetry labels, one row per basis function with its
Cartesian label) and the occupied-orbital counts.

Mechanism: the file is scanned once for the section headers; each section is then parsed as a table on its own.  The
QCinfo is identical to the reference reader's (tests/test_host.py, tests/golden/formaldehyde_gamess.npz) -- including its
way of numbering the orbitals of a symmetry (`<running MO number>.<label>` from the second orbital of a label on,
gamess.py:261-263).  `read_properties` (CIS states, dipole moments, populations) is not built.
"""
import re

import numpy

from .display import display
from .orbitals import AOClass, MOClass
from .qcinfo import QCinfo
from .read import AA_TO_A0, get_atom_symbol
from .read_wf import _text_of
from .tools import lquant

_RE_NUM = re.compile(r'-?\d+\.\d+')


def read_gamess(fname, all_mo=False, spin=None, read_properties=False, **kwargs):
    if read_properties:
        raise NotImplementedError('read_gamess(read_properties=True): CIS states / dipole moments / populations are not built')
    text, name = _text_of(fname)
    lines = text.split('\n')
    n = len(lines)
    geo_key, geo_prev, geo_skip = ' ATOM      ATOMIC                      COORDINATES', '', 1
    mokey, restricted = 'EIGENVECTORS', True
    geo_info, geo_spec, angstrom = [], [], False
    shells_of, order = {}, []                  # atom type -> list of contractions; atom types in order of appearance
    mos, labels, occ = [], [], None
    sym_count, n_seen = {}, 0
    spin_now = None                            # 'alpha' / 'beta' inside an unrestricted orbital set
    has_alpha = has_beta = False
    i = 0
    while i < n:
        line = lines[i]
        tok = line.split()
        if 'RUNTYP=OPTIMIZE' in line:
            geo_key, geo_prev, geo_skip = ' COORDINATES OF ALL ATOMS ARE', '***** EQUILIBRIUM GEOMETRY LOCATED *****', 2
            if 'SCFTYP=UHF' in line:
                mokey, restricted = ' SET ****', False
            else:
                mokey = 'EIGENVECTORS'
        elif geo_key in line and i > 0 and geo_prev in lines[i - 1]:
            angstrom = '(BOHR)' not in line
            i += 1 + geo_skip
            count = 0
            while i < n and len(lines[i]) >= 1 and lines[i].strip():
                t = lines[i].split()
                geo_info.append([t[0], count + 1, t[1]])
                geo_spec.append([float(v) for v in t[2:]])
                count += 1
                i += 1
            continue
        elif 'ATOMIC BASIS SET' in line:
            i += 7                             # header of the table
            cur, new_shell, shell = None, False, ''
            while i < n and ' TOTAL NUMBER OF BASIS SET SHELLS' not in lines[i]:
                t = lines[i].split()
                if len(t) == 1:
                    cur = []
                    order.append((t[0], cur))
                    new_shell = False
                elif not t:
                    if not new_shell:
                        new_shell = True
                elif cur is not None:
                    vals = [float(v) for v in t[3:]]
                    if new_shell:
                        shell = t[1].lower().replace('l', 'sp')
                        for k, ch in enumerate(shell):
                            cur.append({'atom_type': order[-1][0], 'type': ch, 'pnum': 1, 'coeffs': [[vals[0], vals[1 + k]]]})
                        new_shell = False
                    else:
                        for k in range(len(shell)):
                            cur[-len(shell) + k]['coeffs'].append([vals[0], vals[1 + k]])
                            cur[-len(shell) + k]['pnum'] += 1
                i += 1
            continue
        elif '----- ALPHA SET ' in line:
            has_alpha, has_beta, restricted, spin_now = True, False, False, 'alpha'
        elif '----- BETA SET ' in line:
            has_alpha, has_beta, restricted, spin_now = False, True, False, 'beta'
        elif mokey in line and len(tok) < 3:
            skip = 1
            if 'ALPHA' in line:
                has_alpha, spin_now, skip = True, 'alpha', 0
            elif 'BETA' in line:
                has_beta, has_alpha, spin_now, skip = True, False, 'beta', 0
            i += 1 + skip
            # blocks: blank line, orbital numbers, energies, symmetry labels, coefficient rows
            while i < n:
                l = lines[i]
                if ('END OF' in l and 'CALCULATION' in l) or '-----------' in l:
                    break
                if not l.split():
                    nxt = lines[i + 1].split() if i + 1 < n else []
                    if not nxt or not re.fullmatch(r'[+-]?\d+', nxt[0]):
                        break
                    width = len(nxt)
                    energies = lines[i + 2].split()
                    syms = lines[i + 3].split()
                    block = []
                    for k in range(width):
                        n_seen += 1
                        a = syms[k]
                        sym_count[a] = 1 if a not in sym_count else n_seen
                        mo = {'coeffs': [], 'energy': float(energies[k]), 'occ_num': 0.0, 'sym': '%d.%s' % (sym_count[a], a)}
                        if spin_now is not None:
                            mo['sym'] += '_%s' % spin_now[0]
                            mo['spin'] = spin_now
                        block.append(mo)
                    mos += block
                    labels = []
                    i += 4
                    while i < n and lines[i].split():
                        if ('END OF' in lines[i] and 'CALCULATION' in lines[i]) or '-----------' in lines[i]:
                            break
                        labels.append(lines[i][11:17])
                        for k, m in enumerate(_RE_NUM.finditer(lines[i][16:])):
                            block[k]['coeffs'].append(float(m.group()))
                        i += 1
                    continue
                i += 1
            continue
        elif 'NATURAL ORBITALS' in line and len(tok) <= 3:
            display('The natural orbitals are not extracted.')
        elif ' NUMBER OF OCCUPIED ORBITALS (ALPHA)          =' in line or ' NUMBER OF OCCUPIED ORBITALS (ALPHA) KEPT IS    =' in line:
            occ = [int(tok[-1])]
        elif ' NUMBER OF OCCUPIED ORBITALS (BETA )          =' in line or ' NUMBER OF OCCUPIED ORBITALS (BETA ) KEPT IS    =' in line:
            occ.append(int(tok[-1]))
        i += 1
    # ---- basis: one set of contractions per atom type, attached to every atom of that type ------------------------
    basis = {}
    for at_type, shells in order:
        if at_type not in basis:
            basis[at_type] = shells
        elif [s['coeffs'] for s in shells] != [s['coeffs'] for s in basis[at_type]]:
            raise IOError('Different basis sets for the same atom.')
    aos = []
    for k, info in enumerate(geo_info):
        for s in basis[info[0]]:
            aos.append({'atom': info[1] - 1, 'type': s['type'], 'pnum': s['pnum'], 'coeffs': numpy.array(s['coeffs']),
                        'lxlylz': None})
    count = 0
    for ao in aos:
        nfn = (lquant[ao['type']] + 1) * (lquant[ao['type']] + 2) // 2
        ao['lxlylz'] = numpy.array([(s.lower().count('x'), s.lower().count('y'), s.lower().count('z'))
                                    for s in labels[count:count + nfn]], dtype=numpy.int64)
        count += nfn
    for mo in mos:
        mo['coeffs'] = numpy.array(mo['coeffs'])
    # ---- occupations from the numbers of occupied alpha / beta orbitals (gamess.py:349-373) -----------------------
    occ = list(occ) if occ else [0, 0]
    has_alpha = has_beta = False
    if restricted:
        for mo in mos:
            if occ[0] and occ[1]:
                mo['occ_num'] += 2.0
                occ[0] -= 1
                occ[1] -= 1
            if not occ[0] and occ[1]:
                mo['occ_num'] += 1.0
                occ[1] -= 1
            if not occ[1] and occ[0]:
                mo['occ_num'] += 1.0
                occ[0] -= 1
    else:
        for mo in mos:
            if mo.get('spin') == 'alpha' and occ[0] > 0:
                mo['occ_num'] += 1.0
                occ[0] -= 1
                has_alpha = True
            elif mo.get('spin') == 'beta' and occ[1] > 0:
                mo['occ_num'] += 1.0
                occ[1] -= 1
                has_beta = True
    if spin is not None:
        if restricted:
            raise IOError('The keyword `spin` is only supported for unrestricted calculations.')
        if spin not in ('alpha', 'beta'):
            raise IOError('`spin=%s` is not a valid option' % spin)
        if not has_alpha and not has_beta:
            raise IOError('No spin molecular orbitals available')
        if (spin == 'alpha' and not has_alpha) or (spin == 'beta' and not has_beta):
            raise IOError('You requested `%s` orbitals, but None of them are present.' % spin)
        display('Reading only molecular orbitals of spin %s.' % spin)
    if not all_mo:
        mos = [mo for mo in mos if mo['occ_num'] >= 0.0000001]
    if spin is not None:
        mos = [mo for mo in mos if mo['spin'] == spin]
    qc = QCinfo()
    qc.geo_info = numpy.array([[get_atom_symbol(a[0]), a[1], float(a[2])] for a in geo_info])
    qc.geo_spec = numpy.array(geo_spec, dtype=float)
    if angstrom:
        qc.geo_spec *= AA_TO_A0
    qc.ao_spec = AOClass(aos)
    qc.mo_spec = MOClass(mos)
    qc.mo_spec.update()
    qc.ao_spec.update()
    return qc
