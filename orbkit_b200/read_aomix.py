"""Reader for AOMix input files as Turbomole's `tm2aomix` writes them (orbkit/read/aomix.py:11-326): a Molden-like layout
(`[AOMix Format]`, `[SCF Energy / Hartree]`, `[Atoms]`, `[GTO]`, `[MO]`) whose coefficient lines carry the function label
(`dxx`, `dx2`, `fx2y`, ...), from which the Cartesian exponents of every function are taken.

Mechanism: the selected `[AOMix Format]` block is cut into its bracketed sections once; each section is a small table
parsed on its own, instead of the reference's line-by-line state machine.  The resulting QCinfo is identical
(tests/test_host.py compares every flat array with the reference reader's output, tests/golden/h2o_turbomole_aomix.npz).
"""
import re

import numpy

from .display import display
from .orbitals import AOClass, MOClass
from .qcinfo import QCinfo
from .read import AA_TO_A0, get_atom_symbol
from .read_wf import _text_of, select_spin, spin_check
from .tools import lquant

_RE_HEAD = re.compile(r'\[\s*aomix\s+format\s*\]', re.I)
_RE_SECTION = re.compile(r'^\s*\[([^\]]+)\](.*)$')


def _label_exponents(label):
    """'dx2' / 'dxy' / 'fx2y' / 'px' -> (lx, ly, lz): letters name an axis, a digit raises the power of the axis named
    last (aomix.py:262-276); the leading shell letter is skipped"""
    e, last = [0, 0, 0], None
    for ch in label[1:]:
        if ch.isdigit():
            e[last] += int(ch) - 1
        elif ch in 'xyz':
            last = 'xyz'.index(ch)
            e[last] += 1
    return tuple(e)


def read_aomix(fname, all_mo=False, spin=None, i_md=-1, interactive=True, created_by_tmol=True, **kwargs):
    """QCinfo of an AOMix file; `i_md` selects the `[AOMix Format]` block of files that hold several (never asked for
    interactively here); `created_by_tmol`: Cartesian d / f / g coefficients get Turbomole's normalisation
    sqrt((2lx-1)!!(2ly-1)!!(2lz-1)!!) (aomix.py:300-321)"""
    text, name = _text_of(fname)
    lines = text.splitlines()
    starts = [i for i, l in enumerate(lines) if _RE_HEAD.search(l)]
    if '[AOMix Format]' not in [l.rstrip('\n') for l in lines] or not starts:
        raise IOError('The input file %s is no valid aomix file!\n\nIt does not contain the keyword: [AOMix Format]\n' % name)
    if len(starts) > 1:
        try:
            i_md = list(range(len(starts)))[i_md]
        except IndexError:
            raise IOError('\tPlease give an integer from 0 to %d!' % (len(starts) - 1))
        display('\tFound %d [AOMix Format] keywords; selecting the element with index %d.' % (len(starts), i_md))
    else:
        i_md = 0
    block = lines[starts[i_md] + 1:(starts + [len(lines)])[i_md + 1]]
    # ---- cut the block into sections ------------------------------------------------------------------------
    sections, cur = [], None
    for l in block:
        m = _RE_SECTION.match(l)
        if m:
            cur = [m.group(1).strip().lower(), m.group(2), []]
            sections.append(cur)
        elif cur is not None:
            cur[2].append(l)
    sec = {}
    for key, rest, body in sections:
        sec.setdefault(key, (rest, body))
    if 'sto' in sec:
        raise IOError('orbkit does not work for STOs!\nEXIT\n')
    qc = QCinfo()
    if 'scf energy / hartree' in sec:
        try:
            qc.etot = float(sec['scf energy / hartree'][1][0].split()[0])
        except IndexError:
            pass
    # ---- [Atoms] ----------------------------------------------------------------------------------------------
    rest, body = sec['atoms']
    angstrom = 'Angs' in rest
    rows = [l.split() for l in body if l.split()]
    qc.geo_info = numpy.array([[get_atom_symbol(r[0]), r[1], float(r[2])] for r in rows])
    qc.geo_spec = numpy.array([[float(v) for v in r[3:]] for r in rows], dtype=float)
    if angstrom:
        qc.geo_spec *= AA_TO_A0
    # ---- [GTO] ------------------------------------------------------------------------------------------------
    aos, new_atom, at_num, shell, row = [], True, 0, '', 0
    for l in sec['gto'][1]:
        t = l.split()
        if not t:
            new_atom = True
        elif new_atom:
            new_atom, at_num = False, int(t[0]) - 1
        elif len(t) == 3 and re.fullmatch(r'[+-]?\d+', t[1]):
            shell, pnum, row = t[0], int(t[1]), 0
            aos += [{'atom': at_num, 'type': ch, 'pnum': pnum, 'coeffs': numpy.zeros((pnum, 2))} for ch in shell]
        else:
            vals = numpy.array(l.replace('D', 'e').split(), dtype=numpy.float64)
            for i in range(len(shell)):
                aos[-len(shell) + i]['coeffs'][row, :] = [vals[0], vals[1 + i]]
            row += 1
    n_basis = sum((lquant[ao['type']] + 1) * (lquant[ao['type']] + 2) // 2 for ao in aos)
    # ---- [MO] -------------------------------------------------------------------------------------------------
    keys = {'Sym': 'sym', 'Ene': 'energy', 'Occup': 'occ_num', 'Spin': 'spin'}
    mos, labels, fresh = [], [], True
    has_alpha = has_beta = restricted = False
    for l in sec['mo'][1]:
        if '=' in l:
            if fresh:
                mos.append({'coeffs': numpy.zeros(n_basis), 'sym': '%d.1' % (len(mos) + 1)})
                fresh = False
            k, v = l.replace('\n', '').replace(' ', '').split('=')[:2]
            if k not in keys:
                continue
            if k == 'Spin':
                v = v.lower()
                has_alpha |= v == 'alpha'
                has_beta |= v == 'beta'
            elif k != 'Sym':
                v = float(v)
                if k == 'Occup':
                    restricted |= v > 1. + 1e-4
            elif '.' not in v:
                a = re.search(r'\d+', v).group()
                v = '%s.1' % a if a == v else v.replace(a, '%s.' % a, 1)
            mos[-1][keys[k]] = v
        elif '[' in l:
            break
        else:
            t = l.split()
            if not t:
                continue
            fresh = True
            mos[-1]['coeffs'][int(t[0]) - 1] = float(t[-1])
            if len(mos) == 1:
                labels.append(t[-2])
    spin_check(spin, restricted, has_alpha, has_beta)
    # ---- exponents from the function labels of the first orbital ------------------------------------------------
    exps, count = [_label_exponents(s) for s in labels], 0
    for ao in aos:
        n = (lquant[ao['type']] + 1) * (lquant[ao['type']] + 2) // 2
        ao['lxlylz'] = numpy.array(exps[count:count + n], dtype=numpy.int64)
        count += n
    is_tmol_cart = not (len(mos) % len(mos[0]['coeffs']))
    if not all_mo:
        mos = [mo for mo in mos if mo['occ_num'] >= 0.0000001]
    mos = select_spin(mos, restricted, spin=spin)
    if is_tmol_cart and created_by_tmol:
        display('\nFound a Cartesian basis set in the AOMix file.\nWe assume that this file has been created by Turbomole.\n'
                'Applying a conversion to the molecular orbital coefficients, in order to get normalized orbitals.')
        dfact = lambda n: 1 if n <= 0 else n * dfact(n - 2)
        i = 0
        for ao in aos:
            for e in ao['lxlylz']:
                if sum(e) > 1:
                    norm = numpy.sqrt(dfact(2 * e[0] - 1) * dfact(2 * e[1] - 1) * dfact(2 * e[2] - 1))
                    for mo in mos:
                        mo['coeffs'][i] *= norm
                i += 1
    qc.ao_spec = AOClass(aos)
    qc.mo_spec = MOClass(mos)
    qc.mo_spec.update()
    qc.ao_spec.update()
    return qc
