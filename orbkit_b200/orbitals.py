"""AO / MO containers: the INPUT CONTRACT of the grid path.

Host-side mirror of the reference's data model (orbkit/orbitals.py: AOClass :25-409,
MOClass :411-832): a list of per-contraction / per-orbital dictionaries plus cached flat NumPy
views reached through the same getter names, so that `core.ao_creator`, `core.mo_creator` and
`core.rho_compute` accept either these classes or the reference's own objects (duck typing on
the getters).  Only what the hot path needs is implemented; the reference's MO selection
mini-language (`MOClass.select`, orbitals.py:896-1210) is reduced to index lists, slices,
'all_mo', 'homo'/'lumo' offsets and 'a:b' ranges thereof.

AO record:  {'atom': int (0-based), 'type': 's'|'p'|..., 'pnum': int (<0: primitives already
            normalised), 'coeffs': (pnum,2) [alpha, c], optional 'lxlylz': (n_fn,3),
            optional 'lm': [(l,m), ...]}
MO record:  {'coeffs': (n_ao,), 'occ_num': float, 'energy': float, 'sym': str[, 'spin': str]}
"""
from collections import UserList
from copy import copy
import re

import numpy

from .tools import exp, lquant, require


import itertools as _itertools
_STAMPS = _itertools.count(1)

class AOClass(UserList):
    def __init__(self, data=None, seq=(), restart=None):
        if isinstance(data, list):
            seq = data
        elif isinstance(data, dict):
            restart = data
        UserList.__init__(self, list(seq))
        self._up_to_date = False
        self.normalized = False
        self.spherical = False
        self._renorm = None
        for name in ('_cont_types', '_nprim_per_cont', '_prim_coeffs', '_assign_prim_to_cont',
                     '_assign_cont_to_atoms', '_lxlylz', '_assign_lxlylz_to_cont',
                     '_nlxlylz_per_cont', '_lm', '_assign_lm_to_cont'):
            setattr(self, name, None)
        if restart is not None:
            self._from_flat(restart)

    # -- flat <-> list-of-dict ------------------------------------------------------------
    def _from_flat(self, r):
        """Rebuild the records from the flat arrays of todict() (orbitals.py:101-121,318-334)."""
        a2c = numpy.asarray(r['_assign_prim_to_cont'])
        l2c = numpy.asarray(r['_assign_lxlylz_to_cont'])
        self.spherical = bool(r['spherical'])
        self.normalized = bool(r['normalized'])
        lm = r.get('_lm', None)
        lm2c = r.get('_assign_lm_to_cont', None)
        normalized = self.normalized
        self.data = []
        for ic in range(len(r['_assign_cont_to_atoms'])):
            npr = int(numpy.asarray(r['_nprim_per_cont'])[ic])
            rec = {'atom': int(numpy.asarray(r['_assign_cont_to_atoms'])[ic]),
                   'type': str(numpy.asarray(r['_cont_types'])[ic]),
                   'pnum': -npr if normalized else npr,
                   'coeffs': numpy.array(numpy.asarray(r['_prim_coeffs'])[a2c == ic], dtype=float),
                   'lxlylz': numpy.array(numpy.asarray(r['_lxlylz'])[l2c == ic], dtype=numpy.intc)}
            if self.spherical and lm is not None:
                rec['lm'] = [tuple(int(v) for v in t)
                             for t in numpy.asarray(lm)[numpy.asarray(lm2c) == ic]]
            self.data.append(rec)
        if r.get('N', None) is not None:
            self.data[0]['N'] = numpy.asarray(r['N'], dtype=float)
        self._up_to_date = False

    def todict(self):
        self.update()
        return {'normalized': self.normalized, 'spherical': self.spherical,
                '_assign_cont_to_atoms': self._assign_cont_to_atoms,
                '_cont_types': self._cont_types, '_nprim_per_cont': self._nprim_per_cont,
                '_prim_coeffs': self._prim_coeffs,
                '_assign_prim_to_cont': self._assign_prim_to_cont, '_lxlylz': self._lxlylz,
                '_assign_lxlylz_to_cont': self._assign_lxlylz_to_cont,
                '_nlxlylz_per_cont': self._nlxlylz_per_cont, '_lm': self._lm,
                '_assign_lm_to_cont': self._assign_lm_to_cont,
                'parent_class_name': self.__module__ + '.' + self.__class__.__name__}

    # -- list protocol: any mutation invalidates the flat views -----------------------------
    def __getitem__(self, item):
        if isinstance(item, (int, numpy.integer)):
            return self.data[item]
        out = AOClass(seq=self.data[item])
        out.spherical = self.spherical
        out.update()
        return out

    def __setitem__(self, i, item):
        self.data[i] = item
        self._up_to_date = False

    def __delitem__(self, i):
        del self.data[i]
        self._up_to_date = False

    def append(self, item):
        self.data.append(item)
        self._up_to_date = False

    def extend(self, other):
        self.data.extend(other)
        self._up_to_date = False

    def remove(self, item):
        self.data.remove(item)
        self._up_to_date = False

    def __eq__(self, other):
        if other is None or (isinstance(other, list) and other == []):
            return not self.data
        if not isinstance(other, AOClass):
            raise TypeError('Comparing of AOClass to non AOClass object not defined')
        self.update()
        other.update()
        if self.spherical != other.spherical or self.normalized != other.normalized:
            return False
        keys = ['_assign_cont_to_atoms', '_nprim_per_cont', '_prim_coeffs', '_assign_prim_to_cont',
                '_lxlylz', '_assign_lxlylz_to_cont']
        if self.spherical:
            keys += ['_lm', '_assign_lm_to_cont']
        for k in keys:
            a, b = numpy.asarray(getattr(self, k)), numpy.asarray(getattr(other, k))
            if a.shape != b.shape or not numpy.allclose(a, b):
                return False
        return list(self._cont_types) == list(other._cont_types)

    def __str__(self):
        return '\n'.join(self.get_labels())

    # -- derive the flat views (orbitals.py:203-301) -----------------------------------------
    def update(self):
        if not self.data:
            raise ValueError('ao_spec not initialized')
        for i, rec in enumerate(self.data):
            if not isinstance(rec, dict):
                raise ValueError('ao_spec[{0}] has to be a dictionary'.format(i))
            missing = [k for k in ('atom', 'type', 'pnum', 'coeffs') if k not in rec]
            if self.spherical and 'lm' not in rec:
                missing.append('lm')
            if missing:
                raise ValueError('ao_spec[{0}] misses {1}'.format(i, str(missing)))
        atoms, types, nprim, p2c, prim = [], [], [], [], []
        lxlylz, l2c, lm, lm2c, norm_flags = [], [], [], [], []
        for i, rec in enumerate(self.data):
            c = numpy.asarray(rec['coeffs'], dtype=float).reshape(-1, 2)
            atoms.append(rec['atom'])
            types.append(rec['type'])
            nprim.append(len(c))
            prim.append(c)
            p2c.extend([i] * len(c))
            ll = rec['lxlylz'] if 'lxlylz' in rec else exp[lquant[rec['type']]]
            lxlylz.extend([tuple(int(v) for v in t) for t in ll])
            l2c.extend([i] * len(ll))
            norm_flags.append(rec['pnum'] < 0)
            if self.spherical:
                for t in rec['lm']:
                    lm.append((int(t[0]), int(t[1])))
                    lm2c.append(i)
        if all(norm_flags) != any(norm_flags):
            raise ValueError('Either all or none of the atomic orbitals have to be normalized!')
        self.normalized = all(norm_flags)
        self._assign_cont_to_atoms = require(atoms, dtype='i')
        self._cont_types = types
        self._nprim_per_cont = require(nprim, dtype='i')
        self._prim_coeffs = require(numpy.concatenate(prim, axis=0), dtype='f')
        self._assign_prim_to_cont = require(p2c, dtype='i')
        self._lxlylz = require(numpy.asarray(lxlylz).reshape(-1, 3), dtype='i')
        self._assign_lxlylz_to_cont = require(l2c, dtype='i')
        self._nlxlylz_per_cont = require(numpy.bincount(self._assign_lxlylz_to_cont,
                                                        minlength=len(self.data)), dtype='i')
        if self.spherical:
            self._lm = lm
            self._assign_lm_to_cont = require(lm2c, dtype='i')
        else:
            self._lm, self._assign_lm_to_cont = None, None
        self._renorm = (numpy.asarray(self.data[0]['N'], dtype=float)
                        if 'N' in self.data[0] else None)
        self._up_to_date = True
        self._stamp = next(_STAMPS)       # identifies this state of the flat arrays (engine handle caches)

    def is_normlized(self, force=False):
        if force or not self._up_to_date:
            self.update()
        return copy(self.normalized)

    def set_lm_dict(self, p=(1, 0)):
        """(l,m) labels per contraction in the readers' order: m = 0,+1,-1,+2,-2,... and for p
        shells the order given by `p` (default x,y,z = (1,1),(1,-1),(1,0)) (orbitals.py:303-316)."""
        for rec in self.data:
            l = lquant[rec['type']]
            rec['lm'] = []
            for m in (range(0, l + 1) if l != 1 else p):
                rec['lm'].append((l, m))
                if m != 0:
                    rec['lm'].append((l, -m))
        self.spherical = True
        self._up_to_date = False

    def _get(self, name):
        if not self._up_to_date:
            self.update()
        return copy(getattr(self, name))

    def get_assign_cont_to_atoms(self): return self._get('_assign_cont_to_atoms')
    def get_cont_types(self): return self._get('_cont_types')
    def get_nprim_per_cont(self): return self._get('_nprim_per_cont')
    def get_prim_coeffs(self): return self._get('_prim_coeffs')
    def get_assign_prim_to_cont(self): return self._get('_assign_prim_to_cont')
    def get_lxlylz(self): return self._get('_lxlylz')
    def get_assign_lxlylz_to_cont(self): return self._get('_assign_lxlylz_to_cont')
    def get_nlxlylz_per_cont(self): return self._get('_nlxlylz_per_cont')
    def get_lm(self): return self._get('_lm')
    def get_assign_lm_to_cont(self): return self._get('_assign_lm_to_cont')
    def get_renorm(self): return self._get('_renorm')

    def get_normalized(self):
        if not self._up_to_date:
            self.update()
        return int(self.normalized)

    def get_old_ao_spherical(self):
        if not self._up_to_date:
            self.update()
        return list(zip(self.get_assign_lm_to_cont(), self.get_lm())) if self.spherical else []

    def get_labels(self):
        if not self._up_to_date:
            self.update()
        atoms = self._assign_cont_to_atoms
        if self.spherical:
            return ['l,m=%s,atom=%d' % (self._lm[i], atoms[j])
                    for i, j in enumerate(self._assign_lm_to_cont)]
        return ['lxlylz=%s,atom=%d' % (self._lxlylz[i], atoms[j])
                for i, j in enumerate(self._assign_lxlylz_to_cont)]

    def get_ao_num(self):
        if not self._up_to_date:
            self.update()
        return len(self._lm) if self.spherical else len(self._lxlylz)


class MOClass(UserList):
    def __init__(self, seq=(), restart=None, data=None):
        if isinstance(data, list):
            seq = data
        elif isinstance(data, dict):
            restart = data
        elif isinstance(seq, dict):
            restart, seq = seq, ()
        UserList.__init__(self, list(seq))
        self._up_to_date = False
        self.coeffs = self.occ = self.eig = self.sym = self.spin = None
        self.selected_mo = None
        self.selection_string = None
        self.spinpolarized = False
        self.alpha_index, self.beta_index = [], []
        if restart is not None:
            coeffs = numpy.asarray(restart['coeffs'], dtype=float)
            n = len(coeffs)
            occ = numpy.asarray(restart['occ'], dtype=float)
            eig = numpy.asarray(restart.get('eig', numpy.zeros(n)), dtype=float)
            sym = numpy.asarray(restart.get('sym', ['%d.1' % (i + 1) for i in range(n)]), dtype=str)
            spin = numpy.asarray(restart.get('spin', ['alpha'] * n), dtype=str)
            self.data = [{'coeffs': coeffs[i], 'energy': eig[i], 'occ_num': occ[i],
                          'sym': str(sym[i]), 'spin': str(spin[i])} for i in range(n)]

    def todict(self):
        self.update()
        return {'coeffs': self.coeffs, 'occ': self.occ, 'eig': self.eig, 'sym': self.sym,
                'spin': self.spin, 'spinpolarized': self.spinpolarized,
                'alpha_index': self.alpha_index, 'beta_index': self.beta_index,
                'selected_mo': self.selected_mo, 'selection_string': self.selection_string,
                'parent_class_name': self.__module__ + '.' + self.__class__.__name__}

    # -- list protocol ----------------------------------------------------------------------
    def __getitem__(self, item):
        if isinstance(item, (int, numpy.integer)):
            return self.data[item]
        if isinstance(item, slice):
            return MOClass(self.data[item])
        if isinstance(item, str):
            return self.select(item)
        idx = numpy.asarray(list(item))
        if idx.ndim != 1:
            raise ValueError('Only 1D arrays can be used for indexing!')
        if idx.dtype == bool:
            idx = numpy.nonzero(idx)[0]
        elif idx.dtype.kind not in 'iu':
            return self.select(list(item))
        out = MOClass([self.data[int(i)] for i in idx])
        out.selected_mo = idx
        out.update()
        return out

    def __setitem__(self, i, item):
        self.data[i] = item
        self._up_to_date = False

    def __delitem__(self, i):
        del self.data[i]
        self._up_to_date = False

    def append(self, item):
        self.data.append(item)
        self._up_to_date = False

    def extend(self, other):
        self.data.extend(other)
        self._up_to_date = False

    def remove(self, item):
        self.data.remove(item)
        self._up_to_date = False

    def __eq__(self, other):
        if other is None or (isinstance(other, list) and other == []):
            return not self.data
        if not isinstance(other, MOClass):
            raise TypeError('Comparing of MOClass to non MOClass object not defined')
        self.update()
        other.update()
        return (self.coeffs.shape == other.coeffs.shape and
                numpy.allclose(self.coeffs, other.coeffs) and
                numpy.allclose(self.occ, other.occ) and numpy.allclose(self.eig, other.eig) and
                list(self.sym) == list(other.sym))

    def __str__(self):
        return '\n'.join(self.get_labels())

    # -- flat views (orbitals.py:640-832) ----------------------------------------------------
    def update(self):
        n = len(self.data)
        n_ao = len(self.data[0]['coeffs']) if n else 0
        self.coeffs = numpy.zeros((n, n_ao), dtype=numpy.float64)
        self.occ = numpy.zeros(n)
        self.eig = numpy.zeros(n)
        sym, spin = [], []
        for i, mo in enumerate(self.data):
            self.coeffs[i] = mo['coeffs']
            self.occ[i] = mo['occ_num']
            self.eig[i] = mo.get('energy', 0.0)
            s = str(mo.get('sym', '%d.1' % (i + 1)))
            sp = str(mo.get('spin', 'unknown'))
            if s.endswith('_a') or s.endswith('_b'):
                sp = 'alpha' if s.endswith('_a') else 'beta'
                s = s[:-2]
            sym.append(s)
            spin.append(sp)
        self.sym = numpy.array(sym, dtype=str)
        self.spin = numpy.array(spin, dtype=str)
        self.alpha_index = [i for i, s in enumerate(spin) if not s.startswith('b')]
        self.beta_index = [i for i, s in enumerate(spin) if s.startswith('b')]
        self.spinpolarized = len(self.beta_index) != 0
        self._up_to_date = True
        self._stamp = next(_STAMPS)       # identifies this state of the flat arrays (engine handle caches)

    def _get(self, name):
        if not self._up_to_date:
            self.update()
        return copy(getattr(self, name))

    def get_coeffs(self): return self._get('coeffs')
    def get_eig(self): return self._get('eig')
    def get_sym(self): return self._get('sym')

    def get_occ(self, return_alpha_beta=False, return_int=False, tol_int=1e-5, sum_occ=False):
        occ = self._get('occ')
        if sum_occ:
            return sum(numpy.array(occ, dtype=numpy.intc))
        if return_alpha_beta and self.spinpolarized:
            occ = numpy.array([occ[self.alpha_index], occ[self.beta_index]])
        return numpy.array(occ, dtype=numpy.intc) if return_int else occ

    def set_coeffs(self, item):
        item = require(item, dtype=numpy.float64)
        if self.get_coeffs().shape != item.shape:
            raise ValueError('Old and new arrays need to be of the same size!')
        for i, mo in enumerate(self.data):
            mo['coeffs'] = item[i]
        self._up_to_date = False

    def set_occ(self, item):
        item = require(item, dtype=numpy.float64)
        if self.get_occ().shape != item.shape:
            raise ValueError('Old and new arrays need to be of the same size!')
        for i, mo in enumerate(self.data):
            mo['occ_num'] = item[i]
        self._up_to_date = False

    def get_labels(self, format='default'):
        fmt = {'short': '%(sym)s',
               'print': '%(sym)s (Occ = %(occ_num).2f, E = %(energy)+.4f E_h)',
               'cube': '%(sym)s,Occ=%(occ_num).1f,E=%(energy)+.2f'}
        fmt['cb'] = fmt['vmd'] = fmt['cube']
        f = fmt.get(format, '%(sym)s, Occ=%(occ_num).2f, E=%(energy)+.4f E_h')
        return [f % {'sym': mo.get('sym', ''), 'occ_num': mo['occ_num'],
                     'energy': mo.get('energy', 0.0)} for mo in self.data]

    def get_indices(self):
        return self.selected_mo if self.selected_mo is not None else list(range(len(self.data)))

    @property
    def is_energy_sorted(self):
        return bool(numpy.all(numpy.diff(self.get_eig()) >= 0))

    def sort_by_energy(self):
        """orbitals.py:649-659: the MO list itself is reordered (stable argsort of the energies), with a warning"""
        if self.is_energy_sorted:
            return
        import warnings
        warnings.warn('MOs are not sorted by energy. Sorting them...', UserWarning)
        order = numpy.argsort(self.get_eig())
        self.data = [self.data[i] for i in order]
        self._up_to_date = False
        self.update()

    def get_homo(self, tol=1e-5, sort=True):
        """index of the highest occupied MO; like the reference (orbitals.py:566-576) `sort=True` first sorts the MO
        list by energy IN PLACE -- for an unrestricted set read with all_mo=True (alpha block, then beta block) homo and
        lumo would otherwise straddle the two blocks"""
        if sort:
            self.sort_by_energy()
        occ = numpy.nonzero(self.get_occ() > tol)[0]
        return None if not len(occ) else occ[-1]

    def get_lumo(self, tol=1e-5, sort=True):
        if sort:
            self.sort_by_energy()
        un = numpy.nonzero(self.get_occ() < tol)[0]
        return None if not len(un) else un[0]

    # -- reduced MO selection ----------------------------------------------------------------
    def _resolve(self, token):
        token = token.strip().lower()
        m = re.fullmatch(r'(homo|lumo)\s*([+-]\s*\d+)?', token)
        if m:
            base = self.get_homo() if m.group(1) == 'homo' else self.get_lumo()
            if base is None:
                raise ValueError('no %s in this MO set' % m.group(1))
            return int(base) + (int(m.group(2).replace(' ', '')) if m.group(2) else 0)
        if re.fullmatch(r'[+-]?\d+', token):
            return int(token)
        if re.fullmatch(r'\d+\.\w+', token):             # MOLPRO-like label '3.1'
            hits = numpy.nonzero(self.get_sym() == token)[0]
            if len(hits):
                return int(hits[0])
        raise ValueError('MO selection %r not understood' % token)

    def select(self, fid_mo_list, flatten_input=True, sort_indices=True):
        """Subset of orbitals: 'all_mo', index lists, 'homo-1:lumo+2', 'homo', '3.1', ...

        Reduced form of orbitals.py:896-1210 (reading selections from files and the alpha/beta
        syntax are host conveniences outside the hot path).  Ranges follow the reference's
        convention 'a:b' == range(a, b) with optional step 'a:b:s'."""
        if isinstance(fid_mo_list, str):
            if fid_mo_list.lower() == 'all_mo':
                out = MOClass(list(self.data))
                out.selected_mo = list(range(len(self.data)))
                out.update()
                return out
            tokens = re.split(r'[,\s]+', fid_mo_list.strip())
        else:
            tokens = []
            for t in fid_mo_list:
                tokens.extend(t if isinstance(t, (list, tuple)) else [t])
        idx = []
        for tok in tokens:
            if isinstance(tok, (int, numpy.integer)):
                idx.append(int(tok))
            elif ':' in tok:
                parts = tok.split(':')
                a, b = self._resolve(parts[0]), self._resolve(parts[1])
                s = int(parts[2]) if len(parts) > 2 else 1
                idx.extend(range(a, b, s))
            else:
                idx.append(self._resolve(tok))
        if sort_indices:
            idx = sorted(set(idx))
        out = MOClass([self.data[i] for i in idx])
        out.selected_mo = idx
        out.selection_string = str(fid_mo_list)
        out.update()
        return out
