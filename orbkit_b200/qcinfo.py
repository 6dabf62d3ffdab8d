"""QCinfo: geometry + AO + MO container handed to the grid path (input contract).

Mirror of orbkit/qcinfo.py:35-110.  `geo_spec` is (n_atoms,3) in Bohr; `geo_info` rows are
[symbol, index, charge].  The file readers that normally fill this object (orbkit/read/*) are
out of scope; tests feed it from flat fixture arrays via `QCinfo.from_arrays`.
"""
from copy import deepcopy

import numpy

from .orbitals import AOClass, MOClass


class QCinfo:
    def __init__(self, data=None):
        self.geo_info = []
        self.geo_spec = []
        self.ao_spec = AOClass()
        self.mo_spec = MOClass()
        if data:
            self.geo_spec = numpy.array(data['geo_spec'], dtype=float)
            self.geo_info = numpy.array(data['geo_info'])
            ao, mo = data['ao_spec'], data['mo_spec']
            if isinstance(ao, numpy.ndarray):
                ao = ao[numpy.newaxis][0]
            if isinstance(mo, numpy.ndarray):
                mo = mo[numpy.newaxis][0]
            self.ao_spec = ao if isinstance(ao, AOClass) else AOClass(restart=ao)
            self.mo_spec = mo if isinstance(mo, MOClass) else MOClass(restart=mo)

    @classmethod
    def from_arrays(cls, arr):
        """Rebuild from the flat 'ao.*' / 'mo.*' arrays written by tests/golden/make_golden.py."""
        ao = {k[3:]: arr[k] for k in arr.keys() if k.startswith('ao.')}
        ao['spherical'] = bool(ao['spherical'])
        ao['normalized'] = bool(ao['normalized'])
        mo = {k[3:]: arr[k] for k in arr.keys() if k.startswith('mo.')}
        return cls({'geo_spec': arr['geo_spec'], 'geo_info': arr['geo_info'],
                    'ao_spec': ao, 'mo_spec': mo})

    def update(self):
        self.ao_spec.update()
        self.mo_spec.update()

    def copy(self):
        qc = deepcopy(self)
        qc.update()
        return qc

    def __eq__(self, other):
        if not isinstance(other, QCinfo):
            raise TypeError('Comparing of QCinfo to non QCinfo object not defined')
        return (numpy.allclose(self.geo_spec, other.geo_spec) and
                self.ao_spec == other.ao_spec and self.mo_spec == other.mo_spec)

    def todict(self):
        return {'geo_spec': self.geo_spec, 'geo_info': self.geo_info,
                'ao_spec': self.ao_spec.todict(), 'mo_spec': self.mo_spec.todict(),
                'parent_class_name': self.__module__ + '.' + self.__class__.__name__}

    def get_charge(self, nuclear=True, electron=True):
        charge = 0.
        if electron:
            charge -= float(numpy.sum(self.mo_spec.get_occ()))
        if nuclear:
            charge += sum(float(a[2]) for a in self.geo_info)
        return charge
