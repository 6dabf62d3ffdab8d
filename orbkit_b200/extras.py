"""User-facing wrappers over core.rho_compute with the reference's signatures
(orbkit/extras.py: calc_mo :40-103, mo_set :105-205, calc_ao :208-260).

File output (`otype`, orbkit/output/*) is outside the hot path: requesting it raises
NotImplementedError; with otype=None (the library use) the return values are those of the reference.
"""
import numpy

from . import core, options
from .display import display


def _no_output(otype):
    if otype is not None:
        raise NotImplementedError('orbkit_b200 implements the grid-based compute path only; write the '
                                  'returned arrays with the reference\'s output module (otype=%r)' % (otype,))


def calc_mo(qc, fid_mo_list, drv=None, otype=None, ofid=None, numproc=None, slice_length=None):
    """Selected molecular orbitals (or derivatives) on the grid: ((NMO,)+N) or ((NDRV,NMO)+N)."""
    _no_output(otype)
    mo_spec = qc.mo_spec[fid_mo_list] if not isinstance(fid_mo_list, str) or fid_mo_list != 'all_mo' \
        else qc.mo_spec.select('all_mo')
    qc_select = qc.copy()
    qc_select.mo_spec = mo_spec
    return core.rho_compute(qc_select, calc_mo=True, drv=drv,
                            slice_length=options.slice_length if slice_length is None else slice_length,
                            numproc=options.numproc if numproc is None else numproc)


def mo_set(qc, fid_mo_list, drv=None, laplacian=None, otype=None, ofid=None, return_all=True,
           numproc=None, slice_length=None):
    """Density (and derivatives) of selected MO sets; rows: one rho per set, then the derivative
    rows of every set (extras.py:105-205)."""
    _no_output(otype)
    sets = fid_mo_list
    if isinstance(sets, str) or (len(sets) and not isinstance(sets[0], (list, tuple, numpy.ndarray))):
        sets = [sets]
    laplacian = bool(laplacian)
    datasets, delta = [], []
    for sel in sets:
        qc_select = qc.copy()
        qc_select.mo_spec = qc.mo_spec.select(sel) if not isinstance(sel, str) or sel != 'all_mo' \
            else qc.mo_spec.select('all_mo')
        display('\nStarting with the molecular orbital list \n\tmo_set:%s' % (sel,))
        data = core.rho_compute(qc_select, drv=drv, laplacian=laplacian,
                                slice_length=options.slice_length if slice_length is None else slice_length,
                                numproc=options.numproc if numproc is None else numproc)
        if drv is None and not laplacian:
            rho = data
        elif laplacian:
            rho, delta_rho, lap = data
            delta.extend(delta_rho)
            delta.append(lap)
        else:
            rho, delta_rho = data
            delta.extend(delta_rho)
        datasets.append(rho)
    datasets = numpy.array(datasets)
    delta = numpy.array(delta) if delta else numpy.zeros((0,) + datasets.shape[1:])
    return numpy.append(datasets, delta, axis=0)


def calc_ao(qc, drv=None, otype=None, ofid=None, numproc=None, slice_length=None):
    """All atomic orbitals (or derivatives) on the grid: ((NAO,)+N) or ((NDRV,NAO)+N)."""
    _no_output(otype)
    return core.rho_compute(qc, calc_ao=True, drv=drv, slice_length=options.slice_length,
                            numproc=options.numproc)
