"""User-facing wrappers over core.rho_compute with the reference's signatures
(orbkit/extras.py: calc_mo :40-103, mo_set :105-205, calc_ao :208-260, atom2index :262-304,
gross_atomic_density :306-385, calc_jmo :441-493).

File output: `otype` "cb" / "cube" (Gaussian cube files, formatted on the device: orbkit_b200.output) with the
reference's file naming; the other output types of orbkit/output/* are outside the hot path and raise
NotImplementedError.  With otype=None (the library use) the return values are those of the reference.
"""
import os

import numpy

from . import core, options
from .display import display


def _write(data, qc, ofid, otype, **kwargs):
    """main_output as the reference's extras functions call it (skipped with options.no_output)"""
    if otype is None or options.no_output:
        return []
    from .output import main_output
    return main_output(data, qc, outputname=ofid, otype=otype, **kwargs)


def calc_mo(qc, fid_mo_list, drv=None, otype=None, ofid=None, numproc=None, slice_length=None):
    """Selected molecular orbitals (or derivatives) on the grid: ((NMO,)+N) or ((NDRV,NMO)+N)."""
    mo_spec = qc.mo_spec[fid_mo_list] if not isinstance(fid_mo_list, str) or fid_mo_list != 'all_mo' \
        else qc.mo_spec.select('all_mo')
    qc_select = qc.copy()
    qc_select.mo_spec = mo_spec
    mo_list = core.rho_compute(qc_select, calc_mo=True, drv=drv,
                               slice_length=options.slice_length if slice_length is None else slice_length,
                               numproc=options.numproc if numproc is None else numproc)
    if otype is None:
        return mo_list
    if ofid is None:                                          # extras.py:87-93
        outputname, group = options.outputname.split('@') if '@' in options.outputname else (options.outputname, '')
        outputname, autootype = os.path.splitext(outputname)
        ofid = '%s_MO%s@%s' % (outputname, autootype, group)
    _write(mo_list, qc_select, ofid, otype, datalabels=qc_select.mo_spec.get_labels(),
           dataindices=qc_select.mo_spec.get_indices(), drv=drv)
    return mo_list


def mo_set(qc, fid_mo_list, drv=None, laplacian=None, otype=None, ofid=None, return_all=True,
           numproc=None, slice_length=None):
    """Density (and derivatives) of selected MO sets; rows: one rho per set, then the derivative
    rows of every set (extras.py:105-205)."""
    sets = fid_mo_list
    if isinstance(sets, str) or (len(sets) and not isinstance(sets[0], (list, tuple, numpy.ndarray))):
        sets = [sets]
    laplacian = bool(laplacian)
    datasets, delta, labels, delta_labels = [], [], [], []
    for sel in sets:
        qc_select = qc.copy()
        qc_select.mo_spec = qc.mo_spec.select(sel) if not isinstance(sel, str) or sel != 'all_mo' \
            else qc.mo_spec.select('all_mo')
        label = 'mo_set:' + (sel if isinstance(sel, str) else ','.join(str(i) for i in sel))
        display('\nStarting with the molecular orbital list \n\t%s' % label)
        data = core.rho_compute(qc_select, drv=drv, laplacian=laplacian,
                                slice_length=options.slice_length if slice_length is None else slice_length,
                                numproc=options.numproc if numproc is None else numproc)
        if drv is None and not laplacian:
            rho = data
        elif laplacian:
            rho, delta_rho, lap = data
            delta.extend(delta_rho)
            delta.append(lap)
            delta_labels.extend(['d^2/d%s^2 %s' % (i, label) for i in 'xyz'])
            delta_labels.append('laplacian_of_' + label)
        else:
            rho, delta_rho = data
            delta.extend(delta_rho)
            delta_labels.extend(['d/d%s %s' % (i, label) for i in drv])
        datasets.append(rho)
        labels.append(label)
    datasets = numpy.array(datasets)
    delta = numpy.array(delta) if delta else numpy.zeros((0,) + datasets.shape[1:])
    data = numpy.append(datasets, delta, axis=0)
    delta_labels.append('mo_set')
    _write(data, qc, options.outputname if ofid is None else ofid, otype, datalabels=labels + delta_labels, drv=None)
    return data


def calc_ao(qc, drv=None, otype=None, ofid=None, numproc=None, slice_length=None):
    """All atomic orbitals (or derivatives) on the grid: ((NAO,)+N) or ((NDRV,NAO)+N)."""
    ao_list = core.rho_compute(qc, calc_ao=True, drv=drv, slice_length=options.slice_length,
                               numproc=options.numproc)
    _write(ao_list, qc, '%s_AO' % options.outputname if ofid is None else ofid, otype,
           datalabels=qc.ao_spec.get_labels(), drv=drv)
    return ao_list


def atom2index(atom, geo_info=None):
    """atom numbers (counting from one) -> (atom list, indices into geo_info counting from zero)
    (extras.py:262-304)."""
    if not isinstance(atom, (list, numpy.ndarray)):
        atom = [atom]
    if geo_info is not None:
        col = numpy.array(geo_info)[:, 1]
        index = []
        for a in atom:
            i = numpy.argwhere(col.astype(int) == a)
            if len(i) != 1:
                raise ValueError('No or multiple occurence of the atom number %d in geo_info!' % a)
            index.append(int(i[0, 0]))
        index = numpy.array(index, dtype=int)
    else:
        try:
            index = numpy.array(atom, dtype=int)
        except ValueError:
            raise ValueError('Cannot convert atom to integer array!')
    return atom, index


def _ao_atoms(ao_spec):
    """atom index (from zero) of every contracted AO, Cartesian or spherical"""
    cont_atom = numpy.asarray(ao_spec.get_assign_cont_to_atoms(), dtype=int)
    if ao_spec.spherical:
        return cont_atom[numpy.asarray(ao_spec.get_assign_lm_to_cont(), dtype=int)]
    return numpy.repeat(cont_atom, numpy.asarray(ao_spec.get_nlxlylz_per_cont(), dtype=int))


def gross_atomic_density(atom, qc, bReturnmo=False, ao_list=None, mo_list=None, drv=None):
    r"""Gross atomic density  rho^a = sum_i occ_i phi_i^a phi_i  with  phi_i^a = sum_{k on atom a} C_ik chi_k
    for the selected atoms (extras.py:306-385).  Returns a list with one array of the grid's shape per
    atom and, with `bReturnmo`, the gross atomic MOs mo_atom[a][i] as well.

    Mechanism: phi_i^a is an MO with the coefficients of all other atoms zeroed, so [C ; C^a] is
    evaluated as one set of 2 NMO orbitals by the fused AO->MO kernel and contracted on the device by the
    pair kernel with the terms (occ_i, i, NMO+i) -- the reference's expression order, MO values never
    cross PCIe.  Given `ao_list` / `mo_list` (host arrays) are honoured like in the reference; with `drv`
    both factors are the derivative, as the reference computes them (extras.py:350-353).

    Divergence: the reference counts AOs with the CARTESIAN degeneracy of every shell (extras.py:365), which
    mis-assigns the AOs of spherical bases; here every AO belongs to the atom of its shell."""
    from . import cy_core, grid
    from ._lib import OKB_CI_RHO
    from .detci import ci_core
    from .engine import get_engine
    from .tools import require, validate_drv
    if isinstance(atom, str) and atom == 'all' or (not isinstance(atom, (list, numpy.ndarray, str)) and atom == -1):
        atom = list(range(1, len(qc.geo_info) + 1))
    atom, index = atom2index(atom, geo_info=qc.geo_info)
    display('Computing the gross atomic density with respect to the atom(s) (internal numbering)')
    display('\t%s\n' % (list(atom),))
    coeffs = require(qc.mo_spec.get_coeffs(), dtype='f')
    occ = require(qc.mo_spec.get_occ(), dtype='f')
    n_mo, n_ao = coeffs.shape
    ao_atom = _ao_atoms(qc.ao_spec)
    if len(ao_atom) != n_ao:
        raise ValueError('AO/atom assignment has %d entries for %d AOs' % (len(ao_atom), n_ao))
    eng = get_engine()
    terms = (occ, numpy.arange(n_mo, dtype=numpy.intc), numpy.arange(n_mo, 2 * n_mo, dtype=numpy.intc))
    code = validate_drv(drv)
    rho_atom, mo_atom = [], []
    for a in index:
        c_a = numpy.where(ao_atom[None, :] == a, coeffs, 0.0)
        if ao_list is None and mo_list is None and code == 0 and not bReturnmo:
            # fused: [C ; C^a] on the module grid, contracted on the device
            x, y, z, is_vector, N = core._resolve_grid(None, None, None, None)
            basis = eng.basis(require(qc.geo_spec, dtype='f'), qc.ao_spec)
            if int(numpy.prod(N)) == 0:
                rho_atom.append(numpy.zeros(N))
                continue
            stacked = eng.mos(basis, numpy.concatenate([coeffs, c_a]), numpy.concatenate([occ, occ]))
            g = core._grid_handle(eng, x, y, z, is_vector)
            rho_atom.append(eng.eval_ci(OKB_CI_RHO, terms, stacked, g).reshape(N))
            continue
        # components given by the caller, a derivative, or the gross atomic MOs requested: two steps
        if ao_list is None:
            ao = core.ao_creator(qc.geo_spec, qc.ao_spec, drv=drv)
        else:
            ao = require(ao_list, dtype='f')
        shape = ao.shape[1:]
        ao2 = ao.reshape((ao.shape[0], -1))
        mo = cy_core.mocreator(ao2, coeffs) if mo_list is None else require(mo_list, dtype='f').reshape((n_mo, -1))
        mo_a = cy_core.mocreator(ao2, c_a)
        both = numpy.concatenate([mo, mo_a])
        rho_atom.append(eng.ci_contract(OKB_CI_RHO, terms, both).reshape(shape))
        if bReturnmo:
            mo_atom.append([m.reshape(shape) for m in mo_a])
    if bReturnmo:
        display('Returning the gross atomic density and\n\tthe gross atomic molecular orbitals')
        return rho_atom, mo_atom
    display('Returning the gross atomic density')
    return rho_atom


def calc_jmo(qc, ij, drv=['x', 'y', 'z'], numproc=1, otype=None, ofid='', **kwargs):
    """Transition electronic flux density between the molecular orbitals i and j for every pair in `ij`:
    jmo[d, n] = -1/2 (mo_i d_d mo_j - mo_j d_d mo_i), shape ((len(drv), len(ij)) + N) (extras.py:441-493).

    The reference forms the full NMO^2 product matrix of the selected orbitals (core.calc_mo_matrix) and then
    picks the pairs; here the selected MOs and their derivative sets are evaluated slab by slab on the device and
    only the requested pairs are formed there (okb_eval_ci, OKB_CI_JAB_PAIRS), with the reference's expression
    order and no FMA contraction."""
    from ._lib import OKB_CI_JAB_PAIRS, OKB_CI_PAIRS
    from .engine import get_engine
    from .tools import require, validate_drv
    ij = numpy.asarray(ij)
    if ij.ndim == 1 and len(ij) == 2:
        ij = ij.reshape((1, 2))
    assert ij.ndim == 2
    assert ij.shape[1] == 2
    u, indices = numpy.unique(ij, return_inverse=True)
    indices = indices.reshape((-1, 2))
    qc_select = qc.copy()
    qc_select.mo_spec = qc.mo_spec[u]
    drv = list(drv) if isinstance(drv, (list, tuple)) else [drv]
    codes = [validate_drv(d) for d in drv]
    if any(c == 0 for c in codes):
        raise ValueError('`drv` must name derivatives, e.g. ["x","y","z"]')
    x, y, z, is_vector, N = core._resolve_grid(None, None, None, None)
    npts, n = int(numpy.prod(N)), len(indices)
    labels = qc_select.mo_spec.get_labels(format='short')
    datalabels = ['j( %s , %s )' % (labels[i], labels[j]) for i, j in indices]

    def finish(jmo):
        _write(jmo, qc, ofid, otype, datalabels=datalabels, drv=drv)
        return jmo
    if npts == 0 or n == 0:
        return numpy.zeros((len(codes), n) + N)
    eng = get_engine()
    basis = eng.basis(require(qc_select.geo_spec, dtype='f'), qc_select.ao_spec)
    mo = eng.mos_of(basis, qc_select.mo_spec)
    g = core._grid_handle(eng, x, y, z, is_vector)
    ia = numpy.ascontiguousarray(indices[:, 0], dtype=numpy.intc)
    ib = numpy.ascontiguousarray(indices[:, 1], dtype=numpy.intc)
    if len(codes) == 3 and len(set(codes)) == 3:
        out = eng.eval_ci(OKB_CI_JAB_PAIRS, (numpy.zeros(n), ia, ib), mo, g, drv_codes=codes, flags=core._flags())
        return finish(out.reshape((3, n) + N))
    # any other number of components: the two products of every pair per component, combined on the host
    jmo = numpy.zeros((len(codes), n, npts))
    terms = (numpy.zeros(2 * n), numpy.concatenate([ia, ib]), numpy.concatenate([ib, ia]))
    for d, code in enumerate(codes):
        prod = eng.eval_ci(OKB_CI_PAIRS, terms, mo, g, drv_codes=[code], flags=core._flags())
        jmo[d] = - 0.5 * (prod[:n] - prod[n:])
    return finish(jmo.reshape((len(codes), n) + N))
