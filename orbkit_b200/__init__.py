"""orbkit_b200 -- ORBKIT's grid-based hot path (AO -> MO -> rho, grad rho, second derivatives of rho)
on NVIDIA B200 (sm_100a).

Drop-in for the path only: the same importable names and call signatures as the reference's
`orbkit.grid`, `orbkit.QCinfo`, `orbkit.core.{ao_creator, mo_creator, rho_compute,
rho_compute_no_slice}`, `orbkit.extras.{calc_ao, calc_mo, mo_set}` and the compiled module
`orbkit.cy_core`, returning NumPy arrays of identical shapes.  All arithmetic runs in hand-written
CUDA kernels behind the C ABI of include/okb200.h; there is no CPU fallback.
"""
from . import options, grid, tools, cy_grid
from .qcinfo import QCinfo
from .orbitals import AOClass, MOClass
from . import cy_core, core, extras, detci, output, read
from .core import ao_creator, mo_creator, rho_compute, rho_compute_no_slice
from .extras import calc_ao, calc_mo, mo_set
from .output import main_output
from .read import main_read

__version__ = '0.1.0'
__all__ = ['options', 'grid', 'tools', 'cy_grid', 'cy_core', 'core', 'extras', 'detci', 'output', 'main_output', 'read', 'main_read', 'QCinfo', 'AOClass',
           'MOClass', 'ao_creator', 'mo_creator', 'rho_compute', 'rho_compute_no_slice', 'calc_ao',
           'calc_mo', 'mo_set']
