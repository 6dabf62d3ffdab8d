"""detCI@ORBKIT grid contractions on the device (reference: orbkit/detci/ci_core.py:85-267,
cy_ci.pyx:70-240).  Only the grid-based part of detCI is here: `ci_core.rho`, `ci_core.jab`,
`ci_core.a_nabla_b` with the reference's signatures, plus fused `*_from_qc` variants that never move
the MO arrays over PCIe, and `cy_ci.get_rho_full / get_j_full / get_jab_full`, the time-dependent contractions of the
reference's compiled module.  CI-vector readers, occupation-pattern comparison (`occ_check.compare`) and
the analytic-integral expectation values stay with the caller (out of scope)."""
from . import ci_core, cy_ci

__all__ = ['ci_core', 'cy_ci']
