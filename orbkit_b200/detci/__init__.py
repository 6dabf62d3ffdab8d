"""detCI@ORBKIT grid contractions on the device (reference: orbkit/detci/ci_core.py:85-267,
cy_ci.pyx:70-240).  Only the grid-based part of detCI is here: `ci_core.rho`, `ci_core.jab`,
`ci_core.a_nabla_b` with the reference's signatures, plus fused `*_from_qc` variants that never move
the MO arrays over PCIe.  CI-vector readers, occupation-pattern comparison (`occ_check.compare`) and
the analytic-integral expectation values stay with the caller (out of scope)."""
from . import ci_core

__all__ = ['ci_core']
