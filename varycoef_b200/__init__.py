"""varycoef_b200 -- B200-native (sm_100a) maximum-likelihood objective of GP-based SVC models.

Drop-in for the hot path of jakobdambon/varycoef (n2LL -> Sigma_y -> chol -> GLS_chol).
The arithmetic lives in libvarycoef_b200.so (hand-written CUDA behind a C ABI, see
include/varycoef_b200.h); this package is the thin host-side mirror of the reference's
R interface.  There is no CPU path.
"""
from .objective import SVCObjective, n2LL, GLS_chol, pc_penalty, check_cov_lower  # noqa: F401

from .svc_selection import SVC_selection, SVC_selection_control  # noqa: F401,E402

__all__ = ["SVCObjective", "n2LL", "GLS_chol", "pc_penalty", "check_cov_lower", "SVC_selection",
           "SVC_selection_control"]
