"""CG / Lanczos / pivoted-Cholesky solver restated from GPyTorch's behaviour."""
from .linear_cg import NumericalWarning, linear_cg  # noqa: F401
from .preconditioner import PivCholPreconditioner, pivoted_cholesky, slq_logdet  # noqa: F401
