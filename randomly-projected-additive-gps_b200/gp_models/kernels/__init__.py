from .etc import *  # noqa: F401,F403
from .etc import _sample_from_range  # noqa: F401
from .imq_kernel import *  # noqa: F401,F403
from .memory_efficient_gam_kernel import *  # noqa: F401,F403
from .polynomial_projection_kernels import *  # noqa: F401,F403
from .scaled_projection_kernel import *  # noqa: F401,F403
