from .kernels import *  # noqa: F401,F403
from .models import *  # noqa: F401,F403
