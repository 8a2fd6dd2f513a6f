from .infidelity_optimizer import InfidelityOptimizer  # noqa: F401
from .ptvmc import PTVMC  # noqa: F401
