"""netket_fidelity_b200 -- B200-native (sm_100a) hot path of netket_fidelity:
one infidelity-gradient step of p-tVMC, behind the reference's public surface
(netket_fidelity/__init__.py:1-13): `InfidelityOperator`, `operator`, `driver`, `infidelity`,
`renyi2`, `utils`, plus `nk` -- the minimal NetKet stand-ins the surface is written against.

There is no CPU fallback: every numerical entry point calls the hand-written CUDA kernels in
lib/libnkfb200.so through the C ABI of include/nkfb200.h and raises if it is missing.
"""

__version__ = "0.1.0"

from . import nk  # noqa: F401,E402
from . import operator  # noqa: F401,E402
from . import infidelity  # noqa: F401,E402
from . import driver  # noqa: F401,E402
from . import utils  # noqa: F401,E402
from . import renyi2  # noqa: F401,E402
from .infidelity import InfidelityOperator  # noqa: F401,E402
from .renyi2 import Renyi2EntanglementEntropy  # noqa: F401,E402
