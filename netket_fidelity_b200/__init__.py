"""netket_fidelity_b200 -- B200-native (sm_100a) hot path of netket_fidelity:
one infidelity-gradient step of p-tVMC, behind the reference's public surface.

There is no CPU fallback: every numerical entry point calls the hand-written CUDA
kernels in lib/libnkfb200.so through the C ABI of include/nkfb200.h.
"""

__version__ = "0.1.0"
