"""`nk.operator.spin` -- the one NetKet observable the reference's examples monitor during p-tVMC runs:
sums of sigma^z (examples/adiabatic_sweeping/adiabatic_sweeping.py:41, onespin/onespin_rotation.py)."""
from . import spin  # noqa: F401
