from . import sampling_Ustate, mpack  # noqa: F401
