from .operator import InfidelityOperatorUPsi  # noqa: F401
