from .operator import InfidelityOperatorStandard, InfidelityUPsi  # noqa: F401
