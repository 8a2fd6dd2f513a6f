from .logic import InfidelityOperator  # noqa: F401
from .overlap import InfidelityOperatorStandard, InfidelityUPsi  # noqa: F401
from .overlap_U import InfidelityOperatorUPsi  # noqa: F401
