"""netket_fidelity.operator (reference operator/__init__.py:1-9): Rx, Ry, Hadamard, Ising."""
from .singlequbit_gates import Rx, Ry, Hadamard, DiscreteJaxOperator  # noqa: F401
from .ising import Ising, IsingPropagator  # noqa: F401
