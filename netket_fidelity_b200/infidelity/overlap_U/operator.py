"""InfidelityOperatorUPsi -- drop-in for netket_fidelity/infidelity/overlap_U/operator.py:12-67."""

from __future__ import annotations

from typing import Optional

from ...nk.vqs import MCState, FullSumState
from ..overlap.operator import _check_cv


class InfidelityOperatorUPsi:
    def __init__(self, U, state, *, cv_coeff: Optional[float] = None, U_dagger, is_unitary: bool = False, dtype=None):
        if not isinstance(state, (MCState, FullSumState)):
            raise TypeError("The first argument should be a variational state.")
        if not is_unitary and not isinstance(state, FullSumState):
            raise ValueError("Only works with unitary gates. If the gate is non unitary "
                             "then you must sample from it. Use a different operator.")
        cv_coeff = _check_cv(cv_coeff)
        if isinstance(state, FullSumState):
            cv_coeff = None  # overlap_U/operator.py:40-41
        self._hilbert = state.hilbert
        self._target = state
        self._cv_coeff = cv_coeff
        self._dtype = dtype
        self._U = U
        self._U_dagger = U_dagger

    hilbert = property(lambda self: self._hilbert)
    target = property(lambda self: self._target)
    cv_coeff = property(lambda self: self._cv_coeff)
    dtype = property(lambda self: self._dtype)
    is_hermitian = property(lambda self: True)

    def __repr__(self):
        return f"InfidelityOperatorUPsi(target=U@{self.target}, U={self._U}, cv_coeff={self.cv_coeff})"
