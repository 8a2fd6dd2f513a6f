"""InfidelityOperatorStandard and InfidelityUPsi -- drop-in for
netket_fidelity/infidelity/overlap/operator.py:15-82 (same validation and error messages)."""

from __future__ import annotations

import numbers
from typing import Optional

import numpy as np

from ...nk.vqs import MCState, FullSumState
from ...operator.singlequbit_gates import DiscreteJaxOperator


def _check_cv(cv_coeff):
    """overlap/operator.py:28-32: a real scalar or None."""
    if cv_coeff is None:
        return None
    arr = np.asarray(cv_coeff)
    if arr.ndim != 0 or np.iscomplexobj(arr) or not isinstance(arr.item(), numbers.Real):
        raise TypeError("`cv_coeff` should be a real scalar number or None.")
    return float(arr)


class InfidelityOperatorStandard:
    def __init__(self, target, *, cv_coeff: Optional[float] = None, dtype=None):
        if not isinstance(target, (MCState, FullSumState)):
            raise TypeError("The first argument should be a variational target.")
        cv_coeff = _check_cv(cv_coeff)
        if isinstance(target, FullSumState):
            cv_coeff = None  # overlap/operator.py:34-35: no control variate on exact states
        self._hilbert = target.hilbert
        self._target = target
        self._cv_coeff = cv_coeff
        self._dtype = dtype

    hilbert = property(lambda self: self._hilbert)
    target = property(lambda self: self._target)
    cv_coeff = property(lambda self: self._cv_coeff)
    dtype = property(lambda self: self._dtype)
    is_hermitian = property(lambda self: True)

    def __repr__(self):
        return f"InfidelityOperator(target={self.target}, cv_coeff={self.cv_coeff})"


def InfidelityUPsi(U, state, *, cv_coeff: Optional[float] = None, dtype=None):
    """overlap/operator.py:61-82: a NEW MCState whose amplitude is log(U phi), sharing the target's
    sampler and n_samples; the unitary rides in `variables["unitary"]`."""
    if not isinstance(U, DiscreteJaxOperator):
        raise TypeError("In order to sample from the state U|psi>, U must be"
                        "an instance of DiscreteJaxOperator.")
    params = {k: ({kk: vv.clone() for kk, vv in v.items()} if isinstance(v, dict) else v.clone())
              for k, v in state.parameters.items()}
    target = MCState(sampler=state.sampler, model=state.model, n_samples=state.n_samples,
                     n_discard_per_chain=state.n_discard_per_chain,
                     variables={"params": params, "unitary": U}, device=state.device)
    return InfidelityOperatorStandard(target, cv_coeff=cv_coeff, dtype=dtype)
