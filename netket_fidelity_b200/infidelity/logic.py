"""InfidelityOperator factory -- drop-in for netket_fidelity/infidelity/logic.py:11-187
(same keyword arguments, same branch order, same errors)."""

from __future__ import annotations

from typing import Optional

from ..nk.vqs import FullSumState
from .overlap import InfidelityOperatorStandard, InfidelityUPsi
from .overlap_U import InfidelityOperatorUPsi


class Adjoint:
    """Stand-in for netket.operator.Adjoint: what a non-netket_fidelity operator's `.H` returns
    (logic.py:147-152 refuses it)."""


def InfidelityOperator(target, *, U=None, U_dagger=None, cv_coeff: Optional[float] = None, is_unitary: bool = False,
                       dtype=None, sample_Upsi=False):
    r"""Operator I_op computing the infidelity between |psi> and |Phi> = |phi> or U|phi>:

        I = 1 - |<psi|Phi>|^2 / (<psi|psi><Phi|Phi>)

    * `InfidelityOperator(target)`                         -> overlap with |phi>
    * `InfidelityOperator(target, U=U, is_unitary=True)`   -> sample |psi>, |phi>; conns of U, U^dagger
    * `InfidelityOperator(target, U=U, sample_Upsi=True)`  -> sample |psi> and U|phi>
    """
    if U is None:
        return InfidelityOperatorStandard(target, cv_coeff=cv_coeff, dtype=dtype)
    if U_dagger is None:
        U_dagger = U.H
    if isinstance(U_dagger, Adjoint):
        raise TypeError("Must explicitly pass a jax-compatible operator as `U_dagger`. "
                        "You either did not pass `U_dagger` explicitly or you used `U.H` but should "
                        "use operators coming from `netket_fidelity`. ")
    if isinstance(target, FullSumState):
        return InfidelityOperatorUPsi(U, target, U_dagger=U_dagger, cv_coeff=cv_coeff, dtype=dtype)
    if not is_unitary and not sample_Upsi:
        raise ValueError("Non-unitary operators can only be handled by sampling from the state U|ψ⟩. "
                         "This is more expensive and disabled by default. "
                         "If your operator is Unitary, please specify so by passing `is_unitary=True` as a "
                         "keyword argument. "
                         "If your operator is not unitary, please specify `sample_Upsi=True` explicitly to"
                         "sample from that state. "
                         "You can also sample from U|ψ⟩ if your operator is unitary. ")
    if sample_Upsi:
        return InfidelityUPsi(U, target, cv_coeff=cv_coeff, dtype=dtype)
    return InfidelityOperatorUPsi(U, target, U_dagger=U_dagger, cv_coeff=cv_coeff, dtype=dtype, is_unitary=True)
