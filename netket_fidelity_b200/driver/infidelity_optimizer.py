"""InfidelityOptimizer -- drop-in for netket_fidelity/driver/infidelity_optimizer.py:31-252."""

from __future__ import annotations

from typing import Optional

from ..infidelity import InfidelityOperator
from ..nk.callbacks import ConvergenceStopping
from .abstract_variational_driver import AbstractVariationalDriver


def identity_preconditioner(vstate, grad, step=None):
    return grad


def _to_tuple(maybe_iterable):
    if hasattr(maybe_iterable, "__iter__"):
        return tuple(maybe_iterable)
    return (maybe_iterable,)


class InfidelityOptimizer(AbstractVariationalDriver):
    def __init__(self, target_state, optimizer, *, variational_state, U=None, U_dagger=None,
                 preconditioner=identity_preconditioner, is_unitary=False, sample_Upsi=False, cv_coeff=-0.5):
        r"""Driver training `variational_state` to match |Phi> = |phi> or U|phi> (target_state = |phi>)
        by minimising I = 1 - |<psi|Phi>|^2 / (<psi|psi><Phi|Phi>) with the Monte-Carlo estimators and the
        control-variates correction of the reference (driver/infidelity_optimizer.py:45-118).  Same
        keyword arguments and defaults (note cv_coeff defaults to -1/2 here, None in the factory)."""
        super().__init__(variational_state, optimizer, minimized_quantity_name="Infidelity")
        self._cv = cv_coeff
        self._preconditioner = preconditioner
        self._I_op = InfidelityOperator(target_state, U=U, U_dagger=U_dagger, is_unitary=is_unitary,
                                        cv_coeff=cv_coeff, sample_Upsi=sample_Upsi)

    def _forward_and_backward(self):
        # driver/infidelity_optimizer.py:141-150
        self.state.reset()
        self._I_op.target.reset()
        self._loss_stats, self._loss_grad = self.state.expect_and_grad(self._I_op)
        self._dp = self.preconditioner(self.state, self._loss_grad, self.step_count)
        if self._dp is self._loss_grad:
            return self.state._last_flat_grad          # identity preconditioner: keep the flat device vector
        return self.state._flatten_tree(self._dp)

    def run(self, n_iter, out=None, *args, target_infidelity=None, callback=lambda *x: True, show_progress=True,
            **kwargs):
        """driver/infidelity_optimizer.py:152-188: adds ConvergenceStopping(target, smoothing_window=20,
        patience=30) when `target_infidelity` is given."""
        callbacks = _to_tuple(callback)
        if target_infidelity is not None:
            callbacks = callbacks + (ConvergenceStopping(target_infidelity, smoothing_window=20, patience=30),)
        return super().run(n_iter, out, *args, show_progress=show_progress, callback=callbacks, **kwargs)

    @property
    def cv(self) -> Optional[float]:
        """The coefficient of the control variates."""
        return self._cv

    @property
    def infidelity(self):
        """MCMC statistics of the infidelity at the last step."""
        return self._loss_stats

    @property
    def preconditioner(self):
        return self._preconditioner

    @preconditioner.setter
    def preconditioner(self, val):
        self._preconditioner = identity_preconditioner if val is None else val

    def __repr__(self):
        return "InfidelityOptimiser(" + f"\n  step_count = {self.step_count}," + f"\n  state = {self.state})"

    def info(self, depth=0):
        from .infidelity_optimizer_common import info
        lines = ["{}: {}".format(name, info(obj, depth=depth + 1)) for name, obj in [
            ("Propagator    ", getattr(self._I_op, "_U", None)),
            ("Optimizer      ", self.optimizer),
            ("Preconditioner ", self.preconditioner),
            ("State          ", self.state)]]
        return "\n{}".format(" " * 3 * (depth + 1)).join([str(self)] + lines)
