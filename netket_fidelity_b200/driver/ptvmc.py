"""PTVMC -- the outer time loop of projected time-dependent VMC; drop-in for the class of the same name in
netket_fidelity/driver/ptvmc.py:11-60 (same constructor arguments, same `run()`, same banner on stdout).

One time step = one infidelity optimisation of |psi> against U|phi> (`InfidelityOptimizer.run`, i.e. `n_iter`
infidelity-gradient steps on the GPU), after which the optimised parameters are handed to the target state
(reference ptvmc.py:54) -- a device-to-device copy of the flat parameter buffer here, nothing goes through the host.
"""

from __future__ import annotations

import numpy as np

from .infidelity_optimizer import InfidelityOptimizer, identity_preconditioner

_RULE = "#" * 42


class PTVMC:
    def __init__(self, target_state, U, vstate, optimizer, tf, dt, n_iter, obs=None, U_dagger=None,
                 preconditioner=identity_preconditioner, is_unitary=False, sample_Upsi=False, cv_coeff=None,
                 verbose=True):
        # the driver of ONE time step; it owns the infidelity operator whose target is re-pointed after every step
        self._te = InfidelityOptimizer(target_state, optimizer, variational_state=vstate, U=U, U_dagger=U_dagger,
                                       preconditioner=preconditioner, is_unitary=is_unitary, sample_Upsi=sample_Upsi,
                                       cv_coeff=cv_coeff)
        self._tf, self._dt, self._n_iter = tf, dt, n_iter
        self._ts = np.arange(0, tf, dt)          # t = 0, dt, 2 dt, ... < tf (reference ptvmc.py:33)
        self._obs, self._obs_dict = obs, {"obs": []}
        self._verbose = verbose

    @property
    def times(self):
        return self._ts

    @property
    def observables(self):
        """{"obs": [Stats per time step]} -- the dictionary the reference fills (empty list without `obs`)."""
        return self._obs_dict

    def _say(self, *lines):
        if self._verbose:
            for line in lines:
                print(line)

    def advance(self):
        """One time step: optimise, hand the evolved parameters to the target, measure."""
        te = self._te
        te.run(self._n_iter, show_progress=False)
        te._I_op.target.parameters = te.state.parameters
        if self._obs is not None:
            self._obs_dict["obs"].append(te.state.expect(self._obs))

    def run(self):
        for t in self._ts:
            self._say(f"Time t = {t}: ", _RULE)
            self.advance()
            self._say(_RULE, "\n")
        return self._obs_dict
