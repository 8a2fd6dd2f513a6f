"""PTVMC -- drop-in for netket_fidelity/driver/ptvmc.py:11-60 (outer p-tVMC time loop)."""

from __future__ import annotations

import numpy as np

from .infidelity_optimizer import InfidelityOptimizer, identity_preconditioner


class PTVMC:
    def __init__(self, target_state, U, vstate, optimizer, tf, dt, n_iter, obs=None, U_dagger=None,
                 preconditioner=identity_preconditioner, is_unitary=False, sample_Upsi=False, cv_coeff=None,
                 verbose=True):
        self._dt, self._tf = dt, tf
        self._ts = np.arange(0, self._tf, self._dt)
        self._n_iter = n_iter
        self._obs = obs
        self._obs_dict = {"obs": []}
        self._verbose = verbose
        self._te = InfidelityOptimizer(target_state, optimizer, variational_state=vstate, U=U, U_dagger=U_dagger,
                                       preconditioner=preconditioner, is_unitary=is_unitary,
                                       sample_Upsi=sample_Upsi, cv_coeff=cv_coeff)

    def run(self):
        for t in self._ts:
            if self._verbose:
                print(f"Time t = {t}: ")
                print("##########################################")
            self._te.run(self._n_iter, show_progress=False)
            # ptvmc.py:54: the evolved state becomes the next target
            self._te._I_op.target.parameters = self._te.state.parameters
            if self._obs is not None:
                self._obs_dict["obs"].append(self._te.state.expect(self._obs))
            if self._verbose:
                print("##########################################")
                print("\n")
        return self._obs_dict
