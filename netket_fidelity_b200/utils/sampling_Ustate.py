"""make_logpsi_U_afun -- counterpart of netket_fidelity/utils/sampling_Ustate.py:7-47.

In the reference the wrapper turns an apply-fun into `log(U phi)(x) = logsumexp_c(log phi(x'_c), b=mels_c)`
and carries U inside `variables["unitary"]` so that changing its angle does not recompile.  Here the
same contract is met by the kernels themselves: an MCState whose variables hold `"unitary": U` evaluates,
samples and differentiates log(U phi) with the connected elements generated in-kernel; U is passed by
value on every launch, so changing it costs nothing."""


def make_logpsi_U_afun(state, U, variables=None):
    """Returns (logpsiU_fun, new_variables): `logpsiU_fun(variables, x)` evaluates log(U phi)(x) on the GPU
    with the parameters in `variables["params"]` and the operator in `variables["unitary"]`."""
    variables = dict(state.variables if variables is None else variables)
    new_variables = {**variables, "unitary": U}

    def logpsiU_fun(vars_, x):
        from ..nk.vqs import MCState
        tmp = MCState(state.sampler, state.model, n_samples=state.n_samples, variables=vars_, device=state.device)
        return tmp.log_value(x)

    return logpsiU_fun, new_variables
