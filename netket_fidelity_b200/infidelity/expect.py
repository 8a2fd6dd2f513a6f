"""Host side of `MCState.expect` / `MCState.expect_and_grad` for the infidelity operators --
replaces the plum-dispatched overloads of infidelity/overlap/expect.py:14-55 and
infidelity/overlap_U/expect.py:16-67, and the jitted `infidelity_sampling_MCState` bodies
(:58-130 / :70-161) with two CUDA kernels (nkfb_infidelity_fwd, nkfb_infidelity_grad) and,
when torch.distributed is initialised, two all-reduces (8 doubles, then the 2P-double
gradient accumulator) standing in for NetKet's MPI means (expect.py:126 / :157).
"""

from __future__ import annotations

import torch

from .. import _native as nv
from ..nk.stats import Stats
from ..nk.vqs import MCState, FullSumState
from .overlap.operator import InfidelityOperatorStandard
from .overlap_U.operator import InfidelityOperatorUPsi


def _allreduce(t):
    from ..parallel import allreduce_sum_
    allreduce_sum_(t)


_side_streams = {}


def _side_stream(dev):
    key = dev.index
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=dev)
    return _side_streams[key]


def infidelity_expect(vstate: MCState, op, return_grad: bool):
    if op.hilbert != vstate.hilbert:
        raise TypeError("Hilbert spaces should match")
    target = op.target
    if not isinstance(target, MCState):
        raise TypeError("The target of an infidelity operator evaluated on an MCState must be an MCState")
    dev = vstate.device
    if isinstance(op, InfidelityOperatorStandard):
        # overlap/expect.py:85: log Phi(sigma) + log psi(eta) - log psi(sigma) - log Phi(eta)
        op1 = op4 = target._op_spec()
        op2 = None
    else:
        # overlap_U/expect.py:108-116
        if target._unitary is not None:
            raise TypeError("InfidelityOperatorUPsi needs a plain target state")
        for o in (op._U, op._U_dagger):
            if not hasattr(o, "_spec"):
                raise TypeError("U and U_dagger must be operators from netket_fidelity_b200.operator "
                                "(their connected elements are generated inside the CUDA kernels)")
        op1, op2, op4 = op._U._spec(dev), op._U_dagger._spec(dev), None
    if vstate._spec().code != target._spec().code:
        raise TypeError("the variational state and the target must share a parameter precision")
    if vstate._samples_packed is None and target._samples_packed is None and target is not vstate:
        # both chain families need fresh samples (the driver's reset(), driver/infidelity_optimizer.py:142-143):
        # run the two sampler launches on two streams so that the tail of one overlaps the other
        main = torch.cuda.current_stream(dev)
        side = _side_stream(dev)
        side.wait_stream(main)
        vstate._h.set_sampler_share(2)      # two launches of equal size side by side (include/nkfb200.h)
        try:
            vstate.sample()
            with torch.cuda.stream(side):
                target.sample()
        finally:
            vstate._h.set_sampler_share(1)
        main.wait_stream(side)
    sigma = vstate.samples_packed()
    eta = target.samples_packed()
    if sigma.shape != eta.shape:
        raise ValueError(f"both states need the same number of samples per rank, got {tuple(sigma.shape[:2])} "
                         f"and {tuple(eta.shape[:2])}")
    h = vstate._h
    W32 = sigma.shape[-1]
    psi, phi = vstate._spec(), target._spec()
    ls = vstate._ls
    acc = torch.zeros((8 + 2 * vstate._flat.numel(),), dtype=torch.float64, device=dev)
    sums, gacc = acc[:8], acc[8:]
    log_val, L, rho1, _ = h.infidelity_fwd(psi, phi, sigma.view(-1, W32), eta.view(-1, W32), ls, op.cv_coeff,
                                           op1, op2, op4, sums=sums)
    _allreduce(sums)
    # n_chains = sigma_t.shape[-2] (overlap/expect.py:72): the reference hands the CHAIN LENGTH to
    # nkjax.expect as "n_chains"; mean, variance and gradient do not depend on it
    # the statistics kernels go into the stream here; their read-back (a host synchronisation) waits until the
    # gradient has been enqueued behind them, so that the GPU does not idle while the host does the statistics
    stats_ctx = vstate._stats_begin(L, eta.shape[-2], sums=sums)
    flat = None
    if return_grad:
        h.infidelity_grad(psi, sigma.view(-1, W32), eta.view(-1, W32), ls, op.cv_coeff, log_val, rho1, sums,
                          op2=op2, grad_acc=gacc)
        _allreduce(gacc)
        flat = h.grad_finalize(gacc, sums, vstate._flat.numel(), vstate._cdtype)
    F_stats = vstate._stats_end(stats_ctx, L)
    I_stats = F_stats.replace(mean=1.0 - F_stats.mean)
    if not return_grad:
        return I_stats
    vstate._last_flat_grad = flat
    return I_stats, vstate._tree(flat)


def infidelity_exact(vstate: FullSumState, op, return_grad: bool):
    """Exact infidelity of two full-summation states and its gradient -- replaces
    `infidelity_sampling_FullSumState` of infidelity/overlap/exact.py:53-78 (target |Phi> = target.to_array())
    and infidelity/overlap_U/exact.py:62-90 (|Phi> = U target.to_array(normalize=False)):

        F = |<psi|Phi>|^2 / (<psi|psi> <Phi|Phi>),      I = 1 - F,
        nkjax.vjp(conjugate=True):  grad F_k = dF/dRe theta_k + i dF/dIm theta_k = 2 dF/d conj(theta_k)
                                            = sum_x w_x conj O_k(x),
        w_x = 2 [ conj(a) conj(psi_x) Phi_x / (n m) - F |psi_x|^2 / n ],  a = <psi|Phi>, n = <psi|psi>, m = <Phi|Phi>.

    The 2^N amplitudes come from nkfb_logpsi (U Phi: the same kernel with the operator attached, which is what
    `U.to_sparse() @ phi` computes), the sum over x from nkfb_weighted_score."""
    if op.hilbert != vstate.hilbert:
        raise TypeError("Hilbert spaces should match")
    target = op.target
    if not isinstance(target, FullSumState):
        raise TypeError("Can only compute infidelity of exact states.")
    if vstate._logstate is not None:
        raise NotImplementedError("LogStateVector is supported as a target only")
    dev = vstate.device
    if isinstance(op, InfidelityOperatorStandard):
        lt = target.log_amplitudes()
    else:
        if not hasattr(op._U, "_spec"):
            raise TypeError("U must be an operator from netket_fidelity_b200.operator")
        if target._logstate is not None or target._unitary is not None:
            raise TypeError("InfidelityOperatorUPsi on exact states needs a plain RBM target")
        lt = vstate._h.logpsi(target._spec(), target.all_states_packed(), target._ls,
                              op._U._spec(dev)).to(torch.complex128)
    lp = vstate.log_amplitudes()
    psi = torch.exp(lp - lp.real.max())
    T = torch.exp(lt - lt.real.max())
    n = (psi.abs() ** 2).sum()
    m = (T.abs() ** 2).sum()
    a = (psi.conj() * T).sum()
    F = (a.abs() ** 2 / (n * m)).real
    I_stats = Stats(mean=float(1.0 - F.item()), error_of_mean=0.0, variance=0.0)
    if not return_grad:
        return I_stats
    w = 2.0 * (a.conj() * psi.conj() * T / (n * m) - F * psi.abs() ** 2 / n)
    acc = vstate._h.weighted_score(vstate._spec(), vstate.all_states_packed(), vstate._ls, w)
    gF = torch.view_as_complex(acc.view(-1, 2).contiguous())
    flat = (-gF).to(vstate._cdtype)
    vstate._last_flat_grad = flat
    return I_stats, vstate._tree(flat)


def diagonal_expect(vstate, op):
    """<c0 + sum_i c_i sigma^z_i>: sample mean over the packed samples (MCState) or the exact average over
    |psi|^2 (FullSumState).  sigma^z_i = (2 x_i - s0 - s1) / (s1 - s0) in terms of the local-state value x_i."""
    if op.hilbert != vstate.hilbert:
        raise TypeError("Hilbert spaces should match")
    dev = vstate.device
    s0, s1 = vstate._ls
    c = torch.tensor(op.coeffs, dtype=torch.float64, device=dev)
    if isinstance(vstate, FullSumState):
        x = torch.tensor(vstate.hilbert.all_states(), dtype=torch.float64, device=dev)
        vals = op.const + ((2.0 * x - s0 - s1) / (s1 - s0)) @ c
        p = vstate.to_array().abs() ** 2
        mean = float((p * vals).sum().item())
        var = float((p * (vals - mean) ** 2).sum().item())
        return Stats(mean=mean, error_of_mean=0.0, variance=var)
    pk = vstate.samples_packed()
    x = vstate._h.unpack(pk.view(-1, pk.shape[-1]), vstate._N, vstate._ls, torch.float64)
    vals = op.const + ((2.0 * x - s0 - s1) / (s1 - s0)) @ c
    sums = torch.zeros(8, dtype=torch.float64, device=dev)
    sums[0], sums[1], sums[4] = vals.sum(), (vals * vals).sum(), float(vals.numel())
    _allreduce(sums)
    return vstate._stats_of(vals.contiguous(), pk.shape[0], sums=sums)


def expect_dispatch(vstate, op, return_grad):
    from ..renyi2 import Renyi2EntanglementEntropy, renyi2_expect
    if isinstance(op, (InfidelityOperatorStandard, InfidelityOperatorUPsi)):
        if isinstance(vstate, FullSumState):
            return infidelity_exact(vstate, op, return_grad)
        return infidelity_expect(vstate, op, return_grad)
    from ..nk.operator.spin import DiagonalZ
    if isinstance(op, DiagonalZ):
        if return_grad:
            raise NotImplementedError("gradient of a diagonal observable")
        return diagonal_expect(vstate, op)
    if isinstance(op, Renyi2EntanglementEntropy):
        if return_grad:
            raise NotImplementedError("gradient of the Renyi-2 entropy")
        return renyi2_expect(vstate, op)
    raise TypeError(f"no GPU expectation-value kernel for operators of type {type(op).__name__}")
