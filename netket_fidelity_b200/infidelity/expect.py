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
_stream_sets = {}


def _streamed_streams(dev):
    """(psi sampler, phi sampler, estimator) streams of the streamed path: the samplers above the estimator in priority, so
    that the estimator's CTAs only take the SMs the sampler launches leave idle."""
    key = dev.index
    if key not in _stream_sets:
        _stream_sets[key] = (torch.cuda.Stream(device=dev, priority=-1), torch.cuda.Stream(device=dev, priority=-1),
                             torch.cuda.Stream(device=dev, priority=0))
    return _stream_sets[key]


def _streamed_idle_sms(vstate, target):
    """SMs without a sampler CTA during the last wave of the two sampler launches, when pass A can be streamed under the
    sampling (same rule as engine.InfidelityStep._pick_segments); 0 otherwise."""
    import os
    if os.environ.get("NKFB_STREAM_SEGMENTS", "") == "1":
        return 0
    n = vstate._n_chains_local
    cl = vstate._chain_length
    if (vstate._cdtype != torch.complex64 or target._cdtype != torch.complex64 or not (161 <= vstate._M <= 512)
            or target._M != vstate._M or vstate._N > 128 or vstate.sampler.site_mode != "shared"
            or target.sampler.site_mode != "shared" or target._op_spec() is not None or vstate._op_spec() is not None
            or n < 1024 or target._n_chains_local != n or target._chain_length != cl or cl % 2 or n * (cl // 2) < 2048):
        return 0
    sm = torch.cuda.get_device_properties(vstate.device).multi_processor_count
    ctas = 2 * ((n + 31) // 32)
    waves = (ctas + sm - 1) // sm
    idle = waves * sm - ctas if waves <= 2 else 0
    return idle if idle >= 16 else 0


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
    h = vstate._h
    psi, phi = vstate._spec(), target._spec()
    ls = vstate._ls
    fresh = vstate._samples_packed is None and target._samples_packed is None and target is not vstate
    tables_flag = False
    idle = _streamed_idle_sms(vstate, target) if fresh else 0
    if idle:
        # A shard whose sampler launches leave SMs idle (one rank of a 4- / 8-GPU C4 job): sample in two segments and run
        # pass A of the first on the idle SMs while the second is being sampled (engine.py).  Pass A /
        # pass B work on segment-major lists ([segment][chain][sample]); the samples the states keep and the local
        # estimators handed to the chain statistics are put back into NetKet's (chain, sample) order.
        K = 2
        n, cl, W32 = vstate._n_chains_local, vstate._chain_length, vstate._W32
        seg = cl // K
        S_seg = n * seg
        main = torch.cuda.current_stream(dev)
        s_psi, s_phi, s_est = _streamed_streams(dev)
        h.set_sampler_share(2)
        try:
            for k in range(K):
                nd_psi = vstate.n_discard_per_chain if k == 0 else 0
                nd_phi = target.n_discard_per_chain if k == 0 else 0
                if k == 0:
                    # the first sampler launches go out before anything else is allocated: the GPU is idle until they start
                    s_psi.wait_stream(main)
                    sig_seg = torch.empty((K, n, seg, W32), dtype=torch.int32, device=dev)
                with torch.cuda.stream(s_psi):
                    vstate._sample_segment(seg, nd_psi, sig_seg[k], tables_valid=k > 0)
                if k == 0:
                    s_phi.wait_stream(main)
                    eta_seg = torch.empty_like(sig_seg)
                with torch.cuda.stream(s_phi):
                    target._sample_segment(seg, nd_phi, eta_seg[k], tables_valid=k > 0)
                if k == 0:
                    acc = torch.zeros((8 + 2 * vstate._flat.numel(),), dtype=torch.float64, device=dev)
                    sums, gacc = acc[:8], acc[8:]
                    log_val = torch.empty((K * S_seg,), dtype=vstate._cdtype, device=dev)
                    rho1 = torch.empty_like(log_val)
                    L_seg = torch.empty((K * S_seg,), dtype=torch.float32, device=dev)
                    s_est.wait_stream(main)
                    tables_flag = True
                s_est.wait_stream(s_psi)
                s_est.wait_stream(s_phi)
                with torch.cuda.stream(s_est):
                    h.set_fwd_max_ctas(2 * idle if k < K - 1 else 0)
                    # pass B's tables ahead of the first pass A; pass A of the last segment reuses the first one's
                    # (engine.py: the earlier pass A launches keep their preparation kernels on purpose)
                    h.set_tc_tables_valid(k == K - 1)
                    if k == 0:
                        h.tc_prepare(psi, None, ls)
                    sl = slice(k * S_seg, (k + 1) * S_seg)
                    h.infidelity_fwd(psi, phi, sig_seg[k].view(-1, W32), eta_seg[k].view(-1, W32), ls, op.cv_coeff, op1, op2,
                                     op4, sums=sums, out=(log_val[sl], L_seg[sl], rho1[sl]))
        except BaseException:
            h.set_tc_tables_valid(False)
            raise
        finally:
            h.set_sampler_share(1)
            h.set_fwd_max_ctas(0)
        main.wait_stream(s_est)       # s_est waited for both sampler streams: everything above is ordered before `main` now
        sigma, eta = sig_seg.view(-1, W32), eta_seg.view(-1, W32)          # segment-major, as log_val / rho1 are
        vstate._samples_packed = sig_seg.permute(1, 0, 2, 3).reshape(n, cl, W32)
        target._samples_packed = eta_seg.permute(1, 0, 2, 3).reshape(n, cl, W32)
        L = L_seg.view(K, n, seg).permute(1, 0, 2).reshape(-1)              # (chain, sample) order for the statistics
        chain_len = cl
    else:
        if fresh:
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
        acc = torch.zeros((8 + 2 * vstate._flat.numel(),), dtype=torch.float64, device=dev)
        sums, gacc = acc[:8], acc[8:]
        sigma = vstate.samples_packed()
        eta = target.samples_packed()
        if sigma.shape != eta.shape:
            raise ValueError(f"both states need the same number of samples per rank, got {tuple(sigma.shape[:2])} "
                             f"and {tuple(eta.shape[:2])}")
        W32 = sigma.shape[-1]
        chain_len = eta.shape[-2]
        sigma, eta = sigma.view(-1, W32), eta.view(-1, W32)
        log_val, L, rho1, _ = h.infidelity_fwd(psi, phi, sigma, eta, ls, op.cv_coeff, op1, op2, op4, sums=sums)
    flat = None
    try:
        _allreduce(sums)
        # n_chains = sigma_t.shape[-2] (overlap/expect.py:72): the reference hands the CHAIN LENGTH to
        # nkjax.expect as "n_chains"; mean, variance and gradient do not depend on it
        # the statistics kernels go into the stream here; their read-back (a host synchronisation) waits until the
        # gradient has been enqueued behind them, so that the GPU does not idle while the host does the statistics
        stats_ctx = vstate._stats_begin(L, chain_len, sums=sums)
        if return_grad:
            h.infidelity_grad(psi, sigma, eta, ls, op.cv_coeff, log_val, rho1, sums, op2=op2, grad_acc=gacc)
    finally:
        if tables_flag:
            h.set_tc_tables_valid(False)     # set by the streamed branch for the launches of this call only
    if return_grad:
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
