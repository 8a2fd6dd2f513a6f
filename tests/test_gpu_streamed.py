"""Pass A streamed under the sampler (engine.InfidelityStep, StepConfig.stream_segments): the chains are sampled in K
consecutive segments (nkfb_sample with NKFB_SAMPLE_TABLES_VALID on the later ones) and pass A of segment k runs on a
low-priority stream, its CTAs limited by NKFB_OPT_FWD_MAX_CTAS, while segment k + 1 is sampled -- the form the 8-GPU
shard of C4 runs.  Held here: the streamed step is the unstreamed step (reference call order
driver/infidelity_optimizer.py:142-150: sample both states, then estimator and gradient) on the same chains."""
import numpy as np
import pytest
import torch

from netket_fidelity_b200 import _native as nv
from netket_fidelity_b200.engine import InfidelityStep, StepConfig
from netket_fidelity_b200.workloads import make_workload, device_ops, random_rbm

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _specs(N, M, cd, vb, std):
    return [nv.RbmSpec(*[t.to(DEV) if t is not None else None for t in random_rbm(N, M, s, cd, vb, std=std)])
            for s in (1234, 5678)]


def _chain_major(eng):
    """samples as (chains, chain_length, W32) whatever the engine's layout"""
    out = []
    for t in (eng.sigma, eng.eta):
        out.append(t.permute(1, 0, 2, 3).reshape(t.shape[1], -1, t.shape[3]) if t.dim() == 4 else t)
    return out


def test_streamed_step_is_the_unstreamed_step_float64():
    """complex128 (CUDA-core kernels, trajectories independent of how a run is split into calls): K = 2 and K = 4
    segments emit bit-identical samples over two consecutive steps; sums and gradient agree to fp64 rounding."""
    w = make_workload("C2")
    N, M = w["N"], w["M"]
    U, Ud = device_ops(w, torch.device(DEV))
    psi, phi = _specs(N, M, torch.complex128, w["vb"], 0.05)
    res = {}
    for K in (1, 2, 4):
        cfg = StepConfig(96, 8, 3, N, seed=5, cv_coeff=-0.5, stream_segments=K)
        eng = InfidelityStep(N, M, cfg, w["mode"], U, Ud, device=torch.device(DEV), cdtype=torch.complex128)
        assert eng.K == K
        steps = []
        for _ in range(2):
            sums, L, grad = eng.step(psi, phi)
            sig, eta = _chain_major(eng)
            steps.append((sums.cpu().numpy().copy(), grad.cpu().numpy().copy(), sig.cpu().numpy().copy(),
                          eta.cpu().numpy().copy()))
        res[K] = steps
    for K in (2, 4):
        for a, b in zip(res[1], res[K]):
            np.testing.assert_array_equal(a[2], b[2])
            np.testing.assert_array_equal(a[3], b[3])
            np.testing.assert_allclose(a[0], b[0], rtol=1e-12, atol=1e-12)
            assert np.max(np.abs(a[1] - b[1])) <= 1e-12 * max(1.0, np.max(np.abs(a[1])))


@pytest.mark.parametrize("n_chains", [2048, 4096])
def test_streamed_step_on_the_multi_gpu_shards_of_c4(n_chains):
    """complex64, N = 100, M = 400, 2048 / 4096 chains x 64 (one rank of an 8- / 4-GPU C4 job): the engine picks K = 2 by
    itself (the last wave of the two sampler launches leaves 20 / 40 SMs idle); tensor-core pass A per segment with the CTA
    limit, half-warp sampler with reused tables.  Against the unstreamed step from the same seed: the acceptance decisions differ only by single-precision
    rounding (z is refreshed at the start of every call), so nearly all chains emit identical samples, and F and the
    gradient agree within the Monte-Carlo error of the few chains that differ."""
    w = make_workload("C4")
    N, M = w["N"], w["M"]
    U, Ud = device_ops(w, torch.device(DEV))
    psi, phi = _specs(N, M, torch.complex64, w["vb"], 0.01)
    out = {}
    for K in (None, 1):
        cfg = StepConfig(n_chains, 64, 5, N, seed=20261019, cv_coeff=-0.5, stream_segments=K)
        eng = InfidelityStep(N, M, cfg, w["mode"], U, Ud, device=torch.device(DEV))
        assert eng.K == (2 if K is None else 1)
        eng.step(psi, phi)
        sums, L, grad = eng.step(psi, phi)          # second step: chains continued, tables rebuilt once per step
        torch.cuda.synchronize()
        sig, eta = _chain_major(eng)
        out[K] = (sums.cpu().numpy().copy(), grad.cpu().numpy().copy(), sig.cpu(), eta.cpu(), eng.n_accepted.cpu().numpy().copy())
    a, b = out[None], out[1]
    assert a[0][4] == b[0][4] == n_chains * 64
    same = ((a[2] == b[2]).flatten(1).all(dim=1) & (a[3] == b[3]).flatten(1).all(dim=1)).float().mean().item()
    assert same > 0.9, same
    Fa, Fb = a[0][0] / a[0][4], b[0][0] / b[0][4]
    var = b[0][1] / b[0][4] - Fb ** 2
    assert abs(Fa - Fb) < 5 * np.sqrt(8.0 * max(var, 1e-12) / b[0][4]) * np.sqrt(1.0 - same + 1e-3) + 1e-6, (Fa, Fb)
    assert np.max(np.abs(a[1] - b[1])) < 0.1 * np.max(np.abs(b[1])), np.max(np.abs(a[1] - b[1])) / np.max(np.abs(b[1]))
    assert abs(int(a[4].sum()) - int(b[4].sum())) < 1e-3 * b[4].sum()


def test_public_path_streams_pass_a_on_a_shard(monkeypatch):
    """`MCState.expect_and_grad(InfidelityOperator)` on the 8-GPU shard shape of C4 (2048 chains x 64 per family): the
    public path samples in two segments and runs pass A of the first under the second (infidelity/expect.py), exactly
    when the engine does.  Against the same call with NKFB_STREAM_SEGMENTS=1 (one sampling pass, one pass A): the states
    keep their samples in NetKet's (chain, sample) layout, nearly all chains emit identical samples, mean / variance /
    gradient agree within the Monte-Carlo error of the few that differ, and the chain statistics (error_of_mean,
    tau_corr, R_hat: functions of the per-chain order of the local estimators) agree."""
    import netket_fidelity_b200 as nkf
    from netket_fidelity_b200 import nk
    from netket_fidelity_b200.infidelity import expect as ex
    w = make_workload("C4")
    N, M = w["N"], w["M"]
    dev = torch.device(DEV)
    hi = nk.hilbert.Spin(0.5, N)
    out = {}
    for mode in ("streamed", "plain"):
        if mode == "plain":
            monkeypatch.setenv("NKFB_STREAM_SEGMENTS", "1")
        else:
            monkeypatch.delenv("NKFB_STREAM_SEGMENTS", raising=False)
        sa = nk.sampler.MetropolisLocal(hi, n_chains=2048, sweep_size=N)
        ma = nk.models.RBM(alpha=M // N, param_dtype="complex64", use_visible_bias=w["vb"])
        mk = lambda seed: {"params": (lambda t: {"Dense": {"kernel": t[0].to(dev), "bias": t[1].to(dev)},
                                                **({"visible_bias": t[2].to(dev)} if t[2] is not None else {})})(
            random_rbm(N, M, seed, torch.complex64, w["vb"]))}
        psi = nk.vqs.MCState(sa, ma, n_samples=2048 * 64, n_discard_per_chain=5, seed=1, device=dev, variables=mk(1234))
        phi = nk.vqs.MCState(sa, ma, n_samples=2048 * 64, n_discard_per_chain=5, seed=2, device=dev, variables=mk(5678))
        assert (ex._streamed_idle_sms(psi, phi) > 0) == (mode == "streamed")
        U, Ud = nkf.operator.Rx(hi, 0, 0.1), nkf.operator.Rx(hi, 0, -0.1)
        op = nkf.InfidelityOperator(phi, U=U, U_dagger=Ud, cv_coeff=-0.5, is_unitary=True)
        psi.expect_and_grad(op)                      # first step: random initial configurations
        psi.reset(); phi.reset()
        st, grad = psi.expect_and_grad(op)
        torch.cuda.synchronize()
        sig, eta = psi.samples_packed(), phi.samples_packed()
        assert tuple(sig.shape) == (2048, 64, (N + 31) // 32) and tuple(eta.shape) == tuple(sig.shape)
        out[mode] = (st, psi._last_flat_grad.cpu().numpy().copy(), sig.cpu(), eta.cpu(), psi.acceptance)
    a, b = out["streamed"], out["plain"]
    same = ((a[2] == b[2]).flatten(1).all(dim=1) & (a[3] == b[3]).flatten(1).all(dim=1)).float().mean().item()
    assert same > 0.9, same
    sa_, sb_ = a[0], b[0]
    assert abs(sa_.mean - sb_.mean) < 5 * float(sb_.error_of_mean) * np.sqrt(1.0 - same + 1e-3) + 1e-6
    assert abs(sa_.variance - sb_.variance) < 0.05 * sb_.variance + 1e-9
    assert abs(float(sa_.error_of_mean) - float(sb_.error_of_mean)) < 0.1 * float(sb_.error_of_mean)
    assert abs(float(sa_.R_hat) - float(sb_.R_hat)) < 0.01 and abs(float(sa_.tau_corr) - float(sb_.tau_corr)) < 0.1 + 0.1 * abs(float(sb_.tau_corr))
    assert np.max(np.abs(a[1] - b[1])) < 0.1 * np.max(np.abs(b[1]))
    assert abs(a[4] - b[4]) < 1e-3
