"""Statistics of the two proposal-site modes of the Metropolis-local sampler.

site_mode 1 ("per_chain") is NetKet's MetropolisLocal: every chain draws its own site at every step.  site_mode 0
("shared", the fast path: all chains of a launch propose the same random site sequence, which lets a CTA stage each
row of the parameter tables once) changes the JOINT law of the chains, not the law of any single chain: the site
sequence is independent of the states and of the acceptance uniforms, every single-site Metropolis kernel leaves
|psi|^2 invariant, and GIVEN the site sequence the chains are independent.  At stationarity E[f(x_c,t) | sites] is the
same constant for every site sequence, so Cov(f(x_c,t), f(x_c',t')) = 0 for c != c': the chains are uncorrelated and
the batch-mean variance is the sum of the per-chain variances, which is what error_of_mean / R_hat assume.

This test measures it on the local estimator L of the infidelity (the quantity whose error bar the driver reports)
over R independent calls: (i) the empirical variance of the batch mean against the between-chain prediction
var_c(chain mean) / n_chains, (ii) the average pairwise covariance of chain means against zero within 5 standard
errors, in BOTH modes; (iii) through the public API, Var(Stats.mean) against mean(Stats.error_of_mean^2).
With 1024 chains a cross-chain correlation of 1e-3 would already double (i)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N, M, N_CHAINS, CHAIN_LEN, R = 20, 40, 1024, 16, 300


def _engine_calls(site_mode):
    import torch
    from netket_fidelity_b200 import _native as nv
    from netket_fidelity_b200.engine import InfidelityStep, StepConfig
    from netket_fidelity_b200.workloads import gate_spec, random_rbm
    dev = torch.device("cuda", 0)
    # a correlated state (acceptance ~ 0.6) and a target close to it (fidelity ~ 0.9): L fluctuates from sample to
    # sample; two unrelated random states would be orthogonal, L = 1/2 identically, and the test empty
    p = random_rbm(N, M, 11, torch.complex64, True, std=0.1)
    d = random_rbm(N, M, 12, torch.complex64, True, std=0.02)
    psi = nv.RbmSpec(*[t.to(dev) for t in p])
    phi = nv.RbmSpec(*[(t + u).to(dev) for t, u in zip(p, d)])
    U, Ud = gate_spec("rx", 3, 0.3), gate_spec("rx", 3, -0.3)
    chain_means = np.empty((R, N_CHAINS))
    for r in range(R):
        cfg = StepConfig(N_CHAINS, CHAIN_LEN, 5, N, seed=1000 + r, cv_coeff=-0.5, site_mode=site_mode)
        eng = InfidelityStep(N, M, cfg, "upsi", U, Ud, device=dev)
        sums, L, _ = eng.step(psi, phi, return_grad=False)
        chain_means[r] = L.view(N_CHAINS, CHAIN_LEN).double().mean(1).cpu().numpy()
    return chain_means


@pytest.mark.parametrize("site_mode", [0, 1])
def test_chains_are_uncorrelated_and_the_between_chain_error_bar_is_right(site_mode):
    cm = _engine_calls(site_mode)                    # (R, n_chains) chain means of L
    batch = cm.mean(1)
    v_emp = batch.var(ddof=1)
    v_pred = (cm.var(axis=1, ddof=1) / N_CHAINS).mean()
    ratio = v_emp / v_pred
    sigma = np.sqrt(2.0 / (R - 1))                   # relative standard deviation of a variance estimate
    assert abs(ratio - 1.0) < 5 * sigma, (site_mode, ratio)
    # average pairwise covariance of the chain means of one call
    d = cm - cm.mean()
    pair = ((d.sum(1)) ** 2 - (d ** 2).sum(1)) / (N_CHAINS * (N_CHAINS - 1))
    se = pair.std(ddof=1) / np.sqrt(R)
    assert abs(pair.mean()) < 5 * se, (site_mode, pair.mean(), se)
    # scale of what would be detected: correlation = covariance / per-chain variance
    rho = pair.mean() / cm.var()
    assert abs(rho) < 2e-3, (site_mode, rho)
    assert cm.var() > 1e-6 and 0.02 < 1.0 - cm.mean() < 0.5      # the regime the comment above promises


@pytest.mark.parametrize("site_mode", ["shared", "per_chain"])
def test_public_error_of_mean_matches_the_scatter_of_the_mean(site_mode):
    """Var over independent runs of Stats.mean vs the mean of Stats.error_of_mean^2 (MCState.expect of the
    infidelity operator, the number InfidelityOptimizer logs)."""
    import netket_fidelity_b200 as nkf
    from netket_fidelity_b200 import nk
    hi = nk.hilbert.Spin(0.5, N)
    import torch
    from netket_fidelity_b200.workloads import random_rbm
    sa = nk.sampler.MetropolisLocal(hi, n_chains=256, sweep_size=N, site_mode=site_mode)
    ma = nk.models.RBM(alpha=2, param_dtype=complex)
    p = random_rbm(N, M, 11, torch.complex128, True, std=0.1)
    d = random_rbm(N, M, 12, torch.complex128, True, std=0.02)
    tree = lambda t: {"Dense": {"kernel": t[0], "bias": t[1]}, "visible_bias": t[2]}
    means, errs = [], []
    runs = 200
    for r in range(runs):
        vs = nk.vqs.MCState(sa, ma, n_samples=256 * 64, n_discard_per_chain=5, sampler_seed=500 + r,
                            variables={"params": tree(p)})
        tg = nk.vqs.MCState(sa, ma, n_samples=256 * 64, n_discard_per_chain=5, sampler_seed=900 + r,
                            variables={"params": tree([a + b for a, b in zip(p, d)])})
        op = nkf.InfidelityOperator(tg, U=nkf.operator.Rx(hi, 3, 0.3), U_dagger=nkf.operator.Rx(hi, 3, -0.3),
                                    cv_coeff=-0.5, is_unitary=True)
        st = vs.expect(op)
        means.append(st.mean)
        errs.append(st.error_of_mean)
    assert 0.02 < np.mean(means) < 0.5
    v_emp = np.var(means, ddof=1)
    v_rep = np.mean(np.square(errs))
    # the block estimator of error_of_mean is itself noisy and slightly biased for 64-sample chains: factor 2 window
    assert 0.5 < v_emp / v_rep < 2.0, (site_mode, v_emp, v_rep)
