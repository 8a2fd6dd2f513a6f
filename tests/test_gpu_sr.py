"""Stochastic-reconfiguration preconditioner (netket_fidelity_b200.nk.optimizer.SR): the Jacobian-vector product
kernel, the QGT action and the conjugate-gradient solve against the dense NumPy oracle (oracle/sr.py)."""
import numpy as np
import pytest

from oracle.rbm import RBMParams
from oracle import sr as osr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prec,tol", [("c128", 1e-10), ("c64", 2e-5)])
@pytest.mark.parametrize("N,alpha,S", [(6, 2, 500), (20, 2, 2051), (33, 3, 300)])
def test_score_dot_matches_dense_log_derivatives(prec, tol, N, alpha, S):
    import torch
    from netket_fidelity_b200._native import get_handle
    from tests._helpers import to_rbmspec, pack_np, packed_to_torch
    h = get_handle(0)
    cd = torch.complex128 if prec == "c128" else torch.complex64
    p = RBMParams.random(N, alpha, 3, std=0.08)   # away from the poles of tanh, where float32 has no relative accuracy
    rng = np.random.default_rng(1)
    x = rng.choice([-1.0, 1.0], size=(S, N))
    v = rng.normal(size=p.ravel().shape) + 1j * rng.normal(size=p.ravel().shape)
    u = h.score_dot(to_rbmspec(p, cd), packed_to_torch(pack_np(x)), (-1.0, 1.0), torch.tensor(v).to(cd).cuda())
    ref = osr.log_derivatives(p, x) @ v
    assert np.max(np.abs(u.cpu().numpy() - ref)) < tol * np.max(np.abs(ref))


@pytest.mark.parametrize("param_dtype", [complex, float])
def test_sr_solution_matches_dense_solve(param_dtype):
    import torch
    from netket_fidelity_b200 import nk
    N = 8
    hi = nk.hilbert.Spin(0.5, N)
    vs = nk.vqs.MCState(nk.sampler.MetropolisLocal(hi, n_chains=64), nk.models.RBM(alpha=2, param_dtype=param_dtype, stddev=0.2),
                        n_samples=64 * 32, seed=4, sampler_seed=9)
    x = vs.samples.reshape(-1, N).cpu().numpy()
    k, b, a = [t.cpu().numpy() for t in vs._views()]
    cast = (lambda t: t) if param_dtype is complex else (lambda t: t.real.copy())
    p = RBMParams(cast(k), cast(b), cast(a))
    rng = np.random.default_rng(2)
    g = rng.normal(size=vs._flat.shape[0]) + (1j * rng.normal(size=vs._flat.shape[0]) if param_dtype is complex else 0)
    sr = nk.optimizer.SR(diag_shift=0.01, solver_tol=1e-10, maxiter=4000)
    # the QGT action itself
    v = rng.normal(size=g.shape) + (1j * rng.normal(size=g.shape) if param_dtype is complex else 0)
    Sv = sr.qgt_matvec(vs, torch.tensor(v, dtype=torch.complex128).cuda()).cpu().numpy()
    ref_Sv = osr.qgt(p, x) @ v
    assert np.max(np.abs(Sv - ref_Sv)) < 1e-10 * max(1.0, np.max(np.abs(ref_Sv)))
    dp = sr(vs, torch.tensor(g, dtype=torch.complex128).cuda())
    flat = vs._flatten_tree(dp).cpu().numpy()
    ref = osr.sr_solve(p, x, g, 0.01)
    assert np.max(np.abs(flat - ref)) < 1e-7 * np.max(np.abs(ref)), sr.info
    assert sr.info["iterations"] < 4000


def test_infidelity_optimizer_runs_with_the_sr_preconditioner():
    """driver/infidelity_optimizer.py:40,148: `preconditioner=SR(...)`; the infidelity goes down."""
    import netket_fidelity_b200 as nkf
    from netket_fidelity_b200 import nk
    N = 6
    hi = nk.hilbert.Spin(0.5, N)
    sa = nk.sampler.MetropolisLocal(hi, n_chains=64)
    ma = nk.models.RBM(alpha=1, param_dtype=complex)
    phi = nk.vqs.MCState(sa, ma, n_samples=4096, seed=1)
    psi = nk.vqs.MCState(sa, ma, n_samples=4096, seed=2)
    U, Ud = nkf.operator.Rx(hi, 0, 0.4), nkf.operator.Rx(hi, 0, -0.4)
    te = nkf.driver.InfidelityOptimizer(phi, nk.optimizer.Sgd(0.05), variational_state=psi, U=U, U_dagger=Ud,
                                        is_unitary=True, cv_coeff=-0.5, preconditioner=nk.optimizer.SR(diag_shift=1e-3))
    log = nk.logging.RuntimeLog()
    te.run(60, out=log, show_progress=False)
    I = np.asarray(log["Infidelity"]["Mean"])
    assert I[-5:].mean() < 0.5 * I[:5].mean()
