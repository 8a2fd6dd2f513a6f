"""Element-wise parity of the optimiser kernel (nkfb_optimizer_update) with the optax.sgd / optax.adam restatement
of oracle/optim.py over several steps: complex parameters (second moment |g|^2), real-parameter mode (imaginary
parts of gradient and parameters stay out of the update), float32 and float64 buffers."""
import numpy as np
import pytest

from oracle.optim import AdamState, adam_step, sgd_step

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prec,tol", [("c128", 1e-13), ("c64", 2e-6)])
@pytest.mark.parametrize("real_params", [False, True])
def test_adam_matches_optax_restatement(prec, tol, real_params):
    import torch
    from netket_fidelity_b200._native import get_handle
    h = get_handle(0)
    cd = torch.complex128 if prec == "c128" else torch.complex64
    rng = np.random.default_rng(3)
    P = 1237                                                        # not a multiple of the block size
    p0 = rng.normal(size=P) + (0 if real_params else 1j) * rng.normal(size=P)
    p_ref = p0.astype(np.complex128)
    p = torch.tensor(p_ref).to(cd).cuda()
    m = torch.zeros_like(p)
    v = torch.zeros((P,), dtype=p.real.dtype, device="cuda")
    st = AdamState(p_ref)
    lr, b1, b2, eps = 3e-3, 0.9, 0.999, 1e-8
    for t in range(1, 13):
        g = rng.normal(size=P) * 10.0 ** rng.integers(-6, 2, size=P) + 1j * rng.normal(size=P)
        g_used = g.real.astype(np.complex128) if real_params else g
        p_ref = adam_step(p_ref, g_used, st, lr, b1, b2, eps)
        h.optimizer_update(1, p, torch.tensor(g).to(cd).cuda(), m, v, lr, b1, b2, eps, t, real_params)
        got = p.cpu().numpy().astype(np.complex128)
        assert np.max(np.abs(got - p_ref)) < tol * max(1.0, np.max(np.abs(p_ref))), t
        assert np.max(np.abs(m.cpu().numpy() - st.mu)) < tol * max(1.0, np.max(np.abs(st.mu)))
        assert np.max(np.abs(v.cpu().numpy() - st.nu)) < tol * max(1.0, np.max(np.abs(st.nu)))
    if real_params:
        assert np.max(np.abs(got.imag)) == 0


@pytest.mark.parametrize("prec,tol", [("c128", 1e-15), ("c64", 2e-7)])
def test_sgd_matches_restatement(prec, tol):
    import torch
    from netket_fidelity_b200._native import get_handle
    h = get_handle(0)
    cd = torch.complex128 if prec == "c128" else torch.complex64
    rng = np.random.default_rng(4)
    P = 515
    p_ref = rng.normal(size=P) + 1j * rng.normal(size=P)
    p = torch.tensor(p_ref).to(cd).cuda()
    for t in range(5):
        g = rng.normal(size=P) + 1j * rng.normal(size=P)
        p_ref = sgd_step(p_ref, g, 0.05)
        h.optimizer_update(0, p, torch.tensor(g).to(cd).cuda(), None, None, 0.05, 0.0, 0.0, 0.0, 1, False)
    assert np.max(np.abs(p.cpu().numpy() - p_ref)) < tol * np.max(np.abs(p_ref))


def test_public_adam_object_follows_the_restatement():
    """nk.optimizer.Adam applied through an MCState (the path InfidelityOptimizer.run takes)."""
    import torch
    from netket_fidelity_b200 import nk
    hi = nk.hilbert.Spin(0.5, 6)
    vs = nk.vqs.MCState(nk.sampler.MetropolisLocal(hi, n_chains=16), nk.models.RBM(alpha=2, param_dtype=complex),
                        n_samples=64, seed=5)
    opt = nk.optimizer.Adam(learning_rate=0.01)
    p_ref = vs._flat.cpu().numpy().astype(np.complex128)
    st = AdamState(p_ref)
    rng = np.random.default_rng(9)
    for _ in range(4):
        g = (rng.normal(size=p_ref.shape) + 1j * rng.normal(size=p_ref.shape)) * 0.1
        opt.apply(vs, torch.tensor(g).to(vs._flat.dtype).cuda())
        p_ref = adam_step(p_ref, g, st, 0.01)
    assert vs._flat.dtype == torch.complex128                         # param_dtype=complex -> complex128 buffer
    assert np.max(np.abs(vs._flat.cpu().numpy() - p_ref)) < 1e-12 * max(1.0, np.max(np.abs(p_ref)))
