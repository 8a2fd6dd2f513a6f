"""Pin the oracle's estimator and closed-form gradient with the reference's own
test method (test/test_infidelity_operator.py:95-127,155-182): the exact
full-summation infidelity and its central-difference gradient -- but without
Monte-Carlo noise, by enumerating all (sigma, eta) pairs -- and with an
independent torch-autograd restatement of nkjax.expect + nkjax.vjp."""
import numpy as np
import pytest
import torch

from oracle import operators as ops
from oracle.rbm import RBMParams, rbm_logpsi, logpsi_U, make_logfun
from oracle.exact import (all_states, infidelity_exact, infidelity_exact_grad_fd,
                          exact_joint_expectation, exact_distribution)
from oracle.infidelity import infidelity_and_grad, mode_slots

N = 3


def _params(dtype, scale=0.3, vb=True):
    psi = RBMParams.random(N, 1, 11, dtype=dtype, std=scale, use_visible_bias=vb)
    phi = RBMParams.random(N, 1, 22, dtype=dtype, std=scale, use_visible_bias=vb)
    return psi, phi


CASES = [
    ("standard", None),
    ("upsi", ops.Rx(N, 0, 0.37)),
    ("upsi", ops.Ry(N, 1, 0.51)),
    ("upsi", ops.Hadamard(N, 2)),
    ("standard_Uphi", ops.Rx(N, 1, 0.37)),
    ("standard_Uphi", ops.IsingPropagator(N, ops.chain_edges(N), 1.0, 1.0, 0.05)),
]


def _Udag(U):
    if U is None:
        return None
    if isinstance(U, ops.Ry):      # the reference's Ry.H is not the adjoint (appendix C): pass it explicitly
        return ops.Ry(U.N, U.idx, -U.angle)
    return U.H


@pytest.mark.parametrize("dtype", [np.complex128, np.float64])
@pytest.mark.parametrize("mode,U", CASES)
@pytest.mark.parametrize("cv", [None, -0.5, 0.3])
def test_estimator_unbiased_and_gradient_matches_finite_differences(mode, U, cv, dtype):
    psi, phi = _params(dtype)
    if dtype == np.float64 and isinstance(U, ops.Ry):
        pass
    U1, U2, U4 = mode_slots(mode, U, _Udag(U))
    Ut = U if mode == "standard_Uphi" else None
    ex = exact_joint_expectation(psi, phi, cv, U1, U2, U4, Ut)
    I_exact = infidelity_exact(psi, phi, U)
    assert abs(ex["I"] - I_exact) < 1e-12
    assert abs(ex["A2"] - 1.0) < 1e-12          # E|A|^2 = 1: the CV term has zero mean
    g_fd = infidelity_exact_grad_fd(psi, phi, U, eps=1e-5)
    g = ex["grad"].ravel()
    if dtype == np.float64:
        g = g.real
    np.testing.assert_allclose(g, g_fd, atol=2e-9)


def _torch_logcosh(z):
    sgn = torch.where(z.real < 0, -1.0, 1.0)
    z = z * sgn
    return z + torch.log(1 + torch.exp(-2 * z)) - np.log(2.0)


def _torch_logpsi(W, b, a, x):
    th = x.to(W.dtype) @ W + b
    return _torch_logcosh(th).sum(-1) + x.to(W.dtype) @ a


@pytest.mark.parametrize("mode,U", CASES[:5])
def test_gradient_matches_autograd_restatement_of_nkjax_expect(mode, U):
    """nkjax.expect backward: grad = vjp of mean_s[dL_s * log_pdf(theta, s) + f(theta, s)]
    with dL held constant; torch's complex gradient convention is dF/dRe + i dF/dIm (= A.4)."""
    rng = np.random.default_rng(5)
    psi, phi = _params(np.complex128)
    S = 64
    sigma = rng.choice([-1.0, 1.0], size=(S, N))
    eta = rng.choice([-1.0, 1.0], size=(S, N))
    cv = -0.5
    U1, U2, U4 = mode_slots(mode, U, _Udag(U))
    res = infidelity_and_grad(psi, phi, sigma, eta, cv, U1, U2, U4)

    W = torch.tensor(psi.kernel, requires_grad=True)
    b = torch.tensor(psi.bias, requires_grad=True)
    a = torch.tensor(psi.visible_bias, requires_grad=True)
    Wt, bt, at = (torch.tensor(v) for v in (phi.kernel, phi.bias, phi.visible_bias))

    def T(net, x, Uop):
        if Uop is None:
            return _torch_logpsi(*net, torch.tensor(x))
        xp, mels = Uop.get_conn_padded(x)
        lp = _torch_logpsi(*net, torch.tensor(xp.reshape(-1, N))).reshape(mels.shape)
        return torch.log((torch.tensor(mels.astype(np.complex128)) * torch.exp(lp)).sum(-1))

    lv = T((Wt, bt, at), sigma, U1) + T((W, b, a), eta, U2) - T((W, b, a), sigma, None) - T((Wt, bt, at), eta, U4)
    Lloc = torch.exp(lv).real + cv * (torch.exp(2 * lv.real) - 1)
    np.testing.assert_allclose(Lloc.detach().numpy(), res["L"], rtol=1e-11, atol=1e-12)
    # log_pdf_joint: only the psi-dependent part matters for d/dtheta (overlap/expect.py:91-101)
    log_pdf = 2 * _torch_logpsi(W, b, a, torch.tensor(sigma)).real
    dL = (Lloc - Lloc.mean()).detach()
    obj = (dL * log_pdf + Lloc).mean()
    obj.backward()
    g_auto = -np.concatenate([b.grad.numpy().ravel(), W.grad.numpy().ravel(), a.grad.numpy().ravel()])
    np.testing.assert_allclose(res["grad"].ravel(), g_auto, rtol=1e-9, atol=1e-12)


def test_logpsi_U_is_log_of_U_applied_to_state():
    psi, _ = _params(np.complex128)
    for U in (ops.Rx(N, 0, 0.4), ops.Hadamard(N, 1), ops.IsingPropagator(N, ops.chain_edges(N), 0.7, 1.1, 0.03)):
        st = all_states(N)
        vec = np.exp(rbm_logpsi(psi, st))
        np.testing.assert_allclose(np.exp(logpsi_U(psi, U, st)), U.to_dense() @ vec, rtol=1e-12)


def test_cv_does_not_change_the_mean_but_reduces_variance_near_convergence():
    psi = RBMParams.random(N, 1, 3, std=0.3)
    phi = RBMParams(psi.kernel * 1.02, psi.bias + 0.01, psi.visible_bias)
    e0 = exact_joint_expectation(psi, phi, None, return_grad=False)
    e1 = exact_joint_expectation(psi, phi, -0.5, return_grad=False)
    assert abs(e0["F"] - e1["F"]) < 1e-13
    assert e1["var"] < 0.05 * e0["var"]
