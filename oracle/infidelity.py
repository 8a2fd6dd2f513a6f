"""Oracle: infidelity local estimator, control variates and score-function gradient.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates
* netket_fidelity/infidelity/overlap/expect.py:71-130      (Standard: Phi is a state)
* netket_fidelity/infidelity/overlap_U/expect.py:85-161    (UPsi: sample |psi|^2|phi|^2,
                                                            apply U / U^dagger as conns)
* `nkjax.expect` custom VJP + `nkjax.vjp(conjugate=True)` [NetKet 3.10; SURVEY A.3/A.4]
  in closed form.

Conventions: psi = variational state (parameters being differentiated), phi =
target.  sigma ~ |psi|^2, eta ~ |Phi|^2.  All four terms are written as
`T(net, samples, U_or_None)` = log-sum-exp over the connected elements of U
(identity if None):

    Standard           : l = T(phi,sig,Ut) + T(psi,eta,None) - T(psi,sig,None) - T(phi,eta,Ut)
                         (Ut = None for a plain target, Ut = U when the target is U|phi>,
                          overlap/operator.py:61-82)
    UPsi (unitary U)   : l = T(phi,sig,U) + T(psi,eta,Udag) - T(psi,sig,None) - T(phi,eta,None)

so one function with three operator slots (U1 on (phi,sigma), U2 on (psi,eta),
U4 on (phi,eta)) covers every mode.
"""

from __future__ import annotations

import numpy as np

from .rbm import RBMParams, rbm_logpsi, rbm_theta, logsumexp_b


def _T(p, x, U):
    """log sum_c mel_c * net(x'_c); returns (T[S], xp[S,C,N], rho[S,C])."""
    x = np.asarray(x)
    if U is None:
        lp = rbm_logpsi(p, x)
        return lp, x[:, None, :], np.ones((x.shape[0], 1), dtype=np.complex128)
    xp, mels = U.get_conn_padded(x)
    lp = rbm_logpsi(p, xp.reshape(-1, x.shape[-1])).reshape(mels.shape)
    T = logsumexp_b(lp, mels)
    rho = mels * np.exp(lp - T[:, None])
    return T, xp, rho


def local_log_val(psi: RBMParams, phi: RBMParams, sigma, eta, U1=None, U2=None, U4=None):
    """overlap/expect.py:85, overlap_U/expect.py:108-116."""
    T1, _, _ = _T(phi, sigma, U1)
    T2, xp2, rho2 = _T(psi, eta, U2)
    T3 = rbm_logpsi(psi, sigma)
    T4, _, _ = _T(phi, eta, U4)
    return T1 + T2 - T3 - T4, xp2, rho2


def local_estimator(log_val, cv_coeff):
    """overlap/expect.py:86-89: Re exp(l) + c (exp(2 Re l) - 1)."""
    res = np.exp(log_val).real
    if cv_coeff is not None:
        res = res + cv_coeff * (np.exp(2 * log_val.real) - 1)
    return res


def infidelity_and_grad(psi, phi, sigma, eta, cv_coeff=None, U1=None, U2=None, U4=None,
                        return_grad=True):
    """Returns dict(L[S], F, I, grad RBMParams or None).

    Gradient (convention dF/dRe + i dF/dIm for complex parameters,
    test/_finite_diff.py:19-22; ordinary gradient = real part for real parameters):

        A = exp(l), L = Re A + c(|A|^2-1), dL = L - mean L
        w_eta = conj(A) + 2c|A|^2 ;  w_sig = 2 dL - w_eta
        gradF_k = mean_s [ w_sig conj O_k(sigma_s) + w_eta conj D_k(eta_s) ]
        D_k(eta) = sum_c rho_c O_k(eta'_c),  rho_c = U2_c psi(eta'_c) / sum_l U2_l psi(eta'_l)
        I_grad = -gradF          (overlap/expect.py:127, overlap_U/expect.py:158)
    """
    sigma = np.asarray(sigma, dtype=np.float64)
    eta = np.asarray(eta, dtype=np.float64)
    S = sigma.shape[0]
    lv, xp2, rho2 = local_log_val(psi, phi, sigma, eta, U1, U2, U4)
    A = np.exp(lv)
    c = 0.0 if cv_coeff is None else float(cv_coeff)
    L = local_estimator(lv, cv_coeff)
    F = L.mean()
    out = dict(L=L, A=A, log_val=lv, F=F, I=1.0 - F, grad=None)
    if not return_grad:
        return out

    dL = L - F
    w_eta = np.conj(A) + 2 * c * np.abs(A) ** 2
    w_sig = 2 * dL - w_eta

    # O_k(x): kernel x_i tanh(theta_j), bias tanh(theta_j), visible_bias x_i
    th_s = rbm_theta(psi, sigma)                       # (S,M)
    t_s = np.conj(np.tanh(th_s))
    ys = w_sig[:, None] * t_s                          # (S,M)
    gW = sigma.T.astype(np.complex128) @ ys
    gb = ys.sum(axis=0)
    ga = sigma.T.astype(np.complex128) @ w_sig

    C = xp2.shape[1]
    N = sigma.shape[1]
    th_e = rbm_theta(psi, xp2.reshape(-1, N)).reshape(S, C, -1)
    t_e = np.conj(np.tanh(th_e))                       # (S,C,M)
    we = w_eta[:, None] * np.conj(rho2)                # (S,C)
    ye = we[:, :, None] * t_e                          # (S,C,M)
    xc = xp2.reshape(S * C, N).astype(np.complex128)   # sum over (s, c) as one matrix product (BLAS)
    gW = gW + xc.T @ ye.reshape(S * C, -1)
    gb = gb + ye.sum(axis=(0, 1))
    ga = ga + xc.T @ we.reshape(S * C)

    gW, gb, ga = -gW / S, -gb / S, -ga / S
    if not np.issubdtype(psi.dtype, np.complexfloating):
        gW, gb, ga = gW.real, gb.real, ga.real
    out["grad"] = RBMParams(
        gW,
        gb if psi.bias is not None else None,
        ga if psi.visible_bias is not None else None,
    )
    return out


def mode_slots(mode, U=None, U_dagger=None):
    """Map the reference's three operator kinds (infidelity/logic.py:142-187) onto
    the (U1, U2, U4) slots."""
    if mode == "standard":            # InfidelityOperatorStandard, plain target
        return None, None, None
    if mode == "standard_Uphi":       # InfidelityUPsi: target state is U|phi>
        return U, None, U
    if mode == "upsi":                # InfidelityOperatorUPsi: unitary U, conns of U and U^dagger
        return U, U_dagger, None
    raise ValueError(mode)
