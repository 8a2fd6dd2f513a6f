"""Oracle: exact (full-summation) ground truth for N <= ~14.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates
* test/_infidelity_exact.py:4-20          (1 - |<new|U|old>|^2 / norms)
* test/_finite_diff.py:8-31               (central differences, grad = d_Re + i d_Im)
* netket_fidelity/infidelity/overlap/exact.py:62-65 and
  netket_fidelity/infidelity/overlap_U/exact.py:71-76 (FullSumState infidelity;
  the latter normalises by <U phi|U phi>, valid for non-unitary U as well).
Basis ordering: NetKet `hilbert.numbers_to_states`: big-endian binary digits,
digit d -> local_states[d] (test/test_operator.py:38-41,66-72: state #2 of
Spin(1/2, 3) is [-1, 1, -1]).
"""

from __future__ import annotations

import numpy as np

from .rbm import RBMParams, rbm_logpsi


def all_states(N, local_states=(-1.0, 1.0)):
    n = np.arange(2 ** N)
    bits = (n[:, None] >> (N - 1 - np.arange(N))[None, :]) & 1
    ls = np.asarray(local_states, dtype=np.float64)
    return ls[bits]


def states_to_numbers(x, local_states=(-1.0, 1.0)):
    x = np.asarray(x)
    bits = (x == local_states[1]).astype(np.int64)
    N = x.shape[-1]
    return (bits << (N - 1 - np.arange(N))).sum(axis=-1)


def to_array(p: RBMParams, normalize=True, local_states=(-1.0, 1.0)):
    psi = np.exp(rbm_logpsi(p, all_states(p.N, local_states)).astype(np.complex128))
    if normalize:
        psi = psi / np.linalg.norm(psi)
    return psi


def infidelity_exact(psi: RBMParams, phi: RBMParams, U=None, local_states=(-1.0, 1.0)):
    """1 - |<psi|U|phi>|^2 / (<psi|psi><U phi|U phi>)  (overlap_U/exact.py:71-76).
    For unitary U this equals test/_infidelity_exact.py:17-20."""
    a = to_array(psi, False, local_states)
    b = to_array(phi, False, local_states)
    if U is not None:
        b = U.to_dense() @ b
    ov = np.vdot(a, b)
    return 1.0 - np.abs(ov) ** 2 / (np.vdot(a, a).real * np.vdot(b, b).real)


def central_diff_grad(func, x, eps):
    """test/_finite_diff.py:8-31."""
    x = np.asarray(x)
    cplx = np.iscomplexobj(x)
    grad = np.zeros(len(x), dtype=np.complex128 if cplx else np.float64)
    for i in range(len(x)):
        e = np.zeros(len(x), dtype=x.dtype)
        e[i] = eps
        gr = 0.5 * (func(x + e) - func(x - e))
        if cplx:
            gi = 0.5 * (func(x + 1j * e) - func(x - 1j * e))
            grad[i] = gr + 1j * gi
        else:
            grad[i] = gr
        grad[i] /= eps
    return grad


def infidelity_exact_grad_fd(psi: RBMParams, phi: RBMParams, U=None, eps=1e-5):
    """Central-difference gradient of the exact infidelity w.r.t. the flattened
    parameters of psi (test/test_infidelity_operator.py:95-103)."""
    flat = psi.ravel()

    def f(v):
        return infidelity_exact(psi.unravel(v), phi, U)

    return central_diff_grad(f, flat, eps)


def exact_distribution(logfun, N, local_states=(-1.0, 1.0)):
    """|f(x)|^2 / sum over all basis states, for any log-amplitude function."""
    lp = logfun(all_states(N, local_states))
    w = np.exp(2 * (lp.real - lp.real.max()))
    return w / w.sum()


def exact_joint_expectation(psi, phi, cv_coeff, U1=None, U2=None, U4=None, Ut=None,
                            return_grad=True):
    """Exact expectation of the MC estimator and of its closed-form gradient by
    enumerating ALL (sigma, eta) pairs with weights |psi(sigma)|^2 |Phi(eta)|^2,
    Phi = phi or Ut phi.  N <= 6.  Used to prove that the estimator is unbiased
    for the exact infidelity and that the gradient formula equals the
    finite-difference gradient without Monte-Carlo noise."""
    from .infidelity import local_log_val, local_estimator
    from .rbm import make_logfun, rbm_theta

    N = psi.N
    st = all_states(N)
    D = st.shape[0]
    p_s = exact_distribution(make_logfun(psi), N)
    p_e = exact_distribution(make_logfun(phi, Ut), N)
    sig = np.repeat(st, D, axis=0)
    eta = np.tile(st, (D, 1))
    w = np.repeat(p_s, D) * np.tile(p_e, D)
    lv, xp2, rho2 = local_log_val(psi, phi, sig, eta, U1, U2, U4)
    A = np.exp(lv)
    c = 0.0 if cv_coeff is None else float(cv_coeff)
    L = local_estimator(lv, cv_coeff)
    F = (w * L).sum()
    out = dict(F=F, I=1 - F, A2=(w * np.abs(A) ** 2).sum(), var=(w * (L - F) ** 2).sum())
    if not return_grad:
        return out
    dL = L - F
    w_eta = np.conj(A) + 2 * c * np.abs(A) ** 2
    w_sig = 2 * dL - w_eta
    t_s = np.conj(np.tanh(rbm_theta(psi, sig)))
    ys = (w * w_sig)[:, None] * t_s
    gW = sig.T.astype(np.complex128) @ ys
    gb = ys.sum(0)
    ga = sig.T.astype(np.complex128) @ (w * w_sig)
    S, C = xp2.shape[0], xp2.shape[1]
    t_e = np.conj(np.tanh(rbm_theta(psi, xp2.reshape(-1, N)).reshape(S, C, -1)))
    we = (w * w_eta)[:, None] * np.conj(rho2)
    ye = we[:, :, None] * t_e
    gW = gW + np.einsum("sci,scj->ij", xp2.astype(np.complex128), ye)
    gb = gb + ye.sum(axis=(0, 1))
    ga = ga + np.einsum("sci,sc->i", xp2.astype(np.complex128), we)
    g = RBMParams(-gW, -gb if psi.bias is not None else None,
                  -ga if psi.visible_bias is not None else None)
    out["grad"] = g
    return out
