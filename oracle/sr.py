"""Oracle: log-derivatives of the RBM and the stochastic-reconfiguration solve.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

NetKet `nk.optimizer.SR(diag_shift)` [NetKet 3.10, not in /root/reference; call site: the `preconditioner=` argument of
netket_fidelity/driver/infidelity_optimizer.py:40, applied at :148]: dp = (S + diag_shift 1)^-1 grad with the quantum
geometric tensor S_kl = <O_k^* O_l> - <O_k^*><O_l> over the samples (complex holomorphic parameters), Re S for real
parameters.  O_k of the RBM: bias tanh(theta_j), kernel x_i tanh(theta_j), visible bias x_i, in the flat
leaf order [bias | kernel | visible_bias].  parity unpinned (no vector in the reference)."""
import numpy as np

from .rbm import RBMParams, rbm_theta


def log_derivatives(p: RBMParams, x):
    """O (S, P) complex128 in the flat layout of RBMParams.ravel()."""
    x = np.asarray(x, dtype=np.float64)
    t = np.tanh(rbm_theta(p, x).astype(np.complex128))          # (S, M)
    parts = []
    if p.bias is not None:
        parts.append(t)
    parts.append((x[:, :, None] * t[:, None, :]).reshape(x.shape[0], -1))
    if p.visible_bias is not None:
        parts.append(x.astype(np.complex128))
    return np.concatenate(parts, axis=1)


def qgt(p: RBMParams, x):
    O = log_derivatives(p, x)
    Oc = O - O.mean(0, keepdims=True)
    S = Oc.conj().T @ Oc / x.shape[0]
    return S if np.issubdtype(p.dtype, np.complexfloating) else S.real


def sr_solve(p: RBMParams, x, grad_flat, diag_shift):
    S = qgt(p, x)
    return np.linalg.solve(S + diag_shift * np.eye(S.shape[0]), grad_flat)
