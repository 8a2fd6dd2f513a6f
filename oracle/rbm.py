"""Oracle: RBM log-amplitude and log(U phi).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

* RBM: NetKet `nk.models.RBM.__call__` [NetKet 3.10]; the reference carries a
  verbatim copy of it at examples/twospins_GHZ/RBM_Jastrow_ansatz.py:37-62 and a
  parameter fixture at examples/adiabatic_sweeping/initial_state.mpack
  (tree params/Dense/{kernel (N,M), bias (M)}, params/visible_bias (N)).
* log_cosh: NetKet `nknn.log_cosh`.
* log(U phi): netket_fidelity/utils/sampling_Ustate.py:34-47, with
  `jax.scipy.special.logsumexp(a, axis=-1, b=b)` semantics.
"""

from __future__ import annotations

import numpy as np


def log_cosh(x):
    """sgn = -2*signbit(Re x)+1; x = x*sgn; x + log1p(exp(-2x)) - log 2."""
    x = np.asarray(x)
    sgn = -2.0 * np.signbit(x.real) + 1.0
    x = x * sgn
    return x + np.log1p(np.exp(-2.0 * x)) - np.log(2.0)


class RBMParams:
    """kernel (N, M), bias (M) or None, visible_bias (N) or None."""

    def __init__(self, kernel, bias=None, visible_bias=None):
        self.kernel = np.asarray(kernel)
        self.bias = None if bias is None else np.asarray(bias)
        self.visible_bias = None if visible_bias is None else np.asarray(visible_bias)

    @property
    def N(self):
        return self.kernel.shape[0]

    @property
    def M(self):
        return self.kernel.shape[1]

    @property
    def dtype(self):
        return self.kernel.dtype

    def ravel(self):
        """Flatten in pytree-leaf order of the NetKet RBM:
        Dense/bias, Dense/kernel, visible_bias (alphabetical, as jax flattens dicts)."""
        parts = []
        if self.bias is not None:
            parts.append(self.bias.ravel())
        parts.append(self.kernel.ravel())
        if self.visible_bias is not None:
            parts.append(self.visible_bias.ravel())
        return np.concatenate(parts)

    def unravel(self, flat):
        flat = np.asarray(flat)
        o = 0
        bias = None
        if self.bias is not None:
            bias = flat[o:o + self.M]
            o += self.M
        kernel = flat[o:o + self.N * self.M].reshape(self.N, self.M)
        o += self.N * self.M
        vb = None
        if self.visible_bias is not None:
            vb = flat[o:o + self.N]
        return RBMParams(kernel, bias, vb)

    @staticmethod
    def random(N, alpha, seed, dtype=np.complex128, std=0.01, use_hidden_bias=True,
               use_visible_bias=True):
        """normal(stddev=std) per real/imag part (Flax `normal`, RBM default 0.01;
        examples/twospins_GHZ/RBM_Jastrow_ansatz.py:9,30-35)."""
        rng = np.random.default_rng(seed)
        M = int(alpha * N)

        def draw(shape):
            if np.issubdtype(dtype, np.complexfloating):
                return (rng.normal(0, std, shape) + 1j * rng.normal(0, std, shape)).astype(dtype)
            return rng.normal(0, std, shape).astype(dtype)

        k = draw((N, M))
        b = draw((M,)) if use_hidden_bias else None
        a = draw((N,)) if use_visible_bias else None
        return RBMParams(k, b, a)


def rbm_theta(p: RBMParams, x):
    x = np.asarray(x, dtype=p.dtype)
    th = x @ p.kernel
    if p.bias is not None:
        th = th + p.bias
    return th


def rbm_logpsi(p: RBMParams, x):
    """log psi(x) = sum_j log_cosh(b_j + sum_i x_i W_ij) + sum_i a_i x_i."""
    xx = np.asarray(x, dtype=p.dtype)
    out = log_cosh(rbm_theta(p, x)).sum(axis=-1)
    if p.visible_bias is not None:
        out = out + xx @ p.visible_bias
    return out


def logsumexp_b(a, b):
    """jax.scipy.special.logsumexp(a, axis=-1, b=b) for complex inputs:
    shift by max(Re a), sum b*exp(a-shift), principal complex log, add shift."""
    a = np.asarray(a, dtype=np.complex128)
    b = np.asarray(b, dtype=np.complex128)
    amax = np.max(a.real, axis=-1, keepdims=True)
    amax = np.where(np.isfinite(amax), amax, 0.0)
    s = np.sum(b * np.exp(a - amax), axis=-1)
    return np.log(s) + amax[..., 0]


def logpsi_U(p: RBMParams, U, x):
    """utils/sampling_Ustate.py:42-47."""
    x = np.asarray(x)
    xp, mels = U.get_conn_padded(x)
    lp = rbm_logpsi(p, xp.reshape(-1, x.shape[-1])).reshape(mels.shape)
    return logsumexp_b(lp, mels)


def make_logfun(p: RBMParams, U=None):
    if U is None:
        return lambda x: rbm_logpsi(p, x)
    return lambda x: logpsi_U(p, U, x)
