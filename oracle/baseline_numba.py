"""CPU baseline of the sampling part of one infidelity-gradient step on ALL host cores (plan B).

TEST INFRASTRUCTURE / REPORTED BASELINE ONLY (bench.py `cpu_baseline` and `--impl reference`).

The reference's algorithm, not ours: NetKet `MetropolisLocal` (call sites
netket_fidelity/infidelity/overlap_U/expect.py:21-22, driver/infidelity_optimizer.py:142-143) proposes one spin
flip per chain and step and evaluates the new log-probability with a FULL forward pass of the model
(theta = b + x W from scratch: N M complex-by-real multiply-adds, then M log-cosh), in complex128.  Here that is
compiled with numba, `prange` over the Markov chains (they are independent), one thread per host core.  The
random numbers are drawn up front in the order `oracle.sampler.metropolis_local_netket` draws them, so that the
two implementations walk the same trajectories (tests/test_oracle_baseline.py).  Composite targets log(U phi)
(`sample_Upsi=True`, utils/sampling_Ustate.py:34-47 inside the MH loop) are supported for gates (2 connected
elements); Ising-type operators fall back to the NumPy sampler.

Only Re log psi is needed by the acceptance test, so the log-cosh is evaluated as
Re log cosh(x + i y) = |x| - ln 2 + 1/2 ln((1 + t cos 2y)^2 + (t sin 2y)^2),  t = exp(-2 |x|)
(the real part of NetKet's  x + log1p(exp(-2x)) - ln 2  after the sign fold; same value, no complex log).
"""
import math
import os
import time

import numpy as np

try:  # numba is part of this image; the NumPy sampler remains the fallback and the cross-check
    import numba
    from numba import njit, prange
    HAVE_NUMBA = True
except Exception:  # pragma: no cover
    HAVE_NUMBA = False

from .rbm import RBMParams, make_logfun
from .sampler import metropolis_local_netket
from .infidelity import infidelity_and_grad, mode_slots

LN2 = 0.6931471805599453

if HAVE_NUMBA:

    @njit(fastmath=False, cache=True, inline="always")
    def _re_logpsi(x, Wr, Wi, br, bi, ar, thr, thi):
        """Re log psi(x) by a full forward pass (what NetKet's sampler evaluates at every Metropolis step)."""
        N, M = Wr.shape
        for j in range(M):
            thr[j] = br[j]
            thi[j] = bi[j]
        for i in range(N):
            xi = x[i]
            for j in range(M):
                thr[j] += xi * Wr[i, j]
                thi[j] += xi * Wi[i, j]
        acc = 0.0
        for j in range(M):
            ax = abs(thr[j])
            t = math.exp(-2.0 * ax)
            c = t * math.cos(2.0 * thi[j])
            s = t * math.sin(2.0 * thi[j])
            acc += ax - LN2 + 0.5 * math.log((1.0 + c) * (1.0 + c) + s * s)
        for i in range(N):
            acc += ar[i] * x[i]
        return acc

    @njit(fastmath=False, cache=True, inline="always")
    def _logpsi_complex(x, Wr, Wi, br, bi, ar, ai, thr, thi):
        """complex log psi(x) (needed for the log-sum-exp over connected configurations of a composite target)"""
        N, M = Wr.shape
        for j in range(M):
            thr[j] = br[j]
            thi[j] = bi[j]
        for i in range(N):
            xi = x[i]
            for j in range(M):
                thr[j] += xi * Wr[i, j]
                thi[j] += xi * Wi[i, j]
        re, im = 0.0, 0.0
        for j in range(M):
            xr, xi_ = thr[j], thi[j]
            if xr < 0.0:
                xr, xi_ = -xr, -xi_
            t = math.exp(-2.0 * xr)
            c = 1.0 + t * math.cos(2.0 * xi_)
            s = -t * math.sin(2.0 * xi_)
            re += xr - LN2 + 0.5 * math.log(c * c + s * s)
            im += xi_ + math.atan2(s, c)
        for i in range(N):
            re += ar[i] * x[i]
            im += ai[i] * x[i]
        return re, im

    @njit(parallel=True, fastmath=False, cache=True)
    def _mh_plain(Wr, Wi, br, bi, ar, x0, sites, us, n_discard, chain_length, sweep_size, s0, s1, out):
        n_chains, N = x0.shape
        M = Wr.shape[1]
        for c in prange(n_chains):
            x = x0[c].copy()
            thr = np.empty(M)
            thi = np.empty(M)
            logp = 2.0 * _re_logpsi(x, Wr, Wi, br, bi, ar, thr, thi)
            t = 0
            for sw in range(n_discard + chain_length):
                for _ in range(sweep_size):
                    i = sites[t, c]
                    u = us[t, c]
                    t += 1
                    xold = x[i]
                    x[i] = s1 if xold == s0 else s0
                    logpn = 2.0 * _re_logpsi(x, Wr, Wi, br, bi, ar, thr, thi)
                    if u < math.exp(logpn - logp):
                        logp = logpn
                    else:
                        x[i] = xold
                if sw >= n_discard:
                    for i in range(N):
                        out[c, sw - n_discard, i] = x[i]
            for i in range(N):
                x0[c, i] = x[i]

    @njit(fastmath=False, cache=True, inline="always")
    def _re_logpsi_gate(x, Wr, Wi, br, bi, ar, ai, thr, thi, idx, mel, s0, s1):
        """Re log sum_c mel[bit][c] psi(x'_c) for a single-site gate: x'_0 = x, x'_1 = x with site idx flipped."""
        b = 0 if x[idx] == s0 else 1
        r0, i0 = _logpsi_complex(x, Wr, Wi, br, bi, ar, ai, thr, thi)
        xo = x[idx]
        x[idx] = s1 if xo == s0 else s0
        r1, i1 = _logpsi_complex(x, Wr, Wi, br, bi, ar, ai, thr, thi)
        x[idx] = xo
        m = max(r0, r1)
        e0, e1 = math.exp(r0 - m), math.exp(r1 - m)
        sr = e0 * (mel[b, 0].real * math.cos(i0) - mel[b, 0].imag * math.sin(i0)) + \
            e1 * (mel[b, 1].real * math.cos(i1) - mel[b, 1].imag * math.sin(i1))
        si = e0 * (mel[b, 0].real * math.sin(i0) + mel[b, 0].imag * math.cos(i0)) + \
            e1 * (mel[b, 1].real * math.sin(i1) + mel[b, 1].imag * math.cos(i1))
        return m + 0.5 * math.log(sr * sr + si * si)

    @njit(parallel=True, fastmath=False, cache=True)
    def _mh_gate(Wr, Wi, br, bi, ar, ai, x0, sites, us, n_discard, chain_length, sweep_size, s0, s1, idx, mel, out):
        n_chains, N = x0.shape
        M = Wr.shape[1]
        for c in prange(n_chains):
            x = x0[c].copy()
            thr = np.empty(M)
            thi = np.empty(M)
            logp = 2.0 * _re_logpsi_gate(x, Wr, Wi, br, bi, ar, ai, thr, thi, idx, mel, s0, s1)
            t = 0
            for sw in range(n_discard + chain_length):
                for _ in range(sweep_size):
                    i = sites[t, c]
                    u = us[t, c]
                    t += 1
                    xold = x[i]
                    x[i] = s1 if xold == s0 else s0
                    logpn = 2.0 * _re_logpsi_gate(x, Wr, Wi, br, bi, ar, ai, thr, thi, idx, mel, s0, s1)
                    if u < math.exp(logpn - logp):
                        logp = logpn
                    else:
                        x[i] = xold
                if sw >= n_discard:
                    for i in range(N):
                        out[c, sw - n_discard, i] = x[i]
            for i in range(N):
                x0[c, i] = x[i]


def n_threads():
    return numba.get_num_threads() if HAVE_NUMBA else 1


def _split(p: RBMParams):
    N, M = p.N, p.M
    W = np.ascontiguousarray(p.kernel.astype(np.complex128))
    b = np.zeros(M, dtype=np.complex128) if p.bias is None else p.bias.astype(np.complex128)
    a = np.zeros(N, dtype=np.complex128) if p.visible_bias is None else p.visible_bias.astype(np.complex128)
    c = np.ascontiguousarray
    return c(W.real), c(W.imag), c(b.real), c(b.imag), c(a.real), c(a.imag)


def gate_mel_table(U):
    """mel[bit at idx][connected element] of a single-site gate (oracle.operators.Rx / Ry / Hadamard)."""
    N, ls = U.N, U.local_states
    x = np.full((2, N), ls[0])
    x[1, U.idx] = ls[1]
    _, mels = U.get_conn_padded(x)
    return np.ascontiguousarray(np.asarray(mels, dtype=np.complex128))


def metropolis_local_numba(p: RBMParams, N, n_chains, chain_length, n_discard, sweep_size=None, rng=None, init=None,
                           local_states=(-1.0, 1.0), U=None):
    """Same contract and the same random-number order as oracle.sampler.metropolis_local_netket(make_logfun(p, U))."""
    if not HAVE_NUMBA:
        raise RuntimeError("numba is not importable")
    rng = np.random.default_rng(0) if rng is None else rng
    sweep_size = N if sweep_size is None else sweep_size
    ls = np.asarray(local_states, dtype=np.float64)
    x = ls[rng.integers(0, 2, size=(n_chains, N))] if init is None else np.array(init, dtype=np.float64)
    T = (n_discard + chain_length) * sweep_size
    sites = np.empty((T, n_chains), dtype=np.int64)
    us = np.empty((T, n_chains))
    for t in range(T):                                   # the draw order of the NumPy sampler
        sites[t] = rng.integers(0, N, size=n_chains)
        us[t] = rng.random(n_chains)
    Wr, Wi, br, bi, ar, ai = _split(p)
    out = np.empty((n_chains, chain_length, N))
    if U is None:
        _mh_plain(Wr, Wi, br, bi, ar, x, sites, us, n_discard, chain_length, sweep_size, ls[0], ls[1], out)
    else:
        _mh_gate(Wr, Wi, br, bi, ar, ai, x, sites, us, n_discard, chain_length, sweep_size, ls[0], ls[1], int(U.idx),
                 gate_mel_table(U), out)
    return out, x


def cpu_step(w, n_chains, chain_length, n_discard=5, seed=0, psi=None, phi=None, engine="numba"):
    """One full step (sampling of both families + estimator + gradient) on the host cores.
    Returns (seconds, n_sample_pairs, result dict, info dict)."""
    from .baseline import oracle_ops
    N, M = w["N"], w["M"]
    psi = psi or RBMParams.random(N, M // N, 1234, std=0.01, use_visible_bias=w["vb"])
    phi = phi or RBMParams.random(N, M // N, 5678, std=0.01, use_visible_bias=w["vb"])
    U, Ud = oracle_ops(w)
    U1, U2, U4 = mode_slots(w["mode"], U, Ud)
    Ut = U if w["mode"] == "standard_Uphi" else None
    rng = np.random.default_rng(seed)
    use_numba = engine == "numba" and HAVE_NUMBA and (Ut is None or hasattr(Ut, "idx"))
    t0 = time.perf_counter()
    if use_numba:
        sig, _ = metropolis_local_numba(psi, N, n_chains, chain_length, n_discard, rng=rng)
        eta, _ = metropolis_local_numba(phi, N, n_chains, chain_length, n_discard, rng=rng, U=Ut)
    else:
        sig, _ = metropolis_local_netket(make_logfun(psi), N, n_chains, chain_length, n_discard, rng=rng)
        eta, _ = metropolis_local_netket(make_logfun(phi, Ut), N, n_chains, chain_length, n_discard, rng=rng)
    t1 = time.perf_counter()
    res = infidelity_and_grad(psi, phi, sig.reshape(-1, N), eta.reshape(-1, N), -0.5, U1, U2, U4)
    t2 = time.perf_counter()
    info = dict(engine="numba prange over chains" if use_numba else "numpy", threads=n_threads() if use_numba else None,
                sampling_s=t1 - t0, estimator_s=t2 - t1, dtype="complex128")
    return t2 - t0, n_chains * chain_length, res, info


def warm_up():
    """JIT-compile (or load from cache) outside any timed region."""
    if not HAVE_NUMBA:
        return
    p = RBMParams.random(4, 1, 1)
    metropolis_local_numba(p, 4, 2, 1, 0)
    from . import operators as oops
    metropolis_local_numba(p, 4, 2, 1, 0, U=oops.Rx(4, 0, 0.1))
