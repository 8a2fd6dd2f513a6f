"""CPU baseline of one infidelity-gradient step, following the REFERENCE's algorithm:
NetKet MetropolisLocal with a full forward pass per Metropolis step, then the
estimator and score-function gradient of infidelity/overlap_U/expect.py:85-161.

TEST INFRASTRUCTURE / REPORTED BASELINE ONLY (bench.py `cpu_baseline` and `--impl reference`).
NumPy complex128, BLAS threads = all host cores.
"""
import os
import time

import numpy as np

from .rbm import RBMParams, make_logfun
from .sampler import metropolis_local_netket
from .infidelity import infidelity_and_grad, mode_slots
from . import operators as oops


def make_workload(name):
    """The BASELINE.json configurations as concrete inputs."""
    if name == "C1":
        N, alpha, U = 4, 1, ("hadamard", 0)
        return dict(N=N, M=4, n_chains=16, chain_length=7, mode="upsi", U=U, vb=False)
    if name == "C2":
        return dict(N=20, M=40, n_chains=4096, chain_length=16, mode="upsi", U=("rx", 0, 2 * 0.05 * 1.0), vb=True)
    if name == "C3":
        return dict(N=40, M=80, n_chains=4096, chain_length=64, mode="standard_Uphi",
                    U=("ising_prop", 0.01, 1.0, 1.0), vb=True)
    if name == "C4":
        return dict(N=100, M=400, n_chains=16384, chain_length=64, mode="upsi", U=("rx", 0, 0.1), vb=True)
    raise ValueError(name)


def oracle_ops(w):
    N, u = w["N"], w["U"]
    if u[0] == "rx":
        return oops.Rx(N, u[1], u[2]), oops.Rx(N, u[1], -u[2])
    if u[0] == "hadamard":
        return oops.Hadamard(N, u[1]), oops.Hadamard(N, u[1])
    if u[0] == "ising_prop":
        e = oops.chain_edges(N, pbc=True)
        return oops.IsingPropagator(N, e, u[2], u[3], u[1]), None
    raise ValueError(u)


def cpu_step(w, n_chains, chain_length, n_discard=5, seed=0, psi=None, phi=None):
    """One full step on the CPU; returns (seconds, n_sample_pairs, result dict)."""
    N, M = w["N"], w["M"]
    psi = psi or RBMParams.random(N, M // N, 1234, std=0.01, use_visible_bias=w["vb"])
    phi = phi or RBMParams.random(N, M // N, 5678, std=0.01, use_visible_bias=w["vb"])
    U, Ud = oracle_ops(w)
    U1, U2, U4 = mode_slots(w["mode"], U, Ud)
    rng = np.random.default_rng(seed)
    t0 = time.perf_counter()
    sig, _ = metropolis_local_netket(make_logfun(psi), N, n_chains, chain_length, n_discard, rng=rng)
    Ut = U if w["mode"] == "standard_Uphi" else None
    eta, _ = metropolis_local_netket(make_logfun(phi, Ut), N, n_chains, chain_length, n_discard, rng=rng)
    res = infidelity_and_grad(psi, phi, sig.reshape(-1, N), eta.reshape(-1, N), -0.5, U1, U2, U4)
    dt = time.perf_counter() - t0
    return dt, n_chains * chain_length, res
