"""The BASELINE.json configurations as concrete synthetic inputs (BASELINE.md section 4,
SURVEY 8d) for bench.py and the examples.  No oracle imports here."""
import math

import numpy as np
import torch

from . import _native as nv


def make_workload(name):
    if name == "C1":   # README example (README.md:49-69)
        return dict(N=4, M=4, n_chains=16, chain_length=7, mode="upsi", U=("hadamard", 0), vb=False)
    if name == "C2":   # 1D TFI chain, one Trotter gate Rx(i, 2 dt h), dt=0.05, h=1
        return dict(N=20, M=40, n_chains=4096, chain_length=16, mode="upsi", U=("rx", 0, 2 * 0.05 * 1.0), vb=True)
    if name == "C3":   # non-unitary U = 1 - i dt H_Ising, sampled from U|phi>
        return dict(N=40, M=80, n_chains=4096, chain_length=64, mode="standard_Uphi",
                    U=("ising_prop", 0.01, 1.0, 1.0), vb=True)
    if name == "C4":   # north star: 10x10, alpha=4, 2^20 sample pairs
        return dict(N=100, M=400, n_chains=16384, chain_length=64, mode="upsi", U=("rx", 0, 0.1), vb=True)
    if name == "C5":   # Renyi-2 swap estimator on an 8x8 lattice, subsystem = half of the system, no gradient
        return dict(N=64, M=128, n_chains=16384, chain_length=64, mode="renyi2", partition=tuple(range(32)), vb=True)
    raise ValueError(name)


def chain_edges(N, pbc=True):
    e = [(i, i + 1) for i in range(N - 1)]
    if pbc and N > 2:
        e.append((N - 1, 0))
    return e


def gate_spec(kind, idx, angle=None):
    if kind == "rx":
        c, s = math.cos(angle / 2), math.sin(angle / 2)
        return nv.OpSpec(nv.OP_GATE, idx, [[c, -1j * s], [c, -1j * s]])
    if kind == "ry":
        c, s = math.cos(angle / 2), math.sin(angle / 2)
        return nv.OpSpec(nv.OP_GATE, idx, [[c, s], [c, -s]])
    if kind == "hadamard":
        r = 1 / math.sqrt(2)
        return nv.OpSpec(nv.OP_GATE, idx, [[r, r], [-r, r]])
    raise ValueError(kind)


def device_ops(w, device):
    u = w["U"]
    if u[0] == "rx":
        return gate_spec("rx", u[1], u[2]), gate_spec("rx", u[1], -u[2])
    if u[0] == "hadamard":
        return gate_spec("hadamard", u[1]), gate_spec("hadamard", u[1])
    if u[0] == "ising_prop":
        dt, h, J = u[1], u[2], u[3]
        e = torch.tensor(chain_edges(w["N"]), dtype=torch.int32, device=device)
        return nv.OpSpec(nv.OP_ISING, 0, None, 1.0, -1j * dt * J, 1j * dt * h, e), None
    raise ValueError(u)


def random_rbm(N, M, seed, cdtype=torch.complex64, visible_bias=True, std=0.01):
    """normal(stddev=std) per real/imag part, NumPy default_rng(seed); same draw order as the oracle's
    RBMParams.random so both bench arms see identical parameters.  Returns host tensors
    (kernel (N,M), bias (M), visible_bias (N) | None)."""
    rng = np.random.default_rng(seed)

    def draw(shape):
        return torch.tensor(rng.normal(0, std, shape) + 1j * rng.normal(0, std, shape)).to(cdtype).contiguous()

    k = draw((N, M))
    b = draw((M,))
    a = draw((N,)) if visible_bias else None
    return k, b, a
