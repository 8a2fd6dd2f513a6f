"""Shared helpers of the GPU parity tests: oracle objects -> device specs."""
import numpy as np
import torch

from oracle import operators as oops
from oracle.rbm import RBMParams


def to_rbmspec(p: RBMParams, cdtype, device="cuda"):
    from netket_fidelity_b200._native import RbmSpec

    def cv(a):
        if a is None:
            return None
        return torch.tensor(np.asarray(a).astype(np.complex128)).to(cdtype).to(device).contiguous()

    return RbmSpec(cv(p.kernel), cv(p.bias), cv(p.visible_bias))


def to_opspec(op, device="cuda"):
    from netket_fidelity_b200._native import OpSpec, OP_GATE, OP_ISING

    if op is None:
        return None
    if isinstance(op, oops.Rx):
        c, s = np.cos(op.angle / 2), np.sin(op.angle / 2)
        return OpSpec(OP_GATE, op.idx, [[c, -1j * s], [c, -1j * s]])
    if isinstance(op, oops.Ry):
        c, s = np.cos(op.angle / 2), np.sin(op.angle / 2)
        return OpSpec(OP_GATE, op.idx, [[c, s], [c, -s]])
    if isinstance(op, oops.Hadamard):
        r = 1 / np.sqrt(2)
        return OpSpec(OP_GATE, op.idx, [[r, r], [-r, r]])
    if isinstance(op, oops.IsingLike):
        e = torch.tensor(op.edges.astype(np.int32)).to(device).contiguous()
        return OpSpec(OP_ISING, 0, None, op.c0, op.c1, op.c2, e)
    raise TypeError(op)


def pack_np(x, s1=1.0):
    """numpy reference of the bit packing (site i -> bit i&31 of word i>>5)."""
    x = np.asarray(x)
    B, N = x.shape
    W = (N + 31) // 32
    out = np.zeros((B, W), dtype=np.uint32)
    for i in range(N):
        out[:, i >> 5] |= ((x[:, i] == s1).astype(np.uint32) << np.uint32(i & 31))
    return out


def unpack_np(p, N, ls=(-1.0, 1.0)):
    p = np.asarray(p).view(np.uint32) if np.asarray(p).dtype == np.int32 else np.asarray(p).astype(np.uint32)
    i = np.arange(N)
    bits = (p[:, i >> 5] >> (i & 31).astype(np.uint32)) & 1
    return np.where(bits == 1, ls[1], ls[0]).astype(np.float64)


def packed_to_torch(p, device="cuda"):
    return torch.tensor(np.ascontiguousarray(p).view(np.int32)).to(device)


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)
