"""Multi-GPU plumbing of the infidelity step (one process per GPU, torch.distributed).

The path shards over Markov chains exactly like the reference's MPI scheme (NetKet
`n_chains_per_rank`; gradient mean at infidelity/overlap/expect.py:126): rank r owns the global
chains [r * n_local, (r+1) * n_local) of BOTH families.  Two sums cross the interconnect per step:

  1. eight doubles  [sum L, sum L^2, sum |A|^2, sum Re A, n, n_bad, 0, 0]  after pass A, so that every rank
     centres the score term with the GLOBAL mean F (what NetKet's mpi_statistics does);
  2. the 2P-double gradient accumulator after pass B; then I_grad = -acc / n on every rank.

Everything here is device-agnostic so that the algebra is testable with the gloo backend on CPUs
(tests/test_multirank_gloo.py).
"""

from __future__ import annotations

import torch


def world_info(group=None):
    d = torch.distributed
    if d.is_available() and d.is_initialized():
        return d.get_world_size(group), d.get_rank(group)
    return 1, 0


def shard_chains(n_chains_total: int, world: int, rank: int):
    """(first global chain id, number of local chains) of `rank`."""
    if n_chains_total % world:
        raise ValueError(f"n_chains={n_chains_total} must be divisible by the number of ranks {world}")
    n_local = n_chains_total // world
    return rank * n_local, n_local


def broadcast_seed(seed, group=None) -> int:
    """The seed every rank must use for a default-seeded (seed=None) state: rank 0's value.

    NetKet broadcasts the PRNG key of rank 0 before `model.init`, so that the replicated parameters start
    identical on all ranks; an independent os.urandom draw per rank would make every rank sample and
    differentiate a DIFFERENT state while their sums are all-reduced together.  Single rank: returned as is."""
    world, _ = world_info(group)
    if world == 1:
        return int(seed)
    d = torch.distributed
    dev = torch.device("cuda", torch.cuda.current_device()) if d.get_backend(group) == "nccl" else torch.device("cpu")
    box = torch.tensor([int(seed) & ((1 << 62) - 1)], dtype=torch.int64, device=dev)
    d.broadcast(box, src=d.get_global_rank(group, 0) if group is not None else 0, group=group)
    return int(box.item())


def allreduce_sum_(t: torch.Tensor, group=None):
    world, _ = world_info(group)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM, group=group)
    return t


def allgather_rows(t: torch.Tensor, group=None):
    """Concatenates the equally shaped (rows_local, k) tensors of all ranks along dim 0, in rank order -- the
    per-chain partial sums of nk.stats (chain c of rank r is global chain r * rows_local + c), so that
    error_of_mean / tau_corr / R_hat are those of ALL chains, as NetKet's mpi_statistics computes them."""
    world, _ = world_info(group)
    if world == 1:
        return t
    out = [torch.empty_like(t) for _ in range(world)]
    torch.distributed.all_gather(out, t.contiguous(), group=group)
    return torch.cat(out, dim=0)


def finalize_mean_var(sums):
    """sums = [sum L, sum L^2, ., ., n, ...] (global) -> (F, variance)."""
    n = float(sums[4])
    F = float(sums[0]) / n
    return F, max(0.0, float(sums[1]) / n - F * F)


def finalize_grad(grad_acc: torch.Tensor, sums):
    """I_grad = -acc / n  (infidelity/overlap/expect.py:126-127: mean over ranks, then negate)."""
    return -grad_acc / float(sums[4])
