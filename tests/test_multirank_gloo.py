"""world_size-2 gloo test (CPU) of the multi-rank algebra of netket_fidelity_b200/parallel.py: each rank
evaluates pass A / pass B on ITS shard of the sample pairs (the NumPy oracle stands in for the CUDA kernels),
the two all-reduces combine them, and the result must equal the single-rank estimator and gradient.  Also
checks the chain sharding used for the Philox chain ids."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import operators as oops
from oracle.rbm import RBMParams, rbm_theta
from oracle.infidelity import infidelity_and_grad, local_log_val, local_estimator


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _local_pass_B(psi, sigma, eta, A, L, F, cv, xp2, rho2):
    """un-normalised accumulators sum_s [w_sig conj O(sigma) + w_eta conj D(eta)] in the flat layout."""
    c = 0.0 if cv is None else cv
    w_eta = np.conj(A) + 2 * c * np.abs(A) ** 2
    w_sig = 2 * (L - F) - w_eta
    ys = w_sig[:, None] * np.conj(np.tanh(rbm_theta(psi, sigma)))
    S, C, N = xp2.shape
    t_e = np.conj(np.tanh(rbm_theta(psi, xp2.reshape(-1, N)).reshape(S, C, -1)))
    we = w_eta[:, None] * np.conj(rho2)
    ye = we[:, :, None] * t_e
    gW = sigma.T.astype(complex) @ ys + np.einsum("sci,scj->ij", xp2.astype(complex), ye)
    gb = ys.sum(0) + ye.sum((0, 1))
    ga = sigma.T.astype(complex) @ w_sig + np.einsum("sci,sc->i", xp2.astype(complex), we)
    return np.concatenate([gb, gW.ravel(), ga])


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from netket_fidelity_b200 import parallel as par
    N, S, n_chains = 4, 64, 8
    psi, phi = RBMParams.random(N, 2, 1, std=0.2), RBMParams.random(N, 2, 2, std=0.2)
    rng = np.random.default_rng(0)
    sigma, eta = rng.choice([-1.0, 1.0], (S, N)), rng.choice([-1.0, 1.0], (S, N))
    U, Ud, cv = oops.Rx(N, 1, 0.3), oops.Rx(N, 1, -0.3), -0.5
    first, n_local = par.shard_chains(n_chains, world, rank)
    assert (first, n_local) == (rank * 4, 4) and par.world_info() == (world, rank)
    per_chain = S // n_chains
    sl = slice(first * per_chain, (first + n_local) * per_chain)        # this rank's sample pairs
    lv, xp2, rho2 = local_log_val(psi, phi, sigma[sl], eta[sl], U, Ud, None)
    A, L = np.exp(lv), local_estimator(lv, cv)
    sums = torch.tensor([L.sum(), (L ** 2).sum(), (np.abs(A) ** 2).sum(), A.real.sum(), len(L), 0, 0, 0],
                        dtype=torch.float64)
    par.allreduce_sum_(sums)                                             # collective 1
    F, var = par.finalize_mean_var(sums)
    acc = _local_pass_B(psi, sigma[sl], eta[sl], A, L, F, cv, xp2, rho2)
    acc_t = torch.view_as_real(torch.tensor(acc)).contiguous()
    par.allreduce_sum_(acc_t)                                            # collective 2
    grad = par.finalize_grad(torch.view_as_complex(acc_t), sums).numpy()
    ref = infidelity_and_grad(psi, phi, sigma, eta, cv, U, Ud, None)
    ok = (abs(F - ref["F"]) < 1e-13 and abs(var - ref["L"].var()) < 1e-13
          and np.allclose(grad, ref["grad"].ravel(), rtol=1e-11, atol=1e-14))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_rank_allreduce_reproduces_single_rank_result():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def _partials(x, l_block, n_blocks):
    """NumPy restatement of nkfb_chain_stats: per row [sum, sum of the first half, sum of the second half, block sums]."""
    rows, cols = x.shape
    half = cols // 2
    blk = x[:, :l_block * n_blocks].reshape(rows, n_blocks, l_block).sum(-1)
    return np.concatenate([x.sum(1, keepdims=True), x[:, :half].sum(1, keepdims=True),
                           x[:, half:2 * half].sum(1, keepdims=True), blk], axis=1)


def _stats_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from netket_fidelity_b200 import parallel as par
    from netket_fidelity_b200.nk.stats import stats_from_partials
    rows_local, cols, l_block = 24, 64, 2
    n_blocks = cols // l_block
    rng = np.random.default_rng(5)
    x = rng.normal(size=(world * rows_local, cols)).cumsum(1) * 0.1 + rng.normal(size=(world * rows_local, 1))
    mine = x[rank * rows_local:(rank + 1) * rows_local]
    part = par.allgather_rows(torch.tensor(_partials(mine, l_block, n_blocks)))
    sums = torch.tensor([mine.sum(), (mine ** 2).sum(), 0, 0, mine.size, 0, 0, 0], dtype=torch.float64)
    par.allreduce_sum_(sums)
    mean, var = par.finalize_mean_var(sums)
    rows = rows_local * world
    st = stats_from_partials(part.numpy(), rows, cols, l_block, n_blocks, var * rows * cols)
    ref = stats_from_partials(_partials(x, l_block, n_blocks), rows, cols, l_block, n_blocks, ((x - x.mean()) ** 2).sum())
    ok = (part.shape == (rows, 3 + n_blocks) and abs(st.error_of_mean - ref.error_of_mean) < 1e-12
          and abs(st.R_hat - ref.R_hat) < 1e-12 and abs(st.tau_corr - ref.tau_corr) < 1e-9
          and abs(mean - x.mean()) < 1e-12 and abs(var - x.var()) < 1e-12 and st.R_hat > 1.0)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_rank_chain_statistics_are_those_of_all_chains():
    """error_of_mean / tau_corr / R_hat with G > 1: the per-chain partial sums of every rank are all-gathered
    (parallel.allgather_rows), so the statistics equal those of one process holding all chains."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_stats_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_shard_chains_errors():
    from netket_fidelity_b200 import parallel as par
    assert par.shard_chains(16, 4, 3) == (12, 4)
    with pytest.raises(ValueError):
        par.shard_chains(10, 4, 0)


def _seed_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from netket_fidelity_b200 import parallel as par
    from netket_fidelity_b200.nk.models import RBM
    local_draw = 1000 + 17 * rank                       # what os.urandom would give: different on every rank
    seed = par.broadcast_seed(local_draw)
    params = RBM(alpha=2, param_dtype=complex).init(seed, 6, torch.device("cpu"))
    flat = torch.cat([torch.view_as_real(params["Dense"]["kernel"].reshape(-1).to(torch.complex128)).reshape(-1)])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    q.put((rank, seed, bool(all(torch.equal(g, gathered[0]) for g in gathered))))
    dist.destroy_process_group()


def test_default_seed_is_broadcast_from_rank_zero():
    """seed=None states (nk/vqs.py) draw their seed on rank 0 only: every rank initialises identical parameters
    (NetKet broadcasts the PRNG key of rank 0); ADVICE r1, medium."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_seed_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, 1000, True), (1, 1000, True)]


def _one_collective_worker(rank, world, port, q):
    """SURVEY B.3 on two gloo ranks: every rank centres pass B with the SAME constant c0 (not the global mean),
    accumulates v2 = sum_s conj O_k(sigma_s) beside it, ONE all-reduce over [sums | v1 | v2], and
    grad = -(v1 + 2 (c0 - F) v2) / n equals the single-rank gradient exactly."""
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from netket_fidelity_b200 import parallel as par
    N, S = 4, 64
    psi, phi = RBMParams.random(N, 2, 1, std=0.2), RBMParams.random(N, 2, 2, std=0.2)
    rng = np.random.default_rng(0)
    sigma, eta = rng.choice([-1.0, 1.0], (S, N)), rng.choice([-1.0, 1.0], (S, N))
    U, Ud, cv, c0 = oops.Rx(N, 1, 0.3), oops.Rx(N, 1, -0.3), -0.5, 0.37
    sl = slice(rank * S // world, (rank + 1) * S // world)
    lv, xp2, rho2 = local_log_val(psi, phi, sigma[sl], eta[sl], U, Ud, None)
    A, L = np.exp(lv), local_estimator(lv, cv)
    v1 = _local_pass_B(psi, sigma[sl], eta[sl], A, L, c0, cv, xp2, rho2)          # centred with c0, not with F
    t = np.conj(np.tanh(rbm_theta(psi, sigma[sl])))
    v2 = np.concatenate([t.sum(0), (sigma[sl].T.astype(complex) @ t).ravel(), sigma[sl].sum(0).astype(complex)])
    buf = torch.cat([torch.tensor([L.sum(), (L ** 2).sum(), 0, 0, len(L), 0, 0, 0], dtype=torch.float64),
                     torch.view_as_real(torch.tensor(v1)).reshape(-1), torch.view_as_real(torch.tensor(v2)).reshape(-1)])
    par.allreduce_sum_(buf)                                                        # the one collective
    P = v1.size
    sums = buf[:8].numpy()
    V1 = torch.view_as_complex(buf[8:8 + 2 * P].reshape(-1, 2)).numpy()
    V2 = torch.view_as_complex(buf[8 + 2 * P:].reshape(-1, 2)).numpy()
    F = sums[0] / sums[4]
    grad = -(V1 + 2 * (c0 - F) * V2) / sums[4]
    ref = infidelity_and_grad(psi, phi, sigma, eta, cv, U, Ud, None)
    q.put((rank, bool(abs(F - ref["F"]) < 1e-13 and np.allclose(grad, ref["grad"].ravel(), rtol=1e-10, atol=1e-13))))
    dist.destroy_process_group()


def test_one_collective_algebra_on_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_one_collective_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
