"""GPU sampler tests: (1) fp64 kernels reproduce, trajectory by trajectory, the float64
restatement of the device algorithm (oracle/sampler.py: metropolis_local_b200);
(2) fp32 and fp64 kernels are statistically consistent with the exact distribution for
N <= 12 (north_star level 3; mirrors test/test_infidelity_operator.py:119-122)."""
import numpy as np
import pytest
import torch

from oracle import operators as oops
from oracle.rbm import RBMParams, make_logfun
from oracle.exact import exact_distribution, states_to_numbers
from oracle.sampler import metropolis_local_b200
from tests._helpers import to_rbmspec, to_opspec, unpack_np

pytestmark = pytest.mark.gpu
LS = (-1.0, 1.0)


@pytest.fixture(scope="module")
def h():
    from netket_fidelity_b200._native import get_handle
    return get_handle(0)


def _net(N, M, seed, std, vb=True):
    p = RBMParams.random(N, -(-M // N), seed, std=std, use_visible_bias=vb)
    return RBMParams(p.kernel[:, :M].copy(), p.bias[:M].copy(), p.visible_bias)


def _run(h, p, cd, n_chains, chain_length, n_discard, sweep, seed, site_mode=0, U=None, chain_offset=0,
         step_offset=0, state=None):
    W32 = (p.N + 31) // 32
    init = state is None
    if state is None:
        state = torch.zeros((n_chains, W32), dtype=torch.int32, device="cuda")
    nacc = torch.zeros((1,), dtype=torch.int64, device="cuda")
    out = h.sample(to_rbmspec(p, cd), state, chain_length, n_discard, sweep, seed, LS, op=to_opspec(U),
                   chain_offset=chain_offset, step_offset=step_offset, site_mode=site_mode, init_random=init,
                   n_accepted=nacc)
    torch.cuda.synchronize()
    return out.cpu().numpy().view(np.uint32), state, int(nacc.item())


# layouts: (N, M) chosen to hit lane-group widths 4, 8, 16, 32 and the K buckets 4 / 8 / 16 / 32
LAYOUTS = [(4, 4), (5, 7), (6, 12), (10, 20), (20, 40), (12, 80), (33, 70), (9, 300), (40, 130), (7, 600)]


@pytest.mark.parametrize("N,M", LAYOUTS)
@pytest.mark.parametrize("site_mode", [0, 1])
def test_fp64_trajectories_match_device_algorithm_restatement(h, N, M, site_mode):
    p = _net(N, M, 100 + N, 0.15)
    n_chains, cl, nd = 7, 3, 2
    ref, fin, nacc_ref = metropolis_local_b200(p, n_chains, cl, nd, N, seed=0x1234567890AB, site_mode=site_mode,
                                               chain_offset=5, step_offset=3)
    got, state, nacc = _run(h, p, torch.complex128, n_chains, cl, nd, N, 0x1234567890AB, site_mode,
                            chain_offset=5, step_offset=3)
    np.testing.assert_array_equal(got, ref)
    np.testing.assert_array_equal(state.cpu().numpy().view(np.uint32), fin)
    assert nacc == nacc_ref


def test_fp64_chains_continue_across_calls_and_shards(h):
    """Two calls with step_offset == one long call; a shard (chain_offset) == the same chains of the full run."""
    N, M = 10, 20
    p = _net(N, M, 9, 0.2)
    full, _, _ = _run(h, p, torch.complex128, 8, 6, 0, N, 77)
    a, st, _ = _run(h, p, torch.complex128, 8, 3, 0, N, 77)
    b, _, _ = _run(h, p, torch.complex128, 8, 3, 0, N, 77, step_offset=3 * N, state=st)
    np.testing.assert_array_equal(np.concatenate([a, b], axis=1), full)
    shard, _, _ = _run(h, p, torch.complex128, 4, 6, 0, N, 77, chain_offset=4)
    np.testing.assert_array_equal(shard, full[4:])


@pytest.mark.parametrize("U", [oops.Rx(5, 1, 0.8), oops.Hadamard(5, 0),
                               oops.IsingPropagator(5, oops.chain_edges(5), 1.0, 1.0, 0.2)])
@pytest.mark.parametrize("M", [5, 40])
def test_fp64_composite_trajectories(h, U, M):
    N = 5
    p = _net(N, M, 31, 0.3)
    ref, fin, nacc_ref = metropolis_local_b200(p, 5, 3, 1, N, seed=99, U=U)
    got, state, nacc = _run(h, p, torch.complex128, 5, 3, 1, N, 99, U=U)
    np.testing.assert_array_equal(got, ref)
    assert nacc == nacc_ref


def _z_scores(counts, p, inflate):
    S = counts.sum()
    return (counts / S - p) / np.sqrt(p * (1 - p) * inflate / S)


@pytest.mark.parametrize("N,M", [(4, 4), (6, 12), (8, 40), (10, 80), (12, 300)])
@pytest.mark.parametrize("cd", [torch.complex64, torch.complex128])
@pytest.mark.parametrize("site_mode", [0, 1])
def test_sampler_matches_exact_distribution(h, N, M, cd, site_mode):
    p = _net(N, M, 7 * N, 0.5 / np.sqrt(M))
    n_chains, cl = 4096, 64
    got, _, nacc = _run(h, p, cd, n_chains, cl, 30, N, 2024 + site_mode, site_mode)
    x = unpack_np(got.reshape(-1, got.shape[-1]), N)
    counts = np.bincount(states_to_numbers(x), minlength=2 ** N)
    pe = exact_distribution(make_logfun(p), N)
    keep = pe * counts.sum() > 50
    z = _z_scores(counts, pe, inflate=6.0)
    assert np.max(np.abs(z[keep])) < 5.0, np.max(np.abs(z[keep]))
    acc_rate = nacc / (n_chains * (cl + 30) * N)
    assert 0.05 < acc_rate <= 1.0


@pytest.mark.parametrize("U", [oops.Rx(6, 2, 0.9), oops.Ry(6, 0, 0.7), oops.Hadamard(6, 5),
                               oops.IsingPropagator(6, oops.chain_edges(6), 1.0, 1.0, 0.1)])
@pytest.mark.parametrize("cd", [torch.complex64, torch.complex128])
def test_composite_sampler_matches_exact_distribution(h, U, cd):
    N, M = 6, 12
    p = _net(N, M, 5, 0.3)
    got, _, _ = _run(h, p, cd, 2048, 64, 30, N, 4242, U=U)
    x = unpack_np(got.reshape(-1, got.shape[-1]), N)
    counts = np.bincount(states_to_numbers(x), minlength=2 ** N)
    pe = exact_distribution(make_logfun(p, U), N)
    keep = pe * counts.sum() > 50
    z = _z_scores(counts, pe, inflate=6.0)
    assert np.max(np.abs(z[keep])) < 5.0, np.max(np.abs(z[keep]))
