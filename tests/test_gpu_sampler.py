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
@pytest.mark.parametrize("sampler", [1, 2])
def test_fp64_composite_trajectories(h, algo, U, M, sampler):
    """sampler 1: complex log-cosh kernel; 2: for Ising-type operators the multiplicative kernel
    (sampler_composite_ratio.cuh), which must accept exactly the same proposals in float64."""
    algo(sampler)
    N = 5
    p = _net(N, M, 31, 0.3)
    ref, fin, nacc_ref = metropolis_local_b200(p, 5, 3, 1, N, seed=99, U=U)
    got, state, nacc = _run(h, p, torch.complex128, 5, 3, 1, N, 99, U=U)
    np.testing.assert_array_equal(got, ref)
    assert nacc == nacc_ref


@pytest.fixture
def algo(h):
    """Selects the plain-target kernel for one test and restores the automatic choice afterwards."""
    def set_(a):
        h.set_sampler_algo(a)
    yield set_
    h.set_sampler_algo(0)


@pytest.mark.parametrize("N,M", LAYOUTS)
@pytest.mark.parametrize("site_mode", [0, 1])
def test_fp64_multiplicative_sampler_follows_the_same_trajectories(h, algo, N, M, site_mode):
    """The multiplicative kernel (z = exp(-2 theta), sampler_ratio.cuh) consumes the same random numbers and
    accepts the same proposals as the additive one; in float64 the two log-ratios differ by ~1e-13, so the
    trajectories coincide with the float64 restatement of the additive algorithm."""
    algo(2)
    p = _net(N, M, 100 + N, 0.15)
    n_chains, cl, nd = 7, 3, 2
    ref, fin, nacc_ref = metropolis_local_b200(p, n_chains, cl, nd, N, seed=0x1234567890AB, site_mode=site_mode,
                                               chain_offset=5, step_offset=3)
    got, state, nacc = _run(h, p, torch.complex128, n_chains, cl, nd, N, 0x1234567890AB, site_mode,
                            chain_offset=5, step_offset=3)
    np.testing.assert_array_equal(got, ref)
    np.testing.assert_array_equal(state.cpu().numpy().view(np.uint32), fin)
    assert nacc == nacc_ref


@pytest.mark.parametrize("std", [0.05, 0.3, 0.7, 2.5])
def test_fp64_multiplicative_sampler_range_windows(h, algo, std):
    """X = max_j (|Re b_j| + sum_i |Re W_ij|) grows with std through the three windows of the plain-target
    kernels (one log per 4 pairs: X <= 2.67; one per pair: X <= 10.7; additive beyond): whichever kernel
    takes the call, the trajectories are those of the restatement."""
    algo(2)
    N, M = 10, 20
    p = _net(N, M, 77, std)
    X = np.max(np.abs(p.bias.real) + np.abs(p.kernel.real).sum(0))
    assert {0.05: X < 2.0, 0.3: 2.8 < X < 10.0, 0.7: 2.8 < X < 10.0, 2.5: X > 11.5}[std], X
    ref, fin, nacc_ref = metropolis_local_b200(p, 9, 4, 2, N, seed=4711)
    got, state, nacc = _run(h, p, torch.complex128, 9, 4, 2, N, 4711)
    np.testing.assert_array_equal(got, ref)
    assert nacc == nacc_ref


def test_multiplicative_chains_continue_across_calls_and_shards(h, algo):
    algo(2)
    N, M = 10, 20
    p = _net(N, M, 9, 0.2)
    full, _, _ = _run(h, p, torch.complex128, 8, 6, 0, N, 77)
    a, st, _ = _run(h, p, torch.complex128, 8, 3, 0, N, 77)
    b, _, _ = _run(h, p, torch.complex128, 8, 3, 0, N, 77, step_offset=3 * N, state=st)
    np.testing.assert_array_equal(np.concatenate([a, b], axis=1), full)
    shard, _, _ = _run(h, p, torch.complex128, 4, 6, 0, N, 77, chain_offset=4)
    np.testing.assert_array_equal(shard, full[4:])


def test_fp32_samplers_agree_on_most_trajectories(h, algo):
    """Single precision: the additive and the multiplicative kernel share random numbers and differ only in the
    rounding of the log-ratio (~1e-5), so all but a few chains emit identical samples."""
    N, M = 16, 64
    p = _net(N, M, 3, 0.1)
    algo(1)
    a, _, na = _run(h, p, torch.complex64, 256, 8, 2, N, 5)
    algo(2)
    b, _, nb = _run(h, p, torch.complex64, 256, 8, 2, N, 5)
    same = np.all(a.reshape(256, -1) == b.reshape(256, -1), axis=1).mean()
    assert same > 0.9, same
    assert abs(na - nb) < 0.01 * na


@pytest.fixture
def kernel_path(monkeypatch):
    """Selects which multiplicative kernel a 32-lane shape dispatches to: 'cta' (CTA-synchronous shared rows,
    site_mode 0 only), 'pipe' (one chain per warp, pipelined) or 'generic' (lane-group kernel)."""
    def set_(path):
        monkeypatch.delenv("NKFB_RATIO_NOCTA", raising=False)
        monkeypatch.delenv("NKFB_RATIO_NOPIPE", raising=False)
        monkeypatch.delenv("NKFB_RATIO_FORCECTA", raising=False)
        if path == "cta":
            monkeypatch.setenv("NKFB_RATIO_FORCECTA", "1")  # small launches default to the pipelined kernel
        if path in ("pipe", "generic"):
            monkeypatch.setenv("NKFB_RATIO_NOCTA", "1")
        if path == "generic":
            monkeypatch.setenv("NKFB_RATIO_NOPIPE", "1")
    return set_


@pytest.mark.parametrize("N,M", [(9, 300), (40, 130), (100, 400), (7, 500), (150, 200)])
@pytest.mark.parametrize("path", ["cta", "pipe", "generic"])
def test_fp64_multiplicative_kernels_agree_with_restatement(h, algo, kernel_path, N, M, path):
    """The three forms of the multiplicative sampler (sampler_ratio.cuh) on 32-lane shapes, float64: each one
    reproduces the trajectories of the float64 restatement (shared site sequence, chains continued from a
    non-zero step offset so that chunk / RNG-batch phases are exercised)."""
    algo(2)
    kernel_path(path)
    p = _net(N, M, 50 + N, 0.3 / np.sqrt(N))
    n_chains, cl, nd = 5, 3, 1
    ref, fin, nacc_ref = metropolis_local_b200(p, n_chains, cl, nd, N, seed=31337, site_mode=0, chain_offset=2,
                                               step_offset=2 * N + 1)
    got, state, nacc = _run(h, p, torch.complex128, n_chains, cl, nd, N, 31337, 0, chain_offset=2,
                            step_offset=2 * N + 1)
    np.testing.assert_array_equal(got, ref)
    np.testing.assert_array_equal(state.cpu().numpy().view(np.uint32), fin)
    assert nacc == nacc_ref


@pytest.mark.parametrize("path", ["cta", "pipe", "generic"])
def test_fp32_multiplicative_kernels_north_star_shape(h, algo, kernel_path, path):
    """N = 100, M = 400 (the north-star shape), single precision, many more chains than one CTA holds: the three
    kernels share random numbers, so all but a few per cent of the chains emit the samples of the additive kernel,
    and the acceptance rates agree."""
    N, M = 100, 400
    p = _net(N, M, 11, 0.01)
    algo(1)
    a, _, na = _run(h, p, torch.complex64, 700, 4, 1, N, 5)
    algo(2)
    kernel_path(path)
    b, _, nb = _run(h, p, torch.complex64, 700, 4, 1, N, 5)
    same = np.all(a.reshape(700, -1) == b.reshape(700, -1), axis=1).mean()
    assert same > 0.9, same
    assert abs(na - nb) < 0.005 * na


def test_tail_of_a_large_launch_on_the_one_chain_per_warp_kernel(h, algo, monkeypatch):
    """Large launches of the CTA-synchronous kernel hand the chains beyond their last full wave of CTAs to the
    pipelined kernel when those are few (sampler_ratio_inst.cu, NKFB_OPT_SAMPLER_SHARE).  float64: the split launch
    emits exactly the samples, final configurations and acceptance count of the unsplit one; float32 (north-star
    shape, two launches sharing the GPU): all but a few chains agree and every chain of the tail is written."""
    algo(2)
    for v in ("NKFB_RATIO_NOCTA", "NKFB_RATIO_NOPIPE", "NKFB_RATIO_FORCECTA", "NKFB_RATIO_NOSPLIT"):
        monkeypatch.delenv(v, raising=False)
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    # float64: one chain per warp in the CTA kernel (launches of >= 40 chains per SM); head = 2 waves of 16-warp
    # CTAs, tail = 8 sm + 100 chains
    N, M = 40, 130
    p = _net(N, M, 90, 0.3 / np.sqrt(N))
    n_chains = 40 * sm + 100
    got, st, nacc = _run(h, p, torch.complex128, n_chains, 2, 1, N, 4242, 0, chain_offset=3, step_offset=N + 2)
    monkeypatch.setenv("NKFB_RATIO_NOSPLIT", "1")
    ref, st_ref, nacc_ref = _run(h, p, torch.complex128, n_chains, 2, 1, N, 4242, 0, chain_offset=3,
                                 step_offset=N + 2)
    monkeypatch.delenv("NKFB_RATIO_NOSPLIT")
    np.testing.assert_array_equal(got, ref)
    np.testing.assert_array_equal(st.cpu().numpy(), st_ref.cpu().numpy())
    assert nacc == nacc_ref
    # the split really happened: head + tail instead of one kernel
    l0 = h.launches
    _run(h, p, torch.complex128, n_chains, 1, 0, N, 1, 0)
    l_split = h.launches - l0
    monkeypatch.setenv("NKFB_RATIO_NOSPLIT", "1")
    l0 = h.launches
    _run(h, p, torch.complex128, n_chains, 1, 0, N, 1, 0)
    assert l_split > h.launches - l0          # one more kernel per range window that ran split
    monkeypatch.delenv("NKFB_RATIO_NOSPLIT")
    # float32, N = 100, M = 400: two chains per warp, share = 2 -> head = 3 x (sm / 2) CTAs, tail = 700 chains
    N, M = 100, 400
    p = _net(N, M, 11, 0.01)
    n_chains = 3 * (sm // 2) * 32 + 700
    h.set_sampler_share(2)
    try:
        a, _, na = _run(h, p, torch.complex64, n_chains, 2, 1, N, 5)
    finally:
        h.set_sampler_share(1)
    monkeypatch.setenv("NKFB_RATIO_NOSPLIT", "1")
    b, _, nb = _run(h, p, torch.complex64, n_chains, 2, 1, N, 5)
    same = np.all(a.reshape(n_chains, -1) == b.reshape(n_chains, -1), axis=1)
    assert same.mean() > 0.97 and same[-700:].mean() > 0.9, (same.mean(), same[-700:].mean())
    assert abs(na - nb) < 0.005 * na


def _z_scores(counts, p, inflate):
    S = counts.sum()
    return (counts / S - p) / np.sqrt(p * (1 - p) * inflate / S)


@pytest.mark.parametrize("N,M", [(4, 4), (6, 12), (8, 40), (10, 80), (12, 300)])
@pytest.mark.parametrize("cd", [torch.complex64, torch.complex128])
@pytest.mark.parametrize("site_mode", [0, 1])
@pytest.mark.parametrize("sampler", [1, 2])
def test_sampler_matches_exact_distribution(h, algo, N, M, cd, site_mode, sampler):
    algo(sampler)
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


@pytest.mark.parametrize("B", [0.0, 4.0, 14.0])
@pytest.mark.parametrize("sampler", [1, 2])
@pytest.mark.parametrize("M", [24, 100])
def test_fp32_samplers_saturated_hidden_units(h, algo, B, sampler, M):
    """Single precision with hidden biases +-B: X = max_j (|Re b_j| + sum_i |Re W_ij|) = 1.6 / 5.6 / 15.6 falls
    in the first (one log per 4 pairs), second (one log per pair) and third (additive kernel) range window of
    the multiplicative sampler; z = exp(-2 theta) spans e^{+-2X}.  The distribution itself stays broad (large
    biases saturate cosh), so the chains mix within a sweep and the z-scores are meaningful."""
    algo(sampler)
    N = 10                                 # M = 24: lane-group kernel; M = 100: 32-lane (pipelined) kernels
    p = _net(N, M, 21, 0.15 * (24.0 / M) ** 0.5)
    p = RBMParams(p.kernel, p.bias + B * np.where(np.arange(M) % 2 == 0, 1.0, -1.0), p.visible_bias)
    X = np.max(np.abs(p.bias.real) + np.abs(p.kernel.real).sum(0))
    assert {0.0: X < 2.6, 4.0: 2.8 < X < 10.6, 14.0: X > 10.8}[B], X
    got, _, _ = _run(h, p, torch.complex64, 4096, 64, 30, N, 99)
    x = unpack_np(got.reshape(-1, got.shape[-1]), N)
    counts = np.bincount(states_to_numbers(x), minlength=2 ** N)
    pe = exact_distribution(make_logfun(p), N)
    keep = pe * counts.sum() > 50
    z = _z_scores(counts, pe, inflate=6.0)
    assert np.max(np.abs(z[keep])) < 5.0, np.max(np.abs(z[keep]))


@pytest.mark.parametrize("U", [oops.Rx(6, 2, 0.9), oops.Ry(6, 0, 0.7), oops.Hadamard(6, 5),
                               oops.IsingPropagator(6, oops.chain_edges(6), 1.0, 1.0, 0.1)])
@pytest.mark.parametrize("cd", [torch.complex64, torch.complex128])
@pytest.mark.parametrize("sampler", [1, 2])
def test_composite_sampler_matches_exact_distribution(h, algo, U, cd, sampler):
    algo(sampler)
    N, M = 6, 12
    p = _net(N, M, 5, 0.3)
    got, _, _ = _run(h, p, cd, 2048, 64, 30, N, 4242, U=U)
    x = unpack_np(got.reshape(-1, got.shape[-1]), N)
    counts = np.bincount(states_to_numbers(x), minlength=2 ** N)
    pe = exact_distribution(make_logfun(p, U), N)
    keep = pe * counts.sum() > 50
    z = _z_scores(counts, pe, inflate=6.0)
    assert np.max(np.abs(z[keep])) < 5.0, np.max(np.abs(z[keep]))


@pytest.mark.parametrize("N,M,std", [(12, 30, 0.15), (40, 80, 0.02), (33, 70, 0.05)])
def test_fp64_ising_composite_kernels_agree_on_longer_runs(h, algo, N, M, std):
    """Ising-type propagator targets at the C3 shape and beyond one 32-lane round of connected configurations:
    the multiplicative kernel and the log-cosh kernel emit identical float64 trajectories (chains continued over
    two calls, non-zero chain offset)."""
    U = oops.IsingPropagator(N, oops.chain_edges(N), 1.0, 1.0, 0.01)
    p = _net(N, M, 77, std)
    out = {}
    for sampler in (1, 2):
        algo(sampler)
        a, st, na = _run(h, p, torch.complex128, 6, 3, 1, N, 123, U=U, chain_offset=3)
        b, st, nb = _run(h, p, torch.complex128, 6, 2, 0, N, 123, U=U, chain_offset=3, step_offset=4 * N, state=st)
        out[sampler] = (a, b, na + nb)
    np.testing.assert_array_equal(out[1][0], out[2][0])
    np.testing.assert_array_equal(out[1][1], out[2][1])
    assert out[1][2] == out[2][2]


def test_fp32_ising_composite_saturated_units(h, algo):
    """Parameters beyond the range of the multiplicative composite kernel (X > 3): the log-cosh kernel behind it
    takes over; the distribution is still exact."""
    N, M = 6, 12
    p = _net(N, M, 5, 0.2)
    p = RBMParams(p.kernel, p.bias + 6.0 * np.where(np.arange(M) % 2 == 0, 1.0, -1.0), p.visible_bias)
    U = oops.IsingPropagator(N, oops.chain_edges(N), 1.0, 1.0, 0.1)
    algo(2)
    got, _, _ = _run(h, p, torch.complex64, 2048, 64, 30, N, 4242, U=U)
    x = unpack_np(got.reshape(-1, got.shape[-1]), N)
    counts = np.bincount(states_to_numbers(x), minlength=2 ** N)
    pe = exact_distribution(make_logfun(p, U), N)
    keep = pe * counts.sum() > 50
    z = _z_scores(counts, pe, inflate=6.0)
    assert np.max(np.abs(z[keep])) < 5.0, np.max(np.abs(z[keep]))


def test_long_calls_fall_back_from_the_site_table(h, algo):
    """More Metropolis steps in one call than the site table of the CTA-synchronous kernel holds (65536): the
    dispatcher uses the pipelined kernel; float64 samples equal those of the additive kernel."""
    N, M = 40, 130
    p = _net(N, M, 8, 0.05)
    n_sweeps = 1700                      # 68000 steps
    algo(1)
    a, sa, na = _run(h, p, torch.complex128, 3, n_sweeps, 0, N, 11)
    algo(2)
    b, sb, nb = _run(h, p, torch.complex128, 3, n_sweeps, 0, N, 11)
    np.testing.assert_array_equal(a, b)
    assert na == nb


@pytest.fixture
def hw_kernel(monkeypatch):
    """Half-warp lean kernel (sampler_lean_hw.cuh) on / off; small launches are forced onto the CTA-synchronous path."""
    def set_(on):
        for v in ("NKFB_RATIO_NOCTA", "NKFB_RATIO_NOPIPE", "NKFB_RATIO_NOSPLIT", "NKFB_NOHW"):
            monkeypatch.delenv(v, raising=False)
        monkeypatch.setenv("NKFB_RATIO_FORCECTA", "1")
        if not on:
            monkeypatch.setenv("NKFB_NOHW", "1")
    return set_


@pytest.mark.parametrize("N,M,B", [(100, 400, 0.0), (100, 400, 5.0), (64, 330, 0.0), (37, 384, 0.0), (100, 448, 0.0),
                                   (128, 470, 0.0), (40, 512, 0.0), (40, 512, 5.0), (150, 400, 0.0), (7, 416, 0.0)])
def test_half_warp_kernel_agrees_with_the_32_lane_lean_kernel(h, algo, hw_kernel, N, M, B):
    """sampler_lean_hw.cuh (a chain on 16 lanes, both chains of a warp in one instruction stream) against the 32-lane
    lean kernel on every instantiated width (11..16 pairs per lane), both range windows (B = 5: one logarithm per
    pair), a chain count that leaves a half-empty warp and a partial CTA, a non-zero step offset (partial first chunk,
    RNG batch phase) and chain offset, chains continued over a second call: same Philox streams and decisions up to
    single-precision rounding of the log-ratio, so nearly all chains emit identical samples and the acceptance agrees;
    the launch counter proves that the half-warp kernel's table preparation ran."""
    algo(2)
    p = _net(N, M, 300 + M, 0.01)
    if B:
        p = RBMParams(p.kernel, p.bias + B * np.where(np.arange(M) % 2 == 0, 1.0, -1.0), p.visible_bias)
    n_chains = 333
    out = {}
    for on in (False, True):
        hw_kernel(on)
        l0 = h.launches
        a, st, na = _run(h, p, torch.complex64, n_chains, 3, 1, N, 77, chain_offset=5, step_offset=N + 3)
        out[on, "launches"] = h.launches - l0
        b, st, nb = _run(h, p, torch.complex64, n_chains, 2, 0, N, 77, chain_offset=5, step_offset=5 * N + 3, state=st)
        out[on] = (np.concatenate([a, b], axis=1), st.cpu().numpy(), na + nb)
    assert out[True, "launches"] == out[False, "launches"] + 1      # the 16-lane copy of the q tables
    same = np.all(out[True][0].reshape(n_chains, -1) == out[False][0].reshape(n_chains, -1), axis=1).mean()
    assert same > 0.9, same
    assert abs(out[True][2] - out[False][2]) < 0.005 * out[False][2]
    # final configurations are the last samples of the second call
    np.testing.assert_array_equal(out[True][1].view(np.uint32), out[True][0][:, -1])


@pytest.mark.parametrize("M", [400, 512])
def test_half_warp_kernel_matches_exact_distribution(h, algo, hw_kernel, M):
    """N = 10 spins with the north-star number of hidden units: the half-warp kernel samples |psi|^2 exactly (z-scores
    against full enumeration, as test_sampler_matches_exact_distribution)."""
    algo(2)
    hw_kernel(True)
    N = 10
    p = _net(N, M, 7 * N, 0.5 / np.sqrt(M))
    n_chains, cl = 4096, 64
    l0 = h.launches
    got, _, nacc = _run(h, p, torch.complex64, n_chains, cl, 30, N, 2031)
    x = unpack_np(got.reshape(-1, got.shape[-1]), N)
    counts = np.bincount(states_to_numbers(x), minlength=2 ** N)
    pe = exact_distribution(make_logfun(p), N)
    keep = pe * counts.sum() > 50
    z = _z_scores(counts, pe, inflate=6.0)
    assert np.max(np.abs(z[keep])) < 5.0, np.max(np.abs(z[keep]))
    assert 0.05 < nacc / (n_chains * (cl + 30) * N) <= 1.0
