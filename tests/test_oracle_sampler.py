"""Pins the oracle's RNG (Random123 known-answer vectors) and checks both oracle samplers
against the exact distribution (mirrors the statistical check of
test/test_infidelity_operator.py:119-122)."""
import numpy as np

from oracle.philox import philox4x32_10
from oracle.rbm import RBMParams, make_logfun
from oracle.exact import exact_distribution, states_to_numbers
from oracle.sampler import metropolis_local_netket, metropolis_local_b200
from oracle import operators as oops
from tests._helpers import unpack_np


def test_philox_known_answers():
    # Random123 kat_vectors: philox4x32-10
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, out in kat:
        got = philox4x32_10(np.array(ctr, dtype=np.uint32), key)
        assert tuple(int(v) for v in got) == out


def _z_scores(counts, p, inflate):
    S = counts.sum()
    return (counts / S - p) / np.sqrt(p * (1 - p) * inflate / S)


def test_netket_style_sampler_matches_exact_distribution():
    N = 5
    p = RBMParams.random(N, 2, 3, std=0.4)
    lf = make_logfun(p)
    smp, _ = metropolis_local_netket(lf, N, 256, 200, 20, rng=np.random.default_rng(1))
    counts = np.bincount(states_to_numbers(smp.reshape(-1, N)), minlength=2 ** N)
    z = _z_scores(counts, exact_distribution(lf, N), inflate=4.0)
    assert np.max(np.abs(z)) < 5.0


def test_b200_algorithm_matches_exact_distribution_plain_and_composite():
    N = 4
    p = RBMParams.random(N, 2, 5, std=0.4)
    for U in (None, oops.Rx(N, 1, 0.8), oops.IsingPropagator(N, oops.chain_edges(N), 1.0, 1.0, 0.2)):
        for site_mode in (0, 1):
            w, _, nacc = metropolis_local_b200(p, 48, 60, 10, N, seed=11 + site_mode, site_mode=site_mode, U=U)
            x = unpack_np(w.reshape(-1, 1), N)
            counts = np.bincount(states_to_numbers(x), minlength=2 ** N)
            z = _z_scores(counts, exact_distribution(make_logfun(p, U), N), inflate=4.0)
            assert np.max(np.abs(z)) < 5.0, (U, site_mode, z)
            assert nacc > 0
