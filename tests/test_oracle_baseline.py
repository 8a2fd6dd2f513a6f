"""The numba CPU baseline (oracle/baseline_numba.py: prange over chains, full forward per Metropolis step) walks
the same trajectories as the NumPy restatement of NetKet's MetropolisLocal (oracle/sampler.py) when both consume
the same random numbers -- plain targets and gate-composite targets log(U phi)."""
import numpy as np
import pytest

from oracle import operators as oops
from oracle.rbm import RBMParams, make_logfun
from oracle.sampler import metropolis_local_netket
from oracle import baseline_numba as bn

pytestmark = pytest.mark.skipif(not bn.HAVE_NUMBA, reason="numba not importable")


@pytest.mark.parametrize("gate", [None, "rx", "h"])
def test_numba_sampler_equals_numpy_sampler(gate):
    N = 6
    p = RBMParams.random(N, 2, 11, std=0.3)
    U = {None: None, "rx": oops.Rx(N, 2, 0.4), "h": oops.Hadamard(N, 1)}[gate]
    a, fa = metropolis_local_netket(make_logfun(p, U), N, 5, 4, 2, rng=np.random.default_rng(3))
    b, fb = bn.metropolis_local_numba(p, N, 5, 4, 2, rng=np.random.default_rng(3), U=U)
    assert np.array_equal(a, b) and np.array_equal(fa, fb)


def test_cpu_step_runs_and_reports_threads():
    from oracle.baseline import make_workload
    w = make_workload("C1")
    dt, S, res, info = bn.cpu_step(w, 4, 3)
    assert S == 12 and np.isfinite(res["I"]) and info["threads"] >= 1 and info["engine"].startswith("numba")
