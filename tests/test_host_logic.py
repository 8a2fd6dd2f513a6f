"""CPU tests of the host-side mirror of the reference's surface: factory branching and error rules
(infidelity/logic.py:142-187, overlap/operator.py:22-39,61-82, overlap_U/operator.py:25-49), operator
properties (singlequbit_gates.py), Stats finalisation, mpack parameter I/O.  No kernel is launched."""
import math
import os

import numpy as np
import pytest
import torch

import netket_fidelity_b200 as nkf
from netket_fidelity_b200 import nk
from netket_fidelity_b200.nk.vqs import MCState
from netket_fidelity_b200.infidelity.overlap.operator import InfidelityOperatorStandard
from netket_fidelity_b200.infidelity.overlap_U.operator import InfidelityOperatorUPsi
from netket_fidelity_b200.infidelity.logic import Adjoint


def _fake_state(N=3):
    hi = nk.hilbert.Spin(0.5, N)
    st = MCState.__new__(MCState)           # no GPU here: bypass the constructor, the factory only reads these
    st.hilbert = hi
    st.sampler = nk.sampler.MetropolisLocal(hi, n_chains_per_rank=16)
    return hi, st


def test_factory_branches_like_the_reference():
    hi, st = _fake_state()
    assert isinstance(nkf.InfidelityOperator(st), InfidelityOperatorStandard)
    assert nkf.InfidelityOperator(st, cv_coeff=-0.5).cv_coeff == -0.5
    U = nkf.operator.Rx(hi, 0, 0.01)
    op = nkf.InfidelityOperator(st, U=U, is_unitary=True, cv_coeff=-0.5)
    assert isinstance(op, InfidelityOperatorUPsi)
    assert op._U == U and op._U_dagger == nkf.operator.Rx(hi, 0, -0.01)      # U_dagger defaults to U.H
    with pytest.raises(ValueError):      # logic.py:163-175
        nkf.InfidelityOperator(st, U=U)
    with pytest.raises(TypeError):       # logic.py:147-152
        nkf.InfidelityOperator(st, U=U, U_dagger=Adjoint(), is_unitary=True)
    with pytest.raises(TypeError):       # overlap/operator.py:28-32
        nkf.InfidelityOperator(st, cv_coeff=1j)
    with pytest.raises(TypeError):
        nkf.InfidelityOperator(st, cv_coeff=[1.0, 2.0])
    with pytest.raises(TypeError):       # overlap/operator.py:25-26
        nkf.InfidelityOperator("not a state")
    with pytest.raises(ValueError):      # overlap_U/operator.py:28-32
        InfidelityOperatorUPsi(U, st, U_dagger=U.H, is_unitary=False)
    with pytest.raises(TypeError):       # overlap/operator.py:68-72 (test/test_infidelity_operator.py:81-92)
        nkf.InfidelityOperator(st, U=object(), U_dagger=object(), sample_Upsi=True, is_unitary=True)


def test_operator_properties():
    hi = nk.hilbert.Spin(0.5, 4)
    rx, ry, hd = nkf.operator.Rx(hi, 1, 0.3), nkf.operator.Ry(hi, 2, 0.4), nkf.operator.Hadamard(hi, 0)
    assert rx.max_conn_size == ry.max_conn_size == hd.max_conn_size == 2
    assert rx.dtype is complex and ry.dtype is complex and hd.dtype is np.float64
    assert rx.H == nkf.operator.Rx(hi, 1, -0.3)
    assert ry.H == nkf.operator.Ry(hi, 2, -0.8)      # the reference's Ry.H quirk, singlequbit_gates.py:135-136
    assert hd.H == hd and rx.idx == 1 and ry.angle == 0.4
    s = rx._spec()
    assert s.kind == 1 and s.gate_mel[0][1] == -1j * math.sin(0.15)
    isg = nkf.operator.Ising(hi, nk.graph.Chain(4), h=0.7, J=1.3)
    assert isg.max_conn_size == 5 and len(isg.edges) == 4
    assert nk.graph.Grid([10, 10]).n_nodes == 100 and len(nk.graph.Grid([10, 10]).edges()) == 200


def test_hilbert_ordering_matches_reference_goldens():
    # test/test_operator.py:38-41,66-72
    np.testing.assert_array_equal(nk.hilbert.Spin(0.5, 3).numbers_to_states(2), [-1, 1, -1])
    np.testing.assert_array_equal(nk.hilbert.Qubit(3).numbers_to_states(7), [1, 1, 1])
    hi = nk.hilbert.Spin(0.5, 5)
    st = hi.all_states()
    np.testing.assert_array_equal(hi.states_to_numbers(st), np.arange(32))


def test_stats_finalisation_matches_oracle_statistics():
    from netket_fidelity_b200.nk.stats import stats_from_partials
    from oracle.stats import statistics
    rng = np.random.default_rng(0)
    for rows, cols in ((64, 256), (16, 7), (40, 100), (1, 50)):
        data = rng.normal(size=(rows, cols)) + 0.3 * rng.normal(size=(rows, 1))
        l_block = max(1, cols // 32)
        nb = cols // l_block
        half = cols // 2
        part = np.zeros((rows, 3 + nb))
        part[:, 0] = data.sum(1)
        part[:, 1] = data[:, :half].sum(1)
        part[:, 2] = data[:, half:2 * half].sum(1)
        part[:, 3:] = data[:, :nb * l_block].reshape(rows, nb, l_block).sum(2)
        st = stats_from_partials(part, rows, cols, l_block, nb, ((data - data.mean()) ** 2).sum())
        ref = statistics(data)
        for k in ("mean", "variance", "error_of_mean", "tau_corr", "R_hat"):
            a, b = getattr(st, k), ref[k]
            assert (math.isnan(a) and np.isnan(b)) or a == pytest.approx(b, rel=1e-12, abs=1e-14), k


def test_mpack_roundtrip_and_reference_fixture(tmp_path):
    from netket_fidelity_b200.utils.mpack import save_parameters, load_parameters
    params = {"Dense": {"kernel": torch.randn(3, 6, dtype=torch.complex128), "bias": torch.randn(6, dtype=torch.complex128)},
              "visible_bias": torch.randn(3, dtype=torch.complex128)}
    p = str(tmp_path / "a.mpack")
    save_parameters(p, {"params": params})
    back = load_parameters(p)["params"]
    np.testing.assert_array_equal(back["Dense"]["kernel"], params["Dense"]["kernel"].numpy())
    np.testing.assert_array_equal(back["visible_bias"], params["visible_bias"].numpy())
    fix = os.path.join(os.path.dirname(__file__), "golden", "adiabatic_initial_state.mpack")
    t = load_parameters(fix)["params"]      # copied fixture: examples/adiabatic_sweeping/initial_state.mpack
    assert t["Dense"]["kernel"].shape == (10, 10) and t["Dense"]["kernel"].dtype == np.complex128


def test_convergence_stopping():
    from netket_fidelity_b200.nk.callbacks import ConvergenceStopping
    from netket_fidelity_b200.nk.stats import Stats

    class D:
        _loss_name = "Infidelity"
    cb = ConvergenceStopping(1e-3, smoothing_window=2, patience=2)
    cont = [cb(i, {"Infidelity": Stats(mean=v)}, D()) for i, v in enumerate([1.0, 1e-4, 1e-4, 1e-4, 1e-4, 1e-4])]
    assert cont[:4] == [True, True, True, True] and cont[-1] is False


def test_bench_subtracts_the_l2_flushes_from_the_bracket():
    """bench.net_of_flushes: K steps bracketed by one pair of events, K flushes between them timed by their own pairs;
    the step time is the bracket minus K x the median flush (a host hiccup inside one flush interval is not subtracted),
    the raw bracket and the flush are reported beside it."""
    import bench

    class Ev:
        def __init__(self, t):
            self.t = t

        def elapsed_time(self, other):
            return other.t - self.t

    class Ctx:
        def max_over_ranks(self, x):
            return x

    flushes = [(Ev(0.0), Ev(0.05)), (Ev(10.0), Ev(10.05)), (Ev(20.0), Ev(20.95)), (Ev(30.0), Ev(30.05)), (Ev(40.0), Ev(40.05))]
    net, bracket, flush = bench.net_of_flushes(Ctx(), Ev(0.0), Ev(50.0), flushes)
    assert abs(flush - 0.05) < 1e-12 and bracket == 50.0
    assert abs(net - (50.0 - 5 * 0.05)) < 1e-12
    net0, bracket0, flush0 = bench.net_of_flushes(Ctx(), Ev(0.0), Ev(7.0), [])
    assert (net0, bracket0, flush0) == (7.0, 7.0, 0.0)
