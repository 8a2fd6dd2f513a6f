"""GPU tests of the public surface, mirroring the reference's own tests:
test/test_infidelity_operator.py:74-127 (MC mean vs exact within 5 sigma; expect == expect_and_grad;
gradient vs central differences of the exact infidelity, rel 0.5), the README example
(README.md:49-69), get_conn_padded (test/test_operator.py), plus driver / PTVMC / Renyi2 / Stats."""
import math

import numpy as np
import pytest
import torch

import netket_fidelity_b200 as nkf
from netket_fidelity_b200 import nk
from oracle import operators as oops
from oracle.rbm import RBMParams
from oracle.exact import infidelity_exact, infidelity_exact_grad_fd

pytestmark = pytest.mark.gpu


def _to_oracle(vs):
    p = vs.parameters
    f = lambda t: None if t is None else t.detach().cpu().numpy()
    return RBMParams(f(p["Dense"]["kernel"]), f(p["Dense"].get("bias")), f(p.get("visible_bias")))


def _flat(tree):
    parts = []
    if "bias" in tree["Dense"]:
        parts.append(tree["Dense"]["bias"].reshape(-1))
    parts.append(tree["Dense"]["kernel"].reshape(-1))
    if "visible_bias" in tree:
        parts.append(tree["visible_bias"].reshape(-1))
    return torch.cat(parts).cpu().numpy()


def test_get_conn_padded_surface():
    hi = nk.hilbert.Spin(0.5, 3)
    sig = hi.numbers_to_states([2, 7])
    xp, mels = nkf.operator.Rx(hi, 0, np.pi / 2).get_conn_padded(sig)
    np.testing.assert_array_equal(xp, [[[-1, 1, -1], [1, 1, -1]], [[1, 1, 1], [-1, 1, 1]]])
    np.testing.assert_allclose(mels, [[0.70710678, -0.70710678j]] * 2, rtol=1e-7)
    xp, mels = nkf.operator.Hadamard(hi, 0).get_conn_padded(torch.tensor(sig, device="cuda").reshape(2, 1, 3))
    assert xp.shape == (2, 1, 2, 3) and mels.shape == (2, 1, 2) and mels.dtype == torch.float64
    np.testing.assert_allclose(mels.cpu().numpy()[:, 0], [[0.70710678, 0.70710678], [-0.70710678, 0.70710678]], rtol=1e-7)
    # dense == to_local_operator() dense of the reference (test/test_operator.py:19-28), via Pauli matrices
    d = nkf.operator.Rx(nk.hilbert.Spin(0.5, 2), 1, 0.23).to_dense()
    sx = np.array([[0, 1], [1, 0]])
    np.testing.assert_allclose(d, np.cos(0.115) * np.eye(4) - 1j * np.sin(0.115) * np.kron(np.eye(2), sx), atol=1e-14)
    # test/test_inverserot.py
    rx = nkf.operator.Rx(nk.hilbert.Spin(0.5, 4), 0, 0.3)
    np.testing.assert_allclose(rx.H.to_dense() @ rx.to_dense(), np.eye(16), atol=1e-12)
    isg = nkf.operator.Ising(nk.hilbert.Spin(0.5, 4), nk.graph.Chain(4), h=0.7, J=1.3)
    ref = oops.Ising(4, oops.chain_edges(4), 0.7, 1.3).to_dense()
    np.testing.assert_allclose(isg.to_dense(), ref.real, atol=1e-14)


@pytest.mark.parametrize("sample_Upsi", [False, True])
@pytest.mark.parametrize("which", ["Identity", "Rx"])
@pytest.mark.parametrize("pdtype", [float, complex])
def test_MCState_matches_exact_like_the_reference_test(sample_Upsi, which, pdtype):
    """test/test_infidelity_operator.py:74-127 (N=3, RBM alpha=1, MetropolisLocal 16 chains, cv -1/2)."""
    N = 3
    hi = nk.hilbert.Spin(0.5, N)
    sa = nk.sampler.MetropolisLocal(hilbert=hi, n_chains_per_rank=1024)
    ma = nk.models.RBM(alpha=1, param_dtype=pdtype, stddev=0.3)
    vs_t = nk.vqs.MCState(sampler=sa, model=ma, n_samples=2 ** 20, n_discard_per_chain=100, seed=11)
    vs = nk.vqs.MCState(sampler=sa, model=ma, n_samples=2 ** 20, n_discard_per_chain=100, seed=22)
    U = None if which == "Identity" else nkf.operator.Rx(hi, 0, 0.4)
    oU = None if U is None else oops.Rx(N, 0, 0.4)
    psi, phi = _to_oracle(vs), _to_oracle(vs_t)
    I_exact = infidelity_exact(psi, phi, oU)
    g_exact = infidelity_exact_grad_fd(psi, phi, oU, 1e-5)
    I_op = nkf.InfidelityOperator(vs_t, U=U, U_dagger=None if U is None else U.H, sample_Upsi=sample_Upsi,
                                  cv_coeff=-0.5, is_unitary=True)
    I_stat1 = vs.expect(I_op)
    I_stat, I_grad = vs.expect_and_grad(I_op)
    err = 5 * I_stat1.error_of_mean
    assert err > 0 and not math.isnan(err)
    assert abs(I_exact - I_stat1.mean) <= err                       # :122
    assert I_stat1.mean == pytest.approx(I_stat.mean, abs=1e-5)     # :124
    assert I_stat1.variance == pytest.approx(I_stat.variance, abs=1e-5)
    g = _flat(I_grad)
    assert g.dtype == (np.complex128 if pdtype is complex else np.float64)
    # :127 same_derivatives(rel_eps=0.5): with 2^20 samples we can hold a much tighter bound
    np.testing.assert_allclose(g.real, g_exact.real, rtol=0.5, atol=3e-3)
    if pdtype is complex:
        np.testing.assert_allclose(g.imag, g_exact.imag, rtol=0.5, atol=3e-3)


@pytest.mark.parametrize("pdtype", [complex, float])
def test_MCState_with_an_Ising_type_U_and_U_dagger_in_the_overlap_U_estimator(pdtype):
    """overlap_U/expect.py:101-120 takes any DiscreteJaxOperator pair: (N+1)-connected U and U^dagger
    (`is_unitary=True`).  The estimator and the gradient on the state's own samples equal the oracle's on the
    same samples; expect == expect_and_grad."""
    from oracle.infidelity import infidelity_and_grad, mode_slots
    N = 6
    hi = nk.hilbert.Spin(0.5, N)
    sa = nk.sampler.MetropolisLocal(hilbert=hi, n_chains_per_rank=64)
    ma = nk.models.RBM(alpha=2, param_dtype=pdtype, stddev=0.2)
    vs_t = nk.vqs.MCState(sampler=sa, model=ma, n_samples=2048, n_discard_per_chain=5, seed=3)
    vs = nk.vqs.MCState(sampler=sa, model=ma, n_samples=2048, n_discard_per_chain=5, seed=4)
    U = nkf.operator.Ising(hi, nk.graph.Chain(N), h=0.8, J=1.1).propagator(0.05)
    oU = oops.IsingPropagator(N, oops.chain_edges(N), 0.8, 1.1, 0.05)
    I_op = nkf.InfidelityOperator(vs_t, U=U, U_dagger=U.H, cv_coeff=-0.5, is_unitary=True)
    st1 = vs.expect(I_op)
    st, grad = vs.expect_and_grad(I_op)
    assert st1.mean == pytest.approx(st.mean, abs=1e-12)
    sigma = vs.samples.reshape(-1, N).cpu().numpy()
    eta = vs_t.samples.reshape(-1, N).cpu().numpy()
    U1, U2, U4 = mode_slots("upsi", oU, oU.H)
    ref = infidelity_and_grad(_to_oracle(vs), _to_oracle(vs_t), sigma, eta, -0.5, U1, U2, U4)
    assert st.mean == pytest.approx(1.0 - ref["F"], rel=1e-10, abs=1e-12)
    g, gref = _flat(grad), ref["grad"].ravel()
    if pdtype is float:
        gref = gref.real
    np.testing.assert_allclose(g, gref, rtol=1e-8, atol=1e-11 * max(1.0, np.abs(gref).max()))


def test_hilbert_mismatch_and_unsupported_operator_raise():
    hi3, hi4 = nk.hilbert.Spin(0.5, 3), nk.hilbert.Spin(0.5, 4)
    ma = nk.models.RBM(alpha=1, param_dtype=complex)
    a = nk.vqs.MCState(nk.sampler.MetropolisLocal(hi3), ma, n_samples=64)
    b = nk.vqs.MCState(nk.sampler.MetropolisLocal(hi4), ma, n_samples=64)
    with pytest.raises(TypeError, match="Hilbert spaces should match"):     # overlap/expect.py:16-17
        a.expect(nkf.InfidelityOperator(b))
    with pytest.raises(TypeError):
        a.expect(object())


def test_readme_example_runs_and_lowers_the_infidelity():
    """README.md:49-69, verbatim modulo the import line."""
    hi = nk.hilbert.Spin(0.5, 4)
    sampler = nk.sampler.MetropolisLocal(hilbert=hi, n_chains_per_rank=16)
    model = nk.models.RBM(alpha=1, param_dtype=complex, use_visible_bias=False)
    phi = nk.vqs.MCState(sampler=sampler, model=model, n_samples=100, seed=1)
    psi = nk.vqs.MCState(sampler=sampler, model=model, n_samples=100, seed=2)
    assert psi.n_samples == 112 and psi.chain_length == 7            # 100 rounded up to 16 chains x 7
    U = nkf.operator.Hadamard(hi, 0)
    optimizer = nk.optimizer.Adam(learning_rate=0.01)
    te = nkf.driver.InfidelityOptimizer(phi, optimizer, U=U, U_dagger=U, variational_state=psi, is_unitary=True,
                                        cv_coeff=-1 / 2)
    oU = oops.Hadamard(4, 0)
    I0 = infidelity_exact(_to_oracle(psi), _to_oracle(phi), oU)
    log = nk.logging.RuntimeLog()
    te.run(n_iter=300, out=log, show_progress=False)
    I1 = infidelity_exact(_to_oracle(psi), _to_oracle(phi), oU)
    assert te.step_count == 300 and len(log["Infidelity"]["Mean"]) == 300
    assert I1 < 0.5 * I0, (I0, I1)
    assert te.infidelity.mean == pytest.approx(log["Infidelity"]["Mean"][-1])
    assert te.cv == -0.5


def test_target_infidelity_stops_early_and_sgd_and_ptvmc():
    hi = nk.hilbert.Spin(0.5, 4)
    sampler = nk.sampler.MetropolisLocal(hilbert=hi, n_chains_per_rank=256)
    model = nk.models.RBM(alpha=2, param_dtype=complex)
    phi = nk.vqs.MCState(sampler=sampler, model=model, n_samples=4096, seed=5)
    psi = nk.vqs.MCState(sampler=sampler, model=model, n_samples=4096, seed=5)   # psi == phi: I ~ 0 from the start
    te = nkf.driver.InfidelityOptimizer(phi, nk.optimizer.Sgd(0.01), variational_state=psi, cv_coeff=-0.5)
    te.run(n_iter=200, target_infidelity=1e-3, show_progress=False)
    assert te.step_count < 60                                        # patience 30 (+ window) then stop
    # PTVMC: two time steps of Rx on site 0; the target must follow the state (driver/ptvmc.py:54)
    U = nkf.operator.Rx(hi, 0, 0.2)
    pt = nkf.driver.PTVMC(phi, U, psi, nk.optimizer.Adam(0.02), tf=0.2, dt=0.1, n_iter=150, U_dagger=U.H,
                          is_unitary=True, cv_coeff=-0.5, verbose=False)
    pt.run()
    k_t = pt._te._I_op.target.parameters["Dense"]["kernel"]
    assert torch.equal(k_t, psi.parameters["Dense"]["kernel"])


def test_renyi2_matches_exact():
    from oracle.renyi2 import renyi2_exact, renyi2_swap_estimator
    N = 8
    hi = nk.hilbert.Spin(0.5, N)
    vs = nk.vqs.MCState(nk.sampler.MetropolisLocal(hi, n_chains_per_rank=1024), nk.models.RBM(alpha=2, param_dtype=complex, stddev=0.3),
                        n_samples=2 ** 19, n_discard_per_chain=50, seed=3)
    part = [0, 1, 2, 3]
    st = vs.expect(nkf.renyi2.Renyi2EntanglementEntropy(hi, part))
    p = _to_oracle(vs)
    exact = renyi2_exact(p, part)
    assert abs(st.mean - exact) < max(5 * st.error_of_mean, 5e-3), (st.mean, exact, st.error_of_mean)
    # same samples through the oracle's estimator: identical up to fp64 rounding
    S2_ref, _ = renyi2_swap_estimator(p, vs.samples.cpu().numpy(), part)
    assert st.mean == pytest.approx(S2_ref, abs=1e-9)
    assert vs.expect(nkf.renyi2.Renyi2EntanglementEntropy(hi, [])).mean == 0.0


def test_stats_on_device_match_oracle_statistics():
    from oracle.stats import statistics
    from netket_fidelity_b200._native import get_handle
    from netket_fidelity_b200.nk.stats import stats_from_partials
    h = get_handle(0)
    rng = np.random.default_rng(1)
    for rows, cols in ((64, 16384), (7, 100), (16, 7)):
        data = rng.normal(size=(rows, cols)) + 0.2 * rng.normal(size=(rows, 1))
        for dt in (torch.float64, torch.float32):
            L = torch.tensor(data, dtype=dt, device="cuda").reshape(-1)
            l_block = max(1, cols // 32)
            nb = cols // l_block
            part = h.chain_stats(L, rows, cols, l_block, nb).cpu().numpy()
            d = L.double().cpu().numpy().reshape(rows, cols)
            st = stats_from_partials(part, rows, cols, l_block, nb, ((d - d.mean()) ** 2).sum())
            ref = statistics(d)
            for k in ("mean", "variance", "error_of_mean", "tau_corr", "R_hat"):
                a, b = getattr(st, k), ref[k]
                assert (math.isnan(a) and np.isnan(b)) or a == pytest.approx(b, rel=1e-9, abs=1e-12), k


def test_log_value_to_array_and_sample_Upsi_state():
    hi = nk.hilbert.Spin(0.5, 5)
    vs = nk.vqs.MCState(nk.sampler.MetropolisLocal(hi), nk.models.RBM(alpha=2, param_dtype=complex, stddev=0.2),
                        n_samples=64, seed=9)
    from oracle.exact import to_array
    np.testing.assert_allclose(np.abs(vs.to_array().cpu().numpy()), np.abs(to_array(_to_oracle(vs))), rtol=1e-10)
    U = nkf.operator.Rx(hi, 2, 0.7)
    op = nkf.InfidelityOperator(vs, U=U, sample_Upsi=True)
    tgt = op.target
    assert tgt is not vs and tgt.model_state["unitary"] is U and tgt.n_samples == vs.n_samples
    from oracle.rbm import logpsi_U
    st = hi.all_states()
    ref = logpsi_U(_to_oracle(vs), oops.Rx(5, 2, 0.7), st)
    got = tgt.log_value(st).cpu().numpy()
    np.testing.assert_allclose(np.exp(got - ref), 1.0, rtol=1e-10)


# ---------------------------------------------------------------------------------------------- FullSumState
@pytest.mark.parametrize("which", ["Identity", "Rx", "Ry", "Hadamard"])
@pytest.mark.parametrize("pdtype", [float, complex])
@pytest.mark.parametrize("N,alpha", [(3, 1), (6, 2)])
def test_FullSumState_like_the_reference_test(which, pdtype, N, alpha):
    """test/test_infidelity_operator.py:140-182: the exact (full-summation) path.  expect == expect_and_grad,
    mean == exact infidelity, gradient == central differences of the exact infidelity (the reference accepts
    rel 0.5; the closed form here agrees to 1e-6), cv_coeff silently dropped (overlap/operator.py:34-35)."""
    hi = nk.hilbert.Spin(0.5, N)
    ma = nk.models.RBM(alpha=alpha, param_dtype=pdtype, stddev=0.3)
    vs_t = nk.vqs.FullSumState(hi, ma, seed=11)
    vs = nk.vqs.FullSumState(hi, ma, seed=22)
    U, oU = {"Identity": (None, None), "Rx": (nkf.operator.Rx(hi, 1, 0.4), oops.Rx(N, 1, 0.4)),
             "Ry": (nkf.operator.Ry(hi, 0, 0.7), oops.Ry(N, 0, 0.7)),
             "Hadamard": (nkf.operator.Hadamard(hi, 2), oops.Hadamard(N, 2))}[which]
    psi, phi = _to_oracle(vs), _to_oracle(vs_t)
    I_exact = infidelity_exact(psi, phi, oU)
    g_exact = infidelity_exact_grad_fd(psi, phi, oU, 1e-5)
    I_op = nkf.InfidelityOperator(vs_t, U=U, U_dagger=None if U is None else U.H, cv_coeff=-0.5, is_unitary=True)
    assert I_op.cv_coeff is None
    I_stat1 = vs.expect(I_op)
    I_stat, I_grad = vs.expect_and_grad(I_op)
    assert I_stat1.error_of_mean == 0.0 and I_stat1.variance == 0.0
    assert I_stat1.mean == pytest.approx(I_exact, abs=1e-10)
    assert I_stat1.mean == pytest.approx(I_stat.mean, abs=1e-12)
    g = _flat(I_grad)
    if pdtype is float:
        g_exact = g_exact.real
        assert not np.iscomplexobj(g)
    np.testing.assert_allclose(g, g_exact, atol=2e-7 * max(1.0, np.abs(g_exact).max()))


def test_FullSumState_errors_and_to_array():
    hi = nk.hilbert.Spin(0.5, 4)
    ma = nk.models.RBM(alpha=1, param_dtype=complex)
    vs = nk.vqs.FullSumState(hi, ma, seed=3)
    arr = vs.to_array()
    assert arr.shape == (16,) and abs(float((arr.abs() ** 2).sum()) - 1.0) < 1e-12
    np.testing.assert_allclose(arr.cpu().numpy(), oops_to_array(_to_oracle(vs)), atol=1e-12)
    sa = nk.sampler.MetropolisLocal(hilbert=hi, n_chains_per_rank=16)
    mc = nk.vqs.MCState(sampler=sa, model=ma, n_samples=64, seed=1)
    with pytest.raises(TypeError, match="exact states"):
        vs.expect(nkf.InfidelityOperator(mc))
    with pytest.raises(TypeError):
        vs.expect(nkf.InfidelityOperator(nk.vqs.FullSumState(nk.hilbert.Spin(0.5, 3), ma, seed=1)))


def oops_to_array(p):
    from oracle.exact import to_array
    return to_array(p)


def test_state_learning_example_flow():
    """examples/state_learning/state_learning.py: a LogStateVector target (here the exact ground state of a 2x2
    transverse-field Ising model), an RBM FullSumState trained with InfidelityOptimizer + Adam."""
    g = nk.graph.Hypercube(2, 2) if hasattr(nk.graph, "Hypercube") else nk.graph.Chain(4)
    hi = nk.hilbert.Spin(0.5, 4)
    H = oops.Ising(4, oops.chain_edges(4), 1.0, 1.0).to_dense().real
    w, v = np.linalg.eigh(H)
    psi0 = v[:, 0]
    vs_target = nk.vqs.FullSumState(hi, nk.models.LogStateVector(hi))
    vs_target.parameters = {"logstate": torch.tensor(np.log(psi0 + 0j))}
    np.testing.assert_allclose(np.abs(vs_target.to_array().cpu().numpy()), np.abs(psi0), atol=1e-12)
    vs = nk.vqs.FullSumState(hi, nk.models.RBM(alpha=2, param_dtype=complex, stddev=0.05), seed=5)
    inf = nkf.InfidelityOperator(vs_target)
    I0 = vs.expect(inf).mean
    driver = nkf.driver.InfidelityOptimizer(vs_target, nk.optimizer.Adam(learning_rate=0.02), variational_state=vs)
    log = nk.logging.RuntimeLog()
    driver.run(300, out=log, show_progress=False)
    I1 = vs.expect(inf).mean
    assert I1 < 0.02 < I0, (I0, I1)


def test_examples_run_short():
    """The ported reference examples (examples/): a short adiabatic sweep (Trotter sequencing: Z half-steps on the
    visible bias + one Rx optimisation per site, reference examples/adiabatic_sweeping) and state learning on a
    3x3 lattice (reference examples/state_learning)."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def load(name):
        spec = importlib.util.spec_from_file_location(name, os.path.join(root, "examples", name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    ts, obs, infid = load("adiabatic_sweeping").main(["--tf", "0.1", "--n-iter", "40", "--n-samples", "4096",
                                                      "--n-chains", "256", "--quiet"])
    assert len(obs) == 2 and all(abs(o.mean) < 0.2 for o in obs)      # starts in |+>^N: <sigma^z> ~ 0
    assert all(i < 5e-3 for i in infid), infid                         # each small rotation is tracked closely
    log = load("state_learning").main(["--length", "3", "--n-iter", "150"])
    I = np.asarray(log.data["Infidelity"]["Mean"] if isinstance(log.data["Infidelity"], dict) else
                   [v for v in log.data["Infidelity"]])
    assert I.ravel()[-1] < 0.5 * I.ravel()[0]


def test_nonfinite_local_estimators_are_counted():
    """sums[5] of nkfb_infidelity_fwd: a NaN parameter makes every local estimator NaN; the kernels count them and the
    public Stats carries the count (the reference has no such guard; test/_finite_diff.py:18,28 only asserts)."""
    import warnings
    hi = nk.hilbert.Spin(0.5, 6)
    sa = nk.sampler.MetropolisLocal(hi, n_chains=16)
    ma = nk.models.RBM(alpha=1, param_dtype=complex)
    psi = nk.vqs.MCState(sa, ma, n_samples=256, seed=1)
    phi = nk.vqs.MCState(sa, ma, n_samples=256, seed=2)
    op = nkf.InfidelityOperator(phi, cv_coeff=-0.5)
    assert psi.expect(op).n_nonfinite == 0
    psi.samples_packed()                                   # keep these samples, then poison one parameter
    smp = psi._samples_packed
    psi._flat[0] = float("nan")
    psi._samples_packed = smp
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        st = psi.expect(op)
    assert st.n_nonfinite == 256 and any("NaN" in str(x.message) for x in w)
