"""Parity against the REFERENCE'S OWN SOURCE TEXT.

tests/golden/reference_text_vectors.npz holds the outputs of the reference's files executed as they are
(tests/golden/make_reference_text_golden.py under the jax / netket stand-ins of tests/_refshim.py):
operator/singlequbit_gates.py, utils/sampling_Ustate.py, infidelity/overlap/expect.py,
infidelity/overlap_U/expect.py, test/_infidelity_exact.py, test/_finite_diff.py.

* the CPU oracle (oracle/) must reproduce them               -> pins the oracle to the reference's text;
* where /root/reference is present the vectors are re-generated live and must equal the committed file;
* the CUDA kernels (through the C ABI) must reproduce them   -> bit-exact connected elements, 1e-10 relative in
  complex128, 1e-5 relative in complex64 on the per-sample estimator, the infidelity and the max-norm gradient
  (the tolerances BASELINE.json's north_star states).
"""
import os

import numpy as np
import pytest

from oracle import operators as oops
from oracle import exact as oex
from oracle.rbm import RBMParams, logpsi_U, rbm_logpsi
from oracle.infidelity import infidelity_and_grad, mode_slots
from tests import _refshim as shim

G = os.path.join(os.path.dirname(__file__), "golden")
TAGS = (("c1", 4, 4), ("c2", 20, 40), ("odd", 33, 70))
ANGLE = 0.37
LS = (-1.0, 1.0)


def _load(tag, v):
    psi = RBMParams(v[f"{tag}_psi_kernel"], v[f"{tag}_psi_bias"], v[f"{tag}_psi_vb"])
    phi = RBMParams(v[f"{tag}_phi_kernel"], v[f"{tag}_phi_bias"], v[f"{tag}_phi_vb"])
    return psi, phi, v[f"{tag}_sigma"], v[f"{tag}_eta"]


def _vectors():
    return np.load(os.path.join(G, "oracle_vectors.npz")), np.load(os.path.join(G, "reference_text_vectors.npz"))


def _gates(N, ls=LS):
    idx = N // 2
    return dict(rx=oops.Rx(N, idx, ANGLE, ls), ry=oops.Ry(N, idx, ANGLE, ls), h=oops.Hadamard(N, idx, ls))


def _cases(N):
    """(name in the fixture, mode, U, U_dagger, cv) for every estimator case the fixture holds"""
    g = _gates(N)
    rxd, ryd = oops.Rx(N, N // 2, -ANGLE), oops.Ry(N, N // 2, -ANGLE)
    out = []
    for cv, cvn in ((-0.5, "cv"), (None, "nocv")):
        out += [(f"standard_{cvn}", "standard", None, None, cv), (f"standard_Uphi_{cvn}", "standard_Uphi", g["rx"], None, cv),
                (f"upsi_{cvn}", "upsi", g["rx"], rxd, cv)]
    out += [("upsi_ry", "upsi", g["ry"], ryd, -0.5), ("upsi_h", "upsi", g["h"], g["h"], -0.5)]
    return out


def _same_mod_2pi(a, b, tol):
    return np.max(np.abs(np.exp(np.asarray(a) - np.asarray(b)) - 1)) < tol


# ------------------------------------------------------------------------------------------------ oracle vs text
def test_oracle_gates_equal_reference_text():
    v, r = _vectors()
    for tag, N, M in TAGS:
        sigma = v[f"{tag}_sigma"]
        for conv, ls, x in (("", LS, sigma), ("qubit_", (0.0, 1.0), (sigma + 1) / 2)):
            for name, op in _gates(N, ls).items():
                xp, mels = op.get_conn_padded(x)
                assert np.array_equal(xp, r[f"{tag}_ref_{conv}{name}_xp"]), (tag, conv, name)     # bit-exact
                np.testing.assert_allclose(mels, r[f"{tag}_ref_{conv}{name}_mels"], rtol=0, atol=1e-16)
        g = _gates(N)
        assert g["rx"].H.angle == float(r[f"{tag}_ref_rx_H_angle"])          # Rx.H = Rx(-angle)
        assert g["ry"].H.angle == float(r[f"{tag}_ref_ry_H_angle"])          # Ry.H = Ry(-2 angle) (sic)


def test_oracle_logpsiU_equals_reference_text():
    v, r = _vectors()
    for tag, N, M in TAGS:
        psi, phi, sigma, eta = _load(tag, v)
        for name, op in _gates(N).items():
            assert _same_mod_2pi(logpsi_U(phi, op, sigma), r[f"{tag}_ref_logpsiU_{name}"], 1e-12), (tag, name)


def test_oracle_estimator_and_gradient_equal_reference_text():
    v, r = _vectors()
    for tag, N, M in TAGS:
        psi, phi, sigma, eta = _load(tag, v)
        for name, mode, U, Ud, cv in _cases(N):
            o = infidelity_and_grad(psi, phi, sigma, eta, cv, *mode_slots(mode, U, Ud))
            np.testing.assert_allclose(o["L"], r[f"{tag}_ref_{name}_L"], rtol=1e-11, atol=1e-13, err_msg=f"{tag} {name}")
            scale = max(1.0, np.max(np.abs(o["L"])))
            assert abs(o["I"] - float(r[f"{tag}_ref_{name}_I"])) < 1e-12 * scale
            assert abs(o["L"].var() - float(r[f"{tag}_ref_{name}_var"])) < 1e-12 * scale ** 2
            gref = r[f"{tag}_ref_{name}_grad"]
            assert np.max(np.abs(o["grad"].ravel() - gref)) < 1e-11 * max(1.0, np.max(np.abs(gref))), (tag, name)


def test_oracle_real_parameters_equal_reference_text():
    v, r = _vectors()
    N = 20
    psi = RBMParams(v["c2_psi_kernel"].real.copy(), v["c2_psi_bias"].real.copy(), v["c2_psi_vb"].real.copy())
    phi = RBMParams(v["c2_phi_kernel"].real.copy(), v["c2_phi_bias"].real.copy(), v["c2_phi_vb"].real.copy())
    U, Ud = oops.Rx(N, N // 2, ANGLE), oops.Rx(N, N // 2, -ANGLE)
    for mode in ("standard", "upsi"):
        o = infidelity_and_grad(psi, phi, v["c2_sigma"], v["c2_eta"], -0.5, *mode_slots(mode, U, Ud))
        gref = r[f"c2r_ref_{mode}_cv_grad"]
        assert not np.iscomplexobj(gref) or np.max(np.abs(gref.imag)) == 0
        np.testing.assert_allclose(o["L"], r[f"c2r_ref_{mode}_cv_L"], rtol=1e-11, atol=1e-13)
        assert np.max(np.abs(o["grad"].ravel() - gref.real)) < 1e-11 * max(1.0, np.max(np.abs(gref)))


def test_oracle_exact_infidelity_equals_reference_test_helpers():
    """test/_infidelity_exact.py + test/_finite_diff.py (the reference's own ground truth) vs oracle/exact.py"""
    v, r = _vectors()
    psi, phi, _, _ = _load("c1", v)
    for uname, U in (("none", None), ("rx", oops.Rx(4, 2, ANGLE))):
        assert abs(oex.infidelity_exact(psi, phi, U) - float(r[f"c1_ref_exact_{uname}_I"])) < 1e-13
        g = oex.infidelity_exact_grad_fd(psi, phi, U, eps=1e-5)
        np.testing.assert_allclose(g, r[f"c1_ref_exact_{uname}_fdgrad"], rtol=0, atol=1e-9)


# ------------------------------------------------------------------------------------------------ live re-run
@pytest.mark.skipif(not shim.reference_available(), reason="/root/reference is not present on this machine")
def test_fixture_is_what_the_reference_text_computes_now():
    """Re-executes the reference's files (gates, log(U phi), both estimators with gradient) for one tag and compares
    with the committed fixture, so the fixture cannot drift from the text it claims to come from."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk_ref_golden", os.path.join(G, "make_reference_text_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    ref = shim.install()
    v, r = _vectors()
    tag, N = "c1", 4
    psi, phi, sigma, eta = _load(tag, v)
    hi = ref.Hilbert(N, LS)
    U, Ud = ref.gates.Rx(hi, N // 2, ANGLE), ref.gates.Rx(hi, N // 2, -ANGLE)
    xp, mels = U.get_conn_padded(ref.Arr(sigma))
    assert np.array_equal(np.asarray(xp), r[f"{tag}_ref_rx_xp"]) and np.array_equal(np.asarray(mels), r[f"{tag}_ref_rx_mels"])
    for mode in ("standard", "standard_Uphi", "upsi"):
        out = mk.run_mode(ref, mode, psi, phi, sigma, eta, -0.5, U, Ud)
        for k, val in out.items():
            np.testing.assert_allclose(val, r[f"{tag}_ref_{mode}_cv_{k}"], rtol=1e-13, atol=1e-15)


@pytest.mark.skipif(not shim.reference_available(), reason="/root/reference is not present on this machine")
def test_factory_rules_of_the_reference_text_match_the_drop_in():
    """infidelity/logic.py:142-187 executed from the reference vs netket_fidelity_b200.infidelity.logic: which
    operator kind each argument combination yields, and the two error rules."""
    ref = shim.install()
    hi = ref.Hilbert(3, LS)
    vs = ref.MCState(apply_fun=ref.rbm_apply, variables={"params": {}}, hilbert=hi, n_samples=16)
    U, Ud = ref.gates.Rx(hi, 0, 0.1), ref.gates.Rx(hi, 0, -0.1)
    IO = ref.logic.InfidelityOperator
    assert type(IO(vs)).__name__ == "InfidelityOperatorStandard"
    assert type(IO(vs, U=U, U_dagger=Ud, is_unitary=True)).__name__ == "InfidelityOperatorUPsi"
    assert type(IO(vs, U=U, U_dagger=Ud, sample_Upsi=True)).__name__ == "InfidelityOperatorStandard"
    with pytest.raises(ValueError):
        IO(vs, U=U, U_dagger=Ud)                       # non-unitary without sample_Upsi (logic.py:163-175)
    op = IO(vs, U=U, U_dagger=Ud, is_unitary=True, cv_coeff=-0.5)
    assert float(np.asarray(op.cv_coeff)) == -0.5
    with pytest.raises(TypeError):
        IO(vs, U=U, U_dagger=Ud, is_unitary=True, cv_coeff=1j)
    # the drop-in's factory: same branches, same exception types (tests/test_host_logic.py covers it in depth)
    from netket_fidelity_b200.infidelity import logic as mine
    import inspect
    sig_ref = list(inspect.signature(IO).parameters)
    sig_mine = list(inspect.signature(mine.InfidelityOperator).parameters)
    assert sig_ref == sig_mine


# ------------------------------------------------------------------------------------------------ CUDA vs text
@pytest.mark.gpu
def test_cuda_connected_elements_bit_exact_vs_reference_text():
    import torch
    from netket_fidelity_b200._native import get_handle
    from tests._helpers import to_opspec
    h = get_handle(0)
    v, r = _vectors()
    for tag, N, M in TAGS:
        sigma = v[f"{tag}_sigma"]
        for conv, ls, x in (("", LS, sigma), ("qubit_", (0.0, 1.0), (sigma + 1) / 2)):
            for name, op in _gates(N, ls).items():
                xp, mels = h.conn(to_opspec(op), torch.tensor(x, dtype=torch.float64, device="cuda"), ls)
                assert np.array_equal(xp.cpu().numpy(), r[f"{tag}_ref_{conv}{name}_xp"]), (tag, conv, name)
                assert np.array_equal(mels.cpu().numpy(), r[f"{tag}_ref_{conv}{name}_mels"].astype(complex)), (tag, conv, name)


@pytest.mark.gpu
@pytest.mark.parametrize("prec,tol", [("c128", 1e-10), ("c64", 1e-5)])
def test_cuda_estimator_and_gradient_vs_reference_text(prec, tol):
    """tolerances = north_star's: 1e-10 relative (complex128 kernels) / 1e-5 (complex64 kernels), on the per-sample
    local estimator (relative to max |L|), the infidelity, and the gradient in max norm."""
    import torch
    from netket_fidelity_b200._native import get_handle
    from tests._helpers import to_rbmspec, to_opspec, pack_np, packed_to_torch, rel_err
    cd = torch.complex128 if prec == "c128" else torch.complex64
    h = get_handle(0)
    v, r = _vectors()
    for tag, N, M in TAGS:
        psi, phi, sigma, eta = _load(tag, v)
        sp, ep = packed_to_torch(pack_np(sigma)), packed_to_torch(pack_np(eta))
        a, b = to_rbmspec(psi, cd), to_rbmspec(phi, cd)
        for name, op in _gates(N).items():
            lpu = h.logpsi(b, sp, LS, to_opspec(op)).cpu().numpy()
            # (U phi)(x) = sum_c mel_c phi(x'_c) can cancel (Hadamard: +-phi(x) + phi(x')): the tolerance applies to
            # the amplitudes that are summed, i.e. it is scaled by the condition number of the sum, per sample
            xp, mels = op.get_conn_padded(sigma)
            terms = np.abs(mels * np.exp(rbm_logpsi(phi, xp.reshape(-1, N)).reshape(mels.shape)))
            cond = terms.sum(-1) / np.abs(np.exp(r[f"{tag}_ref_logpsiU_{name}"]))
            err = np.abs(np.exp(lpu - r[f"{tag}_ref_logpsiU_{name}"]) - 1)
            assert np.all(err < tol * np.maximum(cond, 1.0)), (tag, name, float(np.max(err / np.maximum(cond, 1.0))))
        for name, mode, U, Ud, cv in _cases(N):
            U1, U2, U4 = mode_slots(mode, U, Ud)
            lv, L, rho1, sums = h.infidelity_fwd(a, b, sp, ep, LS, cv, to_opspec(U1), to_opspec(U2), to_opspec(U4))
            assert rel_err(L.cpu().numpy(), r[f"{tag}_ref_{name}_L"]) < tol, (tag, name)
            s = sums.cpu().numpy()
            assert abs(1 - s[0] / s[4] - float(r[f"{tag}_ref_{name}_I"])) < tol, (tag, name)
            gacc = h.infidelity_grad(a, sp, ep, LS, cv, lv, rho1, sums, to_opspec(U2))
            g = h.grad_finalize(gacc, sums, a.n_params, cd).cpu().numpy()
            assert rel_err(g, r[f"{tag}_ref_{name}_grad"]) < tol, (tag, name)
