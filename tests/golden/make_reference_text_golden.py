"""Writes tests/golden/reference_text_vectors.npz by EXECUTING THE REFERENCE'S SOURCE FILES (read from
/root/reference at generation time, never copied) under the jax/netket stand-ins of tests/_refshim.py:

    python tests/golden/make_reference_text_golden.py

Inputs are the seeded batches of tests/golden/oracle_vectors.npz (tags c1: N=4 M=4, c2: N=20 M=40, odd: N=33 M=70).
Outputs per tag (names `<tag>_ref_*`):
  gates      Rx / Ry / Hadamard.get_conn_padded(sigma)           operator/singlequbit_gates.py:65-71,89-105,160-166,
                                                                  184-201,250-256,270-288
  logpsiU    _logpsi_U_fun(rbm, {params, unitary: Rx}, sigma)     utils/sampling_Ustate.py:34-47
  standard   infidelity_sampling_MCState(...)                     infidelity/overlap/expect.py:58-130
  standard_Uphi   the same with afun_t = log(U phi), U = Rx       overlap/operator.py:61-82 + sampling_Ustate.py
  upsi       infidelity_sampling_MCState(..., U, ..., U^dagger)   infidelity/overlap_U/expect.py:70-161
each with cv_coeff = -1/2 and None, complex parameters (and real parameters for tag c2r), giving the local
estimator values, the infidelity and the gradient pytree flattened in jax leaf order (Dense/bias, Dense/kernel,
visible_bias).  Plus, at N = 4: the exact infidelity (test/_infidelity_exact.py:4-20) and its central-difference
gradient (test/_finite_diff.py:8-31).

What is reference text and what is a stand-in is stated in tests/_refshim.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests import _refshim as shim  # noqa: E402
from oracle.rbm import RBMParams  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
TAGS = (("c1", 4, 4), ("c2", 20, 40), ("odd", 33, 70))
ANGLE = 0.37
N_CHAINS = 8


def flat_grad(g):
    parts = []
    if "bias" in g["Dense"]:
        parts.append(np.asarray(g["Dense"]["bias"]).ravel())
    parts.append(np.asarray(g["Dense"]["kernel"]).ravel())
    if "visible_bias" in g:
        parts.append(np.asarray(g["visible_bias"]).ravel())
    return np.concatenate(parts)


def run_mode(ref, mode, psi, phi, sigma, eta, cv, U=None, Ud=None):
    """One call of the reference's jitted estimator (text of overlap/expect.py or overlap_U/expect.py)."""
    Arr = ref.Arr
    N = sigma.shape[-1]
    p, pt = shim.params_to_arr(psi), shim.params_to_arr(phi)
    sig3 = Arr(sigma.reshape(N_CHAINS, -1, N))
    eta3 = Arr(eta.reshape(N_CHAINS, -1, N))
    afun = ref.rbm_apply
    if mode == "standard":
        stats, grad = ref.expect_std.infidelity_sampling_MCState(afun, afun, p, pt, {}, {}, sig3, eta3, cv, True)
        stats_ng = ref.expect_std.infidelity_sampling_MCState(afun, afun, p, pt, {}, {}, sig3, eta3, cv, False)
    elif mode == "standard_Uphi":
        afun_t, vars_t = ref.sampling_Ustate.make_logpsi_U_afun(afun, U, {"params": pt})
        ms_t = {k: v for k, v in vars_t.items() if k != "params"}
        stats, grad = ref.expect_std.infidelity_sampling_MCState(afun, afun_t, p, pt, {}, ms_t, sig3, eta3, cv, True)
        stats_ng = ref.expect_std.infidelity_sampling_MCState(afun, afun_t, p, pt, {}, ms_t, sig3, eta3, cv, False)
    elif mode == "upsi":
        stats, grad = ref.expect_U.infidelity_sampling_MCState(afun, afun, p, pt, {}, {}, sig3, U, eta3, Ud, cv, True)
        stats_ng = ref.expect_U.infidelity_sampling_MCState(afun, afun, p, pt, {}, {}, sig3, U, eta3, Ud, cv, False)
    else:
        raise ValueError(mode)
    assert abs(float(stats.mean) - float(stats_ng.mean)) < 1e-14          # expect == expect_and_grad
    return dict(L=stats.local_values, I=np.array(float(stats.mean)), var=np.array(stats.variance),
                grad=flat_grad(grad))


def main():
    ref = shim.install()
    Arr, Hilbert = ref.Arr, ref.Hilbert
    v = np.load(os.path.join(HERE, "oracle_vectors.npz"))
    out = {}
    for tag, N, M in TAGS:
        psi = RBMParams(v[f"{tag}_psi_kernel"], v[f"{tag}_psi_bias"], v[f"{tag}_psi_vb"])
        phi = RBMParams(v[f"{tag}_phi_kernel"], v[f"{tag}_phi_bias"], v[f"{tag}_phi_vb"])
        sigma, eta = v[f"{tag}_sigma"], v[f"{tag}_eta"]
        idx = N // 2
        hi = Hilbert(N, (-1.0, 1.0))
        gates = dict(rx=ref.gates.Rx(hi, idx, ANGLE), ry=ref.gates.Ry(hi, idx, ANGLE), h=ref.gates.Hadamard(hi, idx))
        for name, op in gates.items():
            xp, mels = op.get_conn_padded(Arr(sigma))
            out[f"{tag}_ref_{name}_xp"] = np.asarray(xp)
            out[f"{tag}_ref_{name}_mels"] = np.asarray(mels)
        # the same gates on Qubit-valued configurations (local states 0 / 1)
        hq = Hilbert(N, (0.0, 1.0))
        sq = (sigma + 1) / 2
        for name, op in dict(rx=ref.gates.Rx(hq, idx, ANGLE), ry=ref.gates.Ry(hq, idx, ANGLE),
                             h=ref.gates.Hadamard(hq, idx)).items():
            xp, mels = op.get_conn_padded(Arr(sq))
            out[f"{tag}_ref_qubit_{name}_xp"] = np.asarray(xp)
            out[f"{tag}_ref_qubit_{name}_mels"] = np.asarray(mels)
        # .H rules of the reference classes (singlequbit_gates.py:40-41,135-136,223-224)
        out[f"{tag}_ref_rx_H_angle"] = np.array(float(gates["rx"].H.angle))
        out[f"{tag}_ref_ry_H_angle"] = np.array(float(gates["ry"].H.angle))
        # log(U phi)
        for name in ("rx", "ry", "h"):
            lp = ref.sampling_Ustate._logpsi_U_fun(ref.rbm_apply, {"params": shim.params_to_arr(phi),
                                                                   "unitary": gates[name]}, Arr(sigma))
            out[f"{tag}_ref_logpsiU_{name}"] = np.asarray(lp)
        U, Ud = gates["rx"], ref.gates.Rx(hi, idx, -ANGLE)
        for cv, cvn in ((-0.5, "cv"), (None, "nocv")):
            for mode in ("standard", "standard_Uphi", "upsi"):
                r = run_mode(ref, mode, psi, phi, sigma, eta, cv, U, Ud)
                for k, val in r.items():
                    out[f"{tag}_ref_{mode}_{cvn}_{k}"] = val
        # Ry / Hadamard inside the unitary-mode estimator (untested by the reference itself)
        for name, (a, b) in dict(ry=(gates["ry"], ref.gates.Ry(hi, idx, -ANGLE)), h=(gates["h"], gates["h"])).items():
            r = run_mode(ref, "upsi", psi, phi, sigma, eta, -0.5, a, b)
            for k, val in r.items():
                out[f"{tag}_ref_upsi_{name}_{k}"] = val
        print(tag, "done", flush=True)

    # real parameters: the gradient is the ordinary (real) gradient
    tag, N, M = "c2", 20, 40
    psi_r = RBMParams(v["c2_psi_kernel"].real.copy(), v["c2_psi_bias"].real.copy(), v["c2_psi_vb"].real.copy())
    phi_r = RBMParams(v["c2_phi_kernel"].real.copy(), v["c2_phi_bias"].real.copy(), v["c2_phi_vb"].real.copy())
    hi = Hilbert(N, (-1.0, 1.0))
    U, Ud = ref.gates.Rx(hi, N // 2, ANGLE), ref.gates.Rx(hi, N // 2, -ANGLE)
    for mode in ("standard", "upsi"):
        r = run_mode(ref, mode, psi_r, phi_r, v["c2_sigma"], v["c2_eta"], -0.5, U, Ud)
        for k, val in r.items():
            out[f"c2r_ref_{mode}_cv_{k}"] = val

    # exact infidelity and its central-difference gradient at N = 4 (the reference's own test oracle)
    N = 4
    hi = Hilbert(N, (-1.0, 1.0))
    psi = RBMParams(v["c1_psi_kernel"], v["c1_psi_bias"], v["c1_psi_vb"])
    phi = RBMParams(v["c1_phi_kernel"], v["c1_phi_bias"], v["c1_phi_vb"])
    vstate = ref.MCState(apply_fun=ref.rbm_apply, variables={"params": shim.params_to_arr(phi)}, hilbert=hi)
    for uname, U in (("none", None), ("rx", ref.gates.Rx(hi, 2, ANGLE))):
        out[f"c1_ref_exact_{uname}_I"] = np.array(complex(np.asarray(
            ref.exact_helper._infidelity_exact(shim.params_to_arr(psi), vstate, U))).real)

        def fun(flat):
            return np.float64(complex(np.asarray(ref.exact_helper._infidelity_exact(shim.params_to_arr(psi.unravel(flat)), vstate, U))).real)

        out[f"c1_ref_exact_{uname}_fdgrad"] = ref.finite_diff.central_diff_grad(fun, psi.ravel(), 1e-5)
    np.savez_compressed(os.path.join(HERE, "reference_text_vectors.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
