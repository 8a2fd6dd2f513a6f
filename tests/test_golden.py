"""The committed golden fixtures (tests/golden/, written by make_golden.py) against (a) the CPU oracle
-- guards the oracle against drift -- and (b) the CUDA kernels through the C ABI."""
import json
import os

import numpy as np
import pytest

from oracle import operators as oops
from oracle.rbm import RBMParams, rbm_logpsi, logpsi_U
from oracle.infidelity import infidelity_and_grad, mode_slots

G = os.path.join(os.path.dirname(__file__), "golden")
TAGS = (("c1", 4, 4), ("c2", 20, 40), ("odd", 33, 70))


def _load(tag, v):
    psi = RBMParams(v[f"{tag}_psi_kernel"], v[f"{tag}_psi_bias"], v[f"{tag}_psi_vb"])
    phi = RBMParams(v[f"{tag}_phi_kernel"], v[f"{tag}_phi_bias"], v[f"{tag}_phi_vb"])
    return psi, phi, v[f"{tag}_sigma"], v[f"{tag}_eta"]


def _modes(N):
    U, Ud = oops.Rx(N, N // 2, 0.37), oops.Rx(N, N // 2, -0.37)
    prop = oops.IsingPropagator(N, oops.chain_edges(N), 1.0, 1.0, 0.01)
    return (("standard", None, None), ("upsi", U, Ud), ("standard_Uphi", prop, None)), U, prop


def test_oracle_reproduces_golden_vectors():
    v = np.load(os.path.join(G, "oracle_vectors.npz"))
    g = json.load(open(os.path.join(G, "gates_golden.json")))
    c, m = oops.conns_and_mels_ry(np.array([[-1.0, 1, -1], [1, 1, 1]]), 0, np.pi / 2, (-1.0, 1.0))
    np.testing.assert_allclose(c, g["conns_spin"])
    np.testing.assert_allclose(m, np.array(g["mels_ry"])[..., 0] + 1j * np.array(g["mels_ry"])[..., 1])
    for tag, N, M in TAGS:
        psi, phi, sigma, eta = _load(tag, v)
        modes, U, prop = _modes(N)
        np.testing.assert_allclose(rbm_logpsi(psi, sigma), v[f"{tag}_logpsi"], rtol=1e-13)
        np.testing.assert_allclose(logpsi_U(phi, U, sigma), v[f"{tag}_logpsiU_rx"], rtol=1e-13)
        for mode, a, b in modes:
            r = infidelity_and_grad(psi, phi, sigma, eta, -0.5, *mode_slots(mode, a, b))
            np.testing.assert_allclose(r["L"], v[f"{tag}_{mode}_L"], rtol=1e-12)
            np.testing.assert_allclose(r["grad"].ravel(), v[f"{tag}_{mode}_grad"], rtol=1e-10, atol=1e-15)


@pytest.mark.gpu
@pytest.mark.parametrize("prec,tol", [("c128", 1e-10), ("c64", 1e-5)])
def test_cuda_kernels_reproduce_golden_vectors(prec, tol):
    import torch
    from netket_fidelity_b200._native import get_handle
    from tests._helpers import to_rbmspec, to_opspec, pack_np, packed_to_torch, rel_err
    cd = torch.complex128 if prec == "c128" else torch.complex64
    h = get_handle(0)
    v = np.load(os.path.join(G, "oracle_vectors.npz"))
    for tag, N, M in TAGS:
        psi, phi, sigma, eta = _load(tag, v)
        modes, U, prop = _modes(N)
        sp, ep = packed_to_torch(pack_np(sigma)), packed_to_torch(pack_np(eta))
        a, b = to_rbmspec(psi, cd), to_rbmspec(phi, cd)
        lp = h.logpsi(a, sp, (-1.0, 1.0)).cpu().numpy()
        assert np.max(np.abs(np.exp(lp - v[f"{tag}_logpsi"]) - 1)) < tol * 10
        lpu = h.logpsi(b, sp, (-1.0, 1.0), to_opspec(prop)).cpu().numpy()
        assert np.max(np.abs(np.exp(lpu - v[f"{tag}_logpsiU_ising"]) - 1)) < tol * 20
        for mode, ua, ub in modes:
            U1, U2, U4 = mode_slots(mode, ua, ub)
            lv, L, rho1, sums = h.infidelity_fwd(a, b, sp, ep, (-1.0, 1.0), -0.5, to_opspec(U1), to_opspec(U2),
                                                 to_opspec(U4))
            assert rel_err(L.cpu().numpy(), v[f"{tag}_{mode}_L"]) < tol * 10
            s = sums.cpu().numpy()
            assert abs(1 - s[0] / s[4] - float(v[f"{tag}_{mode}_I"])) < tol * 10
            gacc = h.infidelity_grad(a, sp, ep, (-1.0, 1.0), -0.5, lv, rho1, sums, to_opspec(U2))
            g = h.grad_finalize(gacc, sums, a.n_params, cd).cpu().numpy()
            assert rel_err(g, v[f"{tag}_{mode}_grad"]) < tol * 20
