"""Tensor-core (tcgen05) estimator / gradient kernels against the CPU oracle and against the CUDA-core kernels
(fp32; tolerance 1e-5 relative as stated in BASELINE.json's north_star)."""
import os

import numpy as np
import pytest
import torch

from oracle import operators as oops
from oracle.rbm import RBMParams
from oracle.infidelity import infidelity_and_grad, mode_slots
from tests._helpers import to_rbmspec, to_opspec, pack_np, packed_to_torch, rel_err

pytestmark = pytest.mark.gpu
LS = (-1.0, 1.0)


@pytest.fixture(scope="module")
def h():
    from netket_fidelity_b200._native import get_handle
    return get_handle(0)


def _net(N, M, seed, std):
    p = RBMParams.random(N, -(-M // N), seed, std=std)
    return RBMParams(p.kernel[:, :M].copy(), p.bias[:M].copy(), p.visible_bias)


@pytest.mark.parametrize("N,M,S,std", [(4, 4, 2500, 0.2), (20, 40, 4096, 0.05), (33, 70, 2100, 0.05),
                                        (100, 400, 4096, 0.01), (64, 128, 3000, 0.02)])
@pytest.mark.parametrize("cv", [None, -0.5])
def test_tc_estimator_and_gradient(h, N, M, S, std, cv):
    rng = np.random.default_rng(7)
    psi, phi = _net(N, M, 1234, std), _net(N, M, 5678, std)
    sigma, eta = rng.choice([-1.0, 1.0], (S, N)), rng.choice([-1.0, 1.0], (S, N))
    sp, ep = packed_to_torch(pack_np(sigma)), packed_to_torch(pack_np(eta))
    a, b = to_rbmspec(psi, torch.complex64), to_rbmspec(phi, torch.complex64)
    modes = [("standard", None, None), ("upsi", oops.Rx(N, N // 2, 0.37), oops.Rx(N, N // 2, -0.37)),
             ("upsi", oops.Hadamard(N, N - 1), oops.Hadamard(N, N - 1)), ("standard_Uphi", oops.Ry(N, 0, 0.3), None)]
    for mode, U, Ud in modes:
        U1, U2, U4 = mode_slots(mode, U, Ud)
        ref = infidelity_and_grad(psi, phi, sigma, eta, cv, U1, U2, U4)
        out = {}
        for impl in ("tc", "simt"):
            os.environ["NKFB_FWD_IMPL"] = impl
            os.environ["NKFB_GRAD_IMPL"] = impl
            lv, L, rho1, sums = h.infidelity_fwd(a, b, sp, ep, LS, cv, to_opspec(U1), to_opspec(U2), to_opspec(U4))
            gacc = h.infidelity_grad(a, sp, ep, LS, cv, lv, rho1, sums, to_opspec(U2))
            g = h.grad_finalize(gacc, sums, a.n_params, torch.complex64).cpu().numpy()
            torch.cuda.synchronize()
            s = sums.cpu().numpy()
            out[impl] = (L.cpu().numpy(), s, g, rho1.cpu().numpy())
            assert s[4] == S
            # north_star's complex64 tolerance (1e-5) on F and the max-norm gradient, twice that on the per-sample
            # estimator relative to max |L| (profiles/r02_parity_errors.txt: tcgen05 worst 5.3e-6 / 4.3e-6, CUDA cores
            # 8.1e-6 / 1.8e-6 on these batches)
            assert rel_err(out[impl][0], ref["L"]) <= 2e-5, (impl, mode, U, rel_err(out[impl][0], ref["L"]))
            assert abs(s[0] / S - ref["F"]) <= 1e-5 * max(1.0, abs(ref["F"])), (impl, mode, U)
            assert rel_err(g, ref["grad"].ravel()) <= 1e-5, (impl, mode, U, rel_err(g, ref["grad"].ravel()))
        os.environ.pop("NKFB_FWD_IMPL")
        os.environ.pop("NKFB_GRAD_IMPL")
        # the two implementations must also agree with each other on the per-sample estimator
        assert rel_err(out["tc"][0], out["simt"][0]) <= 2e-5


@pytest.mark.parametrize("N,M,S,std", [(100, 400, 4096, 0.01), (20, 200, 2100, 0.05), (37, 90, 2500, 0.05)])
def test_tc_gradient_eight_and_sixteen_warp_forms_agree(h, monkeypatch, N, M, S, std):
    """Pass B on tensor cores has two forms: sixteen warps (four column quarters per sample row, chunks of a multiple
    of 32 columns, the default where that geometry fits) and eight warps (two halves).  Same tiles, same arithmetic
    per element: the gradients agree to the rounding of the fp64 atomics, and both with the oracle.  (20, 200) has no
    sixteen-warp geometry (224-column chunks): both calls take the eight-warp form there."""
    rng = np.random.default_rng(11)
    psi, phi = _net(N, M, 1234, std), _net(N, M, 5678, std)
    sigma, eta = rng.choice([-1.0, 1.0], (S, N)), rng.choice([-1.0, 1.0], (S, N))
    sp, ep = packed_to_torch(pack_np(sigma)), packed_to_torch(pack_np(eta))
    a, b = to_rbmspec(psi, torch.complex64), to_rbmspec(phi, torch.complex64)
    U, Ud = oops.Rx(N, N - 1, 0.37), oops.Rx(N, N - 1, -0.37)
    U1, U2, U4 = mode_slots("upsi", U, Ud)
    ref = infidelity_and_grad(psi, phi, sigma, eta, -0.5, U1, U2, U4)
    monkeypatch.setenv("NKFB_FWD_IMPL", "tc")
    monkeypatch.setenv("NKFB_GRAD_IMPL", "tc")
    lv, L, rho1, sums = h.infidelity_fwd(a, b, sp, ep, LS, -0.5, to_opspec(U1), to_opspec(U2), to_opspec(U4))
    g = {}
    for form in ("wide", "narrow"):
        if form == "narrow":
            monkeypatch.setenv("NKFB_GRAD_TC_NARROW", "1")
        gacc = h.infidelity_grad(a, sp, ep, LS, -0.5, lv, rho1, sums, to_opspec(U2))
        g[form] = h.grad_finalize(gacc, sums, a.n_params, torch.complex64).cpu().numpy()
        assert rel_err(g[form], ref["grad"].ravel()) <= 1e-5, (form, rel_err(g[form], ref["grad"].ravel()))
    assert rel_err(g["wide"], g["narrow"]) <= 1e-6, rel_err(g["wide"], g["narrow"])


@pytest.mark.parametrize("B", [10.0, 100.0])
def test_tc_estimator_large_imaginary_preactivations(h, B):
    """Hidden biases with imaginary parts +-B, so |Im theta| ~ B (a trained complex RBM).  In single
    precision theta itself carries an absolute error ~ 2^-24 B whatever the kernel, and the tensor-core epilogue's
    cos.approx / sin.approx add the rounding of 2 y / (2 pi) (another ~2^-24 2B in the phase) -- both grow linearly
    with B.  Held here: both fp32 forms stay within 4e-5 max(1, B / 4) of the complex128 oracle on the per-sample
    estimator (1e-5-grade up to |Im theta| ~ 4, degrading in proportion beyond), and within that of each other, so the
    switch from the CUDA-core to the tensor-core kernel at S = 2048 is not a discontinuity beyond fp32 rounding."""
    N, M, S = 20, 40, 4096
    rng = np.random.default_rng(3)
    psi, phi = _net(N, M, 1234, 0.05), _net(N, M, 5678, 0.05)
    sign = np.where(np.arange(M) % 2 == 0, 1.0, -1.0)
    psi = RBMParams(psi.kernel, psi.bias + 1j * B * sign, psi.visible_bias)
    phi = RBMParams(phi.kernel, phi.bias - 1j * B * sign, phi.visible_bias)
    sigma, eta = rng.choice([-1.0, 1.0], (S, N)), rng.choice([-1.0, 1.0], (S, N))
    sp, ep = packed_to_torch(pack_np(sigma)), packed_to_torch(pack_np(eta))
    a, b = to_rbmspec(psi, torch.complex64), to_rbmspec(phi, torch.complex64)
    U, Ud = oops.Rx(N, 3, 0.37), oops.Rx(N, 3, -0.37)
    U1, U2, U4 = mode_slots("upsi", U, Ud)
    # the oracle evaluates the complex64-rounded parameters the kernels see
    c64 = lambda p: RBMParams(p.kernel.astype(np.complex64).astype(np.complex128),
                              p.bias.astype(np.complex64).astype(np.complex128),
                              None if p.visible_bias is None else p.visible_bias.astype(np.complex64).astype(np.complex128))
    ref = infidelity_and_grad(c64(psi), c64(phi), sigma, eta, -0.5, U1, U2, U4)
    tol = 4e-5 * max(1.0, B / 4.0)   # measured at B = 10: 5.0e-5 (tensor cores)
    out = {}
    try:
        for impl in ("tc", "simt"):
            os.environ["NKFB_FWD_IMPL"] = impl
            lv, L, rho1, sums = h.infidelity_fwd(a, b, sp, ep, LS, -0.5, to_opspec(U1), to_opspec(U2), to_opspec(U4))
            torch.cuda.synchronize()
            out[impl] = L.cpu().numpy()
            assert rel_err(out[impl], ref["L"]) <= tol, (impl, B, rel_err(out[impl], ref["L"]))
    finally:
        os.environ.pop("NKFB_FWD_IMPL", None)
    assert rel_err(out["tc"], out["simt"]) <= tol, rel_err(out["tc"], out["simt"])


@pytest.mark.parametrize("N,M,S,std", [(100, 400, 4096, 0.01), (20, 40, 2304, 0.05)])
def test_tc_prepared_tables(h, N, M, S, std):
    """nkfb_tc_prepare + NKFB_OPT_TC_TABLES_VALID: the estimator and the gradient launch without their three preparation
    kernels and give bit-identical per-sample results (same tables, same kernels); with the flag set but OTHER parameter
    buffers the tables are re-prepared (the flag is a promise about the buffers they were written from, nothing else);
    after the flag is cleared an in-place parameter update is seen again."""
    rng = np.random.default_rng(3)
    psi, phi = _net(N, M, 1234, std), _net(N, M, 5678, std)
    sigma, eta = rng.choice([-1.0, 1.0], (S, N)), rng.choice([-1.0, 1.0], (S, N))
    sp, ep = packed_to_torch(pack_np(sigma)), packed_to_torch(pack_np(eta))
    a, b = to_rbmspec(psi, torch.complex64), to_rbmspec(phi, torch.complex64)
    U, Ud = to_opspec(oops.Rx(N, 1, 0.2)), to_opspec(oops.Rx(N, 1, -0.2))

    def run(x, y):
        n0 = h.launches
        lv, L, rho1, sums = h.infidelity_fwd(x, y, sp, ep, LS, -0.5, U, Ud, None)
        gacc = h.infidelity_grad(x, sp, ep, LS, -0.5, lv, rho1, sums, Ud)
        torch.cuda.synchronize()
        return L.clone(), lv.clone(), gacc.clone(), h.launches - n0

    L0, lv0, g0, n_plain = run(a, b)
    try:
        h.tc_prepare(a, b, LS)
        h.set_tc_tables_valid(True)
        L1, lv1, g1, n_prep = run(a, b)
        assert n_prep == n_plain - 3, (n_plain, n_prep)
        assert torch.equal(L0, L1) and torch.equal(lv0, lv1)
        assert rel_err(g1.cpu().numpy(), g0.cpu().numpy()) <= 1e-7       # order of the atomic adds only (measured 4e-9)
        # other buffers under the same flag: prepared again, results of THOSE parameters
        a2, b2 = to_rbmspec(phi, torch.complex64), to_rbmspec(psi, torch.complex64)
        L2, lv2, g2, n2 = run(a2, b2)
        assert n2 == n_plain
        h.set_tc_tables_valid(False)
        L2r, lv2r, g2r, _ = run(a2, b2)
        assert torch.equal(L2, L2r) and torch.equal(lv2, lv2r)
        assert not torch.equal(L2, L0)
        # the tables now belong to (a2, b2); with the flag set again (a, b) is re-prepared, not served from them
        h.set_tc_tables_valid(True)
        L3, lv3, g3, n3 = run(a, b)
        assert n3 == n_plain and torch.equal(L3, L0)
    finally:
        h.set_tc_tables_valid(False)
    # flag cleared: an in-place update of the same buffers is seen
    a.kernel.mul_(1.5)
    L4, _, _, n4 = run(a, b)
    assert n4 == n_plain and not torch.equal(L4, L0)
