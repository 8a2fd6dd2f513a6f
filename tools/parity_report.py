"""Prints the measured parity errors of the CUDA kernels against the oracle on the fixed batches of
tests/test_gpu_parity.py / test_gpu_tensorcore.py (per quantity, per shape, both precisions, both implementations):
the numbers behind the tolerances written in the tests and quoted in DESIGN.md section 4.
usage: python tools/parity_report.py > profiles/r02_parity_errors.txt"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from netket_fidelity_b200._native import get_handle  # noqa: E402
from oracle import operators as oops  # noqa: E402
from oracle.rbm import RBMParams  # noqa: E402
from oracle.infidelity import infidelity_and_grad, mode_slots  # noqa: E402
from tests._helpers import to_rbmspec, to_opspec, pack_np, packed_to_torch, rel_err  # noqa: E402

h = get_handle(0)
LS = (-1.0, 1.0)


def net(N, M, seed, std):
    p = RBMParams.random(N, -(-M // N), seed, std=std)
    return RBMParams(p.kernel[:, :M].copy(), p.bias[:M].copy(), p.visible_bias)


print("shape (N, M, S) | dtype | impl | mode | max|log_val| | err L / max|L| | err F | err grad / max|grad|")
for N, M, S, std in [(4, 4, 2500, 0.2), (20, 40, 4096, 0.05), (33, 70, 2100, 0.05), (100, 400, 4096, 0.01), (64, 128, 3000, 0.02)]:
    rng = np.random.default_rng(7)
    psi, phi = net(N, M, 1234, std), net(N, M, 5678, std)
    sigma, eta = rng.choice([-1.0, 1.0], (S, N)), rng.choice([-1.0, 1.0], (S, N))
    sp, ep = packed_to_torch(pack_np(sigma)), packed_to_torch(pack_np(eta))
    modes = [("standard", None, None), ("upsi", oops.Rx(N, N // 2, 0.37), oops.Rx(N, N // 2, -0.37)),
             ("upsi", oops.Hadamard(N, N - 1), oops.Hadamard(N, N - 1)), ("standard_Uphi", oops.Ry(N, 0, 0.3), None)]
    for mode, U, Ud in modes:
        U1, U2, U4 = mode_slots(mode, U, Ud)
        ref = infidelity_and_grad(psi, phi, sigma, eta, -0.5, U1, U2, U4)
        for cd, impls in ((torch.complex128, ("simt",)), (torch.complex64, ("tc", "simt"))):
            a, b = to_rbmspec(psi, cd), to_rbmspec(phi, cd)
            for impl in impls:
                os.environ["NKFB_FWD_IMPL"] = impl
                os.environ["NKFB_GRAD_IMPL"] = impl
                lv, L, rho1, sums = h.infidelity_fwd(a, b, sp, ep, LS, -0.5, to_opspec(U1), to_opspec(U2), to_opspec(U4))
                gacc = h.infidelity_grad(a, sp, ep, LS, -0.5, lv, rho1, sums, to_opspec(U2))
                g = h.grad_finalize(gacc, sums, a.n_params, cd).cpu().numpy()
                s = sums.cpu().numpy()
                print(f"({N},{M},{S}) | {'c128' if cd == torch.complex128 else 'c64'} | {impl} | {mode}/{type(U).__name__ if U else '-'} | "
                      f"{np.max(np.abs(ref['log_val'])):.2f} | {rel_err(L.cpu().numpy(), ref['L']):.2e} | "
                      f"{abs(s[0] / S - ref['F']):.2e} | {rel_err(g, ref['grad'].ravel()):.2e}")
os.environ.pop("NKFB_FWD_IMPL")
os.environ.pop("NKFB_GRAD_IMPL")
