"""Writes the committed golden fixtures (run from the repo root: python tests/golden/make_golden.py).

The reference (NetKet/JAX) cannot be imported in this image, so the vectors come from (a) the literal
arrays of the reference's own test (test/test_operator.py:66-96) and (b) the NumPy oracle, which is pinned
by those literals and by exact full-summation + finite differences (tests/test_oracle_*.py)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import operators as oops  # noqa: E402
from oracle.rbm import RBMParams, rbm_logpsi, logpsi_U  # noqa: E402
from oracle.infidelity import infidelity_and_grad, mode_slots  # noqa: E402
from tests.test_oracle_gates import CONNS_QUBIT, CONNS_SPIN, MELS_RX, MELS_RY, MELS_H  # noqa: E402

here = os.path.dirname(os.path.abspath(__file__))
json.dump({"conns_qubit": CONNS_QUBIT.tolist(), "conns_spin": CONNS_SPIN.tolist(),
           "mels_rx": [[[z.real, z.imag] for z in r] for r in MELS_RX],
           "mels_ry": [[[z.real, z.imag] for z in r] for r in MELS_RY], "mels_h": MELS_H.tolist(),
           "source": "netket_fidelity test/test_operator.py:66-96 (idx 0, angle pi/2, states #2 and #7 of 3 sites)"},
          open(os.path.join(here, "gates_golden.json"), "w"), indent=1)

out = {}
rng = np.random.default_rng(42)
for tag, N, M, S in (("c1", 4, 4, 112), ("c2", 20, 40, 256), ("odd", 33, 70, 96)):
    psi = RBMParams.random(N, -(-M // N), 1234, std=0.05)
    phi = RBMParams.random(N, -(-M // N), 5678, std=0.05)
    psi = RBMParams(psi.kernel[:, :M].copy(), psi.bias[:M].copy(), psi.visible_bias)
    phi = RBMParams(phi.kernel[:, :M].copy(), phi.bias[:M].copy(), phi.visible_bias)
    sigma = rng.choice([-1.0, 1.0], size=(S, N))
    eta = rng.choice([-1.0, 1.0], size=(S, N))
    out[f"{tag}_psi_kernel"], out[f"{tag}_psi_bias"], out[f"{tag}_psi_vb"] = psi.kernel, psi.bias, psi.visible_bias
    out[f"{tag}_phi_kernel"], out[f"{tag}_phi_bias"], out[f"{tag}_phi_vb"] = phi.kernel, phi.bias, phi.visible_bias
    out[f"{tag}_sigma"], out[f"{tag}_eta"] = sigma, eta
    out[f"{tag}_logpsi"] = rbm_logpsi(psi, sigma)
    U, Ud = oops.Rx(N, N // 2, 0.37), oops.Rx(N, N // 2, -0.37)
    out[f"{tag}_logpsiU_rx"] = logpsi_U(phi, U, sigma)
    prop = oops.IsingPropagator(N, oops.chain_edges(N), 1.0, 1.0, 0.01)
    out[f"{tag}_logpsiU_ising"] = logpsi_U(phi, prop, sigma)
    for mode, (a, b) in (("standard", (None, None)), ("upsi", (U, Ud)), ("standard_Uphi", (prop, None))):
        U1, U2, U4 = mode_slots(mode, a, b)
        r = infidelity_and_grad(psi, phi, sigma, eta, -0.5, U1, U2, U4)
        out[f"{tag}_{mode}_L"] = r["L"]
        out[f"{tag}_{mode}_I"] = np.array(r["I"])
        out[f"{tag}_{mode}_grad"] = r["grad"].ravel()
np.savez_compressed(os.path.join(here, "oracle_vectors.npz"), **out)
print("wrote", len(out), "arrays")
