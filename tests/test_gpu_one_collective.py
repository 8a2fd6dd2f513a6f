"""One-collective form of the step (reference: ONE mean over ranks at overlap/expect.py:126):
pass B centred with the previous step's F plus the accumulator sum_s conj O_k(sigma_s) must give the gradient of the
two-collective step on the same samples -- it is exact algebra -- in float64 to rounding and in float32 (tensor-core
path: 4736 chains x 1... a batch above the 2048-sample threshold) within the float32 parity tolerance."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prec,N,M,n_chains,chain_length,tol", [
    ("c128", 12, 24, 64, 8, 1e-10), ("c64", 12, 24, 64, 8, 1e-5), ("c64", 100, 400, 64, 64, 1e-5)])
def test_one_collective_step_equals_two_collective_step(prec, N, M, n_chains, chain_length, tol):
    import torch
    from netket_fidelity_b200 import _native as nv
    from netket_fidelity_b200.engine import InfidelityStep, StepConfig
    from netket_fidelity_b200.workloads import gate_spec, random_rbm
    dev = torch.device("cuda", 0)
    cd = torch.complex128 if prec == "c128" else torch.complex64
    p = random_rbm(N, M, 21, cd, True, std=0.05)
    d = random_rbm(N, M, 22, cd, True, std=0.02)
    psi = nv.RbmSpec(*[t.to(dev) for t in p])
    phi = nv.RbmSpec(*[(t + u).to(dev) for t, u in zip(p, d)])
    U, Ud = gate_spec("rx", 1, 0.3), gate_spec("rx", 1, -0.3)
    engs = [InfidelityStep(N, M, StepConfig(n_chains, chain_length, 5, N, seed=77, cv_coeff=-0.5, one_collective=oc),
                           "upsi", U, Ud, device=dev, cdtype=cd) for oc in (False, True)]
    for step in range(3):
        outs = []
        for e in engs:
            sums, L, grad = e.step(psi, phi)
            outs.append((sums.cpu().numpy().copy(), grad.cpu().numpy().copy()))
        (s2, g2), (s1, g1) = outs
        assert np.allclose(s1[:5], s2[:5], rtol=1e-12)                     # same samples, same estimator
        err = np.max(np.abs(g1 - g2)) / np.max(np.abs(g2))
        # step 0 is centred with c0 = 0 instead of F: the float32 weights then carry 2 L instead of 2 (L - F)
        assert err < (tol if step > 0 or prec == "c128" else 20 * tol), (step, err)
