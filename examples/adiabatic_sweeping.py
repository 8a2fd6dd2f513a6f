"""Adiabatic sweep of a transverse-field Ising chain with Trotterised p-tVMC -- the reference's
examples/adiabatic_sweeping/adiabatic_sweeping.py on the B200 kernels.  Only the imports change
(`import netket as nk` -> `from netket_fidelity_b200 import nk`, `netket_fidelity` -> `netket_fidelity_b200`,
`flax.serialization.from_bytes` -> `nkf.utils.mpack.load_mpack`); the Trotter sequencing is the reference's:
per time step a diagonal ZZ/Z half-step folded into the visible bias, then N single-site Rx optimisations
(one InfidelityOptimizer.run each, the target following the state), then the second half-step.

    python examples/adiabatic_sweeping.py --tf 1.0 --n-iter 200          # a short run
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import netket_fidelity_b200 as nkf  # noqa: E402
from netket_fidelity_b200 import nk  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--tf", type=float, default=15.0)
    ap.add_argument("--dt", type=float, default=0.05)
    ap.add_argument("--n-iter", type=int, default=1000)
    ap.add_argument("--n-samples", type=int, default=1000)
    ap.add_argument("--n-chains", type=int, default=16)
    ap.add_argument("--quiet", action="store_true")
    args = ap.parse_args(argv)

    N, Gamma, gamma, T = 10, -1.0, 1.0, 8
    ts = np.arange(0, args.tf, args.dt)
    hi = nk.hilbert.Spin(0.5, N)
    sampler = nk.sampler.MetropolisLocal(hilbert=hi, n_chains_per_rank=args.n_chains)
    model = nk.models.RBM(alpha=1, param_dtype=complex)
    phi = nk.vqs.MCState(sampler=sampler, model=model, n_samples=args.n_samples, seed=1)
    psi = nk.vqs.MCState(sampler=sampler, model=model, n_samples=args.n_samples, seed=2)
    optimizer = nk.optimizer.Adam(learning_rate=0.01)

    # the initial state |+>^N of the reference (examples/adiabatic_sweeping/initial_state.mpack)
    init = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden",
                        "adiabatic_initial_state.mpack")
    phi.variables = nkf.utils.mpack.load_mpack(init)
    psi.variables = nkf.utils.mpack.load_mpack(init)
    obs = sum([nk.operator.spin.sigmaz(hi, i) for i in range(N)]) / N

    obs_hist, infid_hist = [], []
    for t in ts:
        if t < T:
            gt, Gt = gamma * t / T, Gamma * (1 - t / T)
        else:
            gt, Gt = gamma * (1 - (t - T) / T), Gamma * (t - T) / T
        # Z diagonal half-step: an exact update of the visible bias
        params = phi.parameters
        params["visible_bias"] = params["visible_bias"] + 1j * gt * args.dt / 2
        psi.parameters = params
        phi.parameters = params
        # X terms: one single-site rotation at a time, psi -> Rx_i phi, then phi <- psi
        last = None
        for i in range(N):
            te = nkf.driver.InfidelityOptimizer(phi, optimizer, U=nkf.operator.Rx(hi, i, 2 * args.dt * Gt),
                                                U_dagger=nkf.operator.Rx(hi, i, -2 * args.dt * Gt),
                                                variational_state=psi, is_unitary=True, cv_coeff=-0.5)
            te.run(n_iter=args.n_iter, show_progress=False)
            last = te.infidelity
            phi.parameters = psi.parameters
        params = phi.parameters
        params["visible_bias"] = params["visible_bias"] + 1j * gt * args.dt / 2
        psi.parameters = params
        phi.parameters = params
        psi.reset()
        o = psi.expect(obs)
        obs_hist.append(o)
        infid_hist.append(last.mean)
        if not args.quiet:
            print(f"t = {t:6.3f}   <sigma^z> = {o.mean:+.4f} +- {o.error_of_mean:.4f}   last infidelity {last.mean:.2e}",
                  flush=True)
    torch.cuda.synchronize()
    return ts, obs_hist, infid_hist


if __name__ == "__main__":
    main()
