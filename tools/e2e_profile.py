"""Host-side profile of the end-to-end step of bench.py (public API with pinned host parameters): where the
time between the device-resident step and the e2e step goes.  usage: python tools/e2e_profile.py [workload]"""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import netket_fidelity_b200 as nkf  # noqa: E402
from netket_fidelity_b200 import nk  # noqa: E402
from netket_fidelity_b200.workloads import make_workload, random_rbm  # noqa: E402
import bench  # noqa: E402

w = make_workload(sys.argv[1] if len(sys.argv) > 1 else "C4")
N, M = w["N"], w["M"]
dev = torch.device("cuda", 0)
cd = torch.complex64
host = {k: random_rbm(N, M, seed, cd, w["vb"]) for k, seed in (("psi", 1234), ("phi", 5678))}
pinned = {k: [t.pin_memory() if t is not None else None for t in v] for k, v in host.items()}
mk_tree = lambda t: {"Dense": {"kernel": t[0], "bias": t[1]}, **({"visible_bias": t[2]} if t[2] is not None else {})}
hi = nk.hilbert.Spin(0.5, N)
n_chains = int(os.environ.get("E2E_CHAINS", w["n_chains"]))
sa = nk.sampler.MetropolisLocal(hi, n_chains=n_chains, sweep_size=N)
ma = nk.models.RBM(alpha=M // N, param_dtype="complex64", use_visible_bias=w["vb"])
S = n_chains * w["chain_length"]
vs_psi = nk.vqs.MCState(sa, ma, n_samples=S, n_discard_per_chain=5, seed=1, device=dev,
                        variables={"params": mk_tree([t.to(dev) if t is not None else None for t in host["psi"]])})
vs_phi = nk.vqs.MCState(sa, ma, n_samples=S, n_discard_per_chain=5, seed=2, device=dev,
                        variables={"params": mk_tree([t.to(dev) if t is not None else None for t in host["phi"]])})
Uop, Udop = bench.host_ops(w, hi)
I_op = nkf.InfidelityOperator(vs_phi, U=Uop, U_dagger=Udop, cv_coeff=-0.5, is_unitary=(w["mode"] == "upsi"),
                              sample_Upsi=(w["mode"] == "standard_Uphi"))
pin_tree = {k: mk_tree(pinned[k]) for k in pinned}
grad_host = torch.empty((vs_psi._flat.numel(),), dtype=cd).pin_memory()


def e2e_step():
    vs_psi.parameters = pin_tree["psi"]
    I_op.target.parameters = pin_tree["phi"]
    vs_psi.reset()
    I_op.target.reset()
    st, grad = vs_psi.expect_and_grad(I_op)
    grad_host.copy_(vs_psi._last_flat_grad, non_blocking=True)
    torch.cuda.synchronize()
    return st


for _ in range(3):
    e2e_step()
t0 = time.perf_counter()
for _ in range(10):
    e2e_step()
print(f"e2e step {1e3 * (time.perf_counter() - t0) / 10:.3f} ms  ({n_chains} chains per family)")
# host time until everything is enqueued (no final synchronize) vs the whole step
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    e2e_step()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)

# host timeline of one step (the GPU is idle from the synchronize that ends a step until the first sampler launch of the next)
import numpy as np  # noqa: E402
seg = []
for _ in range(20):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    vs_psi.parameters = pin_tree["psi"]
    I_op.target.parameters = pin_tree["phi"]
    t1 = time.perf_counter()
    vs_psi.reset()
    I_op.target.reset()
    t2 = time.perf_counter()
    st, grad = vs_psi.expect_and_grad(I_op)
    t3 = time.perf_counter()
    grad_host.copy_(vs_psi._last_flat_grad, non_blocking=True)
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    seg.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3))
m = 1e3 * np.median(np.array(seg), axis=0)
print(f"host timeline (ms, median of 20): set parameters {m[0]:.3f} | reset {m[1]:.3f} | expect_and_grad {m[2]:.3f} | "
      f"gradient D2H + synchronize {m[3]:.3f} | total {m.sum():.3f}")
