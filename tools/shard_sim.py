"""Per-rank step time of the strong-scaling runs, simulated on ONE GPU: a rank of a G-GPU job owns n_chains / G chains
of each family, so its step (without the two small all-reduces) is the 1-GPU step on that many chains.
usage: python tools/shard_sim.py [workload]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from netket_fidelity_b200 import _native as nv  # noqa: E402
from netket_fidelity_b200.engine import InfidelityStep, StepConfig  # noqa: E402
from netket_fidelity_b200.workloads import make_workload, device_ops, random_rbm  # noqa: E402

w = make_workload(sys.argv[1] if len(sys.argv) > 1 else "C4")
N, M = w["N"], w["M"]
dev = torch.device("cuda", 0)
U, Ud = device_ops(w, dev)
specs = {k: nv.RbmSpec(*[t.to(dev) if t is not None else None for t in random_rbm(N, M, seed, torch.complex64, w["vb"])])
         for k, seed in (("psi", 1234), ("phi", 5678))}
flush = torch.empty((160 << 20,), dtype=torch.uint8, device=dev)
base = None
for G in [int(g) for g in os.environ.get("SHARD_G", "1,2,4,8").split(",")]:
    cfg = StepConfig(w["n_chains"] // G, w["chain_length"], 5, N, seed=20261019, cv_coeff=-0.5)
    eng = InfidelityStep(N, M, cfg, w["mode"], U, Ud, device=dev)
    for _ in range(3):
        eng.step(specs["psi"], specs["phi"])
    torch.cuda.synchronize()
    ts = []
    for _ in range(4):
        flush.zero_()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        eng.sample(specs["psi"], specs["phi"])
        e1.record()
        eng.estimate(specs["psi"], specs["phi"])
        e2.record()
        torch.cuda.synchronize()
        ts.append((e0.elapsed_time(e2), e0.elapsed_time(e1)))
    t, ts_ = min(ts)
    base = base or t
    print(f"G={G}: {w['n_chains'] // G} chains/family/rank  step {t:.2f} ms (sampling {ts_:.2f})  speed-up {base / t:.2f}x",
          flush=True)
