"""Half-warp lean sampler (sampler_lean_hw.cuh) against the 32-lane lean kernel at the C4 shape: time per launch, share of
chains with identical samples (same Philox streams: only single-precision rounding of the log-ratio differs), acceptance.
usage: python tools/hw_check.py [n_chains ...]     (NKFB_HW_VARIANT variants need `make TUNE=-DNKFB_TUNE`)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from netket_fidelity_b200 import _native as nv  # noqa: E402
from netket_fidelity_b200.workloads import make_workload, random_rbm  # noqa: E402

w = make_workload("C4")
N, M = w["N"], w["M"]
h = nv.get_handle(0)
net = nv.RbmSpec(*[t.cuda() if t is not None else None for t in random_rbm(N, M, 1234)])
W32 = (N + 31) // 32
os.environ["NKFB_SAMPLER_ALGO"] = "2"
os.environ["NKFB_RATIO_FORCECTA"] = "1"
hwv = [v for v in os.environ.get("HW_VARIANTS", "0").split(",")]
for n_chains in [int(x) for x in sys.argv[1:]] or [16384, 4096, 2048]:
    state = torch.zeros((n_chains, W32), dtype=torch.int32, device="cuda")
    out = torch.empty((n_chains, w["chain_length"], W32), dtype=torch.int32, device="cuda")
    nacc = torch.zeros((1,), dtype=torch.int64, device="cuda")
    ref = None
    for name, env in [("lean32", {"NKFB_NOHW": "1"})] + [(f"hw{v}", {"NKFB_HW_VARIANT": v}) for v in hwv]:
        os.environ.pop("NKFB_NOHW", None)
        os.environ.pop("NKFB_HW_VARIANT", None)
        os.environ.update(env)
        ts = []
        for rep in range(4):
            nacc.zero_()
            state.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            h.sample(net, state, w["chain_length"], 5, N, 7, (-1.0, 1.0), init_random=True, out=out, n_accepted=nacc)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        steps = n_chains * (5 + w["chain_length"]) * N
        line = f"{n_chains} chains  {name}: {min(ts[1:]):.3f} ms  {n_chains / min(ts[1:]):.0f} chains/ms  acceptance {nacc.item() / steps:.5f}"
        if ref is None:
            ref = out.clone()
        else:
            same = (out.view(n_chains, -1) == ref.view(n_chains, -1)).all(dim=1).float().mean().item()
            line += f"  chains identical to lean32: {same:.4f}"
        print(line, flush=True)
