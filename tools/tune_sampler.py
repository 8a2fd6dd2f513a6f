"""Times the plain-target sampler launch of a workload for each tuning variant
(NKFB_SAMP_VARIANT, only compiled in with `make TUNE=-DNKFB_TUNE`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from netket_fidelity_b200 import _native as nv  # noqa: E402
from netket_fidelity_b200.workloads import make_workload, random_rbm  # noqa: E402

w = make_workload(sys.argv[1] if len(sys.argv) > 1 else "C4")
if os.environ.get("TUNE_CHAINS"):
    w["n_chains"] = int(os.environ["TUNE_CHAINS"])
variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2, 3, 4, 5, 6]
N, M = w["N"], w["M"]
h = nv.get_handle(0)
net = nv.RbmSpec(*[t.cuda() if t is not None else None for t in random_rbm(N, M, 1234)])
W32 = (N + 31) // 32
state = torch.zeros((w["n_chains"], W32), dtype=torch.int32, device="cuda")
out = torch.empty((w["n_chains"], w["chain_length"], W32), dtype=torch.int32, device="cuda")
nacc = torch.zeros((1,), dtype=torch.int64, device="cuda")
mufu = h.microbench(0, 2048)
# variants >= 100: multiplicative sampler (NKFB_RATIO_VARIANT = v - 100); below: additive sampler variants
print(f"MUFU.EX2 {mufu / 1e12:.3f} T/s  FFMA {h.microbench(3, 4096) / 1e12:.2f} T/s  "
      f"FFMA2 {h.microbench(5, 4096) / 1e12:.2f} T FMA/s", flush=True)
ref_out = None
for v in variants:
    if v >= 100:
        os.environ["NKFB_SAMPLER_ALGO"] = "2"
        os.environ["NKFB_RATIO_VARIANT"] = str(v - 100)
    else:
        os.environ["NKFB_SAMPLER_ALGO"] = "1"
    os.environ["NKFB_SAMP_VARIANT"] = str(v)
    ts = []
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h.sample(net, state, w["chain_length"], 5, N, 7, (-1.0, 1.0), init_random=(rep == 0), out=out,
                 n_accepted=nacc, step_offset=int(os.environ.get("TUNE_STEP_OFFSET", "0")) * rep)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = min(ts[1:])
    evals = w["n_chains"] * (5 + w["chain_length"]) * N * M
    print(f"variant {v}: {t:.3f} ms  {evals / t / 1e6:.1f} G evals/s  frac of MUFU roofline "
          f"{evals * 2.25 / (t * 1e-3) / mufu:.3f}  acc {nacc.item() / (4 * evals / M):.4f}", flush=True)
    nacc.zero_()
    # same random numbers in every variant: all but a few chains must emit identical samples
    state.zero_()
    h.sample(net, state, w["chain_length"], 5, N, 7, (-1.0, 1.0), init_random=True, out=out)
    torch.cuda.synchronize()
    if ref_out is None:
        ref_out = out.clone()
    else:
        same = (out.view(out.shape[0], -1) == ref_out.view(out.shape[0], -1)).all(dim=1).float().mean().item()
        print(f"    chains identical to the first variant: {same:.4f}", flush=True)
