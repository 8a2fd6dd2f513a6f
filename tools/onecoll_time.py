import os, sys, torch
sys.path.insert(0, '/root/repo')
from netket_fidelity_b200 import _native as nv
from netket_fidelity_b200.engine import InfidelityStep, StepConfig
from netket_fidelity_b200.workloads import make_workload, device_ops, random_rbm
w = make_workload("C4"); N, M = w["N"], w["M"]; dev = torch.device("cuda", 0)
U, Ud = device_ops(w, dev)
specs = {k: nv.RbmSpec(*[t.to(dev) if t is not None else None for t in random_rbm(N, M, seed, torch.complex64, w["vb"])]) for k, seed in (("psi", 1234), ("phi", 5678))}
flush = torch.empty((160 << 20,), dtype=torch.uint8, device=dev)
for name, kw in (("default", {}), ("K=1", dict(stream_segments=1)), ("one_collective", dict(one_collective=True)), ("default again", {})):
    cfg = StepConfig(2048, 64, 5, N, seed=20261019, cv_coeff=-0.5, **kw)
    eng = InfidelityStep(N, M, cfg, w["mode"], U, Ud, device=dev)
    for _ in range(3): eng.step(specs["psi"], specs["phi"])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        flush.zero_(); eng.step(specs["psi"], specs["phi"])
    e1.record(); torch.cuda.synchronize()
    print(name, "K =", eng.K, f"{e0.elapsed_time(e1)/10:.3f} ms/step", flush=True)
