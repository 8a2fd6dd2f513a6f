"""Times the two tcgen05 kernels of a C4 step alone (pass A estimator, pass B gradient) on fixed sampled batches, and
checks them against the CUDA-core kernels on the same samples (max-norm errors).  usage: python tools/tc_time.py [S]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from netket_fidelity_b200 import _native as nv  # noqa: E402
from netket_fidelity_b200.workloads import make_workload, device_ops, random_rbm  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
w = make_workload("C4")
N, M = w["N"], w["M"]
dev = torch.device("cuda", 0)
h = nv.get_handle(0)
U, Ud = device_ops(w, dev)
psi, phi = (nv.RbmSpec(*[t.to(dev) if t is not None else None for t in random_rbm(N, M, seed, torch.complex64, w["vb"])])
            for seed in (1234, 5678))
ls = (-1.0, 1.0)
W32 = (N + 31) // 32
n_chains = 16384
cl = S // n_chains
st = torch.zeros((n_chains, W32), dtype=torch.int32, device=dev)
sig = h.sample(psi, st, cl, 5, N, 1, ls, init_random=True).reshape(-1, W32)
st2 = torch.zeros((n_chains, W32), dtype=torch.int32, device=dev)
eta = h.sample(phi, st2, cl, 5, N, 2, ls, init_random=True, chain_offset=n_chains).reshape(-1, W32)
flush = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)


def run(impl, reps):
    os.environ["NKFB_FWD_IMPL"] = impl
    os.environ["NKFB_GRAD_IMPL"] = impl
    tf, tg, out = [], [], None
    for _ in range(reps):
        flush.zero_()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        lv, L, rho1, sums = h.infidelity_fwd(psi, phi, sig, eta, ls, -0.5, U, Ud, None)
        e[1].record()
        gacc = h.infidelity_grad(psi, sig, eta, ls, -0.5, lv, rho1, sums, Ud)
        e[2].record()
        torch.cuda.synchronize()
        tf.append(e[0].elapsed_time(e[1]))
        tg.append(e[1].elapsed_time(e[2]))
        out = (lv, L, sums, h.grad_finalize(gacc, sums, psi.n_params, torch.complex64))
    return min(tf), min(tg), out


tf, tg, tc = run("tc", 6)
print(f"tcgen05: estimator {tf:.3f} ms  gradient {tg:.3f} ms  (S = {S})")
sf, sg, si = run("simt", 2)
print(f"CUDA cores: estimator {sf:.3f} ms  gradient {sg:.3f} ms")
Lt, Ls = tc[1].double(), si[1].double()
dl = torch.exp(tc[0] - si[0]) - 1
print(f"per-sample L: max |dL| / max|L| = {float((Lt - Ls).abs().max() / Ls.abs().max()):.3e}  mean = {float((Lt - Ls).abs().mean()):.3e};"
      f"  max |exp(dl) - 1| = {float(dl.abs().max()):.3e}")
st_, ss_ = tc[2].cpu().numpy(), si[2].cpu().numpy()
print(f"F: tc {st_[0] / st_[4]:.9f} simt {ss_[0] / ss_[4]:.9f}  diff {abs(st_[0] / st_[4] - ss_[0] / ss_[4]):.3e}")
gt, gs = tc[3], si[3]
print(f"gradient: max |dg| / max|g| = {float((gt - gs).abs().max() / gs.abs().max()):.3e}")
os.environ.pop("NKFB_FWD_IMPL")
os.environ.pop("NKFB_GRAD_IMPL")
# complex128 kernels on the same samples and the same (complex64-representable) parameters: the reference both fp32 forms are held to
psi64, phi64 = (nv.RbmSpec(*[t.to(dev).to(torch.complex128) if t is not None else None
                             for t in random_rbm(N, M, seed, torch.complex64, w["vb"])]) for seed in (1234, 5678))
lv, L64, rho1, sums = h.infidelity_fwd(psi64, phi64, sig, eta, ls, -0.5, U, Ud, None)
gacc = h.infidelity_grad(psi64, sig, eta, ls, -0.5, lv, rho1, sums, Ud)
g64 = h.grad_finalize(gacc, sums, psi64.n_params, torch.complex128)
s64 = sums.cpu().numpy()
for name, o in (("tcgen05", tc), ("CUDA cores", si)):
    so = o[2].cpu().numpy()
    print(f"{name} vs complex128 kernels: L max {float((o[1].double() - L64).abs().max() / L64.abs().max()):.3e}  "
          f"F {abs(so[0] / so[4] - s64[0] / s64[4]):.3e}  gradient max-norm {float((o[3].to(torch.complex128) - g64).abs().max() / g64.abs().max()):.3e}")
