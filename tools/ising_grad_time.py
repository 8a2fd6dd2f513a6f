import sys, os, numpy as np, torch
sys.path.insert(0, "/root/repo")
from netket_fidelity_b200 import _native as nv
from netket_fidelity_b200.workloads import random_rbm
from oracle import operators as oops
from tests._helpers import to_opspec
N, M, S = 40, 80, 1 << 18
dev = torch.device("cuda", 0)
h = nv.get_handle(0)
a = nv.RbmSpec(*[t.to(dev) for t in random_rbm(N, M, 1234)])
b = nv.RbmSpec(*[t.to(dev) for t in random_rbm(N, M, 5678)])
U = oops.IsingPropagator(N, oops.chain_edges(N), 1.0, 1.0, 0.01)
u, ud = to_opspec(U), to_opspec(U.H)
W32 = (N + 31) // 32
g = torch.Generator(device=dev); g.manual_seed(1)
sp = torch.randint(-2**31, 2**31 - 1, (S, W32), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
mask = torch.tensor([-1, (1 << (N - 32)) - 1], dtype=torch.int32, device=dev)
sp &= mask
ep = torch.roll(sp, 7, 0).contiguous()
LS = (-1.0, 1.0)
for it in range(3):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    lv, L, rho1, sums = h.infidelity_fwd(a, b, sp, ep, LS, -0.5, u, ud, None)
    e[1].record()
    gacc = h.infidelity_grad(a, sp, ep, LS, -0.5, lv, rho1, sums, ud)
    e[2].record()
    torch.cuda.synchronize()
    print(f"C3-shape upsi with Ising U, U^dagger: fwd {e[0].elapsed_time(e[1]):.2f} ms, grad {e[1].elapsed_time(e[2]):.2f} ms, F = {float(sums[0] / sums[4]):.6f}")
