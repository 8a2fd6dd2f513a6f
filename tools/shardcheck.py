import torch, numpy as np, sys
sys.path.insert(0, "/root/repo")
from netket_fidelity_b200 import _native as nv
from netket_fidelity_b200.workloads import make_workload, random_rbm
w = make_workload("C4"); N, M = w["N"], w["M"]
h = nv.get_handle(0)
net = nv.RbmSpec(*[t.cuda() for t in random_rbm(N, M, 1234)])
W32 = 4
def run(n_chains, off, cl=4):
    st = torch.zeros((n_chains, W32), dtype=torch.int32, device="cuda")
    return h.sample(net, st, cl, 2, N, 99, (-1.0, 1.0), chain_offset=off, init_random=True).cpu().numpy()
full = run(1024, 0)
a = run(512, 0); b = run(512, 512)
print("shard==full:", np.array_equal(np.concatenate([a, b]), full))
c = run(1024, 0)
print("repeatable:", np.array_equal(c, full))
# odd shard sizes
d = run(100, 7)
print("odd shard:", np.array_equal(d, full[7:107]))
