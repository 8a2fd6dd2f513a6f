"""Learning a given state with the infidelity optimiser -- the reference's
examples/state_learning/state_learning.py on the B200 kernels: the target is the ground state of a 4x4
transverse-field Ising model stored as a LogStateVector (2^16 log-amplitudes), the variational state an RBM
summed exactly (FullSumState).  The exact diagonalisation the reference delegates to `nk.exact.lanczos_ed`
is done here with scipy on the host; everything downstream runs on the GPU.

    python examples/state_learning.py --n-iter 300
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import netket_fidelity_b200 as nkf  # noqa: E402
from netket_fidelity_b200 import nk  # noqa: E402


def tfim_ground_state(hi, edges, h=1.0, J=1.0):
    """Lowest eigenvector of H = -h sum_i sx_i + J sum_<ij> sz_i sz_j in `hi.all_states()` order."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    x = np.asarray(hi.all_states())
    n, N = x.shape
    num = {tuple(r): k for k, r in enumerate(x.astype(int).tolist())}
    diag = sum(J * x[:, i] * x[:, j] for i, j in edges)
    rows, cols = [], []
    for i in range(N):
        y = x.copy()
        y[:, i] *= -1
        rows.append(np.arange(n))
        cols.append(np.array([num[tuple(r)] for r in y.astype(int).tolist()]))
    H = sp.diags(diag) + sp.csr_matrix((np.full(n * N, -h), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
    w, v = spl.eigsh(H, k=1, which="SA")
    return v[:, 0]


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--length", type=int, default=4)
    ap.add_argument("--n-iter", type=int, default=300)
    args = ap.parse_args(argv)
    g = nk.graph.Grid([args.length, args.length], pbc=True)
    hi = nk.hilbert.Spin(0.5, g.n_nodes)
    psi0 = tfim_ground_state(hi, g.edges())
    vs_target = nk.vqs.FullSumState(hi, nk.models.LogStateVector(hi))
    vs_target.parameters = {"logstate": torch.tensor(np.log(psi0 + 0j))}
    vs = nk.vqs.FullSumState(hi, nk.models.RBM(alpha=1, param_dtype=complex), seed=0)
    inf = nkf.InfidelityOperator(vs_target)
    print("initial infidelity", vs.expect(inf).mean)
    driver = nkf.driver.InfidelityOptimizer(vs_target, nk.optimizer.Adam(learning_rate=0.01), variational_state=vs)
    log = nk.logging.RuntimeLog()
    driver.run(args.n_iter, out=log, show_progress=False)
    print("final infidelity  ", vs.expect(inf).mean)
    return log


if __name__ == "__main__":
    main()
