"""Oracle: Renyi-2 entanglement entropy.

TEST INFRASTRUCTURE ONLY.  Parity unpinned: the reference re-exports NetKet's observable
(netket_fidelity/renyi2.py:1) and its only exact helper, test/_Renyi2_exact.py:5-27 (qutip partial
trace, log2, absolute value), is imported by no test.  Restated here without qutip: the reduced
density matrix comes from reshaping the state vector.
"""
import numpy as np

from .rbm import rbm_logpsi
from .exact import to_array


def renyi2_exact(p, subsys):
    """|log2 tr rho_A^2 / (1 - 2)| (test/_Renyi2_exact.py:22-27); 0 for the empty / full subsystem."""
    N = p.N
    subsys = sorted(set(int(s) for s in subsys))
    if len(subsys) in (0, N):
        return 0.0
    psi = to_array(p, normalize=True).reshape((2,) * N)       # axis i = site i (big-endian basis ordering)
    rest = [i for i in range(N) if i not in subsys]
    m = np.transpose(psi, subsys + rest).reshape(2 ** len(subsys), -1)
    rho = m @ m.conj().T
    return float(abs(np.log2(np.trace(rho @ rho).real) / (1 - 2)))


def renyi2_swap_estimator(p, samples, subsys):
    """NetKet's MCState estimator: first half sigma, second half sigma';
    q = psi(s'_A s_B) psi(s_A s'_B) / (psi(s) psi(s')).  Returns (S2, q)."""
    x = np.asarray(samples, dtype=np.float64).reshape(-1, p.N)
    n = x.shape[0] // 2
    s, sp = x[:n], x[n:2 * n]
    mask = np.zeros(p.N, dtype=bool)
    mask[list(subsys)] = True
    a = np.where(mask, sp, s)
    b = np.where(mask, s, sp)
    q = np.exp(rbm_logpsi(p, a) + rbm_logpsi(p, b) - rbm_logpsi(p, s) - rbm_logpsi(p, sp))
    return -np.log2(q.mean().real), q
