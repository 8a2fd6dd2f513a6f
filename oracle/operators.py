"""Oracle: connected elements of the operators on the hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, in NumPy:
* `get_conns_and_mels_Rx`        netket_fidelity/operator/singlequbit_gates.py:89-105
* `get_conns_and_mels_Ry`        netket_fidelity/operator/singlequbit_gates.py:184-201
* `get_conns_and_mels_Hadamard`  netket_fidelity/operator/singlequbit_gates.py:270-288
* `.H` rules                     netket_fidelity/operator/singlequbit_gates.py:40-41,135-136,223-224
* `IsingJax.get_conn_padded` [NetKet 3.10, not in /root/reference; exported at
  netket_fidelity/operator/__init__.py:4] -- parity unpinned.
"""

from __future__ import annotations

import numpy as np


def _flip_conns(sigma, idx, local_states):
    """conns[0] = sigma, conns[1] = sigma with site idx swapped between the two
    local states (singlequbit_gates.py:93-99)."""
    sigma = np.asarray(sigma)
    assert sigma.ndim == 2
    s0 = np.asarray(local_states[0], dtype=sigma.dtype)
    s1 = np.asarray(local_states[1], dtype=sigma.dtype)
    conns = np.repeat(sigma[:, None, :], 2, axis=1)
    cur = sigma[:, idx]
    conns[:, 1, idx] = np.where(cur == s0, s1, s0)
    return conns


def conns_and_mels_rx(sigma, idx, angle, local_states):
    """singlequbit_gates.py:89-105: mels = [cos(a/2), -i sin(a/2)]."""
    conns = _flip_conns(sigma, idx, local_states)
    mels = np.zeros((conns.shape[0], 2), dtype=np.complex128)
    mels[:, 0] = np.cos(angle / 2)
    mels[:, 1] = -1j * np.sin(angle / 2)
    return conns, mels


def conns_and_mels_ry(sigma, idx, angle, local_states):
    """singlequbit_gates.py:184-201: mels = [cos(a/2), +-sin(a/2)],
    '+' iff sigma[idx] == local_states[0]."""
    conns = _flip_conns(sigma, idx, local_states)
    mels = np.zeros((conns.shape[0], 2), dtype=np.complex128)
    mels[:, 0] = np.cos(angle / 2)
    phase = np.where(conns[:, 0, idx] == local_states[0], 1, -1)
    mels[:, 1] = phase * np.sin(angle / 2)
    return conns, mels


def conns_and_mels_hadamard(sigma, idx, local_states):
    """singlequbit_gates.py:270-288: mels = [+-1/sqrt2, 1/sqrt2], float64,
    diagonal '+' iff sigma[idx] == local_states[0]."""
    conns = _flip_conns(sigma, idx, local_states)
    mels = np.zeros((conns.shape[0], 2), dtype=np.float64)
    mels[:, 1] = 1 / np.sqrt(2)
    mels[:, 0] = np.where(conns[:, 0, idx] == local_states[0], 1, -1) / np.sqrt(2)
    return conns, mels


class _Op:
    """Tiny operator protocol used throughout the oracle:
    `get_conn_padded(x[..., N]) -> (xp[..., C, N], mels[..., C])`, `max_conn_size`,
    `H`, `to_dense()`."""

    local_states = (-1.0, 1.0)

    def get_conn_padded(self, x):
        x = np.asarray(x)
        xr = x.reshape(-1, x.shape[-1])
        xp, mels = self._conn(xr)
        return (
            xp.reshape(x.shape[:-1] + xp.shape[-2:]),
            mels.reshape(x.shape[:-1] + mels.shape[-1:]),
        )

    def to_dense(self):
        """<x|O|x'> over all basis states (NetKet ordering, see exact.all_states)."""
        from .exact import all_states, states_to_numbers

        st = all_states(self.N, self.local_states)
        xp, mels = self.get_conn_padded(st)
        D = st.shape[0]
        mat = np.zeros((D, D), dtype=np.complex128)
        for c in range(xp.shape[1]):
            cols = states_to_numbers(xp[:, c, :], self.local_states)
            np.add.at(mat, (np.arange(D), cols), mels[:, c])
        return mat


class Rx(_Op):
    max_conn_size = 2

    def __init__(self, N, idx, angle, local_states=(-1.0, 1.0)):
        self.N, self.idx, self.angle, self.local_states = N, idx, angle, tuple(local_states)

    @property
    def H(self):  # singlequbit_gates.py:40-41
        return Rx(self.N, self.idx, -self.angle, self.local_states)

    def _conn(self, x):
        return conns_and_mels_rx(x, self.idx, self.angle, self.local_states)


class Ry(_Op):
    max_conn_size = 2

    def __init__(self, N, idx, angle, local_states=(-1.0, 1.0)):
        self.N, self.idx, self.angle, self.local_states = N, idx, angle, tuple(local_states)

    @property
    def H(self):  # singlequbit_gates.py:135-136 (sic: -2*angle)
        return Ry(self.N, self.idx, -self.angle * 2, self.local_states)

    def _conn(self, x):
        return conns_and_mels_ry(x, self.idx, self.angle, self.local_states)


class Hadamard(_Op):
    max_conn_size = 2

    def __init__(self, N, idx, local_states=(-1.0, 1.0)):
        self.N, self.idx, self.local_states = N, idx, tuple(local_states)

    @property
    def H(self):  # singlequbit_gates.py:223-224
        return Hadamard(self.N, self.idx, self.local_states)

    def _conn(self, x):
        return conns_and_mels_hadamard(x, self.idx, self.local_states)


class IsingLike(_Op):
    """diag = c0 + c1 * sum_<ij> (2[x_i == x_j] - 1),  off-diagonal (site i flipped) = c2.

    * `Ising(N, edges, h, J)`  : H = -h sum_i sx_i + J sum_<ij> sz_i sz_j
      -> (c0, c1, c2) = (0, J, -h)   [NetKet IsingJax, SURVEY A.5]; row 0 is the
      diagonal, row i+1 is site i flipped, rows in site order.
    * `IsingPropagator(N, edges, h, J, dt)` : U = 1 - i dt H
      -> (1, -i dt J, +i dt h)  (BASELINE config C3).
    """

    def __init__(self, N, edges, c0, c1, c2, local_states=(-1.0, 1.0)):
        self.N = N
        self.edges = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
        self.c0, self.c1, self.c2 = c0, c1, c2
        self.local_states = tuple(local_states)
        self.max_conn_size = N + 1

    @property
    def H(self):
        return IsingLike(self.N, self.edges, np.conj(self.c0), np.conj(self.c1),
                         np.conj(self.c2), self.local_states)

    def _conn(self, x):
        B, N = x.shape
        s0 = np.asarray(self.local_states[0], dtype=x.dtype)
        s1 = np.asarray(self.local_states[1], dtype=x.dtype)
        xp = np.repeat(x[:, None, :], N + 1, axis=1)
        ar = np.arange(N)
        xp[:, ar + 1, ar] = np.where(x == s0, s1, s0)
        same = x[:, self.edges[:, 0]] == x[:, self.edges[:, 1]]
        zz = (2.0 * same - 1.0).sum(axis=1)
        dt = np.result_type(np.asarray(self.c0).dtype, np.asarray(self.c1).dtype,
                            np.asarray(self.c2).dtype, np.float64)
        mels = np.empty((B, N + 1), dtype=dt)
        mels[:, 0] = self.c0 + self.c1 * zz
        mels[:, 1:] = self.c2
        return xp, mels


def Ising(N, edges, h, J=1.0, local_states=(-1.0, 1.0)):
    return IsingLike(N, edges, 0.0, float(J), -float(h), local_states)


def IsingPropagator(N, edges, h, J, dt, local_states=(-1.0, 1.0)):
    return IsingLike(N, edges, 1.0 + 0j, -1j * dt * J, 1j * dt * h, local_states)


def chain_edges(N, pbc=True):
    e = [(i, i + 1) for i in range(N - 1)]
    if pbc and N > 2:
        e.append((N - 1, 0))
    return np.asarray(e, dtype=np.int64)


def square_edges(Lx, Ly, pbc=True):
    e = set()
    for x in range(Lx):
        for y in range(Ly):
            i = x * Ly + y
            for dx, dy in ((1, 0), (0, 1)):
                xx, yy = x + dx, y + dy
                if pbc:
                    xx %= Lx
                    yy %= Ly
                elif xx >= Lx or yy >= Ly:
                    continue
                j = xx * Ly + yy
                if i != j:
                    e.add((min(i, j), max(i, j)))
    return np.asarray(sorted(e), dtype=np.int64)
