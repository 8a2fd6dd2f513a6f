"""Spin-1/2 and Qubit Hilbert spaces (NetKet `nk.hilbert.Spin`, `nk.hilbert.Qubit`)."""
import numpy as np


class _TwoLevel:
    def __init__(self, N, local_states):
        self._N = int(N)
        self._ls = tuple(float(v) for v in local_states)

    @property
    def size(self):
        return self._N

    @property
    def local_states(self):
        return list(self._ls)

    @property
    def local_size(self):
        return 2

    @property
    def n_states(self):
        return 2 ** self._N

    def numbers_to_states(self, numbers):
        """Big-endian binary digits -> local states (reference test/test_operator.py:38-41)."""
        n = np.atleast_1d(np.asarray(numbers, dtype=np.int64))
        bits = (n[:, None] >> (self._N - 1 - np.arange(self._N))[None, :]) & 1
        out = np.asarray(self._ls)[bits]
        return out[0] if np.ndim(numbers) == 0 else out

    def states_to_numbers(self, x):
        x = np.asarray(x)
        bits = (x == self._ls[1]).astype(np.int64)
        return (bits << (self._N - 1 - np.arange(self._N))).sum(axis=-1)

    def all_states(self):
        return self.numbers_to_states(np.arange(self.n_states))

    def __eq__(self, o):
        return type(o) is type(self) and o._N == self._N and o._ls == self._ls

    def __hash__(self):
        return hash((type(self).__name__, self._N, self._ls))


class Spin(_TwoLevel):
    def __init__(self, s=0.5, N=1):
        if float(s) != 0.5:
            raise NotImplementedError("only spin-1/2 is on the netket_fidelity hot path")
        super().__init__(N, (-1.0, 1.0))

    def __repr__(self):
        return f"Spin(s=1/2, N={self.size})"


class Qubit(_TwoLevel):
    def __init__(self, N=1):
        super().__init__(N, (0.0, 1.0))

    def __repr__(self):
        return f"Qubit(N={self.size})"
