"""Ising -- drop-in for `netket.operator.IsingJax`, which the reference re-exports as
`netket_fidelity.operator.Ising` (operator/__init__.py:4):
    H = -h sum_i sx_i + J sum_<ij> sz_i sz_j,   max_conn_size = N + 1,
    get_conn_padded: row 0 = x (mel = J sum_<ij> (2[x_i == x_j] - 1)), row i+1 = x with site i
    flipped (mel = -h), rows in site order  [NetKet 3.10 semantics; parity unpinned].

`IsingPropagator` is the first-order propagator U = 1 - i dt H of BASELINE config C3
(non-unitary: use `sample_Upsi=True`); it has the same connected configurations with
mels (1 - i dt J sum_<ij>.., + i dt h).  Both travel to the kernels as three complex
constants (c0, c1, c2) and the edge list.
"""

from __future__ import annotations

import numpy as np
import torch

from .. import _native as nv
from .singlequbit_gates import DiscreteJaxOperator


class _IsingLike(DiscreteJaxOperator):
    def __init__(self, hilbert, graph, c0, c1, c2, dtype):
        super().__init__(hilbert)
        edges = graph.edges() if hasattr(graph, "edges") else list(graph)
        self._edges = np.asarray(edges, dtype=np.int32).reshape(-1, 2)
        if self._edges.size and (self._edges.min() < 0 or self._edges.max() >= hilbert.size):
            raise ValueError("graph does not fit the Hilbert space")
        self._c = (complex(c0), complex(c1), complex(c2))
        self._dtype = dtype
        self._edges_dev = {}

    dtype = property(lambda self: self._dtype)
    max_conn_size = property(lambda self: self.hilbert.size + 1)

    @property
    def edges(self):
        return self._edges

    def _spec(self, device=None):
        h = nv.get_handle(device)
        key = h.device.index
        if key not in self._edges_dev:
            self._edges_dev[key] = torch.tensor(self._edges, dtype=torch.int32, device=h.device).contiguous()
        return nv.OpSpec(nv.OP_ISING, 0, None, self._c[0], self._c[1], self._c[2], self._edges_dev[key])


class Ising(_IsingLike):
    def __init__(self, hilbert, graph, h, J=1.0, dtype=None):
        self._h, self._J = float(h), float(J)
        self._graph = graph
        super().__init__(hilbert, graph, 0.0, self._J, -self._h, np.float64 if dtype is None else dtype)

    h = property(lambda self: self._h)
    J = property(lambda self: self._J)

    @property
    def is_hermitian(self):
        return True

    @property
    def H(self):
        return self

    def propagator(self, dt):
        """U = 1 - i dt H as a DiscreteJaxOperator with the same connected configurations."""
        return IsingPropagator(self.hilbert, self._graph, self._h, self._J, dt)

    def __repr__(self):
        return f"Ising(J={self._J}, h={self._h}; n_edges={len(self._edges)})"


class IsingPropagator(_IsingLike):
    def __init__(self, hilbert, graph, h, J, dt):
        self._h, self._J, self._dt, self._graph = float(h), float(J), complex(dt), graph
        super().__init__(hilbert, graph, 1.0, -1j * self._dt * self._J, 1j * self._dt * self._h, complex)

    @property
    def H(self):
        return IsingPropagator(self.hilbert, self._graph, self._h, self._J, -np.conj(self._dt))

    def __repr__(self):
        return f"IsingPropagator(1 - i*{self._dt}*Ising(J={self._J}, h={self._h}))"
