"""Diagonal spin observables (NetKet `nk.operator.spin.sigmaz`): linear combinations c0 + sum_i c_i sigma^z_i.
They support `+`, `-`, scalar `*` and `/` (so that `sum(sigmaz(hi, i) for i in range(N)) / N` works as in the
reference's examples) and are evaluated by `MCState.expect` as the sample mean of c0 + sum_i c_i x_i."""
from __future__ import annotations

import numpy as np


class DiagonalZ:
    def __init__(self, hilbert, coeffs, const=0.0):
        self.hilbert = hilbert
        self.coeffs = np.asarray(coeffs, dtype=np.float64)
        self.const = float(const)

    def _lin(self, other, sign):
        if isinstance(other, DiagonalZ):
            if other.hilbert != self.hilbert:
                raise TypeError("Hilbert spaces should match")
            return DiagonalZ(self.hilbert, self.coeffs + sign * other.coeffs, self.const + sign * other.const)
        if np.isscalar(other):
            return DiagonalZ(self.hilbert, self.coeffs, self.const + sign * float(other))
        return NotImplemented

    def __add__(self, other):
        return self._lin(other, 1.0)

    __radd__ = __add__

    def __sub__(self, other):
        return self._lin(other, -1.0)

    def __mul__(self, a):
        return DiagonalZ(self.hilbert, self.coeffs * float(a), self.const * float(a))

    __rmul__ = __mul__

    def __truediv__(self, a):
        return self * (1.0 / float(a))

    def __neg__(self):
        return self * -1.0

    def __repr__(self):
        return f"DiagonalZ(n_sites={int((self.coeffs != 0).sum())}, const={self.const})"


def sigmaz(hilbert, site):
    """sigma^z_site: eigenvalue = the local-state value of the site rescaled to +-1 (Spin-1/2: x_i itself)."""
    c = np.zeros(hilbert.size)
    c[site] = 1.0
    return DiagonalZ(hilbert, c)
