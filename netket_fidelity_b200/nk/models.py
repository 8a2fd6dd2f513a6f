"""RBM (NetKet `nk.models.RBM`; verbatim form in the reference at
examples/twospins_GHZ/RBM_Jastrow_ansatz.py:37-62):
    log psi(x) = sum_j log_cosh(b_j + sum_i x_i W_ij) + sum_i a_i x_i
Parameter tree {"Dense": {"kernel" (N, alpha N), "bias" (alpha N)}, "visible_bias" (N)}."""
import numpy as np
import torch


class RBM:
    def __init__(self, alpha=1, param_dtype=float, use_hidden_bias=True, use_visible_bias=True, stddev=0.01):
        self.alpha = alpha
        self.param_dtype = param_dtype
        self.use_hidden_bias = use_hidden_bias
        self.use_visible_bias = use_visible_bias
        self.stddev = stddev

    @property
    def is_complex(self):
        return np.issubdtype(np.dtype(self.param_dtype), np.complexfloating)

    def torch_dtype(self):
        return {np.dtype("complex128"): torch.complex128, np.dtype("complex64"): torch.complex64,
                np.dtype("float64"): torch.float64, np.dtype("float32"): torch.float32}[np.dtype(self.param_dtype)]

    def init(self, seed, N, device):
        """normal(stddev) per real/imag part (Flax `normal`), NumPy default_rng(seed)."""
        rng = np.random.default_rng(seed)
        M = int(self.alpha * N)
        td = self.torch_dtype()

        def draw(shape):
            v = rng.normal(0, self.stddev, shape)
            if self.is_complex:
                v = v + 1j * rng.normal(0, self.stddev, shape)
            return torch.tensor(v).to(td).to(device).contiguous()

        params = {"Dense": {"kernel": draw((N, M))}}
        if self.use_hidden_bias:
            params["Dense"]["bias"] = draw((M,))
        if self.use_visible_bias:
            params["visible_bias"] = draw((N,))
        return params

    def __repr__(self):
        return f"RBM(alpha={self.alpha}, param_dtype={np.dtype(self.param_dtype)})"


class LogStateVector:
    """NetKet `nk.models.LogStateVector(hilbert)`: a state given by its 2^N log-amplitudes (parameter "logstate").
    Used as a full-summation TARGET (examples/state_learning/state_learning.py:19-30)."""

    def __init__(self, hilbert=None, param_dtype=complex):
        self.hilbert = hilbert
        self.param_dtype = param_dtype

    is_complex = True

    def __repr__(self):
        return "LogStateVector()"
