"""Sgd / Adam (NetKet `nk.optimizer.Sgd`, `nk.optimizer.Adam` = optax).  The update itself is a
CUDA kernel (nkfb_optimizer_update) acting on the flat parameter buffer of an MCState."""
import torch


class _Opt:
    kind = 0

    def __init__(self, learning_rate):
        self.learning_rate = float(learning_rate)
        self._state = {}

    def apply(self, vstate, flat_grad):
        """In-place update of vstate's flat parameter buffer with the flat gradient (P,) complex."""
        raise NotImplementedError


class Sgd(_Opt):
    kind = 0

    def apply(self, vstate, flat_grad):
        vstate._h.optimizer_update(0, vstate._flat, flat_grad, None, None, self.learning_rate, 0.0, 0.0, 0.0, 1,
                                   not vstate.model.is_complex)
        vstate._params_changed()

    def __repr__(self):
        return f"Sgd(learning_rate={self.learning_rate})"


class Adam(_Opt):
    kind = 1

    def __init__(self, learning_rate=0.001, b1=0.9, b2=0.999, eps=1e-8):
        super().__init__(learning_rate)
        self.b1, self.b2, self.eps = b1, b2, eps

    def apply(self, vstate, flat_grad):
        st = self._state.get(id(vstate))
        if st is None:
            st = dict(m=torch.zeros_like(vstate._flat),
                      v=torch.zeros(vstate._flat.shape, dtype=vstate._flat.real.dtype, device=vstate._flat.device),
                      t=0)
            self._state[id(vstate)] = st
        st["t"] += 1
        vstate._h.optimizer_update(1, vstate._flat, flat_grad, st["m"], st["v"], self.learning_rate, self.b1,
                                   self.b2, self.eps, st["t"], not vstate.model.is_complex)
        vstate._params_changed()

    def __repr__(self):
        return f"Adam(learning_rate={self.learning_rate}, b1={self.b1}, b2={self.b2}, eps={self.eps})"
