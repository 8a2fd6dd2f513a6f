"""Oracle: the optimiser updates of the driver loop.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

NetKet's `nk.optimizer.Sgd` / `nk.optimizer.Adam` are `optax.sgd` / `optax.adam` (call site
netket_fidelity/driver/infidelity_optimizer.py:32-44 `optimizer=`, README.md:62).  optax 0.1.x `scale_by_adam`:

    mu  <- b1 mu + (1 - b1) g
    nu  <- b2 nu + (1 - b2) |g|^2            (|g|^2 = g conj(g) for complex leaves)
    mu^ = mu / (1 - b1^t),  nu^ = nu / (1 - b2^t)
    p   <- p - lr mu^ / (sqrt(nu^ + eps_root) + eps),  eps_root = 0

parity unpinned (optax is not in this image): restated from the published algorithm (Kingma & Ba, Alg. 1, with
optax's placement of eps outside the square root).
"""
import numpy as np


def sgd_step(p, g, lr):
    return p - lr * g


class AdamState:
    def __init__(self, p):
        self.mu = np.zeros_like(p)
        self.nu = np.zeros(p.shape, dtype=np.float64)
        self.t = 0


def adam_step(p, g, st: AdamState, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    st.t += 1
    st.mu = b1 * st.mu + (1 - b1) * g
    st.nu = b2 * st.nu + (1 - b2) * (g * np.conj(g)).real
    mu_hat = st.mu / (1 - b1 ** st.t)
    nu_hat = st.nu / (1 - b2 ** st.t)
    return p - lr * mu_hat / (np.sqrt(nu_hat) + eps)
