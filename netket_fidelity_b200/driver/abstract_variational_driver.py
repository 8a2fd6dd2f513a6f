"""Minimal `AbstractVariationalDriver` (NetKet): the run/iter/update loop that
InfidelityOptimizer inherits in the reference (driver/infidelity_optimizer.py:3,31)."""

from __future__ import annotations

import sys

from ..nk.logging import JsonLog, RuntimeLog


def _to_iterable(x):
    if x is None:
        return ()
    if hasattr(x, "__iter__") and not isinstance(x, (str, bytes)):
        return tuple(x)
    return (x,)


class AbstractVariationalDriver:
    def __init__(self, variational_state, optimizer, minimized_quantity_name=""):
        self._loss_name = minimized_quantity_name
        self._variational_state = variational_state
        self.optimizer = optimizer
        self._step_count = 0
        self._loss_stats = None

    state = property(lambda self: self._variational_state)
    step_count = property(lambda self: self._step_count)

    def _forward_and_backward(self):
        raise NotImplementedError

    def reset(self):
        self.state.reset()
        self._step_count = 0

    def update_parameters(self, dp_flat):
        """optax update + apply (NetKet `apply_gradient`), as one CUDA kernel on the flat buffer."""
        self.optimizer.apply(self.state, dp_flat)

    def iter(self, n_steps, step=1):
        for _ in range(0, n_steps, step):
            for i in range(step):
                dp = self._forward_and_backward()
                if i == 0:
                    yield self.step_count
                self._step_count += 1
                self.update_parameters(dp)

    def advance(self, steps=1):
        for _ in self.iter(steps):
            pass

    def estimate(self, observables):
        return {k: self.state.expect(o) for k, o in observables.items()}

    def run(self, n_iter, out=None, obs=None, step_size=1, show_progress=True, save_params_every=50,
            write_every=50, callback=lambda *x: True):
        if not isinstance(n_iter, int) or n_iter < 0:
            raise ValueError("n_iter, the first positional argument to `run`, must be a non-negative integer.")
        obs = obs or {}
        loggers = tuple(JsonLog(o, write_every=write_every) if isinstance(o, str) else o for o in _to_iterable(out))
        callbacks = _to_iterable(callback)
        bar = None
        if show_progress and sys.stderr.isatty():
            try:
                from tqdm import tqdm
                bar = tqdm(total=n_iter)
            except Exception:
                bar = None
        for step in self.iter(n_iter, step_size):
            log_data = self.estimate(obs)
            if self._loss_stats is not None:
                log_data[self._loss_name] = self._loss_stats
            cont = True
            for cb in callbacks:
                if not cb(step, log_data, self):
                    cont = False
            for lg in loggers:
                lg(self.step_count, log_data, self.state)
            if bar is not None:
                bar.update(step_size)
                if self._loss_stats is not None:
                    bar.set_postfix_str(f"{self._loss_name}={self._loss_stats}")
            if not cont:
                break
        for lg in loggers:
            lg.flush(self.state)
        if bar is not None:
            bar.close()
        return loggers
