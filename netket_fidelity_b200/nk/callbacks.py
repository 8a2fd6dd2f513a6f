"""ConvergenceStopping (NetKet `nk.callbacks.ConvergenceStopping`), used by
InfidelityOptimizer.run(target_infidelity=...) with smoothing_window=20, patience=30
(reference driver/infidelity_optimizer.py:182-186)."""
from collections import deque


class ConvergenceStopping:
    def __init__(self, target, monitor="mean", *, smoothing_window=10, patience=10):
        self.target, self.monitor = target, monitor
        self.smoothing_window, self.patience = smoothing_window, patience
        self._loss_window = deque([], maxlen=smoothing_window)
        self._patience_counter = 0

    def __call__(self, step, log_data, driver):
        loss = getattr(log_data[driver._loss_name], self.monitor)
        loss = loss.real if isinstance(loss, complex) else loss
        self._loss_window.append(loss)
        loss_smooth = sum(self._loss_window) / len(self._loss_window)
        if loss_smooth <= self.target:
            self._patience_counter += 1
        else:
            self._patience_counter = 0
        return self._patience_counter <= self.patience
