"""RuntimeLog / JsonLog (NetKet `nk.logging`): per-step dictionaries of Stats."""
import json


class RuntimeLog:
    def __init__(self):
        self.data = {}

    def __call__(self, step, item, variational_state=None):
        for k, v in item.items():
            d = v.to_dict() if hasattr(v, "to_dict") else {"value": v}
            slot = self.data.setdefault(k, {"iters": []})
            slot["iters"].append(step)
            for kk, vv in d.items():
                slot.setdefault(kk, []).append(vv)

    def __getitem__(self, k):
        return self.data[k]

    def flush(self, variational_state=None):
        pass


class JsonLog(RuntimeLog):
    def __init__(self, output_prefix, save_params=True, write_every=50):
        super().__init__()
        self.prefix, self.save_params, self.write_every = output_prefix, save_params, write_every
        self._n = 0

    def __call__(self, step, item, variational_state=None):
        super().__call__(step, item, variational_state)
        self._n += 1
        if self._n % self.write_every == 0:
            self.flush(variational_state)

    def flush(self, variational_state=None):
        with open(self.prefix + ".log", "w") as f:
            json.dump(self.data, f)
        if self.save_params and variational_state is not None:
            from ..utils.mpack import save_parameters
            save_parameters(self.prefix + ".mpack", variational_state.variables)
