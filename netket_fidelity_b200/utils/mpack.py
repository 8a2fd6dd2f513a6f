"""Parameter I/O in NetKet's format: Flax msgpack (`flax.serialization.to_bytes/from_bytes`), where an
ndarray is msgpack ext type 1 holding (shape, dtype name, raw bytes).  The reference's examples load
initial states this way (examples/adiabatic_sweeping/adiabatic_sweeping.py:34-38) and NetKet's JsonLog
writes `.mpack` files."""

from __future__ import annotations

import msgpack
import numpy as np
import torch


def _ext_hook(code, data):
    if code == 1:
        shape, dtype_name, buf = msgpack.unpackb(data, raw=False)
        return np.frombuffer(buf, dtype=np.dtype(dtype_name)).reshape(shape).copy()
    if code == 2:  # native complex scalar
        re, im = msgpack.unpackb(data)
        return complex(re, im)
    return msgpack.ExtType(code, data)


def _default(obj):
    if torch.is_tensor(obj):
        obj = obj.detach().cpu().numpy()
    if isinstance(obj, np.ndarray):
        obj = np.ascontiguousarray(obj)
        return msgpack.ExtType(1, msgpack.packb((list(obj.shape), obj.dtype.name, obj.tobytes()), use_bin_type=True))
    if isinstance(obj, (np.floating, np.integer)):
        return obj.item()
    if isinstance(obj, complex):
        return msgpack.ExtType(2, msgpack.packb((obj.real, obj.imag)))
    raise TypeError(type(obj))


def load_mpack(path):
    with open(path, "rb") as f:
        return msgpack.unpackb(f.read(), ext_hook=_ext_hook, raw=False, strict_map_key=False)


def save_parameters(path, variables):
    tree = {k: v for k, v in variables.items() if k != "unitary"}
    with open(path, "wb") as f:
        f.write(msgpack.packb(tree, default=_default, use_bin_type=True))


def load_parameters(path):
    """Returns the `variables`-like dict ({"params": {...}}) stored in a NetKet `.mpack` file."""
    return load_mpack(path)
