"""ctypes binding of libnkfb200.so (the C ABI declared in include/nkfb200.h).

This module never falls back to a CPU or PyTorch implementation: if the shared library
is missing or a kernel call fails, it raises.
"""

from __future__ import annotations

import ctypes as C
import math
import os
import threading

import torch

_LIB_PATH = os.environ.get("NKFB200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libnkfb200.so")

F32, F64 = 0, 1
OP_NONE, OP_GATE, OP_ISING = 0, 1, 2

SYMBOLS = [
    "nkfb_version", "nkfb_create", "nkfb_destroy", "nkfb_last_error", "nkfb_launch_count",
    "nkfb_pack", "nkfb_unpack", "nkfb_conn", "nkfb_logpsi", "nkfb_sample", "nkfb_sample_workspace_bytes",
    "nkfb_infidelity_fwd", "nkfb_infidelity_grad", "nkfb_grad_finalize", "nkfb_chain_stats",
    "nkfb_renyi2", "nkfb_microbench", "nkfb_optimizer_update", "nkfb_set_option", "nkfb_weighted_score",
    "nkfb_score_dot", "nkfb_grad_finalize_centred", "nkfb_tc_prepare",
]

OPT_SAMPLER_ALGO = 1
OPT_SAMPLER_SHARE = 2
OPT_FWD_MAX_CTAS = 3
OPT_TC_TABLES_VALID = 4
SAMPLER_AUTO, SAMPLER_ADDITIVE, SAMPLER_MULTIPLICATIVE = 0, 1, 2


class NkfbError(RuntimeError):
    pass


class RbmT(C.Structure):
    _fields_ = [("n_visible", C.c_int32), ("n_hidden", C.c_int32), ("dtype", C.c_int32),
                ("reserved", C.c_int32), ("kernel", C.c_void_p), ("bias", C.c_void_p),
                ("visible_bias", C.c_void_p)]


class OpT(C.Structure):
    _fields_ = [("kind", C.c_int32), ("idx", C.c_int32), ("gate_mel", C.c_double * 8),
                ("c0", C.c_double * 2), ("c1", C.c_double * 2), ("c2", C.c_double * 2),
                ("n_edges", C.c_int32), ("reserved", C.c_int32), ("edges", C.c_void_p)]


class SampleCfgT(C.Structure):
    _fields_ = [("n_chains", C.c_int64), ("chain_length", C.c_int32), ("n_discard", C.c_int32),
                ("sweep_size", C.c_int32), ("site_mode", C.c_int32), ("seed", C.c_uint64),
                ("chain_offset", C.c_uint64), ("step_offset", C.c_uint64),
                ("local_states", C.c_double * 2), ("init_random", C.c_int32), ("flags", C.c_int32),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_uint64),
                ("step_offset_dev", C.c_void_p)]


_lib = None
_lib_lock = threading.Lock()


def load_library():
    """dlopen libnkfb200.so; raises if it has not been built (python __graft_entry__.py build)."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(_LIB_PATH):
            raise NkfbError(
                f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        lib = C.CDLL(_LIB_PATH)
        vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double
        P = C.POINTER
        lib.nkfb_version.restype = i32
        lib.nkfb_create.argtypes = [i32, P(vp)]
        lib.nkfb_destroy.argtypes = [vp]
        lib.nkfb_last_error.argtypes = [vp]
        lib.nkfb_last_error.restype = C.c_char_p
        lib.nkfb_launch_count.argtypes = [vp]
        lib.nkfb_launch_count.restype = i64
        lib.nkfb_pack.argtypes = [vp, vp, i32, i64, i32, P(dbl), vp, vp]
        lib.nkfb_unpack.argtypes = [vp, vp, i64, i32, P(dbl), vp, i32, vp]
        lib.nkfb_conn.argtypes = [vp, P(OpT), vp, i32, i64, i32, P(dbl), vp, vp, vp]
        lib.nkfb_logpsi.argtypes = [vp, P(RbmT), P(OpT), vp, i64, P(dbl), vp, vp]
        lib.nkfb_sample.argtypes = [vp, P(RbmT), P(OpT), P(SampleCfgT), vp, vp, vp, vp]
        lib.nkfb_sample_workspace_bytes.argtypes = [i32, i32, i32]
        lib.nkfb_sample_workspace_bytes.restype = C.c_size_t
        lib.nkfb_infidelity_fwd.argtypes = [vp, P(RbmT), P(RbmT), P(OpT), P(OpT), P(OpT), vp, vp, i64,
                                            P(dbl), dbl, vp, vp, vp, vp, vp]
        lib.nkfb_infidelity_grad.argtypes = [vp, P(RbmT), P(OpT), vp, vp, i64, P(dbl), dbl, vp, vp, vp, vp, vp]
        lib.nkfb_grad_finalize.argtypes = [vp, vp, vp, i64, i32, vp, vp]
        lib.nkfb_chain_stats.argtypes = [vp, vp, i32, i64, i64, i64, i64, vp, vp]
        lib.nkfb_renyi2.argtypes = [vp, P(RbmT), vp, i64, vp, P(dbl), vp, vp, vp, vp, vp]
        lib.nkfb_microbench.argtypes = [vp, i32, i32, P(dbl), vp]
        lib.nkfb_set_option.argtypes = [vp, i32, i64]
        lib.nkfb_tc_prepare.argtypes = [vp, P(RbmT), P(RbmT), P(dbl), vp]
        lib.nkfb_weighted_score.argtypes = [vp, P(RbmT), vp, i64, P(dbl), vp, vp, vp]
        lib.nkfb_grad_finalize_centred.argtypes = [vp, vp, vp, vp, vp, i64, i32, vp, vp]
        lib.nkfb_score_dot.argtypes = [vp, P(RbmT), vp, i64, P(dbl), vp, vp, vp]
        lib.nkfb_optimizer_update.argtypes = [vp, i32, i32, vp, vp, vp, vp, i64, dbl, dbl, dbl, dbl, i64, i32, vp]
        for name in SYMBOLS:
            getattr(lib, name)  # AttributeError if the ABI drifted
        _lib = lib
        return lib


def real_dtype_code(t: torch.dtype) -> int:
    if t in (torch.complex64, torch.float32):
        return F32
    if t in (torch.complex128, torch.float64):
        return F64
    raise TypeError(f"unsupported dtype {t}")


def complex_dtype(code: int) -> torch.dtype:
    return torch.complex64 if code == F32 else torch.complex128


def real_dtype(code: int) -> torch.dtype:
    return torch.float32 if code == F32 else torch.float64


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device=None):
    """torch's current stream ON `device` (the handle's device, not the process's current device)."""
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ls(local_states):
    return (C.c_double * 2)(float(local_states[0]), float(local_states[1]))


class OpSpec:
    """Host description of a discrete operator by its connected elements (nkfb_op_t)."""

    def __init__(self, kind, idx=0, gate_mel=None, c0=0j, c1=0j, c2=0j, edges=None):
        self.kind = kind
        self.idx = int(idx)
        self.gate_mel = gate_mel  # [[m(bit0,conn0), m(bit0,conn1)], [m(bit1,conn0), m(bit1,conn1)]]
        self.c0, self.c1, self.c2 = complex(c0), complex(c1), complex(c2)
        self.edges = edges        # int32 device tensor (E, 2) or None
        self._struct = None

    def struct(self):
        s = OpT()
        s.kind = self.kind
        s.idx = self.idx
        if self.gate_mel is not None:
            flat = []
            for b in range(2):
                for c in range(2):
                    m = complex(self.gate_mel[b][c])
                    flat += [m.real, m.imag]
            s.gate_mel = (C.c_double * 8)(*flat)
        s.c0 = (C.c_double * 2)(self.c0.real, self.c0.imag)
        s.c1 = (C.c_double * 2)(self.c1.real, self.c1.imag)
        s.c2 = (C.c_double * 2)(self.c2.real, self.c2.imag)
        if self.edges is not None:
            s.n_edges = int(self.edges.shape[0])
            s.edges = self.edges.data_ptr()
        return s

    def n_conn(self, N):
        return {OP_NONE: 1, OP_GATE: 2, OP_ISING: N + 1}[self.kind]


class RbmSpec:
    """Device-resident RBM parameters: kernel (N, M), bias (M) | None, visible_bias (N) | None,
    all complex64 or complex128 CUDA tensors (contiguous)."""

    def __init__(self, kernel, bias=None, visible_bias=None):
        assert kernel.is_cuda and kernel.is_complex() and kernel.is_contiguous()
        self.kernel, self.bias, self.visible_bias = kernel, bias, visible_bias
        for t in (bias, visible_bias):
            if t is not None:
                assert t.is_cuda and t.dtype == kernel.dtype and t.is_contiguous()
        self.N, self.M = kernel.shape
        self.code = real_dtype_code(kernel.dtype)

    def struct(self):
        s = RbmT()
        s.n_visible, s.n_hidden, s.dtype = self.N, self.M, self.code
        s.kernel = self.kernel.data_ptr()
        s.bias = None if self.bias is None else self.bias.data_ptr()
        s.visible_bias = None if self.visible_bias is None else self.visible_bias.data_ptr()
        return s

    @property
    def n_params(self):
        return self.N * self.M + self.M + self.N


class Handle:
    """One per device.  All calls run on torch's current CUDA stream of that device."""

    def __init__(self, device=None):
        lib = load_library()
        if not torch.cuda.is_available():
            raise NkfbError("no CUDA device: netket_fidelity_b200 has no CPU fallback")
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", int(device) if not isinstance(device, torch.device) else device.index or 0)
        self.lib = lib
        h = C.c_void_p()
        rc = lib.nkfb_create(self.device.index, C.byref(h))
        if rc != 0:
            raise NkfbError(f"nkfb_create(device={self.device.index}) failed with code {rc} "
                            "(-2: the device is not an sm_100 GPU; this library ships sm_100a code only)")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.nkfb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return _stream(self.device)

    def _on_device(self, *tensors):
        """Arguments must live on the handle's device: a pointer of another GPU would fault or, worse, alias."""
        for t in tensors:
            if t is not None and (not t.is_cuda or t.device.index != self.device.index):
                raise NkfbError(f"tensor on {t.device} passed to the handle of {self.device}")

    def _check(self, rc, what):
        if rc != 0:
            msg = self.lib.nkfb_last_error(self.h)
            raise NkfbError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")

    @property
    def launches(self):
        return int(self.lib.nkfb_launch_count(self.h))

    def set_option(self, key, value):
        self._check(self.lib.nkfb_set_option(self.h, int(key), int(value)), "nkfb_set_option")

    def set_sampler_algo(self, algo):
        """0 auto, 1 additive (theta tracked by additions), 2 multiplicative (z = exp(-2 theta) tracked by
        complex multiplications); see include/nkfb200.h."""
        self.set_option(OPT_SAMPLER_ALGO, algo)

    def set_sampler_share(self, n):
        """Scheduling hint: `n` sampler launches of equal size run side by side on different streams (the two
        chain families of a step: 2); see include/nkfb200.h."""
        self.set_option(OPT_SAMPLER_SHARE, n)

    def set_fwd_max_ctas(self, n):
        """Upper bound on the CTAs of the next tensor-core estimator launches (0: no limit); include/nkfb200.h."""
        self.set_option(OPT_FWD_MAX_CTAS, n)

    def set_tc_tables_valid(self, flag):
        """The tables of the tensor-core estimator / gradient kernels still match the parameter buffers (launches on the
        same buffers skip their preparation kernels); clear it before the parameters change.  include/nkfb200.h."""
        self.set_option(OPT_TC_TABLES_VALID, 1 if flag else 0)

    def tc_prepare(self, psi: "RbmSpec", phi, local_states):
        """Write the parameter tables of the tensor-core estimator (psi, phi) and gradient (psi) kernels on the current
        stream, ahead of the launches that read them (nkfb_tc_prepare); phi may be None (gradient tables only)."""
        self._on_device(psi.kernel, *([phi.kernel] if phi is not None else []))
        a = psi.struct()
        b = phi.struct() if phi is not None else None
        self._check(self.lib.nkfb_tc_prepare(self.h, C.byref(a), C.byref(b) if b is not None else None,
                                             _ls(local_states), self._stream()), "nkfb_tc_prepare")

    # ---------------------------------------------------------------- pack / unpack
    def pack(self, x, local_states):
        x = x.contiguous()
        N = x.shape[-1]
        B = x.numel() // N
        out = torch.empty((B, (N + 31) // 32), dtype=torch.int32, device=x.device)
        self._check(self.lib.nkfb_pack(self.h, _ptr(x), real_dtype_code(x.dtype), B, N, _ls(local_states),
                                       _ptr(out), self._stream()), "nkfb_pack")
        return out

    def unpack(self, packed, N, local_states, dtype=torch.float64):
        B = packed.shape[0]
        out = torch.empty((B, N), dtype=dtype, device=packed.device)
        self._check(self.lib.nkfb_unpack(self.h, _ptr(packed), B, N, _ls(local_states), _ptr(out),
                                         real_dtype_code(dtype), self._stream()), "nkfb_unpack")
        return out

    # ---------------------------------------------------------------- connected elements
    def conn(self, op: OpSpec, x, local_states):
        x = x.contiguous()
        N = x.shape[-1]
        B = x.numel() // N
        Cn = op.n_conn(N)
        xp = torch.empty((B, Cn, N), dtype=x.dtype, device=x.device)
        mels = torch.empty((B, Cn), dtype=torch.complex128, device=x.device)
        s = op.struct()
        self._check(self.lib.nkfb_conn(self.h, C.byref(s), _ptr(x), real_dtype_code(x.dtype), B, N,
                                       _ls(local_states), _ptr(xp), _ptr(mels), self._stream()), "nkfb_conn")
        return xp, mels

    # ---------------------------------------------------------------- log amplitudes
    def logpsi(self, net: RbmSpec, packed, local_states, op: OpSpec | None = None):
        B = packed.shape[0]
        self._on_device(net.kernel, packed)
        out = torch.empty((B,), dtype=net.kernel.dtype, device=packed.device)
        ns = net.struct()
        os_ = op.struct() if op is not None else None
        self._check(self.lib.nkfb_logpsi(self.h, C.byref(ns), C.byref(os_) if os_ is not None else None,
                                         _ptr(packed), B, _ls(local_states), _ptr(out), self._stream()), "nkfb_logpsi")
        return out

    # ---------------------------------------------------------------- sampler
    def sample(self, net: RbmSpec, chain_state, chain_length, n_discard, sweep_size, seed, local_states,
               op: OpSpec | None = None, chain_offset=0, step_offset=0, site_mode=0, init_random=False,
               n_accepted=None, out=None, workspace=None, step_offset_dev=None, tables_valid=False):
        """step_offset_dev: optional device int64[1] added to step_offset by the kernels (CUDA-graph replays).
        tables_valid: `workspace` still holds the tables an earlier call wrote for the same net (NKFB_SAMPLE_TABLES_VALID)."""
        n_chains, W32 = chain_state.shape
        self._on_device(net.kernel, chain_state, out, n_accepted, workspace, step_offset_dev)
        if out is None:
            out = torch.empty((n_chains, chain_length, W32), dtype=torch.int32, device=chain_state.device)
        cfg = SampleCfgT()
        cfg.n_chains, cfg.chain_length, cfg.n_discard, cfg.sweep_size = n_chains, chain_length, n_discard, sweep_size
        cfg.site_mode, cfg.seed = int(site_mode), int(seed) & 0xFFFFFFFFFFFFFFFF
        cfg.chain_offset, cfg.step_offset = int(chain_offset), int(step_offset)
        cfg.local_states = _ls(local_states)
        cfg.init_random = 1 if init_random else 0
        cfg.flags = 1 if (tables_valid and workspace is not None) else 0
        if workspace is not None:
            cfg.workspace = workspace.data_ptr()
            cfg.workspace_bytes = workspace.numel() * workspace.element_size()
        if step_offset_dev is not None:
            cfg.step_offset_dev = step_offset_dev.data_ptr()
        ns = net.struct()
        os_ = op.struct() if op is not None else None
        self._check(self.lib.nkfb_sample(self.h, C.byref(ns), C.byref(os_) if os_ is not None else None,
                                         C.byref(cfg), _ptr(chain_state), _ptr(out), _ptr(n_accepted), self._stream()),
                    "nkfb_sample")
        return out

    def sample_workspace(self, net: RbmSpec):
        """Scratch for the sampler's parameter tables (one per concurrently sampled state)."""
        nbytes = int(self.lib.nkfb_sample_workspace_bytes(net.N, net.M, net.code))
        return torch.empty((nbytes,), dtype=torch.uint8, device=self.device)

    # ---------------------------------------------------------------- estimator / gradient
    def infidelity_fwd(self, psi: RbmSpec, phi: RbmSpec, sigma, eta, local_states, cv_coeff, op1=None, op2=None,
                       op4=None, sums=None, out=None):
        """out: optional (log_val, L, rho1) buffers of S entries each (slices of larger arrays when pass A runs segment by
        segment); the sums are ACCUMULATED into `sums`."""
        S = sigma.shape[0]
        dev = sigma.device
        self._on_device(psi.kernel, phi.kernel, sigma, eta, sums)
        cd, rd = psi.kernel.dtype, real_dtype(psi.code)
        if out is not None:
            log_val, L, rho1 = out
            self._on_device(log_val, L, rho1)
            if not (log_val.is_contiguous() and L.is_contiguous() and rho1.is_contiguous()) or \
                    (log_val.shape[0], L.shape[0], rho1.shape[0]) != (S, S, S) or \
                    (log_val.dtype, L.dtype, rho1.dtype) != (cd, rd, cd):
                raise ValueError("infidelity_fwd: out buffers must be contiguous (S,) arrays of the state's dtypes")
        else:
            log_val = torch.empty((S,), dtype=cd, device=dev)
            L = torch.empty((S,), dtype=rd, device=dev)
            rho1 = torch.empty((S,), dtype=cd, device=dev)
        if sums is None:
            sums = torch.zeros((8,), dtype=torch.float64, device=dev)
        a, b = psi.struct(), phi.struct()
        o1 = op1.struct() if op1 is not None else None
        o2 = op2.struct() if op2 is not None else None
        o4 = op4.struct() if op4 is not None else None
        ref = lambda o: C.byref(o) if o is not None else None
        cv = float("nan") if cv_coeff is None else float(cv_coeff)
        self._check(self.lib.nkfb_infidelity_fwd(self.h, C.byref(a), C.byref(b), ref(o1), ref(o2), ref(o4),
                                                 _ptr(sigma), _ptr(eta), S, _ls(local_states), cv, _ptr(log_val),
                                                 _ptr(L), _ptr(rho1), _ptr(sums), self._stream()),
                    "nkfb_infidelity_fwd")
        return log_val, L, rho1, sums

    def infidelity_grad(self, psi: RbmSpec, sigma, eta, local_states, cv_coeff, log_val, rho1, sums, op2=None,
                        grad_acc=None):
        S = sigma.shape[0]
        self._on_device(psi.kernel, sigma, eta, log_val, rho1, sums, grad_acc)
        if grad_acc is None:
            grad_acc = torch.zeros((2 * psi.n_params,), dtype=torch.float64, device=sigma.device)
        a = psi.struct()
        o2 = op2.struct() if op2 is not None else None
        cv = float("nan") if cv_coeff is None else float(cv_coeff)
        self._check(self.lib.nkfb_infidelity_grad(self.h, C.byref(a), C.byref(o2) if o2 is not None else None,
                                                  _ptr(sigma), _ptr(eta), S, _ls(local_states), cv, _ptr(log_val),
                                                  _ptr(rho1), _ptr(sums), _ptr(grad_acc), self._stream()),
                    "nkfb_infidelity_grad")
        return grad_acc

    def weighted_score(self, net: RbmSpec, packed, local_states, weights, acc=None):
        """acc_k += sum_s w_s conj O_k(x_s); weights (S,) complex of the net's dtype, acc (2P,) float64."""
        S = packed.shape[0]
        if acc is None:
            acc = torch.zeros((2 * net.n_params,), dtype=torch.float64, device=packed.device)
        ns = net.struct()
        weights = weights.to(complex_dtype(net.code)).contiguous()
        self._check(self.lib.nkfb_weighted_score(self.h, C.byref(ns), _ptr(packed), S, _ls(local_states),
                                                 _ptr(weights), _ptr(acc), self._stream()), "nkfb_weighted_score")
        return acc

    def score_dot(self, net: RbmSpec, packed, local_states, vec):
        """u_s = sum_k O_k(x_s) v_k; vec (P,) complex in the flat parameter layout, returns (S,) complex."""
        S = packed.shape[0]
        self._on_device(net.kernel, packed, vec)
        vec = vec.to(complex_dtype(net.code)).contiguous()
        out = torch.empty((S,), dtype=complex_dtype(net.code), device=packed.device)
        ns = net.struct()
        self._check(self.lib.nkfb_score_dot(self.h, C.byref(ns), _ptr(packed), S, _ls(local_states), _ptr(vec),
                                            _ptr(out), self._stream()), "nkfb_score_dot")
        return out

    def grad_finalize(self, grad_acc, sums, P, dtype):
        out = torch.empty((P,), dtype=dtype, device=grad_acc.device)
        self._check(self.lib.nkfb_grad_finalize(self.h, _ptr(grad_acc), _ptr(sums), P, real_dtype_code(dtype),
                                                _ptr(out), self._stream()), "nkfb_grad_finalize")
        return out

    def grad_finalize_centred(self, acc1, acc2, sums, centre_sums, P, dtype):
        """I_grad = -(acc1 + 2 (c0 - F) acc2) / n, c0 = centre_sums[0] / centre_sums[4] (one-collective step)."""
        out = torch.empty((P,), dtype=dtype, device=acc1.device)
        self._check(self.lib.nkfb_grad_finalize_centred(self.h, _ptr(acc1), _ptr(acc2), _ptr(sums), _ptr(centre_sums), P,
                                                        real_dtype_code(dtype), _ptr(out), self._stream()),
                    "nkfb_grad_finalize_centred")
        return out

    # ---------------------------------------------------------------- statistics
    def chain_stats(self, L, rows, cols, l_block, n_blocks):
        out = torch.empty((rows, 3 + n_blocks), dtype=torch.float64, device=L.device)
        self._check(self.lib.nkfb_chain_stats(self.h, _ptr(L), real_dtype_code(L.dtype), rows, cols, l_block,
                                              n_blocks, _ptr(out), self._stream()), "nkfb_chain_stats")
        return out

    # ---------------------------------------------------------------- Renyi-2
    def renyi2(self, net: RbmSpec, samples, mask_words, local_states):
        S2, W32 = samples.shape
        S_half = S2 // 2
        dev = samples.device
        scratch = torch.empty((2 * S_half, W32), dtype=torch.int32, device=dev)
        lp = torch.empty((4 * S_half,), dtype=net.kernel.dtype, device=dev)
        q = torch.empty((S_half,), dtype=net.kernel.dtype, device=dev)
        sums = torch.zeros((8,), dtype=torch.float64, device=dev)
        ns = net.struct()
        self._check(self.lib.nkfb_renyi2(self.h, C.byref(ns), _ptr(samples), S_half, _ptr(mask_words),
                                         _ls(local_states), _ptr(scratch), _ptr(lp), _ptr(q), _ptr(sums), self._stream()),
                    "nkfb_renyi2")
        return q, sums

    # ---------------------------------------------------------------- optimiser
    def optimizer_update(self, kind, params, grad, m, v, lr, b1, b2, eps, t, real_params):
        P = params.numel()
        self._check(self.lib.nkfb_optimizer_update(self.h, kind, real_dtype_code(params.dtype), _ptr(params),
                                                   _ptr(grad), _ptr(m), _ptr(v), P, lr, b1, b2, eps, t,
                                                   1 if real_params else 0, self._stream()), "nkfb_optimizer_update")

    # ---------------------------------------------------------------- micro-benchmarks
    def microbench(self, kind, iters=4096):
        v = C.c_double()
        self._check(self.lib.nkfb_microbench(self.h, kind, iters, C.byref(v), self._stream()), "nkfb_microbench")
        return v.value


_handles = {}


def get_handle(device=None) -> Handle:
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    idx = device.index if isinstance(device, torch.device) else int(device)
    if idx not in _handles:
        _handles[idx] = Handle(idx)
    return _handles[idx]
