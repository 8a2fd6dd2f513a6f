"""MCState (NetKet `nk.vqs.MCState`) for an RBM ansatz, backed by the CUDA kernels.

Only what the infidelity hot path uses: `.samples`, `.reset()`, `.parameters` (get/set),
`.variables`, `.model_state`, `.n_samples`, `.chain_length`, `.n_discard_per_chain`,
`.log_value(x)`, `.expect(op)`, `.expect_and_grad(op)`, `.to_array()` (small N).

Parameters live in ONE flat complex device buffer laid out [bias | kernel | visible_bias]
(the jax leaf order of the NetKet RBM parameter dict, which is also the layout of the
gradient returned by the kernels); `.parameters` returns views into it.  A real-parameter
model (`param_dtype=float`) keeps zero imaginary parts; its gradient is the real part of
the complex-convention gradient.

`variables["unitary"]` (reference: utils/sampling_Ustate.py:7-31) turns the state into
U|phi>: log_value, sampling and the estimator then use log sum_c mel_c phi(x'_c).
"""

from __future__ import annotations

import itertools
import os

import numpy as np
import torch

from .. import _native as nv
from .. import parallel as par
from .stats import Stats, stats_from_partials

_readback_streams = {}


def _readback_stream(dev):
    """One side stream per device for the asynchronous read-back of the chain statistics (`MCState._stats_begin`)."""
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _readback_streams:
        _readback_streams[key] = torch.cuda.Stream(device=dev)
    return _readback_streams[key]

_seed_counter = itertools.count()


def _world():
    d = torch.distributed
    if d.is_available() and d.is_initialized():
        return d.get_world_size(), d.get_rank()
    return 1, 0


class _RBMStateBase:
    """Parameter storage and amplitudes shared by MCState and FullSumState (flat complex device buffer)."""

    def _init_params(self, hilbert, model, variables, seed, device):
        if not torch.cuda.is_available():
            raise nv.NkfbError("no CUDA device: netket_fidelity_b200 has no CPU fallback")
        self._h = nv.get_handle(device)
        self.device = self._h.device
        self.hilbert = hilbert
        self.model = model
        N = hilbert.size
        self._N = N
        self._ls = tuple(hilbert.local_states)
        if seed is None:
            # default seed: drawn on rank 0 and broadcast, so that every rank initialises the SAME parameters
            seed = par.broadcast_seed(int.from_bytes(os.urandom(4), "little") + next(_seed_counter))
        self._seed = int(seed)
        self._unitary = None
        if variables is not None:
            params = variables["params"]
            self._unitary = variables.get("unitary", None)
            if model is None:
                raise ValueError("`model` is needed to interpret `variables`")
        else:
            params = model.init(self._seed, N, self.device)
        self._M = params["Dense"]["kernel"].shape[1]
        self._W32 = (N + 31) // 32
        self._alloc_flat(params)

    def _params_changed(self):
        pass

    # ------------------------------------------------------------------ parameters
    def _alloc_flat(self, params):
        N, M = self._N, self._M
        cdt = {torch.float64: torch.complex128, torch.float32: torch.complex64,
               torch.complex128: torch.complex128, torch.complex64: torch.complex64}[params["Dense"]["kernel"].dtype]
        self._cdtype = cdt
        self._flat = torch.zeros((M + N * M + N,), dtype=cdt, device=self.device)
        self._has_bias = "bias" in params["Dense"]
        self._has_vbias = "visible_bias" in params
        self._set_params(params)

    def _views(self):
        N, M = self._N, self._M
        b = self._flat[:M]
        k = self._flat[M:M + N * M].view(N, M)
        a = self._flat[M + N * M:]
        return k, b, a

    def _set_params(self, params):
        # straight into the flat device buffer; asynchronous when the source is pinned host memory (stream-ordered
        # before the kernels that read it), an ordinary blocking copy otherwise
        k, b, a = self._views()
        k.copy_(torch.as_tensor(params["Dense"]["kernel"]), non_blocking=True)
        if self._has_bias:
            b.copy_(torch.as_tensor(params["Dense"]["bias"]), non_blocking=True)
        if self._has_vbias:
            a.copy_(torch.as_tensor(params["visible_bias"]), non_blocking=True)
        self._params_changed()

    def _spec(self) -> nv.RbmSpec:
        k, b, a = self._views()
        return nv.RbmSpec(k, b if self._has_bias else None, a if self._has_vbias else None)

    def _tree(self, flat):
        """(P,) complex flat vector -> parameter-shaped dict (real part for real-parameter models)."""
        N, M = self._N, self._M
        real = not self.model.is_complex
        f = (lambda t: t.real) if real else (lambda t: t)
        out = {"Dense": {"kernel": f(flat[M:M + N * M].view(N, M))}}
        if self._has_bias:
            out["Dense"]["bias"] = f(flat[:M])
        if self._has_vbias:
            out["visible_bias"] = f(flat[M + N * M:])
        return out

    def _flatten_tree(self, tree):
        """parameter-shaped dict -> (P,) complex flat device vector (inverse of `_tree`)."""
        N, M = self._N, self._M
        flat = torch.zeros_like(self._flat)
        flat[M:M + N * M].view(N, M).copy_(torch.as_tensor(tree["Dense"]["kernel"]).to(self.device))
        if self._has_bias:
            flat[:M].copy_(torch.as_tensor(tree["Dense"]["bias"]).to(self.device))
        if self._has_vbias:
            flat[M + N * M:].copy_(torch.as_tensor(tree["visible_bias"]).to(self.device))
        return flat

    @property
    def parameters(self):
        return self._tree(self._flat)

    @parameters.setter
    def parameters(self, params):
        self._set_params(params)

    @property
    def model_state(self):
        return {} if self._unitary is None else {"unitary": self._unitary}

    @property
    def variables(self):
        return {"params": self.parameters, **self.model_state}

    @variables.setter
    def variables(self, v):
        self._unitary = v.get("unitary", None)
        self._set_params(v["params"])

    @property
    def n_parameters(self):
        n = self._N * self._M
        return n + (self._M if self._has_bias else 0) + (self._N if self._has_vbias else 0)

    def _op_spec(self):
        return None if self._unitary is None else self._unitary._spec(self.device)

    # ------------------------------------------------------------------ amplitudes
    def log_value(self, x):
        """log psi(x) (or log (U phi)(x) when a unitary is attached).  x: (..., N) tensor/array of local states."""
        x = torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x).to(self.device)
        shape = x.shape[:-1]
        xf = x.reshape(-1, self._N).to(torch.float64).contiguous()
        p = self._h.pack(xf, self._ls)
        return self._h.logpsi(self._spec(), p, self._ls, self._op_spec()).view(shape)

    def to_array(self, normalize=True):
        if self._N > 24:
            raise ValueError("to_array: Hilbert space too large")
        st = torch.tensor(self.hilbert.all_states(), dtype=torch.float64, device=self.device)
        lv = self.log_value(st).to(torch.complex128)
        lv = lv - lv.real.max()
        psi = torch.exp(lv)
        if normalize:
            psi = psi / torch.linalg.vector_norm(psi)
        return psi


class MCState(_RBMStateBase):
    def __init__(self, sampler, model=None, *, n_samples=None, n_samples_per_rank=None, n_discard_per_chain=None,
                 chunk_size=None, variables=None, seed=None, sampler_seed=None, device=None, apply_fun=None):
        if chunk_size is not None:
            raise NotImplementedError("chunk_size: the CUDA kernels tile the samples internally")
        if apply_fun is not None:
            raise TypeError("only the RBM ansatz is implemented on the GPU hot path; pass `model=`")
        if not torch.cuda.is_available():
            raise nv.NkfbError("no CUDA device: netket_fidelity_b200 has no CPU fallback")
        self._h = nv.get_handle(device)
        self.device = self._h.device
        self.sampler = sampler
        self.hilbert = sampler.hilbert
        self.model = model
        self._world, self._rank = _world()
        N = self.hilbert.size
        self._N = N
        self._ls = tuple(self.hilbert.local_states)

        # --- sample counts: n_samples rounded up to a multiple of the total number of chains
        n_chains = sampler.n_chains
        if n_samples is not None and n_samples_per_rank is not None:
            raise ValueError("Only one argument between `n_samples` and `n_samples_per_rank` can be specified")
        if n_samples_per_rank is not None:
            n_samples = int(n_samples_per_rank) * self._world
        if n_samples is None:
            n_samples = 1000
        self._chain_length = max(1, -(-int(n_samples) // n_chains))
        self._n_chains = n_chains
        self._n_chains_local = sampler.n_chains_per_rank
        self.n_discard_per_chain = 5 if n_discard_per_chain is None else int(n_discard_per_chain)

        # --- parameters
        if seed is None:
            # default seed: drawn on rank 0 and broadcast, so that every rank initialises the SAME parameters
            seed = par.broadcast_seed(int.from_bytes(os.urandom(4), "little") + next(_seed_counter))
        self._seed = int(seed)
        self._sampler_seed = int(sampler_seed) if sampler_seed is not None else (self._seed * 2654435761 + 12345) % (1 << 63)
        self._unitary = None
        if variables is not None:
            params = variables["params"]
            self._unitary = variables.get("unitary", None)
            if model is None:
                raise ValueError("`model` is needed to interpret `variables`")
        else:
            params = model.init(self._seed, N, self.device)
        self._M = params["Dense"]["kernel"].shape[1]
        self._alloc_flat(params)

        # --- chains
        self._W32 = (N + 31) // 32
        self._chain_state = torch.zeros((self._n_chains_local, self._W32), dtype=torch.int32, device=self.device)
        self._chains_initialised = False
        self._steps_done = 0
        self._samples_packed = None
        self._n_accepted = torch.zeros((1,), dtype=torch.int64, device=self.device)
        self._n_steps_total = 0
        self._ws = None

    def _params_changed(self):
        self._samples_packed = None

    # ------------------------------------------------------------------ sampling
    @property
    def n_samples(self):
        return self._chain_length * self._n_chains

    @property
    def n_samples_per_rank(self):
        return self._chain_length * self._n_chains_local

    @property
    def chain_length(self):
        return self._chain_length

    def reset(self):
        """Drop the sample cache; chain positions are kept (NetKet `MCState.reset`)."""
        self._samples_packed = None

    def sample(self, *, chain_length=None, n_discard_per_chain=None):
        cl = self._chain_length if chain_length is None else int(chain_length)
        nd = self.n_discard_per_chain if n_discard_per_chain is None else int(n_discard_per_chain)
        net = self._spec()
        if self._ws is None:
            self._ws = self._h.sample_workspace(net)
        out = self._h.sample(net, self._chain_state, cl, nd, self.sampler.sweep_size, self._sampler_seed, self._ls,
                             op=self._op_spec(), chain_offset=self._rank * self._n_chains_local,
                             step_offset=self._steps_done, site_mode=0 if self.sampler.site_mode == "shared" else 1,
                             init_random=not self._chains_initialised, n_accepted=self._n_accepted,
                             workspace=self._ws)
        self._chains_initialised = True
        self._steps_done += (cl + nd) * self.sampler.sweep_size
        self._n_steps_total += (cl + nd) * self.sampler.sweep_size * self._n_chains_local
        self._samples_packed = out
        return out

    def _sample_segment(self, seg_len, n_discard, out, tables_valid):
        """One segment of a sampling pass split over several nkfb_sample calls (infidelity/expect.py, streamed pass A): the
        chains continue from their device-resident configurations exactly as from one `sample()` to the next; `out`
        (n_chains_local, seg_len, W32) receives the segment.  The caller assembles `_samples_packed`."""
        net = self._spec()
        if self._ws is None:
            self._ws = self._h.sample_workspace(net)
        self._h.sample(net, self._chain_state, seg_len, n_discard, self.sampler.sweep_size, self._sampler_seed, self._ls,
                       op=self._op_spec(), chain_offset=self._rank * self._n_chains_local, step_offset=self._steps_done,
                       site_mode=0 if self.sampler.site_mode == "shared" else 1,
                       init_random=not self._chains_initialised, n_accepted=self._n_accepted, workspace=self._ws, out=out,
                       tables_valid=tables_valid)
        self._chains_initialised = True
        self._steps_done += (seg_len + n_discard) * self.sampler.sweep_size
        self._n_steps_total += (seg_len + n_discard) * self.sampler.sweep_size * self._n_chains_local

    def samples_packed(self):
        """(n_chains_local, chain_length, W32) int32: bit-packed samples, sampled lazily."""
        if self._samples_packed is None:
            self.sample()
        return self._samples_packed

    @property
    def samples(self):
        """(n_chains_local, chain_length, N) local-state values (NetKet layout)."""
        p = self.samples_packed()
        x = self._h.unpack(p.view(-1, self._W32), self._N, self._ls, torch.float64)
        return x.view(p.shape[0], p.shape[1], self._N)

    @property
    def acceptance(self):
        return float(self._n_accepted.item()) / max(1, self._n_steps_total)

    # ------------------------------------------------------------------ expectation values
    def expect(self, op):
        from ..infidelity.expect import expect_dispatch
        return expect_dispatch(self, op, return_grad=False)

    def expect_and_grad(self, op, *, mutable=None):
        from ..infidelity.expect import expect_dispatch
        return expect_dispatch(self, op, return_grad=True)

    def _stats_of(self, L, n_chains_arg, mean_override=None, sums=None):
        """nk.stats.statistics(L.reshape(n_chains_arg, -1)) with the O(S) work on the GPU."""
        return self._stats_end(self._stats_begin(L, n_chains_arg, sums), L)

    def _stats_begin(self, L, n_chains_arg, sums=None):
        """Device half of `_stats_of`: enqueues the partial-sum kernel (and the all-gather of the other ranks' rows)
        and their read-back WITHOUT waiting for either, so that a caller can put more GPU work behind them before
        `_stats_end` synchronises.  The read-back (partial sums and `sums` into pinned host memory) runs on a side
        stream ordered after the kernels above: `_stats_end` waits for that copy only -- not for what the caller has
        enqueued on the main stream in between (pass B of an infidelity step), under which the host then does the
        chain statistics."""
        S = L.numel()
        rows = int(n_chains_arg)
        cols = S // rows
        l_block = max(1, cols // 32)
        n_blocks = cols // l_block
        part = self._h.chain_stats(L, rows, cols, l_block, n_blocks)
        if self._world > 1 and sums is not None:
            # chain / block partial sums of every rank (rows x (3 + n_blocks) doubles each): the statistics below
            # are those of all chains of the job, not of this rank's shard
            from ..parallel import allgather_rows
            part = allgather_rows(part)
            rows = rows * self._world
        host = None
        if part.is_cuda:
            main = torch.cuda.current_stream(part.device)
            side = _readback_stream(part.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                part_h = self._pinned_like("part", part)
                part_h.copy_(part, non_blocking=True)
                sums_h = None
                if torch.is_tensor(sums) and sums.is_cuda:
                    sums_h = self._pinned_like("sums", sums)
                    sums_h.copy_(sums, non_blocking=True)
                    sums.record_stream(side)
                done = torch.cuda.Event()
                done.record(side)
            part.record_stream(side)
            host = (part_h, sums_h, done)
        return part, rows, cols, l_block, n_blocks, sums, host

    def _pinned_like(self, name, t):
        """A pinned host buffer of the shape / dtype of `t`, kept per state and purpose (reused from step to step: the
        previous content has been consumed by `_stats_end` before the next `_stats_begin` overwrites it)."""
        cache = self.__dict__.setdefault("_pinned_cache", {})
        b = cache.get(name)
        if b is None or b.shape != t.shape or b.dtype != t.dtype:
            b = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            cache[name] = b
        return b

    def _stats_end(self, ctx, L):
        """Host half of `_stats_of` (waits for the read-back enqueued by `_stats_begin`, not for the whole stream)."""
        part, rows, cols, l_block, n_blocks, sums, host = ctx
        S = L.numel()
        s_host = None
        if host is not None:
            part_h, sums_h, done = host
            done.synchronize()
            part = part_h.numpy()
            s_host = sums_h.numpy() if sums_h is not None else None
        else:
            part = part.cpu().numpy()
        tot, tot2 = float(part[:, 0].sum()), None
        if sums is not None:
            if s_host is not None:
                s = s_host
            else:
                s = sums.cpu().numpy() if torch.is_tensor(sums) else np.asarray(sums)
            n_glob = s[4]
            mean = s[0] / n_glob
            var = max(0.0, s[1] / n_glob - mean * mean)
            if var < 1e-10 * mean * mean:
                # converged regime (F -> 1, cv = -1/2): the variance is far below mean^2 and the one-pass form
                # sum L^2 / n - mean^2 has cancelled; recompute it centred (one more reduction, rarely taken)
                c2 = ((L.double() - mean) ** 2).sum().reshape(1)
                if self._world > 1:
                    from ..parallel import allreduce_sum_
                    allreduce_sum_(c2)
                var = float(c2.item()) / n_glob
            st = stats_from_partials(part, rows, cols, l_block, n_blocks, var * rows * cols)
            st.mean, st.variance = float(mean), float(var)
            st.n_nonfinite = int(s[5])
            if st.n_nonfinite:
                import warnings
                warnings.warn(f"{st.n_nonfinite} local estimators are NaN or infinite (diverged parameters?)",
                              RuntimeWarning, stacklevel=3)
            return st
        mean = tot / S
        tot2 = float(((L.double() - mean) ** 2).sum().item())
        return stats_from_partials(part, rows, cols, l_block, n_blocks, tot2)

    def __repr__(self):
        return (f"MCState(hilbert={self.hilbert}, sampler={self.sampler}, n_samples={self.n_samples}, "
                f"n_parameters={self.n_parameters})")


class FullSumState(_RBMStateBase):
    """Exact (full-summation) variational state, NetKet `nk.vqs.FullSumState(hilbert, model)`, for N <= 24.

    The amplitudes of all 2^N configurations come from the batched log-amplitude kernel (nkfb_logpsi); the exact
    infidelity and its gradient (reference: infidelity/overlap/exact.py:53-78, overlap_U/exact.py:62-90) are a
    handful of reductions over that vector plus ONE weighted sum of log-derivatives (nkfb_weighted_score).
    Models: the RBM (variational or target) and `LogStateVector` (a target given by its 2^N log-amplitudes,
    examples/state_learning/state_learning.py:19-30)."""

    def __init__(self, hilbert, model=None, *, variables=None, seed=None, chunk_size=None, device=None):
        from .models import LogStateVector
        if hilbert.size > 24:
            raise ValueError("FullSumState: Hilbert space too large to enumerate (N > 24)")
        self._logstate = None
        if isinstance(model, LogStateVector):
            if not torch.cuda.is_available():
                raise nv.NkfbError("no CUDA device: netket_fidelity_b200 has no CPU fallback")
            self._h = nv.get_handle(device)
            self.device = self._h.device
            self.hilbert, self.model = hilbert, model
            self._N, self._ls = hilbert.size, tuple(hilbert.local_states)
            self._W32 = (self._N + 31) // 32
            self._unitary = None
            init = variables["params"]["logstate"] if variables is not None else torch.zeros(2 ** self._N)
            self._logstate = torch.as_tensor(init).to(torch.complex128).to(self.device).clone()
        else:
            self._init_params(hilbert, model, variables, seed, device)
        self._all_packed = None

    # LogStateVector targets carry their own parameter tree
    @property
    def parameters(self):
        if self._logstate is not None:
            return {"logstate": self._logstate}
        return self._tree(self._flat)

    @parameters.setter
    def parameters(self, params):
        if self._logstate is not None:
            self._logstate = torch.as_tensor(params["logstate"]).to(torch.complex128).to(self.device).clone()
        else:
            self._set_params(params)

    @property
    def n_parameters(self):
        if self._logstate is not None:
            return self._logstate.numel()
        return _RBMStateBase.n_parameters.fget(self)

    @property
    def n_samples(self):
        return 2 ** self._N

    def reset(self):
        pass

    def all_states_packed(self):
        """(2^N, W32) int32: every configuration, in `hilbert.all_states()` order."""
        if self._all_packed is None:
            st = torch.tensor(self.hilbert.all_states(), dtype=torch.float64, device=self.device)
            self._all_packed = self._h.pack(st.contiguous(), self._ls)
        return self._all_packed

    def log_amplitudes(self):
        """log psi(x) (log (U phi)(x) with a unitary attached) for all configurations, complex128."""
        if self._logstate is not None:
            return self._logstate
        lv = self._h.logpsi(self._spec(), self.all_states_packed(), self._ls, self._op_spec())
        return lv.to(torch.complex128)

    def to_array(self, normalize=True):
        lv = self.log_amplitudes()
        psi = torch.exp(lv - lv.real.max()) if normalize else torch.exp(lv)
        if normalize:
            psi = psi / torch.linalg.vector_norm(psi)
        return psi

    def expect(self, op):
        from ..infidelity.expect import expect_dispatch
        return expect_dispatch(self, op, return_grad=False)

    def expect_and_grad(self, op, *, mutable=None):
        from ..infidelity.expect import expect_dispatch
        return expect_dispatch(self, op, return_grad=True)

    def __repr__(self):
        return f"FullSumState(hilbert={self.hilbert}, n_parameters={self.n_parameters})"
