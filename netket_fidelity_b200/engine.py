"""One infidelity-gradient step of p-tVMC on one GPU (or one rank of a multi-GPU job).

This is the device-side pipeline that `MCState.expect_and_grad(InfidelityOperator...)`
drives (reference call stack: driver/infidelity_optimizer.py:141-150 ->
infidelity/overlap/expect.py:33-55 or infidelity/overlap_U/expect.py:40-67):

    sample sigma ~ |psi|^2, eta ~ |Phi|^2      nkfb_sample      (x2)
    pass A: local estimator + CV, sums         nkfb_infidelity_fwd
    [all-reduce of 8 doubles when world_size > 1]
    pass B: score-function gradient            nkfb_infidelity_grad
    [all-reduce of the 2P-double accumulator when world_size > 1]
    I_grad = -acc / n                          nkfb_grad_finalize

Markov chains are sharded over ranks: rank r owns the global chains
[r * n_chains / G, (r+1) * n_chains / G) of both families; the Philox streams are keyed
by the GLOBAL chain id, so the samples do not depend on G.  torch.distributed is used only
as plumbing for the two all-reduces.
"""

from __future__ import annotations

import os

import torch

from . import _native as nv
from . import parallel as par


class StepConfig:
    def __init__(self, n_chains, chain_length, n_discard, sweep_size=None, seed=0, local_states=(-1.0, 1.0),
                 cv_coeff=-0.5, site_mode=0, one_collective=False, stream_segments=None):
        self.n_chains = int(n_chains)            # GLOBAL number of chains per family
        self.chain_length = int(chain_length)
        self.n_discard = int(n_discard)
        self.sweep_size = sweep_size
        self.seed = int(seed)
        self.local_states = tuple(local_states)
        self.cv_coeff = cv_coeff
        self.site_mode = site_mode
        # one all-reduce per step instead of two: pass B is centred with the PREVIOUS step's F and an
        # extra accumulator sum_s conj O_k(sigma_s) makes the result exact; costs the sigma half of pass B once more
        self.one_collective = bool(one_collective)
        # pass A streamed under the sampler: the chains are sampled in K consecutive segments and pass A of segment k
        # runs, on the SMs a sub-wave sampler launch leaves idle, while segment k + 1 is being sampled.
        # None: automatic (2 when the shard's sampler launches leave >= 16 SMs idle -- the 8-GPU shard of C4), 1: off
        self.stream_segments = stream_segments


class InfidelityStep:
    """Owns the per-rank buffers of one (psi, phi, U) problem and runs steps on them.

    mode: "standard" (Phi = phi), "standard_Uphi" (Phi = U phi, sampled from; reference
    InfidelityUPsi, overlap/operator.py:61-82) or "upsi" (unitary U; conns of U and U^dagger;
    reference InfidelityOperatorUPsi, overlap_U/operator.py:12-67).
    """

    def __init__(self, N, M, cfg: StepConfig, mode="upsi", U: nv.OpSpec | None = None,
                 U_dagger: nv.OpSpec | None = None, device=None, cdtype=torch.complex64, group=None):
        self.h = nv.get_handle(device)
        self.dev = self.h.device
        self.N, self.M, self.cfg, self.mode = N, M, cfg, mode
        self.cdtype = cdtype
        self.group = group
        self.world, self.rank = par.world_info(group)
        self.chain_offset, self.local_chains = par.shard_chains(cfg.n_chains, self.world, self.rank)
        if mode == "standard":
            self.op1 = self.op2 = self.op4 = self.op_target = None
        elif mode == "standard_Uphi":
            self.op1, self.op2, self.op4, self.op_target = U, None, U, U
        elif mode == "upsi":
            if U is None or U_dagger is None:
                raise ValueError("mode 'upsi' needs U and U_dagger")
            self.op1, self.op2, self.op4, self.op_target = U, U_dagger, None, None
        else:
            raise ValueError(mode)
        W32 = (N + 31) // 32
        self.W32 = W32
        dev = self.dev
        self.state_psi = torch.zeros((self.local_chains, W32), dtype=torch.int32, device=dev)
        self.state_phi = torch.zeros((self.local_chains, W32), dtype=torch.int32, device=dev)
        self.S_local = self.local_chains * cfg.chain_length
        self.K = self._pick_segments()
        if self.K > 1:
            # segment-major samples: [segment][chain][sample within the segment]; pass A / pass B see flat lists
            self.sigma = torch.empty((self.K, self.local_chains, cfg.chain_length // self.K, W32), dtype=torch.int32,
                                     device=dev)
            self.psi_stream = torch.cuda.Stream(device=dev, priority=-1)
            self.est_stream = torch.cuda.Stream(device=dev, priority=0)
            self._lv = torch.empty((self.S_local,), dtype=cdtype, device=dev)
            self._rho1 = torch.empty((self.S_local,), dtype=cdtype, device=dev)
            self._L = torch.empty((self.S_local,), dtype=torch.float32 if cdtype == torch.complex64 else torch.float64,
                                  device=dev)
        else:
            self.sigma = torch.empty((self.local_chains, cfg.chain_length, W32), dtype=torch.int32, device=dev)
        self.eta = torch.empty_like(self.sigma)
        self._passA_done = False
        self.S_total = cfg.n_chains * cfg.chain_length
        P = N * M + M + N
        self.P = P
        # one zero-able block: [sums (8) | grad accumulators (2P)]
        self.acc = torch.zeros((8 + 2 * P,), dtype=torch.float64, device=dev)
        self.sums = self.acc[:8]
        self.grad_acc = self.acc[8:]
        if cfg.one_collective:
            # [sums (8) | acc1 (2P) | acc2 (2P)]: the ONE buffer that crosses NVLink; centre = {c0 n_local, ..., n_local}
            self.acc_one = torch.zeros((8 + 4 * P,), dtype=torch.float64, device=dev)
            self.centre = torch.zeros((8,), dtype=torch.float64, device=dev)
            self.ones_w = torch.ones((self.S_local,), dtype=cdtype, device=dev)
        self.n_accepted = torch.zeros((2,), dtype=torch.int64, device=dev)
        # the two chain families are sampled on two streams so that the tail of one launch (a partial last
        # wave of CTAs) overlaps the other launch; each has its own scratch for the pre-scaled parameters
        self.side_stream = torch.cuda.Stream(device=dev, priority=-1 if self.K > 1 else 0)
        self.single_stream = os.environ.get("NKFB_SINGLE_STREAM", "0") == "1"
        self.ws_psi = self.ws_phi = None
        self.steps_done = 0
        self.initialised = False
        # CUDA-graph replay of a whole step (single rank): the Metropolis step counter lives on the device
        self._step_dev = None
        self._graph = None
        self._graph_out = None
        self.sweep = cfg.sweep_size if cfg.sweep_size is not None else N
        self.last = {}

    # ------------------------------------------------------------------
    def _pick_segments(self):
        cfg = self.cfg
        K = cfg.stream_segments
        if os.environ.get("NKFB_STREAM_SEGMENTS"):
            K = int(os.environ["NKFB_STREAM_SEGMENTS"])
        if cfg.one_collective:
            return 1
        if K is None:
            # automatic: single precision, the shape range of the CTA-synchronous samplers (32 chains per CTA), the LAST
            # wave of the two launches' CTAs leaves >= 16 SMs idle (at most two waves: beyond that the dispatcher mixes
            # kernels), segments large enough for the tensor-core estimator
            ok = (self.cdtype == torch.complex64 and 161 <= self.M <= 512 and self.N <= 128 and cfg.site_mode == 0
                  and self.local_chains >= 1024 and self._idle_sms() >= 16 and cfg.chain_length % 2 == 0
                  and self.S_local // 2 >= 2048)
            # two segments: every further one costs ~0.1 ms of pipeline start-up and straggler wait per sampler launch
            # (measured on the 8-GPU shard of C4: K = 1 / 2 / 4 / 8 -> 6.40 / 6.29 / 6.41 / 6.75 ms per step)
            K = 2 if ok else 1
        K = max(1, int(K))
        if K > 1 and cfg.chain_length % K != 0:
            raise ValueError(f"stream_segments = {K} does not divide chain_length = {cfg.chain_length}")
        return K

    def _idle_sms(self):
        """SMs without a sampler CTA during the last wave of the two sampler launches of this rank (one 16-warp CTA of
        32 chains per SM); 0 when the launches need more than two waves."""
        sm = torch.cuda.get_device_properties(self.dev).multi_processor_count
        ctas = 2 * ((self.local_chains + 31) // 32)
        waves = (ctas + sm - 1) // sm
        return waves * sm - ctas if waves <= 2 else 0

    def _sample_streamed(self, psi: nv.RbmSpec, phi: nv.RbmSpec):
        """K segments of chain_length / K samples per chain; the two families of a segment on two high-priority streams,
        pass A of segment k on a low-priority stream behind both (its CTAs limited to the SMs the samplers leave idle),
        concurrent with the sampling of segment k + 1.  Chains continue from segment to segment exactly as they continue
        from call to call (device-resident configurations, Philox counters keyed by the absolute step)."""
        cfg, K = self.cfg, self.K
        seg = cfg.chain_length // K
        S_seg = self.local_chains * seg
        init = not self.initialised
        steps_per_call = (cfg.n_discard + cfg.chain_length) * self.sweep
        base = self.steps_done * steps_per_call
        if self.ws_psi is None:
            self.ws_psi, self.ws_phi = self.h.sample_workspace(psi), self.h.sample_workspace(phi)
        main = torch.cuda.current_stream(self.dev)
        self.acc.zero_()
        for st in (self.psi_stream, self.side_stream, self.est_stream):
            st.wait_stream(main)
        idle = max(1, self._idle_sms())
        # the parameter tables of pass B are written ahead of pass A of the first segment (on the estimator's stream, under
        # the sampling of the second one), those of pass A by its first launch: pass A of the LAST segment and pass B --
        # the exposed ones -- launch without their preparation kernels (NKFB_OPT_TC_TABLES_VALID; estimate() clears it).  The
        # earlier pass A launches keep theirs on purpose: those few microseconds let the sampler CTAs of the next segment
        # (high-priority streams, one per SM) take their SMs before the estimator CTAs spread over the idle ones --
        # without them the estimator got there first and 20 sampler CTAs waited for it (8-GPU shard 6.02 -> 7.36 ms).
        self.h.set_sampler_share(2)
        try:
            for k in range(K):
                nd = cfg.n_discard if k == 0 else 0
                off = base + (0 if k == 0 else (cfg.n_discard + k * seg) * self.sweep)
                with torch.cuda.stream(self.psi_stream):
                    self.h.sample(psi, self.state_psi, seg, nd, self.sweep, cfg.seed, cfg.local_states, op=None,
                                  chain_offset=self.chain_offset, step_offset=off, site_mode=cfg.site_mode,
                                  init_random=init and k == 0, n_accepted=self.n_accepted[0:1], out=self.sigma[k],
                                  workspace=self.ws_psi, tables_valid=k > 0)
                with torch.cuda.stream(self.side_stream):
                    self.h.sample(phi, self.state_phi, seg, nd, self.sweep, cfg.seed, cfg.local_states,
                                  op=self.op_target, chain_offset=cfg.n_chains + self.chain_offset, step_offset=off,
                                  site_mode=cfg.site_mode, init_random=init and k == 0,
                                  n_accepted=self.n_accepted[1:2], out=self.eta[k], workspace=self.ws_phi,
                                  tables_valid=k > 0)
                self.est_stream.wait_stream(self.psi_stream)
                self.est_stream.wait_stream(self.side_stream)
                with torch.cuda.stream(self.est_stream):
                    # two estimator CTAs per idle SM while a sampler segment is still to come; the last one has the GPU
                    self.h.set_fwd_max_ctas(2 * idle if k < K - 1 else 0)
                    self.h.set_tc_tables_valid(k == K - 1)
                    if k == 0:
                        self.h.tc_prepare(psi, None, cfg.local_states)
                    sl = slice(k * S_seg, (k + 1) * S_seg)
                    self.h.infidelity_fwd(psi, phi, self.sigma[k].view(-1, self.W32), self.eta[k].view(-1, self.W32),
                                          cfg.local_states, cfg.cv_coeff, self.op1, self.op2, self.op4, sums=self.sums,
                                          out=(self._lv[sl], self._L[sl], self._rho1[sl]))
        except BaseException:
            self.h.set_tc_tables_valid(False)
            raise
        finally:
            self.h.set_sampler_share(1)
            self.h.set_fwd_max_ctas(0)
        main.wait_stream(self.est_stream)
        self._passA_done = True
        self.initialised = True
        self.steps_done += 1

    def sample(self, psi: nv.RbmSpec, phi: nv.RbmSpec):
        if self.K > 1 and self._step_dev is None:
            return self._sample_streamed(psi, phi)
        cfg = self.cfg
        init = not self.initialised
        steps_per_call = (cfg.n_discard + cfg.chain_length) * self.sweep
        step_offset = self.steps_done * steps_per_call
        if self._step_dev is not None:
            step_offset = 0          # the device counter carries it (frozen host values inside a captured graph)
        if self.ws_psi is None:
            self.ws_psi, self.ws_phi = self.h.sample_workspace(psi), self.h.sample_workspace(phi)
        main = torch.cuda.current_stream(self.dev)
        side = main if self.single_stream else self.side_stream
        if not self.single_stream:
            self.side_stream.wait_stream(main)
        # both launches are of the same size and share the SMs: the dispatcher sizes their last wave for half a GPU
        self.h.set_sampler_share(1 if self.single_stream else 2)
        # the two families use disjoint global chain ids: psi [0, n_chains), phi [n_chains, 2 n_chains)
        self.h.sample(psi, self.state_psi, cfg.chain_length, cfg.n_discard, self.sweep, cfg.seed,
                      cfg.local_states, op=None, chain_offset=self.chain_offset, step_offset=step_offset,
                      site_mode=cfg.site_mode, init_random=init, n_accepted=self.n_accepted[0:1], out=self.sigma,
                      workspace=self.ws_psi, step_offset_dev=self._step_dev)
        with torch.cuda.stream(side):
            self.h.sample(phi, self.state_phi, cfg.chain_length, cfg.n_discard, self.sweep, cfg.seed,
                          cfg.local_states, op=self.op_target, chain_offset=cfg.n_chains + self.chain_offset,
                          step_offset=step_offset, site_mode=cfg.site_mode, init_random=init,
                          n_accepted=self.n_accepted[1:2], out=self.eta, workspace=self.ws_phi,
                          step_offset_dev=self._step_dev)
        self.h.set_sampler_share(1)
        if not self.single_stream:
            main.wait_stream(self.side_stream)
        if self._step_dev is not None:
            self._step_dev += steps_per_call
        self.initialised = True
        self.steps_done += 1

    def _allreduce(self, t):
        par.allreduce_sum_(t, self.group)

    def estimate(self, psi, phi, return_grad=True, events=None):
        """Pass A (+ pass B).  Returns (sums[8] device tensor, L (S_local,), grad (P,) complex | None).
        `events`: optional 3 CUDA events recorded before pass A, after pass A, after pass B (bench.py)."""
        try:
            return self._estimate(psi, phi, return_grad, events)
        finally:
            self.h.set_tc_tables_valid(False)     # set by _sample_streamed for the launches of this step only

    def _estimate(self, psi, phi, return_grad, events):
        cfg = self.cfg
        sig = self.sigma.view(-1, self.W32)
        eta = self.eta.view(-1, self.W32)
        if cfg.one_collective and return_grad:
            return self._estimate_one_collective(psi, phi, sig, eta, events)
        if self._passA_done:
            # pass A ran segment by segment under the sampler (_sample_streamed); its sums are in self.sums
            self._passA_done = False
            log_val, L, rho1 = self._lv, self._L, self._rho1
            if events:
                events[0].record()
                events[1].record()
        else:
            self.acc.zero_()
            if events:
                events[0].record()
            log_val, L, rho1, _ = self.h.infidelity_fwd(psi, phi, sig, eta, cfg.local_states, cfg.cv_coeff, self.op1,
                                                        self.op2, self.op4, sums=self.sums)
            if events:
                events[1].record()
        self._allreduce(self.sums)
        grad = None
        if return_grad:
            self.h.infidelity_grad(psi, sig, eta, cfg.local_states, cfg.cv_coeff, log_val, rho1, self.sums,
                                   op2=self.op2, grad_acc=self.grad_acc)
            if events:
                events[2].record()
            self._allreduce(self.grad_acc)
            grad = self.h.grad_finalize(self.grad_acc, self.sums, self.P, self.cdtype)
        self.last = dict(log_val=log_val, L=L, rho1=rho1)
        return self.sums, L, grad

    def _estimate_one_collective(self, psi, phi, sig, eta, events=None):
        """Pass A, pass B centred with the previous step's F, sum_s conj O_k(sigma_s), ONE all-reduce, finalize."""
        cfg, P = self.cfg, self.P
        buf = self.acc_one
        buf.zero_()
        sums, acc1, acc2 = buf[:8], buf[8:8 + 2 * P], buf[8 + 2 * P:]
        if events:
            events[0].record()
        log_val, L, rho1, _ = self.h.infidelity_fwd(psi, phi, sig, eta, cfg.local_states, cfg.cv_coeff, self.op1,
                                                    self.op2, self.op4, sums=sums)
        if events:
            events[1].record()
        # centre[0] / centre[4] = c0 (kept on the device from the previous step; 0 before the first one)
        self.centre[4] = float(self.S_local)
        self.h.infidelity_grad(psi, sig, eta, cfg.local_states, cfg.cv_coeff, log_val, rho1, self.centre,
                               op2=self.op2, grad_acc=acc1)
        self.h.weighted_score(psi, sig, cfg.local_states, self.ones_w, acc=acc2)
        if events:
            events[2].record()
        self._allreduce(buf)                                       # the one collective of the step
        grad = self.h.grad_finalize_centred(acc1, acc2, sums, self.centre, P, self.cdtype)
        self.centre[0] = sums[0] / sums[4] * float(self.S_local)   # next step's centring constant (device op)
        self.last = dict(log_val=log_val, L=L, rho1=rho1)
        self.sums = sums
        return sums, L, grad

    def step(self, psi, phi, return_grad=True):
        """reset + sample both families + estimator (+ gradient): the unit bench.py times."""
        if self._graph is not None:
            self._graph.replay()
            self.steps_done += 1
            return self._graph_out
        self.sample(psi, phi)
        return self.estimate(psi, phi, return_grad)

    def capture_graph(self, psi, phi, return_grad=True):
        """Capture one whole step (both sampler launches on their two streams, estimator, gradient, finalize) in a
        CUDA graph; later `step()` calls replay it: one launch instead of ~18, which is what small problems
        (BASELINE C1 / C2: a step is a few hundred microseconds of kernels) are bound by.  The parameters are read
        from the buffers of `psi` / `phi` at replay time, so in-place optimiser updates are seen; the Metropolis
        step counter lives on the device (nkfb_sample_cfg_t.step_offset_dev)."""
        # Single rank only.  Capturing the step of a multi-rank job (with torch's NCCL all-reduces inside the graph) was
        # tried in round 2 and hung on 2 GPUs (gpurun call 176: no progress for 15 min), so it stays refused.
        if self.world != 1:
            raise RuntimeError("capture_graph: single-rank only (capturing the NCCL all-reduces hung on 2 GPUs)")
        if self.K > 1:
            raise RuntimeError("capture_graph: not with streamed segments (StepConfig(stream_segments=1) turns them off)")
        if self._graph is not None:
            return
        if not self.initialised:
            self.step(psi, phi, return_grad)           # random initial configurations, workspaces
        steps_per_call = (self.cfg.n_discard + self.cfg.chain_length) * self.sweep
        self._step_dev = torch.full((1,), self.steps_done * steps_per_call, dtype=torch.int64, device=self.dev)
        self.step(psi, phi, return_grad)               # eager warm-up with the device counter
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.sample(psi, phi)
            out = self.estimate(psi, phi, return_grad)
        self.steps_done -= 1                           # the capture recorded a step, it did not run one
        self._graph, self._graph_out = g, out
