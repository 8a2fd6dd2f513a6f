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
                 cv_coeff=-0.5, site_mode=0, one_collective=False):
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
        self.sigma = torch.empty((self.local_chains, cfg.chain_length, W32), dtype=torch.int32, device=dev)
        self.eta = torch.empty_like(self.sigma)
        self.S_local = self.local_chains * cfg.chain_length
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
        self.side_stream = torch.cuda.Stream(device=dev)
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
    def sample(self, psi: nv.RbmSpec, phi: nv.RbmSpec):
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
        cfg = self.cfg
        sig = self.sigma.view(-1, self.W32)
        eta = self.eta.view(-1, self.W32)
        if cfg.one_collective and return_grad:
            return self._estimate_one_collective(psi, phi, sig, eta, events)
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
