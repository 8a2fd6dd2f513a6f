#!/usr/bin/env python
"""bench.py -- one infidelity-gradient step of p-tVMC (BASELINE.json metric) on N B200s.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference            # the reference's CPU algorithm on the host cores

A "step" = reset + Metropolis sampling of both chain families (discard + chain_length sweeps)
+ local estimator + control variates + score-function gradient (+ the all-reduces when
N > 1) on the workload of BASELINE.json config C4 (N=100, alpha=4 complex RBM, 2^20 sample
pairs, U = Rx, is_unitary=True, cv=-1/2), total work fixed as N grows ("strong" scaling).
Prints ONE JSON line on rank 0.  Besides the headline (complex64, the throughput precision the north star
sanctions with a 1e-5 parity tolerance) the line carries
  * "f64": the SAME C4 step in complex128 (the reference's default arithmetic, parity tolerance 1e-10) with its
    own roofline against the DFMA rate measured in this run;
  * "other_workloads": the other BASELINE.json configurations (C1, C2, C3 and the Renyi-2 estimator C5), each
    with the roofline of its dominant kernel; they are sharded over the ranks like the headline;
  * "cpu_baseline": the reference's algorithm (full forward per Metropolis step, complex128) on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "infidelity_gradient_sample_pairs_per_second"
UNIT = "sample_pairs/s"


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def n_in(self, window):
        return sum(1 for t, _ in self.lines if window[0] <= t <= window[1])

    def stop(self, window=None):
        """Summary of the samples that arrived inside `window` (host wall-clock of the timed region, possibly
        extended by the caller with more steps of the same workload when the region is shorter than nvidia-smi's
        100 ms period)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for tstamp, ln in self.lines:
            if window is not None and not (window[0] <= tstamp <= window[1]):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for nme, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1400.0, "fallback"


def read_ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from the committed ncu
    summary of this round (profiles/ncu_sampler_traffic.json, written by tools/ncu_summary.py from an
    `ncu --set full` capture of this same command); null when the file is absent."""
    p = os.path.join(ROOT, "profiles", "ncu_sampler_traffic.json")
    try:
        d = json.load(open(p))
        return int(d["dram_bytes_read"] + d["dram_bytes_write"]), d.get("source", p)
    except Exception:
        return None, "no ncu capture committed for this kernel"


SITE_MODES = {0: "shared (one site sequence per call, every chain an exact random-scan Metropolis chain)",
              1: "per_chain (NetKet's MetropolisLocal: each chain draws its own site)"}


# ---------------------------------------------------------------------------------------------
def cpu_baseline(workload, n_chains, chain_length):
    """The reference's algorithm on all host cores: oracle/baseline_numba.py (numba prange over chains, full forward
    per Metropolis step, complex128)."""
    from oracle.baseline import make_workload
    from oracle import baseline_numba as bn
    w = make_workload(workload)
    bn.warm_up()                                            # JIT outside the timed region
    cl = min(chain_length, w["chain_length"])
    dt, S, _, info = bn.cpu_step(w, n_chains, cl)
    sample = (f"{n_chains} chains x {cl} samples per family (n_discard 5, sweep {w['N']}) of workload {workload}: one "
              f"full step = sampling with a full forward pass per Metropolis step ({info['engine']}, "
              f"{info['sampling_s']:.1f} s) + estimator + gradient (NumPy/BLAS, {info['estimator_s']:.1f} s), "
              f"complex128 -- the reference's arithmetic; the GPU headline is complex64, its complex128 twin is "
              f"under 'f64'")
    return {"value": S / dt, "unit": UNIT, "cores": info["threads"] or os.cpu_count(), "kind": "port",
            "dtype": "f64 (complex128)", "seconds": dt, "sample": sample}, w


def run_reference(args):
    """The reference's own algorithm on the host cores (NetKet/JAX cannot be installed in this
    image: no wheels, no network -- DESIGN.md), i.e. the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        cb, w = cpu_baseline(args.workload, args.ref_chains, args.ref_chain_length)
        if i >= args.warmup:
            vals.append(cb)
        last = cb
    t = sum(c["seconds"] for c in vals) / len(vals)
    val = sum(c["value"] for c in vals) / len(vals)
    last = dict(last, value=val)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(args.workload, w, 1, args.site_mode),
            "cpu_baseline": last,
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


def host_ops(w, hi):
    import netket_fidelity_b200 as nkf
    from netket_fidelity_b200 import nk
    u = w["U"]
    if u[0] == "rx":
        return nkf.operator.Rx(hi, u[1], u[2]), nkf.operator.Rx(hi, u[1], -u[2])
    if u[0] == "hadamard":
        return nkf.operator.Hadamard(hi, u[1]), nkf.operator.Hadamard(hi, u[1])
    if u[0] == "ising_prop":
        return nkf.operator.IsingPropagator(hi, nk.graph.Chain(w["N"]), u[2], u[3], u[1]), None
    raise ValueError(u)


def bench_config(name, w, n_gpus, site_mode=0):
    return {"workload": f"{name}: N={w['N']} spins, RBM alpha={w['M'] // w['N']} complex (M={w['M']}), "
                        f"{w['n_chains'] * w['chain_length']} sample pairs = {w['n_chains']} chains x "
                        f"{w['chain_length']} per family, n_discard 5, sweep_size N, U={w.get('U')}, mode={w['mode']}, "
                        f"cv_coeff=-0.5, site_mode={SITE_MODES[site_mode]}",
            "global_sample_pairs": w["n_chains"] * w["chain_length"], "n_chains_per_family": w["n_chains"],
            "chain_length": w["chain_length"], "site_mode": site_mode,
            "parallelism": f"chains sharded over {n_gpus} GPU(s)",
            "l2": "160 MiB scratch buffer (larger than the 126 MB L2) rewritten between the timed steps (L2 flush); the flushes "
                  "are timed by their own CUDA events and subtracted from the bracket of the K steps (ms_per_step_bracket "
                  "and l2_flush_ms in the line give the raw bracket and the flush)"}


# ---------------------------------------------------------------------------------------------
class Ctx:
    """per-process plumbing shared by the timed sections"""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.flush = torch.empty((160 << 20,), dtype=torch.uint8, device=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def ev(self):
        return self.torch.cuda.Event(enable_timing=True)

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def net_of_flushes(ctx, e0, e1, flush_pairs):
    """(ms_total, ms_bracket, flush_ms): the bracket e0 -> e1 of K steps minus the K L2 flushes between them (K x the
    median of their event-timed intervals: a flush is one fill kernel, the median keeps a host hiccup between the two
    records out of the subtraction), max over the ranks of both."""
    bracket = e0.elapsed_time(e1)
    iv = sorted(a.elapsed_time(b) for a, b in flush_pairs)
    flush = iv[len(iv) // 2] if iv else 0.0
    return ctx.max_over_ranks(bracket - flush * len(iv)), ctx.max_over_ranks(bracket), flush


def make_engine(ctx, name, dtype, site_mode, one_collective=False):
    from netket_fidelity_b200 import _native as nv
    from netket_fidelity_b200.engine import InfidelityStep, StepConfig
    from netket_fidelity_b200.workloads import make_workload, device_ops, random_rbm
    torch = ctx.torch
    w = make_workload(name)
    N, M = w["N"], w["M"]
    cd = torch.complex64 if dtype == "f32" else torch.complex128
    U, Ud = device_ops(w, ctx.dev)
    cfg = StepConfig(w["n_chains"], w["chain_length"], 5, N, seed=20261019, cv_coeff=-0.5, site_mode=site_mode,
                     one_collective=one_collective)
    eng = InfidelityStep(N, M, cfg, w["mode"], U, Ud, device=ctx.dev, cdtype=cd)
    host = {k: random_rbm(N, M, seed, cd, w["vb"]) for k, seed in (("psi", 1234), ("phi", 5678))}
    devp = {k: [t.to(ctx.dev) if t is not None else None for t in v] for k, v in host.items()}
    specs = {k: nv.RbmSpec(*v) for k, v in devp.items()}
    return w, eng, host, devp, specs, cd


def time_engine(ctx, eng, specs, steps, warmup):
    """W untimed + K timed steps of the device-resident engine, bracketed by a barrier + synchronize and one pair of CUDA
    events (on the launching stream) on both sides.  The L2 is flushed BETWEEN the timed steps (a 160 MiB buffer rewritten
    on the same stream); each flush is timed by its own pair of events and the K flush intervals are subtracted from the
    bracket: `ms_total` = max over the ranks of (bracket - flushes) -- the K steps and every gap between them, not the
    flushes, which are the measurement's and not the workload's; `ms_bracket` (max over the ranks of the raw bracket) and
    `flush_ms` are reported beside it."""
    for _ in range(warmup):
        ctx.flush.zero_()
        eng.step(specs["psi"], specs["phi"])
    ctx.barrier()
    samp_ev = [(ctx.ev(), ctx.ev()) for _ in range(steps)]
    est_ev = [(ctx.ev(), ctx.ev(), ctx.ev()) for _ in range(steps)]
    flush_ev = [ctx.ev() for _ in range(steps)]
    launches0 = eng.h.launches
    ctx.barrier()
    e0, e1 = ctx.ev(), ctx.ev()
    t_wall0 = time.time()
    e0.record()
    for i in range(steps):
        flush_ev[i].record()
        ctx.flush.zero_()
        samp_ev[i][0].record()                      # start of step i: the flush before it has completed on this stream
        eng.sample(specs["psi"], specs["phi"])
        samp_ev[i][1].record()
        sums, L, grad = eng.estimate(specs["psi"], specs["phi"], events=est_ev[i])
    e1.record()
    ctx.barrier()
    ms_total, ms_bracket, flush_ms = net_of_flushes(ctx, e0, e1, [(f, s[0]) for f, s in zip(flush_ev, samp_ev)])
    return dict(ms_total=ms_total, ms_per_step=ms_total / steps, ms_bracket=ms_bracket, flush_ms=flush_ms,
                launches=eng.h.launches - launches0,
                samp_ms=sum(a.elapsed_time(b) for a, b in samp_ev) / steps,
                fwd_ms=sum(a.elapsed_time(b) for a, b, c in est_ev) / steps,
                grad_ms=sum(b.elapsed_time(c) for a, b, c in est_ev) / steps,   # incl. the scalar all-reduce when N > 1
                sums=sums, t_wall0=t_wall0)


def sampler_work(w, local_chains):
    """Algorithmic FP32-pipe operations of the sampler launches of one step on this rank:
    plain family: chains x sweeps x N proposals x M hidden-unit updates x 8 operations (complex multiply 4, 1 + z 1,
    |.|^2 2, running product 1); an Ising-type composite target evaluates N + 1 connected products of M factors per
    proposal at 10 operations (two complex multiplies and an add, packed) per (connected configuration, unit)."""
    N, M = w["N"], w["M"]
    sweeps = 5 + w["chain_length"]
    evals = local_chains * sweeps * N * M
    fams = 1 if w["mode"] == "renyi2" else 2
    ops = evals * 8.0 * (fams if w["mode"] != "standard_Uphi" else 1)
    if w["mode"] == "standard_Uphi":
        ops += evals * (N + 1) * 10.0
    return evals * fams, ops


def time_renyi(ctx, name, dtype, steps, warmup):
    """C5: Renyi-2 swap estimator (reference renyi2.py:1): sample 2^20 configurations of psi, four amplitude
    evaluations per pair of samples, one reduction; samples sharded over the ranks."""
    from netket_fidelity_b200 import _native as nv
    from netket_fidelity_b200.workloads import make_workload, random_rbm
    import numpy as np
    torch = ctx.torch
    w = make_workload(name)
    N, M = w["N"], w["M"]
    cd = torch.complex64 if dtype == "f32" else torch.complex128
    h = nv.get_handle(ctx.dev)
    psi = nv.RbmSpec(*[t.to(ctx.dev) if t is not None else None for t in random_rbm(N, M, 1234, cd, w["vb"])])
    local = w["n_chains"] // ctx.world
    W32 = (N + 31) // 32
    state = torch.zeros((local, W32), dtype=torch.int32, device=ctx.dev)
    out = torch.empty((local, w["chain_length"], W32), dtype=torch.int32, device=ctx.dev)
    words = np.zeros(W32, dtype=np.uint32)
    for i in w["partition"]:
        words[i >> 5] |= np.uint32(1 << (i & 31))
    mask = torch.tensor(words.view(np.int32), device=ctx.dev)
    ws = h.sample_workspace(psi)
    per_call = (5 + w["chain_length"]) * N
    k = [0]

    def step(events=None):
        if events:
            events[0].record()
        h.sample(psi, state, w["chain_length"], 5, N, 20261019, (-1.0, 1.0), chain_offset=ctx.rank * local,
                 step_offset=k[0] * per_call, init_random=(k[0] == 0), out=out, workspace=ws)
        if events:
            events[1].record()
        q, sums = h.renyi2(psi, out.view(-1, W32), mask, (-1.0, 1.0))
        if ctx.world > 1:
            ctx.dist.all_reduce(sums)
        k[0] += 1
        return sums

    for _ in range(warmup):
        ctx.flush.zero_()
        step()
    ctx.barrier()
    evs = [(ctx.ev(), ctx.ev()) for _ in range(steps)]
    flush_ev = [ctx.ev() for _ in range(steps)]
    launches0 = h.launches
    e0, e1 = ctx.ev(), ctx.ev()
    e0.record()
    for i in range(steps):
        flush_ev[i].record()
        ctx.flush.zero_()
        sums = step(evs[i])
    e1.record()
    ctx.barrier()
    ms_total, _, _ = net_of_flushes(ctx, e0, e1, [(f, e[0]) for f, e in zip(flush_ev, evs)])
    s = sums.cpu().numpy()
    S2 = float(-np.log2(s[0] / s[3])) if s[0] > 0 else float("nan")
    return w, dict(ms_total=ms_total, ms_per_step=ms_total / steps, launches=h.launches - launches0,
                   samp_ms=sum(a.elapsed_time(b) for a, b in evs) / steps, local_chains=local, S2=S2)


def other_workload_entry(ctx, name, dtype, site_mode, steps, warmup, ffma2):
    """One line of the `other_workloads` block: step time, rate, launches, and the roofline of the dominant kernel
    (the sampler in every configuration) against the FFMA2 rate measured in this run."""
    if name == "C5":
        w, r = time_renyi(ctx, name, dtype, steps, warmup)
        n_units = w["n_chains"] * w["chain_length"]
        unit, local = "samples/s", r["local_chains"]
        extra = {"renyi2_entropy": r["S2"]}
    else:
        w, eng, host, devp, specs, cd = make_engine(ctx, name, dtype, site_mode)
        r = time_engine(ctx, eng, specs, steps, warmup)
        n_units, unit, local = eng.S_total, UNIT, eng.local_chains
        s = r["sums"].cpu().numpy()
        extra = {"infidelity": float(1.0 - s[0] / s[4])}
    evals, ops = sampler_work(w, local)
    t_samp = max(r["samp_ms"], 1e-6) * 1e-3
    return {"workload": bench_config(name, w, ctx.world, site_mode)["workload"] if name != "C5" else
            f"C5: Renyi-2 swap estimator, N={w['N']} (8x8), RBM alpha=2 complex, {n_units} samples = {w['n_chains']} "
            f"chains x {w['chain_length']}, partition = first {len(w['partition'])} sites, no gradient",
            "dtype": dtype, "ms_per_step": r["ms_per_step"], "value": n_units / (r["ms_per_step"] * 1e-3), "unit": unit,
            "gpu_launches_per_step": r["launches"] / steps, "sampler_ms": r["samp_ms"],
            "sampler_share_of_step": r["samp_ms"] / r["ms_per_step"],
            "roofline": {"bound": "fp32", "kernel": "Metropolis sampler launches of the step",
                         "achieved": ops / t_samp / 1e12, "peak": ffma2 / 1e12, "unit": "T FP32-pipe op/s",
                         "frac": ops / t_samp / ffma2,
                         "note": "algorithmic operations (bench.py:sampler_work) / sampler time of the step; small "
                                 "configurations are bound by launch latency, not by this pipe"},
            **extra}


# ---------------------------------------------------------------------------------------------
def run_b200(args):
    ctx = Ctx()
    torch, dist, world, rank, dev = ctx.torch, ctx.dist, ctx.world, ctx.rank, ctx.dev
    w, eng, host, devp, specs, cd = make_engine(ctx, args.workload, args.dtype, args.site_mode)
    N, M = w["N"], w["M"]
    h = eng.h
    pinned = {k: [t.pin_memory() if t is not None else None for t in v] for k, v in host.items()}
    grad_host = torch.empty((eng.P,), dtype=cd).pin_memory()

    clocks = ClockSampler(ctx.local_rank)
    if rank == 0:
        clocks.start()          # started before the warm-up so that nvidia-smi is already reporting when timing starts

    # ------------------------------------------------------------ device-resident timing (the headline `value`)
    r = time_engine(ctx, eng, specs, args.steps, args.warmup)
    ms_total, launches, t_wall0 = r["ms_total"], r["launches"], r["t_wall0"]
    samp_ms, fwd_ms, grad_ms = r["samp_ms"], r["fwd_ms"], r["grad_ms"]
    s = r["sums"].cpu().numpy()          # result of the last TIMED step (the extension below runs further steps)
    infid = 1.0 - s[0] / s[4]
    acc_rate = eng.n_accepted.sum().item() / (2.0 * eng.local_chains * (5 + w["chain_length"]) * N * eng.steps_done)
    graph_ms = None
    if args.graph and world == 1:
        # the same K steps replayed from ONE captured CUDA graph per step (engine.capture_graph)
        eng.capture_graph(specs["psi"], specs["phi"])
        for _ in range(3):
            eng.step(specs["psi"], specs["phi"])
        ctx.barrier()
        g0, g1 = ctx.ev(), ctx.ev()
        gf = [(ctx.ev(), ctx.ev()) for _ in range(args.steps)]
        g0.record()
        for i in range(args.steps):
            gf[i][0].record()
            ctx.flush.zero_()
            gf[i][1].record()
            eng.step(specs["psi"], specs["phi"])
        g1.record()
        ctx.barrier()
        graph_ms = net_of_flushes(ctx, g0, g1, gf)[0] / args.steps
    # clocks under load: samples that arrived during the timed region; when it is shorter than nvidia-smi's period
    # (multi-GPU runs: tens of ms) every rank keeps running the same step, untimed, until rank 0 has two samples
    t_wall1 = time.time()
    for _ in range(200):
        need = torch.tensor([1 if (rank == 0 and clocks.n_in((t_wall0, t_wall1)) < 2) else 0], device=dev)
        if world > 1:
            dist.broadcast(need, 0)
        if int(need.item()) == 0:
            break
        eng.step(specs["psi"], specs["phi"])
        torch.cuda.synchronize()
        t_wall1 = time.time()
    clk = clocks.stop((t_wall0, t_wall1)) if rank == 0 else None

    # ------------------------------------------------------------ end-to-end through the public API (host buffers)
    # psi.parameters <- pinned host copy (H2D), reset both states, psi.expect_and_grad(I_op) [sampling of both
    # families + estimator + gradient + all-reduces], gradient and statistics back to pinned host memory (D2H).
    import netket_fidelity_b200 as nkf
    from netket_fidelity_b200 import nk
    hi = nk.hilbert.Spin(0.5, N)
    sa = nk.sampler.MetropolisLocal(hi, n_chains=w["n_chains"], sweep_size=N,
                                    site_mode="shared" if args.site_mode == 0 else "per_chain")
    ma = nk.models.RBM(alpha=M // N, param_dtype=(complex if args.dtype == "f64" else "complex64"),
                       use_visible_bias=w["vb"])
    S_total = w["n_chains"] * w["chain_length"]
    mk_tree = lambda t: {"Dense": {"kernel": t[0], "bias": t[1]}, **({"visible_bias": t[2]} if t[2] is not None else {})}
    vs_psi = nk.vqs.MCState(sa, ma, n_samples=S_total, n_discard_per_chain=5, seed=1, device=dev,
                            variables={"params": mk_tree(devp["psi"])})
    vs_phi = nk.vqs.MCState(sa, ma, n_samples=S_total, n_discard_per_chain=5, seed=2, device=dev,
                            variables={"params": mk_tree(devp["phi"])})
    Uop, Udop = host_ops(w, hi)
    I_op = nkf.InfidelityOperator(vs_phi, U=Uop, U_dagger=Udop, cv_coeff=-0.5,
                                  is_unitary=(w["mode"] == "upsi"), sample_Upsi=(w["mode"] == "standard_Uphi"))
    pin_tree = {k: mk_tree(pinned[k]) for k in pinned}

    def e2e_step():
        vs_psi.parameters = pin_tree["psi"]            # H2D (pinned -> flat device buffer)
        I_op.target.parameters = pin_tree["phi"]
        vs_psi.reset()
        I_op.target.reset()
        st, grad = vs_psi.expect_and_grad(I_op)
        grad_host.copy_(vs_psi._last_flat_grad, non_blocking=True)   # D2H
        torch.cuda.synchronize()
        return st

    for _ in range(2):
        e2e_step()
    ctx.barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step()
    ctx.barrier()
    e2e_s = ctx.max_over_ranks(time.perf_counter() - t0)
    h2d = sum(p.numel() * p.element_size() for k in pinned for p in pinned[k] if p is not None)
    # gradient + the 8 sums + the chain-statistics partial sums MCState reads back (rows x (3 + n_blocks) doubles with
    # rows = chain_length on every rank, as the reference hands it to nkjax.expect; all ranks' rows after the all-gather)
    st_rows, st_cols = w["chain_length"], w["n_chains"] // world
    st_blocks = st_cols // max(1, st_cols // 32)
    d2h = grad_host.numel() * grad_host.element_size() + 64 + st_rows * world * (3 + st_blocks) * 8
    del vs_psi, vs_phi, I_op

    # ------------------------------------------------------------ the same C4 step in complex128 ("f64") and the other
    # BASELINE.json configurations; every rank takes part (sharded like the headline), rank 0 reports
    ffma2 = max(h.microbench(5, 4096), 1.0)         # FMA/s through FFMA2 measured on this GPU, this run
    ffma = max(h.microbench(3, 4096), 1.0)
    mufu = max(h.microbench(0, 2048), 1.0)
    dfma = max(h.microbench(6, 2048), 1.0)
    f64_block, others, one_coll, per_chain = None, {}, None, None
    if not args.no_others and args.site_mode == 0:
        # the same step with NetKet's per-chain proposal sites (MetropolisLocal: every chain draws its own site; the
        # one-chain-per-warp kernel samples it), beside the shared-site-sequence headline
        _, engp, _, _, specsp, _ = make_engine(ctx, args.workload, args.dtype, 1)
        rp = time_engine(ctx, engp, specsp, max(3, min(5, args.steps)), 3)
        sp = rp["sums"].cpu().numpy()
        per_chain = {"site_mode": SITE_MODES[1], "ms_per_step": rp["ms_per_step"],
                     "value": engp.S_total / (rp["ms_per_step"] * 1e-3), "unit": UNIT, "sampler_ms": rp["samp_ms"],
                     "infidelity": float(1.0 - sp[0] / sp[4])}
        del engp, specsp
    if not args.no_others and world > 1 and w["mode"] == "upsi":
        # the same step with ONE all-reduce (engine option one_collective): pass B centred with the previous
        # step's F plus sum_s conj O_k(sigma_s); reported beside the two-collective headline, not instead of it
        _, eng1, _, _, specs1, _ = make_engine(ctx, args.workload, args.dtype, args.site_mode, one_collective=True)
        r1 = time_engine(ctx, eng1, specs1, max(3, min(5, args.steps)), 5)
        one_coll = {"ms_per_step": r1["ms_per_step"], "value": eng1.S_total / (r1["ms_per_step"] * 1e-3), "unit": UNIT,
                    "collectives_per_step": 1,
                    "note": "one all-reduce over [sums | acc1 | acc2] (8 + 4P doubles) instead of 8 doubles + 2P doubles; "
                            "costs the sigma half of pass B once more"}
        del eng1, specs1
    if not args.no_others:
        if args.dtype == "f32":
            w64, eng64, _, _, specs64, _ = make_engine(ctx, args.workload, "f64", args.site_mode)
            r64 = time_engine(ctx, eng64, specs64, max(2, min(3, args.steps)), 1)
            s64 = r64["sums"].cpu().numpy()
            evals64 = eng64.local_chains * (5 + w64["chain_length"]) * N * M
            t64 = r64["samp_ms"] * 1e-3 / 2
            f64_block = {"dtype": "f64 (complex128: the reference's default arithmetic, README.md:56; parity 1e-10)",
                         "ms_per_step": r64["ms_per_step"], "value": eng64.S_total / (r64["ms_per_step"] * 1e-3),
                         "unit": UNIT, "steps": max(2, min(3, args.steps)), "warmup": 1,
                         "sampler_ms": r64["samp_ms"], "estimator_ms": r64["fwd_ms"], "gradient_ms": r64["grad_ms"],
                         "infidelity": float(1.0 - s64[0] / s64[4]),
                         "roofline": {"bound": "fp64", "kernel": "sample_ratio_pipe_kernel<double>",
                                      "achieved": evals64 / t64 / 1e9, "peak": dfma / 8.0 / 1e9,
                                      "unit": "G hidden-unit updates/s", "frac": evals64 / t64 / (dfma / 8.0),
                                      "note": f"8 FP64-pipe operations per hidden-unit update; peak = DFMA rate measured "
                                              f"by nkfb_microbench in this run ({dfma / 1e12:.2f} T FMA/s) / 8"}}
            del eng64, specs64
        for name in ("C1", "C2", "C3", "C5"):
            from netket_fidelity_b200.workloads import make_workload
            if make_workload(name)["n_chains"] % world:
                continue
            others[name] = other_workload_entry(ctx, name, "f32", args.site_mode if name != "C3" else 0, 3, 2, ffma2)

    value = eng.S_total * args.steps / (ms_total * 1e-3)
    if rank == 0:
        hbm_gbs, bf16_peak, how = read_peaks()
        # roofline of the dominant kernel (the multiplicative sampler): FP32 pipe.  Per hidden unit and proposal:
        # complex multiply z q (4), 1 + z' (1), |.|^2 (2), running product (1) = 8 FP32-pipe operations, issued as
        # packed FFMA2 / FMUL2 / FADD2 (two per instruction).
        sweeps = 5 + w["chain_length"]
        evals_per_launch = eng.local_chains * sweeps * N * M          # hidden-unit updates, one family
        # streamed shards (engine.py: pass A of the first sample segment runs under the second one): the sampling interval
        # holds pass A of all but the last segment and `fwd_ms` only the record-to-record gap; the roofline of the sampler
        # then counts that interval as sampler time (a lower bound of its fraction)
        streamed = getattr(eng, "K", 1)
        t_launch = samp_ms * 1e-3 / 2
        achieved = evals_per_launch / t_launch
        ops_per_eval = 8.0
        peak = ffma2 / ops_per_eval
        bytes_per_launch = eng.local_chains * w["chain_length"] * eng.W32 * 4 + N * M * 8
        traffic, traffic_src = read_ncu_traffic()
        S_loc = eng.S_local
        # algorithmic GEMM FLOPs (one bf16-exact operand; the kernels execute 3x that as bf16 split terms)
        fwd_flops = 4 * 2.0 * (N + 1) * 2 * M * S_loc
        grad_flops = 2 * 2.0 * 2.0 * (N + 1) * 2 * (M + 1) * S_loc
        other = [
            {"kernel": "infid_fwd_tc_kernel (tcgen05)", "ms": fwd_ms, "bound": "tensor+sfu epilogue",
             "achieved": fwd_flops / (fwd_ms * 1e-3) / 1e12, "peak": bf16_peak, "unit": "TFLOP/s",
             "frac": fwd_flops / (fwd_ms * 1e-3) / 1e12 / bf16_peak,
             "note": "algorithmic 4 forwards x 2(N+1)2M FLOP per pair; executed as 3 bf16 split UMMAs; the log-cosh "
                     "epilogue (6 M complex log-cosh per pair), not the MMA, bounds this kernel"},
            {"kernel": "infid_grad_tc_kernel (tcgen05)", "ms": grad_ms, "bound": "tensor+sfu epilogue",
             "achieved": grad_flops / (grad_ms * 1e-3) / 1e12, "peak": bf16_peak, "unit": "TFLOP/s",
             "frac": grad_flops / (grad_ms * 1e-3) / 1e12 / bf16_peak,
             "note": "algorithmic (forward + X^T Y) for sigma and eta tiles; executed as 3 bf16 split UMMAs each"},
        ]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "ms_per_step_bracket": r["ms_bracket"] / args.steps,
            "l2_flush_ms": r["flush_ms"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": bench_config(args.workload, w, world, args.site_mode),
            "clocks": clk,
            "e2e": {"value": eng.S_total * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp32", "kernel": "sample_ratio_lean_hw_kernel", "achieved": achieved / 1e9,
                         "peak": peak / 1e9, "unit": "G hidden-unit updates/s", "frac": achieved / peak,
                         "traffic": traffic,
                         "note": (f"the dominant kernel is bound by the FP32 pipe, neither HBM nor tensor: algorithmic work = "
                                  f"{evals_per_launch:.3e} hidden-unit updates per launch (chains x sweeps x N x M) at "
                                  f"{ops_per_eval:.0f} FP32-pipe operations each (complex multiply, |1+z|^2, product; no "
                                  f"transcendental per unit); peak = FFMA2 rate measured by nkfb_microbench in this run "
                                  f"({ffma2 / 1e12:.2f} T FMA/s; scalar FFMA {ffma / 1e12:.2f} T/s; MUFU.EX2 "
                                  f"{mufu / 1e12:.3f} T/s) / {ops_per_eval:.0f}; avg launch {t_launch * 1e3:.3f} ms (two "
                                  f"launches share the GPU on two streams: half of the time both take); sampler share "
                                  f"of step {samp_ms / (ms_total / args.steps):.3f}"
                                  + (f"; pass A streamed under the sampler in {streamed} segments (its time is inside the "
                                     f"sampling interval, other_kernels[0].ms is not meaningful)" if streamed > 1 else "")),
                         "hbm_view": {"bound": "hbm", "achieved": bytes_per_launch / t_launch / 1e9, "peak": hbm_gbs,
                                      "unit": "GB/s", "frac": bytes_per_launch / t_launch / 1e9 / hbm_gbs,
                                      "peak_source": how},
                         "traffic_note": f"dram__bytes_read + dram__bytes_write of one sampler launch: {traffic_src}",
                         "other_kernels": other},
            "result": {"infidelity": infid, "acceptance": acc_rate},
            **({"cuda_graph": {"ms_per_step": graph_ms, "value": eng.S_total / (graph_ms * 1e-3), "unit": UNIT,
                               "note": "one captured CUDA graph per step, replayed (not the headline `value`)"}}
               if graph_ms else {}),
            **({"f64": f64_block} if f64_block else {}),
            **({"one_collective": one_coll} if one_coll else {}),
            **({"per_chain_sites": per_chain} if per_chain else {}),
            **({"other_workloads": others} if others else {}),
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"], _ = cpu_baseline(args.workload, args.ref_chains, args.ref_chain_length)
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_FD = None


def _quiet_stdout():
    """Library chatter on fd 1 (e.g. NCCL's version banner at init) goes to stderr: stdout carries the one JSON line."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    os.write(_JSON_FD if _JSON_FD is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C4", choices=["C1", "C2", "C3", "C4"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--site-mode", type=int, default=0, choices=[0, 1])
    ap.add_argument("--ref-chains", type=int, default=64)
    ap.add_argument("--ref-chain-length", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the f64 twin and the other BASELINE workloads")
    ap.add_argument("--graph", action="store_true", help="also time the step replayed from a captured CUDA graph")
    args = ap.parse_args()
    _quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
