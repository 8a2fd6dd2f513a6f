#!/usr/bin/env python
"""bench.py -- one infidelity-gradient step of p-tVMC (BASELINE.json metric) on N B200s.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference            # the reference's CPU algorithm on the host cores

A "step" = reset + Metropolis sampling of both chain families (discard + chain_length sweeps)
+ local estimator + control variates + score-function gradient (+ the two all-reduces when
N > 1) on the workload of BASELINE.json config C4 (N=100, alpha=4 complex RBM, 2^20 sample
pairs, U = Rx, is_unitary=True, cv=-1/2), total work fixed as N grows ("strong" scaling).
Prints ONE JSON line on rank 0.
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


# ---------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own algorithm on the host cores (NetKet/JAX cannot be installed in this
    image: no wheels, no network -- DESIGN.md), i.e. the oracle port: full forward pass per
    Metropolis step, NumPy complex128, all BLAS threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np  # noqa: F401
    from oracle.baseline import make_workload, cpu_step
    w = make_workload(args.workload)
    n_chains, cl = args.ref_chains, min(args.ref_chain_length, w["chain_length"])
    times = []
    for i in range(args.warmup + args.steps):
        dt, S, _ = cpu_step(w, n_chains, cl, seed=i)
        if i >= args.warmup:
            times.append(dt)
    t = sum(times) / len(times)
    val = S / t
    sample = (f"{n_chains} chains x {cl} samples per family (n_discard 5, sweep {w['N']}) of workload "
              f"{args.workload}; full step (sampling with a full forward per MH step + estimator + gradient)")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(args.workload, w, 1),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample},
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


def bench_config(name, w, n_gpus):
    return {"workload": f"{name}: N={w['N']} spins, RBM alpha={w['M'] // w['N']} complex (M={w['M']}), "
                        f"{w['n_chains'] * w['chain_length']} sample pairs = {w['n_chains']} chains x "
                        f"{w['chain_length']} per family, n_discard 5, sweep_size N, U={w['U']}, mode={w['mode']}, "
                        f"cv_coeff=-0.5",
            "global_sample_pairs": w["n_chains"] * w["chain_length"], "n_chains_per_family": w["n_chains"],
            "chain_length": w["chain_length"], "parallelism": f"chains sharded over {n_gpus} GPU(s)",
            "l2": "256 MiB scratch buffer rewritten between timed steps (L2 flush, inside the timed region)"}


# ---------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from netket_fidelity_b200 import _native as nv
    from netket_fidelity_b200.engine import InfidelityStep, StepConfig
    from netket_fidelity_b200.workloads import make_workload, device_ops, random_rbm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w = make_workload(args.workload)
    N, M = w["N"], w["M"]
    cd = torch.complex64 if args.dtype == "f32" else torch.complex128
    U, Ud = device_ops(w, dev)
    cfg = StepConfig(w["n_chains"], w["chain_length"], 5, N, seed=20261019, cv_coeff=-0.5, site_mode=args.site_mode)
    eng = InfidelityStep(N, M, cfg, w["mode"], U, Ud, device=dev, cdtype=cd)
    h = eng.h
    # host (pinned) and device copies of the parameters
    host = {k: random_rbm(N, M, seed, cd, w["vb"]) for k, seed in (("psi", 1234), ("phi", 5678))}
    pinned = {k: [t.pin_memory() if t is not None else None for t in v] for k, v in host.items()}
    devp = {k: [t.to(dev) if t is not None else None for t in v] for k, v in host.items()}
    specs = {k: nv.RbmSpec(*v) for k, v in devp.items()}
    flush = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)
    grad_host = torch.empty((eng.P,), dtype=cd).pin_memory()
    sums_host = torch.empty((8,), dtype=torch.float64).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()          # started before the warm-up so that nvidia-smi is already reporting when timing starts
    for _ in range(args.warmup):
        flush.zero_()
        eng.step(specs["psi"], specs["phi"])
    barrier()

    # ------------------------------------------------------------ device-resident timing
    samp_ev = [(ev(), ev()) for _ in range(args.steps)]
    est_ev = [(ev(), ev(), ev()) for _ in range(args.steps)]
    launches0 = h.launches
    barrier()
    e0, e1 = ev(), ev()
    t_wall0 = time.time()
    e0.record()
    for i in range(args.steps):
        flush.zero_()
        samp_ev[i][0].record()
        eng.sample(specs["psi"], specs["phi"])
        samp_ev[i][1].record()
        sums, L, grad = eng.estimate(specs["psi"], specs["phi"], events=est_ev[i])
    e1.record()
    barrier()
    launches = h.launches - launches0
    ms = e0.elapsed_time(e1)
    graph_ms = None
    if args.graph and world == 1:
        # the same K steps replayed from ONE captured CUDA graph per step (engine.capture_graph): what the step
        # costs once the ~18 launches and the Python in between are out of the way (small workloads)
        eng.capture_graph(specs["psi"], specs["phi"])
        for _ in range(3):
            eng.step(specs["psi"], specs["phi"])
        torch.cuda.synchronize()
        g0, g1 = ev(), ev()
        g0.record()
        for i in range(args.steps):
            flush.zero_()
            sums, L, grad = eng.step(specs["psi"], specs["phi"])
        g1.record()
        torch.cuda.synchronize()
        graph_ms = g0.elapsed_time(g1) / args.steps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    s = sums.cpu().numpy()          # result of the last TIMED step (the extension below runs further steps)
    infid = 1.0 - s[0] / s[4]
    acc_rate = eng.n_accepted.sum().item() / (2.0 * eng.local_chains * (5 + w["chain_length"]) * N * eng.steps_done)
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
    samp_ms = sum(a.elapsed_time(b) for a, b in samp_ev) / args.steps   # two sampler launches per step
    fwd_ms = sum(a.elapsed_time(b) for a, b, c in est_ev) / args.steps
    grad_ms = sum(b.elapsed_time(c) for a, b, c in est_ev) / args.steps  # includes the 64-byte all-reduce when N > 1

    # ------------------------------------------------------------ end-to-end through the public API (host buffers)
    # psi.parameters <- pinned host copy (H2D), reset both states, psi.expect_and_grad(I_op) [sampling of both
    # families + estimator + gradient + all-reduces], gradient and statistics back to pinned host memory (D2H).
    import netket_fidelity_b200 as nkf
    from netket_fidelity_b200 import nk
    hi = nk.hilbert.Spin(0.5, N)
    sa = nk.sampler.MetropolisLocal(hi, n_chains=w["n_chains"], sweep_size=N)
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
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        st_e2e = e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    h2d = sum(p.numel() * p.element_size() for k in pinned for p in pinned[k] if p is not None)
    d2h = grad_host.numel() * grad_host.element_size() + 64

    S_total = eng.S_total
    value = S_total * args.steps / (ms_total * 1e-3)
    line = None
    if rank == 0:
        hbm_gbs, _, how = read_peaks()
        # roofline of the dominant kernel (the multiplicative sampler, sampler_ratio.cuh): FP32 pipe.
        # Per hidden unit and proposal: complex multiply z q (4), 1 + z' (1), |.|^2 (2), running product (1) =
        # 8 FP32-pipe operations, issued as packed FFMA2 / FMUL2 / FADD2 (two per instruction).
        ffma2 = max(h.microbench(5, 4096), 1.0)         # FMA/s through FFMA2 measured on this GPU, this run
        ffma = max(h.microbench(3, 4096), 1.0)
        mufu = max(h.microbench(0, 2048), 1.0)
        sweeps = 5 + w["chain_length"]
        evals_per_launch = eng.local_chains * sweeps * N * M          # hidden-unit updates, one family
        t_launch = samp_ms * 1e-3 / 2
        achieved = evals_per_launch / t_launch
        ops_per_eval = 8.0
        peak = ffma2 / ops_per_eval
        bytes_per_launch = eng.local_chains * w["chain_length"] * eng.W32 * 4 + N * M * 8
        _, bf16_peak, _ = read_peaks()
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
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": bench_config(args.workload, w, world),
            "clocks": clk,
            "e2e": {"value": S_total * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp32", "kernel": "sample_ratio_cta_kernel", "achieved": achieved / 1e9,
                         "peak": peak / 1e9, "unit": "G hidden-unit updates/s", "frac": achieved / peak,
                         "traffic": 1209344,
                         "note": (f"the dominant kernel is bound by the FP32 pipe, neither HBM nor tensor: algorithmic work = "
                                  f"{evals_per_launch:.3e} hidden-unit updates per launch (chains x sweeps x N x M) at "
                                  f"{ops_per_eval:.0f} FP32-pipe operations each (complex multiply, |1+z|^2, product; no "
                                  f"transcendental per unit); peak = FFMA2 rate measured by nkfb_microbench in this run "
                                  f"({ffma2 / 1e12:.2f} T FMA/s; scalar FFMA {ffma / 1e12:.2f} T/s; MUFU.EX2 "
                                  f"{mufu / 1e12:.3f} T/s) / {ops_per_eval:.0f}; avg launch {t_launch * 1e3:.3f} ms (two "
                                  f"launches share the GPU on two streams: half of the time both take); sampler share "
                                  f"of step {samp_ms / (ms_total / args.steps):.3f}"),
                         "hbm_view": {"bound": "hbm", "achieved": bytes_per_launch / t_launch / 1e9, "peak": hbm_gbs,
                                      "unit": "GB/s", "frac": bytes_per_launch / t_launch / 1e9 / hbm_gbs,
                                      "peak_source": how},
                         "traffic_note": "dram__bytes_read+write of one sampler launch from profiles/r01_ncu_top_kernels_v3.md "
                                         "(1.21 MB read, 0 B written: tables and packed samples stay in L2)",
                         "other_kernels": other},
            "result": {"infidelity": infid, "acceptance": acc_rate},
            **({"cuda_graph": {"ms_per_step": graph_ms, "value": S_total / (graph_ms * 1e-3), "unit": UNIT,
                               "note": "one captured CUDA graph per step, replayed (not the headline `value`)"}}
               if graph_ms else {}),
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle.baseline import make_workload as omw, cpu_step
            ow = omw(args.workload)
            nc, cl = args.ref_chains, min(args.ref_chain_length, ow["chain_length"])
            dt, Sc, _ = cpu_step(ow, nc, cl)
            line["cpu_baseline"] = {"value": Sc / dt, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"{nc} chains x {cl} samples per family of workload {args.workload}, "
                                              f"one full step, NumPy complex128 oracle (reference algorithm: full "
                                              f"forward per MH step), {dt:.1f} s"}
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
    ap.add_argument("--site-mode", type=int, default=0)
    ap.add_argument("--ref-chains", type=int, default=16)
    ap.add_argument("--ref-chain-length", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", action="store_true", help="also time the step replayed from a captured CUDA graph")
    args = ap.parse_args()
    _quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
