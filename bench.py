#!/usr/bin/env python
"""Benchmark of the GP log-likelihood hot path (BASELINE.json metric: GP LogL evaluations per
second, fp64, walkers x N up to 1024 x 16k, at 1/2/4/8 B200).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU implementation
    python bench.py --scaling strong --workload C4 ...       # fixed ensemble through sharded run_MCMC (profiles/)

Headline workload (config.workload): BASELINE configs[4] "scaling sweep" at its largest size, N = 16384
points, QuasiPer GP, no mean model: 128 walkers per GPU, i.e. exactly 1024 walkers x 16384 points
on 8 GPUs (weak scaling: per-GPU work fixed).  A "step" is one batched evaluation of every walker
(MCMC.compute of one iteration) followed, for N > 1 ranks, by the all-gather of log L / status.

One JSON line on rank 0.  `value` is device-timed (CUDA events on the launching stream, max over
ranks) with walker parameters already resident in HBM; `e2e` is the same metric through the
host-buffer C ABI call (`mrv_logl_batch`: host parameters in, log L out, copies inside the
timed region); `roofline` is the dominant kernel (trailing SYRK/GEMM update on the FP64 tensor pipe)
against the DMMA issue peak measured in the same run; `cpu_baseline` is the UNMODIFIED reference
(oracle/_ref, staged by oracle/stage_ref.py; the oracle port when that is absent) on the host cores for a
bounded sample.  With one GPU the same line also carries `configs`: every other BASELINE configuration at its named
size (C1, C2, C3, C4, C4M, the C5 sweep at N = 1024 ... 8192), each with value / e2e / ms_per_step / roofline /
cpu_baseline measured the same way (parity cases promoted to bench records; `--no-configs` skips them).
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

import numpy as np  # noqa: E402

METRIC = "gp_logl_evals_per_sec"
UNIT = "evals/s"

# (label, workload generator, N, walkers): the BASELINE configurations other than the headline, at their named sizes.
# C5 at N = 8192 runs 128 walkers (one step of 1024 walkers takes 5.7 s; the rate per walker is the same: waves of 128).
SIDE_CONFIGS = [("C1", "C1", 100, 100), ("C2", "C2", 300, 100), ("C3", "C3", 500, 256), ("C4", "C4", 4096, 512),
                ("C4M", "C4M", 4096, 512), ("C5_N1024", "C5", 1024, 1024), ("C5_N2048", "C5", 2048, 1024),
                ("C5_N4096", "C5", 4096, 1024), ("C5_N8192", "C5", 8192, 128)]
# CPU-baseline sample sizes (walkers evaluated by the reference) per label: about 1-12 s each on the box's cores
CPU_WALKERS = {"C1": 100, "C2": 24, "C3": 12, "C4": 1, "C4M": 1, "C5_N1024": 8, "C5_N2048": 3, "C5_N4096": 1, "C5_N8192": 1}
# reference run_MCMC (mcmc.py:676) iterations timed by the --impl reference arm (BASELINE.md section 3 step 3 asks for
# 50 / 100 / 100; C2 and C3 cost 0.07 / 0.13 s per evaluation in the reference, so they are bounded to fit minutes)
REF_MCMC_ITERS = {"C1": 50, "C2": 3, "C3": 2}   # the reference's progress bar divides by iterations - 1: at least 2


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--npoints", dest="n", type=int, default=16384, help="data points N of the C5 sweep")
    ap.add_argument("--walkers-per-gpu", type=int, default=128)
    ap.add_argument("--workload", dest="config", default="C5", help="workload generator (C1..C5)")
    ap.add_argument("--panel", type=int, default=0)
    ap.add_argument("--wave", type=int, default=0)
    ap.add_argument("--cpu-budget-s", type=float, default=None,
                    help="CPU seconds for the reference samples (default: 60 for the cpu_baseline leg, 150 for --impl reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` array (other BASELINE configurations)")
    ap.add_argument("--only-configs", default="", help="comma-separated labels of SIDE_CONFIGS to run (default: all)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--iterations", type=int, default=0, help="--scaling strong: run_MCMC iterations (default by size)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# CPU side: the reference itself (oracle/_ref) or the oracle port on the host cores -- the only place
# bench.py executes oracle/
# ---------------------------------------------------------------------------------------------
def host_threads():
    try:
        from threadpoolctl import threadpool_info
        n = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
        return int(n)
    except Exception:
        return usable_cores()


def blas_info():
    try:
        from threadpoolctl import threadpool_info
        return ["%s %s (%s threads)" % (p.get("internal_api"), p.get("version"), p.get("num_threads")) for p in threadpool_info()]
    except Exception:
        return []


def usable_cores():
    """Cores this process may run on (cgroup / affinity aware), not the machine's core count."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return int(os.cpu_count() or 1)


def set_host_threads(threads=None):
    """torchrun exports OMP_NUM_THREADS=1; the CPU baseline is meant to use every host core."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=usable_cores() if threads is None else int(threads))
    except Exception:
        pass


def _host_mem_bytes():
    try:
        return os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES")
    except Exception:
        return 32 << 30


class CpuArm:
    """The reference's CPU implementation of the path: `GPLikelihood(...).LogL` per walker exactly as MCMC.compute
    calls it (mcmc.py:421-458) from the staged unmodified package (kind "reference"), else the oracle port (kind "port")."""

    def __init__(self):
        from oracle import ref_loader
        self.ref = None
        try:
            if ref_loader.available():
                self.ref = ref_loader.load()
        except Exception:
            self.ref = None
        self.kind = "reference" if self.ref is not None else "port"

    def describe(self):
        if self.ref is not None:
            return "unmodified magpy_rv 1.0.4 GPLikelihood.LogL (LU slogdet + cho_factor/cho_solve, gp_likelihood.py:265-275)"
        return "oracle port (NumPy/SciPy LU slogdet + Cholesky, as the reference)"

    def seconds_per_eval(self, cfg, walkers):
        """Wall seconds per log-likelihood evaluation over the given walker indices (real evaluations)."""
        from oracle import logl_port as port
        textbook = bool(cfg["kernel_variant"])
        if self.ref is not None and not textbook:
            from oracle import ref_eval
            return ref_eval.time_logl(self.ref, cfg, walkers)[0], "reference"
        # textbook Matern 5/2 does not exist in the reference (its as-coded kernel is not PD): port
        t0 = time.perf_counter()
        for w in walkers:
            port.logL_one(cfg["t"], cfg["y"], cfg["yerr"], cfg["kernel"], cfg["hp"][w], cfg["model_list"], cfg["modpar"][w],
                          cfg["prior_list"], cfg["flags"], textbook_matern5=textbook)
        return (time.perf_counter() - t0) / len(walkers), "port"


def cpu_sample(arm, cfg_name, N_target, budget_s, threads=None, walkers=1):
    """Time real evaluations at the largest N_s <= N_target (halving from the target) whose predicted cost fits the
    budget (and host memory); N_s < N_target is scaled by the cubic law of the factorisations and says so.
    Returns (evals_per_s at N_target, description, seconds spent, kind)."""
    import workloads
    set_host_threads(threads)
    t_start = time.perf_counter()
    n = min(N_target, 1024)
    cfg = workloads.make_config(cfg_name, W=max(2, walkers), N=n)
    arm.seconds_per_eval(cfg, [0])                      # warm the BLAS threads
    dt, kind = arm.seconds_per_eval(cfg, [1])
    while n < N_target:
        nxt = min(N_target, n * 2)
        # go / no-go prediction: with threaded BLAS a doubling of N costs ~4x at these sizes (measured on the box:
        # t(16384) / t(8192) = 3.7, t(8192) / t(4096) = 4.0), not the asymptotic 8x
        pred = dt * (nxt / n) ** 2
        mem_ok = 7 * 8 * nxt * nxt < 0.6 * _host_mem_bytes()
        if pred > 0.9 * (budget_s - (time.perf_counter() - t_start)) or not mem_ok:
            break
        cfg = workloads.make_config(cfg_name, W=2, N=nxt)
        dt, kind = arm.seconds_per_eval(cfg, [1])
        n = nxt
    if n == N_target and walkers > 1:
        cfgw = workloads.make_config(cfg_name, W=max(2, walkers), N=n)
        dt, kind = arm.seconds_per_eval(cfgw, list(range(walkers)))
    scale = (N_target / n) ** 3
    desc = "%s, %d walker(s) at N=%d timed %.3f s per evaluation" % (arm.describe() if kind == "reference" else
                                                                      "oracle port (NumPy/SciPy, same algorithm)", walkers if n == N_target else 1, n, dt)
    if n != N_target:
        desc += ", scaled by (N/N_s)^3 = %.0f to N=%d" % (scale, N_target)
    return 1.0 / (dt * scale), desc, time.perf_counter() - t_start, kind


def config_cpu_baseline(arm, label, gen, N, W):
    """cpu_baseline of one side configuration: real evaluations of CPU_WALKERS[label] walkers at the named size."""
    import workloads
    set_host_threads(None)
    nw = min(CPU_WALKERS.get(label, 1), W)
    cfg = workloads.make_config(gen, W=max(2, nw), N=N)
    if nw <= 2 and N >= 1024:
        arm.seconds_per_eval(workloads.make_config(gen, W=2, N=512), [0])     # warm the BLAS threads
    dt, kind = arm.seconds_per_eval(cfg, list(range(nw)))
    return {"value": 1.0 / dt, "unit": UNIT, "cores": host_threads(), "kind": kind,
            "sample": "%d of %d walkers at N=%d, %.4f s per evaluation (%s)" % (nw, W, N, dt, arm.describe() if kind == "reference" else "oracle port")}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores (rank 0 alone).

    Headline steps are REAL evaluations of the unmodified `GPLikelihood.LogL`: step 0 at the full N (16384: ~15 GiB,
    about a minute; only when host RAM >= 32 GiB and N > 8192), the other steps at N_s = min(N, 8192) -- or 4096 once
    the wall budget is spent -- scaled to N by the MEASURED ratio of the step times (not by a cubic law).
    `ms_per_step` is the wall time of a step's sample, so steps x ms_per_step is what the clock around this run sees."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import workloads
    arm = CpuArm()
    set_host_threads(None)
    budget = 150.0 if args.cpu_budget_s is None else args.cpu_budget_s
    t_begin = time.perf_counter()
    N = args.n
    gen = args.config
    cfg_small = workloads.make_config(gen, W=2, N=min(N, 1024))
    for _ in range(max(1, args.warmup)):
        arm.seconds_per_eval(cfg_small, [0])
    sizes = sorted({N, min(N, 8192), min(N, 4096)}, reverse=True)
    cfgs = {n: workloads.make_config(gen, W=2, N=n) for n in sizes}
    t_at = {}
    samples = []
    full_ok = N <= 8192 or _host_mem_bytes() >= (32 << 30)
    kind = arm.kind
    for i in range(args.steps):
        left = budget - (time.perf_counter() - t_begin)
        steps_left = args.steps - i
        if N not in t_at and full_ok and i == 0:
            n = N
        else:
            n = sizes[1] if len(sizes) > 1 else sizes[0]
            if n in t_at and t_at[n] * steps_left > left and len(sizes) > 2:
                n = sizes[2]
            if n == N and not full_ok:
                n = sizes[1]
        dt, kind = arm.seconds_per_eval(cfgs[n], [1])
        t_at.setdefault(n, dt)
        samples.append((n, dt))
    # scale factors to the target size from measured step times: t(N) / t(n)
    def scale_to_target(n):
        if n == N:
            return 1.0, "measured"
        if N in t_at:
            return t_at[N] / t_at[n], "measured ratio t(%d)/t(%d) = %.2f" % (N, n, t_at[N] / t_at[n])
        big = max(t_at)
        if n == big:
            return (N / n) ** 3, "cubic law (N=%d not run: host RAM %.0f GiB)" % (N, _host_mem_bytes() / 2**30)
        s0, how = scale_to_target(big)
        return s0 * t_at[big] / t_at[n], how
    rates = []
    for n, dt in samples:
        sc, _ = scale_to_target(n)
        rates.append(1.0 / (dt * sc))
    value = float(np.mean(rates))
    counts = {n: sum(1 for m, _ in samples if m == n) for n in t_at}
    desc = "%s; steps: " % arm.describe() + ", ".join(
        "%d x one walker at N=%d (%.2f s each, %s)" % (counts[n], n, float(np.mean([d for m, d in samples if m == n])), scale_to_target(n)[1])
        for n in sorted(counts, reverse=True))
    cores = host_threads()
    ms_sample = 1e3 * float(np.mean([dt for _, dt in samples]))
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_sample, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc, "blas": blas_info(),
                         "host_cores": usable_cores(), "host_ram_gib": _host_mem_bytes() / 2**30},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        # a step of this arm is a bounded sample (one walker's evaluation); the whole step of the GPU arm
        # (walkers_per_gpu x gpus evaluations) would take step_ms_at_this_rate
        "step_ms_at_this_rate": 1e3 * args.walkers_per_gpu * args.gpus / value,
        "measured_at_full_size": N in t_at,
    }
    if not args.no_configs:
        out["configs"] = reference_side_configs(arm, args)
    print(json.dumps(out), flush=True)


def reference_side_configs(arm, args):
    """The other BASELINE configurations through the reference: its own run_MCMC (mcmc.py:676) end to end for C1-C3
    (bounded iteration counts), per-walker LogL samples for the rest."""
    import workloads
    only = [x for x in args.only_configs.split(",") if x]
    out = []
    for label, gen, N, W in SIDE_CONFIGS:
        if only and label not in only:
            continue
        entry = {"label": label, "impl": "reference", "N": N, "walkers": W}
        try:
            entry["cpu_baseline"] = config_cpu_baseline(arm, label, gen, N, W)
            entry["value"] = entry["cpu_baseline"]["value"]
            entry["unit"] = UNIT
            if label in REF_MCMC_ITERS and arm.ref is not None:
                from oracle import ref_eval
                cfg = workloads.make_config(gen, W=W, N=N)
                evals, dt, _ = ref_eval.run_mcmc(arm.ref, cfg, REF_MCMC_ITERS[label])
                entry["run_mcmc"] = {"iterations": REF_MCMC_ITERS[label], "evals": evals, "seconds": dt, "evals_per_s": evals / dt,
                                     "ms_per_iteration": 1e3 * dt / (REF_MCMC_ITERS[label] + 1),
                                     "what": "unmodified reference run_MCMC end to end (mcmc.py:676-874), seeded, %d chains" % W}
        except Exception as exc:  # noqa: BLE001
            entry["error"] = "%s: %s" % (type(exc).__name__, exc)
        out.append(entry)
    return out


# ---------------------------------------------------------------------------------------------
def load_captures():
    """ncu --set full captures of the dominant kernel, keyed by launch geometry (profiles/traffic.json)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return {}


def trailing_update_bytes(nt, PW, B):
    """Algorithmic DRAM bytes per trailing-update launch, from the launch geometry: every target
    tile of columns [P1, nt) is read and written once (2 x 128 KiB) and every operand tile row (PW tiles of 128 KiB) of
    rows [P1, nt) is read once; averaged over the launches of one wave of B walkers."""
    tile = 128 * 128 * 8
    per_launch = []
    for P0 in range(0, nt, PW):
        P1 = min(P0 + PW, nt)
        if P1 >= nt:
            break
        targets = sum(nt - j for j in range(P1, nt))
        per_launch.append(B * tile * (2 * targets + (nt - P1) * (P1 - P0)))
    return float(np.mean(per_launch)) if per_launch else None


WORKLOAD_TEXT = {
    "C1": "tutorial-1 style: QuasiPer GP, no mean model, uniform priors",
    "C2": "tutorial-2 style: QuasiPer GP + Polynomial mean model, uniform priors",
    "C3": "51-Peg style: JitterQuasiPer GP + Keplerian mean model (per-point Newton solve), uniform priors",
    "C4": "sun-as-a-star style: QuasiPer GP + two instrument Offsets, Uniform / Jeffrey / Gaussian priors",
    "C4M": "sun-as-a-star style: textbook Matern 5/2 GP + two instrument Offsets",
    "C5": "scaling sweep: QuasiPer GP, no mean model, no priors",
}


def workload_config(args, gen=None, N=None, W=None, gpus=None):
    gen = args.config if gen is None else gen
    N = args.n if N is None else N
    W = args.walkers_per_gpu if W is None else W
    gpus = args.gpus if gpus is None else gpus
    big = N >= 2048
    return {"workload": "%s %s; N=%d points, %d walkers per GPU (%d walkers total); %s"
                        % (gen, WORKLOAD_TEXT.get(gen, ""), N, W, W * gpus,
                           "factor workspace >> L2 so no flush is needed between steps" if big else
                           "K is rebuilt from t / yerr every step and never stored; the per-step inputs are the walker parameters"),
            "N": N, "walkers_per_gpu": W, "walkers_total": W * gpus,
            "kernel": {"C3": "JitterQuasiPer", "C4M": "Matern5/2 (textbook variant)"}.get(gen, "QuasiPer"),
            "l2_policy": "inputs larger than L2" if big else "nothing to flush: covariance generated on chip each step"}


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.lines = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=3)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


PROF_CATS = ("gen", "potrf", "trsm", "update_panel", "update_mid", "update_trail", "mean", "final", "fused", "dag")
def dag_bytes(nt, B):
    """Algorithmic bytes one launch of the dataflow kernel moves between L2 / HBM and the SMs, for a
    wave of B walkers: every off-diagonal tile (i, k) streams the k operand tiles of rows i and k (2 k x 128 KiB) and the
    inverse diagonal factor (128 KiB) and is written once; every diagonal tile streams its own row once (k x 128 KiB) and
    writes the inverse factor.  16 flop per operand byte at every N."""
    tile = 128 * 128 * 8
    per_walker = 0
    for k in range(nt):
        per_walker += (nt - 1 - k) * (2 * k + 2) * tile      # off-diagonal tiles of column k: operands + inv(L_kk) + the tile itself
        per_walker += (k + 1) * tile                          # diagonal tile: its row once + inv(L_kk) out
    return float(B * per_walker)


KERNEL_NAMES = {
    "update_trail": "gemm_tile_kernel (trailing SYRK/GEMM update, DMMA)",
    "update_panel": "gemm_tile_kernel (in-panel SYRK/GEMM update, DMMA)",
    "dag": "tile_dag_kernel (persistent dataflow factorisation: generate + SYRK/GEMM + TRSM + diagonal tiles, DMMA)",
}


def read_prof(prob):
    prof = {}
    for cat in PROF_CATS:
        rec = {k: prob.info("prof_%s_%s" % (k, cat)) for k in ("ns", "alg", "exec", "n")}
        if rec["n"] is not None and rec["n"] > 0:
            prof[cat] = rec
    return prof


def make_roofline(prob, prof, steps, ms, N, W_local, value_per_gpu, peaks, cov_cost, desc):
    """roofline object of one measured configuration.  Tiled paths: bound "tensor", the dominant kernel's algorithmic
    FLOP/s (CUDA events around its launches) against the DMMA peak.  One-CTA-per-walker paths (N <= 512): bound
    "fp64" -- factorisation AND covariance generation share the one FP64 datapath (tools/micro/pipes.cu), so the
    work is N^3/3 + 2 N^2 + c N(N+1)/2 with c the measured cost of one covariance entry in flop-equivalents."""
    import workloads
    path = prob.info("path")
    step_ns = ms * 1e6 / steps
    shares = {k: (v["ns"] / steps) / step_ns for k, v in prof.items()}
    flops_eval = workloads.flops_per_eval(N)
    peak = peaks["dmma"] / 1e12
    n_entries = N * (N + 1) / 2.0
    gen_flops = cov_cost["flops_per_entry"] * n_entries if cov_cost else 0.0
    out = {"peak": peak, "unit": "TFLOP/s",
           "peak_source": "FP64 DMMA issue microbenchmark run in this process (MEASURED_PEAKS.json has no FP64 entry); "
                          "DFMA vector pipe measured %.1f TFLOP/s" % (peaks["dfma"] / 1e12),
           "kernel_share_of_step": shares,
           "whole_step_frac": value_per_gpu * flops_eval / peaks["dmma"],
           "whole_step_frac_with_generation": value_per_gpu * (flops_eval + gen_flops) / peaks["dmma"],
           "flops_per_eval": flops_eval, "cov_entries_per_eval": n_entries,
           "cov_flop_equivalents_per_entry": cov_cost["flops_per_entry"] if cov_cost else None}
    in_bytes = 8.0 * (desc["nh"] + desc["nm"]) * W_local
    out_bytes = 12.0 * W_local
    data_bytes = 8.0 * N * (4 if desc["flags"] else 3)
    if path in (1, 3):
        d = prof.get("fused")
        work = (flops_eval + gen_flops) * W_local
        achieved = (work * d["n"] / (d["ns"] * 1e-9)) / 1e12 if d and d["ns"] > 0 else None
        out.update({
            "bound": "fp64",
            "kernel": {1: "fused_small_blk_kernel (one CTA per walker, K built and factored in shared memory)",
                       3: "mid_logl_kernel (one CTA per walker, streamed panel, factor L2-resident)"}[path],
            "achieved": achieved, "frac": achieved / peak if achieved else None,
            "achieved_factorisation_only": (flops_eval * W_local * d["n"] / (d["ns"] * 1e-9)) / 1e12 if d and d["ns"] > 0 else None,
            "launches": (d["n"] // steps) if d else None, "avg_launch_ms": d["ns"] / max(1, d["n"]) * 1e-6 if d else None,
            "work_model": "per evaluation N^3/3 + 2 N^2 factorisation flops + c N(N+1)/2 with c = %.0f FP64 flop-equivalents per "
                          "covariance entry (DFMA peak / measured evaluation rate of this kernel, mrv_cov_eval_rate)"
                          % (cov_cost["flops_per_entry"] if cov_cost else float("nan")),
            # K never touches HBM on these paths: compulsory traffic is the walker parameters in, 12 B per walker out
            # and one read of t / y / yerr (/ flags); every CTA re-reads those vectors from L2
            "hbm_gbs": (in_bytes + out_bytes + data_bytes) / (step_ns * 1e-9) / 1e9,
            "l2_gbs_data_vectors": data_bytes * W_local / (step_ns * 1e-9) / 1e9,
            "traffic": None, "traffic_note": "no ncu --set full capture of this configuration in this run; profiles/README.md "
                                             "lists the captures (N = 100: fused kernel, N = 300 / 512: streamed panel)"})
        return out
    dom = max((c for c in ("update_trail", "dag", "update_panel") if c in prof), key=lambda c: prof[c]["ns"], default=None)
    if dom is None:
        out.update({"bound": "tensor", "kernel": None, "achieved": None, "frac": None, "traffic": None})
        return out
    d = prof[dom]
    achieved = (d["alg"] / (d["ns"] * 1e-9)) / 1e12
    nt, PW, B = (N + 127) // 128, prob.info("panel_tiles"), prob.info("wave_walkers")
    geom = "N%d_wave%d_panel%d_%s" % (N, B, PW, dom)
    cap = load_captures().get(geom)
    out.update({
        "bound": "tensor", "kernel": KERNEL_NAMES.get(dom, dom), "achieved": achieved, "frac": achieved / peak,
        "launches": d["n"] // steps, "avg_launch_ms": d["ns"] / max(1, d["n"]) * 1e-6,
        "executed_tflops": (d["exec"] / (d["ns"] * 1e-9)) / 1e12,
        "launch_geometry": geom,
        "algorithmic_bytes_per_launch": trailing_update_bytes(nt, PW, B) if dom == "update_trail" else (dag_bytes(nt, B) if dom == "dag" else None),
        "l2_or_hbm_gbs": (dag_bytes(nt, B) / (d["ns"] / max(1, d["n"]) * 1e-9) / 1e9) if dom == "dag" else None,
        "traffic": cap["dram_bytes_per_launch"] if cap else None,
        "traffic_source": ("ncu --set full capture %s" % cap.get("file", "under profiles/")) if cap else
                          "no ncu capture of this launch geometry (profiles/traffic.json); algorithmic_bytes_per_launch is derived from the geometry",
        "hbm_gbs": None})
    return out


def measure_one(torch, engine, lib, dev, local_rank, cfg, W_local, steps, warmup, peaks, cov_costs, opts=None, time_budget_s=None):
    """Device-timed and end-to-end (host buffers) evaluation rate of one configuration on this GPU."""
    prob = engine.LikelihoodProblem(cfg["t"], cfg["y"], cfg["yerr"], cfg["kernel"], cfg["model_list"], cfg["prior_list"],
                                    flags=cfg["flags"], device=local_rank, kernel_variant=cfg["kernel_variant"])
    for k, v in (opts or {}).items():
        prob.set_option(k, v)
    hp_host = torch.from_numpy(cfg["hp"][:W_local].copy()).pin_memory()
    mp_host = torch.from_numpy(cfg["modpar"][:W_local].copy()).pin_memory()
    d_hp, d_mp = hp_host.to(dev), mp_host.to(dev)
    d_l = torch.empty(W_local, dtype=torch.float64, device=dev)
    d_s = torch.empty(W_local, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()

    def step():
        prob.logl_device(d_hp.data_ptr(), d_mp.data_ptr(), W_local, d_l.data_ptr(), d_s.data_ptr(), to_ecc=True,
                         stream=stream.cuda_stream)

    step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step()
    torch.cuda.synchronize()
    t_one = time.perf_counter() - t0
    if time_budget_s is not None:
        steps = int(max(3, min(200, time_budget_s / max(t_one, 1e-6))))
    for _ in range(max(0, warmup - 2)):
        step()
    torch.cuda.synchronize()
    prob.set_option("profile", 1)
    launches0 = prob.info("launches")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = prob.info("launches") - launches0
    prof = read_prof(prob)
    prob.set_option("profile", 0)
    status_ok = bool((d_s == 0).all().item())
    # end to end through the host-buffer C ABI call
    hp_np, mp_np = hp_host.numpy(), mp_host.numpy()
    prob.logl(hp_np, mp_np)
    e2e_steps = max(2, min(steps, int(max(2, (time_budget_s or 2.0) / max(t_one, 1e-6))), 100))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        l_host, s_host = prob.logl(hp_np, mp_np)
    t_e2e = time.perf_counter() - t0
    value = W_local * steps / (ms * 1e-3)
    key = (cfg["kernel"], cfg["kernel_variant"])
    desc = {"nh": cfg["hp"].shape[1], "nm": cfg["modpar"].shape[1], "flags": cfg["flags"] is not None}
    roof = make_roofline(prob, prof, steps, ms, cfg["N"], W_local, value, peaks, cov_costs.get(key), desc)
    out = {"value": value, "unit": UNIT, "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms / steps,
           "e2e": {"value": W_local * e2e_steps / t_e2e, "unit": UNIT, "steps": e2e_steps,
                   "h2d_bytes_per_step": int(hp_np.nbytes + mp_np.nbytes), "d2h_bytes_per_step": int(l_host.nbytes + s_host.nbytes),
                   "matches_device_path": bool(np.array_equal(l_host, d_l.cpu().numpy()))},
           "gpu_launches": int(launches), "roofline": roof, "status_ok": status_ok, "path": prob.info("path"),
           "wave_walkers": prob.info("wave_walkers"), "panel_tiles": prob.info("panel_tiles"),
           "tflops_per_gpu": value * roof["flops_per_eval"] / 1e12}
    prob.close()
    return out


def cov_cost_table(lib, local_rank, peaks, cfgs):
    """FP64 flop-equivalents per covariance entry for each (kernel, variant) in use: DFMA peak / measured rate."""
    from magpy_rv_b200 import _native as nat
    from magpy_rv_b200 import engine
    table = {}
    for cfg in cfgs:
        key = (cfg["kernel"], cfg["kernel_variant"])
        if key in table:
            continue
        kid = engine.KERNEL_IDS[engine.canonical_kernel(cfg["kernel"])]
        hp = np.ascontiguousarray(cfg["hp"][0], dtype=np.float64)
        rate = lib.mrv_cov_eval_rate(local_rank, kid, int(cfg["kernel_variant"]), nat.dptr(hp), hp.shape[0])
        if rate > 0:
            table[key] = {"evals_per_s": rate, "flops_per_entry": peaks["dfma"] / rate}
    return table


def run_ours(args):
    import torch
    import torch.distributed as dist
    import workloads
    from magpy_rv_b200 import _native, engine
    from magpy_rv_b200.sharding import GatherBuffers, partition

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- magpy_rv_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world

    Wl = args.walkers_per_gpu
    Wtot = Wl * n_gpus
    cfg = workloads.make_config(args.config, W=Wtot, N=args.n)     # identical on every rank (seeded)
    a, b = partition(Wtot, n_gpus)[rank]
    prob = engine.LikelihoodProblem(cfg["t"], cfg["y"], cfg["yerr"], cfg["kernel"], cfg["model_list"], cfg["prior_list"],
                                    flags=cfg["flags"], device=local_rank, kernel_variant=cfg["kernel_variant"])
    if world > 1:
        prob.set_option("ensemble_walkers", Wtot)     # kernel choices by the whole ensemble: results independent of N GPUs
    if args.panel:
        prob.set_option("panel_tiles", args.panel)
    if args.wave:
        prob.set_option("wave_walkers", args.wave)
    hp_host = torch.from_numpy(cfg["hp"][a:b].copy()).pin_memory()
    mp_host = torch.from_numpy(cfg["modpar"][a:b].copy()).pin_memory()
    d_hp = hp_host.to(dev)
    d_mp = mp_host.to(dev)
    stream = torch.cuda.current_stream()
    gb = GatherBuffers(Wtot, world, rank, dev) if world > 1 else None
    if gb is not None:
        d_l, d_s = gb.local_views()
    else:
        d_l = torch.empty(b - a, dtype=torch.float64, device=dev)
        d_s = torch.empty(b - a, dtype=torch.int32, device=dev)

    def step():
        prob.logl_device(d_hp.data_ptr(), d_mp.data_ptr(), b - a, d_l.data_ptr(), d_s.data_ptr(), to_ecc=True,
                         stream=stream.cuda_stream)
        if gb is not None:
            return gb.gather(dist)
        return d_l, d_s

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    lib = _native.lib()
    peaks = {"dmma": lib.mrv_fp64_mma_peak(local_rank, 20000) if rank == 0 else 0.0,
             "dfma": lib.mrv_fp64_fma_peak(local_rank, 20000) if rank == 0 else 0.0}

    for _ in range(args.warmup):
        step()
    sync_all()
    prob.set_option("profile", 1)
    launches0 = prob.info("launches")
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record(stream)
    for _ in range(args.steps):
        logl, status = step()
    e1.record(stream)
    sync_all()
    clock_info = clocks.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    launches = prob.info("launches") - launches0
    prof = read_prof(prob)
    prob.set_option("profile", 0)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    status_ok = bool((status == 0).all().item())

    # end to end through the host-buffer C ABI call (host parameters in, log L out)
    hp_np, mp_np = hp_host.numpy(), mp_host.numpy()
    prob.logl(hp_np, mp_np)
    sync_all()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 2))
    for _ in range(e2e_steps):
        l_host, s_host = prob.logl(hp_np, mp_np)
        if gb is not None:
            gb.gather_host(l_host, s_host, dist)
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = Wtot * e2e_steps / float(t_e2e.item())
    e2e_ok = bool(np.array_equal(l_host, d_l.cpu().numpy()))

    out = None
    if rank == 0:
        value = Wtot * args.steps / (ms * 1e-3)
        cov_costs = cov_cost_table(lib, local_rank, peaks, [cfg])
        desc = {"nh": cfg["hp"].shape[1], "nm": cfg["modpar"].shape[1], "flags": cfg["flags"] is not None}
        roofline = make_roofline(prob, prof, args.steps, ms, args.n, b - a, value / n_gpus, peaks, cov_costs.get((cfg["kernel"], cfg["kernel_variant"])), desc)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args, gpus=n_gpus),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(hp_np.nbytes + mp_np.nbytes),
                    "d2h_bytes_per_step": int(l_host.nbytes + s_host.nbytes), "matches_device_path": e2e_ok},
            "gpu_launches": int(launches), "clocks": clock_info, "roofline": roofline,
            "status_ok": status_ok, "path": prob.info("path"), "wave_walkers": prob.info("wave_walkers"),
            "panel_tiles": prob.info("panel_tiles"),
            "tflops_per_gpu": (value / n_gpus) * workloads.flops_per_eval(args.n) / 1e12,
        }
    prob.close()
    del d_hp, d_mp, d_l, d_s
    torch.cuda.empty_cache()

    # the other BASELINE configurations at their named sizes (one GPU only: parity cases promoted to bench records)
    side = []
    if rank == 0 and world == 1 and not args.no_configs:
        only = [x for x in args.only_configs.split(",") if x]
        plan = [(label, gen, N, W) for label, gen, N, W in SIDE_CONFIGS if not only or label in only]
        cfgs = {label: workloads.make_config(gen, W=W, N=N) for label, gen, N, W in plan}
        cov_costs = cov_cost_table(lib, local_rank, peaks, list(cfgs.values()))
        for label, gen, N, W in plan:
            entry = {"label": label, "config": workload_config(args, gen, N, W, 1)}
            try:
                entry.update(measure_one(torch, engine, lib, dev, local_rank, cfgs[label], W, 0, 3, peaks, cov_costs,
                                         time_budget_s=1.5))
            except Exception as exc:  # noqa: BLE001
                entry["error"] = "%s: %s" % (type(exc).__name__, exc)
            side.append(entry)
            torch.cuda.empty_cache()

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        # CPU baseline last: the other ranks have exited, so the host cores are free for the reference
        if not args.no_cpu_baseline:
            arm = CpuArm()
            budget = 60.0 if args.cpu_budget_s is None else args.cpu_budget_s
            v, desc, _, kind = cpu_sample(arm, args.config, args.n, budget)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": host_threads(), "kind": kind, "sample": desc,
                                   "blas": blas_info(), "host_cores": usable_cores(),
                                   "note": "one real evaluation at the full N when it fits the budget and host RAM; any N^3 scaling is stated in `sample`"}
            v1, desc1, _, kind1 = cpu_sample(arm, args.config, args.n, min(8.0, budget), threads=1)
            out["cpu_baseline_1thread"] = {"value": v1, "unit": UNIT, "cores": 1, "kind": kind1, "sample": desc1}
            for entry in side:
                if "error" in entry:
                    continue
                label, gen, N, W = next(x for x in SIDE_CONFIGS if x[0] == entry["label"])
                try:
                    entry["cpu_baseline"] = config_cpu_baseline(arm, label, gen, N, W)
                except Exception as exc:  # noqa: BLE001
                    entry["cpu_baseline"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
        if side:
            out["configs"] = side
        print(json.dumps(out), flush=True)


def run_strong(args):
    """--scaling strong: a FIXED ensemble (C4: 512 walkers x 4096 points, or C5: 1024 x 1024) through the sharded
    run_MCMC end to end (host proposals, batched GPU likelihood on each rank's block, one all-gather per iteration,
    host accept / reject) -- the regime where the collective and the replicated host loop matter."""
    import contextlib
    import io
    import random
    import torch
    import torch.distributed as dist
    import workloads
    from magpy_rv_b200 import mcmc

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    os.environ["MRV_DEVICE"] = str(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    gen = args.config
    N, W = {"C4": (4096, 512), "C5": (1024, 1024), "C3": (500, 256)}.get(gen, (args.n, args.walkers_per_gpu))
    cfg = workloads.make_config(gen, W=W, N=N)
    iters = args.iterations or (5 if N >= 4096 else 50)
    random.seed(cfg["seed"])
    np.random.seed(cfg["seed"])
    with contextlib.redirect_stdout(io.StringIO()):
        m = mcmc.MCMC(cfg["t"], cfg["y"], cfg["yerr"], cfg["hparam0"], cfg["kernel"], cfg["model_par0"], cfg["model_list"],
                      cfg["prior_list"], numb_chains=W, flags=cfg["flags"])
    tt = {"split_step": 0.0, "compute": 0.0, "compare+reset": 0.0}
    warm = max(3, args.warmup)
    for it in range(warm + iters):
        if it == warm:
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            tt = {k: 0.0 for k in tt}
            t_start = time.perf_counter()
        t0 = time.perf_counter()
        m.split_step()
        t1 = time.perf_counter()
        m.compute()
        t2 = time.perf_counter()
        m.compare()
        m.reset()
        t3 = time.perf_counter()
        tt["split_step"] += t1 - t0
        tt["compute"] += t2 - t1
        tt["compare+reset"] += t3 - t2
    torch.cuda.synchronize()
    wall = time.perf_counter() - t_start
    tw = torch.tensor([wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
    wall = float(tw.item())
    import hashlib
    digest = hashlib.sha256(np.ascontiguousarray(m.hparameter_list).tobytes() + np.ascontiguousarray(m.logL_list).tobytes()).hexdigest()[:16]
    if rank == 0:
        out = {"metric": METRIC, "value": W * iters / wall, "unit": UNIT, "n_gpus": world, "steps": iters, "warmup": warm,
               "ms_per_step": 1e3 * wall / iters, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "f64", "data": "synthetic", "config": workload_config(args, gen, N, W // world, world),
               "what": "sharded run_MCMC end to end (wall clock, max over ranks): host split_step + batched GPU compute + all-gather + host compare/reset",
               "host_ms_per_iteration": {k: 1e3 * v / iters for k, v in tt.items()}, "chain_digest": digest}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    env_world = int(os.environ.get("WORLD_SIZE", "1"))
    if env_world > 1:
        args.gpus = env_world
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.gpus > 1 and env_world == 1:
            # convenience: relaunch under torchrun when called directly with --gpus N
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
                   "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
            raise SystemExit(subprocess.call(cmd))
        if args.scaling == "strong":
            run_strong(args)
        else:
            run_ours(args)


if __name__ == "__main__":
    main()
