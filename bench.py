#!/usr/bin/env python
"""Benchmark of the GP log-likelihood hot path (BASELINE.json metric: GP LogL evaluations per
second, fp64, walkers x N up to 1024 x 16k, at 1/2/4/8 B200).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

Workload (config.workload): BASELINE configs[4] "scaling sweep" at its largest size, N = 16384
points, QuasiPer GP, no mean model: 128 walkers per GPU, i.e. exactly 1024 walkers x 16384 points
on 8 GPUs (weak scaling: per-GPU work fixed).  A "step" is one batched evaluation of every walker
(MCMC.compute of one iteration) followed, for N > 1 ranks, by the all-gather of log L / status.

One JSON line on rank 0.  `value` is device-timed (CUDA events on the launching stream, max over
ranks) with walker parameters already resident in HBM; `e2e` is the same metric through the
host-buffer C ABI call (`mrv_logl_batch`: pinned host parameters in, log L out, copies inside the
timed region); `roofline` is the dominant kernel (trailing SYRK/GEMM update on the FP64 tensor pipe)
against the DMMA issue peak measured in the same run; `cpu_baseline` is the oracle port on the
host cores for a bounded sample.
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
    ap.add_argument("--cpu-budget-s", type=float, default=20.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# CPU side: the oracle port on the host cores (the only place bench.py executes oracle/)
# ---------------------------------------------------------------------------------------------
def host_threads():
    try:
        from threadpoolctl import threadpool_info
        n = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
        return int(n)
    except Exception:
        return usable_cores()


def usable_cores():
    """Cores this process may run on (cgroup / affinity aware), not the machine's core count."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return int(os.cpu_count() or 1)


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU baseline is meant to use every host core."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=usable_cores())
    except Exception:
        pass


def cpu_sample(cfg_name, N_target, budget_s, threads=None):
    """Time the oracle's LogL on a bounded sample: the largest N_s <= N_target (halving from the
    target) whose predicted cost fits the budget; scale to N_target by the cubic law of the
    factorisations.  Returns (evals_per_s at N_target, description, seconds spent)."""
    import workloads
    from oracle import logl_port as port
    if threads is None:
        use_all_host_threads()
    else:
        try:
            from threadpoolctl import threadpool_limits
            threadpool_limits(limits=int(threads))
        except Exception:
            pass
    t_start = time.perf_counter()
    n = min(N_target, 1024)
    cfg = workloads.make_config(cfg_name, W=2, N=n)

    def one(cfg, reps):
        t0 = time.perf_counter()
        for r in range(reps):
            port.logL_one(cfg["t"], cfg["y"], cfg["yerr"], cfg["kernel"], cfg["hp"][r % 2], cfg["model_list"],
                          cfg["modpar"][r % 2], cfg["prior_list"], cfg["flags"])
        return (time.perf_counter() - t0) / reps

    one(cfg, 1)
    dt = one(cfg, 2)
    # grow while the cubic prediction fits into the remaining budget (and host memory: ~6 N^2 doubles)
    while n < N_target:
        nxt = min(N_target, n * 2)
        pred = dt * (nxt / n) ** 3
        mem_ok = 6 * 8 * nxt * nxt < 0.5 * _host_mem_bytes()
        if pred > 0.6 * (budget_s - (time.perf_counter() - t_start)) or not mem_ok:
            break
        cfg = workloads.make_config(cfg_name, W=2, N=nxt)
        dt = one(cfg, 1)
        n = nxt
    scale = (N_target / n) ** 3
    evals_per_s = 1.0 / (dt * scale)
    desc = ("oracle port (NumPy/SciPy LU slogdet + Cholesky, as the reference), 1 walker at N=%d timed %.3f s"
            % (n, dt))
    if n != N_target:
        desc += ", scaled by (N/N_s)^3 = %.0f to N=%d" % (scale, N_target)
    return evals_per_s, desc, time.perf_counter() - t_start


def _host_mem_bytes():
    try:
        return os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES")
    except Exception:
        return 32 << 30


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm for the path on the host cores.  Rank 0
    alone works; each step is a bounded sample scaled to the named workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, walls = [], []
    desc = ""
    per_step = max(5.0, args.cpu_budget_s)
    for i in range(args.warmup + args.steps):
        v, desc, wall = cpu_sample(args.config, args.n, per_step if i >= args.warmup else min(per_step, 8.0))
        if i >= args.warmup:
            vals.append(v)
            walls.append(wall)
    value = float(np.mean(vals))
    cores = host_threads()
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * args.walkers_per_gpu * args.gpus / value, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        # ms_per_step is the time the whole step (walkers_per_gpu x gpus evaluations at N) would take at the sampled
        # rate; the sample itself (what the wall clock around this run sees) took sample_wall_ms_per_step
        "sample_wall_ms_per_step": 1e3 * float(np.mean(walls)),
    }
    print(json.dumps(out), flush=True)


def measured_traffic(args, prob):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture of this
    very configuration (profiles/traffic.json); None when no capture matches."""
    try:
        table = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = "%s_N%d_wave%d_panel%d" % (args.config, args.n, prob.info("wave_walkers"), prob.info("panel_tiles"))
        return table[key]["dram_bytes_per_launch"] if key in table else None
    except Exception:
        return None


WORKLOAD_TEXT = {
    "C1": "tutorial-1 style: QuasiPer GP, no mean model, uniform priors",
    "C2": "tutorial-2 style: QuasiPer GP + Polynomial mean model, uniform priors",
    "C3": "51-Peg style: JitterQuasiPer GP + Keplerian mean model (per-point Newton solve), uniform priors",
    "C4": "sun-as-a-star style: QuasiPer GP + two instrument Offsets, Uniform / Jeffrey / Gaussian priors",
    "C4M": "sun-as-a-star style: textbook Matern 5/2 GP + two instrument Offsets",
    "C5": "scaling sweep: QuasiPer GP, no mean model, no priors",
}


def workload_config(args):
    big = args.n >= 2048
    return {"workload": "%s %s; N=%d points, %d walkers per GPU (%d walkers total); %s"
                        % (args.config, WORKLOAD_TEXT.get(args.config, ""), args.n, args.walkers_per_gpu,
                           args.walkers_per_gpu * args.gpus,
                           "factor workspace >> L2 so no flush is needed between steps" if big else
                           "K is rebuilt from t / yerr every step and never stored; the per-step inputs are the walker parameters"),
            "N": args.n, "walkers_per_gpu": args.walkers_per_gpu, "walkers_total": args.walkers_per_gpu * args.gpus,
            "kernel": {"C3": "JitterQuasiPer", "C4M": "Matern5/2 (textbook variant)"}.get(args.config, "QuasiPer"),
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


def run_ours(args):
    import torch
    import torch.distributed as dist
    import workloads
    from magpy_rv_b200 import _native, engine
    from magpy_rv_b200.sharding import gather_blocks, partition

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
    d_l = torch.empty(b - a, dtype=torch.float64, device=dev)
    d_s = torch.empty(b - a, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()

    def step():
        prob.logl_device(d_hp.data_ptr(), d_mp.data_ptr(), b - a, d_l.data_ptr(), d_s.data_ptr(), to_ecc=True,
                         stream=stream.cuda_stream)
        if world > 1:
            return gather_blocks(d_l, d_s, Wtot, dist, dev)
        return d_l, d_s

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    lib = _native.lib()
    dmma_peak = lib.mrv_fp64_mma_peak(local_rank, 20000) if rank == 0 else 0.0
    dfma_peak = lib.mrv_fp64_fma_peak(local_rank, 20000) if rank == 0 else 0.0

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
    prof = {}
    for cat in ("gen", "potrf", "trsm", "update_panel", "update_mid", "update_trail", "mean", "final", "fused"):
        prof[cat] = {k: prob.info("prof_%s_%s" % (k, cat)) for k in ("ns", "alg", "exec", "n")}
    prob.set_option("profile", 0)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    status_ok = bool((status == 0).all().item())

    # end to end through the host-buffer C ABI call (pinned host parameters in, log L out)
    hp_np, mp_np = hp_host.numpy(), mp_host.numpy()
    prob.logl(hp_np, mp_np)
    sync_all()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 2))
    for _ in range(e2e_steps):
        l_host, s_host = prob.logl(hp_np, mp_np)
        if world > 1:
            gather_blocks(torch.from_numpy(l_host).to(dev), torch.from_numpy(s_host).to(dev), Wtot, dist, dev)[0].cpu()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = Wtot * e2e_steps / float(t_e2e.item())
    e2e_ok = bool(np.array_equal(l_host, d_l.cpu().numpy()))

    if rank == 0:
        value = Wtot * args.steps / (ms * 1e-3)
        flops_eval = workloads.flops_per_eval(args.n)
        dom = "update_trail" if prof["update_trail"]["n"] > 0 else ("fused" if prof["fused"]["n"] > 0 else "update_panel")
        d = prof[dom]
        achieved = (d["alg"] / (d["ns"] * 1e-9)) / 1e12 if d["ns"] > 0 else None
        peak = dmma_peak / 1e12
        step_ns = ms * 1e6 / args.steps
        shares = {k: (v["ns"] / args.steps) / step_ns for k, v in prof.items() if v["n"] > 0}
        roofline = {
            "bound": "tensor", "kernel": "gemm_tile_kernel (trailing SYRK/GEMM update, DMMA)" if dom != "fused" else
            {1: "fused_small_blk_kernel (one CTA per walker, K in shared memory)",
             3: "mid_logl_kernel (one CTA per walker, streamed panel)"}.get(prob.info("path"), "fused"),
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": (achieved / peak) if achieved else None,
            "traffic": measured_traffic(args, prob),
            "peak_source": "FP64 DMMA issue microbenchmark run in this process (MEASURED_PEAKS.json has no FP64 entry); "
                           "DFMA vector pipe measured %.1f TFLOP/s" % (dfma_peak / 1e12),
            "launches": d["n"] // args.steps, "avg_launch_ms": d["ns"] / max(1, d["n"]) * 1e-6,
            "executed_tflops": (d["exec"] / (d["ns"] * 1e-9)) / 1e12 if d["ns"] > 0 else None,
            "kernel_share_of_step": shares,
            "whole_step_frac": (value / n_gpus) * flops_eval / dmma_peak,
        }
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(hp_np.nbytes + mp_np.nbytes),
                    "d2h_bytes_per_step": int(l_host.nbytes + s_host.nbytes), "matches_device_path": e2e_ok},
            "gpu_launches": int(launches), "clocks": clock_info, "roofline": roofline,
            "status_ok": status_ok, "path": prob.info("path"), "wave_walkers": prob.info("wave_walkers"),
            "panel_tiles": prob.info("panel_tiles"),
            "tflops_per_gpu": (value / n_gpus) * flops_eval / 1e12,
        }
    prob.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        # CPU baseline last: the other ranks have exited, so the host cores are free for the oracle
        if not args.no_cpu_baseline:
            v, desc, _ = cpu_sample(args.config, args.n, args.cpu_budget_s)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": host_threads(), "kind": "port", "sample": desc}
            # per-core figure beside the all-cores one; small extra budget
            v1, desc1, _ = cpu_sample(args.config, args.n, min(8.0, args.cpu_budget_s), threads=1)
            out["cpu_baseline_1thread"] = {"value": v1, "unit": UNIT, "cores": 1, "kind": "port", "sample": desc1}
        print(json.dumps(out), flush=True)


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
        run_ours(args)


if __name__ == "__main__":
    main()
