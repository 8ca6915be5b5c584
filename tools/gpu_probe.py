"""Quick GPU probe: FP64 DMMA/DFMA issue peaks and coarse LogL timings by size (CUDA events via torch)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import workloads
from magpy_rv_b200 import _native, engine

lib = _native.lib()
print("DMMA peak TF/s:", lib.mrv_fp64_mma_peak(0, 20000) / 1e12, " DFMA peak TF/s:", lib.mrv_fp64_fma_peak(0, 20000) / 1e12, flush=True)
path_opt = int(os.environ.get("MRV_PATH", "0"))
sizes = [(100, 1024), (168, 1024), (300, 512), (500, 512), (1024, 512), (2048, 128), (4096, 64), (8192, 16)]
if len(sys.argv) > 1:
    sizes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]]
for N, W in sizes:
    cfg = workloads.make_config("C5", W=W, N=N)
    prob = engine.LikelihoodProblem(cfg["t"], cfg["y"], cfg["yerr"], cfg["kernel"])
    if path_opt:
        prob.set_option("path", path_opt)
    for panel, impl in (((0, 0),) if N > 168 else ((1, 0),)):
        prob.set_option("panel_tiles", panel)
        for _ in range(3):      # eager call, graph capture, first replay (small-N host-buffer calls are replayed as a CUDA graph)
            prob.logl(cfg["hp"], cfg["modpar"])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 10 if N <= 1024 else 3
        for _ in range(reps):
            got, st = prob.logl(cfg["hp"], cfg["modpar"])
        dt = (time.perf_counter() - t0) / reps
        fl = workloads.flops_per_eval(N) * W
        prob.set_option("profile", 1)
        prob.logl(cfg["hp"], cfg["modpar"])
        cats = ("gen", "potrf", "trsm", "update_panel", "update_mid", "update_trail", "mean", "final", "fused")
        ns = {c: prob.info("prof_ns_" + c) for c in cats}
        nl = {c: prob.info("prof_n_" + c) for c in cats}
        prob.set_option("profile", 0)
        tot = max(1, sum(ns.values()))
        print("    kernel time shares: " + "  ".join("%s=%.1f%%(%d)" % (c, 100.0 * ns[c] / tot, nl[c]) for c in cats if nl[c]) + "  sum=%.3f ms" % (tot * 1e-6))
        print("N=%5d W=%4d panel=%d impl=%d path=%d wave=%d: %.3f ms/batch  %.1f evals/s  %.2f TF/s  status_ok=%s"
              % (N, W, panel, impl, prob.info("path"), prob.info("wave_walkers"), dt * 1e3, W / dt, fl / dt / 1e12, bool(np.all(st == 0))), flush=True)
    prob.close()
