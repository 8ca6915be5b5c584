"""Fused shared-memory kernels at small N: rank-1 sweep vs blocked DMMA sweep (option "small_blk_min")."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, workloads
from magpy_rv_b200 import engine
for N in (16, 32, 48, 64, 72, 79):
    for W in (100, 1024):
        cfg = workloads.make_config("C5", W=W, N=N)
        prob = engine.LikelihoodProblem(cfg["t"], cfg["y"], cfg["yerr"], cfg["kernel"])
        out = []
        for mn in (80, 1):
            prob.set_option("small_blk_min", mn)
            for _ in range(3): prob.logl(cfg["hp"], cfg["modpar"])
            ts = []
            for _ in range(9):
                torch.cuda.synchronize(); t0 = time.perf_counter(); prob.logl(cfg["hp"], cfg["modpar"]); ts.append(time.perf_counter() - t0)
            out.append(np.median(ts))
        print("N=%d W=%d: rank-1 %.3f ms (%.2f M evals/s)   blocked sweep %.3f ms (%.2f M evals/s)" % (N, W, out[0] * 1e3, W / out[0] / 1e6, out[1] * 1e3, W / out[1] / 1e6), flush=True)
        prob.close()
