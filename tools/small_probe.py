"""Fused shared-memory kernel at small N: resident CTAs per SM / register cap variants (option "small_minb")."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, workloads
from magpy_rv_b200 import engine
for N in (32, 48, 64, 80, 96, 112):
    for W in (100, 1024):
        cfg = workloads.make_config("C5", W=W, N=N)
        prob = engine.LikelihoodProblem(cfg["t"], cfg["y"], cfg["yerr"], cfg["kernel"])
        res = {}
        for rnd in range(2):
            for mb in (2, 3, 4, 0):
                prob.set_option("small_minb", mb)
                for _ in range(2): prob.logl(cfg["hp"], cfg["modpar"])
                for _ in range(5):
                    torch.cuda.synchronize(); t0 = time.perf_counter(); prob.logl(cfg["hp"], cfg["modpar"]); res.setdefault(mb, []).append(time.perf_counter() - t0)
        print("N=%d W=%d: " % (N, W) + "  ".join("minb %d: %.3f ms (%.2f M/s)" % (mb, np.median(v) * 1e3, W / np.median(v) / 1e6) for mb, v in res.items()), flush=True)
        prob.close()
