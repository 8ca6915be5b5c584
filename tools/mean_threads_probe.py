"""Threads per walker of the Keplerian mean kernel (option mean_threads, probe only) on C3-style workloads:
measured flat within +-2 % from 256 to 512 threads (1 / 2 points per thread), worse at 128 and above 768 --
the Newton loop is bound by the FP64 instruction count of sincos + two divisions, not by the reduction chain.
   python tools/mean_threads_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import workloads
from magpy_rv_b200 import engine
for N, W in [(500, 256), (300, 256), (700, 256)]:
    cfg = workloads.make_config("C3", W=W, N=N)
    prob = engine.LikelihoodProblem(cfg["t"], cfg["y"], cfg["yerr"], cfg["kernel"], cfg["model_list"], cfg["prior_list"], flags=cfg["flags"])
    prob.set_option("mean_overlap", 0)
    ref = None
    res = {}
    for rep in range(7):
        for mt in (0, 128, 256, 384, 512, 768, 1024):
            prob.set_option("mean_threads", mt)
            l, s = prob.logl(cfg["hp"], cfg["modpar"])
            t0 = time.perf_counter()
            for _ in range(10):
                l, s = prob.logl(cfg["hp"], cfg["modpar"])
            res.setdefault(mt, []).append((time.perf_counter() - t0) / 10)
            if ref is None: ref = l
            d = np.max(np.abs(l - ref) / np.abs(ref))
            assert d < 1e-11, d
    print(N, W, prob.info("path"), {k: round(1e3 * float(np.median(v)), 3) for k, v in res.items()}, flush=True)
    prob.close()
