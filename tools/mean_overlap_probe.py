"""C3-style workload (Keplerian mean model) on the tiled path: Newton kernel on its own stream next to the
generation of tile column 0 (option mean_overlap = 1) against everything on one stream (0).  Interleaved medians.
   python tools/mean_overlap_probe.py [N W] ..."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import workloads
from magpy_rv_b200 import engine

cases = [(500, 256), (500, 512), (700, 256), (1024, 128)]
if len(sys.argv) > 2:
    cases = [(int(sys.argv[i]), int(sys.argv[i + 1])) for i in range(1, len(sys.argv) - 1, 2)]
for N, W in cases:
    cfg = workloads.make_config("C3", W=W, N=N)
    prob = engine.LikelihoodProblem(cfg["t"], cfg["y"], cfg["yerr"], cfg["kernel"], cfg["model_list"], cfg["prior_list"],
                                    flags=cfg["flags"])
    prob.set_option("path", 2)
    ref = None
    times = {0: [], 1: []}
    for rep in range(9):
        for opt in (0, 1):
            prob.set_option("mean_overlap", opt)
            l, s = prob.logl(cfg["hp"], cfg["modpar"])
            t0 = time.perf_counter()
            for _ in range(10):
                l, s = prob.logl(cfg["hp"], cfg["modpar"])
            times[opt].append((time.perf_counter() - t0) / 10)
            if ref is None:
                ref = l
            assert np.array_equal(l, ref), "results must not depend on the stream layout"
    print("N=%d W=%d: one stream %.3f ms, overlapped %.3f ms" % (N, W, 1e3 * np.median(times[0]), 1e3 * np.median(times[1])),
          flush=True)
    prob.close()
