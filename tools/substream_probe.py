"""Blocked path with one stream vs two half-waves on two streams (option "substreams"): median of 9 batches."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, workloads
from magpy_rv_b200 import engine
sizes = [tuple(int(x) for x in a.split('x')) for a in sys.argv[1:]] or [(640, 512), (1024, 128), (1024, 512), (2048, 128), (4096, 64), (4096, 128), (8192, 32)]
for N, W in sizes:
    cfg = workloads.make_config("C5", W=W, N=N)
    prob = engine.LikelihoodProblem(cfg["t"], cfg["y"], cfg["yerr"], cfg["kernel"])
    prob.set_option("path", 2)
    for _ in range(3):
        prob.logl(cfg["hp"], cfg["modpar"])
    res = {}
    for rnd in range(3):                      # interleave the modes so drifts hit all of them alike
        for sub in (1, 2, 11, 12):                 # 11 / 12: one / two streams with the diagonal-first look-ahead
            prob.set_option("substreams", sub % 10)
            prob.set_option("lookahead", 2 * (sub // 10))
            prob.logl(cfg["hp"], cfg["modpar"])
            for _ in range(3):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                prob.logl(cfg["hp"], cfg["modpar"])
                res.setdefault(sub, []).append(time.perf_counter() - t0)
    med = {k: np.median(v) for k, v in res.items()}
    print("N=%d W=%d wave=%d: " % (N, W, prob.info("wave_walkers")) + "  ".join(
        "%d stream(s) %.2f ms %.2f TF/s (%+.1f %%)" % (k, med[k] * 1e3, workloads.flops_per_eval(N) * W / med[k] / 1e12,
                                                       100 * (med[1] / med[k] - 1)) for k in sorted(med)), flush=True)
    prob.close()
