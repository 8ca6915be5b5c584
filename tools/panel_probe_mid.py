import sys, os, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch, workloads
from magpy_rv_b200 import engine
for N, W in ((1024, 512), (1024, 128), (2048, 128), (4096, 64)):
    cfg = workloads.make_config("C5", W=W, N=N)
    prob = engine.LikelihoodProblem(cfg["t"], cfg["y"], cfg["yerr"], cfg["kernel"])
    prob.logl(cfg["hp"], cfg["modpar"])
    for panel in (2, 3, 4, 6, 8, 16, 32):
        if panel > N // 128: continue
        prob.set_option("panel_tiles", panel)
        prob.logl(cfg["hp"], cfg["modpar"])
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3): prob.logl(cfg["hp"], cfg["modpar"])
        dt = (time.perf_counter() - t0) / 3
        print("N=%d W=%d panel=%d: %.2f ms  %.2f TF/s" % (N, W, panel, dt * 1e3, workloads.flops_per_eval(N) * W / dt / 1e12), flush=True)
    prob.close()
