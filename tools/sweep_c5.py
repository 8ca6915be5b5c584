"""The C5 scaling sweep on one GPU: evaluations per second and algorithmic TFLOP/s for N = 1024 .. 16384 at the
per-GPU ensemble of the 8-GPU run (128 walkers) -- and, with an argument, at any other ensemble size (N limited so that a
batch stays below ~10 s) --, medians of 5 batches through mrv_logl_batch.    python tools/sweep_c5.py [walkers]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, workloads
from magpy_rv_b200 import engine, _native
peak = _native.lib().mrv_fp64_mma_peak(0, 20000)
print("DMMA peak %.2f TFLOP/s" % (peak / 1e12))
W = int(sys.argv[1]) if len(sys.argv) > 1 else 128
for N in (512, 1024, 2048, 4096, 8192, 16384):
    if workloads.flops_per_eval(N) * W / 3.0e13 > 10.0:
        continue
    cfg = workloads.make_config("C5", W=W, N=N)
    prob = engine.LikelihoodProblem(cfg["t"], cfg["y"], cfg["yerr"], cfg["kernel"])
    for _ in range(2):
        prob.logl(cfg["hp"], cfg["modpar"])
    ts = []
    for _ in range(5 if N < 16384 else 2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        prob.logl(cfg["hp"], cfg["modpar"])
        ts.append(time.perf_counter() - t0)
    dt = float(np.median(ts))
    fl = workloads.flops_per_eval(N) * W / dt
    print("N=%5d W=%d: %.2f ms per batch, %.1f evals/s, %.2f TFLOP/s = %.1f %% of the DMMA peak (wave %d)"
          % (N, W, dt * 1e3, W / dt, fl / 1e12, 100 * fl / peak, prob.info("wave_walkers")), flush=True)
    prob.close()
