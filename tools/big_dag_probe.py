"""Dataflow kernel against the launch-per-column loop at N = 8192 / 16384 (operand streaming from HBM: 16 flop per byte
at every N) with different wave sizes:   python tools/big_dag_probe.py [N W ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, workloads
from magpy_rv_b200 import engine, _native

peak = _native.lib().mrv_fp64_mma_peak(0, 20000)
args = [int(a) for a in sys.argv[1:]] or [16384, 64]
for N, W in zip(args[0::2], args[1::2]):
    cfg = workloads.make_config("C5", W=W, N=N)
    ref = None
    for impl, wave in ((1, 0), (2, 32), (2, 64), (2, 0)):
        if wave > W:
            continue
        prob = engine.LikelihoodProblem(cfg["t"], cfg["y"], cfg["yerr"], cfg["kernel"])
        prob.set_option("path", 2)
        prob.set_option("tiled_impl", impl)
        if wave:
            prob.set_option("wave_walkers", wave)
        got, st = prob.logl(cfg["hp"], cfg["modpar"])
        ts = []
        for _ in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            got, st = prob.logl(cfg["hp"], cfg["modpar"])
            ts.append(time.perf_counter() - t0)
        dt = min(ts)
        fl = workloads.flops_per_eval(N) * W / dt
        if ref is None:
            ref = got.copy()
        print("N=%d W=%d impl=%d wave=%d: %.1f ms per batch, %.2f evals/s, %.1f %% of the DMMA peak, bitwise equal to impl 1: %s, status ok %s"
              % (N, W, impl, prob.info("wave_walkers"), dt * 1e3, W / dt, 100 * fl / peak, bool(np.array_equal(got, ref)), bool(np.all(st == 0))), flush=True)
        prob.close()
