"""Tiled path: launch-per-column loop (tiled_impl=1) against the persistent dataflow kernel (tiled_impl=2).
Evaluations per second (median of 5 batches through mrv_logl_batch), fraction of the DMMA peak, bitwise comparison of
the two and, for the launch-per-column path, the time share of every kernel category (CUDA events).

    python tools/dag_probe.py [NxW ...]        # default: the C5 sweep at 128 and 1024 walkers, C3
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, workloads
from magpy_rv_b200 import engine, _native

peak = _native.lib().mrv_fp64_mma_peak(0, 20000)
print("DMMA peak %.2f TFLOP/s" % (peak / 1e12), flush=True)
cases = [("C5", 640, 128), ("C5", 1024, 128), ("C5", 1024, 1024), ("C5", 2048, 128), ("C5", 2048, 1024), ("C5", 4096, 128),
         ("C5", 4096, 512), ("C3", 500, 256), ("C5", 8192, 32)]
if len(sys.argv) > 1:
    cases = []
    for a in sys.argv[1:]:
        name, _, rest = a.rpartition(":")
        n, w = rest.split("x")
        cases.append((name or "C5", int(n), int(w)))
cats = ("gen", "potrf", "trsm", "update_panel", "update_mid", "update_trail", "mean", "final", "dag")
for name, N, W in cases:
    cfg = workloads.make_config(name, W=W, N=N)
    res = {}
    for impl in (1, 2):
        prob = engine.LikelihoodProblem(cfg["t"], cfg["y"], cfg["yerr"], cfg["kernel"], cfg["model_list"], cfg["prior_list"],
                                        flags=cfg["flags"], kernel_variant=cfg["kernel_variant"])
        prob.set_option("path", 2)
        prob.set_option("tiled_impl", impl)
        for _ in range(2):
            got, st = prob.logl(cfg["hp"], cfg["modpar"])
        ts = []
        for _ in range(5 if N * W < 4096 * 512 else 3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            got, st = prob.logl(cfg["hp"], cfg["modpar"])
            ts.append(time.perf_counter() - t0)
        dt = float(np.median(ts))
        prob.set_option("profile", 1)
        prob.logl(cfg["hp"], cfg["modpar"])
        ns = {c: prob.info("prof_ns_" + c) for c in cats}
        nl = {c: prob.info("prof_n_" + c) for c in cats}
        prob.set_option("profile", 0)
        tot = max(1, sum(v for v in ns.values() if v > 0))
        shares = "  ".join("%s=%.1f%%(%d)" % (c, 100.0 * ns[c] / tot, nl[c]) for c in cats if nl[c] > 0)
        fl = workloads.flops_per_eval(N) * W / dt
        res[impl] = (got.copy(), st.copy(), dt)
        print("%s N=%5d W=%4d impl=%d: %8.3f ms/batch %10.1f evals/s %6.2f TFLOP/s = %5.1f %% of DMMA peak (wave %d, status ok %s) | kernel-time sum %.3f ms: %s"
              % (name, N, W, impl, dt * 1e3, W / dt, fl / 1e12, 100 * fl / peak, prob.info("wave_walkers"), bool(np.all(st == 0)),
                 tot * 1e-6, shares), flush=True)
        prob.close()
    same = np.array_equal(res[1][0], res[2][0]) and np.array_equal(res[1][1], res[2][1])
    rel = float(np.max(np.abs(res[1][0] - res[2][0]) / np.abs(res[1][0])))
    print("    dataflow vs launch-per-column: bitwise identical = %s (max rel diff %.2e), speed-up %.2fx"
          % (same, rel, res[1][2] / res[2][2]), flush=True)
