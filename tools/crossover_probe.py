"""Streamed-panel kernel (path 3) against the tiled path (path 2, automatic variant) for 144 < N <= 512.

    python tools/crossover_probe.py [N ...]      # walkers: 100 148 256 512 1024

Evaluations per second through mrv_logl_batch (median of 7 calls after 3 warm-up calls)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, workloads
from magpy_rv_b200 import engine

# MRV_PATHS=1,3: fused shared-memory kernel against the streamed-panel kernel (N <= 144 ... 168)
PA, PBp = (int(x) for x in os.environ.get("MRV_PATHS", "2,3").split(","))
Ns = [int(a) for a in sys.argv[1:]] or [256, 288, 320, 352, 384, 416, 448, 480, 512]
Ws = [int(x) for x in os.environ.get("MRV_WS", "100,148,256,512,1024").split(",")]
print("evals/s: path %d | path %d | chosen automatically   (1 fused, 2 tiled, 3 streamed panel)" % (PA, PBp))
for N in Ns:
    row = []
    for W in Ws:
        cfg = workloads.make_config("C5", W=W, N=N)
        r = {}
        for path in (PA, PBp, 0):
            prob = engine.LikelihoodProblem(cfg["t"], cfg["y"], cfg["yerr"], cfg["kernel"])
            prob.set_option("path", path)
            for _ in range(3):
                prob.logl(cfg["hp"], cfg["modpar"])
            ts = []
            for _ in range(7):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                prob.logl(cfg["hp"], cfg["modpar"])
                ts.append(time.perf_counter() - t0)
            r[path] = (W / float(np.median(ts)), prob.info("path"))
            prob.close()
        row.append("W=%4d: %7.0fk | %7.0fk | %d:%7.0fk %s" % (W, r[PA][0] / 1e3, r[PBp][0] / 1e3, r[0][1], r[0][0] / 1e3,
                                                          "" if r[0][0] >= 0.97 * max(r[PA][0], r[PBp][0]) else "<-- wrong"))
    print("N=%4d  " % N + "   ".join(row), flush=True)
