"""Phase breakdown of the persistent dataflow kernel (csrc/dag.cuh).  Needs a build with the phase timers:

    MRV_NVCC_EXTRA=-DDAG_TIMING python -c "from magpy_rv_b200 import build; build.build(force=True)"
    python tools/dag_timing.py 1024x128 2048x128 C3:500x256

Prints, per case, SM-microseconds per task for every phase (thread 0 of each CTA, clock64 at 1.965 GHz) and the share
of the kernel's total SM time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, workloads
from magpy_rv_b200 import engine

NAMES = ["fetch+dep wait", "generation", "off-diag update loop", "C store", "wait diag tile", "solve main loop", "solve epilogue",
         "diag: acc->S", "sweep pivot blocks", "sweep panel rows", "sweep trailing DMMA", "pivots/logdet/z", "inv(L) write-back",
         "#diag tasks", "#off-diag tasks", "diag update loop"]
GHZ = 1.965
cases = sys.argv[1:] or ["1024x128", "2048x128", "C3:500x256"]
for a in cases:
    name, _, rest = a.rpartition(":")
    n, w = (int(x) for x in rest.split("x"))
    cfg = workloads.make_config(name or "C5", W=w, N=n)
    prob = engine.LikelihoodProblem(cfg["t"], cfg["y"], cfg["yerr"], cfg["kernel"], cfg["model_list"], cfg["prior_list"],
                                    flags=cfg["flags"], kernel_variant=cfg["kernel_variant"])
    prob.set_option("path", 2)
    prob.set_option("tiled_impl", 2)
    prob.logl(cfg["hp"], cfg["modpar"])
    prob.info("dag_t16")
    reps = 3
    for _ in range(reps):
        prob.logl(cfg["hp"], cfg["modpar"])
    v = [prob.info("dag_t%d" % i) for i in range(16)]
    if v[13] <= 0:
        print("no timing data: build with MRV_NVCC_EXTRA=-DDAG_TIMING")
        break
    nd, no = v[13] / reps, v[14] / reps
    tot = sum(v[i] for i in range(16) if i not in (13, 14))
    print("%s N=%d W=%d: %d diag + %d off-diag tasks per batch, total %.1f SM-ms per batch (= %.3f ms on 148 SMs)"
          % (name or "C5", n, w, nd, no, tot / reps / GHZ * 1e-6, tot / reps / GHZ * 1e-6 / 148))
    for i in range(16):
        if i in (13, 14):
            continue
        per = {1: nd + no, 2: no, 0: nd + no, 15: nd, 3: no, 4: no, 5: no, 6: no}.get(i, nd)
        print("   %-22s %6.1f %%   %8.2f us per task" % (NAMES[i], 100.0 * v[i] / tot, v[i] / reps / GHZ * 1e-3 / max(per, 1)))
    prob.close()
