"""Phase timers of the fused small-N kernel (CTA 0, thread 0; build with -DSMALL_TIMING):
    python tools/build_variant.py small_timing -DSMALL_TIMING
    MRV_LIB=magpy_rv_b200/_lib/variants/small_timing.so python tools/small_timing.py C1:100x100 100x1024"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads
from magpy_rv_b200 import engine
names = ["prologue + mean + residual", "fill + phase table", "covariance generation", "sweep", "log det / priors / output"]
for a in sys.argv[1:] or ["C1:100x100", "100x1024"]:
    name, _, rest = a.rpartition(":")
    n, w = (int(x) for x in rest.split("x"))
    cfg = workloads.make_config(name or "C5", W=w, N=n)
    prob = engine.LikelihoodProblem(cfg["t"], cfg["y"], cfg["yerr"], cfg["kernel"], cfg["model_list"], cfg["prior_list"], flags=cfg["flags"])
    prob.set_option("path", 1)
    prob.set_option("cuda_graph", 0)
    prob.logl(cfg["hp"], cfg["modpar"])
    prob.info("small_t8")
    reps = 5
    for _ in range(reps):
        prob.logl(cfg["hp"], cfg["modpar"])
    v = [prob.info("small_t%d" % i) / reps / 1965.0 for i in range(5)]
    print("%s: CTA 0 %.1f us: " % (a, sum(v)) + "  ".join("%s %.1f" % (nm, x) for nm, x in zip(names, v)))
    prob.close()
