"""Phase timers of the streamed-panel kernel (build with MRV_NVCC_EXTRA=-DMID_TIMING): python tools/mid_timing.py N W"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads
from magpy_rv_b200 import engine
N, W = int(sys.argv[1]), int(sys.argv[2])
cfg = workloads.make_config("C5", W=W, N=N)
prob = engine.LikelihoodProblem(cfg["t"], cfg["y"], cfg["yerr"], cfg["kernel"])
prob.set_option("path", 3)
prob.logl(cfg["hp"], cfg["modpar"])
prob.info("mid_t32")
prob.logl(cfg["hp"], cfg["modpar"])
names = ["prologue", "mma loop", "sync", "panel->smem+sync", "factorA | genA", "sync", "factorB | genB", "trsmB+fence+sync", "trsmA", "sync", "updB", "sync", "sync", "loop: issue", "loop: wait full", "loop: lds+mma+arrive"]
walkers_cta0 = (W + min(W, 148 * (2 if N <= 256 else 1)) - 1) // min(W, 148 * (2 if N <= 256 else 1))
for base, who in ((0, "warp 0"), (16, "warp 1")):
    v = [prob.info("mid_t%d" % (base + i)) for i in range(16)]
    tot = sum(v)
    print("%s: total %.1f us per walker (CTA 0 ran %d walkers)" % (who, tot / 1.9e3 / walkers_cta0, walkers_cta0))
    print("   " + "  ".join("%s %.1f%%" % (n, 100.0 * x / max(1, tot)) for n, x in zip(names, v)))
