"""Run one batched LogL of the C5 workload (for ncu): python tools/profile_case.py N W [reps] [config]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads
from magpy_rv_b200 import engine
N, W = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
name = sys.argv[4] if len(sys.argv) > 4 else "C5"
cfg = workloads.make_config(name, W=W, N=N)
prob = engine.LikelihoodProblem(cfg["t"], cfg["y"], cfg["yerr"], cfg["kernel"], cfg["model_list"], cfg["prior_list"],
                                flags=cfg["flags"], kernel_variant=cfg["kernel_variant"])
for kv in os.environ.get("MRV_OPTS", "").split(","):
    if "=" in kv:
        prob.set_option(kv.split("=")[0], int(kv.split("=")[1]))
for _ in range(reps):
    l, s = prob.logl(cfg["hp"], cfg["modpar"])
print("ok", l[:2], int((s != 0).sum()), "launches", prob.info("launches"))
