import sys, os, time
sys.path.insert(0, "/root/repo")
import numpy as np, workloads
from magpy_rv_b200 import engine
cfg = workloads.make_config("C5", W=32, N=16384)
prob = engine.LikelihoodProblem(cfg["t"], cfg["y"], cfg["yerr"], cfg["kernel"])
prob.logl(cfg["hp"], cfg["modpar"])
for band, cs in ((1000, 1), (48, 1), (64, 1), (80, 1), (96, 1), (1000, 1)):
    prob.set_option("update_band", band); prob.set_option("cs_store", cs)
    prob.set_option("profile", 1)
    prob.logl(cfg["hp"], cfg["modpar"])
    ns = prob.info("prof_ns_update_trail"); alg = prob.info("prof_alg_update_trail"); tot = sum(prob.info("prof_ns_" + c) for c in ("gen","potrf","trsm","update_panel","update_trail","mean","final"))
    prob.set_option("profile", 0)
    print("band=%4d cs=%d trailing %.1f ms %.2f TF/s  total kernels %.1f ms" % (band, cs, ns*1e-6, alg/ns/1e3, tot*1e-6), flush=True)
