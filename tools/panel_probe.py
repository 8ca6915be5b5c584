"""Whole-batch kernel time vs panel width (CUDA-event profile counters)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads
from magpy_rv_b200 import engine
for N, W in ((16384, 32), (8192, 64), (4096, 128), (2048, 256)):
    cfg = workloads.make_config("C5", W=W, N=N)
    prob = engine.LikelihoodProblem(cfg["t"], cfg["y"], cfg["yerr"], cfg["kernel"])
    prob.logl(cfg["hp"], cfg["modpar"])
    for pw, pwi in ((8, 8), (16, 8), (24, 8), (24, 6), (24, 4), (32, 8), (48, 8), (24, 12), (48, 12)):
        if pw > N // 128:
            continue
        prob.set_option("panel_tiles", pw)
        prob.set_option("panel_inner", pwi)
        prob.set_option("profile", 1)
        prob.logl(cfg["hp"], cfg["modpar"])
        cats = ("gen", "potrf", "trsm", "update_panel", "update_mid", "update_trail", "mean", "final")
        ns = {c: prob.info("prof_ns_" + c) for c in cats}
        prob.set_option("profile", 0)
        tot = sum(ns.values()) * 1e-6
        print("N=%5d W=%3d panel=%2d/%2d total %.1f ms  %.2f TF/s  (trail %.1f mid %.1f panel %.1f trsm %.1f potrf %.1f)"
              % (N, W, pw, pwi, tot, workloads.flops_per_eval(N) * W / tot / 1e9, ns["update_trail"] * 1e-6, ns["update_mid"] * 1e-6, ns["update_panel"] * 1e-6,
                 ns["trsm"] * 1e-6, ns["potrf"] * 1e-6), flush=True)
    prob.close()
