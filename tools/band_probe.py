"""Trailing-update rate vs rasterisation band height and panel width (CUDA-event profile counters)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads
from magpy_rv_b200 import engine
cfg = workloads.make_config("C5", W=32, N=16384)
prob = engine.LikelihoodProblem(cfg["t"], cfg["y"], cfg["yerr"], cfg["kernel"])
prob.logl(cfg["hp"], cfg["modpar"])
for pw, band in ((24, 64), (24, 32), (24, 24), (24, 16), (16, 32), (16, 24), (12, 40), (8, 64), (24, 1000)):
    prob.set_option("panel_tiles", pw)
    prob.set_option("update_band", band)
    prob.set_option("profile", 1)
    prob.logl(cfg["hp"], cfg["modpar"])
    cats = ("gen", "potrf", "trsm", "update_panel", "update_mid", "update_trail", "mean", "final")
    ns = {c: prob.info("prof_ns_" + c) for c in cats}
    alg = prob.info("prof_alg_update_trail")
    prob.set_option("profile", 0)
    tot = sum(ns.values()) * 1e-6
    print("panel=%2d band=%4d trailing %.1f ms %.2f TF/s | total %.1f ms %.2f TF/s"
          % (pw, band, ns["update_trail"] * 1e-6, alg / ns["update_trail"] / 1e3, tot, workloads.flops_per_eval(16384) * 32 / tot / 1e9), flush=True)
