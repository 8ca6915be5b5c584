"""Trailing-update kernel rate for the GEMM implementation variants (CUDA-event profile counters)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads
from magpy_rv_b200 import engine
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
W = int(sys.argv[2]) if len(sys.argv) > 2 else 32
cfg = workloads.make_config("C5", W=W, N=N)
prob = engine.LikelihoodProblem(cfg["t"], cfg["y"], cfg["yerr"], cfg["kernel"])
prob.logl(cfg["hp"], cfg["modpar"])
for impl, cs in ((0, 1), (1, 1), (1, 3), (1, 1), (1, 3)):
    prob.set_option("gemm_impl", impl)
    prob.set_option("cs_store", cs)
    prob.set_option("profile", 1)
    prob.logl(cfg["hp"], cfg["modpar"])
    cats = ("gen", "potrf", "trsm", "update_panel", "update_mid", "update_trail", "mean", "final")
    ns = {c: prob.info("prof_ns_" + c) for c in cats}
    alg = {c: prob.info("prof_alg_" + c) for c in cats}
    prob.set_option("profile", 0)
    print("cs=%d " % cs + "impl=%d trailing %.1f ms %.2f TF/s | panel %.1f ms %.2f TF/s | trsm %.1f ms | potrf %.1f ms | total %.1f ms"
          % (impl, ns["update_trail"] * 1e-6, alg["update_trail"] / max(1, ns["update_trail"]) / 1e3, ns["update_panel"] * 1e-6,
             alg["update_panel"] / max(1, ns["update_panel"]) / 1e3, ns["trsm"] * 1e-6, ns["potrf"] * 1e-6, sum(ns.values()) * 1e-6), flush=True)
