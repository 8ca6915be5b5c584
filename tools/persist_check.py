import sys; sys.path.insert(0, "/root/repo")
import numpy as np, workloads
from magpy_rv_b200 import engine
for N, W in ((700, 9), (2048, 20), (4096, 8)):
    cfg = workloads.make_config("C5", W=W, N=N)
    prob = engine.LikelihoodProblem(cfg["t"], cfg["y"], cfg["yerr"], cfg["kernel"])
    a, sa = prob.logl(cfg["hp"], cfg["modpar"])
    prob.set_option("gemm_impl", 1)
    b, sb = prob.logl(cfg["hp"], cfg["modpar"])
    print(N, W, "bitwise equal:", np.array_equal(a, b), "status", sa.any(), sb.any(), flush=True)
    prob.close()
