"""Accuracy vs condition number: GPU LogL against the oracle for increasingly ill-conditioned K."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from magpy_rv_b200 import engine
from oracle import logl_port as port
rng = np.random.default_rng(7)
for N in (120, 600, 1500):
    t = np.sort(rng.uniform(0, 0.9 * N, N))
    y = 4 * np.sin(t / 7.0) + rng.normal(0, 0.5, N)
    for kern, hp in (("ExpSquared", [4.0, 9.0]), ("QuasiPer", [20., .5, 40., 6.])):
        for sig in (1.0, 0.1, 0.01, 0.003):
            e = np.full(N, sig)
            K = port.compute_kernel(kern, hp, t, t, t, e)
            cond = np.linalg.cond(K)
            want = port.logL_one(t, y, e, kern, hp)
            for path in (1, 2):
                prob = engine.LikelihoodProblem(t, y, e, kern)
                if path == 1 and N > prob.info("small_n_max"):
                    prob.close(); continue
                prob.set_option("path", path)
                got, st = prob.logl([hp], [[0.0]])
                prob.close()
                print("N=%4d %-10s sigma=%.3f cond=%.1e path=%d status=%d rel err=%.2e" % (N, kern, sig, cond, path, st[0], abs(got[0] - want) / abs(want)), flush=True)
