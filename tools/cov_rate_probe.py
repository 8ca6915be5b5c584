"""Covariance evaluations per second: library sinpi / exp on distances (cov_from_dist_n) against the per-point fast
form (cov_pair_n: angle-difference identity + exp_np), and the accuracy of the fast form against numpy."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from magpy_rv_b200 import _native as nat

lib = nat.lib()
dfma = lib.mrv_fp64_fma_peak(0, 20000)
hps = {0: [3.0, 17.0], 1: [4.0, 9.0], 2: [4.0, 1.3, 19.0], 3: [21.0, 0.6, 35.0, 6.0], 4: [21.0, 0.6, 35.0, 6.0, 1.2], 5: [5.0, 12.0], 6: [5.0, 12.0, 0.8]}
for kid, hp in hps.items():
    h = np.array(hp)
    old = lib.mrv_cov_eval_rate(0, kid, 100, nat.dptr(h), len(hp))
    new = lib.mrv_cov_eval_rate(0, kid, 0, nat.dptr(h), len(hp))
    ls = lib.mrv_cov_eval_rate(0, kid, 200, nat.dptr(h), len(hp))
    print("kernel %d: library form %.3e evals/s (%.0f flop-equivalents per entry), per-point form %.3e evals/s (%.0f): %.2fx, "
          "per-point form with the lock-step exp %.3e evals/s (%.0f)" % (kid, old, dfma / old, new, dfma / new, new / old, ls, dfma / ls), flush=True)
