"""Randomised check of the dataflow kernel (tiled_impl=2) against the launch-per-column loop (tiled_impl=1): random N
(3 ... 16 tile columns), ensemble size, kernel, mean models and priors (tests/test_gpu_fuzz.random_case).  log L and
status words must be BITWISE equal between the two, and equal over repeated runs of the dataflow kernel (its task
order differs from run to run, so a race or an order-dependent sum would show).   python tools/dag_stress.py [cases] [seed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from magpy_rv_b200 import engine
from test_gpu_fuzz import random_case

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
n_bad_status = 0
for k in range(cases):
    N = int(rng.choice([rng.integers(257, 513), rng.integers(513, 1100), rng.integers(513, 1100), rng.integers(1100, 2049)]))
    c = random_case(rng, N)
    W = int(rng.choice([1, 3, 37, 128, 149, 300]))
    if N > 1100:
        W = min(W, 40)
    reps = (W + c["W"] - 1) // c["W"]
    hp = np.tile(c["hp"], (reps, 1))[:W] * (1 + 1e-3 * rng.standard_normal((W, c["hp"].shape[1])))
    if rng.random() < 0.2:
        hp[rng.integers(0, W)] *= -1.0           # a walker whose covariance is not positive definite
    mp = np.tile(c["modpar"], (reps, 1))[:W]
    prob = engine.LikelihoodProblem(c["t"], c["y"], c["yerr"], c["kernel"], c["model_list"], c["priors"], flags=c["flags"])
    prob.set_option("path", 2)
    prob.set_option("tiled_impl", 1)
    want, st_want = prob.logl(hp, mp)
    prob.set_option("tiled_impl", 2)
    for r in range(3):
        got, st = prob.logl(hp, mp)
        assert np.array_equal(st, st_want), (k, N, W, c["kernel"], c["model_list"], r, st[:8], st_want[:8])
        assert np.array_equal(got, want, equal_nan=True), (k, N, W, c["kernel"], c["model_list"], r,
                                                           float(np.nanmax(np.abs(got - want))))
    n_bad_status += int((st_want != 0).sum())
    prob.close()
print("dag stress ok: %d cases bitwise equal to the launch-per-column path over 3 runs each (%d walkers with a non-zero status among them)"
      % (cases, n_bad_status))
