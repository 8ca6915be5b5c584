"""Randomised cross-check of the three GPU paths against each other (no CPU oracle: hundreds of cases in
seconds): random N, ensemble size, kernel, model list and priors; every path that accepts N must agree with the
automatic choice to 1e-9 and report the same status words.   python tools/stress_paths.py [cases] [seed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from magpy_rv_b200 import engine
from test_gpu_fuzz import random_case

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
worst, n_checked = 0.0, 0
for k in range(cases):
    N = int(rng.choice([rng.integers(1, 145), rng.integers(145, 513), rng.integers(145, 513), rng.integers(513, 900)]))
    c = random_case(rng, N)
    W = int(rng.choice([1, 2, 7, 33, 149, 300, 420]))
    reps = (W + c["W"] - 1) // c["W"]
    hp = np.tile(c["hp"], (reps, 1))[:W] * (1 + 1e-3 * rng.standard_normal((W, c["hp"].shape[1])))
    mp = np.tile(c["modpar"], (reps, 1))[:W]
    prob = engine.LikelihoodProblem(c["t"], c["y"], c["yerr"], c["kernel"], c["model_list"], c["priors"], flags=c["flags"])
    base, st0 = prob.logl(hp, mp)
    auto = prob.info("path")
    for path, n_max in ((1, prob.info("small_n_max")), (2, 1 << 30), (3, prob.info("mid_n_max"))):
        if N > n_max or path == auto:
            continue
        prob.set_option("path", path)
        got, st = prob.logl(hp, mp)
        assert np.array_equal(st == 0, st0 == 0), (k, N, W, c["kernel"], c["model_list"], path, st[:8], st0[:8])
        ok = (st0 == 0) & np.isfinite(base)
        assert np.array_equal(np.isfinite(got[ok]), np.isfinite(base[ok]))
        if ok.any():
            err = float(np.max(np.abs(got[ok] - base[ok]) / np.abs(base[ok])))
            worst = max(worst, err)
            assert err < 1e-9, (k, N, W, c["kernel"], c["model_list"], auto, path, err)
        n_checked += 1
    prob.close()
print("stress ok: %d cases, %d cross-checks, worst relative difference %.2e" % (cases, n_checked, worst))
