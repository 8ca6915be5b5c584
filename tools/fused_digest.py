"""Digest of log L / status over random small-N configurations on the fused path (tests/test_gpu_fuzz.random_case), three
runs each -- to compare two builds of the library bit for bit (MRV_LIB) and a build with itself across runs:
    python tools/fused_digest.py > a.txt;  MRV_LIB=... python tools/fused_digest.py > b.txt;  diff a.txt b.txt"""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from magpy_rv_b200 import engine
from test_gpu_fuzz import random_case

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 5)
for k in range(int(sys.argv[2]) if len(sys.argv) > 2 else 120):
    N = int(rng.integers(17, 145))
    c = random_case(rng, N)
    W = int(rng.choice([1, 5, 100, 149, 300, 700]))
    reps = (W + c["W"] - 1) // c["W"]
    hp = np.tile(c["hp"], (reps, 1))[:W] * (1 + 1e-3 * rng.standard_normal((W, c["hp"].shape[1])))
    if k % 4 == 0:
        hp[W // 2] *= -1.0
    mp = np.tile(c["modpar"], (reps, 1))[:W]
    prob = engine.LikelihoodProblem(c["t"], c["y"], c["yerr"], c["kernel"], c["model_list"], c["priors"], flags=c["flags"])
    prob.set_option("path", 1)
    prob.set_option("small_minb", int(rng.choice([0, 0, 2, 3, 4])))
    ds = []
    for r in range(3):
        got, st = prob.logl(hp, mp)
        ds.append(hashlib.sha256(got.tobytes() + st.tobytes()).hexdigest()[:16])
    print(k, N, W, c["kernel"], "+".join(c["model_list"]), ds[0], "stable" if len(set(ds)) == 1 else "UNSTABLE %s" % ds, int((st != 0).sum()))
    prob.close()
