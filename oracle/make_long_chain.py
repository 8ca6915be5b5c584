"""Long seeded reference chains as DIGESTS (test infrastructure only): the unmodified reference's run_MCMC for
many more iterations than tests/golden/chain_*.npz hold, stored as the SHA-256 of the position histories plus the
last log-likelihood row, so that a single flipped accept / reject anywhere in thousands of decisions is detected
without committing megabytes.    python oracle/make_long_chain.py   ->  tests/golden/chain_long.json"""
import contextlib, hashlib, io, json, os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import workloads
from oracle import ref_loader

ref = ref_loader.load()
out = {}
for name, W, iters, N in (("C1", 100, 400, None), ("C3", 40, 60, 90)):
    cfg = workloads.make_config(name, W=W, N=N)
    hp0 = ref.par.par_create(cfg["kernel"])
    for k in hp0:
        s = cfg["hparam0"][k]
        hp0[k] = ref.par.parameter(value=s.value, error=s.error, vary=s.vary)
    mp0 = ref.mod.mod_create(cfg["model_list"])
    for k in mp0:
        s = cfg["model_par0"][k]
        mp0[k] = ref.par.parameter(value=s.value, error=s.error, vary=s.vary)
    priors = [ref.par.pri_create(a, b, list(d.values())) for a, b, d in cfg["prior_list"]]
    random.seed(cfg["seed"]); np.random.seed(cfg["seed"])
    with contextlib.redirect_stdout(io.StringIO()):
        logL, hps, mps, done = ref.mcmc.run_MCMC(iters, cfg["t"], cfg["y"], cfg["yerr"], hp0, cfg["kernel"], mp0, cfg["model_list"],
                                                 priors, numb_chains=W, flags=cfg["flags"])
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(hps).tobytes()); h.update(np.ascontiguousarray(mps).tobytes())
    out[name] = {"W": W, "iterations": iters, "N": cfg["N"], "seed": cfg["seed"], "positions_sha256": h.hexdigest(),
                 "last_logL": logL[-1, :, 0].tolist(), "accepted_steps": int((np.any(hps[1:] != hps[:-1], axis=2)).sum())}
    print(name, out[name]["positions_sha256"][:16], out[name]["accepted_steps"], flush=True)
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "chain_long.json"), "w"))
