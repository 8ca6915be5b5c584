"""Time the UNMODIFIED reference's run_MCMC (src/magpy_rv/mcmc.py:676) end to end for the C1..C3
workloads in the BUILD CONTAINER (the reference tree does not travel to the GPU box): wall time per
iteration and log-likelihood evaluations per second including all of its Python overhead
(SURVEY.md section 8(d)).  TEST / MEASUREMENT INFRASTRUCTURE ONLY.

    python oracle/time_reference_mcmc.py            # writes profiles/reference_run_mcmc_container.json
"""
import contextlib, io, json, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import workloads
from oracle import ref_loader

ref = ref_loader.load()
out = {"host": {"cores": len(os.sched_getaffinity(0)), "where": "build container (not the GPU box)"}, "runs": []}
for name, iters in (("C1", 10), ("C2", 4), ("C3", 2)):
    cfg = workloads.make_config(name)
    hp = ref.par.par_create(cfg["kernel"])
    for k in hp:
        hp[k] = ref.par.parameter(value=cfg["hparam0"][k].value, error=cfg["hparam0"][k].error, vary=cfg["hparam0"][k].vary)
    mp = ref.mod.mod_create(cfg["model_list"])
    for k in mp:
        mp[k] = ref.par.parameter(value=cfg["model_par0"][k].value, error=cfg["model_par0"][k].error, vary=cfg["model_par0"][k].vary)
    priors = [ref.par.pri_create(a, b, list(c.values())) for a, b, c in cfg["prior_list"]]
    random.seed(cfg["seed"]); np.random.seed(cfg["seed"])
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        ref.mcmc.run_MCMC(iters, cfg["t"], cfg["y"], cfg["yerr"], hp, cfg["kernel"], mp, cfg["model_list"], priors,
                          numb_chains=cfg["W"], plot_convergence=False)
    dt = time.perf_counter() - t0
    evals = (iters + 1) * cfg["W"]
    rec = {"config": name, "N": cfg["N"], "walkers": cfg["W"], "iterations": iters, "seconds": dt,
           "ms_per_iteration": 1e3 * dt / (iters + 1), "evals_per_s": evals / dt}
    print(rec, flush=True)
    out["runs"].append(rec)
json.dump(out, open(os.path.join(ROOT, "profiles", "reference_run_mcmc_container.json"), "w"), indent=1)
