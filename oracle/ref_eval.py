"""Drive the UNMODIFIED reference (oracle/ref_loader.py) on a workloads.make_config dictionary.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/__init__.py): used by oracle/make_golden.py, tests/ and the
``cpu_baseline`` / ``--impl reference`` legs of bench.py -- never by the product.

``logl_walker``  one walker exactly as the reference's serial loop does it (mcmc.py:421-458: par_create /
                 mod_create / parameter objects, get_model(to_ecc=True), GPLikelihood(...).LogL(prior_list)).
``run_mcmc``     the reference's run_MCMC (mcmc.py:676) on the configuration, seeded, stdout swallowed.
"""
import contextlib
import io
import random
import time

import numpy as np


def _ref_params(ref, cfg):
    hp0 = ref.par.par_create(cfg["kernel"])
    for k in hp0:
        src = cfg["hparam0"][k]
        hp0[k] = ref.par.parameter(value=src.value, error=src.error, vary=src.vary)
    mp0 = ref.mod.mod_create(list(cfg["model_list"]))
    for k in mp0:
        src = cfg["model_par0"][k]
        mp0[k] = ref.par.parameter(value=src.value, error=src.error, vary=src.vary)
    prior_list = [ref.par.pri_create(a, b, [d[x] for x in d]) for a, b, d in cfg["prior_list"]]
    return hp0, mp0, prior_list


def logl_walker(ref, cfg, w, prior_list=None):
    """LogL of walker ``w`` of ``cfg`` through the reference classes (mcmc.py:421-458)."""
    hp = ref.par.par_create(cfg["kernel"])
    for i, k in enumerate(hp):
        hp[k] = ref.par.parameter(value=float(cfg["hp"][w, i]), error=1., vary=True)
    mp = ref.mod.mod_create(list(cfg["model_list"]))
    for i, k in enumerate(mp):
        mp[k] = ref.par.parameter(value=float(cfg["modpar"][w, i]), error=1., vary=True)
    if prior_list is None:
        prior_list = [ref.par.pri_create(a, b, [d[x] for x in d]) for a, b, d in cfg["prior_list"]]
    if cfg["flags"] is None:
        my = ref.mcmcx.get_model(list(cfg["model_list"]), cfg["t"], mp, flags=None, to_ecc=True)
    else:
        my = ref.mcmcx.get_model(list(cfg["model_list"]), cfg["t"], mp, flags=cfg["flags"], to_ecc=True)
    like = ref.gp.GPLikelihood(cfg["t"], cfg["y"], cfg["yerr"], hp, cfg["kernel"], my, mp)
    return like.LogL(list(prior_list))


def time_logl(ref, cfg, walkers):
    """(seconds per evaluation, values) over the given walker indices."""
    prior_list = [ref.par.pri_create(a, b, [d[x] for x in d]) for a, b, d in cfg["prior_list"]]
    vals = []
    t0 = time.perf_counter()
    for w in walkers:
        vals.append(logl_walker(ref, cfg, w, prior_list))
    return (time.perf_counter() - t0) / max(1, len(walkers)), vals


def run_mcmc(ref, cfg, iterations):
    """(evaluations, seconds, final mean logL) of the reference's run_MCMC on ``cfg`` (cfg["W"] chains)."""
    hp0, mp0, prior_list = _ref_params(ref, cfg)
    random.seed(cfg["seed"])
    np.random.seed(cfg["seed"])
    kw = {} if cfg["flags"] is None else {"flags": cfg["flags"]}
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        out = ref.mcmc.run_MCMC(iterations, cfg["t"], cfg["y"], cfg["yerr"], hp0, cfg["kernel"], mp0, list(cfg["model_list"]),
                                prior_list, numb_chains=cfg["W"], plot_convergence=False, **kw)
    dt = time.perf_counter() - t0
    return (iterations + 1) * cfg["W"], dt, float(np.mean(out[0][-1]))
