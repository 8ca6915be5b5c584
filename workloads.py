"""Synthetic RV workloads of the five BASELINE.json configurations.

Shared by tests/, bench.py and __graft_entry__.smoke().  Data come from
``np.random.default_rng(seed)``; walker positions come from the package's own
``initial_pos_creator`` (RNG-order identical to the reference's, mcmc_aux.py:101-147) under
``np.random.seed(seed)``, so the GPU path, the oracle and the reference all see the same numbers.

The "GP draw" component of the reference plan is replaced by a cheap quasi-periodic surrogate
(amplitude-modulated sinusoids): drawing an exact GP sample needs an N x N Cholesky on the host,
which at N = 16384 would dominate the benchmark set-up.  The likelihood cost does not depend on y.
"""
import numpy as np

from magpy_rv_b200 import auxiliary as aux
from magpy_rv_b200 import mcmc_aux
from magpy_rv_b200 import models as mod
from magpy_rv_b200 import parameters as par

BASE_SEED = 20231013


def _activity(rng, t, per, amp):
    """Quasi-periodic surrogate signal: three slowly modulated harmonics."""
    span = max(float(t[-1] - t[0]), 1.0)
    y = np.zeros_like(t)
    for h, a in ((1, 1.0), (2, 0.45), (3, 0.2)):
        ph = rng.uniform(0, 2 * np.pi)
        mod_ph = rng.uniform(0, 2 * np.pi)
        y += a * np.sin(2 * np.pi * h * t / per + ph) * (1.0 + 0.3 * np.sin(2 * np.pi * t / (span / 3.0) + mod_ph))
    return amp * y / 1.2


def _kepler_rv(t, P, K, ecc, omega, t0):
    """Plain Keplerian RV curve for data synthesis (per-point Newton to 1e-13)."""
    M = np.mod(2 * np.pi * (t - t0) / P, 2 * np.pi)
    E = M.copy()
    for _ in range(60):
        dE = (E - ecc * np.sin(E) - M) / (1 - ecc * np.cos(E))
        E -= dE
        if np.max(np.abs(dE)) < 1e-13:
            break
    nu = 2 * np.arctan2(np.sqrt(1 + ecc) * np.sin(E / 2), np.sqrt(1 - ecc) * np.cos(E / 2))
    return K * (np.cos(omega + nu) + ecc * np.cos(omega))


def _hparams(kernel, vals_errs):
    hp = par.par_create(kernel)
    for key, (v, e) in zip(hp.keys(), vals_errs):
        hp[key] = par.parameter(value=v, error=e, vary=True)
    return hp


def _walkers(seed, W, hparam0, model_par0, model_list):
    """[W, nh], [W, nm] starting positions exactly as MCMC.__init__ builds them (hyper-parameters
    first, then model parameters; Keplerian ecc/omega converted to Sk/Ck with propagated errors)."""
    hv = [hparam0[k].value for k in hparam0]
    he = [hparam0[k].error for k in hparam0]
    names = list(model_par0.keys())
    mv = [model_par0[k].value for k in names]
    me = [model_par0[k].error for k in names]
    kinds = []
    for m in model_list:
        kinds += [m] * mcmc_aux.numb_param_per_model(m)
    for g, n in enumerate(names):
        if n.startswith("ecc") and kinds[g].startswith(("kep", "Kep")):
            mv[g], mv[g + 1], me[g], me[g + 1] = aux.to_SkCk(mv[g], mv[g + 1], me[g], me[g + 1])
    state = np.random.get_state()
    np.random.seed(seed)
    hp = mcmc_aux.initial_pos_creator(hv, he, W)[0]
    mp = mcmc_aux.initial_pos_creator(mv, me, W)[0]
    np.random.set_state(state)
    return np.ascontiguousarray(hp), np.ascontiguousarray(mp)


def make_config(name, W=None, N=None, seed=None):
    """Return a dict describing one configuration: data (t, y, yerr, flags), kernel name, model list,
    parameter dictionaries (hparam0, model_par0), prior_list and walker ensembles hp / modpar."""
    idx = {"C1": 1, "C2": 2, "C3": 3, "C4": 4, "C4M": 4, "C5": 5}[name]
    seed = BASE_SEED + idx if seed is None else seed
    rng = np.random.default_rng(seed)
    flags = None
    model_list = ["no_model"]
    model_par0 = mod.mod_create(["no_model"])
    model_par0["no"] = par.parameter(value=0., error=0., vary=False)
    prior_list = []
    kernel_variant = 0

    if name == "C1":
        N = 100 if N is None else N
        W = 100 if W is None else W
        t = np.sort(rng.uniform(0, 100, N))
        y = 8 * np.cos(2 * np.pi * t / 20) + rng.normal(0, 1.5, N)
        kernel = "QuasiPer"
        hparam0 = _hparams(kernel, [(20., 2.), (0.5, 0.1), (40., 5.), (8., 1.)])
        for key, lo, hi in (("gp_per", 0., 50.), ("gp_perlength", 0., 10.), ("gp_explength", 0., 200.), ("gp_amp", 0., 50.)):
            prior_list.append(par.pri_create(key, "Uniform", [lo, hi]))
    elif name == "C2":
        N = 300 if N is None else N
        W = 100 if W is None else W
        t = np.sort(rng.uniform(0, 300, N))
        y = 8 * np.cos(2 * np.pi * t / 20) + rng.normal(0, 1.5, N) + 5 + 0.01 * t + 2e-5 * t**2
        kernel = "QuasiPer"
        hparam0 = _hparams(kernel, [(20., 2.), (0.5, 0.1), (40., 5.), (8., 1.)])
        model_list = ["Polynomial"]
        model_par0 = mod.mod_create(model_list)
        model_par0["a0"] = par.parameter(value=5., error=1., vary=True)
        model_par0["a1"] = par.parameter(value=0.01, error=0.005, vary=True)
        model_par0["a2"] = par.parameter(value=2e-5, error=1e-5, vary=True)
        model_par0["a3"] = par.parameter(value=0., error=1e-9, vary=False)
        for key, lo, hi in (("gp_per", 0., 50.), ("gp_perlength", 0., 10.), ("gp_explength", 0., 200.), ("gp_amp", 0., 50.),
                            ("a0", -10., 20.), ("a1", -1., 1.)):
            prior_list.append(par.pri_create(key, "Uniform", [lo, hi]))
    elif name == "C3":
        N = 500 if N is None else N
        W = 256 if W is None else W
        t = np.sort(rng.uniform(0, 330, N)) + 2500002.
        y = _kepler_rv(t, 4.23, 55., 0.01, 1.0, 2450001.) + _activity(rng, t, 40., 5.) + rng.normal(0, 3., N)
        kernel = "JitterQuasiPer"
        hparam0 = _hparams(kernel, [(40., 10.), (0.5, 0.2), (30., 10.), (5., 5.), (1., 0.5)])
        model_list = ["Keplerian"]
        model_par0 = mod.mod_create(model_list)
        model_par0["P"] = par.parameter(value=4.2, error=1., vary=True)
        model_par0["K"] = par.parameter(value=55., error=5., vary=True)
        model_par0["ecc"] = par.parameter(value=0.013, error=0.1, vary=True)
        model_par0["omega"] = par.parameter(value=1., error=0.1, vary=True)
        model_par0["t0"] = par.parameter(value=2450001., error=10., vary=True)
        for key, lo, hi in (("gp_amp", 0., 20.), ("gp_explength", 0., 60.), ("gp_per", 0., 80.), ("gp_perlength", 0., 10.),
                            ("gp_jit", 0., 10.), ("P", 0., 20.), ("K", 0., 60.), ("ecc", 0., 80.), ("omega", 0., 10.),
                            ("t0", 2450000., 2450010.)):
            prior_list.append(par.pri_create(key, "Uniform", [lo, hi]))
    elif name in ("C4", "C4M"):
        N = 4096 if N is None else N
        W = 512 if W is None else W
        n1, n2 = N // 2, (3 * N) // 8
        n3 = N - n1 - n2
        ts = [np.sort(rng.uniform(0, 1500, n)) for n in (n1, n2, n3)]
        offs = (0.0, 2.0, 1.5)
        ys, es = [], []
        for tt, o in zip(ts, offs):
            ys.append(_activity(rng, tt, 27., 3.) + o + rng.normal(0, 1., len(tt)))
            es.append(rng.uniform(0.5, 1.5, len(tt)))
        t, y, yerr, flags = mod.combine_data(ts, ys, es)
        model_list = ["Offset", "Offset"]
        model_par0 = mod.mod_create(model_list)
        model_par0["offset_0"] = par.parameter(value=2., error=0.5, vary=True)
        model_par0["offset_1"] = par.parameter(value=1.5, error=0.5, vary=True)
        if name == "C4":
            kernel = "QuasiPer"
            hparam0 = _hparams(kernel, [(27., 2.), (0.5, 0.1), (60., 10.), (3., 0.5)])
            prior_list.append(par.pri_create("gp_per", "Uniform", [0., 100.]))
            prior_list.append(par.pri_create("gp_amp", "Jeffrey", [0.01, 100.]))
        else:  # opt-in textbook Matern 5/2 (the reference's as-coded kernel is not PD)
            kernel = "Matern5/2"
            kernel_variant = 1
            hparam0 = _hparams(kernel, [(3., 0.5), (20., 5.)])
            prior_list.append(par.pri_create("gp_amp", "Jeffrey", [0.01, 100.]))
        prior_list.append(par.pri_create("offset_0", "Gaussian", [2., 1.]))
    elif name == "C5":
        N = 4096 if N is None else N
        W = 1024 if W is None else W
        t = np.sort(rng.uniform(0, N, N))
        y = _activity(rng, t, 25., 5.) + rng.normal(0, 1., N)
        kernel = "QuasiPer"
        hparam0 = _hparams(kernel, [(25., 2.), (0.5, 0.1), (50., 5.), (5., 0.5)])
    else:
        raise KeyError(name)

    if name not in ("C4", "C4M"):
        yerr = rng.uniform(0.5, 1.5, N)
    hp, modpar = _walkers(seed, W, hparam0, model_par0, model_list)
    return dict(name=name, N=int(N), W=int(W), seed=seed, t=np.ascontiguousarray(t, dtype=np.float64),
                y=np.ascontiguousarray(y, dtype=np.float64), yerr=np.ascontiguousarray(yerr, dtype=np.float64),
                flags=None if flags is None else np.ascontiguousarray(flags, dtype=np.float64), kernel=kernel,
                kernel_variant=kernel_variant, model_list=model_list, hparam0=hparam0, model_par0=model_par0,
                prior_list=prior_list, hp=hp, modpar=modpar)


def flops_per_eval(N):
    """Algorithmic FLOPs of one log-likelihood: Cholesky N^3/3 + two triangular
    solves 2 N^2.  The kernel evaluations N(N+1)/2 are reported separately."""
    return N**3 / 3.0 + 2.0 * N**2
