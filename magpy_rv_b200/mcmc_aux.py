"""MCMC helpers -- host mirror of the reference's ``magpy_rv.mcmc_aux``.

Same names/signatures as ``/root/reference/src/magpy_rv/mcmc_aux.py`` (numb_param_per_model :27,
get_model :52, initial_pos_creator :101, star_cross :152, parameter_check :187).  ``get_model``
evaluates the curves on the GPU; the walker initialisation and the validity check are host logic
that must consume the NumPy global RNG in exactly the reference's order.
"""
import numpy as np

from . import auxiliary as aux
from . import engine
from . import models as mod


def numb_param_per_model(model_name):
    """Number of parameters of one model, by name prefix (reference mcmc_aux.py:27-49)."""
    n = None
    if model_name.startswith(("no", "No")):
        n = mod.No_Model.numb_param()
    if model_name.startswith(("off", "Off")):
        n = mod.Offset.numb_param()
    if model_name.startswith(("kep", "Kep")):
        n = mod.Keplerian.numb_param()
    if model_name.startswith(("poly", "Poly")):
        n = mod.Polynomial.numb_param()
    if n is None:
        raise UnboundLocalError("cannot access local variable 'model_param_number' where it is not associated with a value")
    return n


def get_model(model_name, time, model_par, to_ecc=False, flags=None):
    """Sum of the model curves (reference mcmc_aux.py:52-97).

    ``model_par`` is sliced positionally by dictionary order (mcmc_aux.py:76).  With ``to_ecc``
    the ecc/omega entries of every Keplerian are overwritten IN PLACE with the physical values
    converted from Sk, Ck (mcmc_aux.py:84-88) -- callers (the priors) rely on that mutation.
    """
    if isinstance(model_name, str):
        model_name = [model_name]
    keys = list(model_par.keys())
    values = [model_par[k].value for k in keys]
    ops, nm = engine.model_ops(model_name)
    # reference semantics: each model sees the dict slice [i, i+n); a short dict fails the
    # per-model length asserts (models.py:335, 416, 573)
    assert len(values) >= nm, "model parameter dictionary is too short for the model list"
    # the key names inside each slice must resolve the way the reference's constructors do
    n_kep, n_off = 0, 0
    fixed_ops = []
    for (tp, off, fv), name in zip(ops, model_name):
        sl = keys[off:off + engine.MODEL_NPAR[tp]]
        if tp == engine.MODEL_OFFSET:
            fv = 1.0 if sl[0] == 'offset' else float(int(sl[0].rsplit('_', 1)[1]) + 1)
        fixed_ops.append((tp, off, fv))
    y, phys = engine.model_eval_ops(fixed_ops, nm, np.asarray(time, dtype=np.float64),
                                    np.array(values[:nm], dtype=np.float64), to_ecc=to_ecc, flags=flags)
    if to_ecc:
        for (tp, off, _fv) in fixed_ops:
            if tp == engine.MODEL_KEPLER:
                model_par[keys[off + 2]].value = float(phys[0, off + 2])
                model_par[keys[off + 3]].value = float(phys[0, off + 3])
    return y[0]


def initial_pos_creator(param, param_err, numb_chains, allow_neg=False, param_names=None):
    """Starting positions [1, numb_chains, n]: walker 0 is the guess itself, the others are
    ``guess + err * U(-1, 1)`` redrawn (whole vector) while any entry is negative
    (reference mcmc_aux.py:101-147).  One ``np.random.uniform(-1, 1, (1, n))`` call per draw, in
    the reference's order, so a shared ``np.random.seed`` gives identical ensembles."""
    n = len(param)
    chains_param = np.zeros(shape=(1, numb_chains, n))
    chains_param[0, 0, ] = param
    for l in range(numb_chains - 1):
        pos = param + param_err * np.random.uniform(-1., 1., (1, n))
        if not allow_neg:
            while np.min(pos) < 0:
                if param_names is None:
                    pos = param + param_err * np.random.uniform(-1., 1., (1, n))
                else:
                    for i in range(len(param_names)):
                        if param_names[i].startswith("ecc") or param_names[i].startswith("omega"):
                            pass
                        else:
                            while pos[0][i] < 0:
                                pos[0][i] = param[i] + param_err[i] * np.random.uniform(-1., 1., (1, len(param[i])))
        chains_param[0, l + 1, ] = pos[0]
    return chains_param


def star_cross(Sk, Ck, Rstar, P, Mstar):
    """True when the orbit never crosses the stellar surface (reference mcmc_aux.py:152-184)."""
    ecc_pl = Sk**2 + Ck**2
    Rsun = 6.95508e8
    AU = 149597871e3
    ratio = (Rstar * Rsun / AU) / (((P / 365.23)**2) * Mstar)**(1 / 3)
    return bool(ecc_pl < 1 - ratio)


def parameter_check(parameters, names, Rstar=None, Mstar=None):
    """Physical validity of one proposed model-parameter vector (reference mcmc_aux.py:187-231).

    Only models whose name starts with 'kepl'/'Kepl' are checked (so 'kep' is NOT -- reference
    quirk): P, K, t0 >= 0 and Sk^2 + Ck^2 <= 1.  With both Rstar and Mstar the reference calls a
    non-existent ``aux.orbit_check`` and raises AttributeError; so does this.
    """
    o = 0
    for name in names:
        if name.startswith(('kepl', 'Kepl')):
            if parameters[o] < 0. or parameters[o + 1] < 0. or parameters[o + 4] < 0.:
                return False
            if (parameters[o + 2]**2 + parameters[o + 3]**2) > 1.:
                return False
            if Rstar is not None and Mstar is not None:
                if not aux.orbit_check(parameters[o + 2], parameters[o + 3], Rstar, parameters[o], Mstar):
                    return False
        o += numb_param_per_model(name)
    return True
