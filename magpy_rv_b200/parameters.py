"""Parameter / prior plumbing -- host mirror of the reference's ``magpy_rv.parameters``.

Same public names, argument meaning and error behaviour as
``/root/reference/src/magpy_rv/parameters.py`` (par_create :29, pri_create :95, parameter :178,
PRIORS :240, Gaussian :282, Jeffrey :339, Modified_Jeffrey :411, Uniform :488).  Pure host code:
these objects only carry values; the batched prior evaluation itself runs on the GPU inside
``mrv_logl_batch`` (csrc/mean_model.cuh, ``prior_logprob``).  The scalar ``logprob`` methods below
exist because users call them directly; they restate the same closed forms.
"""
import abc
import math

import numpy as np

ABC = abc.ABC

# canonical kernel name -> hyper-parameter names in dictionary order (reference par_create :42-61)
_KERNEL_PREFIXES = (
    (("Cos", "cos"), ("gp_amp", "gp_per")),
    (("ExpSquare", "expsquare", "Expsquare", "expSquare"), ("gp_amp", "gp_timescale")),
    (("Periodic", "periodic", "ExpSin", "expsin", "Expsin"), ("gp_amp", "gp_timescale", "gp_per")),
    (("QuasiPer", "quasiper", "Quasiper"), ("gp_per", "gp_perlength", "gp_explength", "gp_amp")),
    (("Jit", "jit"), ("gp_per", "gp_perlength", "gp_explength", "gp_amp", "gp_jit")),
    (("Matern5", "matern5", "Matern 5", "matern 5"), ("gp_amp", "gp_timescale")),
    (("Matern3", "matern3", "Matern 3", "matern 3"), ("gp_amp", "gp_timescale", "gp_jit")),
)


def par_create(kernel):
    """Dictionary of the hyper-parameters a kernel needs, keys in the kernel's canonical order.

    Prefix matching as in the reference (parameters.py:42-61): every matching prefix rule is
    applied in turn, the last one wins; an unknown name raises ``UnboundLocalError`` there and
    here.
    """
    found = None
    for prefixes, names in _KERNEL_PREFIXES:
        if kernel.startswith(prefixes):
            found = names
    if found is None:
        raise UnboundLocalError("cannot access local variable 'hparams' where it is not associated with a value")
    return {n: n for n in found}


def PRINTPRIORDER(pri_name=None):
    """Print the order of the values each prior expects (reference parameters.py:66-91)."""
    if pri_name is None:
        pri_name = "no prior"
    if pri_name.startswith(("Gaus", "gaus")):
        print("List should take the form [mu, sigma] where all values are floats or ints")
    elif pri_name.startswith(("Jef", "jef")):
        print("List should take the form [minval, maxval] where all values are float or ints")
    elif pri_name.startswith(("Mod", "mod")):
        print("List should take the form [minval, maxval, kneeval] where all values are floats or ints")
    elif pri_name.startswith(("Uni", "uni")):
        print("List should take the form [minval, maxval] where all values are floats or ints")
    else:
        print("Gaussian: List should take the form [mu, sigma] where all values are floats or ints \n"
              "Jeffery: List should take the form [minval, maxval] where all values are floats or ints \n"
              "Modified Jeffery: List should take the form [minval, maxval, kneeval] where all values are floats or ints \n"
              "Uniform: List should take the form [minval, maxval] where all values are floats or ints")


PRIORS = {
    "Gaussian": ['hparam', 'mu', 'sigma'],
    "Jeffrey": ['hparam', 'minval', 'maxval'],
    "Modified_Jeffrey": ['hparam', 'minval', 'maxval', 'kneeval'],
    "Uniform": ['hparam', 'minval', 'maxval'],
}


def PrintPriorList():
    print("Implemented priors:")
    print(PRIORS)


def defPriorList():
    return PRIORS


def pri_create(param_name, prior, vals=None):
    """Return the ``(param_name, prior_name, prior_dict)`` tuple a prior_list is made of.

    Mirrors reference parameters.py:95-169: prefix dispatch (several rules may fire, the last
    wins), default values of zero, length asserts, ``minval <= maxval`` asserts.
    """
    if type(vals) == np.ndarray:
        vals = vals.tolist()
    out = None
    canon = prior
    if prior.startswith(("Gauss", "gauss")):
        canon = "Gaussian"
        v = [0., 0.] if vals is None else vals
        assert len(v) == 2, "Gaussian priors require a list with 2 values, mu and sigma"
        out = (param_name, canon, dict(mu=v[0], sigma=v[1]))
    if prior.startswith(("Jeffrey", "jeffrey")):
        canon = "Jeffrey"
        v = [0., 0.] if vals is None else vals
        assert len(v) == 2, "Jeffrey priors require a list with 2 values, minval and maxval"
        out = (param_name, canon, dict(minval=v[0], maxval=v[1]))
        assert v[0] <= v[1], 'Minval must be less than maxval.'
    if prior.startswith(("Mod", "mod")):
        canon = "Modified_Jeffrey"
        v = [0., 0., 0.] if vals is None else vals
        assert len(v) == 3, "Modified Jeffrey priors require a list with 3 values, minval, maxval and kneeval"
        out = (param_name, canon, dict(minval=v[0], maxval=v[1], kneeval=v[2]))
        assert v[0] <= v[1], 'Minval must be less than maxval.'
    if prior.startswith(("Uni", "uni")):
        canon = "Uniform"
        v = [0., 0.] if vals is None else vals
        assert len(v) == 2, "Uniform priors require a list with 2 values, minval and maxval"
        out = (param_name, canon, dict(minval=v[0], maxval=v[1]))
        assert v[0] <= v[1], 'Minval must be less than maxval.'
    assert canon in PRIORS.keys(), 'prior not yet implemented. Pick from available priors: ' + str(PRIORS.keys())
    return out


class parameter:
    """Value / error / vary holder (reference parameters.py:178-233).

    As in the reference the instance attributes shadow the same-named accessor methods, and a
    missing error defaults to 20 % of the value.
    """

    def __init__(self, value=None, error=None, vary=True):
        self.value = value
        self.error = error
        self.vary = vary
        if self.error is None:
            self.error = 0.2 * self.value

    def __repr__(self):
        return ("Parameter object: value = {}, error={} (vary = {}) \n").format(self.value, self.error, self.vary)


class Prior(ABC):
    """Parent class of the priors (reference parameters.py:253-269)."""

    @abc.abstractmethod
    def logprob(self, x):
        pass


class Gaussian(Prior):
    """ln N(x; mu, sigma) (reference parameters.py:282-329)."""

    def __init__(self, hparam, mu, sigma):
        self.hparam = hparam
        self.mu = float(mu)
        self.sigma = float(sigma)

    @property
    def __repr__(self):
        message = ("Gaussian prior on the parameter {}, with mu = {} and sigma = {}").format(self.hparam, self.mu, self.sigma)
        print(message)
        return message

    def logprob(self, x):
        return -0.5 * ((x - self.mu) / self.sigma) ** 2 - np.log(np.sqrt(2 * np.pi) * self.sigma)


class _Bounded(Prior):
    _label = ""

    def __init__(self, hparam, minval, maxval):
        self.hparam = hparam
        self.minval = minval
        self.maxval = maxval
        assert self.minval < self.maxval, "Minimum value {} must be smaller than the maximum value {}".format(self.minval, self.maxval)

    def _outside(self, x):
        return x < self.minval or x > self.maxval

    @property
    def __repr__(self):
        message = ("{} prior on the paramter {}, with minimum and maximum values of x = ({}, {})").format(self._label, self.hparam, self.minval, self.maxval)
        print(message)
        return message


class Jeffrey(_Bounded):
    """p(x) ~ 1/x on [minval, maxval]; -1e6 outside (reference parameters.py:339-402)."""
    _label = "Jeffrey"

    def logprob(self, x):
        if self._outside(x):
            return -10e5
        return np.log(1. / (np.log(self.maxval / self.minval)) * 1. / x)


class Modified_Jeffrey(_Bounded):
    """p(x) ~ 1/(x - knee) on [minval, maxval]; -1e6 outside (reference parameters.py:411-479)."""
    _label = "Modified Jeffrey"

    def __init__(self, hparam, minval, maxval, kneeval):
        super().__init__(hparam, minval, maxval)
        self.kneeval = kneeval

    @property
    def __repr__(self):
        message = ("Modified Jeffrey prior on the paramter {}, with minimum and maximum values of x = ({}, {}) and knee = {}").format(self.hparam, self.minval, self.maxval, self.kneeval)
        print(message)
        return message

    def logprob(self, x):
        if self._outside(x):
            return -10e5
        return np.log(1. / (np.log((self.maxval + self.kneeval) / (self.kneeval))) * 1. / (x - self.kneeval))


class Uniform(_Bounded):
    """0 inside [minval, maxval], -1e6 outside (reference parameters.py:488-544)."""
    _label = "Uniform"

    def logprob(self, x):
        return -10e5 if self._outside(x) else 0


# ---------------------------------------------------------------------------------------------
# flattening for the C ABI (include/magpy_rv_b200.h, mrv_prior)
# ---------------------------------------------------------------------------------------------

PRIOR_GAUSSIAN, PRIOR_JEFFREY, PRIOR_MODJEFFREY, PRIOR_UNIFORM = 0, 1, 2, 3


def prior_code(prior_name, prior_parameters):
    """(type code, p0, p1, p2) for one prior, with the reference's dispatch
    (gp_likelihood.py:316-349: every matching branch runs, the last one wins) and its
    constructor asserts (min < max)."""
    out = None
    if prior_name.startswith(("Gaussian", "gaussian")):
        out = (PRIOR_GAUSSIAN, float(prior_parameters["mu"]), float(prior_parameters["sigma"]), 0.0)
    if prior_name.startswith(("Jeffrey", "jeffrey")):
        lo, hi = prior_parameters["minval"], prior_parameters["maxval"]
        assert lo < hi, "Minimum value {} must be smaller than the maximum value {}".format(lo, hi)
        out = (PRIOR_JEFFREY, float(lo), float(hi), 0.0)
    if prior_name.startswith(("Modified", "modified")):
        lo, hi, knee = prior_parameters["minval"], prior_parameters["maxval"], prior_parameters["kneeval"]
        assert lo < hi, "Minimum value {} must be smaller than the maximum value {}".format(lo, hi)
        out = (PRIOR_MODJEFFREY, float(lo), float(hi), float(knee))
    if prior_name.startswith(("Uni", "uni")):
        lo, hi = prior_parameters["minval"], prior_parameters["maxval"]
        assert hi > lo, "Minimum value {} must be smaller than the maximum value {}".format(lo, hi)
        out = (PRIOR_UNIFORM, float(lo), float(hi), 0.0)
    if out is None:
        raise AssertionError('Prior not yet implemented. Pick from available priors: ' + str(PRIORS.keys()))
    return out
