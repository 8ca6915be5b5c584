"""Mean models -- host mirror of the reference's ``magpy_rv.models``.

Same names and signatures as ``/root/reference/src/magpy_rv/models.py`` (mod_create :46,
combine_data :114, No_Model :212, Polynomial :275, Offset :380, Keplerian :488).  The curves are
evaluated on the GPU (csrc/mean_model.cuh through ``mrv_model_eval``): Polynomial
``a0 + a1 t + a2 t^2 + a3 t^3``, Offset selected by ``flags``, Keplerian with the reference's
whole-vector Newton schedule (models.py:618-650).  No CPU fallback.
"""
import abc

import numpy as np

from . import engine
from . import parameters as par

ABC = abc.ABC

MODELS = {"No_Model": ['rvs'],
          "Offset": ['rvs', 'offset'],
          "Polynomial": ["a0", "a1", "a2", "a3"],
          "Keplerian": ['time', 'P', 'K', 'ecc', 'omega', 't0'],
          }


def PrintModelList():
    print("Implemented models")
    print(MODELS)


def defModelList():
    return MODELS


_KEP_DESC = ('period', 'semi-amplitude', 'eccentricity', 'angle of periastron')


def mod_create(model):
    """Dictionary of the parameters the chosen model(s) need, keys in evaluation order
    (reference models.py:46-106: bare names for a single model, ``name_j`` for several, j
    counting models of the same kind)."""
    if isinstance(model, str) or (isinstance(model, list) and len(model) == 1):
        single = True
    elif isinstance(model, list) and len(model) > 1:
        single = False
    else:
        raise ValueError("Model must be a string or a list of strings")
    if single:
        # as in the reference, a bare string is indexed at [0] (its first character)
        first = model[0]
        out = None
        if first.startswith(("Kep", "kep")):
            out = dict(P='period', K='semi-amplitude', ecc='eccentricity', omega='angle of periastron', t0='t of per pass')
        if first.startswith(("No_Model", "No", "no")):
            out = dict(no=par.parameter(value=0., error=0., vary=False))
        if first.startswith(("Off", "off")):
            out = dict(offset='offset')
        if first.startswith(("Poly", "poly")):
            out = dict(a0="a0", a1="a1", a2="a2", a3="a3")
        if out is None:
            raise UnboundLocalError("cannot access local variable 'model_params' where it is not associated with a value")
        return out
    counts = dict(kep=0, no=0, off=0, poly=0)
    out = {}
    for name in model:
        if name.startswith(("Kep", "kep")):
            j = counts['kep']
            out.update({'P_%d' % j: 'period', 'K_%d' % j: 'semi-amplitude', 'ecc_%d' % j: 'eccentricity',
                        'omega_%d' % j: 'angle of periastron', 't0_%d' % j: 't of periastron passage'})
            counts['kep'] += 1
        if name.startswith(("No_Model", "No", "no")):
            out.update({'no_%d' % counts['no']: 'no'})
            counts['no'] += 1
        if name.startswith(("Off", "off")):
            out.update({'offset_%d' % counts['off']: 'offset'})
            counts['off'] += 1
        if name.startswith(("Poly", "poly")):
            j = counts['poly']
            out.update({'a0_%d' % j: 'a0', 'a1_%d' % j: 'a1', 'a2_%d' % j: 'a2', 'a3_%d' % j: 'a3'})
            counts['poly'] += 1
    return out


def combine_data(times, ys, y_errs):
    """Merge several instruments' series into one time-sorted series plus a ``flags`` array
    (0 for the first instrument, 1 for the second, ...) -- reference models.py:114-177.  Host-side
    data preparation (runs once, before any likelihood)."""
    if type(times) == np.ndarray:
        times = times.tolist()
    if type(ys) == np.ndarray:
        ys = ys.tolist()
    if type(y_errs) == np.ndarray:
        y_errs = y_errs.tolist()
    flags = [np.zeros_like(np.asarray(tt)) + k for k, tt in enumerate(times)]
    rows = sorted(zip(np.concatenate(times), np.concatenate(ys), np.concatenate(y_errs), np.concatenate(flags)))
    time, y, y_err, flag = (np.array(col) for col in zip(*rows))
    return (time, y, y_err, flag)


class Model(ABC):
    """Parent class of the mean models (reference models.py:185-207)."""

    @abc.abstractmethod
    def model(self):
        pass


def _numbered(names, model_num, plotting):
    if model_num is None:
        return list(names)
    if plotting is True:
        return [(n + "$_{}$").format(model_num) for n in names]
    return ["{}_{}".format(n, model_num) for n in names]


def _lookup(model_params, names, err):
    """Values for ``names`` from a single-model dict ('P') or a numbered one ('P_j', first j in
    0..9 that exists) -- the try/except search of reference models.py:339-357, 419-433, 581-600.
    Returns (values, j or None)."""
    try:
        return [model_params[n].value for n in names], None
    except KeyError:
        for j in range(10):
            try:
                return [model_params['%s_%d' % (n, j)].value for n in names], j
            except KeyError:
                continue
    raise KeyError(err)


class No_Model(Model):
    """Null model: zeros (reference models.py:212-267)."""

    def __init__(self, y, no):
        self.y = y

    @staticmethod
    def numb_param():
        return 1

    @staticmethod
    def params(model_num=None, plotting=True):
        return _numbered(["no"], model_num, plotting)

    def model(self):
        out, _ = engine.model_eval(["no"], np.asarray(self.y, dtype=np.float64), [0.0])
        return out[0]


class Polynomial(Model):
    """a3 t^3 + a2 t^2 + a1 t + a0 (reference models.py:275-371)."""

    def __init__(self, time, model_params):
        self.time = time
        self.model_params = model_params
        assert len(self.model_params) == 4, "Offset Model requires 4 parameters"
        (self.a0, self.a1, self.a2, self.a3), _ = _lookup(self.model_params, ("a0", "a1", "a2", "a3"),
                                                          "Offset Model requires 4 parameters")

    @staticmethod
    def numb_param():
        return 4

    @staticmethod
    def params(model_num=None, plotting=True):
        return _numbered(["a0", "a1", "a2", "a3"], model_num, plotting)

    def model(self):
        out, _ = engine.model_eval(["poly"], self.time, [self.a0, self.a1, self.a2, self.a3])
        return out[0]


class Offset(Model):
    """Constant offset applied where ``flags == flag_val`` (reference models.py:380-479)."""

    def __init__(self, flags, model_params):
        self.flags = flags
        self.model_params = model_params
        assert len(self.model_params) == 1, "Offset Model requires 1 parameter:" + "'offset'"
        (self.offset,), j = _lookup(self.model_params, ("offset",), "Offset Model requires 1 parameter:'offset'")
        self.flag_val = 1. if j is None else j + 1.

    @staticmethod
    def numb_param():
        return 1

    @staticmethod
    def params(model_num=None, plotting=True):
        return _numbered(["offset"], model_num, plotting)

    def model(self):
        flags = np.asarray(self.flags, dtype=np.float64)
        # one OFFSET op whose flag value is this instance's
        out, _ = engine.model_eval_ops([(engine.MODEL_OFFSET, 0, float(self.flag_val))], 1,
                                       np.zeros_like(flags), [self.offset], flags=flags)
        return out[0]


class Keplerian(Model):
    """Keplerian RV curve K (cos(omega + nu) + e cos omega) (reference models.py:488-697)."""

    def __init__(self, time, model_params):
        self.time = time
        self.model_params = model_params
        assert len(self.model_params) == 5, "Keplerian Model requires 5 parameters:" + "'P', 'K', 'ecc', 'omega', 't0'"
        (self.P, self.K, self.ecc, self.omega, self.t0), _ = _lookup(
            self.model_params, ("P", "K", "ecc", "omega", "t0"),
            "Keplerian Model requires 5 parameters:'P', 'K', 'ecc', 'omega', 't0'")

    def __repr__(self):
        return "Keplerian RV model with the following set of parameters: \nP = {} \nK = {} \necc = {} \nomega = {} \nt0 = {}".format(
            self.P, self.K, self.ecc, self.omega, self.t0)

    @staticmethod
    def numb_param():
        return 5

    @staticmethod
    def params(model_num=None, plotting=True, SkCk=False):
        names = ["P", "K", "Sk", "Ck", "t0"] if SkCk is True else ["P", "K", "ecc", "omega", "t0"]
        return _numbered(names, model_num, plotting)

    def ecc_anomaly(self, M, ecc, max_itr=200):
        """Eccentric anomaly by Newton iteration on the whole vector M, stopping when ||E - E0||_1 <= 1e-10
        (reference models.py:618-650), on the GPU."""
        E = engine.kepler_anomaly(M, ecc, max_itr)
        return E if np.ndim(M) else E[()]

    def true_anomaly(self, E, ecc):
        """nu = 2 arctan(sqrt((1 + ecc) / (1 - ecc)) tan(E / 2)) (reference models.py:653-669), on the GPU."""
        nu = engine.true_anomaly(E, ecc)
        return nu if np.ndim(E) else nu[()]

    def model(self):
        out, _ = engine.model_eval(["kep"], self.time, [self.P, self.K, self.ecc, self.omega, self.t0], to_ecc=False)
        return out[0]
