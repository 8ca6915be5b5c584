"""Auxiliary helpers -- host mirror of the reference's ``magpy_rv.auxiliary``.

Same names and behaviour as ``/root/reference/src/magpy_rv/auxiliary.py`` (printProgressBar :45,
transit_to_periastron :81, periastron_to_transit :109, to_SkCk :137, to_ecc :181, mass_calc :223,
transpose :255, model_param_names :277, hparam_names :349, phasefold :386).  Scalar host helpers;
the per-walker Sk,Ck -> ecc,omega conversion inside the sampling loop runs on the GPU
(csrc/mean_model.cuh), these versions serve user code and ``MCMC.__init__``.
"""
import numpy as np


def printProgressBar(iteration, total, prefix='Progress: ', suffix='Complete', decimals=1, length=100, fill='█', printEnd="\r"):
    """Terminal progress bar (reference auxiliary.py:45-78)."""
    percent = ("{0:." + str(decimals) + "f}").format(100 * (iteration / float(total)))
    filled = int(length * iteration // total)
    bar = fill * filled + '-' * (length - filled)
    print(f'\r{prefix} |{bar}| {percent}% {suffix}', end=printEnd)
    if iteration == total:
        print()


def _transit_E(ecc, omega):
    v_tr = np.pi / 2 - omega
    return 2 * np.arctan(np.sqrt((1 - ecc) / (1 + ecc)) * np.tan(v_tr / 2))


def transit_to_periastron(t_tr, P, ecc, omega):
    """Time of periastron from time of transit (reference auxiliary.py:81-106)."""
    E_tr = _transit_E(ecc, omega)
    return t_tr - (P / (2 * np.pi) * (E_tr - ecc * np.sin(E_tr)))


def periastron_to_transit(t_0, P, ecc, omega):
    """Time of transit from time of periastron (reference auxiliary.py:109-134)."""
    E_tr = _transit_E(ecc, omega)
    return t_0 + (P / (2 * np.pi) * (E_tr - ecc * np.sin(E_tr)))


def to_SkCk(ecc, omega, ecc_err=None, omega_err=None):
    """Sk = sqrt(e) sin(omega), Ck = sqrt(e) cos(omega), with first-order error propagation
    when both errors are given (reference auxiliary.py:137-177)."""
    root = np.sqrt(ecc)
    s, c = np.sin(omega), np.cos(omega)
    Sk = root * s
    Ck = root * c
    if ecc_err is None or omega_err is None:
        return Sk, Ck
    if ecc == 0.:
        return Sk, Ck, ecc_err, ecc_err
    Sk_err = np.sqrt((ecc_err**2 * s**2 / (4 * ecc)) + (omega_err**2 * ecc * c**2))
    Ck_err = np.sqrt((ecc_err**2 * c**2 / (4 * ecc)) + (omega_err**2 * ecc * s**2))
    return Sk, Ck, Sk_err, Ck_err


def to_ecc(Sk, Ck, errSk=None, errCk=None):
    """e = Sk^2 + Ck^2, omega = atan2(Sk, Ck) (pi/2 when e == 0) (reference auxiliary.py:181-219;
    the error formula is the reference's, as coded)."""
    ecc = Sk**2 + Ck**2
    omega = np.pi / 2 if ecc == 0. else np.arctan2(Sk, Ck)
    if errSk is not None and errCk is not None:
        errecc = np.sqrt(errSk**2 * 4 * Sk**2 + errCk**2 * 4 * Ck**2)
        erromega = np.sqrt(errSk**2 * (Ck / (Ck**2 + Sk**2)**2 + errCk**2 * (-Sk / (Ck**2 + Sk**2)**2)))
        return ecc, omega, errecc, erromega
    return ecc, omega


def mass_calc(model_param, Mstar, earth_mass=False):
    """Minimum planet mass (Jupiter masses; Earth masses if ``earth_mass``) from
    ``[P, K, ecc, omega]`` (reference auxiliary.py:223-252).  Vectorised over walkers when the
    entries are arrays; the sampler passes the list in the reference's (quirky) order, see
    mcmc.py / SURVEY.md F9."""
    P, K, ecc, omega = model_param      # four entries, as the reference unpacks them (a shorter list raises ValueError)
    Mpl_sini = 4.9191 * 10**(-3) * K * np.sqrt(1 - ecc**2) * P**(1 / 3) * Mstar**(2 / 3)
    if earth_mass == False:  # noqa: E712  (reference semantics)
        return Mpl_sini
    if earth_mass == True:  # noqa: E712
        return Mpl_sini * 317.9


def transpose(lst):
    return np.array(lst).T.tolist()


def model_param_names(model_list, SkCk=False, plotting=True):
    """Parameter names of a model list (reference auxiliary.py:277-346)."""
    from . import models as mod
    if isinstance(model_list, str):
        model_list = [model_list]
    if not (isinstance(model_list, list) and len(model_list) >= 1):
        raise ValueError("Model must be a string or a list of strings")
    if len(model_list) == 1:
        name = model_list[0]
        out = None
        if name.startswith(("Kep", "kep")):
            out = mod.Keplerian.params(plotting=plotting, SkCk=SkCk)
        if name.startswith(("No_Model", "No", "no")):
            out = mod.No_Model.params(plotting=plotting)
        if name.startswith(("Off", "off")):
            out = mod.Offset.params(plotting=plotting)
        if name.startswith(("Polynomial", "polynomial")):
            out = mod.Polynomial.params(plotting=plotting)
        if out is None:
            raise UnboundLocalError("cannot access local variable 'param_names' where it is not associated with a value")
        return out
    counts = dict(kep=0, no=0, off=0, poly=0)
    out = []
    for name in model_list:
        if name.startswith(("Kep", "kep")):
            out.extend(mod.Keplerian.params(model_num=counts['kep'], plotting=plotting, SkCk=SkCk))
            counts['kep'] += 1
        if name.startswith(("No_Model", "No", "no")):
            out.extend(mod.No_Model.params(model_num=counts['no'], plotting=plotting))
            counts['no'] += 1
        if name.startswith(("Off", "off")):
            out.extend(mod.Offset.params(model_num=counts['off'], plotting=plotting))
            counts['off'] += 1
        if name.startswith(("Poly", "poly")):
            out.extend(mod.Polynomial.params(model_num=counts['poly'], plotting=plotting))
            counts['poly'] += 1
    return out


def hparam_names(kernel_name, plotting=True):
    """Hyper-parameter names of a kernel (reference auxiliary.py:349-383)."""
    from . import kernels as ker
    rules = ((("Cos", "cos"), ker.Cosine),
             (("expsquare", "ExpSquare", "Expsquare", "expSquare"), ker.ExpSquared),
             (("ExpSin", "expsin", "expSin", "Expsin"), ker.ExpSinSquared),
             (("Quas", "quas"), ker.QuasiPer),
             (("Jit", "jit"), ker.JitterQuasiPer),
             (("Matern5", "matern5"), ker.Matern5),
             (("Matern3", "matern3"), ker.Matern3))
    out = None
    for prefixes, cls in rules:
        if kernel_name.startswith(prefixes):
            out = cls.hparams(plotting)
    if out is None:
        raise UnboundLocalError("cannot access local variable 'hparam_names' where it is not associated with a value")
    return out


def phasefold(time, period, t0, zerocentre=True, returnepoch=False):
    """Phase-fold a time array (reference auxiliary.py:386-423)."""
    phase = (time - t0) / period
    epoch = np.floor(phase)
    true_phase = phase - epoch
    if zerocentre:
        true_phase[np.where(true_phase >= 0.5)[0]] -= 1.0
    if returnepoch:
        return true_phase, epoch
    return true_phase
