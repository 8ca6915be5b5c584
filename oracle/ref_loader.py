"""Import the UNMODIFIED reference package under matplotlib/seaborn stubs.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference lives at /root/reference (read
only, build container only); ``oracle/stage_ref.py`` installs its unmodified wheel under
``oracle/_ref/`` (git-ignored, travels to the GPU box with the snapshot).  ``load()`` returns a
namespace with the reference modules from whichever of the two is present, or raises
``ReferenceUnavailable``.
"""
import importlib
import os
import sys
import types

REF_SRC_CANDIDATES = ("/root/reference/src", os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref"))


class ReferenceUnavailable(RuntimeError):
    pass


def available():
    return any(os.path.isdir(os.path.join(p, "magpy_rv")) for p in REF_SRC_CANDIDATES)


def _stub(name):
    if name in sys.modules:
        return
    try:
        importlib.import_module(name)
    except Exception:
        sys.modules[name] = types.ModuleType(name)


def load():
    """Return a SimpleNamespace(par, ker, mod, gp, mcmc, mcmcx, aux) of reference modules."""
    src = next((p for p in REF_SRC_CANDIDATES if os.path.isdir(os.path.join(p, "magpy_rv"))), None)
    if src is None:
        raise ReferenceUnavailable("reference package not found (expected /root/reference/src or oracle/_ref)")
    # gp_likelihood.py:14 and mcmc.py:19-20 import matplotlib.pyplot / seaborn at module top;
    # neither is installed here and neither is touched unless plotting is requested.
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        _stub(name)
    if "matplotlib" in sys.modules and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if src not in sys.path:
        sys.path.insert(0, src)
    ns = types.SimpleNamespace()
    ns.par = importlib.import_module("magpy_rv.parameters")
    ns.ker = importlib.import_module("magpy_rv.kernels")
    ns.mod = importlib.import_module("magpy_rv.models")
    ns.aux = importlib.import_module("magpy_rv.auxiliary")
    ns.mcmcx = importlib.import_module("magpy_rv.mcmc_aux")
    ns.gp = importlib.import_module("magpy_rv.gp_likelihood")
    ns.mcmc = importlib.import_module("magpy_rv.mcmc")
    return ns
