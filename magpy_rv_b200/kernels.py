"""Covariance kernels -- host mirror of the reference's ``magpy_rv.kernels``.

Same classes, constructor checks and ``compute_covmatrix(dist_e, dist_se, errors)`` signature as
``/root/reference/src/magpy_rv/kernels.py`` (Cosine :119, ExpSquared :246, ExpSinSquared :374,
QuasiPer :507, JitterQuasiPer :648, Matern5 :799, Matern3 :927).  The N1 x N2 map itself runs on
the GPU (csrc/dense_ops.cuh through ``mrv_covmatrix_from_dist`` / ``mrv_distances``); this file
only resolves names and the reference's diagonal-broadcast rules.  No CPU fallback.
"""
import abc

import numpy as np

from . import engine

ABC = abc.ABC

KERNELS = {
    "Cosine": ['gp_amp', 'gp_per'],
    "ExpSquared": ['gp_amp', 'gp_timescale'],
    "ExpSinSquared": ['gp_amp', 'gp_timescale', 'gp_per'],
    "QuasiPer": ['gp_per', 'gp_perlength', 'gp_explength', 'gp_amp'],
    "JitterQuasiPer": ['gp_per', 'gp_perlength', 'gp_explength', 'gp_amp', 'gp_jit'],
    "Matern5/2": ['gp_amp', 'gp_timescale'],
    "Matern3/2": ['gp_amp', 'gp_timescale', 'gp_jit'],
}


def PrintKernelList():
    print("Implemented kernels:")
    print(KERNELS)


def defKernelList():
    return KERNELS


def compute_distances(t1, t2):
    """|t1_i - t2_j| and (t1_i - t2_j)^2 as dense [N1, N2] arrays (reference kernels.py:51-78,
    two scipy ``cdist`` calls); evaluated on the GPU (``mrv_distances``)."""
    return engine.distances(np.asarray(t1, dtype=np.float64).ravel(), np.asarray(t2, dtype=np.float64).ravel())


class Kernel(ABC):
    """Parent class of the kernels (reference kernels.py:86-112).  Subclasses set ``_name``,
    ``_keys`` (canonical hyper-parameter order), ``_labels`` (plot labels) and ``_jitter``."""

    _name = None
    _keys = ()
    _labels = ()
    _jitter = False
    _swallow_all = False     # Cosine has a bare ``except`` around the error term (kernels.py:234-237)
    _variant = 0

    def __init__(self, hparams):
        self.covmatrix = None
        self.hparams = hparams          # shadows the static accessor, as in the reference
        msg = "{} kernel requires {} hyperparameters:".format(self._name, len(self._keys)) + ", ".join(
            "'%s'" % k for k in self._keys)
        assert len(self.hparams) == len(self._keys), msg
        try:
            for k in self._keys:
                self.hparams[k].value
        except KeyError:
            raise KeyError(msg)

    @classmethod
    def name(cls):
        return cls._name

    @abc.abstractmethod
    def compute_covmatrix(self, dist_e, dist_se, errors=None):
        pass

    def _hp_vector(self):
        return np.array([float(self.hparams[k].value) for k in self._keys], dtype=np.float64)

    def _compute(self, dist_e, dist_se, errors):
        """Shared body: the element-wise kernel map on the GPU plus the reference's diagonal rules.

        Reference order of operations (e.g. kernels.py:774-789): K, then ``K + identity*jit**2``
        (kept only if the shapes broadcast), then ``+= errors**2 * identity`` (skipped on
        ValueError; a missing ``errors`` raises TypeError except in Cosine's bare except).
        """
        dist_e = np.asarray(dist_e, dtype=np.float64)
        dist_se = np.asarray(dist_se, dtype=np.float64)
        n1, n2 = dist_e.shape
        kid = engine.KERNEL_IDS[self._name]
        hp = self._hp_vector()
        err2 = None
        if errors is None:
            if not self._swallow_all:
                raise TypeError("unsupported operand type(s) for ** or pow(): 'NoneType' and 'int'")
        else:
            try:
                err2 = np.asarray(errors, dtype=np.float64) ** 2
            except Exception:
                if not self._swallow_all:
                    raise
        if n1 == n2:
            mode = 1 if self._jitter else 0
            diag = None
            if err2 is not None:
                try:
                    diag = np.broadcast_to(err2, (n2,)) if err2.ndim <= 1 else None
                except ValueError:
                    diag = None
                if diag is not None:
                    mode |= 2
                elif err2.ndim > 1:
                    # exotic 2-D ``errors``: resolve with the reference's own broadcasting on host
                    K = engine.covmatrix_from_dist(kid, hp, dist_e, dist_se, mode, None, self._variant)
                    try:
                        K += err2 * np.identity(n1)
                    except ValueError:
                        pass
                    self.covmatrix = K
                    return K
            self.covmatrix = engine.covmatrix_from_dist(kid, hp, dist_e, dist_se, mode, diag, self._variant)
            return self.covmatrix
        # non-square K: the identity(n1) terms only survive through NumPy broadcasting corner cases
        # (n1 == 1, or n2 == 1 for the jitter term); reproduce them literally on the GPU result.
        K = engine.covmatrix_from_dist(kid, hp, dist_e, dist_se, 0, None, self._variant)
        cov = K
        if self._jitter:
            sigma = np.identity(n1) * float(self.hparams['gp_jit'].value) ** 2
            try:
                cov = K + sigma
            except ValueError:
                cov = K
        if err2 is not None:
            try:
                cov += err2 * np.identity(n1)
            except ValueError:
                pass
        self.covmatrix = cov
        return cov


def _make_hparams_accessor(keys, labels):
    def hparams(plotting=True):
        if plotting is False:
            return list(keys)
        if plotting is True:
            return list(labels)
    return staticmethod(hparams)


class Cosine(Kernel):
    """K = amp^2 cos(2 pi |t-t'| / per) (reference kernels.py:119-238)."""
    _name = "Cosine"
    _keys = ('gp_amp', 'gp_per')
    _labels = (r'gp$_{amp}$', r'gp$_{per}$')
    _swallow_all = True
    hparams = _make_hparams_accessor(_keys, _labels)

    @property
    def __repr__(self):
        message = "Cosine Kernel with amp: {}, per: {}".format(self.hparams['gp_amp'].value, self.hparams['gp_per'].value)
        print(message)

    def compute_covmatrix(self, dist_e, dist_se, errors=None):
        return self._compute(dist_e, dist_se, errors)


class ExpSquared(Kernel):
    """K = amp^2 exp(-0.5 d^2 / timescale^2) (reference kernels.py:246-366)."""
    _name = "ExpSquared"
    _keys = ('gp_amp', 'gp_timescale')
    _labels = (r'gp$_{amp}$', r'gp$_{timescale}$')
    hparams = _make_hparams_accessor(_keys, _labels)

    @property
    def __repr__(self):
        message = "Exponential Squared Kernel with amp: {}, timescale: {}".format(
            self.hparams['gp_amp'].value, self.hparams['gp_timescale'].value)
        print(message)

    def compute_covmatrix(self, dist_e, dist_se, errors=None):
        return self._compute(dist_e, dist_se, errors)


class ExpSinSquared(Kernel):
    """K = amp^2 exp(-2 sin^2(pi d / per) / timescale^2) (reference kernels.py:374-499)."""
    _name = "ExpSinSquared"
    _keys = ('gp_amp', 'gp_timescale', 'gp_per')
    _labels = (r'gp$_{amp}$', r'gp$_{timescale}$', r'gp$_{per}$')
    hparams = _make_hparams_accessor(_keys, _labels)

    @property
    def __repr__(self):
        message = "Periodic Kernel with amp: {}, timescale: {}, per: {}".format(
            self.hparams['gp_amp'].value, self.hparams['gp_timescale'].value, self.hparams['gp_per'].value)
        print(message)

    def compute_covmatrix(self, dist_e, dist_se, errors=None):
        return self._compute(dist_e, dist_se, errors)


class QuasiPer(Kernel):
    """K = amp^2 exp(-d^2/explength^2) exp(-sin^2(pi d/per)/perlength^2) (reference
    kernels.py:507-640)."""
    _name = "QuasiPer"
    _keys = ('gp_per', 'gp_perlength', 'gp_explength', 'gp_amp')
    _labels = (r'gp$_{per}$', r'gp$_{perlength}$', r'gp$_{explength}$', r'gp$_{amp}$')
    hparams = _make_hparams_accessor(_keys, _labels)

    @property
    def __repr__(self):
        h = self.hparams
        message = "QuasiPeriodic Kernel with amp: {}, per length: {}, per: {}, exp length: {}".format(
            h['gp_amp'].value, h['gp_perlength'].value, h['gp_per'].value, h['gp_explength'].value)
        print(message)
        return message

    def compute_covmatrix(self, dist_e, dist_se, errors=None):
        return self._compute(dist_e, dist_se, errors)


class JitterQuasiPer(Kernel):
    """QuasiPer + jit^2 on the diagonal (reference kernels.py:648-791)."""
    _name = "JitterQuasiPer"
    _keys = ('gp_per', 'gp_perlength', 'gp_explength', 'gp_amp', 'gp_jit')
    _labels = (r'gp$_{per}$', r'gp$_{perlength}$', r'gp$_{explength}$', r'gp$_{amp}$', r'gp$_{jit}$')
    _jitter = True
    # the reference's text list misspells gp_perlength (kernels.py:717); kept for drop-in output
    hparams = _make_hparams_accessor(('gp_per', 'gp_perlegth', 'gp_explength', 'gp_amp', 'gp_jit'), _labels)

    @property
    def __repr__(self):
        h = self.hparams
        message = "QuasiPeriodic Kernel with amp: {}, per length: {}, per: {}, exp length: {}, jit: {}".format(
            h['gp_amp'].value, h['gp_perlength'].value, h['gp_per'].value, h['gp_explength'].value, h['gp_jit'].value)
        print(message)
        return message

    def compute_covmatrix(self, dist_e, dist_se, errors=None):
        return self._compute(dist_e, dist_se, errors)


class Matern5(Kernel):
    """Matern 5/2 AS CODED in the reference (kernels.py:909): amp^2 (1 + sqrt5 d/ts +
    (5 d^4 / 3 ts^2) exp(-sqrt5 d/ts)) -- not a valid covariance, kept verbatim
    for parity.  ``Matern5Textbook`` is the opt-in valid kernel."""
    _name = "Matern5/2"
    _keys = ('gp_amp', 'gp_timescale')
    _labels = (r'gp$_{amp}$', r'gp$_{timescale}$')
    hparams = _make_hparams_accessor(_keys, _labels)

    @property
    def __repr__(self):
        message = "Matern 5/2 Kernel with amp: {}, timescale: {}".format(
            self.hparams['gp_amp'].value, self.hparams['gp_timescale'].value)
        print(message)
        return message

    def compute_covmatrix(self, dist_e, dist_se, errors=None):
        return self._compute(dist_e, dist_se, errors)


class Matern5Textbook(Matern5):
    """Opt-in textbook Matern 5/2: amp^2 (1 + sqrt5 d/ts + 5 d^2/(3 ts^2)) exp(-sqrt5 d/ts).
    No counterpart in the reference; never selected by a reference kernel name."""
    _variant = 1


class Matern3(Kernel):
    """K = amp^2 (1 + sqrt3 d/ts) exp(-sqrt3 d/ts) + jit^2 I (reference kernels.py:927-1057)."""
    _name = "Matern3/2"
    _keys = ('gp_amp', 'gp_timescale', 'gp_jit')
    _labels = (r'gp$_{amp}$', r'gp$_{timescale}$', r'gp$_{jit}$')
    _jitter = True
    hparams = _make_hparams_accessor(_keys, _labels)

    @property
    def __repr__(self):
        h = self.hparams
        message = "Matern 3/2 Kernel with amp: {}, timescale: {}, jit: {}".format(
            h['gp_amp'].value, h['gp_timescale'].value, h['gp_jit'].value)
        print(message)
        return message

    def compute_covmatrix(self, dist_e, dist_se, errors=None):
        return self._compute(dist_e, dist_se, errors)


KERNEL_CLASSES = {"Cosine": Cosine, "ExpSquared": ExpSquared, "ExpSinSquared": ExpSinSquared,
                  "QuasiPer": QuasiPer, "JitterQuasiPer": JitterQuasiPer, "Matern5/2": Matern5,
                  "Matern3/2": Matern3}
