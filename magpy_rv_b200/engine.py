"""Host-side problem descriptor: flattens the reference's Python world (kernel name, model list,
prior list, parameter dictionaries) into the POD structs of the C ABI and wraps the batched calls.

This is the boundary SURVEY.md section 8(b) describes: it sits exactly where the reference's
per-walker loop is (``/root/reference/src/magpy_rv/mcmc.py:410-461``) and behind
``GPLikelihood.logprob/LogL`` (``gp_likelihood.py:250,357``).
"""
import ctypes

import numpy as np

from . import _native as nat
from . import parameters as par

KERNEL_IDS = {"Cosine": 0, "ExpSquared": 1, "ExpSinSquared": 2, "QuasiPer": 3, "JitterQuasiPer": 4,
              "Matern5/2": 5, "Matern3/2": 6}
KERNEL_NH = {0: 2, 1: 2, 2: 3, 3: 4, 4: 5, 5: 2, 6: 3}
HAS_JITTER = {4: 4, 6: 2}   # kernel id -> index of gp_jit in the hyper-parameter vector

MODEL_NONE, MODEL_POLY, MODEL_OFFSET, MODEL_KEPLER = 0, 1, 2, 3
MODEL_NPAR = {MODEL_NONE: 1, MODEL_POLY: 4, MODEL_OFFSET: 1, MODEL_KEPLER: 5}


def canonical_kernel(kernel_name):
    """Name dispatch of GPLikelihood.compute_kernel (reference gp_likelihood.py:198-221): each
    prefix rule is tested in turn and the last match wins; no match -> AssertionError."""
    rules = (
        (("Cos", "cos"), "Cosine"),
        (("ExpSquare", "expsquare", "Expsquare", "expSquare"), "ExpSquared"),
        (("Periodic", "periodic", "ExpSin", "Expsin", "expsin"), "ExpSinSquared"),
        (("QuasiPer", "quasiper", "Quasiper"), "QuasiPer"),
        (("jit", "Jit"), "JitterQuasiPer"),
        (("Matern5", "matern5", "Matern 5", "matern 5"), "Matern5/2"),
        (("Matern3", "matern3", "Matern 3", "matern 3"), "Matern3/2"),
    )
    name = kernel_name
    for prefixes, canon in rules:
        if name.startswith(prefixes):
            name = canon
    assert name in KERNEL_IDS, 'Kernel not yet implemented. Pick from available kernels: ' + str(KERNEL_IDS.keys())
    return name


def model_type(name):
    """if/elif chain of get_model (reference mcmc_aux.py:77-92)."""
    if name.startswith(("no", "No")):
        return MODEL_NONE
    if name.startswith(("off", "Off")):
        return MODEL_OFFSET
    if name.startswith(("poly", "Poly")):
        return MODEL_POLY
    if name.startswith(("kep", "Kep")):
        return MODEL_KEPLER
    raise KeyError("model not yet implemented, please from currently implemented models: "
                   "dict_keys(['No_Model', 'Offset', 'Polynomial', 'Keplerian'])")


def model_ops(model_list):
    """[(type, param_offset, flag_val)] in model_list order; Offset flag values follow the key the
    reference's mod_create would give that model ('offset' / 'offset_j' -> 1 / j+1,
    models.py:419-433)."""
    if isinstance(model_list, str):
        model_list = [model_list]
    ops, off, n_off = [], 0, 0
    for name in model_list:
        t = model_type(name)
        flag_val = 0.0
        if t == MODEL_OFFSET:
            flag_val = n_off + 1.0
            n_off += 1
        ops.append((t, off, flag_val))
        off += MODEL_NPAR[t]
    return ops, off


class LikelihoodProblem:
    """One (t, y, yerr, flags, kernel, model list, prior list) problem resident on one GPU.

    ``logl(hp, modpar)`` evaluates every walker of a [W, nh] / [W, nm] ensemble in one call.
    """

    def __init__(self, t, y, yerr, kernel_name, model_list=("no",), prior_list=(), hparam_names=None,
                 model_param_names=None, flags=None, device=None, kernel_variant=0):
        self.t = nat.as_f64(t)
        self.y = nat.as_f64(y)
        self.yerr = nat.as_f64(yerr)
        self.N = int(self.t.shape[0])
        assert self.y.shape == (self.N,) and self.yerr.shape == (self.N,), "t, y, yerr must have equal length"
        self.flags = None if flags is None else nat.as_f64(flags)
        self.kernel = canonical_kernel(kernel_name)
        self.kernel_id = KERNEL_IDS[self.kernel]
        self.kernel_variant = int(kernel_variant)
        self.nh = KERNEL_NH[self.kernel_id]
        if model_list is None:
            # caller supplies model curves itself (logl_given_model); only the parameter count matters
            assert model_param_names is not None
            self.model_list, self.ops, self.nm = None, [], len(model_param_names)
        else:
            self.model_list = [model_list] if isinstance(model_list, str) else list(model_list)
            self.ops, self.nm = model_ops(self.model_list)
        if any(o[0] == MODEL_OFFSET for o in self.ops) and self.flags is None:
            raise TypeError("Offset models need a flags array (object of type 'NoneType' has no len())")
        self.hparam_names = list(hparam_names) if hparam_names is not None else list(par.par_create(kernel_name).keys())
        if model_param_names is None:
            from .models import mod_create
            model_param_names = list(mod_create(self.model_list).keys())
        self.model_param_names = list(model_param_names)
        self.priors = self._flatten_priors(prior_list)
        self.device = nat.default_device() if device is None else int(device)
        self._handle = ctypes.c_void_p()
        self._create()

    def _pod_arrays(self):
        ops_arr = _ops_array(self.ops)
        pri_arr = (nat.mrv_prior * max(1, len(self.priors)))()
        for i, (tp, src, idx, p0, p1, p2) in enumerate(self.priors):
            pri_arr[i].type, pri_arr[i].source, pri_arr[i].index = tp, src, idx
            pri_arr[i].p0, pri_arr[i].p1, pri_arr[i].p2 = p0, p1, p2
        return ops_arr, pri_arr

    def _create(self):
        ops_arr, pri_arr = self._pod_arrays()
        rc = nat.lib().mrv_problem_create(ctypes.byref(self._handle), self.device, nat.dptr(self.t), nat.dptr(self.y),
                                          nat.dptr(self.yerr), nat.dptr(self.flags), self.N, self.kernel_id,
                                          self.kernel_variant, self.nh, ops_arr, len(self.ops), self.nm, pri_arr,
                                          len(self.priors))
        nat.check(rc, "mrv_problem_create")

    # -- flattening -----------------------------------------------------------------------------
    def _flatten_priors(self, prior_list):
        out = []
        for item in prior_list:
            pname, prior_name, pdict = item[0], item[1], item[2]
            # hparams first, then model params (reference gp_likelihood.py:308-313)
            if pname in self.hparam_names:
                src, idx = 0, self.hparam_names.index(pname)
            elif pname in self.model_param_names:
                src, idx = 1, self.model_param_names.index(pname)
            else:
                raise KeyError(pname)
            code, p0, p1, p2 = par.prior_code(prior_name, pdict)
            out.append((code, src, idx, p0, p1, p2))
        return out

    # -- calls ----------------------------------------------------------------------------------
    def _check_shapes(self, hp, modpar):
        hp = nat.as_f64(hp)
        modpar = nat.as_f64(modpar)
        if hp.ndim == 1:
            hp = hp[None, :]
        if modpar.ndim == 1:
            modpar = modpar[None, :]
        assert hp.shape[1] == self.nh, "expected %d hyper-parameters per walker, got %d" % (self.nh, hp.shape[1])
        assert modpar.shape[1] == self.nm, "expected %d model parameters per walker, got %d" % (self.nm, modpar.shape[1])
        assert hp.shape[0] == modpar.shape[0]
        return hp, modpar

    def logl(self, hp, modpar, to_ecc=True, raise_on_status=False):
        """(logL [W], status [W]) for the ensemble; see mrv_logl_batch."""
        hp, modpar = self._check_shapes(hp, modpar)
        W = hp.shape[0]
        out = np.empty(W, dtype=np.float64)
        status = np.empty(W, dtype=np.int32)
        rc = nat.lib().mrv_logl_batch(self._handle, nat.dptr(hp), nat.dptr(modpar), W, int(bool(to_ecc)),
                                      nat.dptr(out), status.ctypes.data_as(nat.c_int32_p))
        nat.check(rc, "mrv_logl_batch")
        if raise_on_status:
            raise_for_status(status)
        return out, status

    def logl_device(self, d_hp, d_modpar, W, d_logl, d_status, to_ecc=True, stream=0):
        """Device-pointer variant (integers / ``tensor.data_ptr()``), asynchronous on ``stream``."""
        rc = nat.lib().mrv_logl_batch_device(self._handle, ctypes.c_void_p(d_hp), ctypes.c_void_p(d_modpar), int(W),
                                             int(bool(to_ecc)), ctypes.c_void_p(d_logl), ctypes.c_void_p(d_status),
                                             ctypes.c_void_p(stream))
        nat.check(rc, "mrv_logl_batch_device")

    def logl_given_model(self, hp, modpar, model_y):
        hp, modpar = self._check_shapes(hp, modpar)
        W = hp.shape[0]
        model_y = nat.as_f64(model_y).reshape(W, self.N)
        out = np.empty(W, dtype=np.float64)
        status = np.empty(W, dtype=np.int32)
        rc = nat.lib().mrv_logl_given_model(self._handle, nat.dptr(hp), nat.dptr(modpar), nat.dptr(model_y), W,
                                            nat.dptr(out), status.ctypes.data_as(nat.c_int32_p))
        nat.check(rc, "mrv_logl_given_model")
        return out, status

    def solve(self, hp, rhs):
        """(z [W, N] = L^-1 rhs, logdet [W], status [W]) -- see mrv_solve_batch."""
        hp = nat.as_f64(hp)
        if hp.ndim == 1:
            hp = hp[None, :]
        rhs = nat.as_f64(rhs).reshape(-1, self.N)
        W = rhs.shape[0]
        if hp.shape[0] == 1 and W > 1:
            hp = np.ascontiguousarray(np.broadcast_to(hp, (W, self.nh)))
        assert hp.shape == (W, self.nh)
        z = np.empty((W, self.N), dtype=np.float64)
        logdet = np.empty(W, dtype=np.float64)
        status = np.empty(W, dtype=np.int32)
        rc = nat.lib().mrv_solve_batch(self._handle, nat.dptr(hp), nat.dptr(rhs), W, nat.dptr(z), nat.dptr(logdet),
                                       status.ctypes.data_as(nat.c_int32_p))
        nat.check(rc, "mrv_solve_batch")
        return z, logdet, status

    def solve_multi(self, hp, rhs):
        """(z [R, N] = L^-1 rhs_r, logdet, status) for ONE hyper-parameter set: one factorisation, R forward
        substitutions through the stored factor -- see mrv_solve_multi."""
        hp = nat.as_f64(hp).ravel()
        assert hp.shape == (self.nh,)
        rhs = nat.as_f64(rhs).reshape(-1, self.N)
        R = rhs.shape[0]
        z = np.empty((R, self.N), dtype=np.float64)
        logdet = np.empty(1, dtype=np.float64)
        status = np.empty(1, dtype=np.int32)
        rc = nat.lib().mrv_solve_multi(self._handle, nat.dptr(hp), nat.dptr(rhs), R, nat.dptr(z), nat.dptr(logdet),
                                       status.ctypes.data_as(nat.c_int32_p))
        nat.check(rc, "mrv_solve_multi")
        return z, float(logdet[0]), status

    def set_option(self, key, value):
        nat.check(nat.lib().mrv_set_option(self._handle, key.encode(), int(value)), "mrv_set_option(%s)" % key)

    def info(self, key):
        return int(nat.lib().mrv_get_info(self._handle, key.encode()))

    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            nat.lib().mrv_problem_destroy(self._handle)
            self._handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LibraryShardedProblem(LikelihoodProblem):
    """The same problem replicated on several GPUs of the box by the library itself (``mrv_shard_*``): ONE host
    process and thread, no process group, no torch.  ``logl`` cuts the ensemble into contiguous walker blocks,
    keeps every device busy and returns the full vectors; per-walker values are bitwise those of one GPU.
    (The one-process-per-GPU form with NCCL is ``sharding.ShardedEvaluator``.)"""

    def __init__(self, *args, devices=(0,), **kw):
        self.devices = [int(d) for d in devices]
        assert len(self.devices) >= 1
        kw["device"] = self.devices[0]
        self._shard = ctypes.c_void_p()
        super().__init__(*args, **kw)

    def _create(self):
        ops_arr, pri_arr = self._pod_arrays()
        devs = (ctypes.c_int * len(self.devices))(*self.devices)
        rc = nat.lib().mrv_shard_create(ctypes.byref(self._shard), len(self.devices), devs, nat.dptr(self.t), nat.dptr(self.y),
                                        nat.dptr(self.yerr), nat.dptr(self.flags), self.N, self.kernel_id, self.kernel_variant,
                                        self.nh, ops_arr, len(self.ops), self.nm, pri_arr, len(self.priors))
        nat.check(rc, "mrv_shard_create")
        self._handle = ctypes.c_void_p(nat.lib().mrv_shard_part(self._shard, 0))     # borrowed: info / single-device calls

    def logl(self, hp, modpar, to_ecc=True, raise_on_status=False):
        hp, modpar = self._check_shapes(hp, modpar)
        W = hp.shape[0]
        out = np.empty(W, dtype=np.float64)
        status = np.empty(W, dtype=np.int32)
        rc = nat.lib().mrv_shard_logl_batch(self._shard, nat.dptr(hp), nat.dptr(modpar), W, int(bool(to_ecc)), nat.dptr(out),
                                            status.ctypes.data_as(nat.c_int32_p))
        nat.check(rc, "mrv_shard_logl_batch")
        if raise_on_status:
            raise_for_status(status)
        return out, status

    def set_option(self, key, value):
        for i in range(len(self.devices)):
            part = ctypes.c_void_p(nat.lib().mrv_shard_part(self._shard, i))
            nat.check(nat.lib().mrv_set_option(part, key.encode(), int(value)), "mrv_set_option(%s)" % key)

    def close(self):
        if getattr(self, "_shard", None) is not None and self._shard.value:
            nat.lib().mrv_shard_destroy(self._shard)
            self._shard = ctypes.c_void_p()
            self._handle = ctypes.c_void_p()


def raise_for_status(status):
    """Turn per-walker status into the exception the reference would have raised
    (cho_factor: LinAlgError for a non-PD matrix, ValueError for non-finite input)."""
    status = np.asarray(status)
    bad = np.nonzero(status != 0)[0]
    if bad.size == 0:
        return
    w = int(bad[0])
    s = int(status[w])
    if s == -3:
        raise RuntimeError("magpy_rv_b200 internal error: the dataflow factorisation kernel gave up waiting for a tile "
                           "dependency (walker %d); no result was produced" % w)
    if s > 0:
        raise np.linalg.LinAlgError("%d-th leading minor of the array is not positive definite (walker %d)" % (s, w))
    raise ValueError("array must not contain infs or NaNs (walker %d)" % w)


# ---------------------------------------------------------------------------------------------
# dense helpers (no problem handle)
# ---------------------------------------------------------------------------------------------

def _ops_array(ops):
    arr = (nat.mrv_model_op * max(1, len(ops)))()
    for i, (tp, off, fv) in enumerate(ops):
        arr[i].type, arr[i].param_offset, arr[i].flag_val = tp, off, fv
    return arr


def model_eval(model_list, t, modpar, to_ecc=False, flags=None, device=None):
    """(model_y [W, N], modpar_phys [W, nm]) for a model list -- see mrv_model_eval."""
    ops, nm = model_ops(model_list)
    return model_eval_ops(ops, nm, t, modpar, to_ecc, flags, device)


def model_eval_ops(ops, nm, t, modpar, to_ecc=False, flags=None, device=None):
    """Same with an explicit op list [(type, param_offset, flag_val)]."""
    t = nat.as_f64(t).ravel()
    modpar = nat.as_f64(modpar)
    if modpar.ndim == 1:
        modpar = modpar[None, :]
    assert modpar.shape[1] == nm, "expected %d model parameters, got %d" % (nm, modpar.shape[1])
    fl = None
    if any(o[0] == MODEL_OFFSET for o in ops):
        if flags is None:
            raise TypeError("object of type 'NoneType' has no len()")
        fl = nat.as_f64(flags).ravel()
        assert fl.shape[0] == t.shape[0]
    W, N = modpar.shape[0], t.shape[0]
    out = np.empty((W, N), dtype=np.float64)
    phys = np.empty((W, nm), dtype=np.float64)
    device = nat.default_device() if device is None else device
    rc = nat.lib().mrv_model_eval(device, nat.dptr(t), nat.dptr(fl), N, _ops_array(ops), len(ops), nm,
                                  nat.dptr(modpar), W, int(bool(to_ecc)), nat.dptr(out), nat.dptr(phys))
    nat.check(rc, "mrv_model_eval")
    return out, phys


def kepler_anomaly(M, ecc, max_itr=200, device=None):
    """Eccentric anomaly of a mean-anomaly vector -- see mrv_kepler_anomaly (reference models.py:618-650)."""
    M = np.asarray(M, dtype=np.float64)
    flat = np.ascontiguousarray(M.ravel())
    out = np.empty_like(flat)
    device = nat.default_device() if device is None else device
    rc = nat.lib().mrv_kepler_anomaly(device, nat.dptr(flat), flat.shape[0], float(ecc), int(max_itr), nat.dptr(out))
    nat.check(rc, "mrv_kepler_anomaly")
    return out.reshape(M.shape)


def true_anomaly(E, ecc, device=None):
    """True anomaly of an eccentric-anomaly vector -- see mrv_true_anomaly (reference models.py:653-669)."""
    E = np.asarray(E, dtype=np.float64)
    flat = np.ascontiguousarray(E.ravel())
    out = np.empty_like(flat)
    device = nat.default_device() if device is None else device
    rc = nat.lib().mrv_true_anomaly(device, nat.dptr(flat), flat.shape[0], float(ecc), nat.dptr(out))
    nat.check(rc, "mrv_true_anomaly")
    return out.reshape(E.shape)


def predict_reduce(V, z, Kss, full_cov=False, device=None):
    """(mean [Np], cov [Np, Np] or var [Np]) -- see mrv_predict_reduce."""
    V = nat.as_f64(V)
    z = nat.as_f64(z).ravel()
    Kss = nat.as_f64(Kss)
    Np, N = V.shape
    assert z.shape[0] == N and Kss.shape == (Np, Np)
    mean = np.empty(Np, dtype=np.float64)
    cov = np.empty((Np, Np) if full_cov else (Np,), dtype=np.float64)
    device = nat.default_device() if device is None else device
    rc = nat.lib().mrv_predict_reduce(device, nat.dptr(V), nat.dptr(z), nat.dptr(Kss), Np, N, int(bool(full_cov)),
                                      nat.dptr(mean), nat.dptr(cov))
    nat.check(rc, "mrv_predict_reduce")
    return mean, cov


def _diag_args(diag_mode, diag_add, n2):
    d = None
    if diag_mode & 2:
        d = nat.as_f64(np.broadcast_to(np.asarray(diag_add, dtype=np.float64), (n2,)))
    return d


def covmatrix(kernel_id, hp, x1, x2, diag_mode=0, diag_add=None, kernel_variant=0, device=None):
    x1 = nat.as_f64(x1).ravel()
    x2 = nat.as_f64(x2).ravel()
    hp = nat.as_f64(hp).ravel()
    n1, n2 = x1.shape[0], x2.shape[0]
    K = np.empty((n1, n2), dtype=np.float64)
    d = _diag_args(diag_mode, diag_add, n2)
    device = nat.default_device() if device is None else device
    rc = nat.lib().mrv_covmatrix(device, kernel_id, kernel_variant, nat.dptr(hp), hp.shape[0], nat.dptr(x1), n1,
                                 nat.dptr(x2), n2, diag_mode, nat.dptr(d), nat.dptr(K))
    nat.check(rc, "mrv_covmatrix")
    return K


def covmatrix_from_dist(kernel_id, hp, dist_e, dist_se, diag_mode=0, diag_add=None, kernel_variant=0, device=None):
    dist_e = nat.as_f64(dist_e)
    dist_se = nat.as_f64(dist_se)
    assert dist_e.ndim == 2 and dist_e.shape == dist_se.shape
    hp = nat.as_f64(hp).ravel()
    n1, n2 = dist_e.shape
    K = np.empty((n1, n2), dtype=np.float64)
    d = _diag_args(diag_mode, diag_add, n2)
    device = nat.default_device() if device is None else device
    rc = nat.lib().mrv_covmatrix_from_dist(device, kernel_id, kernel_variant, nat.dptr(hp), hp.shape[0],
                                           nat.dptr(dist_e), nat.dptr(dist_se), n1, n2, diag_mode, nat.dptr(d),
                                           nat.dptr(K))
    nat.check(rc, "mrv_covmatrix_from_dist")
    return K


def distances(t1, t2, device=None):
    t1 = nat.as_f64(t1).ravel()
    t2 = nat.as_f64(t2).ravel()
    n1, n2 = t1.shape[0], t2.shape[0]
    de = np.empty((n1, n2), dtype=np.float64)
    dse = np.empty((n1, n2), dtype=np.float64)
    device = nat.default_device() if device is None else device
    rc = nat.lib().mrv_distances(device, nat.dptr(t1), n1, nat.dptr(t2), n2, nat.dptr(de), nat.dptr(dse))
    nat.check(rc, "mrv_distances")
    return de, dse
