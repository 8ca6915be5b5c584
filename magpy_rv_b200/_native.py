"""ctypes binding of the C ABI in ``include/magpy_rv_b200.h`` (csrc/abi.cu).

The shared library is built in-tree by ``__graft_entry__.build()`` / ``magpy_rv_b200.build``
(nvcc, sm_100a) as ``magpy_rv_b200/_lib/libmagpy_rv_b200.so``.  There is NO CPU fallback: if the
library is missing, or no CUDA device is present when a compute entry point is called, the call
raises.
"""
import ctypes
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MRV_LIB: another build of the same library (A/B probes of kernel variants, tools/build_variant.py); unset in normal use
LIB_PATH = os.environ.get("MRV_LIB") or os.path.join(_HERE, "_lib", "libmagpy_rv_b200.so")

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int32_p = ctypes.POINTER(ctypes.c_int32)


class mrv_model_op(ctypes.Structure):
    _fields_ = [("type", ctypes.c_int32), ("param_offset", ctypes.c_int32), ("flag_val", ctypes.c_double)]


class mrv_prior(ctypes.Structure):
    _fields_ = [("type", ctypes.c_int32), ("source", ctypes.c_int32), ("index", ctypes.c_int32),
                ("pad", ctypes.c_int32), ("p0", ctypes.c_double), ("p1", ctypes.c_double),
                ("p2", ctypes.c_double)]


# name -> (restype, argtypes); also the list the "exports every declared symbol" test walks
SIGNATURES = {
    "mrv_problem_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, c_double_p, c_double_p,
                                          c_double_p, c_double_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_int, ctypes.POINTER(mrv_model_op), ctypes.c_int, ctypes.c_int,
                                          ctypes.POINTER(mrv_prior), ctypes.c_int]),
    "mrv_problem_destroy": (None, [ctypes.c_void_p]),
    "mrv_logl_batch": (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_double_p, ctypes.c_int64, ctypes.c_int,
                                      c_double_p, c_int32_p]),
    "mrv_logl_batch_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                             ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "mrv_shard_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.POINTER(ctypes.c_int), c_double_p,
                                        c_double_p, c_double_p, c_double_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.POINTER(mrv_model_op), ctypes.c_int, ctypes.c_int,
                                        ctypes.POINTER(mrv_prior), ctypes.c_int]),
    "mrv_shard_destroy": (None, [ctypes.c_void_p]),
    "mrv_shard_count": (ctypes.c_int, [ctypes.c_void_p]),
    "mrv_shard_part": (ctypes.c_void_p, [ctypes.c_void_p, ctypes.c_int]),
    "mrv_shard_logl_batch": (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_double_p, ctypes.c_int64, ctypes.c_int,
                                            c_double_p, c_int32_p]),
    "mrv_logl_given_model": (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_double_p, c_double_p, ctypes.c_int64,
                                            c_double_p, c_int32_p]),
    "mrv_solve_batch": (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_double_p, ctypes.c_int64, c_double_p, c_double_p,
                                       c_int32_p]),
    "mrv_solve_multi": (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_double_p, ctypes.c_int64, c_double_p, c_double_p,
                                       c_int32_p]),
    "mrv_predict_reduce": (ctypes.c_int, [ctypes.c_int, c_double_p, c_double_p, c_double_p, ctypes.c_int64,
                                          ctypes.c_int64, ctypes.c_int, c_double_p, c_double_p]),
    "mrv_model_eval": (ctypes.c_int, [ctypes.c_int, c_double_p, c_double_p, ctypes.c_int64,
                                      ctypes.POINTER(mrv_model_op), ctypes.c_int, ctypes.c_int, c_double_p,
                                      ctypes.c_int64, ctypes.c_int, c_double_p, c_double_p]),
    "mrv_covmatrix": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, c_double_p, ctypes.c_int,
                                     c_double_p, ctypes.c_int64, c_double_p, ctypes.c_int64, ctypes.c_int,
                                     c_double_p, c_double_p]),
    "mrv_covmatrix_from_dist": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, c_double_p, ctypes.c_int,
                                               c_double_p, c_double_p, ctypes.c_int64, ctypes.c_int64,
                                               ctypes.c_int, c_double_p, c_double_p]),
    "mrv_distances": (ctypes.c_int, [ctypes.c_int, c_double_p, ctypes.c_int64, c_double_p, ctypes.c_int64,
                                     c_double_p, c_double_p]),
    "mrv_mcmc_device": (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_double_p, ctypes.c_int64, ctypes.c_int64,
                                       ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64,
                                       c_double_p, c_double_p, c_double_p, c_double_p, c_int32_p]),
    "mrv_gelman_rubin": (ctypes.c_int, [ctypes.c_int, c_double_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                        ctypes.c_int64, c_double_p]),
    # host-only helper called twice per MCMC iteration: raw addresses (c_void_p) keep the ctypes overhead low
    "mrv_kepler_anomaly": (ctypes.c_int, [ctypes.c_int, c_double_p, ctypes.c_int64, ctypes.c_double, ctypes.c_int, c_double_p]),
    "mrv_true_anomaly": (ctypes.c_int, [ctypes.c_int, c_double_p, ctypes.c_int64, ctypes.c_double, c_double_p]),
    "mrv_stretch_scan": (ctypes.c_int64, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int32,
                                          ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_double, c_int32_p, ctypes.c_int32, ctypes.POINTER(ctypes.c_int64),
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "mrv_stretch_split": (ctypes.c_int64, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int32,
                                           ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, c_int32_p,
                                           ctypes.c_int32, ctypes.POINTER(ctypes.c_int64), ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_void_p]),
    "mrv_set_option": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64]),
    "mrv_get_info": (ctypes.c_int64, [ctypes.c_void_p, ctypes.c_char_p]),
    "mrv_device_count": (ctypes.c_int, []),
    "mrv_last_error": (ctypes.c_char_p, []),
    "mrv_abi_version": (ctypes.c_int, []),
    "mrv_fp64_mma_peak": (ctypes.c_double, [ctypes.c_int, ctypes.c_int]),
    "mrv_fp64_fma_peak": (ctypes.c_double, [ctypes.c_int, ctypes.c_int]),
    "mrv_cov_eval_rate": (ctypes.c_double, [ctypes.c_int, ctypes.c_int, ctypes.c_int, c_double_p, ctypes.c_int]),
}

_lib = None
_lock = threading.Lock()


class NativeLibraryMissing(ImportError):
    pass


def lib():
    """Load (once) and return the ctypes library; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise NativeLibraryMissing(
                    "magpy_rv_b200: %s not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
            handle = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(handle, name)
                fn.restype = res
                fn.argtypes = args
            _lib = handle
    return _lib


def last_error():
    msg = lib().mrv_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc, what):
    if rc != 0:
        raise RuntimeError("magpy_rv_b200: %s failed (rc=%d): %s" % (what, rc, last_error()))


def as_f64(a):
    """C-contiguous float64 view/copy of ``a``."""
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def dptr(a):
    return a.ctypes.data_as(c_double_p) if a is not None else None


def default_device():
    return int(os.environ.get("MRV_DEVICE", os.environ.get("LOCAL_RANK", "0")))
