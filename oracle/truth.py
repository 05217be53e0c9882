"""ctypes front end of oracle/ekld_truth.c (80-bit long-double arbiter) -- TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libekld_truth.so")


def build(force=False):
    src = os.path.join(_HERE, "ekld_truth.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        dp = ctypes.POINTER(ctypes.c_double)
        _lib.ekld_truth.restype = ctypes.c_int
        _lib.ekld_truth.argtypes = [dp, dp, ctypes.c_int, ctypes.c_int, dp, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_double, ctypes.c_double, dp, ctypes.c_long,
                                    dp, dp, dp, dp, dp, dp, dp, dp, ctypes.c_long]
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def ekld_truth(X, Y, X_design, hyp, nugget=1e-3, jitter=1e-8, quad=None):
    """dict(kld (S,Nd), sigma_x (S,Nd), v (S,Nd), mu_1 (S), sigma_1 (S), loglik (S)) rounded from long double."""
    lib = _load()
    X = np.ascontiguousarray(X, dtype=np.float64)
    Y = np.ascontiguousarray(Y, dtype=np.float64).ravel()
    Xd = np.ascontiguousarray(np.atleast_2d(X_design), dtype=np.float64)
    hyp = np.ascontiguousarray(hyp, dtype=np.float64)
    n, d = X.shape
    S, ndim = hyp.shape
    Nd = Xd.shape[0]
    kld = np.empty((S, Nd)); sx = np.empty((S, Nd)); v = np.empty((S, Nd))
    mu1 = np.empty(S); s1 = np.empty(S); ll = np.empty(S)
    Xq = wq = None
    Nq = 0
    if quad is not None:
        Xq = np.ascontiguousarray(np.atleast_2d(quad[0]), dtype=np.float64)
        Nq = Xq.shape[0]
        wq = None if quad[1] is None else np.ascontiguousarray(quad[1], dtype=np.float64)
    rc = lib.ekld_truth(_p(X), _p(Y), n, d, _p(hyp), S, ndim, float(nugget) ** 2, float(jitter), _p(Xd), Nd,
                        _p(kld), _p(sx), _p(v), _p(mu1), _p(s1), _p(ll),
                        _p(Xq) if Xq is not None else None, _p(wq) if wq is not None else None, Nq)
    if rc != 0:
        raise np.linalg.LinAlgError("long-double Cholesky failed at sample %d" % (rc - 1))
    return dict(kld=kld, sigma_x=sx, v=v, mu_1=mu1, sigma_1=s1, loglik=ll)
