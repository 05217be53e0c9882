"""ctypes binding of ``libbode_b200.so`` (the C ABI declared in ``include/bode_b200.h``).

There is no CPU fallback: if the shared library is missing or no CUDA device is usable every compute call raises.
PyTorch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BODE_B200_LIB", os.path.join(_HERE, "lib", "libbode_b200.so"))  # env override: developer builds

BODE_FORM_REFERENCE = 0
BODE_FORM_LOG1P = 1
BODE_MAX_N = 1024
BODE_MAX_D = 16

EXPORTS = (
    "bode_version", "bode_strerror", "bode_last_cuda_error", "bode_ekld_workspace_bytes", "bode_ekld_prepare_dev",
    "bode_ekld_sample_terms_dev", "bode_ekld_mu_sigma_dev", "bode_ekld_score_scratch_bytes", "bode_ekld_score_dev", "bode_ekld_rescore_dev",
    "bode_ekld_score_dev_timed", "bode_ekld_score_dev_events", "bode_us_variance_dev", "bode_quad_scratch_bytes", "bode_quad_rowmean_dev",
    "bode_ekld_prepare_quad_dev", "bode_ekld_score_quad_dev", "bode_ekld_score_host", "bode_gp_loglik_host", "bode_gp_loglik_dev", "bode_ekld_plan_describe", "bode_fp64_peak",
    "bode_tuning_reload", "bode_ctx_create", "bode_ctx_destroy", "bode_ctx_alloc_count", "bode_ekld_score_host_ctx", "bode_gp_loglik_host_ctx",
    "bode_nccl_unique_id", "bode_nccl_comm_create", "bode_nccl_comm_destroy", "bode_ekld_argmax_nccl", "bode_argmax_combine",
    "bode_xchg_create", "bode_xchg_connect", "bode_xchg_destroy", "bode_ekld_argmax_p2p",
    "bode_ekld_prepare_jitter_dev", "bode_ekld_prepare_quad_rm_dev",
)


class BodeError(RuntimeError):
    """A non-zero status returned across the C ABI."""

    def __init__(self, status, detail=""):
        self.status = status
        super().__init__("bode_b200 status %d: %s" % (status, detail))


_lib = None


def load():
    """Load the shared library (built in-tree by ``__graft_entry__.build()`` / ``make -C bode_b200/csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "bode_b200: %s not found -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    c_dp = ctypes.c_void_p
    lib.bode_version.restype = ctypes.c_char_p
    lib.bode_strerror.restype = ctypes.c_char_p
    lib.bode_strerror.argtypes = [ctypes.c_int]
    lib.bode_last_cuda_error.restype = ctypes.c_char_p
    lib.bode_ekld_workspace_bytes.restype = ctypes.c_size_t
    lib.bode_ekld_workspace_bytes.argtypes = [ctypes.c_int] * 3
    lib.bode_ekld_prepare_dev.restype = ctypes.c_int
    lib.bode_ekld_prepare_dev.argtypes = [c_dp, c_dp, c_dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_double, ctypes.c_double, c_dp, ctypes.c_size_t, c_dp, c_dp]
    lib.bode_ekld_prepare_jitter_dev.restype = ctypes.c_int
    lib.bode_ekld_prepare_jitter_dev.argtypes = [c_dp, c_dp, c_dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                 ctypes.c_double, c_dp, c_dp, ctypes.c_size_t, c_dp, c_dp]
    lib.bode_ekld_sample_terms_dev.restype = ctypes.c_int
    lib.bode_ekld_sample_terms_dev.argtypes = [c_dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_dp, c_dp]
    lib.bode_ekld_mu_sigma_dev.restype = ctypes.c_int
    lib.bode_ekld_mu_sigma_dev.argtypes = [c_dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_dp, c_dp]
    lib.bode_ekld_score_scratch_bytes.restype = ctypes.c_size_t
    lib.bode_ekld_score_scratch_bytes.argtypes = [ctypes.c_int64, ctypes.c_int]
    lib.bode_ekld_score_dev.restype = ctypes.c_int
    lib.bode_ekld_score_dev.argtypes = [c_dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_dp, ctypes.c_int64,
                                        ctypes.c_int64, ctypes.c_int, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]
    lib.bode_ekld_rescore_dev.restype = ctypes.c_int
    lib.bode_ekld_rescore_dev.argtypes = [c_dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_dp, ctypes.c_int64,
                                          ctypes.c_int64, ctypes.c_int, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]
    lib.bode_ekld_score_dev_timed.restype = ctypes.c_int
    lib.bode_ekld_score_dev_timed.argtypes = list(lib.bode_ekld_score_dev.argtypes) + [c_dp]
    lib.bode_ekld_score_dev_events.restype = ctypes.c_int
    lib.bode_ekld_score_dev_events.argtypes = list(lib.bode_ekld_score_dev.argtypes) + [c_dp, c_dp]
    lib.bode_quad_scratch_bytes.restype = ctypes.c_size_t
    lib.bode_quad_scratch_bytes.argtypes = [ctypes.c_int] * 3
    lib.bode_quad_rowmean_dev.restype = ctypes.c_int
    lib.bode_quad_rowmean_dev.argtypes = [c_dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_dp, ctypes.c_int64, c_dp, c_dp,
                                          ctypes.c_int, c_dp, c_dp, c_dp]
    lib.bode_ekld_prepare_quad_dev.restype = ctypes.c_int
    lib.bode_ekld_prepare_quad_dev.argtypes = [c_dp, c_dp, c_dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_double, ctypes.c_double, c_dp, c_dp, ctypes.c_int, c_dp,
                                               ctypes.c_size_t, c_dp, c_dp, c_dp]
    lib.bode_ekld_prepare_quad_rm_dev.restype = ctypes.c_int
    lib.bode_ekld_prepare_quad_rm_dev.argtypes = [c_dp, c_dp, c_dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                  ctypes.c_double, ctypes.c_double, c_dp, c_dp, ctypes.c_int, c_dp, c_dp,
                                                  ctypes.c_size_t, c_dp, c_dp, c_dp]
    lib.bode_ekld_score_quad_dev.restype = ctypes.c_int
    lib.bode_ekld_score_quad_dev.argtypes = [c_dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_dp, ctypes.c_int, c_dp,
                                             ctypes.c_int64, ctypes.c_int64, ctypes.c_int, c_dp, c_dp, ctypes.c_int,
                                             c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]
    lib.bode_us_variance_dev.restype = ctypes.c_int
    lib.bode_us_variance_dev.argtypes = [c_dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_dp, ctypes.c_int64,
                                         ctypes.c_int64, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]
    lib.bode_ekld_score_host.restype = ctypes.c_int
    lib.bode_ekld_score_host.argtypes = [c_dp, c_dp, c_dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_double, ctypes.c_double, c_dp, ctypes.c_int64, ctypes.c_int,
                                         c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]
    lib.bode_gp_loglik_host.restype = ctypes.c_int
    lib.bode_gp_loglik_host.argtypes = [c_dp, c_dp, c_dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_double, ctypes.c_double, c_dp]
    lib.bode_gp_loglik_dev.restype = ctypes.c_int
    lib.bode_gp_loglik_dev.argtypes = [c_dp, c_dp, c_dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_double, ctypes.c_double, c_dp, ctypes.c_size_t, c_dp, c_dp, c_dp]
    lib.bode_ekld_plan_describe.restype = ctypes.c_int
    lib.bode_ekld_plan_describe.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int,
                                            ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
    lib.bode_fp64_peak.restype = ctypes.c_int
    lib.bode_fp64_peak.argtypes = [c_dp, c_dp]
    lib.bode_tuning_reload.restype = ctypes.c_int
    lib.bode_tuning_reload.argtypes = []
    lib.bode_ctx_create.restype = ctypes.c_int
    lib.bode_ctx_create.argtypes = [ctypes.POINTER(ctypes.c_void_p)]
    lib.bode_ctx_destroy.restype = ctypes.c_int
    lib.bode_ctx_destroy.argtypes = [c_dp]
    lib.bode_ctx_alloc_count.restype = ctypes.c_longlong
    lib.bode_ctx_alloc_count.argtypes = [c_dp]
    lib.bode_ekld_score_host_ctx.restype = ctypes.c_int
    lib.bode_ekld_score_host_ctx.argtypes = [c_dp] + list(lib.bode_ekld_score_host.argtypes)
    lib.bode_gp_loglik_host_ctx.restype = ctypes.c_int
    lib.bode_gp_loglik_host_ctx.argtypes = [c_dp] + list(lib.bode_gp_loglik_host.argtypes)
    lib.bode_nccl_unique_id.restype = ctypes.c_int
    lib.bode_nccl_unique_id.argtypes = [c_dp]
    lib.bode_nccl_comm_create.restype = ctypes.c_int
    lib.bode_nccl_comm_create.argtypes = [c_dp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
    lib.bode_nccl_comm_destroy.restype = ctypes.c_int
    lib.bode_nccl_comm_destroy.argtypes = [c_dp]
    lib.bode_ekld_argmax_nccl.restype = ctypes.c_int
    lib.bode_ekld_argmax_nccl.argtypes = [c_dp, c_dp, c_dp, c_dp]
    lib.bode_argmax_combine.restype = ctypes.c_int
    lib.bode_argmax_combine.argtypes = [c_dp, ctypes.c_int, c_dp, c_dp]
    lib.bode_xchg_create.restype = ctypes.c_int
    lib.bode_xchg_create.argtypes = [ctypes.c_int, ctypes.c_int, c_dp, c_dp]
    lib.bode_xchg_connect.restype = ctypes.c_int
    lib.bode_xchg_connect.argtypes = [c_dp, c_dp]
    lib.bode_xchg_destroy.restype = ctypes.c_int
    lib.bode_xchg_destroy.argtypes = [c_dp]
    lib.bode_ekld_argmax_p2p.restype = ctypes.c_int
    lib.bode_ekld_argmax_p2p.argtypes = [c_dp, c_dp, c_dp, c_dp]
    _lib = lib
    return lib


def check(status):
    if status != 0:
        lib = load()
        detail = lib.bode_strerror(status).decode()
        if status in (-4, -6):
            detail += " (" + lib.bode_last_cuda_error().decode() + ")"
        raise BodeError(status, detail)


def _np_ptr(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else None


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def reload_tuning():
    """Re-read the BODE_* tuning environment variables (they are otherwise read once, when the library is loaded)."""
    check(load().bode_tuning_reload())


class HostContext(object):
    """Explicit reusable context for the host-buffer entry points (``bode_ctx_create`` / ``bode_ctx_destroy``)."""

    def __init__(self):
        self.lib = load()
        self.handle = ctypes.c_void_p()
        check(self.lib.bode_ctx_create(ctypes.byref(self.handle)))

    def alloc_count(self):
        return int(self.lib.bode_ctx_alloc_count(self.handle))

    def close(self):
        if self.handle:
            self.lib.bode_ctx_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def default_ctx_alloc_count():
    return int(load().bode_ctx_alloc_count(None))


def argmax_combine(records):
    """np.argmax-style winner among packed 16-byte (float64 value, int64 index) records (host bytes / int64 array)."""
    lib = load()
    buf = np.ascontiguousarray(records).view(np.uint8).reshape(-1)
    assert buf.size % 16 == 0
    best = ctypes.c_double(0.0)
    idx = ctypes.c_int64(-1)
    check(lib.bode_argmax_combine(_np_ptr(buf), buf.size // 16, ctypes.byref(best), ctypes.byref(idx)))
    return best.value, idx.value


def ekld_score_host(X, Y, hyp, X_design, nugget_var, jitter=1e-8, form=BODE_FORM_REFERENCE, want_var=True,
                    want_kld=False, raise_not_pd=True, ctx=None):
    """Host-buffer call (``bode_ekld_score_host``): NumPy in, NumPy out; copies both ways inside the call."""
    lib = load()
    X = _f64(X); n, d = X.shape
    Y = _f64(Y).reshape(-1)
    hyp = _f64(hyp); S, ndim = hyp.shape
    Xd = _f64(np.atleast_2d(X_design)); Nd = Xd.shape[0]
    assert Xd.shape[1] == d and Y.shape[0] == n
    score = np.empty(Nd)
    var = np.empty(Nd) if want_var else None
    kld = np.empty((S, Nd)) if want_kld else None
    best = np.zeros(1, dtype=np.int64)
    ms = np.empty(2)
    info = np.zeros(S, dtype=np.int32)
    st = lib.bode_ekld_score_host_ctx(ctx.handle if ctx is not None else None, _np_ptr(X), _np_ptr(Y), _np_ptr(hyp), n, d, S,
                                      ndim, float(nugget_var), float(jitter), _np_ptr(Xd), Nd, int(form), _np_ptr(score),
                                      _np_ptr(var), _np_ptr(kld), _np_ptr(best), _np_ptr(ms), _np_ptr(info))
    if st != 0 and not (st == -5 and not raise_not_pd):
        check(st)
    return dict(score=score, var=var, kld=kld, argmax=int(best[0]), mu_Q=float(ms[0]), sigma_Q=float(ms[1]), info=info)


def gp_loglik_host(X, Y, hyp, nugget_var, jitter=1e-8):
    lib = load()
    X = _f64(X); n, d = X.shape
    Y = _f64(Y).reshape(-1)
    hyp = _f64(np.atleast_2d(hyp)); S, ndim = hyp.shape
    out = np.empty(S)
    check(lib.bode_gp_loglik_host(_np_ptr(X), _np_ptr(Y), _np_ptr(hyp), n, d, S, ndim, float(nugget_var), float(jitter),
                                  _np_ptr(out)))
    return out


def plan_describe(n, d, S, Nd, sm_count=148, smem_optin=232448):
    """Launch plan (dict) of the scoring kernel for a shape; host-only, works without a GPU."""
    import json
    lib = load()
    buf = ctypes.create_string_buffer(1 << 16)
    check(lib.bode_ekld_plan_describe(int(n), int(d), int(S), int(Nd), int(sm_count), int(smem_optin), buf, len(buf)))
    return json.loads(buf.value.decode())


def fp64_peak():
    lib = load()
    a = ctypes.c_double(0.0)
    b = ctypes.c_double(0.0)
    check(lib.bode_fp64_peak(ctypes.byref(a), ctypes.byref(b)))
    return dict(dmma_tflops=a.value, dfma_tflops=b.value)
