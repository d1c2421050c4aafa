"""ctypes front-end of oracle/smc_oracle.c (TEST INFRASTRUCTURE, not product code).

Every function takes/returns numpy arrays and forwards to the C restatement; the reference lines
each one follows are cited in smc_oracle.c.  ``build()`` compiles the library with gcc (the same
recipe as oracle/Makefile) when it is missing.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsmc_oracle.so")
_lib = None

OK, NAN_INPUT, DEGENERATE = 0, 1, 2


def build(force=False):
    src = os.path.join(_HERE, "smc_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        f32 = ctypes.c_float
        for name in ("aesmc_oracle_np_expf", "aesmc_oracle_np_logf", "aesmc_oracle_log1pf"):
            getattr(L, name).restype = f32
            getattr(L, name).argtypes = [f32]
        L.aesmc_oracle_pairwise_sum.restype = f32
        L.aesmc_oracle_sample_ancestral_index.restype = ctypes.c_int
        L.aesmc_oracle_digitize.restype = ctypes.c_int
        L.aesmc_oracle_core_pass.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


_i64 = ctypes.c_int64


def _map_scalar(fn, x):
    x = _f32(x)
    out = np.empty_like(x)
    f = getattr(lib(), fn)
    flat_in, flat_out = x.ravel(), out.ravel()
    for i in range(flat_in.size):
        flat_out[i] = f(ctypes.c_float(float(flat_in[i])))
    return out


def np_expf(x):
    return _map_scalar("aesmc_oracle_np_expf", x)


def np_logf(x):
    return _map_scalar("aesmc_oracle_np_logf", x)


def log1pf(x):
    return _map_scalar("aesmc_oracle_log1pf", x)


def pairwise_sum_rows(a):
    a = _f32(a)
    B, K = a.shape
    out = np.empty(B, np.float32)
    f = lib().aesmc_oracle_pairwise_sum
    for b in range(B):
        out[b] = f(_p(a[b]), _i64(K))
    return out


def logsumexp_rows(a):
    a = _f32(a)
    B, K = a.shape
    out = np.empty(B, np.float32)
    lib().aesmc_oracle_logsumexp_rows(_p(a), _i64(B), _i64(K), _p(out))
    return out


def log_weight(a, b=None, c=None):
    a = _f32(a)
    b = None if b is None else _f32(b)
    c = None if c is None else _f32(c)
    out = np.empty_like(a)
    lib().aesmc_oracle_log_weight(_p(a), _p(b), _p(c), _i64(a.size), _p(out))
    return out


def normalized_weights(lw, lse_inject=None):
    lw = _f32(lw)
    B, K = lw.shape
    w = np.empty_like(lw)
    lse = np.empty(B, np.float32)
    inj = None if lse_inject is None else _f32(lse_inject).reshape(B)
    lib().aesmc_oracle_normalized_weights(_p(lw), _i64(B), _i64(K), _p(inj), _p(w), _p(lse))
    return w, lse


def cdf(w):
    w = _f32(w)
    B, K = w.shape
    out = np.empty_like(w)
    lib().aesmc_oracle_cdf(_p(w), _i64(B), _i64(K), _p(out))
    return out


def digitize(u, cdf_):
    cdf_ = _f32(cdf_)
    B, K = cdf_.shape
    u = np.ascontiguousarray(u, dtype=np.float64).reshape(B)
    idx = np.empty((B, K), np.int64)
    st = lib().aesmc_oracle_digitize(_p(u), _p(cdf_), _i64(B), _i64(K), _p(idx))
    return idx, st


def sample_ancestral_index(lw, u, lse_inject=None, return_parts=False):
    """inference.py:234-269 with the per-row uniforms injected.  Returns (idx int64 [B,K], status)
    or, with return_parts, (idx, status, lse, w, cdf)."""
    lw = _f32(lw)
    B, K = lw.shape
    u = np.ascontiguousarray(u, dtype=np.float64).reshape(B)
    idx = np.empty((B, K), np.int64)
    lse = np.empty(B, np.float32)
    w = np.empty_like(lw) if return_parts else None
    c = np.empty_like(lw) if return_parts else None
    inj = None if lse_inject is None else _f32(lse_inject).reshape(B)
    st = lib().aesmc_oracle_sample_ancestral_index(_p(lw), _p(u), _i64(B), _i64(K), _p(inj), _p(idx),
                                                   _p(lse), _p(w), _p(c))
    if return_parts:
        return idx, st, lse, w, c
    return idx, st


def resample(x, idx):
    x = _f32(x)
    B, K = x.shape[:2]
    D = int(np.prod(x.shape[2:])) if x.ndim > 2 else 1
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    out = np.empty_like(x)
    lib().aesmc_oracle_resample(_p(x), _p(idx), _i64(B), _i64(K), _i64(D), _p(out))
    return out


def resample_bwd(g, idx):
    g = _f32(g)
    B, K = g.shape[:2]
    D = int(np.prod(g.shape[2:])) if g.ndim > 2 else 1
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    out = np.empty_like(g)
    lib().aesmc_oracle_resample_bwd(_p(g), _p(idx), _i64(B), _i64(K), _i64(D), _p(out))
    return out


def compose_index(prev, cur):
    prev = np.ascontiguousarray(prev, dtype=np.int64)
    cur = np.ascontiguousarray(cur, dtype=np.int64)
    B, K = prev.shape
    out = np.empty_like(prev)
    lib().aesmc_oracle_compose_index(_p(prev), _p(cur), _i64(B), _i64(K), _p(out))
    return out


def lse_f64(lw):
    lw = _f32(lw)
    B, K = lw.shape
    out = np.empty(B, np.float64)
    lib().aesmc_oracle_lse_f64(_p(lw), _i64(B), _i64(K), _p(out))
    return out


def log_ess_f64(lw):
    lw = _f32(lw)
    B, K = lw.shape
    out = np.empty(B, np.float64)
    lib().aesmc_oracle_log_ess_f64(_p(lw), _i64(B), _i64(K), _p(out))
    return out


def lse_bwd_f64(lw, lse, g):
    lw = _f32(lw)
    B, K = lw.shape
    lse = np.ascontiguousarray(lse, dtype=np.float64)
    g = _f32(g)
    out = np.empty((B, K), np.float64)
    lib().aesmc_oracle_lse_bwd_f64(_p(lw), _p(lse), _p(g), _i64(B), _i64(K), _p(out))
    return out


def core_pass(a, b, c, u, x, T):
    """T steps of the core (see aesmc_oracle_core_pass).  a,b,c: [R,B,K] ring; u: [T-1,B];
    x: [B,K,D] (modified in place).  Returns (lml [B] float64, idx_last [B,K], status)."""
    a, b, c = _f32(a), _f32(b), _f32(c)
    R, B, K = a.shape
    D = int(np.prod(x.shape[2:])) if x.ndim > 2 else 1
    u = np.ascontiguousarray(u, dtype=np.float64)
    assert x.dtype == np.float32 and x.flags.c_contiguous
    lml = np.empty(B, np.float64)
    idx = np.empty((B, K), np.int64)
    st = lib().aesmc_oracle_core_pass(_p(a), _p(b), _p(c), _i64(R), _p(u), _i64(T), _i64(B), _i64(K),
                                      _i64(D), _p(x), _p(lml), _p(idx))
    return lml, idx, st
