"""ctypes binding of libaesmc_b200.so (the C ABI declared in include/aesmc_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (plain nvcc, sm_100a).  There is no CPU
fallback: if the shared object is missing, or no CUDA device is present, every hot-path call raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# AESMC_B200_LIB: measure an alternative build of the same library (kernel experiments); default: the in-tree build
LIB_PATH = os.environ.get("AESMC_B200_LIB") or os.path.join(_HERE, "libaesmc_b200.so")

OK, ERR_BAD_ARG, ERR_LAUNCH, ERR_UNSUPPORTED = 0, 1, 2, 3
FLAG_NAN, FLAG_DEGENERATE, FLAG_INDEX_RANGE = 1, 2, 4
MODE_EXACT, MODE_FAST = 0, 1

_vp, _i64, _int = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int

# name -> argtypes; every function returns int unless listed in _RESTYPES
_PROTOTYPES = {
    "aesmc_smc_step_f32": [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _int, _vp],
    "aesmc_smc_step_ws_f32": [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _int, _vp, _i64, _vp],
    "aesmc_smc_step_workspace_bytes": [_i64, _i64],
    "aesmc_smc_step_lg_f32": [_vp, _vp, _vp, _vp, _vp, ctypes.c_float, ctypes.c_uint64, _vp, ctypes.c_uint64, _i64,
                              _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, _vp],
    "aesmc_smc_step_lg_dev_f32": [_vp, _vp, _vp, _vp, _vp, ctypes.c_float, ctypes.c_uint64, _vp, ctypes.c_uint64, _i64,
                                  _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, _vp],
    "aesmc_lgv_propose_f32": [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _int, ctypes.c_uint64, ctypes.c_uint64, _i64, _i64,
                              _vp, _vp, _vp],
    "aesmc_lg_step_bwd_f32": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp],
    "aesmc_resample_from_weights_f32": [_vp, _vp, _i64, _i64, _vp, _vp, _int, _vp],
    "aesmc_resample_from_cdf_f32": [_vp, _vp, _i64, _i64, _vp, _vp, _vp],
    "aesmc_is_accumulate_f32": [_vp, _vp, _vp, _vp, _vp, _i64, _int, _vp],
    "aesmc_logsumexp_f32": [_vp, _i64, _i64, _vp, _vp, _vp],
    "aesmc_logsumexp_f64": [_vp, _i64, _i64, _vp, _vp, _vp],
    "aesmc_step_bwd_f32": [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp],
    "aesmc_lognormexp_f32": [_vp, _i64, _i64, _vp, _int, _vp],
    "aesmc_gather_bytes": [_vp, _vp, _int, _i64, _i64, _i64, _vp, _vp, _vp],
    "aesmc_gather_bwd_f32": [_vp, _vp, _int, _i64, _i64, _i64, _vp, _int, _vp],
    "aesmc_gather_bwd_f64": [_vp, _vp, _int, _i64, _i64, _i64, _vp, _int, _vp],
    "aesmc_compose_index_i32": [_vp, _vp, _i64, _i64, _vp, _vp],
    "aesmc_iota_index_i32": [_i64, _i64, _vp, _vp],
    "aesmc_index_widen": [_vp, _vp, _i64, _vp],
    "aesmc_index_narrow": [_vp, _vp, _i64, _vp],
    "aesmc_normal_log_prob_f32": [_vp, _int, _vp, _int, ctypes.c_float, _vp, ctypes.c_float, ctypes.c_float,
                                  ctypes.c_float, _i64, _i64, _vp, _vp],
    "aesmc_normal_log_prob_bwd_f32": [_vp, _int, _vp, _int, ctypes.c_float, _vp, ctypes.c_float, _vp, _i64, _i64, _vp, _vp,
                                      _vp, _vp],
    "aesmc_selftest_expf": [_vp, _vp],
    "aesmc_debug_force_rare_paths": [_int],
    "aesmc_log_ess_f32": [_vp, _i64, _i64, _vp, _vp],
    "aesmc_log_ess_f64": [_vp, _i64, _i64, _vp, _vp],
    "aesmc_weighted_moments_f32": [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp],
    "aesmc_version": [],
    "aesmc_last_error_string": [],
    "aesmc_launch_count": [],
    "aesmc_max_particles_single_cta": [],
}
_RESTYPES = {
    "aesmc_last_error_string": ctypes.c_char_p,
    "aesmc_launch_count": _i64,
    "aesmc_max_particles_single_cta": _i64,
    "aesmc_smc_step_workspace_bytes": _i64,
}
EXPORTED_SYMBOLS = tuple(_PROTOTYPES)

_lib = None


class ExtensionMissingError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises ExtensionMissingError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ExtensionMissingError(
                "aesmc_b200: %s not found. Build it with `python -c \"import __graft_entry__ as g; g.build()\"` "
                "from the repo root. There is no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, argtypes in _PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, _int)
        _lib = lib
    return _lib


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("aesmc_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")


def last_error():
    msg = load().aesmc_last_error_string()
    return msg.decode() if msg else ""


def launch_count():
    return int(load().aesmc_launch_count())


def max_particles_single_cta():
    return int(load().aesmc_max_particles_single_cta())


def check(status):
    if status == OK:
        return
    msg = last_error()
    if status == ERR_BAD_ARG:
        raise ValueError(msg)
    raise RuntimeError(msg)


class DevicePointer(int):
    """A raw device address that remembers which CUDA device it lives on, so that call() can launch on
    that device's current stream whatever the process's current device is."""
    device_index = -1


def ptr(t):
    if t is None:
        return None
    p = DevicePointer(t.data_ptr())
    p.device_index = t.device.index if t.is_cuda else -1
    return p


def stream():
    return torch.cuda.current_stream().cuda_stream


def call(name, *args):
    """Invoke an exported function; maps error codes to exceptions.

    The kernels are enqueued on torch's current stream OF THE DEVICE THE POINTER ARGUMENTS LIVE ON (not of
    the process's current device): tensors on cuda:1 while cuda:0 is current launch on cuda:1, ordered after
    the torch ops that produced them.  All pointer arguments must share one device."""
    lib = load()
    dev = -1
    for a in args:
        if type(a) is DevicePointer and a.device_index >= 0:
            if dev < 0:
                dev = a.device_index
            elif a.device_index != dev:
                raise ValueError("%s: pointer arguments live on different CUDA devices (cuda:%d and cuda:%d)"
                                 % (name, dev, a.device_index))
    if dev < 0 or dev == torch.cuda.current_device():
        check(getattr(lib, name)(*args, torch.cuda.current_stream().cuda_stream))
        return
    with torch.cuda.device(dev):
        check(getattr(lib, name)(*args, torch.cuda.current_stream(dev).cuda_stream))
