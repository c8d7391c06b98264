"""ctypes binding of libsoftpool_b200.so (the C ABI in include/softpool_b200.h).

There is NO fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsoftpool_b200.so")
ABI_VERSION = 3

_p = ctypes.c_void_p
_i = ctypes.c_int
_SIGS = {
    # name: (restype, argtypes)
    "spk_abi_version": (_i, []),
    "spk_last_error": (ctypes.c_char_p, []),
    "sp_topk_f32": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _p]),
    "sp_argmax_i64": (_i, [_p, _i, _i, _i, _p, _p]),
    "sp_gather_fwd_f32": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "sp_gather_bwd_f32": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p]),
    "sp_cabins_fwd_f32": (_i, [_p, ctypes.c_longlong, _i, _i, _p, _p, _p]),
    "sp_cabins_bwd_f32": (_i, [_p, _p, ctypes.c_longlong, _i, _i, _p, _p]),
    "chamfer_fwd_workspace_bytes": (ctypes.c_size_t, [_i, _i, _i]),
    "chamfer_fwd_f32": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _p, _p, ctypes.c_size_t, _p]),
    "chamfer_bwd_f32": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p]),
    "chamfer_loss_f32": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "chamfer_fwd_loss_f32": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _p, ctypes.c_size_t, _p]),
    "chamfer_fwd_multi_f32": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, ctypes.c_size_t, _p]),
}
EXPORTS = tuple(_SIGS)

_lib = None


def lib():
    """Load (once) and return the ctypes handle; raises if the CUDA library is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "softpool_b200: %s not found -- build it with `python -m softpool_b200.build` "
                "(there is no CPU/PyTorch fallback)" % LIB_PATH)
        h = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(h, name)      # AttributeError here = ABI mismatch, loud by design
            fn.restype = res
            fn.argtypes = args
        v = h.spk_abi_version()
        if v != ABI_VERSION:
            raise RuntimeError("softpool_b200: ABI version %d, expected %d" % (v, ABI_VERSION))
        _lib = h
    return _lib


def check(code, what):
    if code != 0:
        msg = lib().spk_last_error().decode("utf-8", "replace")
        raise RuntimeError("softpool_b200.%s failed (code %d): %s" % (what, code, msg))


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_of(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def require_cuda(t, name, dtype):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("softpool_b200: %s must be a CUDA tensor (got %s); the B200 kernels are "
                           "the only implementation" % (name, t.device))
    if t.dtype != dtype:
        raise RuntimeError("softpool_b200: %s must be %s (got %s)" % (name, dtype, t.dtype))
