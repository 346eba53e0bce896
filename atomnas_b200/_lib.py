"""ctypes binding of the C ABI declared in include/atomnas_b200.h.

The library is built in-tree (atomnas_b200/libatomnas_b200.so, see _build.py).  There is no
Python or PyTorch fallback: if the library is missing, `lib()` raises, and every entry point
returns an error on a machine without an sm_100 GPU.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libatomnas_b200.so")

c_int, c_ll, c_vp, c_fp = ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p

ACT = {None: 0, "none": 0, "nn.ReLU": 1, "relu": 1, "nn.ReLU6": 2, "relu6": 2, "nn.Swish": 3,
       "swish": 3}


class PwGemmArgs(ctypes.Structure):
    _fields_ = [
        ("x1", c_vp), ("ldx1", c_ll), ("x2", c_vp), ("ldx2", c_ll), ("w", c_vp), ("ldw", c_ll),
        ("M", c_int), ("N", c_int), ("K", c_int), ("a_mode", c_int), ("a_act", c_int),
        ("a1", c_fp), ("a2", c_fp), ("a3", c_fp),
        ("epi_mode", c_int), ("y", c_vp), ("ldy", c_ll), ("y_is_f32", c_int), ("bias", c_fp),
        ("stats_part", c_fp), ("r", c_vp), ("ldr", c_ll),
        ("r_scale", c_fp), ("r_shift", c_fp), ("r_mean", c_fp), ("r_rstd", c_fp),
        ("r_act", c_int),
    ]


class PwWgradArgs(ctypes.Structure):
    _fields_ = [
        ("p1", c_vp), ("ldp1", c_ll), ("p2", c_vp), ("ldp2", c_ll), ("q", c_vp), ("ldq", c_ll),
        ("Mrows", c_int), ("Cbig", c_int), ("Csmall", c_int), ("p_mode", c_int), ("p_act", c_int),
        ("a1", c_fp), ("a2", c_fp), ("a3", c_fp), ("partials", c_fp), ("splits", c_int),
    ]


# name -> (restype, argtypes); must list every symbol of include/atomnas_b200.h
SIGNATURES = {
    "an_last_error": (ctypes.c_char_p, []),
    "an_version": (c_int, []),
    "an_sm_count": (c_int, []),
    "an_pw_gemm": (c_int, [ctypes.POINTER(PwGemmArgs), c_vp]),
    "an_pw_gemm_stats_rows": (c_int, [c_int, c_int]),
    "an_pw_wgrad_splits": (c_int, [c_int, c_int, c_int]),
    "an_pw_wgrad": (c_int, [ctypes.POINTER(PwWgradArgs), c_vp]),
    "an_wgrad_reduce": (c_int, [c_fp, c_int, c_int, c_int, c_fp, c_ll, c_ll, c_vp]),
}

_lib = None


def lib():
    """Load (once) and return the shared library; raise loudly if it was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "atomnas_b200: %s is missing -- run `python -m atomnas_b200._build` "
                "(there is no CPU / PyTorch fallback for the hot path)" % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


class AtomnasError(RuntimeError):
    pass


def check(code, what=""):
    if code != 0:
        msg = lib().an_last_error()
        raise AtomnasError("%s failed (%d): %s" % (what, code, msg.decode() if msg else "?"))


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
