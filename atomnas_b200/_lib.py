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


class DwFwdArgs(ctypes.Structure):
    _fields_ = [
        ("x", c_vp), ("ldx", c_ll), ("y", c_vp), ("ldy", c_ll), ("w", c_fp), ("scale", c_fp),
        ("shift", c_fp), ("stats_part", c_fp), ("act", c_int),
        ("B", c_int), ("H", c_int), ("W", c_int), ("C", c_int), ("k", c_int), ("stride", c_int),
    ]


class DwBwdArgs(ctypes.Structure):
    _fields_ = [
        ("g", c_vp), ("ldg", c_ll), ("draw", c_vp), ("ldd", c_ll), ("eraw", c_vp), ("lde", c_ll),
        ("ge", c_vp), ("ldge", c_ll), ("w", c_fp), ("ca", c_fp), ("cb", c_fp), ("cc", c_fp),
        ("e_scale", c_fp), ("e_shift", c_fp), ("e_mean", c_fp), ("e_rstd", c_fp),
        ("stats_part", c_fp), ("act", c_int),
        ("B", c_int), ("H", c_int), ("W", c_int), ("C", c_int), ("k", c_int), ("stride", c_int),
    ]


class BnSeg(ctypes.Structure):
    _fields_ = [("part", c_fp), ("rows", c_int), ("c0", c_int), ("C", c_int)]


class WgSeg(ctypes.Structure):
    _fields_ = [("row0", c_int), ("nrows", c_int), ("dst", c_fp), ("dst_sb", c_ll), ("dst_ss", c_ll)]


c_f, c_d, c_ull = ctypes.c_float, ctypes.c_double, ctypes.c_ulonglong

# name -> (restype, argtypes); must list every symbol of include/atomnas_b200.h
SIGNATURES = {
    "an_last_error": (ctypes.c_char_p, []),
    "an_version": (c_int, []),
    "an_sm_count": (c_int, []),
    "an_launch_count": (c_ll, []),
    "an_pw_gemm": (c_int, [ctypes.POINTER(PwGemmArgs), c_vp]),
    "an_pw_gemm_stats_rows": (c_int, [c_int, c_int]),
    "an_pw_wgrad_splits": (c_int, [c_int, c_int, c_int]),
    "an_pw_wgrad": (c_int, [ctypes.POINTER(PwWgradArgs), c_vp]),
    "an_wgrad_reduce": (c_int, [c_fp, c_int, c_int, c_int, c_int, c_int, c_int, c_fp, c_ll, c_ll,
                                c_vp]),
    "an_bn_eval_coeffs": (c_int, [c_fp, c_fp, c_fp, c_fp, c_f, c_int, c_fp, c_fp, c_vp]),
    "an_act_bwd_reduce": (c_int, [c_vp, c_ll, c_vp, c_ll, c_fp, c_fp, c_fp, c_fp, c_int, c_vp, c_ll,
                                  c_ll, c_int, c_fp, c_vp]),
    "an_sum_rows": (c_int, [c_fp, c_int, c_ll, c_ll, c_fp, c_vp]),
    "an_dw_fwd": (c_int, [ctypes.POINTER(DwFwdArgs), c_vp]),
    "an_dw_fwd_stats_rows": (c_int, [c_int] * 7),
    "an_dw_bwd_data": (c_int, [ctypes.POINTER(DwBwdArgs), c_vp]),
    "an_dw_bwd_data_stats_rows": (c_int, [c_int] * 7),
    "an_dw_bwd_weight": (c_int, [ctypes.POINTER(DwBwdArgs), c_fp, c_fp, c_vp]),
    "an_dw_bwd_weight_rows": (c_int, [c_int] * 7),
    "an_dw_on_tensor_cores": (c_int, [c_int] * 5),
    "an_bn_finalize": (c_int, [c_fp, c_int, c_int, c_d, c_fp, c_fp, c_fp, c_fp, c_f, c_f, c_fp,
                               c_fp, c_fp, c_fp, c_vp]),
    "an_bn_bwd_finalize": (c_int, [c_fp, c_int, c_int, c_d, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp,
                                   c_fp, c_fp, c_int, c_vp]),
    "an_affine_act": (c_int, [c_vp, c_ll, c_vp, c_ll, c_fp, c_fp, c_fp, c_int, c_vp, c_ll, c_ll,
                              c_int, c_vp]),
    "an_col_reduce_rows": (c_int, [c_ll, c_int]),
    "an_col_reduce": (c_int, [c_vp, c_ll, c_vp, c_ll, c_fp, c_fp, c_ll, c_int, c_fp, c_vp]),
    "an_stem_fwd": (c_int, [c_fp, c_fp, c_vp, c_ll, c_fp, c_int, c_int, c_int, c_int, c_vp]),
    "an_stem_fwd_stats_rows": (c_int, [c_int, c_int, c_int]),
    "an_stem_wgrad": (c_int, [c_fp, c_vp, c_ll, c_vp, c_ll, c_fp, c_fp, c_fp, c_fp, c_fp, c_int,
                              c_int, c_int, c_int, c_vp]),
    "an_stem_wgrad_rows": (c_int, [c_int, c_int, c_int]),
    "an_pool_fwd": (c_int, [c_vp, c_ll, c_fp, c_fp, c_int, c_vp, c_ll, c_int, c_int, c_int, c_f,
                            c_ull, c_vp, c_vp]),
    "an_pool_bwd": (c_int, [c_vp, c_ll, c_vp, c_ll, c_fp, c_fp, c_fp, c_fp, c_int, c_vp, c_ll,
                            c_fp, c_int, c_int, c_int, c_f, c_ull, c_vp, c_vp]),
    "an_pool_bwd_stats_rows": (c_int, [c_int]),
    "an_ce_loss": (c_int, [c_fp, c_ll, c_vp, c_int, c_int, c_f, c_f, c_fp, c_vp, c_ll, c_int,
                           c_vp]),
    "an_f32_to_bf16": (c_int, [c_fp, c_ll, c_vp, c_ll, c_ll, c_int, c_int, c_vp]),
    "an_rmsprop_ema_step": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_ll, c_f, c_f, c_f,
                                    c_f, c_int, c_f, c_f, c_f, c_f, c_fp, c_vp]),
    "an_ema_update": (c_int, [c_fp, c_fp, c_ll, c_f, c_fp, c_vp]),
    "an_reg_losses": (c_int, [c_fp, c_fp, c_fp, c_ll, c_fp, c_vp]),
    "an_refresh_weights": (c_int, [c_vp, c_int, c_ll, c_fp, c_vp, c_fp, c_vp]),
    "an_gather_rows": (c_int, [c_fp, c_vp, c_ll, c_ll, c_fp, c_vp]),
    "an_bn_finalize_multi": (c_int, [ctypes.POINTER(BnSeg), c_int, c_d, c_fp, c_fp, c_fp, c_fp, c_f, c_f, c_fp, c_fp,
                                     c_fp, c_fp, c_vp]),
    "an_bn_bwd_finalize_multi": (c_int, [ctypes.POINTER(BnSeg), c_int, c_d, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp,
                                         c_fp, c_int, c_vp]),
    "an_wgrad_reduce_multi": (c_int, [c_fp, c_int, c_int, c_int, c_int, ctypes.POINTER(WgSeg), c_int, c_vp]),
    "an_se_pool": (c_int, [c_vp, c_ll, c_vp, c_ll, c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_fp, c_vp]),
    "an_se_gate": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, c_vp, c_int, c_int, c_int, c_int, c_int, c_fp, c_fp, c_vp]),
    "an_se_scale": (c_int, [c_vp, c_ll, c_fp, c_fp, c_int, c_fp, c_int, c_int, c_int, c_vp, c_ll, c_vp]),
    "an_se_gate_bwd": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int,
                               c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_vp]),
    "an_se_bwd_apply_rows": (c_int, [c_ll, c_int]),
    "an_se_bwd_apply": (c_int, [c_vp, c_ll, c_vp, c_ll, c_fp, c_fp, c_fp, c_fp, c_int, c_fp, c_fp, c_int, c_int,
                                c_int, c_vp, c_ll, c_fp, c_vp]),
}

_lib = None
_profiler = None
TAG = [""]            # set by the engine at stage boundaries; only read while profiling


class _Profiler(object):
    """Brackets every kernel-launching entry point with CUDA events on the launching stream
    (bench.py roofline measurement; serialises nothing, but adds two event records per call)."""

    def __init__(self, handle):
        self._h = handle
        self.records = []

    def __getattr__(self, name):
        fn = getattr(self._h, name)
        if not name.startswith("an_") or fn.restype is not c_int or name.endswith(("_rows", "_splits")) \
                or name in ("an_version", "an_sm_count"):
            return fn

        def timed(*args):
            s = torch.cuda.Event(enable_timing=True)
            e = torch.cuda.Event(enable_timing=True)
            s.record()
            r = fn(*args)
            e.record()
            self.records.append((name, TAG[0], s, e))
            return r

        return timed


def start_profile():
    global _profiler
    lib()
    _profiler = _Profiler(_lib)


def stop_profile():
    """Return {(entry point, tag): [calls, total ms]} since start_profile()."""
    global _profiler
    torch.cuda.synchronize()
    out = {}
    for name, tag, s, e in _profiler.records:
        k = (name, tag)
        c = out.setdefault(k, [0, 0.0])
        c[0] += 1
        c[1] += s.elapsed_time(e)
    _profiler = None
    return out


def lib():
    """Load (once) and return the shared library; raise loudly if it was not built."""
    global _lib
    if _profiler is not None:
        return _profiler
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
