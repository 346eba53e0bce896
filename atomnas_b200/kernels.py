"""Thin Python wrappers over the C ABI (one function per entry point).

Tensors are only used as device-memory handles here: every wrapper passes raw pointers,
sizes and the current CUDA stream to the shared library.
"""
import torch

from . import _lib as L


def _c(t):
    assert t is None or t.is_cuda, "device tensors only"
    return L.ptr(t)


def pw_gemm(x1, w, y, M, N, K, *, ldx1=None, ldw=None, ldy=None, x2=None, ldx2=None, a_mode=0,
            a_act=0, a1=None, a2=None, a3=None, epi_mode=0, bias=None, stats_part=None, r=None,
            ldr=None, r_scale=None, r_shift=None, r_mean=None, r_rstd=None, r_act=0):
    """D[M,N] = Aeff[M,K] @ W[N,K]^T with fused prologue / epilogue (see atomnas_b200.h)."""
    a = L.PwGemmArgs()
    a.x1, a.ldx1 = _c(x1), ldx1 if ldx1 is not None else K
    a.x2, a.ldx2 = _c(x2), (ldx2 if ldx2 is not None else K)
    a.w, a.ldw = _c(w), ldw if ldw is not None else K
    a.M, a.N, a.K = M, N, K
    a.a_mode, a.a_act = a_mode, a_act
    a.a1, a.a2, a.a3 = _c(a1), _c(a2), _c(a3)
    a.epi_mode = epi_mode
    a.y, a.ldy = _c(y), ldy if ldy is not None else N
    a.y_is_f32 = 1 if y.dtype == torch.float32 else 0
    a.bias = _c(bias)
    a.stats_part = _c(stats_part)
    a.r, a.ldr = _c(r), (ldr if ldr is not None else N)
    a.r_scale, a.r_shift, a.r_mean, a.r_rstd = _c(r_scale), _c(r_shift), _c(r_mean), _c(r_rstd)
    a.r_act = r_act
    L.check(L.lib().an_pw_gemm(a, L.stream_ptr()), "an_pw_gemm")


def pw_gemm_stats_rows(M, N):
    return L.lib().an_pw_gemm_stats_rows(M, N)


def pw_wgrad_splits(Mrows, Cbig, Csmall):
    return L.lib().an_pw_wgrad_splits(Mrows, Cbig, Csmall)


def pw_wgrad(p1, q, partials, Mrows, Cbig, Csmall, *, ldp1=None, ldq=None, p2=None, ldp2=None,
             p_mode=0, p_act=0, a1=None, a2=None, a3=None):
    a = L.PwWgradArgs()
    a.p1, a.ldp1 = _c(p1), ldp1 if ldp1 is not None else Cbig
    a.p2, a.ldp2 = _c(p2), (ldp2 if ldp2 is not None else Cbig)
    a.q, a.ldq = _c(q), ldq if ldq is not None else Csmall
    a.Mrows, a.Cbig, a.Csmall = Mrows, Cbig, Csmall
    a.p_mode, a.p_act = p_mode, p_act
    a.a1, a.a2, a.a3 = _c(a1), _c(a2), _c(a3)
    a.partials = _c(partials)
    a.splits = pw_wgrad_splits(Mrows, Cbig, Csmall)
    L.check(L.lib().an_pw_wgrad(a, L.stream_ptr()), "an_pw_wgrad")
    return a.splits


def wgrad_reduce(partials, splits, Cbig, Csmall, dst, sb, ss):
    L.check(L.lib().an_wgrad_reduce(_c(partials), splits, Cbig, Csmall, _c(dst), sb, ss,
                                    L.stream_ptr()), "an_wgrad_reduce")
