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


def wgrad_reduce(partials, splits, Cbig, Csmall, dst, sb, ss, row0=0, nrows=None, ncols=None):
    nrows = Cbig - row0 if nrows is None else nrows
    ncols = Csmall if ncols is None else ncols
    L.check(L.lib().an_wgrad_reduce(_c(partials), splits, Cbig, Csmall, row0, nrows, ncols, _c(dst),
                                    sb, ss, L.stream_ptr()), "an_wgrad_reduce")


# ------------------------------------------------------------------------------------------------
# depthwise
# ------------------------------------------------------------------------------------------------
def dw_out_hw(H, W, k, stride):
    p = (k - 1) // 2
    return (H + 2 * p - k) // stride + 1, (W + 2 * p - k) // stride + 1


def dw_fwd_stats_rows(B, C, Ho, Wo, k, stride, fused=True):
    return L.lib().an_dw_fwd_stats_rows(B, C, Ho, Wo, k, stride, 1 if fused else 0)


def dw_fwd(x, y, w, B, H, W, C, k, stride, *, ldx, ldy, scale=None, shift=None, act=0,
           stats_part=None):
    a = L.DwFwdArgs()
    a.x, a.ldx, a.y, a.ldy = _c(x), ldx, _c(y), ldy
    a.w, a.scale, a.shift, a.stats_part = _c(w), _c(scale), _c(shift), _c(stats_part)
    a.act = act
    a.B, a.H, a.W, a.C, a.k, a.stride = B, H, W, C, k, stride
    L.check(L.lib().an_dw_fwd(a, L.stream_ptr()), "an_dw_fwd")


def _dw_bwd_args(g, draw, eraw, ge, w, ca, cb, cc, e_scale, e_shift, e_mean, e_rstd, stats_part,
                 act, B, H, W, C, k, stride, ldg, ldd, lde, ldge):
    a = L.DwBwdArgs()
    a.g, a.ldg, a.draw, a.ldd, a.eraw, a.lde = _c(g), ldg, _c(draw), ldd, _c(eraw), lde
    a.ge, a.ldge = _c(ge), ldge
    a.w, a.ca, a.cb, a.cc = _c(w), _c(ca), _c(cb), _c(cc)
    a.e_scale, a.e_shift, a.e_mean, a.e_rstd = _c(e_scale), _c(e_shift), _c(e_mean), _c(e_rstd)
    a.stats_part = _c(stats_part)
    a.act = act
    a.B, a.H, a.W, a.C, a.k, a.stride = B, H, W, C, k, stride
    return a


def dw_bwd_data_stats_rows(B, C, H, W, k, stride, fused=True):
    return L.lib().an_dw_bwd_data_stats_rows(B, C, H, W, k, stride, 1 if fused else 0)


def dw_bwd_data(g, draw, eraw, ge, w, ca, cb, cc, e_scale, e_shift, e_mean, e_rstd, stats_part, act,
                B, H, W, C, k, stride, *, ldg, ldd, lde, ldge):
    a = _dw_bwd_args(g, draw, eraw, ge, w, ca, cb, cc, e_scale, e_shift, e_mean, e_rstd, stats_part,
                     act, B, H, W, C, k, stride, ldg, ldd, lde, ldge)
    L.check(L.lib().an_dw_bwd_data(a, L.stream_ptr()), "an_dw_bwd_data")


def dw_on_tensor_cores(which, k, stride, B, W):
    """Dispatch rule of the depthwise kernels (host logic only, no GPU needed): which = 0 forward /
    data gradient, 1 weight gradient."""
    return bool(L.lib().an_dw_on_tensor_cores(which, k, stride, B, W))


def dw_bwd_weight_rows(B, C, H, W, k, stride, fused=True):
    return L.lib().an_dw_bwd_weight_rows(B, C, H, W, k, stride, 1 if fused else 0)


def dw_bwd_weight(g, draw, eraw, ca, cb, cc, e_scale, e_shift, act, partials, dw, B, H, W, C, k,
                  stride, *, ldg, ldd, lde):
    a = _dw_bwd_args(g, draw, eraw, None, None, ca, cb, cc, e_scale, e_shift, None, None, None, act,
                     B, H, W, C, k, stride, ldg, ldd, lde, 8)
    L.check(L.lib().an_dw_bwd_weight(a, _c(partials), _c(dw), L.stream_ptr()), "an_dw_bwd_weight")


# ------------------------------------------------------------------------------------------------
# batch norm pieces
# ------------------------------------------------------------------------------------------------
def bn_finalize(part, rows, C, count, gamma, beta, rm, rv, momentum, eps, mean, rstd, scale, shift):
    L.check(L.lib().an_bn_finalize(_c(part), rows, C, float(count), _c(gamma), _c(beta), _c(rm),
                                   _c(rv), momentum, eps, _c(mean), _c(rstd), _c(scale), _c(shift),
                                   L.stream_ptr()), "an_bn_finalize")


def bn_bwd_finalize(part, rows, C, count, gamma, mean, rstd, dgamma, dbeta, ca, cb, cc, raw_second=False):
    L.check(L.lib().an_bn_bwd_finalize(_c(part), rows, C, float(count), _c(gamma), _c(mean),
                                       _c(rstd), _c(dgamma), _c(dbeta), _c(ca), _c(cb), _c(cc),
                                       1 if raw_second else 0, L.stream_ptr()), "an_bn_bwd_finalize")


def affine_act(x1, out, M, C, *, ld1, ldo, x2=None, ld2=0, a1=None, a2=None, a3=None, act=0):
    L.check(L.lib().an_affine_act(_c(x1), ld1, _c(x2), ld2, _c(a1), _c(a2), _c(a3), act, _c(out),
                                  ldo, M, C, L.stream_ptr()), "an_affine_act")


def col_reduce_rows(M, C):
    return L.lib().an_col_reduce_rows(M, C)


def col_reduce(g, part, M, C, *, ldg, x=None, ldx=0, mean=None, rstd=None):
    L.check(L.lib().an_col_reduce(_c(g), ldg, _c(x), ldx, _c(mean), _c(rstd), M, C, _c(part),
                                  L.stream_ptr()), "an_col_reduce")


# ------------------------------------------------------------------------------------------------
# stem / head / loss
# ------------------------------------------------------------------------------------------------
def stem_fwd_stats_rows(B, Ho, Wo):
    return L.lib().an_stem_fwd_stats_rows(B, Ho, Wo)


def stem_fwd(x, w, y, part, B, H, W, C, *, ldy):
    L.check(L.lib().an_stem_fwd(_c(x), _c(w), _c(y), ldy, _c(part), B, H, W, C, L.stream_ptr()),
            "an_stem_fwd")


def stem_wgrad_rows(B, Ho, Wo):
    return L.lib().an_stem_wgrad_rows(B, Ho, Wo)


def stem_wgrad(x, g, sraw, ca, cb, cc, partials, dw, B, H, W, C, *, ldg, lds):
    L.check(L.lib().an_stem_wgrad(_c(x), _c(g), ldg, _c(sraw), lds, _c(ca), _c(cb), _c(cc),
                                  _c(partials), _c(dw), B, H, W, C, L.stream_ptr()), "an_stem_wgrad")


def pool_fwd(h, scale, shift, act, pooled, B, HW, C, *, ldh, ldp, drop_p=0.0, seed=0, seed_dev=None):
    L.check(L.lib().an_pool_fwd(_c(h), ldh, _c(scale), _c(shift), act, _c(pooled), ldp, B, HW, C,
                                drop_p, seed, _c(seed_dev), L.stream_ptr()), "an_pool_fwd")


def pool_bwd_stats_rows(B):
    return L.lib().an_pool_bwd_stats_rows(B)


def pool_bwd(dpooled, h, scale, shift, mean, rstd, act, gh, part, B, HW, C, *, ldd, ldh, ldg,
             drop_p=0.0, seed=0, seed_dev=None):
    L.check(L.lib().an_pool_bwd(_c(dpooled), ldd, _c(h), ldh, _c(scale), _c(shift), _c(mean),
                                _c(rstd), act, _c(gh), ldg, _c(part), B, HW, C, drop_p, seed,
                                _c(seed_dev), L.stream_ptr()), "an_pool_bwd")


def ce_loss(logits, labels, B, NC, eps, grad_scale, loss, dlogits, *, ldl, ldd, nd_pad):
    L.check(L.lib().an_ce_loss(_c(logits), ldl, _c(labels), B, NC, eps, grad_scale, _c(loss),
                               _c(dlogits), ldd, nd_pad, L.stream_ptr()), "an_ce_loss")


def f32_to_bf16(src, dst, rows, cols, cols_pad, *, lds, ldd):
    L.check(L.lib().an_f32_to_bf16(_c(src), lds, _c(dst), ldd, rows, cols, cols_pad,
                                   L.stream_ptr()), "an_f32_to_bf16")


# ------------------------------------------------------------------------------------------------
# optimizer / weights
# ------------------------------------------------------------------------------------------------
def rmsprop_ema_step(p, g, sq, buf, ema, wd_flag, l1_pen, n, lr, alpha, momentum, eps,
                     eps_inside_sqrt, grad_scale, weight_decay, rho, ema_decay, hparams_dev=None):
    L.check(L.lib().an_rmsprop_ema_step(_c(p), _c(g), _c(sq), _c(buf), _c(ema), _c(wd_flag),
                                        _c(l1_pen), n, lr, alpha, momentum, eps,
                                        1 if eps_inside_sqrt else 0, grad_scale, weight_decay, rho,
                                        ema_decay, _c(hparams_dev), L.stream_ptr()), "an_rmsprop_ema_step")


def ema_update(ema, x, n, decay, hparams_dev=None):
    L.check(L.lib().an_ema_update(_c(ema), _c(x), n, decay, _c(hparams_dev), L.stream_ptr()),
            "an_ema_update")


def reg_losses(p, wd_flag, l1_pen, n, part):
    L.check(L.lib().an_reg_losses(_c(p), _c(wd_flag), _c(l1_pen), n, _c(part), L.stream_ptr()),
            "an_reg_losses")


def refresh_weights(table, n_entries, total_tiles, src, dst16, dst32):
    L.check(L.lib().an_refresh_weights(_c(table), n_entries, total_tiles, _c(src), _c(dst16),
                                       _c(dst32), L.stream_ptr()), "an_refresh_weights")


def gather_rows(src, idx, n_out, inner, dst):
    L.check(L.lib().an_gather_rows(_c(src), _c(idx), n_out, inner, _c(dst), L.stream_ptr()),
            "an_gather_rows")


def bn_eval_coeffs(gamma, beta, rm, rv, eps, C, scale, shift):
    L.check(L.lib().an_bn_eval_coeffs(_c(gamma), _c(beta), _c(rm), _c(rv), eps, C, _c(scale),
                                      _c(shift), L.stream_ptr()), "an_bn_eval_coeffs")


def act_bwd_reduce(dy, x, scale, shift, mean, rstd, act, g, part, M, C, *, ldy, ldx, ldg):
    L.check(L.lib().an_act_bwd_reduce(_c(dy), ldy, _c(x), ldx, _c(scale), _c(shift), _c(mean),
                                      _c(rstd), act, _c(g), ldg, M, C, _c(part), L.stream_ptr()),
            "an_act_bwd_reduce")


def sum_rows(part, rows, row_stride, n, dst):
    L.check(L.lib().an_sum_rows(_c(part), rows, row_stride, n, _c(dst), L.stream_ptr()),
            "an_sum_rows")


# ------------------------------------------------------------------------------------------------
# squeeze-and-excitation (fused block)
# ------------------------------------------------------------------------------------------------
def se_pool(x, ldx, scale, shift, act, B, HW, C, out, D=None, ldD=0, mean=True):
    L.check(L.lib().an_se_pool(_c(x), ldx, _c(D), ldD, _c(scale), _c(shift), act, B, HW, C, 1 if mean else 0,
                               _c(out), L.stream_ptr()), "an_se_pool")


def se_gate(pooled, W1, b1, W2, b2, chmap, act, B, C, total, mid, hpre, gate):
    L.check(L.lib().an_se_gate(_c(pooled), _c(W1), _c(b1), _c(W2), _c(b2), _c(chmap), act, B, C, total, mid,
                               _c(hpre), _c(gate), L.stream_ptr()), "an_se_gate")


def se_scale(x, ldx, scale, shift, act, gate, B, HW, C, out, ldo):
    L.check(L.lib().an_se_scale(_c(x), ldx, _c(scale), _c(shift), act, _c(gate), B, HW, C, _c(out), ldo,
                                L.stream_ptr()), "an_se_scale")


def se_gate_bwd(dgate, gate, hpre, pooled, W1, W2, chmap, inv_map, act, B, C, total, mid, dz, dh, ds, dW1, db1,
                dW2, db2):
    L.check(L.lib().an_se_gate_bwd(_c(dgate), _c(gate), _c(hpre), _c(pooled), _c(W1), _c(W2), _c(chmap),
                                   _c(inv_map), act, B, C, total, mid, _c(dz), _c(dh), _c(ds), _c(dW1), _c(db1),
                                   _c(dW2), _c(db2), L.stream_ptr()), "an_se_gate_bwd")


def se_bwd_apply_rows(M, C):
    return L.lib().an_se_bwd_apply_rows(M, C)


def se_bwd_apply(D, ldD, x, ldx, scale, shift, mean, rstd, act, gate, ds, B, HW, C, g, ldg, part):
    L.check(L.lib().an_se_bwd_apply(_c(D), ldD, _c(x), ldx, _c(scale), _c(shift), _c(mean), _c(rstd), act,
                                    _c(gate), _c(ds), B, HW, C, _c(g), ldg, _c(part), L.stream_ptr()),
            "an_se_bwd_apply")


# ------------------------------------------------------------------------------------------------
# multi-segment variants (one launch for the three branches of a block)
# ------------------------------------------------------------------------------------------------
def _bn_segs(segs):
    """segs: [(part tensor, rows, c0, C)]"""
    arr = (L.BnSeg * 4)()
    for i, (part, rows, c0, C) in enumerate(segs):
        arr[i].part, arr[i].rows, arr[i].c0, arr[i].C = _c(part), rows, c0, C
    return arr


def bn_finalize_multi(segs, count, gamma, beta, rm, rv, momentum, eps, mean, rstd, scale, shift):
    L.check(L.lib().an_bn_finalize_multi(_bn_segs(segs), len(segs), float(count), _c(gamma), _c(beta), _c(rm),
                                         _c(rv), momentum, eps, _c(mean), _c(rstd), _c(scale), _c(shift),
                                         L.stream_ptr()), "an_bn_finalize_multi")


def bn_bwd_finalize_multi(segs, count, gamma, mean, rstd, dgamma, dbeta, ca, cb, cc, raw_second=False):
    L.check(L.lib().an_bn_bwd_finalize_multi(_bn_segs(segs), len(segs), float(count), _c(gamma), _c(mean),
                                             _c(rstd), _c(dgamma), _c(dbeta), _c(ca), _c(cb), _c(cc),
                                             1 if raw_second else 0, L.stream_ptr()), "an_bn_bwd_finalize_multi")


def wgrad_reduce_multi(partials, splits, Cbig, Csmall, ncols, segs):
    """segs: [(row0, nrows, dst tensor view, sb, ss)]"""
    arr = (L.WgSeg * 4)()
    for i, (row0, nrows, dst, sb, ss) in enumerate(segs):
        arr[i].row0, arr[i].nrows, arr[i].dst, arr[i].dst_sb, arr[i].dst_ss = row0, nrows, _c(dst), sb, ss
    L.check(L.lib().an_wgrad_reduce_multi(_c(partials), splits, Cbig, Csmall, ncols, arr, len(segs),
                                          L.stream_ptr()), "an_wgrad_reduce_multi")
