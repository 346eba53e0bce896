"""Unit parity of the non-GEMM CUDA kernels (depthwise, BN pieces, stem, pool, loss, optimizer).

Floating-point kernels are compared with a plain fp32 torch reference of the same op on the
same (bf16-rounded) inputs; index/copy kernels must be bit exact.  The optimizer and the loss
are compared with the CPU oracle (oracle/atomnas_oracle.py).
"""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from atomnas_b200 import kernels as K  # noqa: E402
from oracle import atomnas_oracle as O  # noqa: E402

DEV = "cuda"


def bf(x):
    return x.to(torch.bfloat16)


def act_f(v, act):
    return {0: lambda t: t, 1: torch.relu, 2: lambda t: t.clamp(0, 6), 3: lambda t: t * torch.sigmoid(t)}[act](v)


def act_grad(v, act):
    v = v.detach().clone().requires_grad_(True)
    act_f(v, act).sum().backward()
    return v.grad


def nchw(t):       # [B,H,W,C] -> [B,C,H,W] fp32
    return t.float().permute(0, 3, 1, 2).contiguous()


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def tap_major(w):  # [C,1,k,k] -> [k*k, C]
    C, _, k, _ = w.shape
    return w.reshape(C, k * k).t().contiguous()


DW_CASES = [
    # B, H, W, Ctot, c_off, C, k, s
    (2, 16, 16, 64, 0, 64, 3, 1), (3, 14, 14, 208, 48, 96, 5, 1), (2, 28, 28, 144, 0, 144, 7, 1),
    (2, 32, 32, 96, 32, 64, 3, 2), (2, 28, 28, 96, 0, 96, 5, 2), (9, 14, 14, 96, 0, 96, 7, 2),
    (65, 7, 7, 384, 128, 256, 7, 1), (2, 56, 56, 32, 0, 32, 3, 1), (4, 20, 20, 16, 0, 16, 3, 1),
    (2, 7, 7, 40, 8, 24, 5, 1), (70, 16, 16, 72, 0, 72, 5, 2),
]
# stride 1, k = 5 / 7, batch >= 8: the tensor-core path (dw_mma.cu) -- batch / width / channel remainders,
# both column-tile configurations (W <= 8: one tile x 32 channels, else two tiles x 16 channels)
MMA_CASES = [
    (16, 14, 14, 208, 48, 96, 5, 1), (20, 28, 28, 144, 0, 144, 7, 1), (17, 56, 56, 24, 0, 24, 5, 1),
    (33, 9, 9, 40, 8, 32, 7, 1), (16, 12, 20, 48, 0, 48, 7, 1), (8, 7, 7, 96, 0, 96, 5, 1), (40, 5, 5, 16, 0, 16, 7, 1),
]


@pytest.mark.parametrize("B,H,W,Ct,co,C,k,s", DW_CASES + MMA_CASES + [(3, 56, 56, 144, 0, 144, 7, 1), (3, 14, 14, 480, 0, 480, 5, 1), (2, 56, 56, 144, 0, 144, 5, 2)])
@pytest.mark.parametrize("act", [0, 1, 2, 3])
def test_dw_forward(B, H, W, Ct, co, C, k, s, act):
    torch.manual_seed(B * H + C + k)
    x = bf(torch.randn(B, H, W, Ct, device=DEV))
    w = torch.randn(C, 1, k, k, device=DEV) * (2.0 / k / k) ** 0.5
    scale = torch.rand(Ct, device=DEV) + 0.5
    shift = torch.randn(Ct, device=DEV) * 0.3
    Ho, Wo = K.dw_out_hw(H, W, k, s)
    y = torch.zeros(B, Ho, Wo, Ct, device=DEV, dtype=torch.bfloat16)
    rows = K.dw_fwd_stats_rows(B, C, Ho, Wo, k, s)
    part = torch.full((rows, 2, C), float("nan"), device=DEV)
    K.dw_fwd(x[..., co:], y[..., co:], tap_major(w), B, H, W, C, k, s, ldx=Ct, ldy=Ct,
             scale=scale[co:], shift=shift[co:], act=act, stats_part=part)
    u = act_f(scale[co:co + C] * x[..., co:co + C].float() + shift[co:co + C], act)     # fp32 on load
    ref = nhwc(F.conv2d(nchw(u), w, None, s, (k - 1) // 2, 1, C))
    got = y[..., co:co + C].float()
    assert (got - ref).abs().max().item() <= 1e-2 * ref.abs().max().item() + 1e-3
    if co > 0:
        assert y[..., :co].abs().max().item() == 0          # neighbours untouched
    st = part.sum(0)
    torch.testing.assert_close(st[0], got.sum((0, 1, 2)), rtol=1e-3, atol=0.05)
    torch.testing.assert_close(st[1], (got * got).sum((0, 1, 2)), rtol=1e-3, atol=0.05)


@pytest.mark.parametrize("B,H,W,Ct,co,C,k,s", DW_CASES + MMA_CASES + [(3, 56, 56, 144, 0, 144, 7, 1), (2, 56, 56, 144, 0, 144, 5, 2)])
@pytest.mark.parametrize("act", [1, 2, 3])
def test_dw_backward(B, H, W, Ct, co, C, k, s, act):
    if s == 2 and (H % 2 or W % 2):
        pytest.skip("stride 2 needs even sizes")
    torch.manual_seed(B + H + C + k + s)
    Ho, Wo = K.dw_out_hw(H, W, k, s)
    eraw = bf(torch.randn(B, H, W, Ct, device=DEV))
    g = bf(torch.randn(B, Ho, Wo, Ct, device=DEV))
    draw = bf(torch.randn(B, Ho, Wo, Ct, device=DEV))
    w = torch.randn(C, 1, k, k, device=DEV) * (2.0 / k / k) ** 0.5
    ca, cb, cc = (torch.randn(Ct, device=DEV) for _ in range(3))
    es = torch.rand(Ct, device=DEV) + 0.5
    eh = torch.randn(Ct, device=DEV) * 0.3
    em = torch.randn(Ct, device=DEV) * 0.1
    er = torch.rand(Ct, device=DEV) + 0.5
    sl = slice(co, co + C)
    # dd = BN-backward applied gradient, materialised by the element-wise kernel (the depthwise
    # kernels take it as their `g` operand; the producer's BN + activation is applied ON LOAD)
    ddf = torch.zeros(B, Ho, Wo, Ct, device=DEV, dtype=torch.bfloat16)
    K.affine_act(g[..., co:], ddf[..., co:], B * Ho * Wo, C, ld1=Ct, ldo=Ct, x2=draw[..., co:], ld2=Ct,
                 a1=ca[co:], a2=cb[co:], a3=cc[co:])
    dd = bf(ca[sl] * g[..., sl].float() + cb[sl] * draw[..., sl].float() + cc[sl])
    assert (ddf[..., sl].float() - dd.float()).abs().max().item() <= 1e-2 * dd.float().abs().max().item()
    dd = ddf[..., sl]
    pre = es[sl] * eraw[..., sl].float() + eh[sl]
    u = bf(act_f(pre, act))
    # ---- data gradient
    ge = torch.zeros(B, H, W, Ct, device=DEV, dtype=torch.bfloat16)
    rows = K.dw_bwd_data_stats_rows(B, C, H, W, k, s)
    part = torch.full((rows, 2, C), float("nan"), device=DEV)
    K.dw_bwd_data(ddf[..., co:], draw[..., co:], eraw[..., co:], ge[..., co:], tap_major(w), None,
                  None, None, es[co:], eh[co:], em[co:], er[co:], part, act, B, H, W, C, k, s,
                  ldg=Ct, ldd=Ct, lde=Ct, ldge=Ct)
    du = nhwc(torch.nn.grad.conv2d_input((B, C, H, W), w, nchw(dd), s, (k - 1) // 2, 1, C))
    ref = du * act_grad(pre, act)
    got = ge[..., sl].float()
    assert (got - ref).abs().max().item() <= 1e-2 * ref.abs().max().item() + 1e-3
    xh = (eraw[..., sl].float() - em[sl]) * er[sl]
    st = part.sum(0)
    n = B * H * W
    torch.testing.assert_close(st[0], got.sum((0, 1, 2)), rtol=2e-3, atol=0.02 * n ** 0.5)
    torch.testing.assert_close(st[1], (got * xh).sum((0, 1, 2)), rtol=2e-3, atol=0.02 * n ** 0.5)
    # ---- weight gradient
    wrows = K.dw_bwd_weight_rows(B, C, H, W, k, s)
    wpart = torch.full((wrows, k * k, C), float("nan"), device=DEV)
    dw = torch.zeros(C, 1, k, k, device=DEV)
    K.dw_bwd_weight(ddf[..., co:], draw[..., co:], eraw[..., co:], None, None, None, es[co:],
                    eh[co:], act, wpart, dw, B, H, W, C, k, s, ldg=Ct, ldd=Ct, lde=Ct)
    # the kernel applies BN + activation in fp32 on load (no bf16 rounding of the operand)
    uf = act_f(pre, act)
    wref = torch.nn.grad.conv2d_weight(nchw(uf), (C, 1, k, k), nchw(dd), s, (k - 1) // 2, 1, C)
    assert (dw - wref).abs().max().item() <= 2e-3 * wref.abs().max().item() + 1e-3


def test_dw_tensor_core_path_forced():
    """The per-shape dispatch (dw_mma.cu use_conv / use_wgrad) sends only the shapes where the banded MMA
    wins to the tensor cores; AN_DW_MMA=357 forces every stride-1 case of this file with >= 8 images
    through it (the switch is read once per process, hence the subprocess)."""
    import subprocess
    env = dict(os.environ, AN_DW_MMA="357")
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
                          "-k", "test_dw_forward or test_dw_backward"], env=env, capture_output=True, text=True, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]


def test_bn_finalize_and_backward_coefficients():
    torch.manual_seed(0)
    M, C = 5000, 72
    x = bf(torch.randn(M, C, device=DEV) * 2 + 0.5)
    rows = K.col_reduce_rows(M, C)
    part = torch.zeros(rows, 2, C, device=DEV)
    K.col_reduce(x, part, M, C, ldg=C)             # sum x, sum x^2
    gamma = torch.rand(C, device=DEV) + 0.5
    beta = torch.randn(C, device=DEV)
    rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    mean, rstd, scale, shift = (torch.empty(C, device=DEV) for _ in range(4))
    K.bn_finalize(part, rows, C, M, gamma, beta, rm, rv, 0.01, 1e-3, mean, rstd, scale, shift)
    xf = x.float()
    rm2, rv2 = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    ref = F.batch_norm(xf, rm2, rv2, gamma, beta, True, 0.01, 1e-3)
    torch.testing.assert_close(scale * xf + shift, ref, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(rm, rm2, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(rv, rv2, rtol=1e-4, atol=1e-6)
    # backward: dx = ca*g + cb*x + cc must equal autograd of batch_norm
    g = bf(torch.randn(M, C, device=DEV))
    part2 = torch.zeros(rows, 2, C, device=DEV)
    K.col_reduce(g, part2, M, C, ldg=C, x=x, ldx=C, mean=mean, rstd=rstd)
    dgamma, dbeta, ca, cb, cc = (torch.empty(C, device=DEV) for _ in range(5))
    K.bn_bwd_finalize(part2, rows, C, M, gamma, mean, rstd, dgamma, dbeta, ca, cb, cc)
    xa = xf.clone().requires_grad_(True)
    ga, ba = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    F.batch_norm(xa, None, None, ga, ba, True, 0.0, 1e-3).backward(g.float())
    torch.testing.assert_close(dgamma, ga.grad, rtol=1e-3, atol=1e-2)
    torch.testing.assert_close(dbeta, ba.grad, rtol=1e-3, atol=1e-2)
    torch.testing.assert_close(ca * g.float() + cb * xf + cc, xa.grad, rtol=1e-3, atol=2e-4)
    out = torch.zeros(M, C, device=DEV, dtype=torch.bfloat16)
    K.affine_act(g, out, M, C, ld1=C, ldo=C, x2=x, ld2=C, a1=ca, a2=cb, a3=cc)
    torch.testing.assert_close(out.float(), bf(ca * g.float() + cb * xf + cc).float(), rtol=1e-2, atol=1e-2)
    # residual form: NULL a2 => coefficient 1
    K.affine_act(g, out, M, C, ld1=C, ldo=C, x2=x, ld2=C, a1=scale, a3=shift)
    torch.testing.assert_close(out.float(), bf(scale * g.float() + shift + xf).float(), rtol=1e-2, atol=2e-2)


@pytest.mark.parametrize("B,H,C", [(3, 64, 8), (5, 32, 32), (2, 224, 16)])
def test_stem(B, H, C):
    torch.manual_seed(B + H)
    x = torch.randn(B, 3, H, H, device=DEV)
    w = torch.randn(C, 3, 3, 3, device=DEV) * 0.2
    Ho = H // 2
    y = torch.zeros(B, Ho, Ho, C, device=DEV, dtype=torch.bfloat16)
    rows = K.stem_fwd_stats_rows(B, Ho, Ho)
    part = torch.full((rows, 2, C), float("nan"), device=DEV)
    K.stem_fwd(x, w, y, part, B, H, H, C, ldy=C)
    ref = nhwc(F.conv2d(x, w, None, 2, 1))
    assert (y.float() - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()
    st = part.sum(0)
    torch.testing.assert_close(st[0], y.float().sum((0, 1, 2)), rtol=1e-3, atol=0.05)
    torch.testing.assert_close(st[1], (y.float() ** 2).sum((0, 1, 2)), rtol=1e-3, atol=0.05)
    # weight gradient with the BN-backward apply on load
    g = bf(torch.randn(B, Ho, Ho, C, device=DEV))
    ca, cb, cc = (torch.randn(C, device=DEV) for _ in range(3))
    dy = ca * g.float() + cb * y.float() + cc
    wrows = K.stem_wgrad_rows(B, Ho, Ho)
    wpart = torch.zeros(wrows, C, 27, device=DEV)
    dw = torch.zeros(C, 3, 3, 3, device=DEV)
    K.stem_wgrad(x, g, y, ca, cb, cc, wpart, dw, B, H, H, C, ldg=C, lds=C)
    wref = torch.nn.grad.conv2d_weight(x, (C, 3, 3, 3), nchw(dy), 2, 1)
    assert (dw - wref).abs().max().item() <= 1e-3 * wref.abs().max().item() + 1e-3


@pytest.mark.parametrize("drop_p", [0.0, 0.3])
def test_pool_forward_backward(drop_p):
    torch.manual_seed(4)
    B, HW, C, act = 19, 49, 200, 1
    h = bf(torch.randn(B, HW, C, device=DEV))
    scale = torch.rand(C, device=DEV) + 0.5
    shift = torch.randn(C, device=DEV) * 0.2
    mean = torch.randn(C, device=DEV) * 0.1
    rstd = torch.rand(C, device=DEV) + 0.5
    pooled = torch.zeros(B, C, device=DEV, dtype=torch.bfloat16)
    K.pool_fwd(h, scale, shift, act, pooled, B, HW, C, ldh=C, ldp=C, drop_p=drop_p, seed=77)
    pre = scale * h.float() + shift
    ref = act_f(pre, act).mean(1)
    if drop_p == 0.0:
        torch.testing.assert_close(pooled.float(), ref, rtol=1e-2, atol=1e-2)
        keep = torch.ones(B, C, device=DEV)
    else:
        keep = (pooled.float().abs() > 0).float() / (1 - drop_p)
        frac = (keep > 0).float().mean().item()
        assert abs(frac - (1 - drop_p)) < 0.03
        nz = keep > 0
        torch.testing.assert_close(pooled.float()[nz], (ref * keep)[nz], rtol=1e-2, atol=1e-2)
    dp = bf(torch.randn(B, C, device=DEV))
    gh = torch.zeros(B, HW, C, device=DEV, dtype=torch.bfloat16)
    rows = K.pool_bwd_stats_rows(B)
    part = torch.full((rows, 2, C), float("nan"), device=DEV)
    K.pool_bwd(dp, h, scale, shift, mean, rstd, act, gh, part, B, HW, C, ldd=C, ldh=C, ldg=C,
               drop_p=drop_p, seed=77)
    gref = (dp.float() * keep / HW).unsqueeze(1) * act_grad(pre, act)
    mask = (ref.abs() > 1e-3).unsqueeze(1).expand_as(gref) if drop_p > 0 else torch.ones_like(gref, dtype=torch.bool)
    assert ((gh.float() - gref).abs() * mask).max().item() <= 1e-2 * gref.abs().max().item() + 1e-4
    st = part.sum(0)
    xh = (h.float() - mean) * rstd
    torch.testing.assert_close(st[0], gh.float().sum((0, 1)), rtol=1e-3, atol=1e-2)
    torch.testing.assert_close(st[1], (gh.float() * xh).sum((0, 1)), rtol=1e-3, atol=1e-2)


@pytest.mark.parametrize("B,NC", [(7, 10), (64, 1000)])
def test_ce_loss_matches_oracle(B, NC):
    torch.manual_seed(B)
    pad = (NC + 7) // 8 * 8
    logits = torch.randn(B, pad, device=DEV) * 3
    labels = torch.randint(0, NC, (B,), device=DEV)
    loss = torch.zeros(B, device=DEV)
    dl = torch.full((B, pad), 7.0, device=DEV, dtype=torch.bfloat16)
    K.ce_loss(logits, labels, B, NC, 0.1, 1.0 / B, loss, dl, ldl=pad, ldd=pad, nd_pad=pad)
    lc = logits[:, :NC].cpu().requires_grad_(True)
    ref = O.cross_entropy_label_smooth(lc, labels.cpu(), NC, 0.1)
    ref.mean().backward()
    torch.testing.assert_close(loss.cpu(), ref.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(dl[:, :NC].float().cpu(), lc.grad, rtol=1e-2, atol=1e-5)
    assert dl[:, NC:].abs().max().item() == 0 if pad > NC else True


def test_rmsprop_ema_matches_oracle():
    torch.manual_seed(1)
    n = 256 * 40
    p = torch.randn(n)
    p[5] = 0.0
    g = torch.randn(n) * 0.1
    sq, buf = torch.rand(n) * 0.01, torch.randn(n) * 0.1
    ema = p.clone()
    wd_flag = (torch.arange(40) % 2).float().repeat_interleave(256)
    l1_pen = (((torch.arange(40) % 3) == 0).float() * 2.5).repeat_interleave(256)
    wd_flag[7] = -1.0      # frozen padding element
    lr, wd, rho, decay, world = 0.128, 1e-5, 1e-4, 0.99995, 8
    d = [t.clone().to(DEV) for t in (p, g, sq, buf, ema, wd_flag, l1_pen)]
    K.rmsprop_ema_step(d[0], d[1], d[2], d[3], d[4], d[5], d[6], n, lr, 0.9, 0.9, 1e-3, True,
                       1.0 / world, wd, rho, decay)
    gg = g / world + wd * wd_flag * p + rho * l1_pen * torch.sign(p)
    p7, s7, b7 = p[7].clone(), sq[7].clone(), buf[7].clone()
    O.rmsprop_step(p, gg, sq, buf, lr)
    p[7], sq[7], buf[7] = p7, s7, b7
    O.ema_update(ema, p, decay)
    torch.testing.assert_close(d[0].cpu(), p, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(d[2].cpu(), sq, rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(d[3].cpu(), buf, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(d[4].cpu(), ema, rtol=1e-6, atol=1e-6)
    part = torch.zeros(128, 2, device=DEV)
    K.reg_losses(d[0], d[5], d[6], n, part)
    s = part.sum(0).cpu()
    torch.testing.assert_close(s[0], (wd_flag.clamp(min=0) * p * p).sum(), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(s[1], (l1_pen * p.abs()).sum(), rtol=1e-4, atol=1e-4)


def test_refresh_and_gather_bit_exact():
    torch.manual_seed(2)
    src = torch.randn(5000, device=DEV)
    # entry 0: [24, 40] -> bf16 [24, ld 48]; entry 1: same matrix transposed -> bf16 [40, ld 24]
    # entry 2: depthwise [16, 9] -> fp32 tap-major [9, 16]
    ents = [(100, 0, 24, 40, 48, 0, 0), (100, 2000, 24, 40, 24, 1, 0), (1500, 10, 16, 9, 16, 1, 1)]
    table, blk = [], 0
    for so, do, r, c, ld, tr, f32 in ents:
        table.append([c, so, do, r, c, ld, tr, f32, blk])
        blk += ((r + 31) // 32) * ((c + 31) // 32)
    tt = torch.tensor(table, dtype=torch.int64, device=DEV)
    d16 = torch.zeros(4000, device=DEV, dtype=torch.bfloat16)
    d32 = torch.zeros(400, device=DEV)
    K.refresh_weights(tt, len(ents), blk, src, d16, d32)
    m = src[100:100 + 960].view(24, 40)
    assert torch.equal(d16[:24 * 48].view(24, 48)[:, :40], bf(m))
    assert torch.equal(d16[2000:2000 + 960].view(40, 24), bf(m.t()))
    dwm = src[1500:1500 + 144].view(16, 9)
    assert torch.equal(d32[10:10 + 144].view(9, 16), dwm.t())
    mask = torch.rand(24, device=DEV) > 0.5
    idx = torch.nonzero(mask).flatten()
    out = torch.zeros(int(mask.sum()), 40, device=DEV)
    K.gather_rows(m.contiguous(), idx, idx.numel(), 40, out)
    assert torch.equal(out, m[mask])


# ------------------------------------------------------------------------------------------------
# materialised operands (scale == NULL): same row-streaming kernels (dw2.cu / dw_stream.cu), both strides
# ------------------------------------------------------------------------------------------------
DW_REAL = [   # the AtomNAS-C layer shapes (4 / 2 / 1 columns per thread, 7 column-group warps)
    (3, 56, 56, 144, 0, 144, 7, 1), (2, 28, 28, 200, 8, 136, 5, 1), (3, 14, 14, 480, 0, 480, 7, 1),
    (5, 7, 7, 1152, 0, 1152, 3, 1), (9, 7, 7, 1160, 8, 1152, 7, 1), (2, 112, 112, 32, 0, 32, 3, 1),
    (2, 112, 112, 96, 0, 96, 7, 2), (2, 56, 56, 144, 0, 144, 5, 2), (3, 28, 28, 240, 0, 240, 3, 2),
    (3, 28, 28, 240, 0, 240, 7, 2), (5, 14, 14, 576, 0, 576, 7, 2), (5, 14, 14, 584, 8, 576, 5, 2),
    (17, 14, 14, 96, 0, 96, 3, 1),
]


@pytest.mark.parametrize("B,H,W,Ct,co,C,k,s", DW_CASES + DW_REAL)
def test_dw_forward_plain_operand(B, H, W, Ct, co, C, k, s):
    torch.manual_seed(B * H + C + k + 1)
    u = bf(torch.randn(B, H, W, Ct, device=DEV))
    w = torch.randn(C, 1, k, k, device=DEV) * (2.0 / k / k) ** 0.5
    Ho, Wo = K.dw_out_hw(H, W, k, s)
    y = torch.zeros(B, Ho, Wo, Ct, device=DEV, dtype=torch.bfloat16)
    rows = K.dw_fwd_stats_rows(B, C, Ho, Wo, k, s, fused=False)
    part = torch.full((rows, 2, C), float("nan"), device=DEV)
    K.dw_fwd(u[..., co:], y[..., co:], tap_major(w), B, H, W, C, k, s, ldx=Ct, ldy=Ct, stats_part=part)
    ref = nhwc(F.conv2d(nchw(u[..., co:co + C]), w, None, s, (k - 1) // 2, 1, C))
    got = y[..., co:co + C].float()
    assert (got - ref).abs().max().item() <= 1e-2 * ref.abs().max().item() + 1e-3
    if co > 0:
        assert y[..., :co].abs().max().item() == 0
    assert y[..., co + C:].abs().max().item() == 0 if co + C < Ct else True
    st = part.sum(0)
    torch.testing.assert_close(st[0], got.sum((0, 1, 2)), rtol=1e-3, atol=0.05)
    torch.testing.assert_close(st[1], (got * got).sum((0, 1, 2)), rtol=1e-3, atol=0.05)


@pytest.mark.parametrize("B,H,W,Ct,co,C,k,s", DW_CASES + DW_REAL)
@pytest.mark.parametrize("act", [1, 3])
def test_dw_backward_plain_operands(B, H, W, Ct, co, C, k, s, act):
    if s == 2 and (H % 2 or W % 2):
        pytest.skip("stride 2 needs even sizes")
    torch.manual_seed(B + H + C + k + s + 1)
    Ho, Wo = K.dw_out_hw(H, W, k, s)
    eraw = bf(torch.randn(B, H, W, Ct, device=DEV))
    dd = bf(torch.randn(B, Ho, Wo, Ct, device=DEV))
    w = torch.randn(C, 1, k, k, device=DEV) * (2.0 / k / k) ** 0.5
    es = torch.rand(Ct, device=DEV) + 0.5
    eh = torch.randn(Ct, device=DEV) * 0.3
    em = torch.randn(Ct, device=DEV) * 0.1
    er = torch.rand(Ct, device=DEV) + 0.5
    sl = slice(co, co + C)
    pre = es[sl] * eraw[..., sl].float() + eh[sl]
    u = bf(act_f(es * eraw.float() + eh, act))
    ge = torch.zeros(B, H, W, Ct, device=DEV, dtype=torch.bfloat16)
    rows = K.dw_bwd_data_stats_rows(B, C, H, W, k, s, fused=False)
    part = torch.full((rows, 2, C), float("nan"), device=DEV)
    K.dw_bwd_data(dd[..., co:], dd[..., co:], eraw[..., co:], ge[..., co:], tap_major(w), None, None, None,
                  es[co:], eh[co:], em[co:], er[co:], part, act, B, H, W, C, k, s, ldg=Ct, ldd=Ct, lde=Ct, ldge=Ct)
    du = nhwc(torch.nn.grad.conv2d_input((B, C, H, W), w, nchw(dd[..., sl]), s, (k - 1) // 2, 1, C))
    ref = du * act_grad(pre, act)
    got = ge[..., sl].float()
    assert (got - ref).abs().max().item() <= 1e-2 * ref.abs().max().item() + 1e-3
    xh = (eraw[..., sl].float() - em[sl]) * er[sl]
    st = part.sum(0)
    n = B * H * W
    torch.testing.assert_close(st[0], got.sum((0, 1, 2)), rtol=2e-3, atol=0.02 * n ** 0.5)
    torch.testing.assert_close(st[1], (got * xh).sum((0, 1, 2)), rtol=2e-3, atol=0.02 * n ** 0.5)
    wrows = K.dw_bwd_weight_rows(B, C, H, W, k, s, fused=False)
    wpart = torch.full((wrows, k * k, C), float("nan"), device=DEV)
    dw = torch.zeros(C, 1, k, k, device=DEV)
    K.dw_bwd_weight(dd[..., co:], dd[..., co:], u[..., co:], None, None, None, None, None, 0, wpart, dw, B, H, W, C,
                    k, s, ldg=Ct, ldd=Ct, lde=Ct)
    wref = torch.nn.grad.conv2d_weight(nchw(u[..., sl]), (C, 1, k, k), nchw(dd[..., sl]), s, (k - 1) // 2, 1, C)
    assert (dw - wref).abs().max().item() <= 2e-3 * wref.abs().max().item() + 1e-3
