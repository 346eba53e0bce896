"""tcgen05 pointwise GEMM / wgrad kernels vs a plain fp32 torch reference of the same op."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from atomnas_b200 import kernels as K


def _act(v, act):
    if act == 1:
        return torch.relu(v)
    if act == 2:
        return torch.clamp(v, 0, 6)
    if act == 3:
        return v * torch.sigmoid(v)
    return v


def _bf(x):
    return x.to(torch.bfloat16)


def _aeff(x1, x2, mode, act, a1, a2, a3):
    if mode == 0:
        return x1.float()
    if mode == 1:
        return _bf(_act(a1 * x1.float() + a3, act)).float()
    return _bf(a1 * x1.float() + a2 * x2.float() + a3).float()


SHAPES = [
    # M, N, K
    (128, 16, 32), (300, 24, 96), (1000, 96, 16), (4096, 432, 24), (777, 40, 144),
    (2500, 720, 40), (640, 80, 240), (12544 // 8, 1152, 192), (392, 320, 1152), (256, 1000, 1280),
]


@pytest.mark.parametrize("M,N,K_", SHAPES)
@pytest.mark.parametrize("a_mode", [0, 1, 2])
def test_pw_gemm_forward_stats(M, N, K_, a_mode):
    torch.manual_seed(M + N + K_ + a_mode)
    dev = "cuda"
    x1 = _bf(torch.randn(M, K_, device=dev))
    x2 = _bf(torch.randn(M, K_, device=dev))
    w = _bf(torch.randn(N, K_, device=dev) / K_ ** 0.5)
    a1 = torch.rand(K_, device=dev) + 0.5
    a2 = torch.randn(K_, device=dev)
    a3 = torch.randn(K_, device=dev) * 0.3
    y = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    rows = K.pw_gemm_stats_rows(M, N)
    part = torch.full((rows, 2, N), float("nan"), device=dev)
    K.pw_gemm(x1, w, y, M, N, K_, x2=x2, a_mode=a_mode, a_act=1, a1=a1, a2=a2, a3=a3,
              stats_part=part)
    torch.cuda.synchronize()
    ref = _aeff(x1, x2, a_mode, 1, a1, a2, a3) @ w.float().t()
    err = (y.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 1e-2 * scale + 1e-3, (err, scale)
    # statistics are those of the rounded output
    s = part.sum(0)
    assert torch.isfinite(s).all()
    yf = y.float()
    torch.testing.assert_close(s[0], yf.sum(0), rtol=2e-3, atol=2e-2 * M ** 0.5)
    torch.testing.assert_close(s[1], (yf * yf).sum(0), rtol=2e-3, atol=1e-2)


def test_pw_gemm_f32_bias():
    torch.manual_seed(0)
    M, N, K_ = 256, 1000, 1280
    x = _bf(torch.randn(M, K_, device="cuda"))
    w = _bf(torch.randn(N, K_, device="cuda") / K_ ** 0.5)
    b = torch.randn(N, device="cuda")
    y = torch.zeros(M, N, device="cuda")
    K.pw_gemm(x, w, y, M, N, K_, bias=b)
    ref = x.float() @ w.float().t() + b
    torch.testing.assert_close(y, ref, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("M,N,K_", [(300, 96, 24), (5000, 432, 24), (1000, 24, 144)])
@pytest.mark.parametrize("act", [1, 2, 3])
def test_pw_gemm_dgrad_act_bn_epilogue(M, N, K_, act):
    torch.manual_seed(1 + M)
    dev = "cuda"
    dy = _bf(torch.randn(M, K_, device=dev))
    wt = _bf(torch.randn(N, K_, device=dev) / K_ ** 0.5)
    r = _bf(torch.randn(M, N, device=dev) * 2)
    scale = torch.rand(N, device=dev) + 0.5
    shift = torch.randn(N, device=dev)
    mean = torch.randn(N, device=dev) * 0.1
    rstd = torch.rand(N, device=dev) + 0.5
    g = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    rows = K.pw_gemm_stats_rows(M, N)
    part = torch.zeros(rows, 2, N, device=dev)
    K.pw_gemm(dy, wt, g, M, N, K_, epi_mode=1, r=r, r_scale=scale, r_shift=shift, r_mean=mean,
              r_rstd=rstd, r_act=act, stats_part=part)
    d = dy.float() @ wt.float().t()
    pre = (scale * r.float() + shift).requires_grad_(True)
    _act(pre, act).backward(torch.ones_like(pre))
    gref = d * pre.grad
    err = (g.float() - gref).abs().max().item()
    assert err <= 1e-2 * gref.abs().max().item() + 1e-3
    xh = (r.float() - mean) * rstd
    s = part.sum(0)
    torch.testing.assert_close(s[0], g.float().sum(0), rtol=2e-3, atol=0.05 * M ** 0.5)
    torch.testing.assert_close(s[1], (g.float() * r.float()).sum(0), rtol=2e-3, atol=0.1 * M ** 0.5)


def test_pw_gemm_residual_epilogue():
    torch.manual_seed(2)
    M, N, K_ = 1000, 24, 432
    a = _bf(torch.randn(M, K_, device="cuda"))
    w = _bf(torch.randn(N, K_, device="cuda") / K_ ** 0.5)
    r = _bf(torch.randn(M, N, device="cuda"))
    y = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    K.pw_gemm(a, w, y, M, N, K_, epi_mode=2, r=r)
    ref = a.float() @ w.float().t() + r.float()
    assert (y.float() - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()


WG_SHAPES = [
    # Mrows, Cbig, Csmall
    (640, 128, 16), (1000, 96, 24), (5000, 432, 24), (3000, 288, 16), (2048, 720, 40),
    (1568, 1152, 192), (392, 3456, 320), (256, 1280, 1000), (12544, 1280, 320),
]


@pytest.mark.parametrize("Mr,Cb,Cs", WG_SHAPES)
@pytest.mark.parametrize("p_mode", [0, 1, 2])
def test_pw_wgrad(Mr, Cb, Cs, p_mode):
    torch.manual_seed(Mr + Cb + Cs + p_mode)
    dev = "cuda"
    p1 = _bf(torch.randn(Mr, Cb, device=dev))
    p2 = _bf(torch.randn(Mr, Cb, device=dev))
    q = _bf(torch.randn(Mr, Cs, device=dev))
    a1 = torch.rand(Cb, device=dev) + 0.5
    a2 = torch.randn(Cb, device=dev)
    a3 = torch.randn(Cb, device=dev) * 0.3
    splits = K.pw_wgrad_splits(Mr, Cb, Cs)
    partials = torch.full((splits, Cb, Cs), float("nan"), device=dev)
    K.pw_wgrad(p1, q, partials, Mr, Cb, Cs, p2=p2, p_mode=p_mode, p_act=1, a1=a1, a2=a2, a3=a3)
    out = torch.zeros(Cb, Cs, device=dev)
    K.wgrad_reduce(partials, splits, Cb, Cs, out, Cs, 1)
    out_t = torch.zeros(Cs, Cb, device=dev)
    K.wgrad_reduce(partials, splits, Cb, Cs, out_t, 1, Cb)
    ref = _aeff(p1, p2, p_mode, 1, a1, a2, a3).t() @ q.float()
    tol = 2e-3 * ref.abs().max().item() + 1e-3
    assert (out - ref).abs().max().item() <= tol
    assert torch.equal(out_t, out.t())
    if Cs > 8:
        sub = torch.zeros(Cs - 8, 40, device=dev)       # rows [16,56) of the big side, transposed
        K.wgrad_reduce(partials, splits, Cb, Cs, sub, 1, 40, row0=16, nrows=40, ncols=Cs - 8)
        assert torch.equal(sub, out[16:56, :Cs - 8].t())
