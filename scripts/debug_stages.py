"""Teacher-forced per-stage check of the engine's forward buffers (debug helper, GPU)."""
import collections
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from atomnas_b200 import mobilenet_supernet as ms  # noqa: E402
from atomnas_b200 import engine as E  # noqa: E402

g = torch.load(os.path.join(ROOT, "tests/golden", "tiny_supernet_relu6.pt"), weights_only=False)
kw = dict(g["kwargs"])
kw["active_fn"] = "nn.ReLU"
B, S = 8, 64
model = ms.MobileNetV2(**kw, input_size=S)
model.load_state_dict(g["state0"])
model = model.cuda().train()
torch.manual_seed(1)
x = torch.randn(B, 3, S, S, device="cuda")
eng = E.get_engine(model)
with torch.no_grad():
    eng.forward(x, True, True, 1)
bf = next(iter(eng.bufs.values()))


def bfr(t):
    return t.to(torch.bfloat16).float()


def nchw(t, b, h, w):
    return t.float().view(b, h, w, -1).permute(0, 3, 1, 2)


def flat(t):
    return t.permute(0, 2, 3, 1).reshape(-1, t.shape[1])


def rep(name, got, ref):
    err = (got.float() - ref).abs().max().item()
    print("%-28s max|ref| %.3f  max err %.5f  rel %.5f" % (name, ref.abs().max().item(), err, err / (ref.abs().max().item() + 1e-9)))


def bn_coeffs(raw, bn):
    m = raw.mean(0)
    v = raw.var(0, unbiased=False)
    sc = bn.weight / torch.sqrt(v + bn.eps)
    return sc, bn.bias - m * sc


feats = list(model.features.children())
h, w = bf.s_hw
s_ref = flat(F.conv2d(x, feats[0][0].weight, None, 2, 1))
rep("stem raw", bf.s_raw, s_ref)
sc, sh = bn_coeffs(bf.s_raw.float(), feats[0][1])
rep("stem y0", bf.y0, torch.relu(sc * bf.s_raw.float() + sh))
rep("stem scale", eng.bn_vec(eng.stem.bn, "scale"), sc)
cur = bf.y0
for p, bb in zip(eng.blocks, bf.blk):
    xin = cur.float()
    if p.expand:
        rep(p.name + " e_raw", bb.e_raw, xin @ bfr(eng.W16[p.we16:p.we16 + p.Kh * p.inp].float().view(p.Kh, p.inp)).t())
        wcat = torch.cat([br.ex_conv.weight.view(br.h, p.inp) for br in p.branches])
        rep(p.name + " We16", eng.W16[p.we16:p.we16 + p.Kh * p.inp].float().view(p.Kh, p.inp), bfr(wcat))
        sc, sh = bn_coeffs(bb.e_raw.float(), None) if False else (None, None)
        m = bb.e_raw.float().mean(0); v = bb.e_raw.float().var(0, unbiased=False)
        gam = torch.cat([br.ex_bn.weight for br in p.branches]); bet = torch.cat([br.ex_bn.bias for br in p.branches])
        sc = gam / torch.sqrt(v + 1e-3); sh = bet - m * sc
        rep(p.name + " bn_e scale", eng.bn_vec(p.bn_e, "scale"), sc)
        u = bfr(torch.relu(sc * bb.e_raw.float() + sh))
    else:
        u = xin
    outs = []
    for br in p.branches:
        ui = nchw(u[:, br.off:br.off + br.h], B, bb.hin, bb.win)
        outs.append(flat(F.conv2d(ui, br.dw_conv.weight, None, p.stride, (br.k - 1) // 2, 1, br.h)))
    rep(p.name + " d_raw", bb.d_raw, torch.cat(outs, 1))
    m = bb.d_raw.float().mean(0); v = bb.d_raw.float().var(0, unbiased=False)
    gam = torch.cat([br.dw_bn.weight for br in p.branches]); bet = torch.cat([br.dw_bn.bias for br in p.branches])
    sc = gam / torch.sqrt(v + 1e-3); sh = bet - m * sc
    rep(p.name + " bn_d scale", eng.bn_vec(p.bn_d, "scale"), sc)
    rep(p.name + " bn_d shift", eng.bn_vec(p.bn_d, "shift"), sh)
    vv = bfr(torch.relu(sc * bb.d_raw.float() + sh))
    wp = bfr(torch.cat([br.proj.weight.view(p.oup, br.h) for br in p.branches], 1))
    rep(p.name + " p_raw", bb.p_raw, vv @ wp.t())
    m = bb.p_raw.float().mean(0); v = bb.p_raw.float().var(0, unbiased=False)
    sc = p.mod.pw_bn.weight / torch.sqrt(v + 1e-3); sh = p.mod.pw_bn.bias - m * sc
    yy = sc * bb.p_raw.float() + sh
    if p.res:
        yy = yy + xin
    rep(p.name + " y", bb.y, yy)
    cur = bb.y
hd = eng.head
rep("head raw", bf.h_raw, cur.float() @ bfr(feats[-2][0].weight.view(hd.C, hd.cin)).t())
m = bf.h_raw.float().mean(0); v = bf.h_raw.float().var(0, unbiased=False)
sc = feats[-2][1].weight / torch.sqrt(v + 1e-3); sh = feats[-2][1].bias - m * sc
pooled = torch.relu(sc * bf.h_raw.float() + sh).view(B, -1, hd.C).mean(1)
rep("pooled", bf.pooled, pooled)
fc = model.classifier[1]
rep("logits", bf.logits[:, :hd.NC], bf.pooled.float() @ bfr(fc.weight).t() + fc.bias)
