import os, sys, torch, collections
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from atomnas_b200 import mobilenet_base as mb
from oracle import atomnas_oracle as O
torch.manual_seed(0)
for (inp, oup, stride, chans, ks, expand, actn, B, H) in [
        (16, 16, 1, [32, 24, 40], [3, 5, 7], True, "nn.ReLU", 8, 24),
        (16, 24, 2, [48, 48, 48], [3, 5, 7], True, "nn.ReLU", 8, 24),
        (16, 16, 1, [16], [3], False, "nn.ReLU", 8, 24),
        (16, 16, 1, [32, 24, 40], [3, 5, 7], True, "nn.Swish", 8, 24)]:
    blk = mb.InvertedResidualChannels(inp, oup, stride, chans, ks, expand, active_fn=mb.get_active_fn(actn),
                                      batch_norm_kwargs={"momentum": 0.01, "eps": 1e-3})
    blk.apply(mb.init_weights_mnas)
    with torch.no_grad():
        for m in blk.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.copy_(1 + 0.2 * torch.randn_like(m.weight)); m.bias.copy_(0.2 * torch.randn_like(m.bias))
    spec = O.NetSpec(inverted_residual_setting=[[1, 16, 1, 1, [3]]], input_channel=16, input_size=32,
                     active_fn=actn, batch_norm_momentum=0.01, batch_norm_epsilon=1e-3)
    ob = O.BlockSpec("b", inp, oup, stride, chans, ks, expand)
    x = torch.randn(B, inp, H, H).to(torch.bfloat16).float()
    w = torch.randn(B, oup, H // stride, H // stride)
    res = {}
    for mode in ("fp32", "bf16pts"):
        st = collections.OrderedDict(("b." + k, v.detach().clone()) for k, v in blk.state_dict().items())
        for k in st:
            if st[k].dtype == torch.float32 and "running" not in k: st[k].requires_grad_(True)
        xr = x.clone().requires_grad_(True)
        out = (O.block_forward if mode == "fp32" else O.block_forward_bf16_points)(spec, st, ob, xr, True)
        (out * w).sum().backward()
        res[mode] = (out.detach(), xr.grad, {k[2:]: v.grad for k, v in st.items() if v.requires_grad})
    b2 = blk.cuda().train()
    xg = x.cuda().requires_grad_(True)
    out = b2(xg)
    (out * w.cuda()).sum().backward()
    print("block", inp, oup, stride, chans, ks, expand, actn)
    for mode in ("fp32", "bf16pts"):
        o, dx, gr = res[mode]
        gmax = max(v.norm().item() for v in gr.values())
        line = "  vs %-7s out %.4f dx %.4f |" % (mode, ((out.cpu() - o).norm() / o.norm()).item(), ((xg.grad.cpu() - dx).norm() / dx.norm()).item())
        for n, p in b2.named_parameters():
            r = gr[n]
            line += " %s:%.3f" % (n.replace("ops.", "").replace(".weight", "w").replace(".bias", "b"), ((p.grad.cpu() - r).norm() / max(r.norm().item(), 1e-3 * gmax)).item())
        print(line)
