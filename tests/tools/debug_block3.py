import os, sys, torch, collections
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from atomnas_b200 import mobilenet_base as mb
from oracle import atomnas_oracle as O
CFG = [(16, 16, 1, [32, 24, 40], [3, 5, 7], True, "nn.ReLU", 8, 24), (16, 24, 2, [48, 48, 48], [3, 5, 7], True, "nn.ReLU6", 8, 24),
       (16, 16, 1, [16], [3], False, "nn.ReLU", 8, 24)]
for (inp, oup, stride, chans, ks, expand, actn, B, H) in CFG:
    torch.manual_seed(inp + oup + len(chans))
    blk = mb.InvertedResidualChannels(inp, oup, stride, chans, ks, expand, active_fn=mb.get_active_fn(actn), batch_norm_kwargs={"momentum": 0.01, "eps": 1e-3})
    blk.apply(mb.init_weights_mnas)
    spec = O.NetSpec(inverted_residual_setting=[[1, 16, 1, 1, [3]]], input_channel=16, input_size=32, active_fn=actn, batch_norm_momentum=0.01, batch_norm_epsilon=1e-3)
    ob = O.BlockSpec("b", inp, oup, stride, chans, ks, expand)
    x = torch.randn(B, inp, H, H).to(torch.bfloat16).float()
    w = torch.randn(B, oup, H // stride, H // stride)
    st = collections.OrderedDict(("b." + k, v.detach().clone()) for k, v in blk.state_dict().items())
    ref = O.block_forward_bf16_points(spec, st, ob, x, True)
    blk = blk.cuda().train()
    xg = x.cuda().requires_grad_(True)
    out = blk(xg)
    (out * w.cuda()).sum().backward()
    torch.cuda.synchronize()
    print(chans, "out err", (out.detach().cpu() - ref).abs().max().item())
    for n, b in blk.named_buffers():
        d = (b.cpu().float() - st["b." + n].float()).abs().max().item()
        if d > 1e-4: print("   MISMATCH", n, d, b.cpu().float().flatten()[:6].tolist())
