import os, sys, torch, collections
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from atomnas_b200 import mobilenet_base as mb
from oracle import atomnas_oracle as O
torch.manual_seed(16 + 16 + 1)
inp, oup, stride, chans, ks, expand, actn, B, H = 16, 16, 1, [16], [3], False, "nn.ReLU", 8, 24
blk = mb.InvertedResidualChannels(inp, oup, stride, chans, ks, expand, active_fn=mb.get_active_fn(actn), batch_norm_kwargs={"momentum": 0.01, "eps": 1e-3})
blk.apply(mb.init_weights_mnas)
spec = O.NetSpec(inverted_residual_setting=[[1, 16, 1, 1, [3]]], input_channel=16, input_size=32, active_fn=actn, batch_norm_momentum=0.01, batch_norm_epsilon=1e-3)
ob = O.BlockSpec("b", inp, oup, stride, chans, ks, expand)
x = torch.randn(B, inp, H, H).to(torch.bfloat16).float()
st = collections.OrderedDict(("b." + k, v.detach().clone()) for k, v in blk.state_dict().items())
ref = O.block_forward_bf16_points(spec, st, ob, x, True)
blk = blk.cuda().train()
out = blk(x.cuda())
print("out err", (out.cpu() - ref).abs().max().item())
for n, b in blk.named_buffers():
    print(n, b.cpu().float().flatten()[:4].tolist(), st["b." + n].float().flatten()[:4].tolist())
print("---- with backward")
w = torch.randn(B, oup, H // stride, H // stride)
xg = x.cuda().requires_grad_(True)
out = blk(xg)
before = {n: b.clone() for n, b in blk.named_buffers()}
(out * w.cuda()).sum().backward()
torch.cuda.synchronize()
for n, b in blk.named_buffers():
    print(n, (b.float() - before[n].float()).abs().max().item())
from atomnas_b200 import engine as E
eng = E.get_engine(blk)
print("S arena", eng.S.numel(), "P", eng.P.numel(), "V", eng.V.numel())
