import os, sys, torch, collections
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from atomnas_b200 import mobilenet_supernet as ms
from oracle import atomnas_oracle as O
g = torch.load(os.path.join(ROOT, "tests/golden", "tiny_supernet_relu6.pt"), weights_only=False)
torch.set_num_threads(16)
verbose = len(sys.argv) > 1
for act in ["nn.ReLU", "nn.Swish"]:
  for B, S in [(4, 64), (32, 64)]:
    kw = dict(g["kwargs"]); kw["active_fn"] = act
    model = ms.MobileNetV2(**kw, input_size=S)
    model.load_state_dict(g["state0"])
    spec = O.NetSpec(**kw, input_size=S)
    torch.manual_seed(B + S)
    x = torch.randn(B, 3, S, S); y = torch.randint(0, 10, (B,))
    refs = {}
    for mode in ("fp32", "bf16pts"):
        st = collections.OrderedDict((k, v.clone()) for k, v in g["state0"].items())
        names = O.param_names(spec)
        for n in names: st[n].requires_grad_(True)
        lo = (O.forward if mode == "fp32" else O.forward_bf16_points)(spec, st, x, True)
        O.cross_entropy_label_smooth(lo, y, 10, 0.1).mean().backward()
        refs[mode] = (lo.detach(), {n: st[n].grad for n in names})
    model = model.cuda().train()
    lg = model(x.cuda())
    O.cross_entropy_label_smooth(lg, y.cuda(), 10, 0.1).mean().backward()
    for mode in ("fp32", "bf16pts"):
        lo, gr = refs[mode]
        errs = []
        gmax = max(v.norm().item() for v in gr.values())
        for n, p in model.named_parameters():
            r = gr[n].cuda()
            if r.norm().item() > 1e-3 * gmax:
                e = ((p.grad - r).norm() / r.norm()).item()
                errs.append(e)
                if verbose and mode == "bf16pts" and e > 0.02: print("   ", n, e, r.norm().item())
        e = torch.tensor(errs)
        print("%s B=%d S=%d vs %-7s logits err %.4f | grad rel err median %.4f mean %.4f max %.4f (n=%d)" % (
            act, B, S, mode, (lg.cpu() - lo).abs().max().item(), e.median().item(), e.mean().item(), e.max().item(), len(errs)))
