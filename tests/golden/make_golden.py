"""Generate golden vectors by running the UNMODIFIED reference (imported from /root/reference).

Run in the build container only (the reference does not exist on the GPU box):

    python tests/golden/make_golden.py

Writes tests/golden/*.pt (small, committed).  Each file holds the model kwargs, the initial
state_dict, the inputs and what the reference's own modules / optimizer / losses produced.
"""
import collections
import copy
import os
import sys

import torch

REF = "/root/reference"
sys.argv = sys.argv[:1]
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))

import models.mobilenet_base as mb  # noqa: E402
import models.mobilenet_supernet as ms  # noqa: E402
import models.searched_network as sn  # noqa: E402
from utils import optim as roptim  # noqa: E402
from utils import prune as rprune  # noqa: E402
from utils.model_profiling import model_profiling  # noqa: E402
from utils.rmsprop import RMSprop  # noqa: E402

torch.set_num_threads(4)


def clone_sd(sd):
    return collections.OrderedDict((k, v.detach().clone()) for k, v in sd.items())


TINY = dict(
    num_classes=10, input_channel=8, last_channel=64, width_mult=1.0, round_nearest=8,
    inverted_residual_setting=[[1, 8, 1, 1, [3]], [3, 16, 2, 2, [3, 5, 7]], [3, 16, 1, 2, [3, 5]],
                               [3, 24, 1, 2, [3, 5, 7]], [3, 24, 1, 2, [3]]],
    dropout_ratio=0.0, batch_norm_momentum=0.01, batch_norm_epsilon=1e-3, active_fn="nn.ReLU",
    block="InvertedResidualChannels")


def build(kwargs, input_size, seed, lib=ms):
    torch.manual_seed(seed)
    model = lib.Model(**kwargs, input_size=input_size)
    model.apply(mb.init_weights_mnas)
    # de-trivialise BN affine parameters so that gamma/beta paths are exercised
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.copy_(1.0 + 0.2 * torch.randn(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
    return model


def train_case(name, kwargs, input_size, batch, steps, seed, rho, lr=0.016):
    model = build(kwargs, input_size, seed)
    model_profiling(model, input_size, input_size, use_cuda=False, num_forwards=0, verbose=False)
    n_macs = int(model.n_macs)
    block_macs = {k: [int(op.n_macs) for op in b.ops] for k, b in model.get_named_block_list().items()}
    model.train()
    flags = {"bn_prune_filter": "expansion_only_skip_expand1"}
    prune_info = rprune.get_bn_to_prune(model, flags, verbose=False)
    opt = RMSprop(model.parameters(), lr=lr, alpha=0.9, momentum=0.9, eps=1e-3, eps_inside_sqrt=True,
                  weight_decay=0)
    crit = roptim.CrossEntropyLabelSmooth(kwargs["num_classes"], 0.1, reduction="none")
    ema = roptim.ExponentialMovingAverage(0.9999 ** 0.5)
    for n, p in model.named_parameters():
        ema.register(n, p)
    for n, b in model.named_buffers():
        if "running_var" in n or "running_mean" in n:
            ema.register(n, b)
    out = {"kwargs": kwargs, "input_size": input_size, "state0": clone_sd(model.state_dict()),
           "n_macs": n_macs, "block_macs": block_macs, "prune_names": prune_info.weight,
           "penalties": prune_info.penalty,
           "per_channel_flops": prune_info.get_info_list("per_channel_flops"),
           "rho": rho, "lr": lr, "weight_decay": 1e-5, "ema_decay": 0.9999 ** 0.5, "steps": []}
    g = torch.Generator().manual_seed(seed + 7)
    named = dict(model.named_parameters())
    for step in range(steps):
        x = torch.randn(batch, 3, input_size, input_size, generator=g)
        y = torch.randint(0, kwargs["num_classes"], (batch,), generator=g)
        opt.zero_grad()
        logits = model(x)
        ce = crit(logits, y)
        loss = torch.mean(ce)
        l2 = roptim.cal_l2_loss(model, 1e-5, "mnas")
        l1 = rprune.cal_bn_l1_loss([named[n] for n in prune_info.weight], prune_info.penalty, rho)
        (loss + l2 + l1).backward()
        grads = collections.OrderedDict((n, p.grad.detach().clone()) for n, p in named.items())
        opt.step()
        for n in ema.average_names():
            v = named[n] if n in named else dict(model.named_buffers())[n]
            ema(n, v, step + 1)
        rec = {"x": x, "y": y, "logits": logits.detach().clone(), "ce": ce.detach().clone(),
               "loss": float(loss), "l2": float(l2), "l1": float(l1),
               "state_after": clone_sd(model.state_dict())}
        if step == 0:
            rec["grads"] = grads
        out["steps"].append(rec)
    out["square_avg"] = collections.OrderedDict(
        (n, opt.state[p]["square_avg"].clone()) for n, p in named.items())
    out["momentum_buffer"] = collections.OrderedDict(
        (n, opt.state[p]["momentum_buffer"].clone()) for n, p in named.items())
    out["ema"] = collections.OrderedDict((n, ema.average(n).clone()) for n in ema.average_names())
    torch.save(out, os.path.join(HERE, name + ".pt"))
    print(name, "loss", [s["loss"] for s in out["steps"]], "macs", n_macs)
    return model, opt, ema, prune_info


def fwd_bwd_case(name, kwargs, input_size, batch, seed, lib=ms):
    model = build(kwargs, input_size, seed, lib)
    model.train()
    g = torch.Generator().manual_seed(seed + 3)
    x = torch.randn(batch, 3, input_size, input_size, generator=g)
    y = torch.randint(0, kwargs["num_classes"], (batch,), generator=g)
    state0 = clone_sd(model.state_dict())
    logits = model(x)
    ce = roptim.CrossEntropyLabelSmooth(kwargs["num_classes"], 0.1, reduction="none")(logits, y)
    torch.mean(ce).backward()
    out = {"kwargs": kwargs, "input_size": input_size, "state0": state0, "x": x, "y": y,
           "logits": logits.detach().clone(), "ce": ce.detach().clone(),
           "grads": collections.OrderedDict((n, p.grad.clone()) for n, p in model.named_parameters()),
           "state_after": clone_sd(model.state_dict())}
    model.eval()
    with torch.no_grad():
        out["logits_eval"] = model(x).clone()
    torch.save(out, os.path.join(HERE, name + ".pt"))
    print(name, "ce mean", float(ce.mean()))


def shrink_case(name):
    """compress_by_mask on a trained-for-2-steps tiny supernet, with optimizer/EMA/PruneInfo."""
    model, opt, ema, prune_info = train_case("_tmp_shrink", TINY, 64, 4, 2, 11, 1e-4)
    os.remove(os.path.join(HERE, "_tmp_shrink.pt"))
    before = clone_sd(model.state_dict())
    sq_before = collections.OrderedDict(
        (n, opt.state[p]["square_avg"].clone()) for n, p in model.named_parameters())
    ema_before = collections.OrderedDict((n, ema.average(n).clone()) for n in ema.average_names())
    g = torch.Generator().manual_seed(5)
    all_masks = collections.OrderedDict()
    for bname, block in model.get_named_block_list().items():
        masks = []
        for i, bn in enumerate(block.get_depthwise_bn()):
            m = torch.rand(bn.weight.numel(), generator=g) > 0.4
            if not block.expand:
                m[:] = True           # features.1 is never pruned (skip_expand1); see train.py:46-64
            if bname == "features.3" and i == 1:
                m[:] = False          # a dead branch: dropped, survivors renumbered
            masks.append(m)
        all_masks[bname] = [m.clone() for m in masks]
        block.compress_by_mask(masks, ema=ema, optimizer=opt, prune_info=prune_info, prefix=bname,
                               verbose=False)
    assert set(opt.param_groups[0]["params"]) == set(model.parameters())
    model_profiling(model, 64, 64, use_cuda=False, num_forwards=0, verbose=False)
    x = torch.randn(4, 3, 64, 64, generator=g)
    model.train()
    state_after = clone_sd(model.state_dict())
    logits = model(x)
    out = {"kwargs": TINY, "input_size": 64, "state_before": before, "masks": all_masks,
           "sq_before": sq_before, "ema_before": ema_before,
           "state_after": state_after,
           "sq_after": collections.OrderedDict(
               (n, opt.state[p]["square_avg"].clone()) for n, p in model.named_parameters()),
           "ema_after": collections.OrderedDict((n, ema.average(n).clone()) for n in ema.average_names()),
           "prune_names_after": prune_info.weight, "penalties_after": prune_info.penalty,
           "n_macs_after": int(model.n_macs), "network_after": mb.output_network(model),
           "x": x, "logits_train_after": logits.detach().clone()}
    torch.save(out, os.path.join(HERE, name + ".pt"))
    print(name, "macs after", out["n_macs_after"], out["network_after"]["inverted_residual_setting"])


def known_answers(name):
    """Scalars of the reference's schedules and losses at known points."""
    out = {}
    params = {"epoch_free": 1, "epoch_warmup": 5, "scheduler": "linear", "rho": 1e-4, "stepwise": True}
    f = rprune.get_rho_scheduler(params, 10)
    out["rho_stepwise"] = [f(i) for i in range(0, 70, 3)]
    params["stepwise"] = False
    f = rprune.get_rho_scheduler(params, 10)
    out["rho_epochwise"] = [f(i) for i in range(0, 70, 3)]
    torch.manual_seed(0)
    logits = torch.randn(6, 10)
    y = torch.randint(0, 10, (6,))
    out["ce_logits"], out["ce_y"] = logits, y
    out["ce"] = roptim.CrossEntropyLabelSmooth(10, 0.1, reduction="none")(logits, y)
    e = roptim.ExponentialMovingAverage(0.99)
    e.register("a", torch.ones(3))
    vals = []
    for i in range(1, 6):
        e("a", torch.full((3,), float(i)), i)
        vals.append(e.average("a").clone())
    out["ema_vals"] = vals
    ws = [torch.tensor([0.5, -2.0, 3.0]), torch.tensor([1.6, -1.7, 1.5, 0.0])]
    out["mask_w"] = ws
    out["mask"] = rprune.cal_mask_network_slimming_by_threshold(ws, 1.5)
    torch.save(out, os.path.join(HERE, name + ".pt"))
    print(name, "ok")


def fused_block_case(name, inp, oup, stride, channels, kernel_sizes, expand, active_fn, se_ratio, hw, batch, seed):
    """One InvertedResidualChannelsFused (+ SqueezeAndExcitation) block of the reference in train mode:
    forward, input gradient, every parameter gradient and the BN buffers after the step
    (models/mobilenet_base.py:93-117,145-268; the block of the AtomNAS-C+ searched nets)."""
    torch.manual_seed(seed)
    bn_kw = {"momentum": 0.01, "eps": 1e-3}
    blk = mb.InvertedResidualChannelsFused(inp, oup, stride, channels, kernel_sizes, expand,
                                           active_fn=mb.get_active_fn(active_fn), batch_norm_kwargs=bn_kw,
                                           se_ratio=se_ratio)
    blk.apply(mb.init_weights_mnas)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for m in blk.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.copy_(1.0 + 0.2 * torch.randn(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
        if se_ratio is not None:                      # init_weights_mnas zeroes conv biases: make them matter
            blk.se_op.se_reduce.bias.copy_(0.1 * torch.randn(blk.se_op.se_reduce.bias.shape, generator=g))
            blk.se_op.se_expand.bias.copy_(0.1 * torch.randn(blk.se_op.se_expand.bias.shape, generator=g))
    blk.train()
    state0 = clone_sd(blk.state_dict())
    x = torch.randn(batch, inp, hw, hw, generator=g, requires_grad=True)
    y = blk(x)
    wgt = torch.randn(y.shape, generator=g)
    (y * wgt).sum().backward()
    out = {"args": dict(inp=inp, oup=oup, stride=stride, channels=channels, kernel_sizes=kernel_sizes,
                        expand=expand, active_fn=active_fn, se_ratio=se_ratio, bn_momentum=0.01, bn_eps=1e-3),
           "state0": state0, "x": x.detach().clone(), "w": wgt, "y": y.detach().clone(),
           "dx": x.grad.detach().clone(),
           "grads": collections.OrderedDict((n, p.grad.detach().clone()) for n, p in blk.named_parameters()),
           "state_after": clone_sd(blk.state_dict())}
    torch.save(out, os.path.join(HERE, name + ".pt"))
    print(name, "ok", tuple(y.shape))


def calibration_case(name, kwargs, input_size, batch, nbatches, seed):
    """BN re-calibration as train.py:419-458 runs it: model.eval(); model.apply(bn_calibration)
    (utils/common.py:213-226: reset stats, BN modules back to train(), momentum=None -> cumulative
    average), a few no-grad forward passes, then eval-mode logits with the calibrated statistics."""
    from utils.common import bn_calibration
    model = build(kwargs, input_size, seed)
    g = torch.Generator().manual_seed(seed + 3)
    with torch.no_grad():                               # start from non-trivial running statistics
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(0.3 * torch.randn(m.running_mean.shape, generator=g))
                m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))
    state0 = clone_sd(model.state_dict())
    model.eval()
    model.apply(bn_calibration)
    xs = [torch.randn(batch, 3, input_size, input_size, generator=g) for _ in range(nbatches)]
    with torch.no_grad():
        for x in xs:
            model(x)
    state_cal = clone_sd(model.state_dict())
    model.eval()
    with torch.no_grad():
        logits = model(xs[0])
    torch.save({"kwargs": kwargs, "input_size": input_size, "state0": state0, "xs": xs,
                "state_calibrated": state_cal, "logits_eval": logits}, os.path.join(HERE, name + ".pt"))
    print(name, "ok")


if __name__ == "__main__":
    if os.environ.get("GOLDEN_ONLY") == "calibration":
        calibration_case("tiny_bn_calibration", TINY, 64, 4, 3, 31)
        sys.exit(0)
    if os.environ.get("GOLDEN_ONLY") == "fused":           # only the fused / SE block fixtures
        fused_block_case("tiny_fused_se_swish", 16, 16, 1, [16, 24, 8], [3, 5, 7], True, "nn.Swish", 0.5, 12, 3, 21)
        fused_block_case("tiny_fused_relu_s2", 16, 24, 2, [24, 40], [3, 5], True, "nn.ReLU", None, 12, 3, 22)
        sys.exit(0)
    train_case("tiny_supernet_train", TINY, 64, 4, 2, 3, 1e-4)
    for fn in ("nn.ReLU6", "nn.Swish"):
        kw = copy.deepcopy(TINY)
        kw["active_fn"] = fn
        fwd_bwd_case("tiny_supernet_" + fn.split(".")[1].lower(), kw, 64, 4, 5)
    searched = dict(TINY)
    searched["inverted_residual_setting"] = [[8, 1, 1, [3], [8], False], [16, 1, 2, [3, 5, 7], [24, 16, 40], True],
                                             [16, 1, 1, [3, 7], [32, 8], True], [24, 1, 2, [5], [48], True],
                                             [24, 1, 2, [3, 5], [16, 24], True], [32, 1, 2, [7], [40], True]]
    fwd_bwd_case("tiny_searched", searched, 64, 2, 9, lib=sn)
    shrink_case("tiny_shrink")
    known_answers("known_answers")
