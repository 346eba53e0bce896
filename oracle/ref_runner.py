"""Run the UNMODIFIED reference staged under oracle/_ref/AtomNAS (see oracle/make_ref.py).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's reference arms /
cpu_baseline.  Nothing under atomnas_b200/ may import it.

`train_state(...)` builds exactly what train.py builds for one replica -- the reference's own
`models.mobilenet_supernet.Model`, `utils.rmsprop.RMSprop`, `CrossEntropyLabelSmooth`,
`ExponentialMovingAverage`, `get_bn_to_prune` penalties -- and `train_step` is the loop body of
train.py:174-236 (forward_loss common.py:67-80, cal_l2_loss, cal_bn_l1_loss, backward, optimizer
step, per-name EMA) written against those objects.  Works on CPU (fp32, the reference as shipped)
and on a CUDA device (fp32 as shipped, or bf16 autocast + channels_last = its best light).
"""
import contextlib
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref", "AtomNAS")
STUBS = os.path.join(HERE, "_ref", "stubs")

_mods = None


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "models", "mobilenet_base.py"))


def load():
    """Import the reference's modules from oracle/_ref/AtomNAS (once)."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError("oracle/_ref/AtomNAS is missing: run `python oracle/make_ref.py` where /root/reference exists")
    for name in ("models", "utils"):
        m = sys.modules.get(name)
        if m is not None and not os.path.abspath(getattr(m, "__file__", "") or "").startswith(REF_ROOT):
            raise RuntimeError("a different top-level `%s` package is already imported; the reference needs its own" % name)
    sys.path.insert(0, STUBS)
    sys.path.insert(0, REF_ROOT)
    import importlib

    class M(object):
        pass

    m = M()
    m.mb = importlib.import_module("models.mobilenet_base")
    m.ms = importlib.import_module("models.mobilenet_supernet")
    m.sn = importlib.import_module("models.searched_network")
    m.optim = importlib.import_module("utils.optim")
    m.prune = importlib.import_module("utils.prune")
    m.rmsprop = importlib.import_module("utils.rmsprop")
    m.profiling = importlib.import_module("utils.model_profiling")
    m.ucommon = importlib.import_module("utils.common")
    _mods = m
    return m


class RefState(object):
    pass


def train_state(kwargs, input_size, device="cpu", seed=1995, lr=0.128, ema_decay=0.9999 ** 0.5, state_dict=None,
                searched=False, channels_last=False):
    R = load()
    torch.manual_seed(seed)
    lib = R.sn if searched else R.ms
    model = lib.Model(**kwargs, input_size=input_size)
    model.apply(R.mb.init_weights_mnas)
    if state_dict is not None:
        model.load_state_dict(state_dict)
    R.profiling.model_profiling(model, input_size, input_size, use_cuda=False, num_forwards=0, verbose=False)
    model = model.to(device)
    if channels_last:
        model = model.to(memory_format=torch.channels_last)
    model.train()
    st = RefState()
    st.R, st.model, st.device, st.channels_last = R, model, device, channels_last
    st.prune_info = R.prune.get_bn_to_prune(model, {"bn_prune_filter": "expansion_only_skip_expand1"}, verbose=False)
    st.opt = R.rmsprop.RMSprop(model.parameters(), lr=lr, alpha=0.9, momentum=0.9, eps=1e-3, eps_inside_sqrt=True,
                               weight_decay=0)
    st.crit = R.optim.CrossEntropyLabelSmooth(kwargs["num_classes"], 0.1, reduction="none")
    st.ema = R.optim.ExponentialMovingAverage(ema_decay)
    for n, p in model.named_parameters():
        st.ema.register(n, p)
    for n, b in model.named_buffers():
        if "running_var" in n or "running_mean" in n:
            st.ema.register(n, b)
    st.global_step = 0
    return st


def train_step(st, x, y, weight_decay=1e-5, rho=1e-4, autocast_bf16=False):
    """train.py:174-236 for one batch; returns the mean CE loss (python float is NOT taken here)."""
    R, model = st.R, st.model
    st.opt.zero_grad()
    if st.channels_last:
        x = x.contiguous(memory_format=torch.channels_last)
    ctx = torch.autocast("cuda", dtype=torch.bfloat16) if autocast_bf16 else contextlib.nullcontext()
    with ctx:
        out = model(x)
    loss = torch.mean(st.crit(out.float(), y))
    loss_l2 = R.optim.cal_l2_loss(model, weight_decay, "mnas")
    weights = R.ucommon.get_params_by_name(model, st.prune_info.weight)
    loss_l1 = R.prune.cal_bn_l1_loss(weights, st.prune_info.penalty, rho)
    (loss + loss_l2 + loss_l1).backward()
    st.opt.step()
    st.global_step += 1
    names = st.ema.average_names()
    params = R.ucommon.get_params_by_name(model, names)
    for name, param in zip(names, params):
        st.ema(name, param, st.global_step)
    return loss.detach()
