"""Optimisation utilities with the reference's surface (mirror of utils/optim.py).

ExponentialMovingAverage (:15-177), CrossEntropyLabelSmooth (:180-207), cal_l2_loss (:210-249),
get_lr_scheduler (:252-306), get_optimizer (:309-332) keep their names, arguments and
semantics so train.py / common.py run unchanged.  What changes is how the per-step work is done:

  * EMA: train.py:231-236 calls `ema(name, param, step)` once per tensor (799 calls per step).
    The first call of a step updates EVERY shadow of the owning engine with two launches of
    an_ema_update over the flat arenas; the remaining calls of that step only return the shadow.
  * the loss, L2 and schedules are tiny host-side / autograd expressions as in the reference (the
    fused path -- trainer.Trainer -- folds CE, L2 and L1 into kernels instead).

EMA updates of CPU tensors raise: there is no CPU path.
"""
from __future__ import division

import copy
import functools
import importlib
import logging
import warnings
import weakref
from collections import OrderedDict

import torch
from torch import nn

from . import engine as E
from . import kernels as K
from .rmsprop import RMSprop


class ExponentialMovingAverage(nn.Module):
    """tf.train.ExponentialMovingAverage: shadow = m*shadow + (1-m)*x, m = min(momentum,
    (1+n)/(10+n)) when the update count n is given."""

    def __init__(self, momentum, zero_debias=False):
        if zero_debias:
            raise NotImplementedError("zero_debias")
        if momentum < 0.0 or momentum > 1.0:
            raise ValueError("Invalid momentum value: {}".format(momentum))
        super(ExponentialMovingAverage, self).__init__()
        self._momentum, self._zero_debias = momentum, zero_debias
        self.clear()

    # ---- registry ----------------------------------------------------------------------------
    def clear(self):
        self._shadow, self._info = OrderedDict(), OrderedDict()
        self._source, self._flats = {}, {}

    def _check_exist(self, name):
        if name not in self._shadow:
            raise RuntimeError("{} has not been registered".format(name))

    def register(self, name, val, zero_init=False):
        if name in self._shadow:
            raise ValueError("Should not register twice for {}".format(name))
        if val.dtype not in (torch.float16, torch.float32, torch.float64):
            raise TypeError("The variables must be half, float, or double: {}".format(name))
        self._shadow[name] = torch.zeros_like(val) if zero_init else val.detach().clone()
        self._info[name] = {"num_updates": 0, "last_momemtum": None, "zero_init": zero_init,
                            "compress_masked": False}
        self._source[name] = weakref.ref(val)

    def pop(self, name):
        self._check_exist(name)
        self._source.pop(name, None)
        self._flats = {}
        return self._shadow.pop(name), self._info.pop(name)

    def average_names(self):
        return list(self._shadow.keys())

    def average(self, name):
        self._check_exist(name)
        return self._shadow[name]

    # ---- fused update ------------------------------------------------------------------------
    def _flat(self, eng):
        """Flat shadow arenas mirroring the engine's parameter / running-stat arenas."""
        flat = self._flats.get(id(eng))
        if flat is not None and flat["P"] is eng.P:
            return flat
        slots = {id(p): ("P", off, n) for p, off, n, _, _ in eng.param_slots}
        for mod, attr, off, n in eng.buf_slots:
            slots[id(getattr(mod, attr))] = ("S", off, n)
        flat = {"P": eng.P, "ema": {"P": eng.P.clone(), "S": eng.S.clone()}, "key": None, "names": set()}
        with torch.no_grad():
            for name, ref in self._source.items():
                t = ref()
                if t is None or id(t) not in slots or name not in self._shadow:
                    continue
                arena, off, n = slots[id(t)]
                view = flat["ema"][arena][off:off + n].view(t.shape)
                view.copy_(self._shadow[name])
                self._shadow[name] = view
                flat["names"].add(name)
        self._flats[id(eng)] = flat
        return flat

    def forward(self, name, x, num_updates=None):
        self._check_exist(name)
        m = self._momentum if num_updates is None else min(self._momentum,
                                                           (1.0 + num_updates) / (10.0 + num_updates))
        info = self._info[name]
        info["num_updates"] += 1
        info["last_momemtum"] = m
        if not x.is_cuda:
            raise RuntimeError("atomnas_b200 EMA: CPU tensor {} -- updates only run as sm_100a kernels".format(name))
        self._source[name] = weakref.ref(x)
        eng = E.engine_of_tensor(x)
        if eng is not None:
            flat = self._flat(eng)
            if name in flat["names"]:
                key = (info["num_updates"], num_updates)
                if flat["key"] != key:             # first tensor of this step: update the whole arena
                    K.ema_update(flat["ema"]["P"], eng.P, eng.P.numel(), m)
                    K.ema_update(flat["ema"]["S"], eng.S, eng.S.numel(), m)
                    flat["key"] = key
                return self._shadow[name]
        sh = self._shadow[name]
        K.ema_update(sh, x.detach().contiguous(), sh.numel(), m)
        return sh

    # ---- checkpointing -----------------------------------------------------------------------
    def state_dict(self):
        return {"info": self._info, "shadow": self._shadow,
                "param": {"momentum": self._momentum, "zero_debias": self._zero_debias}}

    def load_state_dict(self, state_dict):
        for key, val in state_dict["param"].items():
            cur = getattr(self, "_{}".format(key))
            if val != cur:
                msg = "EMA {} mismatch: current {} vs previous {}".format(key, cur, val)
                warnings.warn(msg, RuntimeWarning)
                logging.warning(msg)
        self._shadow = copy.deepcopy(state_dict["shadow"])
        self._info = copy.deepcopy(state_dict["info"])
        self._flats = {}

    def to(self, *args, **kwargs):
        device, dtype, non_blocking = torch._C._nn._parse_to(*args, **kwargs)[:3]
        for k in list(self._shadow.keys()):
            v = self._shadow[k]
            self._shadow[k] = v.to(device, dtype if v.is_floating_point() else None, non_blocking)
        self._flats = {}
        return self

    # ---- dynamic network shrinkage ---------------------------------------------------------------
    def compress_start(self):
        for val in self._info.values():
            val["compress_masked"] = False

    def compress_mask(self, info, verbose=False):
        old, new = info["var_old_name"], info["var_new_name"]
        if verbose:
            logging.info("EMA compress: {} -> {}".format(old, new))
        if self._info[old]["compress_masked"]:
            raise RuntimeError("May have dependencies in compress")
        if new in self._info and self._info[new]["compress_masked"]:
            raise RuntimeError("Compress {} twice".format(new))
        ema_old = self._shadow.pop(old)
        ema_new = torch.zeros_like(info["var_new"], device=ema_old.device)
        info["mask_hook"](ema_new, ema_old, info["mask"])
        entry = self._info.pop(old)
        entry["compress_masked"] = True
        self._info[new] = entry
        self._shadow[new] = ema_new
        self._source.pop(old, None)
        self._source[new] = weakref.ref(info["var_new"])
        self._flats = {}

    def compress_drop(self, info, verbose=False):
        name = info["var_old_name"]
        if verbose:
            logging.info("EMA drop: {}".format(name))
        self._check_exist(name)
        if self._info[name]["compress_masked"]:
            return None
        return self.pop(name)

    @staticmethod
    def adjust_momentum(momentum, steps_multi):
        """Keep the averaging horizon when the number of steps per sample changes."""
        return momentum ** (1.0 / steps_multi)


class CrossEntropyLabelSmooth(nn.Module):
    """-sum_c ((1-eps)*onehot + eps/C) * log_softmax(x); `reduction` in none|mean|sum."""

    def __init__(self, num_classes, label_smoothing, reduction="none"):
        super(CrossEntropyLabelSmooth, self).__init__()
        self.num_classes, self.label_smoothing = num_classes, label_smoothing
        reducers = {"none": lambda t: t, "mean": torch.mean, "sum": torch.sum}
        if reduction not in reducers:
            raise ValueError("Unknown reduction: {}".format(reduction))
        self.reduce_fun = reducers[reduction]

    def forward(self, inputs, targets):
        assert inputs.size(1) == self.num_classes
        logp = torch.log_softmax(inputs, dim=1)
        eps = self.label_smoothing
        nll = -logp.gather(1, targets.unsqueeze(1)).squeeze(1)
        return self.reduce_fun((1 - eps) * nll - eps / self.num_classes * logp.sum(1))


def cal_l2_loss(model, weight_decay, method):
    """0.5 * sum wd_i * |p_i|^2 ('mnas': convs, fc weight and fc bias; 'slimmable': non-depthwise
    convs and fc weight)."""
    loss = 0.0
    if method == "slimmable":
        for p in model.parameters():
            shape = list(p.size())
            if (len(shape) == 4 and shape[1] != 1) or len(shape) == 2:
                loss += weight_decay * (p ** 2).sum()
    elif method == "mnas":
        n_bias = 0
        for name, p in model.named_parameters():
            if p.dim() in (4, 2):
                loss += weight_decay * (p ** 2).sum()
            else:
                assert p.dim() == 1
                if "classifier" in name:
                    loss += weight_decay * (p ** 2).sum()
                    n_bias += 1
        assert n_bias == 1
    elif method == "mnas_no_bias":
        raise NotImplementedError()
    else:
        raise ValueError("Unknown weight_decay method: {}".format(method))
    return loss * 0.5


def get_lr_scheduler(optimizer, FLAGS):
    stepwise = FLAGS.get("lr_stepwise", True)
    spe = FLAGS._steps_per_epoch
    warm_iters = FLAGS.get("epoch_warmup", 5) * spe
    use_warmup = FLAGS.lr > FLAGS.base_lr

    def with_warmup(fn, i):
        if use_warmup and i <= warm_iters:
            r = FLAGS.base_lr / FLAGS.lr
            return r + i / warm_iters * (1 - r)
        return fn(i)

    name = FLAGS.lr_scheduler
    if name == "multistep":
        if use_warmup:
            raise NotImplementedError("Warmup not implemented for multistep")
        return torch.optim.lr_scheduler.MultiStepLR(
            optimizer, milestones=[spe * v for v in FLAGS.multistep_lr_milestones],
            gamma=FLAGS.multistep_lr_gamma)
    if name in ("exp_decaying", "exp_decaying_trunc"):
        floor = 0.05 if "trunc" in name else 0.0

        def decay(i):
            interval = spe * FLAGS.exp_decay_epoch_interval
            if not stepwise:
                i = (i // interval) * interval
            return max(FLAGS.exp_decaying_lr_gamma ** (i / interval), floor)

        return torch.optim.lr_scheduler.LambdaLR(optimizer, lr_lambda=functools.partial(with_warmup, decay))
    if name == "linear_decaying":
        assert stepwise
        return torch.optim.lr_scheduler.LambdaLR(
            optimizer, lr_lambda=functools.partial(with_warmup, lambda i: 1 - i / (FLAGS.num_epochs * spe)))
    raise NotImplementedError("Learning rate scheduler {} is not yet implemented.".format(name))


def get_optimizer(model, FLAGS):
    if FLAGS.optimizer == "sgd":
        return torch.optim.SGD(model.parameters(), lr=FLAGS.lr, momentum=FLAGS.momentum,
                               nesterov=FLAGS.nesterov, weight_decay=0)
    if FLAGS.optimizer == "rmsprop":
        return RMSprop(model.parameters(), lr=FLAGS.lr, alpha=FLAGS.alpha, momentum=FLAGS.momentum,
                       eps=FLAGS.epsilon, eps_inside_sqrt=FLAGS.eps_inside_sqrt, weight_decay=0)
    try:
        return importlib.import_module(FLAGS.optimizer).get_optimizer(model)
    except ImportError:
        raise NotImplementedError("Optimizer {} is not yet implemented.".format(FLAGS.optimizer))
