"""Host-side mirror of the reference's block zoo (models/mobilenet_base.py) on the B200 engine.

The classes keep the reference's nn.Module surface -- class names, constructor arguments,
child-module indices (hence state_dict keys `ops.{i}.{0|1}.{0|1}.*`, `ops.{i}.2.weight`,
`pw_bn.*`), the attributes the drivers read (`ops`, `pw_bn`, `channels`, `kernel_sizes`,
`expand`, `stride`, `input_dim`, `output_dim`, `use_res_connect`) and the methods they call
(`get_depthwise_bn`, `get_named_depthwise_bn`, `compress_by_mask`, `compress_by_threshold`) -- so
train.py / utils/prune.py / utils/model_profiling.py of the reference work unchanged.

What differs is execution: the nn.Conv2d / nn.BatchNorm2d children are PARAMETER HOLDERS.  A
block's forward never calls them; it runs the sm_100a kernels of atomnas_b200/csrc through
`engine.Engine` (there is no PyTorch/CPU fallback: a CPU tensor raises).

Reference lines: ConvBNReLU :120-142, InvertedResidualChannels :277-404, get_active_fn :407-415,
get_block :418-423, init_weights_* :426-459, output_network :462-479, _get_named_block_list
:482-488, _make_divisible :17-31.
"""
import collections
import functools
import math

import torch
from torch import nn

__all__ = [
    "ConvBNReLU", "InvertedResidualChannels", "Identity", "Narrow", "Swish", "HSwish",
    "get_active_fn", "get_block", "init_weights_mnas", "init_weights_slimmable",
    "output_network", "_get_named_block_list", "_make_divisible",
]


def _make_divisible(v, divisor, min_value=None):
    """Round a channel count to a multiple of `divisor` without dropping more than 10 %."""
    floor = divisor if min_value is None else min_value
    rounded = max(floor, int(v + divisor / 2) // divisor * divisor)
    return rounded + divisor if rounded < 0.9 * v else rounded


class Identity(nn.Module):
    def forward(self, x):
        return x


class Narrow(nn.Module):
    def __init__(self, dimension, start, length):
        super().__init__()
        self.dimension, self.start, self.length = dimension, start, length

    def forward(self, x):
        return x.narrow(self.dimension, self.start, self.length)


class Swish(nn.Module):
    """x * sigmoid(x); only a marker here -- the kernels apply activation id 3."""

    def forward(self, x):
        return x * torch.sigmoid(x)


class HSwish(object):
    def forward(self, x):
        return x * torch.nn.functional.relu6(x + 3, True).div_(6)


_ACTIVE_FNS = {
    "nn.ReLU6": functools.partial(nn.ReLU6, inplace=True),
    "nn.ReLU": functools.partial(nn.ReLU, inplace=True),
    "nn.Swish": Swish,
    "nn.HSwish": HSwish,
}


def get_active_fn(name):
    return _ACTIVE_FNS[name]


def active_fn_id(fn):
    """Kernel activation id (include/atomnas_b200.h AN_ACT_*) of a get_active_fn factory."""
    probe = fn() if fn is not None else None
    if probe is None or isinstance(probe, Identity):
        return 0
    if isinstance(probe, nn.ReLU6):
        return 2
    if isinstance(probe, nn.ReLU):
        return 1
    if isinstance(probe, Swish):
        return 3
    raise NotImplementedError("activation %r has no sm_100a kernel path" % (probe,))


class ConvBNReLU(nn.Sequential):
    """Conv2d (no bias) -> BatchNorm2d -> activation; children [0], [1], [2] hold parameters."""

    def __init__(self, in_planes, out_planes, kernel_size=3, stride=1, groups=1, active_fn=None,
                 batch_norm_kwargs=None):
        bn_kw = {} if batch_norm_kwargs is None else batch_norm_kwargs
        super().__init__(
            nn.Conv2d(in_planes, out_planes, kernel_size, stride, (kernel_size - 1) // 2,
                      groups=groups, bias=False),
            nn.BatchNorm2d(out_planes, **bn_kw),
            active_fn())

    def forward(self, x):
        from . import engine
        return engine.run_module(self, x)


class InvertedResidualChannels(nn.Module):
    """Atomic-block inverted residual: per branch 1x1 expand -> k x k depthwise -> 1x1 project,
    branch outputs summed, one BatchNorm, optional residual."""

    def __init__(self, inp, oup, stride, channels, kernel_sizes, expand, active_fn=None,
                 batch_norm_kwargs=None):
        super().__init__()
        assert stride in [1, 2]
        assert len(channels) == len(kernel_sizes)
        self.input_dim, self.output_dim = inp, oup
        self.expand, self.stride = expand, stride
        self.kernel_sizes, self.channels = kernel_sizes, channels
        self.use_res_connect = stride == 1 and inp == oup
        self.batch_norm_kwargs = batch_norm_kwargs
        self.active_fn = active_fn
        self.ops, self.pw_bn = self._build(channels, kernel_sizes, expand)

    def _branch(self, hidden, k, expand, bn_kw):
        mods = []
        if expand:
            mods.append(ConvBNReLU(self.input_dim, hidden, kernel_size=1, batch_norm_kwargs=bn_kw,
                                   active_fn=self.active_fn))
        elif hidden != self.input_dim:
            raise RuntimeError("non-expanding branch must keep the input width")
        mods.append(ConvBNReLU(hidden, hidden, kernel_size=k, stride=self.stride, groups=hidden,
                               batch_norm_kwargs=bn_kw, active_fn=self.active_fn))
        mods.append(nn.Conv2d(hidden, self.output_dim, 1, 1, 0, bias=False))
        return nn.Sequential(*mods)

    def _build(self, hidden_dims, kernel_sizes, expand):
        bn_kw = self.batch_norm_kwargs if self.batch_norm_kwargs is not None else {}
        ops = nn.ModuleList(self._branch(h, k, expand, bn_kw)
                            for k, h in zip(kernel_sizes, hidden_dims))
        if not expand and sum(hidden_dims) != self.input_dim:
            raise ValueError("Part of input are not used")
        return ops, nn.BatchNorm2d(self.output_dim, **bn_kw)

    # ---- what utils/prune.py and train.py read ------------------------------------------------
    def get_named_depthwise_bn(self, prefix=None):
        idx = 1 if self.expand else 0
        res = collections.OrderedDict()
        for i, op in enumerate(self.ops):
            name = "ops.{}.{}.1".format(i, idx)
            res[name if prefix is None else "{}.{}".format(prefix, name)] = op[idx][1]
        return res

    def get_depthwise_bn(self):
        return list(self.get_named_depthwise_bn().values())

    def forward(self, x):
        from . import engine
        return engine.run_module(self, x)

    def __repr__(self):
        return ("{}({}, {}, channels={}, kernel_sizes={}, expand={}, stride={})").format(
            self._get_name(), self.input_dim, self.output_dim, self.channels, self.kernel_sizes,
            self.expand, self.stride)

    def compress_by_mask(self, masks, **kwargs):
        from . import compress_utils as cu
        device = self.pw_bn.weight.device
        cu.copmress_inverted_residual_channels(self, masks, **kwargs)
        self.to(device)

    def compress_by_threshold(self, threshold, **kwargs):
        masks = [bn.weight.detach().abs() > threshold for bn in self.get_depthwise_bn()]
        self.compress_by_mask(masks, **kwargs)


def get_block(name):
    table = {"InvertedResidualChannels": InvertedResidualChannels}
    if name == "InvertedResidualChannelsFused":
        from .fused_block import InvertedResidualChannelsFused
        return InvertedResidualChannelsFused
    return table[name]


def init_weights_slimmable(m):
    if isinstance(m, nn.Conv2d):
        nn.init.kaiming_normal_(m.weight, mode="fan_out")
        if m.bias is not None:
            nn.init.zeros_(m.bias)
    elif isinstance(m, nn.BatchNorm2d):
        nn.init.ones_(m.weight)
        nn.init.zeros_(m.bias)
    elif isinstance(m, nn.Linear):
        nn.init.normal_(m.weight, 0, 0.01)
        nn.init.zeros_(m.bias)


def init_weights_mnas(m):
    """MnasNet init: conv N(0, sqrt(2/fan_out)) with fan_out = k*k for depthwise; BN (1, 0);
    Linear U(+-1/sqrt(out_features)), bias 0."""
    if isinstance(m, nn.Conv2d):
        kh, kw = m.kernel_size
        fan_out = kh * kw if m.groups == m.in_channels else m.out_channels * kh * kw
        nn.init.normal_(m.weight, 0.0, math.sqrt(2.0) / math.sqrt(fan_out))
        if m.bias is not None:
            nn.init.zeros_(m.bias)
    elif isinstance(m, nn.BatchNorm2d):
        nn.init.ones_(m.weight)
        nn.init.zeros_(m.bias)
    elif isinstance(m, nn.Linear):
        bound = 1.0 / math.sqrt(m.weight.shape[0])
        nn.init.uniform_(m.weight, -bound, bound)
        nn.init.zeros_(m.bias)


def output_network(model):
    """Network kwargs in the 6-tuple `searched_network` style."""
    kw = {k: getattr(model, k) for k in
          ("input_channel", "last_channel", "width_mult", "round_nearest", "active_fn", "num_classes")}
    kw["inverted_residual_setting"] = [
        [b.output_dim, 1, b.stride, b.kernel_sizes, b.channels, b.expand]
        for b in model.get_named_block_list().values()]
    return kw


def _get_named_block_list(m):
    """features[1:-2] keyed by their state_dict prefix."""
    children = list(m.features.named_children())[1:-2]
    return collections.OrderedDict(("features.{}".format(n), b) for n, b in children)


# the fused sibling and its SE module live in fused_block.py; exported here under the reference's
# names (models.mobilenet_base.InvertedResidualChannelsFused / SqueezeAndExcitation are looked up
# by utils/model_profiling.py:105-127 and models/searched_network.py:86-92)
from .fused_block import InvertedResidualChannelsFused, SqueezeAndExcitation  # noqa: E402,F401
