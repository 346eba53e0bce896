"""InvertedResidualChannelsFused + SqueezeAndExcitation module surface.

Mirrors models/mobilenet_base.py:93-117 (SE) and :145-274 (fused block): constructor arguments,
child names (`expand_conv`, `depth_ops`, `project_conv`, `se_op`) and therefore state_dict keys,
so AtomNAS-A+/B+/C+ checkpoints and YAMLs load.  Execution: `forward` hands the block (or the whole
network it sits in) to engine.Engine, which maps the single expand conv / BatchNorm and the single
projection conv onto the same branch-fused kernels as InvertedResidualChannels (8-padded channel
slots per branch) and runs the SE gate with the kernels of csrc/se.cu.  There is no PyTorch / CPU
fallback.  A stand-alone SqueezeAndExcitation call is not part of the hot path and raises.
"""
import collections

import torch
from torch import nn

from .mobilenet_base import ConvBNReLU, Identity, Narrow


class SqueezeAndExcitation(nn.Module):

    def __init__(self, n_feature, n_hidden, spatial_dims=[2, 3], active_fn=None):
        super().__init__()
        self.n_feature, self.n_hidden, self.spatial_dims = n_feature, n_hidden, spatial_dims
        self.se_reduce = nn.Conv2d(n_feature, n_hidden, 1, bias=True)
        self.se_expand = nn.Conv2d(n_hidden, n_feature, 1, bias=True)
        self.active_fn = active_fn()

    def forward(self, x):
        raise NotImplementedError("stand-alone SqueezeAndExcitation is not on the B200 hot path: it runs inside "
                                  "InvertedResidualChannelsFused (csrc/se.cu)")

    def __repr__(self):
        return "{}({}, {}, spatial_dims={}, active_fn={})".format(
            self._get_name(), self.n_feature, self.n_hidden, self.spatial_dims, self.active_fn)


class InvertedResidualChannelsFused(nn.Module):
    supports_se = True

    def __init__(self, inp, oup, stride, channels, kernel_sizes, expand, active_fn=None,
                 batch_norm_kwargs=None, se_ratio=None):
        super().__init__()
        assert stride in [1, 2]
        assert len(channels) == len(kernel_sizes)
        self.input_dim, self.output_dim, self.expand, self.stride = inp, oup, expand, stride
        self.kernel_sizes, self.channels = kernel_sizes, channels
        self.use_res_connect = stride == 1 and inp == oup
        self.batch_norm_kwargs, self.active_fn, self.se_ratio = batch_norm_kwargs, active_fn, se_ratio
        bn_kw = batch_norm_kwargs if batch_norm_kwargs is not None else {}
        total = sum(channels)
        self.expand_conv = (ConvBNReLU(inp, total, kernel_size=1, batch_norm_kwargs=bn_kw, active_fn=active_fn)
                            if expand else Identity())
        start, ops = 0, nn.ModuleList()
        for k, h in zip(kernel_sizes, channels):
            layers = []
            if expand:
                layers.append(Narrow(1, start, h))
                start += h
            elif h != inp:
                raise RuntimeError("non-expanding branch must keep the input width")
            layers.append(ConvBNReLU(h, h, kernel_size=k, stride=stride, groups=h, batch_norm_kwargs=bn_kw,
                                     active_fn=active_fn))
            ops.append(nn.Sequential(*layers))
        self.depth_ops = ops
        self.project_conv = nn.Sequential(nn.Conv2d(total, oup, 1, 1, 0, bias=False),
                                          nn.BatchNorm2d(oup, **bn_kw))
        if expand and start != total:
            raise ValueError("Part of expanded are not used")
        self.se_op = (SqueezeAndExcitation(total, int(round(inp * se_ratio)), active_fn=active_fn)
                      if se_ratio is not None else Identity())

    def get_named_depthwise_bn(self, prefix=None):
        if not self.expand:
            raise RuntimeError("Not search_first")
        res = collections.OrderedDict()
        for i, op in enumerate(self.depth_ops):
            name = "depth_ops.{}.1.1".format(i)
            res[name if prefix is None else "{}.{}".format(prefix, name)] = op[1][1]
        return res

    def get_depthwise_bn(self):
        return list(self.get_named_depthwise_bn().values())

    def forward(self, x):
        from . import engine
        return engine.run_module(self, x)

    def __repr__(self):
        return ("{}({}, {}, channels={}, kernel_sizes={}, expand={}, stride={}, se_ratio={})").format(
            self._get_name(), self.input_dim, self.output_dim, self.channels, self.kernel_sizes,
            self.expand, self.stride, self.se_ratio)
