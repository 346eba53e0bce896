"""Supernet / searched-network builders on the B200 engine.

Mirrors models/mobilenet_supernet.py:58-176 (5-tuple settings `[t, c, n, s, ks]`, `Model`) and
models/searched_network.py:14-147 (6-tuple settings `[c, n, s, ks, hiddens, expand]`): same
constructor keywords, same `features` / `classifier` children and hence state_dict keys, same
attributes read by utils/prune.py and mobilenet_base.output_network.  `forward` hands the
whole network to `engine.Engine` (one autograd node for the full forward/backward).
"""
import numbers
import warnings

from torch import nn

from .mobilenet_base import ConvBNReLU, _get_named_block_list, _make_divisible, get_active_fn, get_block

__all__ = ["MobileNetV2", "MobileNetSearched", "Model"]


def _hidden_dims(inp, expand_ratio, kernel_sizes):
    """expand_ratio (number or per-branch list) -> (hidden widths, expand?)  (:30-41)."""
    if isinstance(expand_ratio, list):
        assert len(expand_ratio) == len(kernel_sizes)
        ratios, expand = expand_ratio, True
    elif isinstance(expand_ratio, numbers.Number):
        ratios, expand = [expand_ratio] * len(kernel_sizes), expand_ratio != 1
    else:
        raise ValueError("Unknown expand_ratio type: {}".format(expand_ratio))
    return [int(round(inp * e)) for e in ratios], expand


def get_block_wrapper(block_str):
    """Block subclass taking `expand_ratio` instead of explicit hidden widths."""
    base = get_block(block_str)

    class InvertedResidual(base):

        def __init__(self, inp, oup, stride, expand_ratio, kernel_sizes, active_fn=None,
                     batch_norm_kwargs=None):
            hidden, expand = _hidden_dims(inp, expand_ratio, kernel_sizes)
            super().__init__(inp, oup, stride, hidden, kernel_sizes, expand, active_fn=active_fn,
                             batch_norm_kwargs=batch_norm_kwargs)
            self.expand_ratio = expand_ratio

    return InvertedResidual


class _Net(nn.Module):
    """Common tail of both builders: stem / blocks / head / pool / classifier + engine forward."""

    def _assemble(self, blocks, input_channel, head_in, last_channel, num_classes, input_size,
                  dropout_ratio, active_fn, bn_kw):
        feats = [ConvBNReLU(3, input_channel, stride=2, batch_norm_kwargs=bn_kw, active_fn=active_fn)]
        feats.extend(blocks)
        feats.append(ConvBNReLU(head_in, last_channel, kernel_size=1, batch_norm_kwargs=bn_kw,
                                active_fn=active_fn))
        feats.append(nn.AvgPool2d(input_size // 32))
        self.features = nn.Sequential(*feats)
        self.classifier = nn.Sequential(nn.Dropout(dropout_ratio),
                                        nn.Linear(last_channel, num_classes))

    def get_named_block_list(self):
        return _get_named_block_list(self)

    def forward(self, x):
        from . import engine
        return engine.run_module(self, x)


class MobileNetV2(_Net):

    def __init__(self, num_classes=1000, input_size=224, input_channel=32, last_channel=1280,
                 width_mult=1.0, inverted_residual_setting=None, dropout_ratio=0.2,
                 batch_norm_momentum=0.1, batch_norm_epsilon=1e-5, active_fn="nn.ReLU6",
                 block="InvertedResidualChannels", round_nearest=8):
        super().__init__()
        bn_kw = {"momentum": batch_norm_momentum, "eps": batch_norm_epsilon}
        self.input_size, self.input_channel, self.last_channel = input_size, input_channel, last_channel
        self.num_classes, self.width_mult, self.round_nearest = num_classes, width_mult, round_nearest
        self.inverted_residual_setting = inverted_residual_setting
        self.active_fn, self.block, self.batch_norm_kwargs = active_fn, block, bn_kw
        if len(inverted_residual_setting) == 0 or len(inverted_residual_setting[0]) != 5:
            raise ValueError("inverted_residual_setting should be non-empty or a 5-element list, "
                             "got {}".format(inverted_residual_setting))
        if input_size % 32 != 0:
            raise ValueError("Input size must divide 32")
        act = get_active_fn(active_fn)
        block_cls = get_block_wrapper(block)
        c_in = _make_divisible(input_channel * width_mult, round_nearest)
        c_last = _make_divisible(last_channel * max(1.0, width_mult), round_nearest)
        stem_c, blocks = c_in, []
        for t, c, n, s, ks in inverted_residual_setting:
            c_out = _make_divisible(c * width_mult, round_nearest)
            for i in range(n):
                blocks.append(block_cls(c_in, c_out, s if i == 0 else 1, t, ks, active_fn=act,
                                        batch_norm_kwargs=bn_kw))
                c_in = c_out
        self._assemble(blocks, stem_c, c_in, c_last, num_classes, input_size, dropout_ratio, act, bn_kw)


class MobileNetSearched(_Net):

    def __init__(self, num_classes=1000, input_size=224, input_channel=32, last_channel=1280,
                 width_mult=1.0, inverted_residual_setting=None, dropout_ratio=0.2, se_ratio=None,
                 batch_norm_momentum=0.1, batch_norm_epsilon=1e-5, active_fn="nn.ReLU6",
                 block="InvertedResidualChannels", round_nearest=8):
        super().__init__()
        bn_kw = {"momentum": batch_norm_momentum, "eps": batch_norm_epsilon}
        if width_mult != 1.0:
            raise ValueError("Searched model should have width 1")
        self.input_size, self.input_channel, self.last_channel = input_size, input_channel, last_channel
        self.num_classes, self.width_mult, self.round_nearest = num_classes, width_mult, round_nearest
        self.inverted_residual_setting = inverted_residual_setting
        self.active_fn, self.block, self.batch_norm_kwargs = active_fn, block, bn_kw
        if len(inverted_residual_setting) == 0 or len(inverted_residual_setting[0]) != 6:
            raise ValueError("inverted_residual_setting should be non-empty or a 6-element list, "
                             "got {}".format(inverted_residual_setting))
        if input_size % 32 != 0:
            raise ValueError("Input size must divide 32")
        for what, ch in (("Input", input_channel), ("Last", last_channel)):
            if (ch * width_mult) % round_nearest:
                warnings.warn("{} channel could not divide {}".format(what, round_nearest))
        act = get_active_fn(active_fn)
        block_cls = get_block(block)
        extra = {}
        if se_ratio is not None:
            if getattr(block_cls, "supports_se", False):
                extra["se_ratio"] = se_ratio
            else:
                raise NotImplementedError("SE module not supported for block: {}".format(block_cls))
        c_in, blocks = input_channel, []
        for c, n, s, ks, hiddens, expand in inverted_residual_setting:
            for i in range(n):
                blocks.append(block_cls(c_in, c, s if i == 0 else 1, hiddens, ks, expand,
                                        active_fn=act, batch_norm_kwargs=bn_kw, **extra))
                c_in = c
        self._assemble(blocks, input_channel, c_in, last_channel, num_classes, input_size,
                       dropout_ratio, act, bn_kw)


Model = MobileNetV2
