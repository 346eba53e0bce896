"""Resource-aware atomic-block selection (host logic; mirrors utils/prune.py of the reference).

Same public names and semantics as the reference module so train.py can import it unchanged:
PruneInfo (:11-86), get_bn_to_prune (:89-158), cal_bn_l1_loss (:161-167),
cal_mask_network_slimming_by_flops (:170-187) / _by_threshold (:190-195), cal_pruned_flops
(:198-212), get_rho_scheduler (:215-245), output_searched_network (:248-289).

Everything here is integer / scalar bookkeeping on tiny tensors; the per-step L1 term itself is
applied analytically inside the fused optimizer kernel (rho * penalty * sign(gamma)), see
trainer.Trainer and rmsprop.RMSprop.
"""
import collections
import itertools
import logging

import torch

from . import mobilenet_base as mb


class PruneInfo(object):
    """Ordered {bn weight name: {'penalty', 'compress_masked', ...}} table."""

    def __init__(self, names, penalties):
        assert len(names) == len(penalties)
        self._info = collections.OrderedDict(
            (n, {"compress_masked": False, "penalty": p}) for n, p in zip(names, penalties))

    @property
    def weight(self):
        return list(self._info.keys())

    @property
    def penalty(self):
        return self.get_info_list("penalty")

    def add_info_list(self, name, values):
        assert len(values) == len(self._info)
        for entry, v in zip(self._info.values(), values):
            entry[name] = v

    def get_info_list(self, name):
        return [entry[name] for entry in self._info.values()]

    # ---- shrink protocol (models/compress_utils.py drives these) -------------------------------
    def compress_start(self):
        for entry in self._info.values():
            entry["compress_masked"] = False

    def compress_check_exist(self, info):
        return info["var_old_name"] in self._info

    def compress_mask(self, info, verbose=False):
        old, new = info["var_old_name"], info["var_new_name"]
        if verbose:
            logging.info("PruneInfo compress: {} -> {}".format(old, new))
        if self._info[old]["compress_masked"]:
            raise RuntimeError("May have dependencies in compress")
        if new in self._info and self._info[new]["compress_masked"]:
            raise RuntimeError("Compress {} twice".format(new))
        entry = self._info.pop(old)
        entry["compress_masked"] = True
        self._info[new] = entry

    def compress_drop(self, info, verbose=False):
        name = info["var_old_name"]
        if verbose:
            logging.info("PruneInfo drop: {}".format(name))
        if self._info[name]["compress_masked"]:
            return None
        return self._info.pop(name)


def _prunable(model, bn_prune_filter):
    """Yield (bn weight name, width, MACs per channel) for every prunable depthwise-BN gamma."""
    skip_first = bn_prune_filter.endswith("skip_expand1")
    for name, block in model.get_named_block_list().items():
        if not isinstance(block, mb.InvertedResidualChannels):
            continue
        if skip_first and not block.expand:
            continue
        named_bn = block.get_named_depthwise_bn(prefix=name)
        for op, (bn_name, bn) in zip(block.ops, named_bn.items()):
            width = bn.weight.numel()
            yield "{}.weight".format(bn_name), width, op.n_macs / width


def get_bn_to_prune(model, flags, verbose=True):
    """FLOPs-normalised L1 penalties: penalty_i = mpc_i / (sum_j n_j mpc_j / (sum_j n_j + 1e-5))."""
    mode = flags.get("bn_prune_filter", None)
    names, penalties, per_channel = [], [], []
    if mode in ("expansion_only", "expansion_only_skip_expand1"):
        rows = list(_prunable(model, mode))
        total = sum(n for _, n, _ in rows)
        norm = sum(n * mpc for _, n, mpc in rows) / (total + 1e-5)
        names = [r[0] for r in rows]
        per_channel = [r[2] for r in rows]
        penalties = [mpc / norm for mpc in per_channel]
    elif mode == "equal_penalty_skip_expand1":
        rows = list(_prunable(model, mode))
        names = [r[0] for r in rows]
        per_channel = [r[2] for r in rows]
        penalties = [1 for _ in rows]
    elif mode is not None:
        raise NotImplementedError()
    info = PruneInfo(names, penalties)
    info.add_info_list("per_channel_flops", per_channel)
    if verbose:
        for n, p in zip(info.weight, info.penalty):
            logging.info("{} penalty: {}".format(n, p))
    known = set(k for k, _ in model.named_parameters())
    assert all(n in known for n in info.weight)
    return info


def cal_bn_l1_loss(bn_weights, penalties, rho):
    """sum_i rho * penalty_i * |gamma_i|_1 (autograd form; the fused step applies its gradient
    analytically instead)."""
    assert len(bn_weights) == len(penalties)
    loss = 0.0
    for w, p in zip(bn_weights, penalties):
        loss += rho * p * w.abs().sum()
    return loss


def cal_mask_network_slimming_by_flops(weights, prune_info, flops_to_prune, incremental=False):
    mags = [w.detach().abs() for w in weights]
    flat, order = torch.sort(torch.cat(mags))
    cost = torch.cat([torch.full_like(m, c) for m, c in
                      zip(mags, prune_info.get_info_list("per_channel_flops"))])[order]
    idx = torch.nonzero(torch.cumsum(cost, 0) > flops_to_prune)[0].item()
    threshold = flat[idx].item()
    return [m > threshold for m in mags], threshold


def cal_mask_network_slimming_by_threshold(weights, threshold):
    """alive = |gamma| > threshold (strict)."""
    return [w.detach().abs() > threshold for w in weights]


def cal_pruned_flops(prune_info):
    rows, pruned = [], 0
    for name, mpc, mask in zip(prune_info.weight, prune_info.get_info_list("per_channel_flops"),
                               prune_info.get_info_list("mask")):
        dead = (~mask.detach()).sum().item()
        total = mask.numel()
        rows.append([name, total, dead, total * mpc, dead * mpc, dead / total])
        pruned += dead * mpc
    return pruned, rows


def get_rho_scheduler(prune_params, steps_per_epoch):
    """Linear warm-up of the L1 weight: 0 until epoch_free, rho from epoch_warmup on."""
    free = prune_params["epoch_free"] * steps_per_epoch
    warm = prune_params["epoch_warmup"] * steps_per_epoch
    rho, stepwise = prune_params["rho"], prune_params["stepwise"]
    if prune_params["scheduler"] != "linear":
        raise ValueError("Unknown sparsity scheduler {}".format(prune_params["scheduler"]))

    def at(i):
        if not stepwise:
            i = (i // steps_per_epoch) * steps_per_epoch
        if i < free:
            return 0.0
        if i >= warm:
            return rho
        return (i - free) / (warm - free) * rho

    return at


def output_searched_network(model, infos, flags):
    """6-tuple network description with the dead atomic blocks removed."""
    kw = {k: getattr(model, k) for k in
          ("input_channel", "last_channel", "width_mult", "round_nearest", "active_fn", "num_classes")}
    setting = model.inverted_residual_setting
    blocks = list(model.get_named_block_list().values())
    rows = []
    if "skip_expand1" in flags.get("bn_prune_filter", None):
        t, c, n, s, ks = setting[0]
        assert t == 1 and n == 1 and len(ks) == 1 and ks[0] == 3
        rows.append([c, n, s, ks, [kw["input_channel"]], False])
        blocks = blocks[n:]
    it = iter(infos)
    used = 0
    for block in blocks:
        remain = []
        for k, c in zip(block.kernel_sizes, block.channels):
            info = next(it)
            used += 1
            assert c == info[1], "{}, {}, {}, {}".format(block, k, c, str(info))
            remain.append(info[1] - info[2])
        alive = [c != 0 for c in remain]
        rows.append([block.output_dim, 1, block.stride,
                     list(itertools.compress(block.kernel_sizes, alive)),
                     list(itertools.compress(remain, alive)), block.expand])
    assert used == len(infos)
    kw["inverted_residual_setting"] = rows
    return kw


def profile_macs(model, input_size):
    """Set `n_macs` on every block / branch / conv exactly as utils/model_profiling.py:82-127 does
    (integer formula Cin*Cout*k*k*Ho*Wo // groups, batch 1), without running a forward pass."""
    feats = list(model.features.children())
    h = (input_size + 1) // 2

    def conv_macs(conv, ho):
        kh, kw = conv.kernel_size
        return conv.in_channels * conv.out_channels * kh * kw * ho * ho // conv.groups

    total = conv_macs(feats[0][0], h)
    feats[0][0].n_macs = total
    for block in feats[1:-2]:
        ho = (h + 1) // 2 if block.stride == 2 else h
        block.n_macs = 0
        if hasattr(block, "depth_ops"):                        # InvertedResidualChannelsFused (:121-130)
            parts = []
            if block.expand:
                block.expand_conv[0].n_macs = conv_macs(block.expand_conv[0], h)
                block.expand_conv.n_macs = block.expand_conv[0].n_macs
                parts.append(block.expand_conv.n_macs)
            for op in block.depth_ops:
                conv = op[-1][0]
                conv.n_macs = conv_macs(conv, ho)
                op.n_macs = conv.n_macs
                parts.append(op.n_macs)
            proj = block.project_conv[0]
            proj.n_macs = conv_macs(proj, ho)
            block.project_conv.n_macs = proj.n_macs
            parts.append(proj.n_macs)
            se = block.se_op
            if hasattr(se, "se_reduce"):
                se.se_reduce.n_macs = conv_macs(se.se_reduce, 1)
                se.se_expand.n_macs = conv_macs(se.se_expand, 1)
                se.n_macs = se.n_feature * ho * ho + se.se_reduce.n_macs + se.se_expand.n_macs
                parts.append(se.n_macs)
            block.n_macs = sum(parts)
            total += block.n_macs
            h = ho
            continue
        for op in block.ops:
            d = 1 if block.expand else 0
            convs = [(op[d][0], ho), (op[d + 1], ho)]          # depthwise, projection: output res
            if block.expand:
                convs.append((op[0][0], h))                    # 1x1 expand runs at the input res
            op.n_macs = 0
            for conv, res in convs:
                conv.n_macs = conv_macs(conv, res)
                op.n_macs += conv.n_macs
            block.n_macs += op.n_macs
        total += block.n_macs
        h = ho
    head = feats[-2][0]
    head.n_macs = conv_macs(head, h)
    fc = model.classifier[1]
    total += head.n_macs + head.out_channels * h * h + fc.in_features * fc.out_features
    model.n_macs = total
    return total
