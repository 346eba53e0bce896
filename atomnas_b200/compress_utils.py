"""Dynamic network shrinkage: physically drop dead atomic blocks (mirror of models/compress_utils.py).

Index semantics are the reference's and are bit exact: `new = old[mask]` along dim 0 (expand conv
/ BN, depthwise conv / BN and their optimizer / EMA state) or dim 1 (projection conv), dead
branches are removed and the survivors renumbered `ops.{old}` -> `ops.{new}`; optimizer, EMA and
PruneInfo are kept in lock-step through their `compress_mask` / `compress_drop` callbacks
(models/compress_utils.py:180-301, _mask_along_dim :31-37).  On CUDA tensors the masked copy is the
an_gather_rows kernel (with idx = nonzero(mask), i.e. exactly `rhs[mask]` order).

After a shrink the engine that owned the block is dropped; the next forward rebuilds arenas,
kernel-layout weights and launch plan for the new widths (widths that are not multiples of 8
run in padded channel slots).  Nothing is recompiled.
"""
import functools
import itertools
import warnings

import torch
import torch.nn as nn

from . import kernels as K


def add_prefix(name, prefix=None, split="."):
    return name if prefix is None else "{}{}{}".format(prefix, split, name)


def _scatter_by_bool(values, mask, pad=None):
    it = iter(values)
    out = [next(it) if keep else pad for keep in mask]
    assert next(it, None) is None
    return out


def _find_only_one(func, iterable):
    hits = [x for x in iterable if func(x)]
    assert len(hits) == 1
    return hits[0]


def _masked_rows(src, mask):
    """src[mask] along dim 0, via the gather kernel on the GPU."""
    if not src.is_cuda:
        return src[mask]
    idx = torch.nonzero(mask.to(src.device)).flatten()
    flat = src.contiguous().view(src.shape[0], -1)
    if flat.dtype != torch.float32 or idx.numel() == 0:
        return src[mask.to(src.device)]
    out = torch.empty((idx.numel(), flat.shape[1]), device=src.device)
    K.gather_rows(flat, idx, idx.numel(), flat.shape[1], out)
    return out.view((idx.numel(),) + tuple(src.shape[1:]))


def _mask_along_dim(lhs, rhs, mask, dim=0):
    with torch.no_grad():
        if dim == 0:
            lhs.data.copy_(_masked_rows(rhs.data, mask))
        elif dim == 1:
            picked = _masked_rows(rhs.data.transpose(0, 1).contiguous(), mask)
            lhs.data.copy_(picked.transpose(0, 1))
        else:
            raise NotImplementedError()


def _copy(lhs, rhs, *args):
    lhs.data.copy_(rhs.data)


def build_default_info(m_new, m_old, mask, attr, mask_hook, var_type="variable", prefix_new=None,
                       prefix_old=None):
    assert var_type in ("variable", "buffer")
    info = {"var_old_name": add_prefix(attr, prefix_old), "var_old": getattr(m_old, attr),
            "type": var_type, "mask": mask, "module_class": type(m_old)}
    if m_new is not None:
        info.update({"var_new_name": add_prefix(attr, prefix_new), "var_new": getattr(m_new, attr),
                     "mask_hook": mask_hook})
    return info


def compress_conv(m_new, m_old, mask, dim, prefix_new=None, prefix_old=None):
    assert m_new is None or isinstance(m_new, nn.Conv2d)
    assert isinstance(m_old, nn.Conv2d) and dim in (0, 1)
    mk = functools.partial(build_default_info, m_new, m_old, mask, prefix_new=prefix_new, prefix_old=prefix_old)
    infos = [mk("weight", functools.partial(_mask_along_dim, dim=dim))]
    if m_old.bias is not None:
        infos.append(mk("bias", _mask_along_dim if dim == 0 else _copy))
    return infos


def compress_bn(m_new, m_old, mask, prefix_new=None, prefix_old=None):
    assert m_new is None or isinstance(m_new, nn.BatchNorm2d)
    assert isinstance(m_old, nn.BatchNorm2d)
    assert m_new is None or m_new.affine == m_old.affine
    infos = []
    if m_old.affine:
        mk = functools.partial(build_default_info, m_new, m_old, mask, prefix_new=prefix_new, prefix_old=prefix_old)
        infos += [mk("weight", _mask_along_dim), mk("bias", _mask_along_dim)]
    if m_old.track_running_stats:
        mk = functools.partial(build_default_info, m_new, m_old, mask, var_type="buffer",
                               prefix_new=prefix_new, prefix_old=prefix_old)
        infos += [mk("running_var", _mask_along_dim), mk("running_mean", _mask_along_dim),
                  mk("num_batches_tracked", _copy)]
    return infos


def adjust_bn(m_new, m_old, post_hook_params, **kwargs):
    """pw_bn is copied whole (all-true mask).  The reference keeps a running-mean correction hook
    here but disables it ("Do not adjust bn mean!!!", compress_utils.py:201-204); BN calibration
    before every evaluation repairs the statistics instead."""
    mask = torch.ones_like(m_new.weight, dtype=torch.bool)
    infos = compress_bn(m_new, m_old, mask, **kwargs)
    info_mean = _find_only_one(lambda i: "running_mean" in i["var_old_name"], infos)
    info_mean["post_hook_params"] = post_hook_params
    info_mean["post_hook"] = None
    return infos


def compress_conv_bn_relu(m_new, m_old, mask, prefix_new=None, prefix_old=None, dim=0):
    from . import mobilenet_base as mb
    assert m_new is None or isinstance(m_new, mb.ConvBNReLU)
    assert isinstance(m_old, mb.ConvBNReLU)
    old = list(m_old.children())
    new = [None] * len(old) if m_new is None else list(m_new.children())
    return (compress_conv(new[0], old[0], mask, dim=dim, prefix_new="{}.0".format(prefix_new),
                          prefix_old="{}.0".format(prefix_old)) +
            compress_bn(new[1], old[1], mask, prefix_new="{}.1".format(prefix_new),
                        prefix_old="{}.1".format(prefix_old)))


def copmress_inverted_residual_channels(m, masks, ema=None, optimizer=None, prune_info=None, prefix=None,
                                        verbose=False):
    """(sic -- the reference spells it this way.)  Rebuild `m.ops` / `m.pw_bn` for the alive atoms."""
    from . import engine as E

    def is_bn_param(info):
        return issubclass(info["module_class"], nn.BatchNorm2d) and info["type"] == "variable"

    def apply(infos, drop):
        for info in infos:
            if optimizer is not None and info["type"] != "buffer":
                (optimizer.compress_drop if drop else optimizer.compress_mask)(info, verbose=verbose)
            if ema is not None and "num_batches_tracked" not in info["var_old_name"]:
                (ema.compress_drop if drop else ema.compress_mask)(info, verbose=verbose)
            if prune_info is not None and is_bn_param(info) and prune_info.compress_check_exist(info):
                (prune_info.compress_drop if drop else prune_info.compress_mask)(info, verbose=verbose)
            if not drop:
                info["mask_hook"](info["var_new"], info["var_old"], info["mask"])
                if "post_hook" in info:
                    warnings.warn("Do not adjust bn mean!!!")

    assert len(m.kernel_sizes) == len(masks)
    device = m.pw_bn.weight.device
    widths = [int(mask.detach().sum().item()) for mask in masks]
    keeps = [w > 0 for w in widths]
    m.channels = list(itertools.compress(widths, keeps))
    m.kernel_sizes = list(itertools.compress(m.kernel_sizes, keeps))
    new_ops, new_pw_bn = m._build(m.channels, m.kernel_sizes, m.expand)
    new_ops.to(device)
    new_pw_bn.to(device)
    idx_depth, idx_proj = (1, 2) if m.expand else (0, 1)
    padded_ops = _scatter_by_bool(list(new_ops), keeps)
    padded_idx = _scatter_by_bool(list(range(len(new_ops))), keeps)

    keep_infos, drop_infos, adjust = [], [], []
    for new_op, new_i, (old_i, old_op), mask in zip(padded_ops, padded_idx, enumerate(m.ops), masks):
        old_children = list(old_op.children())
        new_children = [None] * len(old_children) if new_op is None else list(new_op.children())
        bucket = drop_infos if new_op is None else keep_infos

        def names(sub):
            return (add_prefix("ops.{}.{}".format(new_i, sub), prefix),
                    add_prefix("ops.{}.{}".format(old_i, sub), prefix))

        if m.expand:
            bucket.append(compress_conv_bn_relu(new_children[0], old_children[0], mask, *names(0)))
        depth = compress_conv_bn_relu(new_children[idx_depth], old_children[idx_depth], mask, *names(idx_depth))
        bucket.append(depth)
        pn, po = names(idx_proj)
        proj = compress_conv(new_children[idx_proj], old_children[idx_proj], mask, dim=1, prefix_new=pn,
                             prefix_old=po)
        bucket.append(proj)
        adjust.append({"active_fn": m.active_fn,
                       "bias": _find_only_one(lambda i: is_bn_param(i) and "bias" in i["var_old_name"], depth),
                       "following_conv_weight": _find_only_one(
                           lambda i: issubclass(i["module_class"], nn.Conv2d) and "weight" in i["var_old_name"], proj)})
    pw_name = add_prefix("pw_bn", prefix)
    pw_infos = adjust_bn(new_pw_bn, m.pw_bn, adjust, prefix_new=pw_name, prefix_old=pw_name)

    if ema is not None:
        ema.compress_start()
    if prune_info is not None:
        prune_info.compress_start()
    apply(pw_infos, drop=False)            # pw_bn first (EMA name bookkeeping depends on the order)
    for infos in keep_infos:
        apply(infos, drop=False)
    for infos in drop_infos:               # dead branches last
        apply(infos, drop=True)

    E.invalidate(m)
    del m.ops
    del m.pw_bn
    m.ops, m.pw_bn = new_ops, new_pw_bn
