"""Data-parallel backend: one process per GPU, NCCL over NVLink (mirror of utils/distributed.py).

Same entry points as the reference (init_dist :25-32, AllReduceDistributedDataParallel :172-199,
allreduce_grads :155-161, allreduce_bn :164-169, is_master / master_only :50-80, world-size
helpers) so common.py / train.py run unchanged.  The reference flattens ~500 gradient tensors
into a fresh buffer, all-reduces it and copies back (:131-139); here the gradients already live
in ONE arena, so `allreduce_grads` is a single in-place NCCL all-reduce with no flatten / unflatten
copies (and trainer.Trainer overlaps it with backward in buckets).  BatchNorm statistics stay
per-GPU by design (allreduce_bn is the optional running-stat averaging used after calibration).
"""
import functools
import os

import torch
import torch.distributed as dist
import torch.nn as nn
from torch.distributed import get_rank, get_world_size  # noqa: F401  (re-exported like the reference)

from . import engine as E


def _get_env(name):
    if name not in os.environ:
        raise RuntimeError("${} should be set".format(name))
    return os.environ[name]


def init_dist(backend="nccl", **kwargs):
    if dist.is_initialized():
        raise RuntimeError("Should not init distributed twice")
    rank, local_rank = int(_get_env("RANK")), int(_get_env("LOCAL_RANK"))
    if torch.cuda.is_available():
        assert rank % torch.cuda.device_count() == local_rank
        torch.cuda.set_device(local_rank)
    dist.init_process_group(backend=backend, **kwargs)


def assert_initialized():
    if not dist.is_initialized():
        raise RuntimeError("Default process group is not initialized")


def get_local_rank():
    assert_initialized()
    return int(_get_env("LOCAL_RANK"))


def get_local_size():
    assert_initialized()
    return torch.cuda.device_count()


def get_rank_fallback():
    return get_rank() if dist.is_initialized() else 0


def get_world_size_fallback():
    return get_world_size() if dist.is_initialized() else 1


def is_master():
    return get_rank_fallback() == 0


def master_only(func):
    @functools.wraps(func)
    def wrapper(*args, **kwargs):
        return func(*args, **kwargs) if is_master() else None
    return wrapper


def dist_reduce_tensor(tensor, dst=0):
    world = get_world_size()
    if world < 2:
        return tensor
    with torch.no_grad():
        dist.reduce(tensor, dst=dst)
        if get_rank() == dst:
            tensor.div_(world)
    return tensor


def dist_all_reduce_tensor(tensor):
    world = get_world_size()
    if world < 2:
        return tensor
    with torch.no_grad():
        dist.all_reduce(tensor)
        tensor.div_(world)
    return tensor


def _mean_allreduce_flat(flat, world):
    dist.all_reduce(flat)
    flat.div_(world)


def _allreduce(tensors, coalesce=True, bucket_size_mb=-1):
    """Mean over ranks of every tensor.  Tensors that are views of one arena are reduced in place
    as a single range; anything else goes through a flat staging buffer per dtype."""
    world = get_world_size()
    if not tensors:
        return
    by_storage = {}
    for t in tensors:
        by_storage.setdefault((t.untyped_storage().data_ptr(), t.dtype), []).append(t)
    loose = {}                                   # tensors with private storage, coalesced per dtype
    for (_, dtype), group in by_storage.items():
        if len(group) > 1 and all(t.is_contiguous() for t in group):
            base = min(t.data_ptr() for t in group)
            end = max(t.data_ptr() + t.numel() * t.element_size() for t in group)
            first = min(group, key=lambda t: t.data_ptr())
            n = (end - base) // first.element_size()
            flat = first.as_strided((n,), (1,), first.storage_offset())
            _mean_allreduce_flat(flat, world)            # the arena range itself: no copies
        else:
            loose.setdefault(dtype, []).extend(group)
    for dtype, group in loose.items():           # ONE flat buffer per dtype (utils/distributed.py:131-139)
        flat = torch.cat([t.reshape(-1) for t in group])
        _mean_allreduce_flat(flat, world)
        off = 0
        for t in group:
            t.copy_(flat[off:off + t.numel()].view_as(t))
            off += t.numel()


def allreduce_grads(model, *args, **kwargs):
    grads = [p.grad.data for p in model.parameters() if p.requires_grad and p.grad is not None]
    _allreduce(grads, *args, **kwargs)


def allreduce_bn(model, *args, **kwargs):
    stats = [b for n, b in model.named_buffers() if "running_var" in n or "running_mean" in n]
    _allreduce(stats, *args, **kwargs)


def _broadcast_all(tensors):
    for t in tensors:
        dist.broadcast(t, 0)


class AllReduceDistributedDataParallel(nn.Module):
    """Thin wrapper: broadcast rank 0's state at construction, forward to `.module`; gradients are
    reduced explicitly by `allreduce_grads` after backward (train.py:184-185)."""

    def __init__(self, module, dim=0, broadcast_buffers=True, bucket_cap_mb=25):
        super(AllReduceDistributedDataParallel, self).__init__()
        self.module, self.dim, self.broadcast_buffers = module, dim, broadcast_buffers
        self.broadcast_bucket_size_mb = bucket_cap_mb
        self._sync_params()

    def _sync_params(self):
        if not dist.is_initialized() or get_world_size() < 2:
            return
        eng = None
        p0 = next(self.module.parameters(), None)
        if p0 is not None and p0.is_cuda:
            try:
                eng = E.get_engine(self.module)
            except (NotImplementedError, RuntimeError):
                eng = None
        if eng is not None:                      # three arena broadcasts instead of ~1000 tensors
            _broadcast_all([eng.P, eng.S, eng.NBT])
            eng.refresh_weights(force=True)
        else:
            _broadcast_all([t.data for t in self.module.state_dict().values()])

    def forward(self, *inputs, **kwargs):
        return self.module(*inputs, **kwargs)
