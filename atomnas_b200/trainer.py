"""Fused training step: the public fast path of the package.

One call = train.py:174-236 of the reference for one batch: forward, label-smoothed CE, backward,
gradient all-reduce (NCCL over NVLink), RMSprop with the analytic L2 / L1 terms, EMA of parameters
and BN running statistics, refresh of the bf16 weight copies -- with no per-parameter Python loop
and no host synchronisation.

    trainer = Trainer(model, lr=0.128, weight_decay=1e-5, ema_decay=0.9999 ** 0.5)
    loss = trainer.step_host(images_pinned, labels_pinned)     # H2D + step; device scalar
    trainer.shrink(threshold=1e-3, prune_info=info)            # train.py:27-81 on the fast path

The drop-in route (reference train.py + utils.rmsprop.RMSprop + ExponentialMovingAverage) uses
the same engine through atomnas_b200.rmsprop / atomnas_b200.optim instead.

CUDA graphs.  One rank: the whole step is one graph.  Several ranks: the step is two graphs
(forward + backward, optimizer) around ONE in-place NCCL all-reduce of the gradient arena issued
from Python (45 MB over NVSwitch is ~0.2 ms against a >30 ms step, so overlapping it buys < 1 %:
measured 53.8 ms/step on 2 GPUs vs 53.5 on one).  Capturing the NCCL all-reduce INSIDE the graph was
tried and hangs with torch 2.11 / NCCL 2.28 async process groups (tests/test_multirank_gpu.py,
round 2), so it is not offered.  Eager mode (use_cuda_graph=False) keeps the bucketed overlap.
"""
import os
import time

import torch
import torch.distributed as dist

from . import engine as E
from . import kernels as K


class _ShrinkRecorder(object):
    """Stands in for `optimizer` and `ema` in compress_by_mask (models/compress_utils.py:188-217):
    records which new tensor is which old tensor under which mask, so the flat optimizer / EMA
    arenas can be repacked in one pass afterwards."""

    def __init__(self):
        self.masked, self.dropped = {}, {}

    def compress_start(self):
        pass

    def compress_mask(self, info, verbose=False):
        self.masked[id(info["var_new"])] = info

    def compress_drop(self, info, verbose=False):
        self.dropped[id(info["var_old"])] = info


class Trainer(object):

    def __init__(self, model, lr, alpha=0.9, momentum=0.9, eps=1e-3, eps_inside_sqrt=True,
                 weight_decay=1e-5, label_smoothing=0.1, ema_decay=None, prune_penalties=None,
                 bucket_bytes=8 << 20, use_cuda_graph=False, weight_decay_method="mnas", distributed=True):
        self.model = model
        self.lr, self.alpha, self.momentum, self.eps = lr, alpha, momentum, eps
        self.eps_inside_sqrt = eps_inside_sqrt
        self.weight_decay, self.label_smoothing = weight_decay, label_smoothing
        self.weight_decay_method = weight_decay_method
        self.ema_decay = ema_decay
        self.num_updates = 0
        self.world = dist.get_world_size() if distributed and dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.bucket_elems = max(1, bucket_bytes // 4)
        self.use_cuda_graph = use_cuda_graph
        self._bind(E.get_engine(model))
        self.sq = torch.zeros_like(self.eng.P)
        self.buf = torch.zeros_like(self.eng.P)
        self.ema_P = self.eng.P.clone() if ema_decay else None
        self.ema_S = self.eng.S.clone() if ema_decay else None
        self._penalties = None
        if prune_penalties:
            self.set_prune_penalties(prune_penalties)
        self.hp_dev = torch.zeros(4, device=self.eng.dev)                       # lr, rho, ema decay
        self.seed_dev = torch.zeros(1, device=self.eng.dev, dtype=torch.int64)  # dropout stream position
        self._x_dev = None
        self._y_dev = None

    def _bind(self, eng):
        if not eng.is_net:
            raise ValueError("Trainer needs a whole network")
        self.eng = eng
        eng.set_weight_decay_method(self.weight_decay_method)
        self.launches_per_step = None
        self._handles = []
        self._pending_hi = None
        self._graph = None
        self._graphs2 = None
        self._static = None

    def set_prune_penalties(self, named_penalties):
        """[(bn weight name, penalty)] of utils/prune.py:112-123 -> per-element L1 weights."""
        self._penalties = list(named_penalties)
        self.eng.set_l1_penalties(self._penalties, dict(self.model.named_parameters()))

    # ---- gradient all-reduce, bucketed in reverse arena order as backward releases ranges --------
    def _grad_ready(self, offset):
        if self.world == 1:
            return
        if self._pending_hi is None:
            self._pending_hi = self.eng.n_params
        if self._pending_hi - offset >= self.bucket_elems or offset == 0:
            seg = self.eng.G[offset:self._pending_hi]
            self._handles.append(dist.all_reduce(seg, async_op=True))
            self._pending_hi = offset

    def _finish_allreduce(self):
        for h in self._handles:
            h.wait()
        self._handles = []
        self._pending_hi = None

    # ---- one step on device tensors ---------------------------------------------------------------
    def _push_hparams(self, rho, lr):
        """Per-step scalars live in device memory ({lr, rho, ema_decay}) so the same launch
        sequence -- or a captured CUDA graph of it -- serves every step of the schedules."""
        self.num_updates += 1
        d = 0.0
        if self.ema_decay:
            d = min(self.ema_decay, (1.0 + self.num_updates) / (10.0 + self.num_updates))
        self.hp_dev.copy_(torch.tensor([lr, rho, d, 0.0]))

    def _fwd_bwd(self, x, target, overlap):
        eng = self.eng
        self.seed_dev.add_(1)                      # dropout stream advances on the device
        eng.forward(x, True, True, seed=0x5EED + 7919 * self.rank, seed_dev=self.seed_dev)
        loss = eng.loss_and_dlogits(target, self.label_smoothing, 1.0)
        eng.backward(None, grad_ready=self._grad_ready if overlap else None)
        return loss

    def _optimizer(self):
        eng = self.eng
        K.rmsprop_ema_step(eng.P, eng.G, self.sq, self.buf, self.ema_P, eng.wd_flag, eng.l1_pen, eng.n_params,
                           0.0, self.alpha, self.momentum, self.eps, self.eps_inside_sqrt, 1.0 / self.world,
                           self.weight_decay, 0.0, 0.0, hparams_dev=self.hp_dev)
        if self.ema_decay:
            K.ema_update(self.ema_S, eng.S, eng.S.numel(), 0.0, hparams_dev=self.hp_dev)
        eng.refresh_weights(force=True)

    def _step_impl(self, x, target):
        loss = self._fwd_bwd(x, target, overlap=True)
        self._finish_allreduce()
        self._optimizer()
        return loss

    def _check_engine(self):
        if E.get_engine(self.model) is not self.eng:
            raise RuntimeError("atomnas_b200.Trainer: the model was rebuilt (shrunk) outside Trainer.shrink(); the "
                               "fast path's optimizer / EMA arenas no longer match it -- use trainer.shrink(...)")

    def step(self, x, target, rho=0.0, lr=None):
        """x: fp32 NCHW on the device, target: int64 labels.  Returns per-sample losses (device)."""
        self._check_engine()
        lr = self.lr if lr is None else lr
        self._push_hparams(rho, lr)
        if not self.use_cuda_graph:
            return self._step_impl(x, target)
        return self._step_graphed(x, target)

    def _record(self, x, target, split):
        """Capture the step (records only, executes nothing; every kernel has run eagerly before) as
        one graph, or as two (forward+backward | optimizer) with the all-reduce issued between."""
        from ._lib import lib
        sx, sy = x.clone(), target.clone()
        torch.cuda.synchronize()
        c0 = lib().an_launch_count()
        if split:
            g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1):
                out = self._fwd_bwd(sx, sy, overlap=False)
            with torch.cuda.graph(g2, pool=g1.pool()):
                self._optimizer()
            self._graphs2, self._graph = (g1, g2), None
        else:
            self._graph, self._graphs2 = torch.cuda.CUDAGraph(), None
            with torch.cuda.graph(self._graph):
                out = self._step_impl(sx, sy)
        self.launches_per_step = lib().an_launch_count() - c0     # kernel nodes replayed per step
        self._static = (sx, sy, tuple(x.shape), out, split)

    def _capture(self, x, target, split):
        """First graphed step of a shape: run it eagerly (its result is returned), then record."""
        first = self._step_impl(x, target).clone()
        self._record(x, target, split)
        return first

    def _step_graphed(self, x, target):
        """Step replayed from CUDA graphs (the ~900 launches of a step are launch-bound from
        Python); lr / rho / EMA decay / dropout seed are read from device memory."""
        split = self.world > 1
        if self._static is None or self._static[2] != tuple(x.shape) or self._static[4] != split:
            return self._capture(x, target, split)
        sx, sy, _, out, _ = self._static
        sx.copy_(x, non_blocking=True)
        sy.copy_(target, non_blocking=True)
        if split:
            self._graphs2[0].replay()
            dist.all_reduce(self.eng.G)                    # in place on the arena; 1/world is in the optimizer
            self._graphs2[1].replay()
        else:
            self._graph.replay()
        return out

    def step_accum(self, batches, rho=0.0, lr=None, _fetch=None):
        """(S) series of SURVEY 8(d): gradient accumulation over len(batches) micro-batches (BN sees
        one micro-batch at a time, as per-GPU BN does in the reference), then ONE all-reduce and ONE
        optimizer / EMA step.  Returns the mean loss over the micro-batches (device scalar)."""
        self._check_engine()
        self._push_hparams(rho, self.lr if lr is None else lr)
        eng, n = self.eng, len(batches)
        if getattr(self, "_gacc", None) is None or self._gacc.shape != eng.G.shape:
            self._gacc = torch.zeros_like(eng.G)
        total = None
        ready = (self.use_cuda_graph and self._static is not None and self._static[2] == tuple(batches[0][0].shape)
                 and self._static[4])
        for i in range(n):
            x, y = batches[i] if _fetch is None else _fetch(i)
            if ready:
                sx, sy, _, out, _ = self._static
                sx.copy_(x, non_blocking=True)
                sy.copy_(y, non_blocking=True)
                self._graphs2[0].replay()
            else:
                out = self._fwd_bwd(x, y, overlap=False)
            if i == 0:
                self._gacc.copy_(eng.G)
            else:
                self._gacc.add_(eng.G)
            total = out.mean() if total is None else total + out.mean()
        torch.mul(self._gacc, 1.0 / n, out=eng.G)
        if self.world > 1:
            dist.all_reduce(eng.G)
        if ready:
            self._graphs2[1].replay()
        else:
            self._optimizer()
            if self.use_cuda_graph:                        # this call ran eagerly; record for the next ones
                total = total.clone()
                self._record(x, y, True)
        return total / n

    def step_accum_host(self, host_batches, rho=0.0, lr=None):
        """step_accum from pinned host buffers: micro-batch i+1 is copied on a side stream into the
        other of two device buffers while micro-batch i computes."""
        dev = self.eng.dev
        x0, y0 = host_batches[0]
        if getattr(self, "_acc_dev", None) is None or self._acc_dev[0][0].shape != x0.shape:
            self._acc_dev = [(torch.empty(x0.shape, device=dev), torch.empty(y0.shape, device=dev, dtype=torch.int64))
                             for _ in range(2)]
            self._acc_stream = torch.cuda.Stream(device=dev)
        cur, cs = torch.cuda.current_stream(), self._acc_stream
        copied, consumed = {}, {}

        def issue_copy(i):
            xd, yd = self._acc_dev[i % 2]
            if i - 2 in consumed:
                cs.wait_event(consumed[i - 2])
            else:
                cs.wait_stream(cur)                        # buffers may still be read by the previous call
            with torch.cuda.stream(cs):
                xd.copy_(host_batches[i][0], non_blocking=True)
                yd.copy_(host_batches[i][1], non_blocking=True)
                copied[i] = torch.cuda.Event()
                copied[i].record(cs)

        def fetch(i):
            if i == 0:
                issue_copy(0)
            if i > 0:                                      # micro-batch i-1 has been enqueued: mark its buffer
                consumed[i - 1] = torch.cuda.Event()
                consumed[i - 1].record(cur)
            if i + 1 < len(host_batches):
                issue_copy(i + 1)
            cur.wait_event(copied[i])
            return self._acc_dev[i % 2]

        return self.step_accum(host_batches, rho, lr, _fetch=fetch)

    def step_host(self, x_host, y_host, rho=0.0, lr=None, prefetch=None):
        """End-to-end step from (pinned) host buffers: H2D copies, the step, mean loss (device).

        prefetch=(x_next, y_next) starts the NEXT batch's host->device copies on a side stream
        while this step computes, into the other of two device buffers -- the double-buffered
        side-stream input pipeline of the reference's DataPrefetcher (utils/dataflow.py:13-58).
        A later step_host call with those same host tensors picks the prefetched buffer up."""
        dev = self.eng.dev
        if self._x_dev is None or self._x_dev[0].shape != x_host.shape:
            self._x_dev = [torch.empty(x_host.shape, device=dev, dtype=torch.float32) for _ in range(2)]
            self._y_dev = [torch.empty(y_host.shape, device=dev, dtype=torch.int64) for _ in range(2)]
            self._pf_key, self._pf_slot = None, 0
            self._pf_ready = [None, None]          # copy finished (recorded on the copy stream)
            self._consumed = [None, None]          # step that read the buffer finished (compute stream)
            self._copy_stream = torch.cuda.Stream(device=dev)
        cur = torch.cuda.current_stream()
        key = (x_host.data_ptr(), y_host.data_ptr(), tuple(x_host.shape))
        if self._pf_key == key and self._pf_ready[self._pf_slot] is not None:
            slot = self._pf_slot
            cur.wait_event(self._pf_ready[slot])
        else:                                      # not prefetched: copy on the compute stream
            slot = 1 - self._pf_slot
            self._x_dev[slot].copy_(x_host, non_blocking=True)
            self._y_dev[slot].copy_(y_host, non_blocking=True)
        self._pf_key = None
        loss = self.step(self._x_dev[slot], self._y_dev[slot], rho, lr)
        ev = torch.cuda.Event()
        ev.record(cur)
        self._consumed[slot] = ev
        if prefetch is not None:
            xn, yn = prefetch
            other = 1 - slot
            cs = self._copy_stream
            if self._consumed[other] is not None:
                cs.wait_event(self._consumed[other])
            with torch.cuda.stream(cs):
                self._x_dev[other].copy_(xn, non_blocking=True)
                self._y_dev[other].copy_(yn, non_blocking=True)
                rdy = torch.cuda.Event()
                rdy.record(cs)
            self._pf_ready[other] = rdy
            self._pf_key = (xn.data_ptr(), yn.data_ptr(), tuple(xn.shape))
            self._pf_slot = other
        return loss.mean()

    def reg_losses(self, rho):
        """(l2, l1) regularisation loss values for logging (train.py:203-208)."""
        part = torch.zeros(128, 2, device=self.eng.dev)
        K.reg_losses(self.eng.P, self.eng.wd_flag, self.eng.l1_pen, self.eng.n_params, part)
        s = part.sum(0)
        return 0.5 * self.weight_decay * s[0], rho * s[1]

    # ---- dynamic network shrinkage on the fast path (train.py:27-81) -------------------------------
    def ema_named(self):
        """{state_dict-style name: EMA shadow view} of every parameter and BN running statistic."""
        if self.ema_P is None:
            return {}
        eng = self.eng
        by_param = {id(p): n for n, p in self.model.named_parameters()}
        by_buf = {id(b): n for n, b in self.model.named_buffers()}
        out = {}
        for p, off, n, _, _ in eng.param_slots:
            out[by_param[id(p)]] = self.ema_P[off:off + n].view(p.shape)
        for mod, attr, off, n in eng.buf_slots:
            out[by_buf[id(getattr(mod, attr))]] = self.ema_S[off:off + n]
        return out

    def shrink_masks(self, threshold=1e-3, ema_only=False):
        """Alive masks per block exactly as train.py:48-63: |gamma| > thr, OR-ed with the EMA's
        (EMA only on the last epoch)."""
        ema = self.ema_named()
        masks = {}
        for name, block in self.model.get_named_block_list().items():
            m = [bn.weight.detach().abs() > threshold for bn in block.get_depthwise_bn()]
            if ema:
                m_ema = [ema["{}.{}.weight".format(name, bn_name)].abs() > threshold
                         for bn_name in block.get_named_depthwise_bn().keys()]
                m = m_ema if ema_only else [a | b for a, b in zip(m, m_ema)]
            masks[name] = m
        return masks

    def shrink(self, masks=None, threshold=1e-3, ema_only=False, prune_info=None):
        """Physically drop dead atomic blocks and carry the flat optimizer / EMA arenas across,
        bit-exactly `new = old[mask]` (compress_utils.py:31-37, rmsprop.py:134-165, optim.py:134-153).
        Returns {'seconds', 'params_before', 'params_after'}; the CUDA graph is re-captured by the
        next step.  Runs identically on every rank (masks are a function of replicated state)."""
        self._check_engine()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if masks is None:
            masks = self.shrink_masks(threshold, ema_only)
        old = self.eng
        old_slot = {id(p): (off, n, tuple(p.shape)) for p, off, n, _, _ in old.param_slots}
        old_buf = {id(getattr(m, a)): (off, n) for m, a, off, n in old.buf_slots}
        old_state = (self.sq, self.buf, self.ema_P, self.ema_S)
        n_before = sum(n for _, _, n, _, _ in old.param_slots)
        rec = _ShrinkRecorder()
        for name, block in self.model.get_named_block_list().items():
            block.compress_by_mask(masks[name], ema=rec, optimizer=rec, prune_info=prune_info, prefix=name,
                                   verbose=False)
        new = E.get_engine(self.model)
        self._bind(new)
        sq, buf = torch.zeros_like(new.P), torch.zeros_like(new.P)
        ema_P = new.P.clone() if old_state[2] is not None else None
        ema_S = new.S.clone() if old_state[3] is not None else None

        def carry(dst_arena, src_arena, off_new, n_new, shape_new, info, off_old, n_old, shape_old):
            dst = dst_arena[off_new:off_new + n_new].view(shape_new)
            src = src_arena[off_old:off_old + n_old].view(shape_old)
            if info is None:
                dst.copy_(src)
            else:
                info["mask_hook"](dst, src, info["mask"])

        with torch.no_grad():
            for p, off, n, _, _ in new.param_slots:
                info = rec.masked.get(id(p))
                src_id = id(info["var_old"]) if info is not None else id(p)
                if src_id not in old_slot:
                    continue
                o_off, o_n, o_shape = old_slot[src_id]
                for dst_a, src_a in ((sq, old_state[0]), (buf, old_state[1]), (ema_P, old_state[2])):
                    if dst_a is not None:
                        carry(dst_a, src_a, off, n, tuple(p.shape), info, o_off, o_n, o_shape)
            if ema_S is not None:
                for mod, attr, off, n in new.buf_slots:
                    t = getattr(mod, attr)
                    info = rec.masked.get(id(t))
                    src_id = id(info["var_old"]) if info is not None else id(t)
                    if src_id in old_buf:
                        o_off, o_n = old_buf[src_id]
                        carry(ema_S, old_state[3], off, n, (n,), info, o_off, o_n, (o_n,))
        self.sq, self.buf, self.ema_P, self.ema_S = sq, buf, ema_P, ema_S
        if prune_info is not None:
            self.set_prune_penalties(list(zip(prune_info.weight, prune_info.penalty)))
        elif self._penalties:
            alive = dict(self.model.named_parameters())
            self.set_prune_penalties([(n, p) for n, p in self._penalties if n in alive])
        torch.cuda.synchronize()
        return {"seconds": time.perf_counter() - t0, "params_before": n_before,
                "params_after": sum(n for _, _, n, _, _ in new.param_slots)}

    # ---- checkpointing of the fast path's optimizer / EMA state --------------------------------------
    def state_dict(self):
        """Per-tensor state keyed like the model's state_dict (compacted shapes after a shrink)."""
        eng = self.eng
        names = {id(p): n for n, p in self.model.named_parameters()}
        out = {"num_updates": self.num_updates, "seed": int(self.seed_dev.item()), "square_avg": {},
               "momentum_buffer": {}, "ema": {k: v.clone() for k, v in self.ema_named().items()}}
        for p, off, n, _, _ in eng.param_slots:
            out["square_avg"][names[id(p)]] = self.sq[off:off + n].view(p.shape).clone()
            out["momentum_buffer"][names[id(p)]] = self.buf[off:off + n].view(p.shape).clone()
        return out

    def load_state_dict(self, state):
        self._check_engine()
        eng = self.eng
        names = {id(p): n for n, p in self.model.named_parameters()}
        bufs = {id(b): n for n, b in self.model.named_buffers()}
        with torch.no_grad():
            for p, off, n, _, _ in eng.param_slots:
                name = names[id(p)]
                self.sq[off:off + n].copy_(state["square_avg"][name].reshape(-1))
                self.buf[off:off + n].copy_(state["momentum_buffer"][name].reshape(-1))
                if self.ema_P is not None and name in state["ema"]:
                    self.ema_P[off:off + n].copy_(state["ema"][name].reshape(-1))
            if self.ema_S is not None:
                for mod, attr, off, n in eng.buf_slots:
                    name = bufs[id(getattr(mod, attr))]
                    if name in state["ema"]:
                        self.ema_S[off:off + n].copy_(state["ema"][name].reshape(-1))
            self.seed_dev.fill_(int(state.get("seed", 0)))
        self.num_updates = int(state["num_updates"])
