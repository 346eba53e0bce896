"""Fused training step: the public fast path of the package.

One call = train.py:174-236 of the reference for one batch: forward, label-smoothed CE, backward,
(bucketed NCCL all-reduce of the gradient arena overlapped with backward), RMSprop with the
analytic L2 / L1 terms, EMA of parameters and BN running statistics, refresh of the bf16 weight
copies -- with no per-parameter Python loop and no host synchronisation.

    trainer = Trainer(model, lr=0.128, weight_decay=1e-5, ema_decay=0.9999 ** 0.5)
    loss = trainer.step_host(images_pinned, labels_pinned)     # H2D + step; device scalar

The drop-in route (reference train.py + utils.rmsprop.RMSprop + ExponentialMovingAverage) uses
the same engine through atomnas_b200.rmsprop / atomnas_b200.optim instead.
"""
import torch
import torch.distributed as dist

from . import engine as E
from . import kernels as K


class Trainer(object):

    def __init__(self, model, lr, alpha=0.9, momentum=0.9, eps=1e-3, eps_inside_sqrt=True,
                 weight_decay=1e-5, label_smoothing=0.1, ema_decay=None, prune_penalties=None,
                 bucket_bytes=8 << 20, use_cuda_graph=False):
        self.model = model
        self.eng = E.get_engine(model)
        if not self.eng.is_net:
            raise ValueError("Trainer needs a whole network")
        eng = self.eng
        self.lr, self.alpha, self.momentum, self.eps = lr, alpha, momentum, eps
        self.eps_inside_sqrt = eps_inside_sqrt
        self.weight_decay, self.label_smoothing = weight_decay, label_smoothing
        self.ema_decay = ema_decay
        self.sq = torch.zeros_like(eng.P)
        self.buf = torch.zeros_like(eng.P)
        self.ema_P = eng.P.clone() if ema_decay else None
        self.ema_S = eng.S.clone() if ema_decay else None
        if prune_penalties:
            eng.set_l1_penalties(prune_penalties, dict(model.named_parameters()))
        self.num_updates = 0
        self.hp_dev = torch.zeros(4, device=eng.dev)                       # lr, rho, ema decay
        self.seed_dev = torch.zeros(1, device=eng.dev, dtype=torch.int64)  # dropout stream position
        self.launches_per_step = None
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.bucket_elems = max(1, bucket_bytes // 4)
        self._handles = []
        self._pending_hi = None
        self._x_dev = None
        self._y_dev = None
        self.use_cuda_graph = use_cuda_graph and self.world == 1
        self._graph = None
        self._static = None

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

    def _step_impl(self, x, target):
        eng = self.eng
        self.seed_dev.add_(1)                      # dropout stream advances on the device
        eng.forward(x, True, True, seed=0x5EED, seed_dev=self.seed_dev)
        loss = eng.loss_and_dlogits(target, self.label_smoothing, 1.0)
        eng.backward(None, grad_ready=self._grad_ready)
        self._finish_allreduce()
        K.rmsprop_ema_step(eng.P, eng.G, self.sq, self.buf, self.ema_P, eng.wd_flag, eng.l1_pen, eng.n_params,
                           0.0, self.alpha, self.momentum, self.eps, self.eps_inside_sqrt, 1.0 / self.world,
                           self.weight_decay, 0.0, 0.0, hparams_dev=self.hp_dev)
        if self.ema_decay:
            K.ema_update(self.ema_S, eng.S, eng.S.numel(), 0.0, hparams_dev=self.hp_dev)
        eng.refresh_weights(force=True)
        return loss

    def step(self, x, target, rho=0.0, lr=None):
        """x: fp32 NCHW on the device, target: int64 labels.  Returns per-sample losses (device)."""
        lr = self.lr if lr is None else lr
        self._push_hparams(rho, lr)
        if not self.use_cuda_graph:
            return self._step_impl(x, target)
        return self._step_graphed(x, target)

    def _step_graphed(self, x, target):
        """Whole step replayed from a CUDA graph (the ~900 launches of a step are launch-bound
        from Python); lr / rho / EMA decay / dropout seed are read from device memory."""
        key = tuple(x.shape)
        if self._graph is None or self._static[2] != key:
            sx, sy = x.clone(), target.clone()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                first = self._step_impl(sx, sy).clone()    # this step runs eagerly (also the warm-up)
            torch.cuda.current_stream().wait_stream(side)
            self._graph = torch.cuda.CUDAGraph()
            from ._lib import lib
            c0 = lib().an_launch_count()
            with torch.cuda.graph(self._graph):            # capture records, it does not execute
                out = self._step_impl(sx, sy)
            self.launches_per_step = lib().an_launch_count() - c0     # kernel nodes replayed per step
            self._static = (sx, sy, key, out)
            return first
        sx, sy, _, out = self._static
        sx.copy_(x, non_blocking=True)
        sy.copy_(target, non_blocking=True)
        self._graph.replay()
        return out

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
