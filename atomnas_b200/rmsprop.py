"""TF-style RMSprop on the fused sm_100a optimizer kernel (mirror of utils/rmsprop.py:70-182).

Same constructor, `param_groups`, `state` keys (`step`, `square_avg`, `momentum_buffer`),
`compress_mask` / `compress_drop` protocol as the reference class, so utils.optim.get_optimizer,
LambdaLR, checkpointing and the shrink code work unchanged.  `step()` is ONE launch of
an_rmsprop_ema_step per flat arena instead of ~6 ATen kernels per parameter tensor:

  * parameters owned by an atomnas_b200 engine use the engine's parameter / gradient arenas;
  * any other CUDA parameters are packed into a private arena on first use.

There is no CPU path: stepping CPU parameters raises.
"""
import logging

import torch
from torch.optim.optimizer import Optimizer

from . import engine as E
from . import kernels as K


class _Flat(object):
    """A flat (P, G, square_avg, momentum) quadruple covering a fixed list of parameters."""

    def __init__(self, params, eng, old_state):
        self.eng = eng
        self.sig = tuple((id(p), p.data_ptr()) for p in params)
        self.params = list(params)
        if eng is not None:
            self.P, self.G, self.n = eng.P, eng.G, eng.n_params
            slots = {id(p): (off, n) for p, off, n, _, _ in eng.param_slots}
            self.slots = [slots[id(p)] for p in params]
            self.wd_flag, self.l1_pen = eng.wd_flag, eng.l1_pen
            if set(id(p) for p in params) != set(slots.keys()):
                # this param_group covers only part of the engine's arena (several groups, frozen
                # parameters, params without gradients): every element outside it is frozen (-1) for
                # this launch, as the reference's per-parameter loop would skip it (rmsprop.py:81-85)
                self.wd_flag = torch.full_like(eng.wd_flag, -1.0)
                for o, k in self.slots:
                    self.wd_flag[o:o + k] = eng.wd_flag[o:o + k]
        else:
            dev = params[0].device
            offs, n = [], 0
            for p in params:
                n = (n + 7) // 8 * 8
                offs.append(n)
                n += p.numel()
            self.n = (n + 7) // 8 * 8
            self.P = torch.zeros(self.n, device=dev)
            self.G = torch.zeros(self.n, device=dev)
            self.slots = [(o, p.numel()) for o, p in zip(offs, params)]
            self.wd_flag = torch.full((self.n,), -1.0, device=dev)
            self.l1_pen = torch.zeros(self.n, device=dev)
            with torch.no_grad():
                for p, (o, k) in zip(params, self.slots):
                    view = self.P[o:o + k].view(p.shape)
                    view.copy_(p.data)
                    p.data = view
                    self.wd_flag[o:o + k] = 1.0            # optimizer-level weight_decay hits every element
        self.sq = torch.zeros(self.n, device=self.P.device)
        self.buf = torch.zeros(self.n, device=self.P.device)
        with torch.no_grad():
            for p, (o, k) in zip(params, self.slots):
                st = old_state.get(p, {})
                if "square_avg" in st:
                    self.sq[o:o + k].copy_(st["square_avg"].reshape(-1))
                if "momentum_buffer" in st:
                    self.buf[o:o + k].copy_(st["momentum_buffer"].reshape(-1))

    def views(self, p_index):
        o, k = self.slots[p_index]
        shape = self.params[p_index].shape
        return self.sq[o:o + k].view(shape), self.buf[o:o + k].view(shape), self.G[o:o + k].view(shape)


class RMSprop(Optimizer):

    def __init__(self, params, lr=1e-2, alpha=0.99, eps=1e-8, eps_inside_sqrt=False, weight_decay=0,
                 momentum=0, centered=False):
        for name, v in (("learning rate", lr), ("epsilon value", eps), ("momentum value", momentum),
                        ("weight_decay value", weight_decay), ("alpha value", alpha)):
            if not 0.0 <= v:
                raise ValueError("Invalid {}: {}".format(name, v))
        if centered:
            raise NotImplementedError("centered RMSprop is not on the fused sm_100a path")
        defaults = dict(lr=lr, momentum=momentum, alpha=alpha, eps=eps, eps_inside_sqrt=eps_inside_sqrt,
                        centered=centered, weight_decay=weight_decay)
        super(RMSprop, self).__init__(params, defaults)
        self._flats = {}

    def __setstate__(self, state):
        super(RMSprop, self).__setstate__(state)
        for group in self.param_groups:
            group.setdefault("momentum", 0)
            group.setdefault("centered", False)
        self._flats = {}

    # ---- arena binding -----------------------------------------------------------------------
    def _bind(self, gi, group):
        params = [p for p in group["params"] if p.grad is not None]
        if not params:
            return []
        if any(not p.is_cuda for p in params):
            raise RuntimeError("atomnas_b200.RMSprop: CPU parameters -- the optimizer only runs as the "
                               "fused sm_100a kernel (no CPU fallback)")
        by_eng, free = {}, []
        for p in params:
            eng = E.engine_of_param(p)
            (by_eng.setdefault(eng, []) if eng is not None else free).append(p)
        flats = []
        for key, plist, eng in [(("e", id(e)), pl, e) for e, pl in by_eng.items()] + \
                               ([(("free",), free, None)] if free else []):
            sig = tuple((id(p), p.data_ptr()) for p in plist)
            flat = self._flats.get((gi,) + key)
            if flat is None or flat.sig != sig or (eng is not None and flat.P is not eng.P):
                flat = _Flat(plist, eng, self.state)
                self._flats[(gi,) + key] = flat
                for i, p in enumerate(plist):
                    sq, buf, _ = flat.views(i)
                    st = self.state[p]
                    st.setdefault("step", 0)
                    st["square_avg"] = sq
                    if group["momentum"] > 0:
                        st["momentum_buffer"] = buf
            flats.append(flat)
        return flats

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            for flat in self._bind(gi, group):
                if flat.eng is not None and group["weight_decay"] != 0:
                    raise NotImplementedError("engine-owned parameters take L2 through cal_l2_loss / Trainer "
                                              "(the reference always constructs RMSprop with weight_decay=0)")
                for i, p in enumerate(flat.params):
                    _, _, gview = flat.views(i)
                    if p.grad.data_ptr() != gview.data_ptr():
                        gview.copy_(p.grad)               # gradient produced outside the arena
                    self.state[p]["step"] += 1
                K.rmsprop_ema_step(flat.P, flat.G, flat.sq, flat.buf if group["momentum"] > 0 else None, None,
                                   flat.wd_flag, flat.l1_pen, flat.n, group["lr"], group["alpha"],
                                   group["momentum"], group["eps"], group["eps_inside_sqrt"], 1.0,
                                   group["weight_decay"], 0.0, 0.0)
                if flat.eng is not None:
                    flat.eng.refresh_weights(force=True)
        return loss

    # ---- dynamic network shrinkage (models/compress_utils.py drives these) ----------------------
    def _find(self, var):
        for group in self.param_groups:
            for i, p in enumerate(group["params"]):
                if p is var:
                    return group, i
        return None, None

    def compress_mask(self, info, verbose=False):
        """new state = old state[mask] along the masked dim; swap the parameter in its group."""
        var_old, var_new = info["var_old"], info["var_new"]
        if verbose:
            logging.info("RMSProp compress: {} -> {}".format(info["var_old_name"], info["var_new_name"]))
        group, idx = self._find(var_old)
        assert group is not None, "Var: {} not in RMSProp".format(info["var_old_name"])
        state = self.state.pop(var_old, None)
        if state:
            new_state = {"step": state["step"]}
            for key in ("square_avg", "momentum_buffer", "grad_avg"):
                if key in state:
                    new_state[key] = torch.zeros_like(var_new.data, device=var_old.device)
                    info["mask_hook"](new_state[key], state[key], info["mask"])
            self.state[var_new] = new_state
        del group["params"][idx]
        group["params"].append(var_new)
        self._flats = {}

    def compress_drop(self, info, verbose=False):
        if verbose:
            logging.info("RMSProp drop: {}".format(info["var_old_name"]))
        assert info["type"] == "variable"
        group, idx = self._find(info["var_old"])
        assert group is not None, "Var: {} not in RMSProp".format(info["var_old_name"])
        self.state.pop(info["var_old"], None)
        del group["params"][idx]
        self._flats = {}
