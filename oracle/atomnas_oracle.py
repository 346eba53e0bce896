"""CPU oracle for the AtomNAS supernet training step -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  The product path (atomnas_b200/) never does: it fails loudly without the
CUDA library.

It restates, as plain fp32 functional PyTorch on the CPU, the algorithm of the reference's hot
path.  The arithmetic of the reference lives in PyTorch itself (nn.Conv2d / nn.BatchNorm2d /
autograd -- `requirements.txt:1` pins only torch>=1.0), so this restatement calls the same
F.conv2d / F.batch_norm primitives and follows the reference's module code for the structure:

  network       models/mobilenet_supernet.py:58-173, models/searched_network.py:14-144
  block         models/mobilenet_base.py:277-382 (InvertedResidualChannels), :120-142 (ConvBNReLU)
  fused/SE blk  models/mobilenet_base.py:145-268 (InvertedResidualChannelsFused), :93-117 (SE) -- the
                next scope row (8f-1); restated and pinned ahead of its kernels
  activations   models/mobilenet_base.py:407-415
  init          models/mobilenet_base.py:440-459 (init_weights_mnas)
  loss          utils/optim.py:180-207 (CrossEntropyLabelSmooth), :210-249 (cal_l2_loss 'mnas'),
                utils/prune.py:161-167 (cal_bn_l1_loss), :89-158 (penalties), :215-245 (rho)
  optimizer     utils/rmsprop.py:70-132, EMA utils/optim.py:54-65
  masks/shrink  utils/prune.py:190-195, train.py:27-81, models/compress_utils.py:31-37,180-301
  MAC counts    utils/model_profiling.py:82-104

PINNING: the reference holds no golden activation/gradient vectors.  The
oracle is pinned against outputs of the reference itself, generated in the build container by
tests/golden/make_golden.py (imports /root/reference) and committed as tests/golden/*.pt;
tests/test_oracle_golden.py replays them.  The known-answer values of the reference's own
unit tests (tests/utils/prune_test.py, optim_test.py) are replayed by tests/test_oracle_golden.py::test_known_answers.
"""
import collections
import math
import numbers

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------
# architecture description
# ------------------------------------------------------------------------------------------------
def make_divisible(v, divisor, min_value=None):
    """models/mobilenet_base.py:17-31."""
    if min_value is None:
        min_value = divisor
    new_v = max(min_value, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


BlockSpec = collections.namedtuple(
    "BlockSpec", "name inp oup stride channels kernel_sizes expand")


class NetSpec(object):
    """Static description of a MobileNetV2-style AtomNAS network."""

    def __init__(self, **kw):
        self.num_classes = kw.get("num_classes", 1000)
        self.input_size = kw.get("input_size", 224)
        self.width_mult = kw.get("width_mult", 1.0)
        self.round_nearest = kw.get("round_nearest", 8)
        self.dropout_ratio = kw.get("dropout_ratio", 0.2)
        self.bn_momentum = kw.get("batch_norm_momentum", 0.1)
        self.bn_eps = kw.get("batch_norm_epsilon", 1e-5)
        self.active_fn = kw.get("active_fn", "nn.ReLU6")
        setting = kw["inverted_residual_setting"]
        if self.input_size % 32 != 0:
            raise ValueError("Input size must divide 32")
        c = make_divisible(kw.get("input_channel", 32) * self.width_mult, self.round_nearest)
        self.stem_channels = c
        self.last_channel = make_divisible(
            kw.get("last_channel", 1280) * max(1.0, self.width_mult), self.round_nearest)
        self.blocks = []
        idx = 1
        for row in setting:
            if len(row) == 5:   # supernet: [t, c, n, s, ks]  (mobilenet_supernet.py:135-146)
                t, ch, n, s, ks = row
                oup = make_divisible(ch * self.width_mult, self.round_nearest)
                for i in range(n):
                    if isinstance(t, list):
                        ratios, expand = t, True
                    elif isinstance(t, numbers.Number):
                        ratios, expand = [t for _ in ks], t != 1
                    else:
                        raise ValueError("Unknown expand_ratio type")
                    hidden = [int(round(c * e)) for e in ratios]
                    self.blocks.append(BlockSpec("features.%d" % idx, c, oup, s if i == 0 else 1,
                                                 hidden, list(ks), expand))
                    c = oup
                    idx += 1
            elif len(row) == 6:  # searched: [c, n, s, ks, hiddens, expand] (searched_network.py)
                ch, n, s, ks, hiddens, expand = row
                oup = make_divisible(ch * self.width_mult, self.round_nearest)
                for i in range(n):
                    self.blocks.append(BlockSpec("features.%d" % idx, c, oup, s if i == 0 else 1,
                                                 list(hiddens), list(ks), expand))
                    c = oup
                    idx += 1
            else:
                raise ValueError("bad inverted_residual_setting row")
        self.head_index = idx
        self.head_in = c

    def idx_depth(self, b):
        return 1 if b.expand else 0


def act(x, name):
    """get_active_fn, models/mobilenet_base.py:407-415."""
    if name == "nn.ReLU":
        return torch.relu(x)
    if name == "nn.ReLU6":
        return torch.clamp(x, 0.0, 6.0)
    if name == "nn.Swish":
        return x * torch.sigmoid(x)
    raise ValueError(name)


# ------------------------------------------------------------------------------------------------
# parameters
# ------------------------------------------------------------------------------------------------
def param_shapes(spec):
    """OrderedDict name -> shape, in the reference's named_parameters()/buffers naming."""
    P, B = collections.OrderedDict(), collections.OrderedDict()

    def conv(n, co, ci, k):
        P[n + ".weight"] = (co, ci, k, k)

    def bn(n, c):
        P[n + ".weight"] = (c,)
        P[n + ".bias"] = (c,)
        B[n + ".running_mean"] = (c,)
        B[n + ".running_var"] = (c,)
        B[n + ".num_batches_tracked"] = ()

    conv("features.0.0", spec.stem_channels, 3, 3)
    bn("features.0.1", spec.stem_channels)
    for b in spec.blocks:
        for j, (h, k) in enumerate(zip(b.channels, b.kernel_sizes)):
            base = "%s.ops.%d" % (b.name, j)
            d = 0
            if b.expand:
                conv(base + ".0.0", h, b.inp, 1)
                bn(base + ".0.1", h)
                d = 1
            P["%s.%d.0.weight" % (base, d)] = (h, 1, k, k)
            bn("%s.%d.1" % (base, d), h)
            conv("%s.%d" % (base, d + 1), b.oup, h, 1)
        bn(b.name + ".pw_bn", b.oup)
    hn = "features.%d" % spec.head_index
    conv(hn + ".0", spec.last_channel, spec.head_in, 1)
    bn(hn + ".1", spec.last_channel)
    P["classifier.1.weight"] = (spec.num_classes, spec.last_channel)
    P["classifier.1.bias"] = (spec.num_classes,)
    return P, B


def init_state(spec, seed=0):
    """init_weights_mnas (models/mobilenet_base.py:440-459) with a private generator."""
    g = torch.Generator().manual_seed(seed)
    P, B = param_shapes(spec)
    state = collections.OrderedDict()
    for n, s in P.items():
        if len(s) == 4:
            fan_out = s[2] * s[3] if s[1] == 1 and s[0] > 1 and s[2] > 1 else s[0] * s[2] * s[3]
            if s[1] == 1 and s[2] == 1:   # 1x1 conv with one input channel is not depthwise here
                fan_out = s[0] * s[2] * s[3]
            state[n] = torch.randn(s, generator=g) * (math.sqrt(2.0) / math.sqrt(fan_out))
        elif len(s) == 2:
            r = 1.0 / math.sqrt(s[0])
            state[n] = (torch.rand(s, generator=g) * 2 - 1) * r
        elif n.startswith("classifier"):
            state[n] = torch.zeros(s)
        elif n.endswith(".weight"):
            state[n] = torch.ones(s)
        else:
            state[n] = torch.zeros(s)
    for n, s in B.items():
        if n.endswith("running_var"):
            state[n] = torch.ones(s)
        elif n.endswith("num_batches_tracked"):
            state[n] = torch.zeros((), dtype=torch.long)
        else:
            state[n] = torch.zeros(s)
    return state


def param_names(spec):
    return list(param_shapes(spec)[0].keys())


# ------------------------------------------------------------------------------------------------
# forward
# ------------------------------------------------------------------------------------------------
def _bn(spec, st, name, x, training):
    """nn.BatchNorm2d semantics: biased var to normalise, unbiased into
    running_var, momentum update, momentum=None -> cumulative average."""
    rm, rv = st[name + ".running_mean"], st[name + ".running_var"]
    mom = spec.bn_momentum
    if training:
        nbt = st[name + ".num_batches_tracked"]
        nbt += 1
        if mom is None:
            mom = 1.0 / float(nbt)
    return F.batch_norm(x, rm, rv, st[name + ".weight"], st[name + ".bias"], training,
                        0.0 if mom is None else mom, spec.bn_eps)


def _conv_bn_act(spec, st, name, x, stride, groups, training, k):
    x = F.conv2d(x, st[name + ".0.weight"], None, stride, (k - 1) // 2, 1, groups)
    return act(_bn(spec, st, name + ".1", x, training), spec.active_fn)


def fused_block_forward(spec, st, name, x, inp, oup, stride, channels, kernel_sizes, expand,
                        se_ratio=None, training=True):
    """InvertedResidualChannelsFused.forward (models/mobilenet_base.py:258-268) with the optional
    SqueezeAndExcitation (:93-112) -- the block of the searched AtomNAS-C+ nets (scope table 8f-1).

    One expand conv over the concatenated hidden width (:186-193), each depthwise branch on its
    narrow slice (:196-216), concat, SE gate = sigmoid(expand(act(reduce(mean_HW(t))))) with biased
    1x1 convs of widths hidden_total -> round(inp * se_ratio) -> hidden_total (:225-230, :109-112),
    linear projection + BN (:217-219), residual iff stride == 1 and inp == oup (:171).
    State names are the reference's: {name}.expand_conv.{0,1}, .depth_ops.{i}.{idx}.{0,1} (idx = 1
    behind the Narrow when expanding), .se_op.se_reduce / .se_expand, .project_conv.{0,1}."""
    t = x
    if expand:
        t = _conv_bn_act(spec, st, name + ".expand_conv", t, 1, 1, training, 1)
    outs, start = [], 0
    for i, (h, k) in enumerate(zip(channels, kernel_sizes)):
        if expand:
            u = t.narrow(1, start, h)
            start += h
            base = "%s.depth_ops.%d.1" % (name, i)
        else:
            u = t
            base = "%s.depth_ops.%d.0" % (name, i)
        outs.append(_conv_bn_act(spec, st, base, u, stride, h, training, k))
    t = outs[0] if len(outs) == 1 else torch.cat(outs, 1)
    if se_ratio is not None:
        g = t.mean((2, 3), keepdim=True)
        g = F.conv2d(g, st[name + ".se_op.se_reduce.weight"], st[name + ".se_op.se_reduce.bias"])
        g = F.conv2d(act(g, spec.active_fn), st[name + ".se_op.se_expand.weight"],
                     st[name + ".se_op.se_expand.bias"])
        t = torch.sigmoid(g) * t
    t = F.conv2d(t, st[name + ".project_conv.0.weight"])
    t = _bn(spec, st, name + ".project_conv.1", t, training)
    if stride == 1 and inp == oup:
        return x + t
    return t


def block_forward(spec, st, b, x, training=True):
    """InvertedResidualChannels.forward, models/mobilenet_base.py:371-382."""
    if len(b.channels) == 0:
        return x
    tmp = 0
    for j, (h, k) in enumerate(zip(b.channels, b.kernel_sizes)):
        base = "%s.ops.%d" % (b.name, j)
        y, d = x, 0
        if b.expand:
            y = _conv_bn_act(spec, st, base + ".0", y, 1, 1, training, 1)
            d = 1
        y = _conv_bn_act(spec, st, "%s.%d" % (base, d), y, b.stride, h, training, k)
        y = F.conv2d(y, st["%s.%d.weight" % (base, d + 1)])
        tmp = tmp + y
    tmp = _bn(spec, st, b.name + ".pw_bn", tmp, training)
    if b.stride == 1 and b.inp == b.oup:
        return x + tmp
    return tmp


def forward(spec, st, x, training=True, dropout_mask=None):
    """MobileNetV2.forward, models/mobilenet_supernet.py:169-173.

    dropout_mask: optional [B, last_channel] 0/1 keep mask (training); None => no dropout
    (the reference draws from the torch generator, which no other implementation can replay).
    """
    x = _conv_bn_act(spec, st, "features.0", x, 2, 1, training, 3)
    for b in spec.blocks:
        x = block_forward(spec, st, b, x, training)
    x = _conv_bn_act(spec, st, "features.%d" % spec.head_index, x, 1, 1, training, 1)
    x = F.avg_pool2d(x, spec.input_size // 32)
    x = x.squeeze(3).squeeze(2)
    if training and dropout_mask is not None:
        x = x * dropout_mask / (1.0 - spec.dropout_ratio)
    return F.linear(x, st["classifier.1.weight"], st["classifier.1.bias"])


# ------------------------------------------------------------------------------------------------
# forward with the product's bf16 storage points (still CPU fp32 arithmetic)
# ------------------------------------------------------------------------------------------------
class _Round(torch.autograd.Function):
    """Round to bf16 in forward and/or backward (straight-through)."""

    @staticmethod
    def forward(ctx, x, fwd, bwd):
        ctx.bwd = bwd
        return x.to(torch.bfloat16).to(x.dtype) if fwd else x.clone()

    @staticmethod
    def backward(ctx, g):
        return (g.to(torch.bfloat16).to(g.dtype) if ctx.bwd else g), None, None


def _q2(x): return _Round.apply(x, True, True)       # tensor stored in bf16, so is its gradient
def _qf(x): return _Round.apply(x, True, False)      # operand rounded on the fly (GEMM A / dw input)
def _qb(x): return _Round.apply(x, False, True)      # only the gradient is stored in bf16
def _qw(w): return w + (w.to(torch.bfloat16).to(w.dtype) - w).detach()   # bf16 weight copy


def dw_on_tensor_cores(k, stride, batch, width):
    """Which depthwise layers the engine's FORWARD runs as banded tensor-core products (csrc/dw_mma.cu
    use_conv: stride 1, k = 7, at least 8 images, maps at least 14 wide): there BOTH operands are bf16 --
    the activated input is rounded after the on-load BN + activation and the taps are the bf16
    rounding of the fp32 master (exactly what an autocast convolution multiplies).  The other layers
    keep fp32 taps and an unrounded operand."""
    return stride == 1 and k == 7 and batch >= 8 and width >= 14


def block_forward_bf16_points(spec, st, b, y, training=True):
    """One InvertedResidualChannels block with the engine's bf16 storage points; `y` is the
    (bf16-valued) block input."""
    a = spec.active_fn
    if len(b.channels) == 0:
        return y
    tmp = 0
    for j, (h, k) in enumerate(zip(b.channels, b.kernel_sizes)):
        base = "%s.ops.%d" % (b.name, j)
        t, d = y, 0
        if b.expand:
            t = _q2(F.conv2d(y, _qw(st[base + ".0.0.weight"])))
            # ReLU / ReLU6 / identity: the depthwise kernels apply BN + activation to the raw (bf16)
            # expand output ON LOAD, in fp32 -- the activated operand is not rounded.  Swish: the engine
            # materialises act(bn(.)) as a bf16 tensor first (exp on the load path is too dear).
            t = act(_qb(_bn(spec, st, base + ".0.1", t, training)), a)
            if a == "nn.Swish" or dw_on_tensor_cores(k, b.stride, y.shape[0], y.shape[3]):
                t = _qf(t)
            d = 1
        wd = st["%s.%d.0.weight" % (base, d)]
        if dw_on_tensor_cores(k, b.stride, y.shape[0], y.shape[3]):
            wd = _qw(wd)
        t = _q2(F.conv2d(t, wd, None, b.stride, (k - 1) // 2, 1, h))
        t = _qf(act(_qb(_bn(spec, st, "%s.%d.1" % (base, d), t, training)), a))
        tmp = tmp + F.conv2d(t, _qw(st["%s.%d.weight" % (base, d + 1)]))
    tmp = _bn(spec, st, b.name + ".pw_bn", _q2(tmp), training)
    return _q2(y + tmp) if (b.stride == 1 and b.inp == b.oup) else _q2(tmp)


def forward_bf16_points(spec, st, x, training=True):
    """Same network as `forward`, in fp32 arithmetic, but rounding to bf16 exactly where the B200
    engine stores bf16 (DESIGN.md 'data layout'): raw conv outputs and their gradients, the
    activated operands fed to the next conv, block outputs, gradients w.r.t. BN outputs, and the
    bf16 copies of the 1x1 / classifier weights.  Differences to the engine that remain are fp32
    accumulation order and fast-math exp -- so this is the tight parity target, while `forward`
    (pure fp32 == the reference) measures the bf16 error itself."""
    a = spec.active_fn

    def bn_act(name, t, fwd_round):
        t = act(_qb(_bn(spec, st, name, t, training)), a)
        return _qf(t) if fwd_round else t

    s = _q2(F.conv2d(x, st["features.0.0.weight"], None, 2, 1))
    y = _q2(bn_act("features.0.1", s, False))
    for b in spec.blocks:
        y = block_forward_bf16_points(spec, st, b, y, training)
    hn = "features.%d" % spec.head_index
    t = _q2(F.conv2d(y, _qw(st[hn + ".0.weight"])))
    t = bn_act(hn + ".1", t, False)
    t = _q2(F.avg_pool2d(t, spec.input_size // 32).squeeze(3).squeeze(2))
    return _qb(F.linear(t, _qw(st["classifier.1.weight"]), st["classifier.1.bias"]))


# ------------------------------------------------------------------------------------------------
# losses
# ------------------------------------------------------------------------------------------------
def cross_entropy_label_smooth(logits, target, num_classes, eps):
    """utils/optim.py:199-207 with reduction='none'."""
    logp = F.log_softmax(logits, dim=1)
    t = torch.zeros_like(logp).scatter_(1, target.unsqueeze(1), 1)
    t = (1 - eps) * t + eps / num_classes
    return torch.sum(-t * logp, 1)


def l2_loss_mnas(named_params, weight_decay):
    """utils/optim.py:226-249, method 'mnas'."""
    loss = 0.0
    for name, p in named_params:
        if p.dim() in (4, 2) or (p.dim() == 1 and "classifier" in name):
            loss = loss + weight_decay * (p ** 2).sum()
    return loss * 0.5


def conv_macs(cin, cout, k, ho, wo, groups=1, batch=1):
    """utils/model_profiling.py:82-85."""
    return (cin * cout * k * k * ho * wo // groups) * batch


def block_op_macs(spec, b, hin):
    """n_macs of every branch Sequential of a block (model_profiling.py:82-85, :121-127)."""
    ho = hin // b.stride
    res = []
    for h, k in zip(b.channels, b.kernel_sizes):
        m = 0
        if b.expand:
            m += conv_macs(b.inp, h, 1, hin, hin)
        m += conv_macs(h, h, k, ho, ho, groups=h)
        m += conv_macs(h, b.oup, 1, ho, ho)
        res.append(m)
    return res


def network_macs(spec):
    s = spec.input_size
    total = conv_macs(3, spec.stem_channels, 3, s // 2, s // 2)
    h = s // 2
    for b in spec.blocks:
        total += sum(block_op_macs(spec, b, h))
        h //= b.stride
    total += conv_macs(spec.head_in, spec.last_channel, 1, h, h)
    total += spec.last_channel * h * h          # AvgPool2d, model_profiling.py:101-104
    total += spec.last_channel * spec.num_classes
    return total


def prune_targets(spec, bn_prune_filter="expansion_only_skip_expand1"):
    """get_bn_to_prune, utils/prune.py:89-158: names, penalties, per-channel MACs."""
    names, pen = [], []
    h = spec.input_size // 2
    for b in spec.blocks:
        macs = block_op_macs(spec, b, h)
        h //= b.stride
        if bn_prune_filter == "expansion_only_skip_expand1" and not b.expand:
            continue
        for j, (c, m) in enumerate(zip(b.channels, macs)):
            names.append("%s.ops.%d.%d.1.weight" % (b.name, j, spec.idx_depth(b)))
            pen.append((c, m / c))
    per_channel = [v for _, v in pen]
    numel = sum(n for n, _ in pen)
    norm = sum(n * v for n, v in pen) / (numel + 1e-5)
    return names, [v / norm for _, v in pen], per_channel


def bn_l1_loss(weights, penalties, rho):
    """utils/prune.py:161-167."""
    loss = 0.0
    for w, p in zip(weights, penalties):
        loss = loss + rho * p * w.abs().sum()
    return loss


def rho_at(step, steps_per_epoch, rho, epoch_free, epoch_warmup, stepwise=True):
    """get_rho_scheduler 'linear', utils/prune.py:215-245."""
    free, warm = epoch_free * steps_per_epoch, epoch_warmup * steps_per_epoch
    i = step if stepwise else (step // steps_per_epoch) * steps_per_epoch
    if i < free:
        return 0.0
    if i >= warm:
        return rho
    return (i - free) / (warm - free) * rho


def masks_by_threshold(weights, threshold):
    """utils/prune.py:190-195 (strict >)."""
    return [w.detach().abs() > threshold for w in weights]


# ------------------------------------------------------------------------------------------------
# optimizer / EMA
# ------------------------------------------------------------------------------------------------
def rmsprop_step(p, g, sq, buf, lr, alpha=0.9, momentum=0.9, eps=1e-3, eps_inside_sqrt=True):
    """One tensor of utils/rmsprop.py:100-130 (weight_decay=0, centered=False); in place."""
    sq.mul_(alpha).addcmul_(g, g, value=1 - alpha)
    avg = sq.add(eps).sqrt_() if eps_inside_sqrt else sq.sqrt().add_(eps)
    if momentum > 0:
        buf.mul_(momentum).addcdiv_(g, avg)
        p.add_(buf, alpha=-lr)
    else:
        p.addcdiv_(g, avg, value=-lr)


def ema_decay(base, num_updates):
    """utils/optim.py:56-60."""
    if num_updates is None:
        return base
    return min(base, (1.0 + num_updates) / (10.0 + num_updates))


def ema_update(shadow, x, decay):
    """utils/optim.py:64-65."""
    shadow.mul_(decay).add_(x.detach(), alpha=1.0 - decay)


# ------------------------------------------------------------------------------------------------
# one full training step (the thing bench.py's cpu_baseline times)
# ------------------------------------------------------------------------------------------------
class TrainState(object):
    def __init__(self, spec, state):
        self.spec = spec
        self.state = state
        self.names = param_names(spec)
        for n in self.names:
            state[n].requires_grad_(True)
        self.sq = {n: torch.zeros_like(state[n]) for n in self.names}
        self.buf = {n: torch.zeros_like(state[n]) for n in self.names}
        self.ema = {n: state[n].detach().clone() for n in state
                    if not n.endswith("num_batches_tracked")}
        self.step = 0
        self.prune_names, self.penalties, _ = prune_targets(spec)


def train_step(ts, x, target, lr, weight_decay=1e-5, rho=0.0, label_smoothing=0.1,
               ema_base=None, dropout_mask=None, world_grads=None):
    """zero_grad -> forward -> CE+L2+L1 -> backward -> RMSprop -> EMA (train.py:174-236)."""
    spec, st = ts.spec, ts.state
    for n in ts.names:
        st[n].grad = None
    logits = forward(spec, st, x, True, dropout_mask)
    ce = cross_entropy_label_smooth(logits, target, spec.num_classes, label_smoothing)
    loss = ce.mean()
    loss_l2 = l2_loss_mnas([(n, st[n]) for n in ts.names], weight_decay)
    loss_l1 = bn_l1_loss([st[n] for n in ts.prune_names], ts.penalties, rho)
    total = loss + loss_l2 + loss_l1
    total.backward()
    with torch.no_grad():
        for n in ts.names:
            rmsprop_step(st[n], st[n].grad, ts.sq[n], ts.buf[n], lr)
        ts.step += 1
        if ema_base is not None:
            d = ema_decay(ema_base, ts.step)
            for n, sh in ts.ema.items():
                ema_update(sh, st[n], d)
    return loss.detach(), logits.detach(), ce.detach()


# ------------------------------------------------------------------------------------------------
# shrink (index semantics only)
# ------------------------------------------------------------------------------------------------
def compress_block_state(spec_block, st, masks, expand):
    """new = old[mask] along dim 0 (expand conv/BN, dw conv/BN) or dim 1 (project conv), dead
    branches dropped and the survivors renumbered (models/compress_utils.py:220-301).
    Returns (new BlockSpec, {new_name: tensor}) for the block's variables."""
    b = spec_block
    keep = [int(m.sum().item()) > 0 for m in masks]
    new_channels = [int(m.sum().item()) for m, k in zip(masks, keep) if k]
    new_ks = [ks for ks, k in zip(b.kernel_sizes, keep) if k]
    out = collections.OrderedDict()
    nj = 0
    d = 1 if expand else 0
    for j, (m, k) in enumerate(zip(masks, keep)):
        if not k:
            continue
        old, new = "%s.ops.%d" % (b.name, j), "%s.ops.%d" % (b.name, nj)
        for sub in ([".0"] if expand else []) + [".%d" % d]:
            out[new + sub + ".0.weight"] = st[old + sub + ".0.weight"][m]
            for a in ("weight", "bias", "running_mean", "running_var"):
                out[new + sub + ".1." + a] = st[old + sub + ".1." + a][m]
            out[new + sub + ".1.num_batches_tracked"] = st[old + sub + ".1.num_batches_tracked"]
        out["%s.%d.weight" % (new, d + 1)] = st["%s.%d.weight" % (old, d + 1)][:, m]
        nj += 1
    for a in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked"):
        out[b.name + ".pw_bn." + a] = st[b.name + ".pw_bn." + a]
    nb = BlockSpec(b.name, b.inp, b.oup, b.stride, new_channels, new_ks, b.expand)
    return nb, out
