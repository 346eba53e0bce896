#!/usr/bin/env python
"""Benchmark of the hot path: AtomNAS supernet training step on B200.

    python bench.py --gpus N --steps K --warmup W              # this repo (sm_100a kernels)
    python bench.py --impl reference --gpus N --steps K ...    # the unmodified reference on the host cores
    python bench.py --impl reference-gpu [--ref-mode fp32|bf16]  # the unmodified reference modules on the same B200
    python bench.py --model atomnas_a                          # BASELINE config 2 (AtomNAS-A supernet)
    python bench.py --masks random|searched                    # BASELINE config 3 fixtures (ii)/(iii): shrink first
    python bench.py --accum 8                                  # (S) series: 8 micro-batches of 256 per optimizer step

One "step" = forward + label-smoothed CE + backward + gradient all-reduce + RMSprop (with the
analytic L2 / BN-gamma L1 terms) + EMA, on one synthetic ImageNet-shaped batch of 256 images per
GPU (weak scaling: global batch 256*N, 2048 at N=8 -- BASELINE.json configs[2]).  Prints ONE JSON
line (contract in the task statement / DESIGN.md "Measurement").  Under torchrun every rank runs
every collective of every phase; rank 0 prints.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_SETTING = [[6, 24, 4, 2, [3, 5, 7]], [6, 40, 4, 2, [3, 5, 7]], [6, 80, 4, 2, [3, 5, 7]], [6, 96, 4, 1, [3, 5, 7]],
            [6, 192, 4, 2, [3, 5, 7]], [6, 320, 1, 1, [3, 5, 7]]]
ATOMNAS_C_SUPERNET = dict(      # apps/slimming/shrink/atomnas_c.yml -> supernet_single_path_nas_relu.yml
    num_classes=1000, input_channel=32, last_channel=1280, width_mult=1.0, round_nearest=8,
    inverted_residual_setting=[[1, 16, 1, 1, [3]]] + _SETTING,
    dropout_ratio=0.2, batch_norm_momentum=0.01, batch_norm_epsilon=1e-3, active_fn="nn.ReLU",
    block="InvertedResidualChannels")
ATOMNAS_A_SUPERNET = dict(      # apps/slimming/shrink/atomnas_a.yml -> supernet_single_path_nas_relu_small.yml
    ATOMNAS_C_SUPERNET, input_channel=16)
MODELS = {"atomnas_c": ("AtomNAS-C supernet (apps/slimming/shrink/atomnas_c.yml)", ATOMNAS_C_SUPERNET),
          "atomnas_a": ("AtomNAS-A supernet (apps/slimming/shrink/atomnas_a.yml)", ATOMNAS_A_SUPERNET)}
# apps/searched/models/atomnas_c.yml:14-35 -- per-branch widths of the searched AtomNAS-C (the end
# state of the shrink): (kernel sizes, hidden widths) per block; mask fixture (iii) keeps the first n atoms.
SEARCHED_C = [[[3], [32]], [[3, 5, 7], [15, 23, 13]], [[3, 5, 7], [20, 3, 2]], [[3, 5, 7], [21, 8, 1]],
              [[3, 5], [21, 4]], [[3, 5, 7], [64, 65, 70]], [[3, 5, 7], [40, 23, 28]], [[3, 5, 7], [36, 21, 9]],
              [[3, 5, 7], [40, 22, 7]], [[3, 5, 7], [169, 170, 157]], [[3, 5, 7], [45, 55, 74]],
              [[3, 5, 7], [58, 21, 37]], [[3, 5, 7], [77, 18, 39]], [[3, 5, 7], [393, 263, 247]],
              [[3, 5, 7], [57, 40, 36]], [[3, 5, 7], [61, 34, 26]], [[3, 5, 7], [87, 38, 33]],
              [[3, 5, 7], [454, 468, 468]], [[3, 5, 7], [188, 137, 306]], [[3, 5, 7], [253, 187, 314]],
              [[3, 5, 7], [345, 202, 233]], [[3, 5, 7], [823, 738, 749]]]
_C_OUT = [16, 24, 24, 24, 24, 40, 40, 40, 40, 80, 80, 80, 80, 96, 96, 96, 96, 192, 192, 192, 192, 320]
_C_STRIDE = [1, 2, 1, 1, 1, 2, 1, 1, 1, 2, 1, 1, 1, 1, 1, 1, 1, 2, 1, 1, 1, 1]
ATOMNAS_C_PLUS = dict(           # apps/searched/atomnas_c/atomnas_c+.yml (BASELINE config 4): searched widths,
    num_classes=1000, input_channel=32, last_channel=1280, width_mult=1.0, round_nearest=8,   # SE 0.5, Swish
    inverted_residual_setting=[[c, 1, s, ks, hid, i > 0] for i, (c, s, (ks, hid)) in
                               enumerate(zip(_C_OUT, _C_STRIDE, SEARCHED_C))],
    dropout_ratio=0.28, se_ratio=0.5, batch_norm_momentum=0.01, batch_norm_epsilon=1e-3, active_fn="nn.Swish",
    block="InvertedResidualChannelsFused")
MODELS["atomnas_c_plus"] = ("AtomNAS-C+ (apps/searched/atomnas_c/atomnas_c+.yml: searched widths, SE 0.5, Swish, "
                            "fused block)", ATOMNAS_C_PLUS)
PER_GPU_BATCH = 256
IMAGE = 224
RHO = 1.0e-4            # post-warm-up L1 weight (atomnas_c.yml prune_params.rho): pruning term active
WEIGHT_DECAY = 1.0e-5
METRIC = "images/sec AtomNAS-C supernet fwd+bwd+RMSprop bf16"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([f.strip() for f in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm = sorted(float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit())
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names)
                   if any(len(s) > 3 + i and s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def build_model(kwargs=ATOMNAS_C_SUPERNET):
    from atomnas_b200 import mobilenet_base as mb
    from atomnas_b200 import mobilenet_supernet as ms
    torch.manual_seed(1995)
    cls = ms.MobileNetSearched if len(kwargs["inverted_residual_setting"][0]) == 6 else ms.MobileNetV2
    model = cls(**kwargs, input_size=IMAGE)
    model.apply(mb.init_weights_mnas)
    return model


def build_trainer(model, world, use_graph):
    from atomnas_b200 import prune
    from atomnas_b200.trainer import Trainer
    prune.profile_macs(model, IMAGE)
    searched = hasattr(list(model.get_named_block_list().values())[0], "depth_ops")
    info = prune.get_bn_to_prune(model, {"bn_prune_filter": None if searched else "expansion_only_skip_expand1"},
                                 verbose=False)
    gb = PER_GPU_BATCH * world
    tr = Trainer(model, lr=0.016 * gb / 256, alpha=0.9, momentum=0.9, eps=1e-3, eps_inside_sqrt=True,
                 weight_decay=WEIGHT_DECAY, label_smoothing=0.1, ema_decay=0.9999 ** (gb / 4096.0),
                 prune_penalties=list(zip(info.weight, info.penalty)), use_cuda_graph=use_graph)
    return tr, info


def mask_fixture(model, kind):
    """BASELINE config 3 mask fixtures: 'random' = keep-probability 0.5 per atom with
    manual_seed(1995); 'searched' = first n atoms of every branch alive, n from
    apps/searched/models/atomnas_c.yml:14-35 (1521 M -> 360 M MACs)."""
    g = torch.Generator().manual_seed(1995)
    masks = {}
    for i, (name, block) in enumerate(model.get_named_block_list().items()):
        ms_ = []
        for j, (k, c) in enumerate(zip(block.kernel_sizes, block.channels)):
            if kind == "random":
                m = torch.rand(c, generator=g) < 0.5
                if not block.expand:
                    m[:] = True
            else:
                ks, hid = SEARCHED_C[i]
                n = hid[ks.index(k)] if k in ks else 0
                m = torch.arange(c) < n
            ms_.append(m.cuda())
        masks[name] = ms_
    return masks


def timed_loop(fn, steps, world):
    """barrier + sync, CUDA events around exactly `steps` calls, max over ranks; returns seconds."""
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        fn()
    e.record()
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / 1e3], device="cuda")
    if world > 1:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


# ------------------------------------------------------------------------------------------------
# algorithmic work per kernel class, SURVEY 8(d) convention: every tensor a pass touches counted
# ONCE per pass (forward, data gradient, weight gradient); re-reads of saved tensors by fused
# epilogues / prologues are NOT billed.
# ------------------------------------------------------------------------------------------------
CLASSES = {"an_dw_fwd": "depthwise", "an_dw_bwd_data": "depthwise", "an_dw_bwd_weight": "depthwise",
           "an_dw_bwd": "depthwise",
           "an_pw_gemm": "pointwise", "an_pw_wgrad": "pointwise", "an_wgrad_reduce": "pointwise",
           "an_wgrad_reduce_multi": "pointwise",
           "an_affine_act": "elementwise", "an_col_reduce": "elementwise", "an_act_bwd_reduce": "elementwise",
           "an_bn_finalize": "bn_finalize", "an_bn_bwd_finalize": "bn_finalize",
           "an_bn_finalize_multi": "bn_finalize", "an_bn_bwd_finalize_multi": "bn_finalize"}


def algorithmic_work(eng, B):
    """{class: (bytes, flops)} per training step for the 22 blocks + head."""
    work = {"depthwise": [0, 0], "pointwise": [0, 0]}
    h = IMAGE // 2
    for p in eng.blocks:
        ho = h // 2 if p.stride == 2 else h
        M, Mo = B * h * h, B * ho * ho
        if not p.branches:
            h = ho
            continue
        if p.expand:
            work["pointwise"][0] += 3 * 2 * M * (p.inp + p.Kh)
            work["pointwise"][1] += 3 * 2 * M * p.Kh * p.inp
        for br in p.branches:
            work["depthwise"][0] += 3 * 2 * (M + Mo) * br.hp
            work["depthwise"][1] += 3 * 2 * Mo * br.hp * br.k * br.k
        work["pointwise"][0] += 3 * 2 * Mo * (p.Kh + p.oup)
        work["pointwise"][1] += 3 * 2 * Mo * p.Kh * p.oup
        h = ho
    hd = eng.head
    Mh = B * h * h
    work["pointwise"][0] += 3 * 2 * Mh * (hd.cin + hd.C)
    work["pointwise"][1] += 3 * 2 * Mh * hd.cin * hd.C
    return work


def profile_kernels(trainer, x, y, peaks, steps=2):
    from atomnas_b200 import _lib as L
    L.start_profile()
    for _ in range(steps):
        trainer.step(x, y, rho=RHO)
    rec = L.stop_profile()
    total_ms = sum(v[1] for v in rec.values())
    work = algorithmic_work(trainer.eng, x.shape[0])
    families, classes = {}, {}
    for (name, tag), (calls, ms) in rec.items():
        f = families.setdefault(name, {"ms": 0.0, "calls": 0})
        f["ms"] += ms / steps
        f["calls"] += calls // steps
        c = classes.setdefault(CLASSES.get(name, "other"), {"ms": 0.0, "calls": 0})
        c["ms"] += ms / steps
        c["calls"] += calls // steps
    table = [{"kernel": name, "class": CLASSES.get(name, "other"), "ms_per_step": round(f["ms"], 3),
              "launches_per_step": f["calls"], "share": round(f["ms"] * steps / total_ms, 4)}
             for name, f in sorted(families.items(), key=lambda kv: -kv[1]["ms"])]
    ctable = []
    for name, c in sorted(classes.items(), key=lambda kv: -kv[1]["ms"]):
        row = {"class": name, "ms_per_step": round(c["ms"], 3), "launches_per_step": c["calls"],
               "share": round(c["ms"] * steps / total_ms, 4)}
        if name in work:
            nbytes, flops = work[name]
            row["algorithmic_GB"] = round(nbytes / 1e9, 3)
            row["algorithmic_GBps"] = round(nbytes / (c["ms"] * 1e-3) / 1e9, 1)
            row["frac_hbm"] = round(row["algorithmic_GBps"] / peaks["hbm_gbs"], 4)
            row["algorithmic_TFLOPs"] = round(flops / (c["ms"] * 1e-3) / 1e12, 2)
            row["frac_tensor_sustained"] = round(row["algorithmic_TFLOPs"] / peaks["bf16_tflops_sustained"], 4)
        ctable.append(row)
    by_tag = {"%s|%s" % k: round(v[1] / steps, 3) for k, v in sorted(rec.items(), key=lambda kv: -kv[1][1])[:24]}
    return table, ctable, by_tag, total_ms / steps


def measured_traffic(cls, model_key, batch):
    """dram__bytes_read.sum + dram__bytes_write.sum per step of a kernel class, from the committed
    ncu capture of this same command (profiles/r02_dram_by_class.json), or None."""
    path = os.path.join(ROOT, "profiles", "r02_dram_by_class.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        d = json.load(f)
    if d.get("model") != model_key or d.get("batch") != batch:
        return None
    return d.get("classes", {}).get(cls, {}).get("dram_bytes_per_step")


# ------------------------------------------------------------------------------------------------
# reference arms: the UNMODIFIED reference (oracle/_ref, staged by oracle/make_ref.py) -- on the host
# cores (the CPU arm / cpu_baseline) or on the same B200 (the cuDNN bar); the oracle port only when
# oracle/_ref is absent
# ------------------------------------------------------------------------------------------------
def cpu_reference(kwargs, steps, warmup, batch, budget_s=200.0):
    from oracle import ref_runner as RR
    cores = os.cpu_count() or 1
    g = torch.Generator().manual_seed(0)
    kind = "reference" if RR.available() else "port"
    if kind == "reference":
        st = RR.train_state(kwargs, IMAGE, device="cpu", lr=0.016 * 8)

        def run(x, y):
            RR.train_step(st, x, y, WEIGHT_DECAY, RHO)
    else:
        from oracle import atomnas_oracle as O
        spec = O.NetSpec(**kwargs, input_size=IMAGE)
        ts = O.TrainState(spec, O.init_state(spec, 1995))

        def run(x, y):
            O.train_step(ts, x, y, 0.016 * 8, WEIGHT_DECAY, RHO, 0.1, 0.9999 ** 0.5)

    def data(b):
        return torch.randn(b, 3, IMAGE, IMAGE, generator=g), torch.randint(0, 1000, (b,), generator=g)

    def one(x, y):
        t0 = time.perf_counter()
        run(x, y)
        return time.perf_counter() - t0

    # Thread count: the oneDNN depthwise / 1x1 primitives at this batch stop scaling well before a
    # large host's core count (and oversubscribed runs are many times slower), so probe ASCENDING
    # thread counts {cores/16 .. cores} on a small batch and stop at the first one that is slower
    # or when 35 % of the time budget is gone.  Everything below is bounded by `budget_s`.
    t_start = time.perf_counter()
    xs, ys = data(min(16, batch))
    best_t, best_n = None, max(1, min(cores, 8))
    for n in sorted({max(min(cores, 8), cores // d) for d in (16, 8, 4, 2, 1)}):
        torch.set_num_threads(n)
        one(xs, ys)                                    # primitive creation per shape happens here
        t = one(xs, ys)
        if best_t is None or t < best_t:
            best_t, best_n = t, n
        elif t > 1.15 * best_t:
            break
        if time.perf_counter() - t_start > 0.35 * budget_s:
            break
    torch.set_num_threads(best_n)
    per_image = best_t / xs.shape[0]
    left = budget_s - (time.perf_counter() - t_start)
    # bounded sample: the largest batch in {64, 32, 16, 8} whose warm-up + steps fit what is left
    while batch > 8 and (steps + warmup + 1) * per_image * batch * 1.15 > left:
        batch //= 2
    x, y = data(batch)
    one(x, y)                                          # shapes of this batch
    for _ in range(warmup):
        one(x, y)
    times = []
    for _ in range(steps):
        times.append(one(x, y))
        if time.perf_counter() - t_start > 1.5 * budget_s:      # hard stop: report what was measured
            break
    sec = sum(times) / len(times)
    steps = len(times)
    what = ("unmodified reference modules (oracle/_ref: models.mobilenet_supernet, utils.rmsprop, utils.optim, "
            "utils.prune), fp32 as shipped") if kind == "reference" else "oracle port (fp32 torch CPU ops)"
    return {"value": batch / sec, "unit": "images/sec", "cores": best_n, "host_cores": cores, "kind": kind,
            "sample": "%d timed steps of batch %d after %d warm-up steps; %s; train.py:174-236 loop body; "
                      "thread count = fastest of cores/{1,2,4,8,16}" % (steps, batch, warmup, what)}, sec, batch


def gpu_reference(kwargs, steps, warmup, mode):
    """The reference's own nn.Modules / RMSprop / EMA loop on cuda:0 (cuDNN, ATen): 'fp32' = as
    shipped (train.py:255 sets cudnn.benchmark; TF32 convs are torch's default), 'bf16' = its best
    light (autocast bf16 + channels_last)."""
    from oracle import ref_runner as RR
    torch.backends.cudnn.benchmark = True
    st = RR.train_state(kwargs, IMAGE, device="cuda", lr=0.016, channels_last=(mode == "bf16"))
    gen = torch.Generator(device="cuda").manual_seed(1995)
    x = torch.randn(PER_GPU_BATCH, 3, IMAGE, IMAGE, device="cuda", generator=gen)
    y = torch.randint(0, 1000, (PER_GPU_BATCH,), device="cuda", generator=gen)
    for _ in range(warmup):
        RR.train_step(st, x, y, WEIGHT_DECAY, RHO, autocast_bf16=(mode == "bf16"))
    return timed_loop(lambda: RR.train_step(st, x, y, WEIGHT_DECAY, RHO, autocast_bf16=(mode == "bf16")), steps, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-gpu"])
    ap.add_argument("--ref-mode", default="bf16", choices=["fp32", "bf16"])
    ap.add_argument("--model", default="atomnas_c", choices=sorted(MODELS))
    ap.add_argument("--masks", default="dense", choices=["dense", "random", "searched"])
    ap.add_argument("--accum", type=int, default=1, help="micro-batches of 256 per optimizer step ((S) series)")
    ap.add_argument("--no-graph", action="store_true", help="do not replay the step from a CUDA graph")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(3, args.warmup)
    model_name, kwargs = MODELS[args.model]
    masks_txt = {"dense": "dense (no shrink yet)", "random": "shrunk with the random keep-0.5 mask fixture",
                 "searched": "shrunk to the searched AtomNAS-C widths (apps/searched/models/atomnas_c.yml)"}[args.masks]
    config = {"workload": "%s training step, synthetic 3x224x224, %d images per GPU%s, dynamic BN-gamma pruning "
                          "term active (rho=1e-4), %s" % (model_name, PER_GPU_BATCH,
                                                         " x %d micro-batches" % args.accum if args.accum > 1 else "",
                                                         masks_txt),
              "global_batch": PER_GPU_BATCH * world * args.accum, "per_gpu_batch": PER_GPU_BATCH, "image": IMAGE,
              "parallelism": "dp%d" % world, "model": args.model, "masks": args.masks, "accum": args.accum,
              "cache": "inputs larger than L2 (154 MB images + ~14 GB of saved activations per step)"}

    if args.impl == "reference":
        if rank != 0:
            return
        base, sec, batch = cpu_reference(kwargs, max(1, args.steps), max(0, args.warmup), 64)
        config = dict(config, sample="reference arm: each step is a bounded sample of %d images of the %d-image "
                                     "workload on the host cores" % (batch, PER_GPU_BATCH))
        out = {"impl": "reference", "metric": METRIC, "value": base["value"],
               "unit": "images/sec", "n_gpus": args.gpus, "steps": max(1, args.steps), "warmup": max(0, args.warmup),
               "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f32", "data": "synthetic", "config": config, "cpu_baseline": base,
               "e2e": {"value": base["value"], "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(out))
        return

    torch.cuda.set_device(local_rank)
    if args.impl == "reference-gpu":
        if rank != 0:
            return
        sampler = ClockSampler(local_rank)
        sampler.start()
        sec = gpu_reference(kwargs, args.steps, warmup, args.ref_mode)
        clocks = sampler.stop()
        out = {"impl": "reference-gpu", "ref_mode": args.ref_mode, "metric": METRIC,
               "value": PER_GPU_BATCH * args.steps / sec, "unit": "images/sec", "n_gpus": 1, "steps": args.steps,
               "warmup": warmup, "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f32 (TF32 convs)" if args.ref_mode == "fp32" else "bf16 autocast",
               "data": "synthetic", "config": config, "clocks": clocks,
               "note": "unmodified reference nn.Modules + utils.rmsprop + EMA loop (oracle/_ref) on cuda:0 through "
                       "PyTorch eager / cuDNN; %s" % ("as shipped (fp32 NCHW, cudnn.benchmark)" if args.ref_mode == "fp32"
                                                      else "autocast bf16 + channels_last + cudnn.benchmark"),
               "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 1e9, 2)}
        print(json.dumps(out))
        return

    if world > 1:
        dist.init_process_group("nccl")
    from atomnas_b200 import _lib as L
    peaks = load_peaks()
    model = build_model(kwargs).cuda()
    if world > 1:                                     # replicas start identical (utils/distributed.py:183-190)
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, 0)
    model.train()
    trainer, info = build_trainer(model, world, use_graph=not args.no_graph)
    shrink_info = None
    if args.masks != "dense":
        shrink_info = trainer.shrink(masks=mask_fixture(model, args.masks), prune_info=info)
        from atomnas_b200 import prune
        shrink_info["macs_after"] = int(prune.profile_macs(model, IMAGE))
    gen = torch.Generator(device="cuda").manual_seed(1995 + rank)
    x = torch.randn(PER_GPU_BATCH, 3, IMAGE, IMAGE, device="cuda", generator=gen)
    y = torch.randint(0, 1000, (PER_GPU_BATCH,), device="cuda", generator=gen)
    x_host = x.cpu().pin_memory()
    y_host = y.cpu().pin_memory()

    if args.accum > 1:
        def one_step():
            return trainer.step_accum([(x, y)] * args.accum, rho=RHO)
    else:
        def one_step():
            return trainer.step(x, y, rho=RHO)

    for _ in range(warmup):
        one_step()
    torch.cuda.synchronize()
    launches0 = L.lib().an_launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    sec = timed_loop(one_step, args.steps, world)
    clocks = sampler.stop() if rank == 0 else None
    launches = L.lib().an_launch_count() - launches0
    per_step_launches = trainer.launches_per_step if getattr(trainer, "launches_per_step", None) else launches // args.steps
    per_step_launches *= args.accum if trainer.launches_per_step and args.accum > 1 else 1

    # end to end through the public API: pinned host -> device copies and a loss read every step
    last = [None]

    def e2e_step():
        # the next batch's H2D copy runs on a side stream under this step (DataPrefetcher pattern)
        if args.accum > 1:
            last[0] = trainer.step_accum_host([(x_host, y_host)] * args.accum, rho=RHO).item()
        else:
            last[0] = trainer.step_host(x_host, y_host, rho=RHO, prefetch=(x_host, y_host)).item()

    for _ in range(2):
        e2e_step()
    sec_e2e = timed_loop(e2e_step, args.steps, world)
    loss_value = last[0]

    # per-kernel pass: eager launches bracketed by CUDA events.  EVERY rank runs it (the step holds
    # collectives); only rank 0 reports.
    prof = None
    if not args.no_profile and args.accum == 1:
        was = trainer.use_cuda_graph
        trainer.use_cuda_graph = False               # per-kernel CUDA events need real launches
        prof = profile_kernels(trainer, x, y, peaks)
        trainer.use_cuda_graph = was
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    imgs = PER_GPU_BATCH * world * args.accum
    out = {"metric": METRIC, "value": imgs * args.steps / sec,
           "unit": "images/sec", "n_gpus": world, "steps": args.steps, "warmup": warmup,
           "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "bf16", "data": "synthetic", "config": config, "clocks": clocks,
           "e2e": {"value": imgs * args.steps / sec_e2e, "unit": "images/sec",
                   "h2d_bytes_per_step": (x_host.numel() * 4 + y_host.numel() * 8) * args.accum, "d2h_bytes_per_step": 4,
                   "ms_per_step": sec_e2e / args.steps * 1e3},
           "gpu_launches": int(per_step_launches * args.steps), "gpu_launches_per_step": int(per_step_launches),
           "cuda_graph": bool(trainer.use_cuda_graph), "final_loss": loss_value, "peaks": peaks}
    if shrink_info is not None:
        out["shrink"] = shrink_info
    if prof is not None:
        table, ctable, by_tag, prof_ms = prof
        out["kernels"] = table
        out["kernel_classes"] = ctable
        out["kernel_stage_ms"] = by_tag
        top = next((r for r in ctable if "algorithmic_GBps" in r), None)
        if top is not None:
            traffic = measured_traffic(top["class"], args.model, PER_GPU_BATCH) if args.masks == "dense" else None
            out["roofline"] = {
                "kernel": top["class"], "bound": "hbm", "achieved": top["algorithmic_GBps"],
                "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": top["frac_hbm"], "traffic": traffic,
                "peak_source": peaks["source"] + " (STREAM-style copy, MEASURED_PEAKS.json)",
                "share_of_step": top["share"], "algorithmic_bytes_per_step": int(top["algorithmic_GB"] * 1e9),
                "accounting": "SURVEY 8(d): every tensor once per pass (forward, data gradient, weight gradient)",
                "note": ("depthwise k3/5/7 family (an_dw_*): CUDA cores, HBM roof; k=7 is FP32-FMA bound at 49 FMA per "
                         "4 bytes (FMA roof in DESIGN.md / profiles/r02_fma_probe.txt)") if top["class"] == "depthwise"
                else "pointwise GEMMs: arithmetic intensity below the machine balance, HBM is the binding roof; "
                     "tensor-pipe fraction %.4f of sustained bf16" % top["frac_tensor_sustained"]}
    if world == 1 and not args.no_cpu_baseline:
        try:
            out["cpu_baseline"], _, _ = cpu_reference(kwargs, 2, 1, 64, budget_s=60.0)
        except Exception as ex:                         # the baseline must never sink the GPU number
            out["cpu_baseline"] = {"error": repr(ex)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
