"""CPU-side checks: the C ABI library loads and exports every declared symbol, and the host logic
(module surface, pruning bookkeeping, shrink index semantics, data-parallel helpers) matches the
reference's golden outputs.  No compute kernels are launched here."""
import collections
import os
import re
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from atomnas_b200 import _lib as L  # noqa: E402
from atomnas_b200 import compress_utils as cu  # noqa: E402
from atomnas_b200 import mobilenet_base as mb  # noqa: E402
from atomnas_b200 import mobilenet_supernet as ms  # noqa: E402
from atomnas_b200 import optim, prune  # noqa: E402
from atomnas_b200.rmsprop import RMSprop  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")


def load(name):
    return torch.load(os.path.join(G, name + ".pt"), weights_only=False)


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "atomnas_b200.h")).read()
    declared = set(re.findall(r"\b(an_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    handle = L.lib()
    for name in declared:
        assert getattr(handle, name) is not None
    assert handle.an_version() >= 100
    # planning helpers are pure host code and must agree with themselves without a GPU
    assert handle.an_pw_gemm_stats_rows(3211264, 432) > 0
    assert handle.an_pw_wgrad_splits(3211264, 432, 24) > 0
    assert handle.an_dw_fwd_stats_rows(256, 96, 56, 56, 7, 2, 1) > 0
    # every AtomNAS-C depthwise shape has a streaming-kernel plan (materialised operands, fused = 0)
    for C, H, s in [(32, 112, 1), (96, 112, 2), (144, 56, 1), (144, 56, 2), (240, 28, 1), (240, 28, 2),
                    (480, 14, 1), (576, 14, 1), (576, 14, 2), (1152, 7, 1)]:
        for k in (3, 5, 7):
            ho = H // s
            assert handle.an_dw_fwd_stats_rows(256, C, ho, ho, k, s, 0) > 0
            assert handle.an_dw_bwd_data_stats_rows(256, C, H, H, k, s, 0) > 0
            assert handle.an_dw_bwd_weight_rows(256, C, H, H, k, s, 0) > 0


def test_header_prototypes_match_ctypes_signatures():
    """Argument counts of include/atomnas_b200.h and of the ctypes binding must agree (an ABI drift
    between the two would corrupt the stack silently)."""
    header = open(os.path.join(ROOT, "include", "atomnas_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    protos = dict(re.findall(r"\b(an_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S))
    assert set(protos) == set(L.SIGNATURES)
    for name, args in protos.items():
        args = args.strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        assert n == len(L.SIGNATURES[name][1]), (name, n, len(L.SIGNATURES[name][1]))


def test_no_cpu_fallback():
    g = load("tiny_supernet_train")
    model = ms.MobileNetV2(**g["kwargs"], input_size=64)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(torch.zeros(1, 3, 64, 64))
    p = torch.nn.Parameter(torch.ones(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        RMSprop([p], lr=0.1).step()


@pytest.mark.parametrize("name,searched", [("tiny_supernet_train", False), ("tiny_searched", True)])
def test_module_surface_and_state_dict_keys(name, searched):
    g = load(name)
    cls = ms.MobileNetSearched if searched else ms.MobileNetV2
    model = cls(**g["kwargs"], input_size=g["input_size"])
    sd = model.state_dict()
    assert list(sd.keys()) == list(g["state0"].keys())
    for k, v in g["state0"].items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
    assert model.load_state_dict(g["state0"]).missing_keys == []
    blocks = model.get_named_block_list()
    assert list(blocks.keys()) == ["features.%d" % i for i in range(1, len(blocks) + 1)]
    b = blocks["features.2"]
    assert isinstance(b, mb.InvertedResidualChannels) and b.expand and b.stride == 2
    assert list(b.get_named_depthwise_bn("features.2").keys())[0] == "features.2.ops.0.1.1"
    net = mb.output_network(model)
    assert len(net["inverted_residual_setting"]) == len(blocks)


def test_penalties_and_macs_match_reference():
    g = load("tiny_supernet_train")
    model = ms.MobileNetV2(**g["kwargs"], input_size=g["input_size"])
    assert prune.profile_macs(model, g["input_size"]) == g["n_macs"]
    for name, block in model.get_named_block_list().items():
        assert [op.n_macs for op in block.ops] == g["block_macs"][name]
    info = prune.get_bn_to_prune(model, {"bn_prune_filter": "expansion_only_skip_expand1"}, verbose=False)
    assert info.weight == g["prune_names"]
    assert info.penalty == pytest.approx(g["penalties"], rel=1e-12)
    assert info.get_info_list("per_channel_flops") == pytest.approx(g["per_channel_flops"], rel=1e-12)
    k = load("known_answers")
    f = prune.get_rho_scheduler({"epoch_free": 1, "epoch_warmup": 5, "scheduler": "linear", "rho": 1e-4,
                                 "stepwise": True}, 10)
    assert [f(i) for i in range(0, 70, 3)] == k["rho_stepwise"]
    m = prune.cal_mask_network_slimming_by_threshold(k["mask_w"], 1.5)
    assert [x.tolist() for x in m] == [[False, True, True], [True, True, False, False]]
    torch.testing.assert_close(optim.CrossEntropyLabelSmooth(10, 0.1)(k["ce_logits"], k["ce_y"]), k["ce"])


def test_shrink_bookkeeping_matches_reference_bit_exact():
    """compress_by_mask with optimizer / EMA / PruneInfo on CPU tensors (pure index bookkeeping)."""
    g = load("tiny_shrink")
    model = ms.MobileNetV2(**g["kwargs"], input_size=g["input_size"])
    model.load_state_dict(g["state_before"])
    prune.profile_macs(model, g["input_size"])
    info = prune.get_bn_to_prune(model, {"bn_prune_filter": "expansion_only_skip_expand1"}, verbose=False)
    opt = RMSprop(model.parameters(), lr=0.016, alpha=0.9, momentum=0.9, eps=1e-3, eps_inside_sqrt=True)
    named = dict(model.named_parameters())
    for n, p in named.items():      # optimizer state as the reference had it before the shrink
        opt.state[p] = {"step": 2, "square_avg": g["sq_before"][n].clone(),
                        "momentum_buffer": torch.zeros_like(p)}
    ema = optim.ExponentialMovingAverage(0.99)
    for n, v in g["ema_before"].items():
        src = named[n] if n in named else dict(model.named_buffers())[n]
        ema.register(n, src)
        ema._shadow[n] = v.clone()
    for bname, block in model.get_named_block_list().items():
        block.compress_by_mask([m.clone() for m in g["masks"][bname]], ema=ema, optimizer=opt,
                               prune_info=info, prefix=bname, verbose=False)
    assert set(opt.param_groups[0]["params"]) == set(model.parameters())
    sd = model.state_dict()
    assert list(sd.keys()) == list(g["state_after"].keys())
    for k, v in g["state_after"].items():
        assert torch.equal(sd[k], v), k
    named = dict(model.named_parameters())
    for n, v in g["sq_after"].items():
        assert torch.equal(opt.state[named[n]]["square_avg"], v), n
    assert set(ema.average_names()) == set(g["ema_after"].keys())
    for n, v in g["ema_after"].items():
        assert torch.equal(ema.average(n), v), n
    assert set(info.weight) == set(g["prune_names_after"])
    assert dict(zip(info.weight, info.penalty)) == pytest.approx(dict(zip(g["prune_names_after"], g["penalties_after"])))
    assert mb.output_network(model)["inverted_residual_setting"] == g["network_after"]["inverted_residual_setting"]
    assert prune.profile_macs(model, g["input_size"]) == g["n_macs_after"]


# ------------------------------------------------------------------------------------------------
# world_size-2 gloo: the data-parallel helpers
# ------------------------------------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _dp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world))
    from atomnas_b200 import distributed as udist
    udist.init_dist(backend="gloo")
    torch.manual_seed(rank)
    lin = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.BatchNorm1d(5))
    wrapped = udist.AllReduceDistributedDataParallel(lin)           # broadcast from rank 0
    w0 = lin[0].weight.detach().clone()
    arena = torch.arange(24, dtype=torch.float32) * (rank + 1)      # gradient "arena" with two views
    lin[0].weight.grad = torch.full((5, 6), float(rank + 1))
    lin[0].bias.grad = torch.full((5,), float(rank + 1))
    lin[1].weight.grad = arena[0:5]
    lin[1].bias.grad = arena[8:13]
    udist.allreduce_grads(wrapped)
    lin[1].running_mean.fill_(rank + 1.0)
    udist.allreduce_bn(wrapped)
    out[rank] = (w0, lin[0].weight.grad.clone(), lin[1].weight.grad.clone(), lin[1].bias.grad.clone(),
                 lin[1].running_mean.clone(), udist.is_master(), udist.get_world_size())
    dist.destroy_process_group()


def test_data_parallel_helpers_gloo_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    a, b = out[0], out[1]
    assert torch.equal(a[0], b[0])                                   # parameters broadcast from rank 0
    for x, y in zip(a[1:5], b[1:5]):
        assert torch.equal(x, y)                                     # identical reduced values on both ranks
    assert torch.allclose(a[1], torch.full((5, 6), 1.5))             # mean of 1 and 2
    assert torch.allclose(a[2], torch.arange(5, dtype=torch.float32) * 1.5)      # arena views reduced in place
    assert torch.allclose(a[3], torch.arange(8, 13, dtype=torch.float32) * 1.5)
    assert torch.allclose(a[4], torch.full((5,), 1.5))
    assert a[5] is True and b[5] is False and a[6] == 2


_DROPIN_PROBE = r'''
import importlib, json, os, sys
os.chdir(sys.argv[1])
os.environ.setdefault("ARNOLD_OUTPUT", "/tmp/an_out")
os.environ.setdefault("DATA_LMDB", "/tmp/an_data")
sys.argv = ["train.py", "app:apps/slimming/shrink/atomnas_c.yml"]
from utils.config import FLAGS                      # always the REFERENCE's config loader
m = importlib.import_module(FLAGS.model)             # common.py:129-130
model = m.Model(**FLAGS.model_kwparams, input_size=FLAGS.image_size)
out = {"model_module": type(model).__module__,
       "keys": [[k, list(v.shape)] for k, v in model.state_dict().items()],
       "blocks": list(model.get_named_block_list().keys()),
       "rmsprop": importlib.import_module("utils.rmsprop").RMSprop.__module__,
       "ema": importlib.import_module("utils.optim").ExponentialMovingAverage.__module__,
       "prune": importlib.import_module("utils.prune").get_bn_to_prune.__module__,
       "dist": importlib.import_module("utils.distributed").allreduce_grads.__module__,
       "config": importlib.import_module("utils.config").__file__}
print("RESULT " + json.dumps(out))
'''


@pytest.mark.skipif(not os.path.isdir("/root/reference/apps"), reason="reference checkout not present")
def test_dropin_shims_build_the_yaml_model_like_the_reference():
    """train.py's model factory (common.py:129-130) on the UNCHANGED atomnas_c.yml: with dropin/ first
    on sys.path the reference's config loader and YAML produce OUR module tree, with exactly the
    reference's state_dict keys and shapes; the optimizer / EMA / prune / distributed imports resolve
    to this repo while utils.config still is the reference's file."""
    import json
    import subprocess
    ref = "/root/reference"

    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        script = os.path.join(tmp, "probe_train.py")      # stands in for train.py (same imports, same factory)
        with open(script, "w") as f:
            f.write(_DROPIN_PROBE)

        def probe(cmd, env_extra):
            env = dict(os.environ, **env_extra)
            r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
            lines = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")]
            assert lines, r.stderr[-2000:]
            return json.loads(lines[-1][7:])

        # through the launcher: dropin/ first, whatever directory the script lives in
        ours = probe([sys.executable, os.path.join(ROOT, "dropin", "launch.py"), script, ref],
                     {"ATOMNAS_REFERENCE": ref, "PYTHONPATH": ""})
        theirs = probe([sys.executable, script, ref], {"PYTHONPATH": ref})
    assert ours["model_module"] == "atomnas_b200.mobilenet_supernet"
    assert theirs["model_module"] == "models.mobilenet_supernet"
    assert ours["keys"] == theirs["keys"]
    assert ours["blocks"] == theirs["blocks"]
    for k in ("rmsprop", "ema", "prune", "dist"):
        assert ours[k].startswith("atomnas_b200."), (k, ours[k])
    assert ours["config"].startswith(ref)


@pytest.mark.parametrize("name", ["tiny_fused_se_swish", "tiny_fused_relu_s2"])
def test_fused_block_surface_matches_reference_state_dict(name):
    """The AtomNAS-C+ block's module surface (scope row 8f-1): same state_dict keys / shapes as the
    reference's InvertedResidualChannelsFused (+SE), loads its weights, and refuses CPU tensors (the
    block runs on the sm_100a kernels only -- no silent PyTorch fallback)."""
    from atomnas_b200 import fused_block as fb
    g = torch.load(os.path.join(G, name + ".pt"), weights_only=False)
    a = g["args"]
    blk = fb.InvertedResidualChannelsFused(a["inp"], a["oup"], a["stride"], a["channels"], a["kernel_sizes"],
                                           a["expand"], active_fn=mb.get_active_fn(a["active_fn"]),
                                           batch_norm_kwargs={"momentum": a["bn_momentum"], "eps": a["bn_eps"]},
                                           se_ratio=a["se_ratio"])
    sd = blk.state_dict()
    assert list(sd.keys()) == list(g["state0"].keys())
    for k, v in g["state0"].items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
    assert blk.load_state_dict(g["state0"]).missing_keys == []
    assert list(blk.get_named_depthwise_bn("p").keys()) == ["p.depth_ops.%d.1.1" % i for i in range(len(a["channels"]))]
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        blk(g["x"])


def test_oracle_bf16_points_follow_the_depthwise_dispatch_rule():
    """The oracle rounds the depthwise operand and taps to bf16 exactly on the layers the engine's forward
    runs on the tensor cores (host-side rule of the C-ABI library; no kernel is launched)."""
    from atomnas_b200 import kernels as K
    from oracle import atomnas_oracle as O
    if os.environ.get("AN_DW_MMA") is not None:
        pytest.skip("dispatch forced by AN_DW_MMA")
    for k in (3, 5, 7):
        for stride in (1, 2):
            for B in (1, 4, 8, 16, 256):
                for W in (4, 7, 8, 14, 28, 56, 112):
                    assert K.dw_on_tensor_cores(0, k, stride, B, W) == O.dw_on_tensor_cores(k, stride, B, W), (k, stride, B, W)
                    if stride == 2 or B < 8:
                        assert not K.dw_on_tensor_cores(1, k, stride, B, W)
