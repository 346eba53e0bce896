"""Pin the CPU oracle (oracle/atomnas_oracle.py) to outputs of the reference itself.

The fixtures in tests/golden/ were produced by tests/golden/make_golden.py, which imports the
unmodified reference from /root/reference.  Everything here runs on CPU.
"""
import collections
import copy
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import atomnas_oracle as O  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")


def load(name):
    return torch.load(os.path.join(G, name + ".pt"), weights_only=False)


def clone(sd):
    return collections.OrderedDict((k, v.clone()) for k, v in sd.items())


def close(a, b, rtol=2e-4, atol=2e-5):
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol)


@pytest.mark.parametrize("name", ["tiny_supernet_relu6", "tiny_supernet_swish", "tiny_searched"])
def test_forward_backward_matches_reference(name):
    g = load(name)
    spec = O.NetSpec(**g["kwargs"], input_size=g["input_size"])
    P, B = O.param_shapes(spec)
    assert list(P.keys()) + list(B.keys()) != []
    assert set(P.keys()) | set(B.keys()) == set(g["state0"].keys())
    for n, s in list(P.items()) + list(B.items()):
        assert tuple(g["state0"][n].shape) == tuple(s), n
    st = clone(g["state0"])
    for n in P:
        st[n].requires_grad_(True)
    logits = O.forward(spec, st, g["x"], True)
    close(logits, g["logits"])
    ce = O.cross_entropy_label_smooth(logits, g["y"], spec.num_classes, 0.1)
    close(ce, g["ce"])
    ce.mean().backward()
    for n in P:
        close(st[n].grad, g["grads"][n], rtol=1e-3, atol=1e-5)
    for n in B:      # running statistics and counters after one training forward
        close(st[n].float(), g["state_after"][n].float())
    st2 = clone(g["state_after"])
    with torch.no_grad():
        close(O.forward(spec, st2, g["x"], False), g["logits_eval"])


def test_train_steps_match_reference():
    g = load("tiny_supernet_train")
    spec = O.NetSpec(**g["kwargs"], input_size=g["input_size"])
    assert O.network_macs(spec) == g["n_macs"]
    names, pen, pcf = O.prune_targets(spec)
    assert names == g["prune_names"]
    assert pen == pytest.approx(g["penalties"], rel=1e-12)
    assert pcf == pytest.approx(g["per_channel_flops"], rel=1e-12)
    h = spec.input_size // 2
    for b in spec.blocks:
        assert O.block_op_macs(spec, b, h) == g["block_macs"][b.name]
        h //= b.stride
    ts = O.TrainState(spec, clone(g["state0"]))
    for i, rec in enumerate(g["steps"]):
        loss, logits, ce = O.train_step(ts, rec["x"], rec["y"], g["lr"], g["weight_decay"], g["rho"],
                                        0.1, g["ema_decay"])
        close(logits, rec["logits"], rtol=1e-3, atol=1e-4)
        assert float(loss) == pytest.approx(rec["loss"], rel=1e-4)
        for n, v in rec["state_after"].items():
            close(ts.state[n].detach().float(), v.float(), rtol=2e-3, atol=2e-5)
    for n in ts.names:
        close(ts.sq[n], g["square_avg"][n], rtol=2e-3, atol=1e-9)
        close(ts.buf[n], g["momentum_buffer"][n], rtol=2e-3, atol=1e-4)  # g~0 tensors are rounding noise / sqrt(eps)
    for n, v in g["ema"].items():
        close(ts.ema[n].float(), v.float(), rtol=1e-3, atol=1e-5)


def test_shrink_index_semantics():
    g = load("tiny_shrink")
    spec = O.NetSpec(**g["kwargs"], input_size=g["input_size"])
    st = g["state_before"]
    new_state = collections.OrderedDict()
    new_blocks = []
    for b in spec.blocks:
        nb, out = O.compress_block_state(b, st, g["masks"][b.name], b.expand)
        new_blocks.append(nb)
        new_state.update(out)
    for n, v in g["state_after"].items():
        if n in new_state:
            assert torch.equal(new_state[n], v), n       # index ops: bit exact
    assert set(k for k in g["state_after"] if ".ops." in k or "pw_bn" in k) == set(new_state.keys())
    got = [[b.oup, 1, b.stride, b.kernel_sizes, b.channels, b.expand] for b in new_blocks]
    assert got == g["network_after"]["inverted_residual_setting"]
    # the compacted network evaluates like the reference's rebuilt modules
    kw = dict(g["kwargs"])
    kw["inverted_residual_setting"] = got
    spec2 = O.NetSpec(**kw, input_size=g["input_size"])
    assert O.network_macs(spec2) == g["n_macs_after"]
    st2 = clone(g["state_after"])
    with torch.no_grad():
        close(O.forward(spec2, st2, g["x"], True), g["logits_train_after"])
    names, pen, _ = O.prune_targets(spec2)
    # PruneInfo re-inserts renamed entries, so only the SET of names is defined (utils/prune.py:60-74)
    assert set(names) == set(g["prune_names_after"])
    # ... and it keeps the ORIGINAL penalties through a shrink, under the renumbered names
    old_names, old_pen, _ = O.prune_targets(spec)
    old = dict(zip(old_names, old_pen))
    after = dict(zip(g["prune_names_after"], g["penalties_after"]))
    for b in spec.blocks:
        if not b.expand:
            continue
        nj = 0
        for j, m in enumerate(g["masks"][b.name]):
            if int(m.sum()) == 0:
                continue
            assert after["%s.ops.%d.1.1.weight" % (b.name, nj)] == old["%s.ops.%d.1.1.weight" % (b.name, j)]
            nj += 1


def test_known_answers():
    g = load("known_answers")
    got = [O.rho_at(i, 10, 1e-4, 1, 5, True) for i in range(0, 70, 3)]
    assert got == pytest.approx(g["rho_stepwise"], abs=0)
    got = [O.rho_at(i, 10, 1e-4, 1, 5, False) for i in range(0, 70, 3)]
    assert got == pytest.approx(g["rho_epochwise"], abs=0)
    close(O.cross_entropy_label_smooth(g["ce_logits"], g["ce_y"], 10, 0.1), g["ce"])
    sh = torch.ones(3)
    for i in range(1, 6):
        O.ema_update(sh, torch.full((3,), float(i)), O.ema_decay(0.99, i))
        close(sh, g["ema_vals"][i - 1], rtol=1e-6, atol=0)
    m = O.masks_by_threshold(g["mask_w"], 1.5)
    assert all(torch.equal(a, b) for a, b in zip(m, g["mask"]))
    # reference tests/utils/prune_test.py:54-64 known answer
    assert [x.tolist() for x in m] == [[False, True, True], [True, True, False, False]]


def test_full_size_macs():
    """[probed] MAC counts of SURVEY.md section 6 (utils/model_profiling.py on the real YAMLs)."""
    setting = [[1, 16, 1, 1, [3]], [6, 24, 4, 2, [3, 5, 7]], [6, 40, 4, 2, [3, 5, 7]],
               [6, 80, 4, 2, [3, 5, 7]], [6, 96, 4, 1, [3, 5, 7]], [6, 192, 4, 2, [3, 5, 7]],
               [6, 320, 1, 1, [3, 5, 7]]]
    c = O.NetSpec(inverted_residual_setting=setting, input_channel=32, input_size=224)
    a = O.NetSpec(inverted_residual_setting=setting, input_channel=16, input_size=224)
    assert O.network_macs(c) == 1521035776
    assert O.network_macs(a) == 1510599168
    P, _ = O.param_shapes(c)
    n = sum(int(torch.tensor(s).prod()) if len(s) else 1 for s in P.values())
    assert n == 11242152
    names, pen, pcf = O.prune_targets(c)
    assert len(names) == 63
    i = names.index("features.2.ops.0.1.1.weight")
    assert pcf[i] == 304192 and pen[i] == pytest.approx(6.457615, rel=1e-6)
    j = names.index("features.22.ops.2.1.1.weight")
    assert pcf[j] == 27489 and pen[j] == pytest.approx(0.583557, rel=1e-6)


@pytest.mark.parametrize("name", ["tiny_fused_se_swish", "tiny_fused_relu_s2"])
def test_fused_se_block_matches_reference(name):
    """Oracle restatement of InvertedResidualChannelsFused + SqueezeAndExcitation (scope table 8f-1,
    the AtomNAS-C+ block) against the reference's own module: output, input gradient, every
    parameter gradient, BN buffers after the step."""
    import types
    g = torch.load(os.path.join(G, name + ".pt"), weights_only=False)
    a = g["args"]
    spec = types.SimpleNamespace(bn_momentum=a["bn_momentum"], bn_eps=a["bn_eps"], active_fn=a["active_fn"])
    st = {"blk." + k: v.clone() for k, v in g["state0"].items()}
    params = [k for k, v in st.items() if v.dtype.is_floating_point and "running" not in k]
    for k in params:
        st[k].requires_grad_(True)
    x = g["x"].clone().requires_grad_(True)
    y = O.fused_block_forward(spec, st, "blk", x, a["inp"], a["oup"], a["stride"], a["channels"],
                              a["kernel_sizes"], a["expand"], a["se_ratio"], training=True)
    torch.testing.assert_close(y, g["y"], rtol=1e-5, atol=1e-5)
    (y * g["w"]).sum().backward()
    torch.testing.assert_close(x.grad, g["dx"], rtol=1e-4, atol=1e-5)
    assert set("blk." + n for n in g["grads"]) == set(params)
    for n, ref in g["grads"].items():
        torch.testing.assert_close(st["blk." + n].grad, ref, rtol=1e-4, atol=1e-5, msg=n)
    for k, v in g["state_after"].items():
        if "running" in k or "num_batches" in k:
            torch.testing.assert_close(st["blk." + k].detach(), v, rtol=1e-5, atol=1e-6, msg=k)


def test_bn_calibration_matches_reference():
    """BN re-calibration (train.py:419-458 -> utils/common.py:213-226): running statistics reset,
    momentum=None (cumulative average) over a few forward passes, then eval logits."""
    g = torch.load(os.path.join(G, "tiny_bn_calibration.pt"), weights_only=False)
    spec = O.NetSpec(**g["kwargs"], input_size=g["input_size"])
    st = {k: v.clone() for k, v in g["state0"].items()}
    for k in st:                                        # nn.BatchNorm2d.reset_running_stats()
        if k.endswith("running_mean"):
            st[k].zero_()
        elif k.endswith("running_var"):
            st[k].fill_(1.0)
        elif k.endswith("num_batches_tracked"):
            st[k].zero_()
    spec.bn_momentum = None
    with torch.no_grad():
        for x in g["xs"]:
            O.forward(spec, st, x, training=True)
    for k, v in g["state_calibrated"].items():
        if "running" in k:
            torch.testing.assert_close(st[k], v, rtol=1e-4, atol=1e-5, msg=k)
        elif "num_batches" in k:
            assert int(st[k]) == int(v) == len(g["xs"])
    with torch.no_grad():
        logits = O.forward(spec, st, g["xs"][0], training=False)
    torch.testing.assert_close(logits, g["logits_eval"], rtol=1e-4, atol=1e-4)
