"""The bench.py JSON contract, checked on the committed lines of the last GPU run (profiles/) and on
the argument parser -- no GPU needed.  (The reference arm itself runs in minutes, not in this suite.)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def last_line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_committed_bench_line_has_every_contract_key():
    d = last_line("r02_bench_n1.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches",
              "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["unit"] == "images/sec" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["dtype"] == "bf16" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "AtomNAS-C supernet" in d["config"]["workload"] and d["warmup"] >= 3
    assert abs(d["value"] - d["config"]["global_batch"] / d["ms_per_step"] * 1e3) / d["value"] < 1e-6
    e = d["e2e"]
    assert e["unit"] == "images/sec" and e["h2d_bytes_per_step"] == 256 * 3 * 224 * 224 * 4 + 256 * 8
    assert e["d2h_bytes_per_step"] > 0 and e["value"] > 0
    assert d["gpu_launches"] > 0 and d["gpu_launches"] == d["gpu_launches_per_step"] * d["steps"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3 and "traffic" in r
    c = d["cpu_baseline"]
    assert c["kind"] == "reference" and c["cores"] >= 1 and c["value"] > 0 and "unmodified reference" in c["sample"]
    assert r["kernel"] == "depthwise" and r["algorithmic_bytes_per_step"] > 0
    classes = {k["class"]: k for k in d["kernel_classes"]}
    assert {"depthwise", "pointwise", "elementwise", "bn_finalize"} <= set(classes)
    assert abs(sum(k["ms_per_step"] for k in d["kernel_classes"]) - d["ms_per_step"]) / d["ms_per_step"] < 0.15
    assert d["clocks"]["reasons"] == [] and d["clocks"]["sm_mhz"] > 0.8 * d["clocks"]["sm_max_mhz"]


def test_committed_reference_arm_line():
    d = last_line("r02_bench_reference.json")
    assert d["impl"] == "reference" and d["unit"] == "images/sec" and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"]
    assert d["config"]["workload"] == last_line("r02_bench_n1.json")["config"]["workload"]


def test_committed_series_lines():
    """AtomNAS-A, the shrink series (mask fixtures), the (S) fixed-global-batch series and the cuDNN bar."""
    base = last_line("r02_bench_n1.json")
    a = last_line("r02_bench_atomnas_a.json")
    assert a["config"]["model"] == "atomnas_a" and a["value"] > 0 and a["gpu_launches"] > 0
    rnd, srch = last_line("r02_bench_masks_random.json"), last_line("r02_bench_masks_searched.json")
    assert rnd["config"]["masks"] == "random" and srch["config"]["masks"] == "searched"
    assert srch["value"] > rnd["value"] > base["value"]          # shrinking speeds the step up
    acc = last_line("r02_bench_accum8.json")
    assert acc["config"]["accum"] == 8 and acc["config"]["global_batch"] == 8 * 256
    assert abs(acc["value"] - base["value"]) / base["value"] < 0.05
    for name, mode in (("r02_bench_refgpu_bf16.json", "bf16"), ("r02_bench_refgpu_fp32.json", "fp32")):
        g = last_line(name)
        assert g["impl"] == "reference-gpu" and g["value"] > 0 and base["value"] / g["value"] > 3.0


def test_committed_multi_gpu_line():
    base, d = last_line("r02_bench_n1.json"), last_line("r02_bench_n8.json")
    assert d["n_gpus"] == 8 and d["scaling"] == "weak" and d["config"]["global_batch"] == 8 * 256
    assert d["config"]["parallelism"] == "dp8" and d["gpu_launches"] > 0 and d["e2e"]["value"] > 0
    assert 0.95 <= d["value"] / (8 * base["value"]) <= 1.02          # weak scaling: max over ranks of the step time
    r = last_line("r02_bench_reference_n8.json")
    assert r["impl"] == "reference" and r["value"] > 0 and r["e2e"]["h2d_bytes_per_step"] == 0


def test_bench_defaults_and_workload():
    import bench
    assert bench.PER_GPU_BATCH == 256 and bench.IMAGE == 224
    # the workload is the reference's AtomNAS-C supernet (supernet_single_path_nas_relu.yml)
    irs = bench.ATOMNAS_C_SUPERNET["inverted_residual_setting"]
    assert [r[1] for r in irs] == [16, 24, 40, 80, 96, 192, 320] and irs[1][4] == [3, 5, 7]
    assert bench.ATOMNAS_C_SUPERNET["input_channel"] == 32 and bench.ATOMNAS_C_SUPERNET["active_fn"] == "nn.ReLU"
