"""Aggregate an ncu launch list (csv with gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum
per launch) of ONE eager training step into per-class totals:

    python scripts/ncu_by_class.py gpurun_out/r02_launches.csv profiles/r02_dram_by_class.json atomnas_c 256 [steps]

The csv is produced on the GPU box by
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file gpurun_out/r02_launches.csv python scripts/profile_step.py 256 1
bench.py reads the json for `roofline.traffic` (DRAM bytes per step of the dominant kernel class)."""
import csv
import json
import re
import sys

CLASS_OF = [("dwm_conv", "depthwise"), ("dwm_wgrad", "depthwise"), ("dw2_conv", "depthwise"), ("dw2_wgrad", "depthwise"), ("dw_dgrad2", "depthwise"), ("dw2::wgrad_reduce", "depthwise"),
            ("pw_gemm", "pointwise"), ("pw_wgrad", "pointwise"), ("wgrad_reduce", "pointwise"),
            ("affine_act", "elementwise"), ("col_reduce", "elementwise"), ("act_bwd_reduce", "elementwise"),
            ("bn_finalize", "bn_finalize"), ("bn_bwd_finalize", "bn_finalize")]


def main():
    src, dst, model, batch = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
    steps = int(sys.argv[5]) if len(sys.argv) > 5 else 1
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    hdr = rows[0]
    k_name, k_metric, k_val, k_unit = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    k_id = hdr.index("ID")
    per = {}
    for r in rows[1:]:
        e = per.setdefault(r[k_id], {"name": r[k_name]})
        v = float(r[k_val].replace(",", ""))
        unit = r[k_unit].lower()
        if "byte" in unit:
            v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]
        elif unit in ("usecond", "us"):
            v *= 1e3
        elif unit in ("msecond", "ms"):
            v *= 1e6
        e[r[k_metric]] = v
    classes, kernels = {}, {}
    for e in per.values():
        cls = next((c for pat, c in CLASS_OF if pat in e["name"]), "other")
        if "at::" in e["name"] or "at_cuda_detail" in e["name"]:
            cls = "torch_setup"          # arena / optimizer-state / input initialisation of the process, not the step
        short = re.sub(r"\(.*", "", e["name"]).replace("void ", "").replace("an::", "")[:60]
        for table, key in ((classes, cls), (kernels, short)):
            t = table.setdefault(key, {"launches": 0, "time_ns": 0.0, "dram_read": 0.0, "dram_write": 0.0})
            t["launches"] += 1
            t["time_ns"] += e.get("gpu__time_duration.sum", 0.0)
            t["dram_read"] += e.get("dram__bytes_read.sum", 0.0)
            t["dram_write"] += e.get("dram__bytes_write.sum", 0.0)
    total = sum(t["time_ns"] for c, t in classes.items() if c != "torch_setup")
    out = {"model": model, "batch": batch, "steps_captured": steps,
           "how": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                  "on scripts/profile_step.py (one eager step; per-launch times are cold-cache and serialised: compare shares)",
           "classes": {c: {"launches_per_step": t["launches"] // steps, "time_ms_per_step": t["time_ns"] / 1e6 / steps,
                           "share": t["time_ns"] / total, "dram_read_bytes_per_step": t["dram_read"] / steps,
                           "dram_write_bytes_per_step": t["dram_write"] / steps,
                           "dram_bytes_per_step": (t["dram_read"] + t["dram_write"]) / steps}
                       for c, t in sorted(classes.items(), key=lambda kv: -kv[1]["time_ns"])},
           "kernels": {k: {"launches_per_step": t["launches"] // steps, "time_ms_per_step": t["time_ns"] / 1e6 / steps,
                           "dram_GB_per_step": (t["dram_read"] + t["dram_write"]) / 1e9 / steps}
                       for k, t in sorted(kernels.items(), key=lambda kv: -kv[1]["time_ns"])[:40]}}
    json.dump(out, open(dst, "w"), indent=1)
    for c, t in out["classes"].items():
        print("%-12s %6.2f ms  share %.3f  dram %.2f GB" % (c, t["time_ms_per_step"], t["share"], t["dram_bytes_per_step"] / 1e9))


if __name__ == "__main__":
    main()
