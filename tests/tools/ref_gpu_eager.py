"""The bar of SURVEY.md section 8(d): the reference's arithmetic as PyTorch eager ops (cuDNN / cuBLAS)
on the SAME GPU, same network / batch / step as bench.py.  The reference checkout does not exist on
the GPU box, so this runs the oracle's functional restatement of it (identical F.conv2d /
F.batch_norm / autograd calls, Python-loop RMSprop + EMA as in train.py) on CUDA tensors.
Not part of the product or of bench.py's contract; prints one JSON line per mode.

    python tests/tools/ref_gpu_eager.py [batch] [steps]
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import atomnas_oracle as O  # noqa: E402


def run(mode, batch, steps, warmup=3):
    torch.backends.cudnn.benchmark = True                 # train.py:255
    spec = O.NetSpec(**bench.ATOMNAS_C_SUPERNET, input_size=bench.IMAGE)
    state = {k: v.cuda() for k, v in O.init_state(spec, 1995).items()}
    ts = O.TrainState(spec, state)
    g = torch.Generator(device="cuda").manual_seed(1995)
    x = torch.randn(batch, 3, bench.IMAGE, bench.IMAGE, device="cuda", generator=g)
    y = torch.randint(0, 1000, (batch,), device="cuda", generator=g)
    if mode == "bf16_autocast_channels_last":
        x = x.contiguous(memory_format=torch.channels_last)

    def one():
        if mode == "fp32_tf32":
            O.train_step(ts, x, y, 0.016, bench.WEIGHT_DECAY, bench.RHO, 0.1, 0.9999 ** (batch / 4096.0))
        else:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                O.train_step(ts, x, y, 0.016, bench.WEIGHT_DECAY, bench.RHO, 0.1, 0.9999 ** (batch / 4096.0))

    for _ in range(warmup):
        one()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        one()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / steps
    return {"impl": "pytorch_eager_" + mode, "metric": "images/sec AtomNAS-C supernet fwd+bwd+RMSprop",
            "value": batch / ms * 1e3, "unit": "images/sec", "ms_per_step": ms, "batch": batch, "steps": steps,
            "max_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30, "torch": torch.__version__}


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    for mode in ("bf16_autocast_channels_last", "fp32_tf32"):
        t0 = time.time()
        try:
            out = run(mode, batch, steps)
        except Exception as ex:                               # one mode failing must not lose the other
            out = {"impl": "pytorch_eager_" + mode, "error": repr(ex)[:300]}
        out["wall_s"] = round(time.time() - t0, 1)
        print(json.dumps(out), flush=True)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
