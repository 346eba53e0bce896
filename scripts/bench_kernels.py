"""Kernel / shape sweep (BASELINE.json configs[4]): every (C, H, k, stride) depthwise layer and every
pointwise GEMM shape of the AtomNAS-C supernet at 256 images, each kernel timed alone with CUDA
events on rotating buffers larger than L2, reported as algorithmic GB/s and fraction of the measured
HBM peak.   python scripts/bench_kernels.py [dw|pw|all] [batch] [--json out.json] [--res 224] [--plain]
--res R scales every feature map by R/224 (BASELINE config 5: 128 / 160 / 192 / 224); --plain feeds
the depthwise kernels materialised operands instead of BN + activation on load (the engine's path).
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from atomnas_b200 import kernels as K  # noqa: E402

BF = torch.bfloat16
DEV = "cuda"
PEAK = bench.load_peaks()["hbm_gbs"]

# (C_in of block, hidden per branch, H, stride, C_out) for the AtomNAS-C supernet @224
BLOCKS = [(16, 96, 112, 2, 24), (24, 144, 56, 1, 24), (24, 144, 56, 2, 40), (40, 240, 28, 1, 40),
          (40, 240, 28, 2, 80), (80, 480, 14, 1, 80), (80, 480, 14, 1, 96), (96, 576, 14, 1, 96),
          (96, 576, 14, 2, 192), (192, 1152, 7, 1, 192), (192, 1152, 7, 1, 320)]


def timeit(fn, nrot, iters=None):
    fn(0)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = iters or max(4, 2 * nrot)
    s.record()
    for i in range(iters):
        fn(i % nrot)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def nrot_for(nbytes):
    return max(2, min(8, int(400e6 // max(nbytes, 1)) + 1))


def dw_sweep(B, out, res=224, plain=False):
    seen = set()
    for cin, h, H, stride, cout in BLOCKS:
        H = H * res // 224
        for k in (3, 5, 7):
            key = (h, H, k, stride)
            if key in seen:
                continue
            seen.add(key)
            C = h
            Ho = H // stride
            M, Mo = B * H * H, B * Ho * Ho
            nr = nrot_for(2 * M * C)
            x = [torch.randn(M, C, device=DEV).to(BF) for _ in range(nr)]
            y = [torch.empty(Mo, C, device=DEV, dtype=BF) for _ in range(nr)]
            dd = [torch.randn(Mo, C, device=DEV).to(BF) for _ in range(nr)]
            ge = [torch.empty(M, C, device=DEV, dtype=BF) for _ in range(nr)]
            w = torch.randn(k * k, C, device=DEV)
            dwt = torch.empty(C * k * k, device=DEV)
            sc = torch.ones(C, device=DEV)
            sh = torch.zeros(C, device=DEV)
            part = torch.empty(max(K.dw_fwd_stats_rows(B, C, Ho, Ho, k, stride, False),
                                   K.dw_bwd_data_stats_rows(B, C, H, H, k, stride, False)) * 2 * C + 8, device=DEV)
            wpart = torch.empty(K.dw_bwd_weight_rows(B, C, H, H, k, stride, False) * k * k * C + 8, device=DEV)

            def f_fwd(i):
                K.dw_fwd(x[i], y[i], w, B, H, H, C, k, stride, ldx=C, ldy=C, act=1, stats_part=part,
                         scale=None if plain else sc, shift=None if plain else sh)

            def f_bd(i):
                K.dw_bwd_data(dd[i], y[i], x[i], ge[i], w, None, None, None, sc, sh, sh, sc, part, 1,
                              B, H, H, C, k, stride, ldg=C, ldd=C, lde=C, ldge=C)

            def f_bw(i):
                K.dw_bwd_weight(dd[i], y[i], x[i], None, None, None, None if plain else sc, None if plain else sh,
                                0 if plain else 1, wpart, dwt, B, H, H, C, k, stride, ldg=C, ldd=C, lde=C)

            for name, fn, nbytes in (("dw_fwd", f_fwd, 2 * (M + Mo) * C), ("dw_bwd_data", f_bd, 2 * (Mo + 2 * M) * C),
                                     ("dw_bwd_weight", f_bw, 2 * (Mo + M) * C)):
                ms = timeit(fn, nr)
                gbs = nbytes / ms / 1e6
                gfma = Mo * C * k * k / ms / 1e6
                row = {"kernel": name, "C": C, "H": H, "k": k, "stride": stride, "res": res, "ms": round(ms, 4),
                       "GBps": round(gbs, 1), "frac_hbm": round(gbs / PEAK, 3), "GFMAps": round(gfma, 1)}
                out.append(row)
                print(row, flush=True)
            del x, y, dd, ge


def pw_sweep(B, out):
    for cin, h, H, stride, cout in BLOCKS:
        Kh = 3 * h
        Ho = H // stride
        M, Mo = B * H * H, B * Ho * Ho
        nr = nrot_for(2 * M * Kh)
        # expand forward: [M, cin] x [Kh, cin]^T -> e_raw [M, Kh] + stats
        x = [torch.randn(M, cin, device=DEV).to(BF) for _ in range(nr)]
        e = [torch.empty(M, Kh, device=DEV, dtype=BF) for _ in range(nr)]
        we = torch.randn(Kh, cin, device=DEV).to(BF)
        weT = torch.randn(cin, Kh, device=DEV).to(BF)
        part = torch.empty(max(K.pw_gemm_stats_rows(M, Kh), K.pw_gemm_stats_rows(Mo, Kh)) * 2 * Kh + 8, device=DEV)
        v = torch.ones(Kh, device=DEV)
        z = torch.zeros(Kh, device=DEV)
        ms = timeit(lambda i: K.pw_gemm(x[i], we, e[i], M, Kh, cin, stats_part=part), nr)
        rows = [("pw_expand_fwd", ms, 2 * M * (cin + Kh), M, Kh, cin)]
        ms = timeit(lambda i: K.pw_gemm(x[i], we, e[i], M, Kh, cin), nr)       # store-only ceiling of the epilogue
        rows.append(("pw_expand_fwd_nostats", ms, 2 * M * (cin + Kh), M, Kh, cin))
        # expand dgrad: a_mode 2 on (g_e, e_raw): [M, Kh] x [cin, Kh]^T -> dx [M, cin]
        ms = timeit(lambda i: K.pw_gemm(e[i], weT, x[i], M, cin, Kh, x2=e[(i + 1) % nr], a_mode=2, a1=v, a2=z, a3=z), nr)
        rows.append(("pw_expand_dgrad", ms, 2 * M * (2 * Kh + cin), M, cin, Kh))
        del x, e
        d = [torch.randn(Mo, Kh, device=DEV).to(BF) for _ in range(nr)]
        gd = [torch.empty(Mo, Kh, device=DEV, dtype=BF) for _ in range(nr)]
        pr = [torch.empty(Mo, cout, device=DEV, dtype=BF) for _ in range(nr)]
        wp = torch.randn(cout, Kh, device=DEV).to(BF)
        wpT = torch.randn(Kh, cout, device=DEV).to(BF)
        ms = timeit(lambda i: K.pw_gemm(d[i], wp, pr[i], Mo, cout, Kh, a_mode=1, a_act=1, a1=v, a3=z, stats_part=part), nr)
        rows.append(("pw_project_fwd", ms, 2 * Mo * (Kh + cout), Mo, cout, Kh))
        ms = timeit(lambda i: K.pw_gemm(pr[i], wpT, gd[i], Mo, Kh, cout, epi_mode=1, r=d[i], ldr=Kh, r_scale=v, r_shift=z,
                                        r_mean=z, r_rstd=v, r_act=1, stats_part=part), nr)
        rows.append(("pw_project_dgrad", ms, 2 * Mo * (cout + 2 * Kh), Mo, Kh, cout))
        wpart = torch.empty(K.pw_wgrad_splits(Mo, Kh, cout) * Kh * cout + 8, device=DEV)
        ms = timeit(lambda i: K.pw_wgrad(d[i], pr[i], wpart, Mo, Kh, cout, p_mode=1, p_act=1, a1=v, a3=z), nr)
        rows.append(("pw_project_wgrad", ms, 2 * Mo * (Kh + cout), Mo, Kh, cout))
        ms = timeit(lambda i: K.affine_act(d[i], gd[i], Mo, Kh, ld1=Kh, ldo=Kh, a1=v, a3=z, act=1), nr)
        rows.append(("affine_act", ms, 2 * Mo * 2 * Kh, Mo, Kh, 0))
        for name, ms, nbytes, m, n, k in rows:
            gbs = nbytes / ms / 1e6
            row = {"kernel": name, "M": m, "N": n, "K": k, "ms": round(ms, 4), "GBps": round(gbs, 1),
                   "frac_hbm": round(gbs / PEAK, 3), "TFLOPs": round(2.0 * m * n * k / ms / 1e9, 1)}
            out.append(row)
            print(row, flush=True)
        del d, gd, pr


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    out = []
    res = int(sys.argv[sys.argv.index("--res") + 1]) if "--res" in sys.argv else 224
    if which in ("dw", "all"):
        dw_sweep(B, out, res, "--plain" in sys.argv)
    if which in ("pw", "all"):
        pw_sweep(B, out)
    tot = {}
    for r in out:
        tot[r["kernel"]] = tot.get(r["kernel"], 0.0) + r["ms"]
    print("sum of ms per kernel family (one launch per distinct shape):", {k: round(v, 3) for k, v in tot.items()})
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
            json.dump({"batch": B, "res": res, "peak_hbm_gbs": PEAK, "rows": out,
                       "dw_occupancy": os.environ.get("AN_DW_OCC", "default")}, f, indent=1)


if __name__ == "__main__":
    main()
