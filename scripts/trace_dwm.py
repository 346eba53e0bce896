"""One tensor-core depthwise forward launch with the per-step clock trace on (AN_DWM_EXP=8 AN_DWM_WS=0 AN_DW_MMA=7)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from atomnas_b200 import kernels as K
B, C, H, k = 256, 144, 56, int(sys.argv[1]) if len(sys.argv) > 1 else 7
BF = torch.bfloat16
x = torch.randn(B * H * H, C, device="cuda").to(BF); y = torch.empty(B * H * H, C, device="cuda", dtype=BF)
w = torch.randn(k * k, C, device="cuda"); sc = torch.ones(C, device="cuda"); sh = torch.zeros(C, device="cuda")
part = torch.empty(K.dw_fwd_stats_rows(B, C, H, H, k, 1, False) * 2 * C + 8, device="cuda")
for _ in range(2):
    K.dw_fwd(x, y, w, B, H, H, C, k, 1, ldx=C, ldy=C, act=1, stats_part=part, scale=sc, shift=sh)
    torch.cuda.synchronize()
