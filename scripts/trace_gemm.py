"""Per-tile clock trace of the pointwise GEMM (diagnostic build libatomnas_b200_trace.so, -DAN_GEMM_TRACE)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from atomnas_b200 import kernels as K
BF = torch.bfloat16
which = sys.argv[1]
M, Kh, cout, cin = 256 * 56 * 56, 432, 24, 24
v = torch.ones(Kh, device="cuda"); z = torch.zeros(Kh, device="cuda")
part = torch.empty(K.pw_gemm_stats_rows(M, Kh) * 2 * Kh + 8, device="cuda")
if which == "project_dgrad":
    pr = torch.randn(M, cout, device="cuda").to(BF); wpT = torch.randn(Kh, cout, device="cuda").to(BF)
    d = torch.randn(M, Kh, device="cuda").to(BF); gd = torch.empty(M, Kh, device="cuda", dtype=BF)
    f = lambda: K.pw_gemm(pr, wpT, gd, M, Kh, cout, epi_mode=1, r=d, ldr=Kh, r_scale=v, r_shift=z, r_mean=z, r_rstd=v, r_act=1, stats_part=part)
elif which == "expand_fwd":
    x = torch.randn(M, cin, device="cuda").to(BF); we = torch.randn(Kh, cin, device="cuda").to(BF); e = torch.empty(M, Kh, device="cuda", dtype=BF)
    f = lambda: K.pw_gemm(x, we, e, M, Kh, cin, stats_part=part)
elif which == "project_fwd":
    d = torch.randn(M, Kh, device="cuda").to(BF); wp = torch.randn(cout, Kh, device="cuda").to(BF); pr = torch.empty(M, cout, device="cuda", dtype=BF)
    f = lambda: K.pw_gemm(d, wp, pr, M, cout, Kh, a_mode=1, a_act=1, a1=v, a3=z, stats_part=part)
elif which == "expand_dgrad":
    e = torch.randn(M, Kh, device="cuda").to(BF); e2 = torch.randn(M, Kh, device="cuda").to(BF); weT = torch.randn(cin, Kh, device="cuda").to(BF)
    x = torch.empty(M, cin, device="cuda", dtype=BF)
    f = lambda: K.pw_gemm(e, weT, x, M, cin, Kh, x2=e2, a_mode=2, a1=v, a2=z, a3=z)
else:
    x = torch.randn(M, cin, device="cuda").to(BF); we = torch.randn(Kh, cin, device="cuda").to(BF); e = torch.empty(M, Kh, device="cuda", dtype=BF)
    f = lambda: K.pw_gemm(x, we, e, M, Kh, cin)
f(); torch.cuda.synchronize()
