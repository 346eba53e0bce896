import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from atomnas_b200 import kernels as K
torch.manual_seed(0)
for (M, N, Kk) in [(128, 16, 32), (128, 64, 32), (256, 128, 64), (384, 288, 16)]:
    x = torch.randn(M, Kk, device="cuda").to(torch.bfloat16)
    w = torch.randn(N, Kk, device="cuda").to(torch.bfloat16)
    y = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    rows = K.pw_gemm_stats_rows(M, N)
    part = torch.full((rows, 2, N), float("nan"), device="cuda")
    dbg = torch.full((128, 32), -7.0, device='cuda')
    K.pw_gemm(x, w, y, M, N, Kk, stats_part=part, r_act=77, r_mean=dbg)
    torch.cuda.synchronize()
    yf = y.float()
    ref = x.float() @ w.float().t()
    print("shape", M, N, Kk, "out err", (yf - ref).abs().max().item())
    s = part.sum(0)
    exp2 = (yf * yf).sum(0)
    print(" s1 err", (s[0] - yf.sum(0)).abs().max().item())
    rel = ((s[1] - exp2).abs() / exp2.abs().clamp(min=1)).cpu()
    print(" s2 rel err per 32-col block:", [round(rel[i:i+32].max().item(), 4) for i in range(0, N, 32)])
    print(" s2 got", s[1][:8].tolist())
    print(" s2 exp", exp2[:8].tolist())
    G = yf[:128].t() @ yf[:128]          # first tile
    torch.cuda.synchronize()
    for qq in range(min(4, N // 32 if N >= 32 else 1)):
        blk = dbg[qq * 32:(qq + 1) * 32]
        ref = G[qq * 32:(qq + 1) * 32, qq * 32:(qq + 1) * 32] if N >= 32 else None
        print("  q", qq, "dbg[0,:4]", blk[0, :4].tolist(), "G", (G[qq * 32, qq * 32:qq * 32 + 4].tolist() if N > qq * 32 else None))
        if N >= 32 * (qq + 1):
            print("    block err", (blk - G[qq * 32:(qq + 1) * 32, qq * 32:(qq + 1) * 32]).abs().max().item())
    # which entry of the Gram matrix did we get for column c?
    for c in range(min(N, 8)):
        d = (G[c] - s[1][c]).abs()
        j = int(d.argmin())
        print("  col", c, "matches G[c][%d]" % j, "diff", d.min().item())
