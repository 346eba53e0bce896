"""One AtomNAS-C supernet training step (for ncu): python scripts/profile_step.py [batch] [steps]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
bench.PER_GPU_BATCH = B
model = bench.build_model().cuda().train()
tr, _ = bench.build_trainer(model, 1, use_graph=False)
x = torch.randn(B, 3, 224, 224, device="cuda"); y = torch.randint(0, 1000, (B,), device="cuda")
for _ in range(steps):
    tr.step(x, y, rho=bench.RHO)
torch.cuda.synchronize()
print("done")
