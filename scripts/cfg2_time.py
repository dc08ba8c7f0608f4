"""README training step (BASELINE config 2: 64 x N=5, 6 layers hidden 192) as a CUDA graph: ms per step with the fused
edge kernels (default at this size) and with the GEMM-form edge layer forced (argv[1] = gemm)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import segnn_b200 as S
import bench

if len(sys.argv) > 1 and sys.argv[1] == "gemm":
    S.ops.GEMM_FORM_MIN_ROWS = 0
torch.manual_seed(0)
dev = torch.device("cuda", 0)
B, N = 64, 5
model = S.SEGNN(hidden_features=192, num_layers=6).to(dev).train()
ts = S.TrainStep(model, B, N, use_cuda_graph=True)
pos, vel, charge = bench.synthetic_system(B, N, seed=77)
y = torch.randn(B * N, 6)
for i in range(5):
    loss = ts.step(pos, vel, charge, y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 50
e0.record()
for i in range(steps):
    loss = ts.step(pos, vel, charge, y)
e1.record()
torch.cuda.synchronize()
print(f"cfg2 training step: {e0.elapsed_time(e1) / steps:.3f} ms, loss {float(loss):.5f}, gemm form {S.ops._use_gemm_form(B, N, model.n)}")
