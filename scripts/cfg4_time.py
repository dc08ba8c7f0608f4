"""BASELINE config 4 (one N=1000 graph, hidden 128, 6 layers) forward + backward: ms per step with the GEMM-form edge
layer (default from ops.GEMM_FORM_MIN_ROWS rows on) and, with `fused` as argv[1], with the fused fp32 kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import segnn_b200 as S
import bench

if len(sys.argv) > 1 and sys.argv[1] == "fused":
    S.ops.GEMM_FORM_MIN_ROWS = 1 << 62
torch.manual_seed(0)
dev = torch.device("cuda", 0)
N4, H4 = 1000, 128
m4 = S.SEGNN(hidden_features=H4, num_layers=6, lmax_h=1).to(dev).train()
p4, v4, c4 = bench.synthetic_system(1, N4, seed=5)
g = S.GraphBatch(pos=p4.reshape(-1, 3).to(dev), vel=v4.reshape(-1, 3).to(dev), mass=c4.reshape(-1, 1).to(dev),
                 num_graphs=1, n_nodes=N4)
y4 = torch.randn(N4, 6, device=dev)


def step():
    m4.zero_grad(set_to_none=True)
    loss = S.target_common_loss(m4(g), y4)
    loss.backward()
    return loss


for _ in range(3):
    loss = step()
torch.cuda.synchronize()
torch.cuda.reset_peak_memory_stats()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    loss = step()
e1.record()
torch.cuda.synchronize()
print(f"cfg4 fwd+bwd: {e0.elapsed_time(e1) / reps:.2f} ms per step, loss {float(loss):.6f}, "
      f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.2f} GB, gemm form: {S.ops._use_gemm_form(1, N4, m4.n)}")
# forward only
with torch.no_grad():
    m4(g)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        m4(g)
    e1.record()
    torch.cuda.synchronize()
print(f"cfg4 train-mode forward only (no grad): {e0.elapsed_time(e1) / reps:.2f} ms")
