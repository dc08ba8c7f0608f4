"""Timeline of the K3 roles for the first tiles of CTA 0 (needs the trace build: SEGNN_NVCC_EXTRA=-DSEGNN_K3_TRACE
csrc/build.sh). Prints, per tile, the clock offsets of every role's events."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, ROOT)
import torch
import segnn_b200 as S
import bench

lib = ctypes.CDLL(S._lib.LIB_PATH)
dev = torch.device("cuda", 0)
buf = torch.zeros(16 * 64 * 8, dtype=torch.int64, device=dev)
B, N = 1024, 100
torch.manual_seed(0)
model = S.SEGNN(hidden_features=192, num_layers=1, compute_mode="bf16").to(dev).eval()
pos, vel, charge = bench.synthetic_system(B, N, seed=1)
p, v, m = pos.reshape(-1, 3).to(dev), vel.reshape(-1, 3).to(dev), charge.reshape(-1).to(dev)
with torch.no_grad():
    model.forward_state(p, v, m, B, N)
    torch.cuda.synchronize()
    assert lib.segnn_debug_set_k3_trace(ctypes.c_void_p(buf.data_ptr())) == 0
    model.forward_state(p, v, m, B, N)
    torch.cuda.synchronize()
t = buf.cpu().reshape(16, 64, 8)
t0 = int(t[t > 0].min())
for tile in range(20, 34):
    print(f"--- tile {tile}")
    for wp in (0, 5, 10, 7, 11, 3):
        ev = [int(x) - t0 if x > 0 else -1 for x in t[wp, tile]]
        print(f"  warp {wp:2d}: {ev[:7]}")
