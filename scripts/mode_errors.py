"""Error of every compute mode against the float64 oracle on a few shapes (even N so that 'fp16p' applies)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_gpu_parity as T
S, O = T.S, T.O
for H, L, B, N in ((64, 2, 4, 6), (192, 6, 2, 6), (192, 6, 2, 100), (128, 3, 2, 38)):
    om, m = T.make_pair(H, L, seed=3)
    pos, vel, mass = O.synthetic_system(B, N, seed=5)
    with torch.no_grad():
        ref, ref_layers = om(O.make_graph(pos.reshape(-1, 3), vel.reshape(-1, 3), mass.reshape(-1, 1), B, N)), None
        out = {}
        for mode in ("fp32", "bf16", "fp16", "fp16p"):
            m.compute_mode = mode
            out[mode] = T.rel(m(T.gpu_graph(pos, vel, mass, B, N)), ref)
    print(f"H={H} L={L} B={B} N={N}: " + "  ".join(f"{k} {v:.2e}" for k, v in out.items()), flush=True)
