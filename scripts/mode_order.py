import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_gpu_parity as T
S, O = T.S, T.O
for order in (["bf16"], ["fp16"], ["fp32", "bf16"], ["bf16", "fp32", "bf16", "fp16", "bf16"]):
    om, m = T.make_pair(64, 2, seed=3)
    B, N = 4, 5
    pos, vel, mass = O.synthetic_system(B, N, seed=5)
    with torch.no_grad():
        ref = om(O.make_graph(pos.reshape(-1, 3), vel.reshape(-1, 3), mass.reshape(-1, 1), B, N))
        errs = []
        for mode in order:
            m.compute_mode = mode
            out = m(T.gpu_graph(pos, vel, mass, B, N))
            errs.append(f"{mode}:{T.rel(out, ref):.2e}")
    print(order, errs)
