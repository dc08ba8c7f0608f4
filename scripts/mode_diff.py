"""bf16 tensor-core mode vs fp32 FFMA mode, per layer, for growing N (diagnostic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_gpu_parity as T
S, O = T.S, T.O
for H in (128, 192):
    _, m = T.make_pair(H, 2, seed=11)
    for N in (100, 250, 500, 1000):
        pos, vel, mass = O.synthetic_system(1, N, seed=13)
        p, v, ms = pos.reshape(-1, 3).float().cuda(), vel.reshape(-1, 3).float().cuda(), mass.reshape(-1).float().cuda()
        with torch.no_grad():
            m.compute_mode = "fp32"; ref, rl = m.forward_state(p, v, ms, 1, N, return_layers=True)
            m.compute_mode = "bf16"; out, ol = m.forward_state(p, v, ms, 1, N, return_layers=True)
        errs = [float((a - b).abs().max() / b.abs().max()) for a, b in zip(ol, rl)]
        mags = [float(b.abs().max()) for b in rl]
        print(f"H={H} N={N}: per-layer rel err {['%.2e' % e for e in errs]} out {float((out-ref).abs().max()/ref.abs().max()):.2e} | layer max {['%.1f' % x for x in mags]}")
