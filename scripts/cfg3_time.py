import sys, os, time
sys.path.insert(0, "/root/repo")
import torch, bench
import segnn_b200 as S
dev = torch.device("cuda", 0)
for B in (16, 64):
    N = 100
    torch.manual_seed(0)
    m = S.SEGNN(hidden_features=192, num_layers=6, lmax_h=2).to(dev).eval()
    bench.perturb_batchnorm(m)
    pos, vel, charge = bench.synthetic_system(B, N, seed=1)
    p, v, c = pos.reshape(-1, 3).to(dev), vel.reshape(-1, 3).to(dev), charge.reshape(-1).to(dev)
    with torch.no_grad():
        for _ in range(2):
            m.forward_state(p, v, c, B, N)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            m.forward_state(p, v, c, B, N)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"cfg3 lmax_h=2 B={B} N={N}: {ms:.1f} ms per step = {B*N/ms*1e3:.0f} particle-steps/s, peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GB")
