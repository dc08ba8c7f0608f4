import sys
sys.path.insert(0, "/root/repo")
import torch, bench
import segnn_b200 as S
dev = torch.device("cuda", 0)
B, N = 16, 100
torch.manual_seed(0)
m = S.SEGNN(hidden_features=192, num_layers=6, lmax_h=2).to(dev).eval()
bench.perturb_batchnorm(m)
pos, vel, charge = bench.synthetic_system(B, N, seed=1)
p, v, c = pos.reshape(-1, 3).to(dev), vel.reshape(-1, 3).to(dev), charge.reshape(-1).to(dev)
with torch.no_grad():
    for _ in range(2):
        m.forward_state(p, v, c, B, N)
torch.cuda.synchronize()
