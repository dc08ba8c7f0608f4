"""A few eager self-feed steps of BASELINE config 1 (100 x 5 bodies, 4 layers, hidden 64) for an ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import segnn_b200 as S

dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = S.SEGNN(hidden_features=64, num_layers=4, compute_mode=sys.argv[1] if len(sys.argv) > 1 else "bf16").to(dev).eval()
bench.perturb_batchnorm(m)
pos, vel, charge = bench.synthetic_system(100, 5, seed=1)
r = S.SelfFeedRollout(m, 100, 5, dev, max_frames=16, use_cuda_graph=False)
r.reset(pos, vel, charge)
r.capture()
for _ in range(8):
    r.step()
torch.cuda.synchronize()
