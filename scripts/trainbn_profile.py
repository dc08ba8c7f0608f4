"""Two self-feed steps of the cfg5 shard with train-mode (batch-statistic) BatchNorm, eager, for an ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import segnn_b200 as S

dev = torch.device("cuda", 0)
torch.manual_seed(0)
B, N = 1024, 100
m = S.SEGNN(hidden_features=192, num_layers=6, compute_mode="fp16p").to(dev).train()
pos, vel, charge = bench.synthetic_system(B, N, seed=1)
r = S.SelfFeedRollout(m, B, N, dev, max_frames=8, use_cuda_graph=False, allow_train_mode=True)
r.reset(pos, vel, charge)
r.capture()
for _ in range(3):
    r.step()
torch.cuda.synchronize()
print("ok")
