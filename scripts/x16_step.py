"""One eager cfg5-shard rollout step (1024 x N=100, fp16p) for an ncu launch list; X16=0/1 selects the feature path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import segnn_b200 as S
dev = torch.device("cuda", 0)
B, N = 1024, 100
S.ops.X16_FEATURES = os.environ.get("X16", "1") == "1"
torch.manual_seed(0)
m = S.SEGNN(hidden_features=192, num_layers=6, compute_mode="fp16p").to(dev).eval()
bench.perturb_batchnorm(m)
pos, vel, charge = bench.synthetic_system(B, N, seed=1)
roll = S.SelfFeedRollout(m, B, N, dev, max_frames=8, use_cuda_graph=False)
roll.reset(pos, vel, charge)
for _ in range(3):
    roll.step()
torch.cuda.synchronize()
print("done")
