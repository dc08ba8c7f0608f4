"""Runs a few eager training steps of the README configuration (for an ncu launch list)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import segnn_b200 as S
import bench

torch.manual_seed(0)
dev = torch.device("cuda", 0)
B, N = 64, 5
model = S.SEGNN(hidden_features=192, num_layers=6).to(dev).train()
ts = S.TrainStep(model, B, N, use_cuda_graph=False)
pos, vel, charge = bench.synthetic_system(B, N, seed=77)
y = torch.randn(B * N, 6)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    print(i, float(ts.step(pos, vel, charge, y)))
