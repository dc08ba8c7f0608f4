"""Which torch ops (and from where) run in one eager training step of the README configuration."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import segnn_b200 as S
import bench

torch.manual_seed(0)
dev = torch.device("cuda", 0)
B, N = 64, 5
model = S.SEGNN(hidden_features=192, num_layers=6).to(dev).train()
ts = S.TrainStep(model, B, N, use_cuda_graph=False)
pos, vel, charge = bench.synthetic_system(B, N, seed=77)
y = torch.randn(B * N, 6)
for _ in range(2):
    ts.step(pos, vel, charge, y)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    ts.step(pos, vel, charge, y)
    torch.cuda.synchronize()
cnt = collections.Counter()
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CPU and ev.name.startswith("aten::") and ev.cuda_time_total > 0:
        frame = next((f for f in ev.stack if "_b200/" in f or "segnn_b200" in f), "?")
        cnt[(ev.name, frame.split("/")[-1][:70])] += 1
for (name, frame), c in cnt.most_common(40):
    print(f"{c:4d} {name:28s} {frame}")
