"""README training step (cfg2) as a CUDA graph under torch.profiler: the kernels of ONE replay + optimizer step in
launch order with start times, so gaps and the critical path show (no ncu serialisation)."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import segnn_b200 as S
import bench
from torch.profiler import profile, ProfilerActivity

torch.manual_seed(0)
dev = torch.device("cuda", 0)
B, N = 64, 5
model = S.SEGNN(hidden_features=192, num_layers=6).to(dev).train()
ts = S.TrainStep(model, B, N, use_cuda_graph=True)
pos, vel, charge = bench.synthetic_system(B, N, seed=77)
y = torch.randn(B * N, 6)
for i in range(8):
    ts.step(pos, vel, charge, y)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(3):
        ts.step(pos, vel, charge, y)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
print("cuda events:", len(ev))
# split into steps by the largest gaps
t0 = ev[0].time_range.start
third = len(ev) // 3
step = ev[third:2 * third]
span = step[-1].time_range.end - step[0].time_range.start
busy = sum(e.time_range.end - e.time_range.start for e in step)
print(f"middle step: {len(step)} events, span {span:.1f} us, sum of kernel durations {busy:.1f} us")
cnt = collections.Counter()
dur = collections.Counter()
for e in step:
    cnt[e.name[:70]] += 1
    dur[e.name[:70]] += e.time_range.end - e.time_range.start
for k, v in sorted(dur.items(), key=lambda kv: -kv[1])[:45]:
    print(f"{v:8.1f} us {cnt[k]:4d} x  {k}")
print("---- sequence (start us, dur us, stream, name)")
base = step[0].time_range.start
for e in step[:400]:
    print(f"{e.time_range.start - base:8.1f} {e.time_range.end - e.time_range.start:6.1f} {getattr(e, 'device_index', 0)} {e.name[:60]}")
