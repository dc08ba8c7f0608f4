"""BASELINE config 1 (SEGNN 4 layers hidden 64, 5 bodies, batch 100, 100-step self-feed rollout) on the GPU path:
CUDA-graph rollout, every compute mode that applies (N = 5 is odd: no packed-half mode)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import segnn_b200 as S
dev = torch.device("cuda", 0)
B, N, steps = 100, 5, 100
torch.manual_seed(0)
m = S.SEGNN(hidden_features=64, num_layers=4).to(dev).eval()
bench.perturb_batchnorm(m)
pos, vel, charge = bench.synthetic_system(B, N, seed=1)
for mode in ("fp32", "bf16", "fp16"):
    m.compute_mode = mode
    roll = S.SelfFeedRollout(m, B, N, dev, max_frames=steps + 1, use_cuda_graph=True)
    roll.reset(pos, vel, charge)
    roll.capture()
    roll.reset(pos, vel, charge)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    roll.run(steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"cfg1 [{mode}] 100-step rollout of 100 x 5 bodies: {ms:.1f} ms = {ms / steps * 1e3:.0f} us per step, "
          f"{B * N * steps / ms * 1e3:.0f} particle-steps/s")
