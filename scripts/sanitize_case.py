"""Small ragged shapes through every kernel family, for compute-sanitizer (one --tool per gpurun call):

    compute-sanitizer --tool memcheck  python scripts/sanitize_case.py
    compute-sanitizer --tool racecheck python scripts/sanitize_case.py

Shapes are chosen so that tiles are partially filled: N in {5, 6, 7, 10} against 4 receivers x 8 senders per tile,
node counts that are not multiples of the 128-row GEMM tile, an odd and an even number of graphs."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import segnn_b200 as S  # noqa: E402
from oracle import segnn_oracle as O  # noqa: E402

torch.manual_seed(0)
worst = 0.0
for H, L, B, N in [(64, 1, 3, 6), (64, 1, 2, 7), (128, 1, 1, 10), (192, 1, 2, 6), (64, 1, 5, 5)]:
    om = O.SEGNN(hidden_features=H, num_layers=L).eval()
    O.perturb_bn_buffers(om)
    m = S.SEGNN(hidden_features=H, num_layers=L)
    m.load_state_dict(om.state_dict())
    m = m.float().cuda().eval()
    pos, vel, mass = O.synthetic_system(B, N, seed=H + N)
    with torch.no_grad():
        ref = om(O.make_graph(pos.reshape(-1, 3), vel.reshape(-1, 3), mass.reshape(-1, 1), B, N))
        modes = ["fp32", "bf16", "fp16"] + (["fp16p"] if N % 2 == 0 else [])
        for mode in modes:
            m.compute_mode = mode
            g = S.GraphBatch(pos=pos.reshape(-1, 3).float().cuda(), vel=vel.reshape(-1, 3).float().cuda(),
                             mass=mass.reshape(-1, 1).float().cuda(), num_graphs=B, n_nodes=N)
            out = m(g)
            torch.cuda.synchronize()
            err = float((out.double().cpu() - ref).abs().max() / ref.abs().max())
            worst = max(worst, err)
            print(f"H={H} B={B} N={N} {mode}: rel err {err:.2e}", flush=True)
# backward kernels (fp32) on a ragged shape
om = O.SEGNN(hidden_features=64, num_layers=1).train()
m = S.SEGNN(hidden_features=64, num_layers=1)
m.load_state_dict(om.state_dict())
m = m.float().cuda().train()
pos, vel, mass = O.synthetic_system(3, 7, seed=2)
g = S.GraphBatch(pos=pos.reshape(-1, 3).float().cuda(), vel=vel.reshape(-1, 3).float().cuda(),
                 mass=mass.reshape(-1, 1).float().cuda(), num_graphs=3, n_nodes=7)
S.target_common_loss(m(g), torch.randn(21, 6).cuda()).backward()
torch.cuda.synchronize()
# rollout (integrate, counter), macros, simulators
m.eval()
roll = S.SelfFeedRollout(m, 3, 7, "cuda", max_frames=4, use_cuda_graph=False)
roll.reset(pos, vel, mass)
tp, tv = roll.run(3)
S.macros.event_counters(tp, tv, 3, 7)
S.macros.group_collisions(tp, 3, 7)
S.macros.energy_momentum(tp, tv, 3, 7, 2.0, 0.2)
S.simulator.GravitySim(n_balls=7, interaction_strength=2.0, dt=0.01, softening=0.2).sample_trajectories(
    2, T=20, sample_freq=5)
torch.cuda.synchronize()
print(f"sanitize_case done, worst rel err {worst:.2e}")
assert worst < 3e-2
