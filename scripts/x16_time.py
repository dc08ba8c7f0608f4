"""cfg5 shard rollout step (1024 x N=100, 6 layers, hidden 192, fp16p) with and without the fp16 operand copies of the
node features (ops.X16_FEATURES), CUDA-graph replay; plus the node-level kernels timed one by one."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import segnn_b200 as S
dev = torch.device("cuda", 0)
B, N, steps = int(os.environ.get("X16_B", 1024)), 100, 8
torch.manual_seed(0)
m = S.SEGNN(hidden_features=192, num_layers=6, compute_mode="fp16p").to(dev).eval()
bench.perturb_batchnorm(m)
pos, vel, charge = bench.synthetic_system(B, N, seed=1)
outs = {}
for x16 in (False, True, False, True):
    S.ops.X16_FEATURES = x16
    roll = S.SelfFeedRollout(m, B, N, dev, max_frames=steps + 4, use_cuda_graph=True)
    roll.reset(pos, vel, charge)
    roll.capture()
    roll.reset(pos, vel, charge)
    for _ in range(3):
        roll.step()
    roll.reset(pos, vel, charge)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        roll.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    outs[x16] = roll.pos.clone()
    print(f"x16={x16}: {ms:.3f} ms per step = {B * N / ms * 1e3 / 1e6:.3f} M particle-steps/s", flush=True)
print("rollout states bit-identical:", torch.equal(outs[False], outs[True]))

# node-level kernels one by one
ops = S.ops
n, nodes = 96, B * N
w = m.packed(N - 1)
lw = w["layers"][0]
x_in, attr = ops.prep(roll.pos, roll.vel, B, N)
h, h16 = ops.embed(x_in, attr, w["embed"]["w"], w["embed"]["bias"], n, want16=True)
m1, u1, u2 = lw["msg1"], lw["upd1"], lw["upd2"]
wpq = dict(wt_s=m1["wt_s_h"], wt_v=m1["wt_v_h"], operand=1)
agg = torch.randn(nodes, 4, n, device=dev)
agg16 = agg.half()


def timeit(name, fn, nbytes):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tot = 0.0
    for _ in range(5):
        flush.zero_()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    us = tot / 5 * 1e3
    print(f"{name:34s} {us:8.1f} us  {nbytes / us / 1e6:6.2f} TB/s")


rows = nodes * 4
timeit("P/Q GEMM fp32 rows", lambda: ops.node_gemm_pair16(h, wpq, 6 * n, m1["bias_tc_h"], 3 * n, 3 * n), rows * (384 + 1152))
timeit("P/Q GEMM fp16 rows", lambda: ops.node_gemm_pair16(h16, wpq, 6 * n, m1["bias_tc_h"], 3 * n, 3 * n), rows * (192 + 1152))
timeit("update_1 GEMM fp32 rows", lambda: ops.node_gemm_out16(h, agg, u1, 3 * n), rows * (768 + 576))
timeit("update_1 GEMM fp16 rows", lambda: ops.node_gemm_out16(h16, agg16, u1, 3 * n), rows * (384 + 576))
y1 = ops.node_gemm_out16(h16, agg16, u1, 3 * n)
timeit("combine gate -> fp32", lambda: ops.tp_combine(y1, attr, n, True, bias=u1["bias"]), rows * (576 + 384))
timeit("combine gate -> fp16 only", lambda: ops.tp_combine(y1, attr, n, True, bias=u1["bias"], out16="only"), rows * (576 + 192))
g1 = ops.tp_combine(y1, attr, n, True, bias=u1["bias"])
g1_16 = g1.half()
timeit("update_2 GEMM fp32 rows", lambda: ops.node_gemm_out16(g1, None, u2, 2 * n), rows * (384 + 384))
timeit("update_2 GEMM fp16 rows", lambda: ops.node_gemm_out16(g1_16, None, u2, 2 * n), rows * (192 + 384))
y2 = ops.node_gemm_out16(g1_16, None, u2, 2 * n)
timeit("combine residual+BN -> fp32", lambda: ops.tp_combine(y2, attr, n, False, bias=u2["bias"], residual=h, bn_mul=lw["bn_feat"][0], bn_add=lw["bn_feat"][1]), rows * (384 + 384 + 384))
timeit("combine residual+BN -> fp32+fp16", lambda: ops.tp_combine(y2, attr, n, False, bias=u2["bias"], residual=h, bn_mul=lw["bn_feat"][0], bn_add=lw["bn_feat"][1], out16="both"), rows * (384 + 384 + 384 + 192))
