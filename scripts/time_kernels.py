"""Per-kernel timing of one SEGNN layer on the bench workload (cfg5 shard: 1024 x N=100, hidden 192) with CUDA
events: node GEMMs (P/Q, update 1, update 2), the two attribute-combine passes and the fused edge kernel, each with
its HBM figure. Usage: python scripts/time_kernels.py [mode] [sims]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
import segnn_b200 as S
from segnn_b200 import ops

mode_name = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
N = 100
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = S.SEGNN(hidden_features=192, num_layers=1, compute_mode=mode_name).to(dev).eval()
bench.perturb_batchnorm(model)
pos, vel, charge = bench.synthetic_system(B, N, seed=1)
p, v, m = pos.reshape(-1, 3).to(dev), vel.reshape(-1, 3).to(dev), charge.reshape(-1).to(dev)
n = model.n
nodes = B * N
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=10):
    fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


with torch.no_grad():
    w = model.packed(N - 1)
    lw = w["layers"][0]
    mode = S.segnn._MODES[mode_name]
    tc = mode in S.segnn._TC_MODES
    x_in, node_attr = ops.prep(p, v, B, N)
    h = ops.embed(x_in, node_attr, w["embed"]["w"], w["embed"]["bias"], n)
    m1, u1, u2 = lw["msg1"], lw["upd1"], lw["upd2"]
    pq_bias, pq_nb = (m1["bias_tc"], 3 * n) if tc else (m1["bias"], 2 * n)
    packed = mode == ops.MODE_FP16_PACKED
    m1h = dict(wt_s=m1.get("wt_s_h"), wt_v=m1.get("wt_v_h"), operand=1)
    if packed:
        pq_fn = lambda: ops.node_gemm_pair16(h, m1h, 6 * n, m1["bias_tc_h"], 3 * n, 3 * n)
    else:
        pq_fn = lambda: ops.node_gemm(h, None, m1, 6 * n, bias=pq_bias, n_bias=pq_nb, split=3 * n, tc=tc)
    pq = pq_fn()
    agg = ops.edge_layer(mode, p, m, B, N, n, pq[0], pq[1], m1["w_edge"], lw["msg2"], lw["bn_msg"][0], lw["bn_msg"][1])
    g16 = (lambda a, b, w_, no: ops.node_gemm_out16(a, b, w_, no)) if tc else (lambda a, b, w_, no: ops.node_gemm(a, b, w_, no))
    ob = 2 if tc else 4  # bytes per element of the node-GEMM outputs that feed the combine passes
    y1 = g16(h, agg, u1, 3 * n)
    g1 = ops.tp_combine(y1, node_attr, n, True, bias=u1["bias"])
    y2 = g16(g1, None, u2, 2 * n)
    rows = nodes * 4
    cases = [
        ("node_gemm P/Q  (K=n,  out 6n)", pq_fn, rows * 4 * n + rows * 6 * n * (2 if packed else 4)),
        ("edge_layer K3", lambda: ops.edge_layer(mode, p, m, B, N, n, pq[0], pq[1], m1["w_edge"], lw["msg2"], lw["bn_msg"][0],
                                                 lw["bn_msg"][1]), rows * 6 * n * (2 if packed else 4) + rows * 4 * n),
        ("node_gemm upd1 (K=2n, out 3n)", lambda: g16(h, agg, u1, 3 * n), rows * (4 * 2 * n + ob * 3 * n)),
        ("tp_combine gate", lambda: ops.tp_combine(y1, node_attr, n, True, bias=u1["bias"]), rows * (ob * 3 * n + 4 * n)),
        ("node_gemm upd2 (K=n,  out 2n)", lambda: g16(g1, None, u2, 2 * n), rows * (4 * n + ob * 2 * n)),
        ("tp_combine residual+BN", lambda: ops.tp_combine(y2, node_attr, n, False, bias=u2["bias"], residual=h,
                                                          bn_mul=lw["bn_feat"][0], bn_add=lw["bn_feat"][1]),
         rows * (ob * 2 * n + 4 * n + 4 * n)),
    ]
    total = 0.0
    for name, fn, nbytes in cases:
        ms = timed(fn)
        total += ms
        print(f"{name:34s} {ms * 1e3:8.1f} us   {nbytes / ms / 1e6:7.0f} GB/s algorithmic ({nbytes / 1e6:.0f} MB)")
    print(f"layer total {total:.3f} ms  (mode {mode_name}, {B} sims x N={N})")
