"""Debug: edge part of one lmax_h = 2 layer, l2 GEMM form vs the table-driven generic kernels, with pieces switched off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import segnn_b200 as S
import segnn_b200.generic as G
from segnn_b200 import ops
from oracle import segnn_oracle as O

torch.manual_seed(0)
H, B, N = 32, 2, 6
m = S.SEGNN(hidden_features=H, num_layers=1, lmax_h=2, compute_mode="generic").float().cuda().eval()
pos, vel, mass = O.synthetic_system(B, N, seed=6)
p, v, c = pos.reshape(-1, 3).float().cuda(), vel.reshape(-1, 3).float().cuda(), mass.reshape(-1).float().cuda()
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())


def both(tag):
    r = G.GenericRunner(m, p.device)
    layer, pl = m.layers[0], r.layers[0]
    x_in, attr = ops.prep(p, v, B, N)
    x = r.embed.run(x_in, attr)
    ea, add = ops.edge_attr(p, c, B, N)
    msg = pl["g_msg2"].run(pl["msg2"].run(pl["g_msg1"].run(pl["msg1h"].run(x, ea, add, B, N)), ea))
    agg = torch.empty((B * N, msg.shape[1]), device=p.device)
    G.check(G.lib.segnn_generic_aggregate(G._p(msg), B, N, msg.shape[1], G._p(agg), ops._stream()), "agg")
    if layer.message_norm is not None:
        mul, addc = G._bn_eval_columns(layer.message_norm, layer.hidden_irreps)
        agg = ops.lincomb(agg, None, mul, None, addc * float(N - 1))
    got = pl["l2"].run(x, p, c, B, N)
    n = pl["l2"].n
    print(tag, "total", f"{rel(got, agg):.2e}", "0e", f"{rel(got[:, :n], agg[:, :n]):.2e}", "1o",
          f"{rel(got[:, n:4*n], agg[:, n:4*n]):.2e}", "2e", f"{rel(got[:, 4*n:], agg[:, 4*n:]):.2e}")


both("full")
L = m.layers[0]
norm = L.message_norm
L.message_norm = None
both("no message_norm")
with torch.no_grad():
    L.message_layer_1.biases.zero_(); L.message_layer_2.biases.zero_()
both("+ no biases")
with torch.no_grad():
    for ins in L.message_layer_1.tp.instructions:
        if ins["i1"] == 6:
            k = ins["shape"][0] * ins["shape"][2]
            L.message_layer_1.tp.weight[ins["offset"]: ins["offset"] + k].zero_()
both("+ no additional scalars")
