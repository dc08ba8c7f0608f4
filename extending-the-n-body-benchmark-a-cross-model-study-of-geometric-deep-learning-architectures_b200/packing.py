"""Re-layout of the reference's flat e3nn ``tp.weight`` parameters into the kernel operand blocks.

The flat parameter is the concatenation of the [mul1, 1, mul_out] instruction views in e3nn's instruction order
(models/segnn/o3_building_blocks.py:86,96). Everything here is plain differentiable torch
indexing, so weight gradients computed in kernel layout flow back to the reference-named parameters.

Net coupling constants (e3nn 'component' path weight x the reference's sqrt_k_correction, o3_building_blocks.py:150-162):
(l,0,l) and (0,l,l) paths are identities, (1,1,0) is a dot product / sqrt(3). For edge tensor products the constant
Y_0 of the edge attribute is folded into the weights that multiply it.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from .irreps import Irreps, tp_instructions

Y0 = 0.28209479177387814
INV_SQRT3 = 1.0 / math.sqrt(3.0)
ATTR = Irreps("1x0e+1x1o")


def _views(flat: torch.Tensor, irreps_in1: Irreps, irreps_out: Irreps) -> Dict[tuple, torch.Tensor]:
    instr, numel = tp_instructions(irreps_in1, ATTR, irreps_out)
    if flat.numel() != numel:
        raise ValueError(f"tp.weight has {flat.numel()} elements, irreps need {numel}")
    out = {}
    for ins in instr:
        m1, m2, mo = ins["shape"]
        out[(ins["i1"], ins["i2"], ins["io"])] = flat[ins["offset"]: ins["offset"] + m1 * m2 * mo].view(m1, mo)
    return out


def hidden_irreps(n: int) -> Irreps:
    return Irreps([(n, 0, 1), (n, 1, -1)])


def _hidden_blocks(flat, n_blocks: int, n: int, n0: int, extra_scalars: int = 0):
    """Views of a TP whose in1 is n_blocks copies of (n x0e + n x1o) [+ extra x0e], out = n0 x0e + n x1o.
    Returns per block (ss [n,n0], sv [n,n], vv [n,n], vs [n,n0]) and (ss_add [extra,n0], sv_add [extra,n])."""
    in1 = Irreps([(n, 0, 1), (n, 1, -1)] * n_blocks + ([(extra_scalars, 0, 1)] if extra_scalars else []))
    out = Irreps([(n0, 0, 1), (n, 1, -1)])
    v = _views(flat, in1, out)
    blocks = []
    for b in range(n_blocks):
        blocks.append((v[(2 * b, 0, 0)], v[(2 * b, 1, 1)], v[(2 * b + 1, 0, 1)], v[(2 * b + 1, 1, 0)]))
    add = (v[(2 * n_blocks, 0, 0)], v[(2 * n_blocks, 1, 1)]) if extra_scalars else None
    return blocks, add


def pack_msg1(flat: torch.Tensor, biases: torch.Tensor, n: int):
    """message_layer_1 (segnn.py:212-214; in1 = x_i | x_j | (dist, m_i m_j)) -> node-GEMM weights producing the
    hoisted receiver/sender projections pq [.., 6n] = (P0 2n, P1 n, Q0 2n, Q1 n) and the per-edge scalar terms."""
    (ss_i, sv_i, vv_i, vs_i), (ss_j, sv_j, vv_j, vs_j) = _hidden_blocks(flat, 2, n, 2 * n, 2)[0]
    ss_a, sv_a = _hidden_blocks(flat, 2, n, 2 * n, 2)[1]
    w_s = torch.cat([Y0 * ss_i, sv_i, Y0 * ss_j, sv_j], dim=1)
    w_v = torch.cat([INV_SQRT3 * vs_i, Y0 * vv_i, INV_SQRT3 * vs_j, Y0 * vv_j], dim=1)
    w_edge = torch.cat([Y0 * ss_a[0], Y0 * ss_a[1], sv_a[0], sv_a[1]])
    return dict(w_s=w_s.contiguous(), w_v=w_v.contiguous(), bias=biases.contiguous(), w_edge=w_edge.contiguous())


def pack_msg2(flat: torch.Tensor, biases: torch.Tensor, n: int):
    """message_layer_2 (segnn.py:215-217) blocks for the fused edge kernel (Y_0 and 1/sqrt3 folded)."""
    (ss, sv, vv, vs), = _hidden_blocks(flat, 1, n, 2 * n)[0]
    return dict(ss=(Y0 * ss).contiguous(), vs=(INV_SQRT3 * vs).contiguous(), sv=sv.contiguous(),
                vv=(Y0 * vv).contiguous(), b=biases.contiguous())


def pack_node_tp(flat: torch.Tensor, biases: torch.Tensor, n_blocks: int, n: int, n0: int):
    """Node-level TP (update_layer_1/2, pre_pool1): GEMM weights w_s/w_v [n_blocks*n, n0+n] + bias for the combine."""
    blocks, _ = _hidden_blocks(flat, n_blocks, n, n0)
    w_s = torch.cat([torch.cat([ss, sv], dim=1) for ss, sv, vv, vs in blocks], dim=0)
    w_v = torch.cat([torch.cat([INV_SQRT3 * vs, vv], dim=1) for ss, sv, vv, vs in blocks], dim=0)
    return dict(w_s=w_s.contiguous(), w_v=w_v.contiguous(), bias=biases.contiguous())


def pack_embedding(flat: torch.Tensor, biases: torch.Tensor, n: int):
    """embedding_layer (segnn.py:63-65): in1 = 2x1o + 1x0e -> w_embed [6][n]."""
    v = _views(flat, Irreps("2x1o+1x0e"), hidden_irreps(n))
    vec_v, vec_s, sc_s, sc_v = v[(0, 0, 1)], v[(0, 1, 0)], v[(1, 0, 0)], v[(1, 1, 1)]
    w = torch.stack([vec_v[0], vec_v[1], INV_SQRT3 * vec_s[0], INV_SQRT3 * vec_s[1], sc_s[0], sc_v[0]])
    return dict(w=w.contiguous(), bias=biases.contiguous())


def pack_head(flat: torch.Tensor, n: int):
    """pre_pool2 (segnn.py:104-106): h -> 2x1o, no bias. w_head [2][n][2]."""
    v = _views(flat, hidden_irreps(n), Irreps("2x1o"))
    return torch.stack([v[(0, 1, 0)], v[(1, 0, 0)]]).contiguous()


def fold_batchnorm(weight, bias, running_mean, running_var, n: int, eps: float = 1e-5, degree: float = 1.0):
    """Eval-mode e3nn BatchNorm (segnn.py:233-235) as out_s*mul[:n]+add, out_v*mul[n:]. With degree = N-1 it is the
    BatchNorm of every message folded through the sum over the N-1 senders (affine commutes with the sum)."""
    mul = weight * (running_var + eps).rsqrt()
    add = degree * (bias - running_mean * mul[:n])
    return mul.contiguous(), add.contiguous()


def to_planar(x: torch.Tensor, n: int) -> torch.Tensor:
    """e3nn mul-major [rows, 4n] = [s | v(u,k)] -> planar [rows, 4, n] = (s, vx, vy, vz)."""
    rows = x.shape[0]
    return torch.cat([x[:, :n].reshape(rows, 1, n), x[:, n:].reshape(rows, n, 3).permute(0, 2, 1)], dim=1).contiguous()


def from_planar(x: torch.Tensor) -> torch.Tensor:
    """planar [rows, 4, n] -> e3nn mul-major [rows, 4n]."""
    rows, _, n = x.shape
    return torch.cat([x[:, 0, :], x[:, 1:, :].permute(0, 2, 1).reshape(rows, 3 * n)], dim=1).contiguous()
