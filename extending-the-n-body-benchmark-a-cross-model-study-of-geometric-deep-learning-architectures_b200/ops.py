"""Tensor-level wrappers over the C ABI: validate, allocate outputs with torch, launch on torch's current
stream. PyTorch is plumbing here (device memory + streams); all arithmetic happens in libsegnn_b200.so."""
from __future__ import annotations

import ctypes
from typing import Optional

import os

import torch

from ._lib import (MODE_BF16_TC, MODE_FP16_PACKED, MODE_FP16_TC, MODE_FP32, OPERAND_BF16, OPERAND_FP16, check,
                   lib)

__all__ = ["edge_index", "edge_attr", "prep", "embed", "node_gemm", "node_gemm_out16", "node_gemm_pair16", "pack_node_weight_tc", "tp_combine", "edge_layer", "head",
           "integrate", "counter_add", "launch_count", "gemm_tf32x3", "MODE_FP32", "MODE_BF16_TC", "MODE_FP16_TC", "MODE_FP16_PACKED"]

_launches = 0  # kernels launched through this module (bench.py reports it as gpu_launches)


def launch_count() -> int:
    return _launches


# e3nn normalize2mom constants of SiLU / sigmoid folded into the gate kernels (csrc/segnn_common.cuh)
C_SILU, C_SIG = 1.6791767923989418, 1.8467055342154763


def _bump(k: int = 1) -> None:
    global _launches
    _launches += k


def _p(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the SEGNN hot path has no CPU fallback")
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    return t.contiguous()


def edge_index(batch_size: int, num_nodes: int, device) -> torch.Tensor:
    E = batch_size * num_nodes * (num_nodes - 1)
    out = torch.empty((2, E), dtype=torch.int64, device=device)
    if not out.is_cuda:
        raise RuntimeError("edge_index needs a CUDA device")
    with torch.cuda.device(out.device):
        check(lib.segnn_edge_index(batch_size, num_nodes, _p(out), _stream()), "segnn_edge_index")
    _bump()
    return out


def knn_edge_index(loc: torch.Tensor, batch_size: int, num_nodes: int, num_neighbors: int, device) -> torch.Tensor:
    """utils/build_fully_connected_graph.py:42-80 on the device (float64 distances)."""
    if loc is None:
        raise ValueError("a kNN graph needs the node positions")
    loc = torch.as_tensor(loc).to(device=device, dtype=torch.float64).reshape(batch_size * num_nodes, -1).contiguous()
    if not loc.is_cuda:
        raise RuntimeError("knn_edge_index needs a CUDA device")
    out = torch.empty((2, batch_size * num_nodes * num_neighbors), dtype=torch.int64, device=loc.device)
    with torch.cuda.device(loc.device):
        check(lib.segnn_knn_edge_index(_p(loc), batch_size, num_nodes, loc.shape[1], num_neighbors, _p(out), _stream()),
              "segnn_knn_edge_index")
    _bump()
    return out


def edge_list_csr(edge_index: torch.Tensor, nodes: int):
    """(order, ptr) of the edges grouped by target (edge_index[1]) in their original order: index plumbing for
    segnn_segment_reduce; stable sort + searchsorted, no host synchronisation (CUDA-graph capturable)."""
    dst = edge_index[1].contiguous()
    sorted_dst, order = torch.sort(dst, stable=True)
    ptr = torch.searchsorted(sorted_dst, torch.arange(nodes + 1, dtype=torch.int64, device=dst.device))
    return order.contiguous(), ptr.contiguous()


def edge_attr_list(pos: torch.Tensor, mass: torch.Tensor, edge_index: torch.Tensor, lmax_attr: int = 1):
    """O3Transform's edge part (o3_building_blocks.py:237-245,277) on an explicit edge list."""
    pos, mass = _f32(pos, "pos"), _f32(mass, "mass").reshape(-1)
    assert edge_index.dtype == torch.int64 and edge_index.dim() == 2 and edge_index.shape[0] == 2
    edge_index = edge_index.contiguous()
    E = edge_index.shape[1]
    ea = torch.empty((E, (lmax_attr + 1) ** 2), dtype=torch.float32, device=pos.device)
    add = torch.empty((E, 2), dtype=torch.float32, device=pos.device)
    with torch.cuda.device(pos.device):
        check(lib.segnn_edge_attr_list(_p(pos), _p(mass), _p(edge_index), E, int(lmax_attr), _p(ea), _p(add), _stream()),
              "segnn_edge_attr_list")
    _bump()
    return ea, add


def segment_reduce(values: torch.Tensor, order: torch.Tensor, ptr: torch.Tensor, mean: bool = False):
    """out[node] = sum / mean of values[order[ptr[node]:ptr[node + 1]]] in that order (deterministic scatter)."""
    values = _f32(values, "values")
    nodes, D = ptr.numel() - 1, values.shape[1]
    out = torch.empty((nodes, D), dtype=torch.float32, device=values.device)
    with torch.cuda.device(values.device):
        check(lib.segnn_segment_reduce(_p(values), _p(order), _p(ptr), nodes, D, int(bool(mean)), _p(out), _stream()),
              "segnn_segment_reduce")
    _bump()
    return out


def prep_list(pos: torch.Tensor, vel: torch.Tensor, edge_attr: torch.Tensor, order: torch.Tensor, ptr: torch.Tensor,
              lmax_attr: int = 1):
    """x, node_attr of O3Transform (o3_building_blocks.py:253-276) for an explicit edge list: scatter-mean of the
    incoming edge attributes + harmonics of the velocity, l = 0 slot 1 (segnn.py:148)."""
    pos, vel = _f32(pos, "pos"), _f32(vel, "vel")
    nodes = pos.shape[0]
    mean_attr = segment_reduce(edge_attr, order, ptr, mean=True)
    x_in = torch.empty((nodes, 7), dtype=torch.float32, device=pos.device)
    attr = torch.empty((nodes, (lmax_attr + 1) ** 2), dtype=torch.float32, device=pos.device)
    with torch.cuda.device(pos.device):
        check(lib.segnn_prep_fwd_list(_p(pos), _p(vel), _p(mean_attr), nodes, int(lmax_attr), _p(x_in), _p(attr),
                                      _stream()), "segnn_prep_fwd_list")
    _bump()
    return x_in, attr


def add_vector_harmonics(node_attr: torch.Tensor, vec: torch.Tensor, lmax_attr: int = 1):
    """node_attr + Y_0..lmax(vec) with the l = 0 slot kept at 1 (o3_building_blocks.py:267-271, use_force_input: the
    'integral' harmonics of a per-node vector added to the node attributes)."""
    node_attr, vec = _f32(node_attr, "node_attr"), _f32(vec, "vec")
    nodes = node_attr.shape[0]
    assert vec.shape == (nodes, 3) and node_attr.shape[1] == (lmax_attr + 1) ** 2
    scratch = torch.empty((nodes, 7), dtype=torch.float32, device=vec.device)
    out = torch.empty_like(node_attr)
    with torch.cuda.device(vec.device):
        check(lib.segnn_prep_fwd_list(_p(vec), _p(vec), _p(node_attr), nodes, int(lmax_attr), _p(scratch), _p(out),
                                      _stream()), "segnn_prep_fwd_list")
    _bump()
    return out


def instance_norm(x: torch.Tensor, graph_ptr: torch.Tensor, blocks: torch.Tensor, weight, bias, eps: float):
    """models/segnn/instance_norm.py:53-129; x [rows, dim] fp32, graph_ptr int64 [graphs + 1], blocks int32 [nb, 6]."""
    x = _f32(x, "x")
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(lib.segnn_instance_norm(_p(x), _p(graph_ptr), graph_ptr.numel() - 1, x.shape[1], _p(blocks),
                                      blocks.shape[0], _p(weight), _p(bias), float(eps), _p(out), _stream()),
              "segnn_instance_norm")
    _bump()
    return out


def edge_attr(pos: torch.Tensor, mass: torch.Tensor, batch_size: int, num_nodes: int, lmax_attr: int = 1):
    pos, mass = _f32(pos, "pos"), _f32(mass, "mass").reshape(-1)
    E = batch_size * num_nodes * (num_nodes - 1)
    ea = torch.empty((E, (lmax_attr + 1) ** 2), dtype=torch.float32, device=pos.device)
    add = torch.empty((E, 2), dtype=torch.float32, device=pos.device)
    with torch.cuda.device(pos.device):
        if lmax_attr == 1:
            check(lib.segnn_edge_attr(_p(pos), _p(mass), batch_size, num_nodes, _p(ea), _p(add), _stream()),
                  "segnn_edge_attr")
        else:
            check(lib.segnn_edge_attr_lmax(_p(pos), _p(mass), batch_size, num_nodes, int(lmax_attr), _p(ea), _p(add),
                                           _stream()), "segnn_edge_attr_lmax")
    _bump()
    return ea, add


def prep(pos: torch.Tensor, vel: torch.Tensor, batch_size: int, num_nodes: int, lmax_attr: int = 1):
    pos, vel = _f32(pos, "pos"), _f32(vel, "vel")
    nodes = batch_size * num_nodes
    assert pos.shape == (nodes, 3) and vel.shape == (nodes, 3), (pos.shape, vel.shape, nodes)
    x_in = torch.empty((nodes, 7), dtype=torch.float32, device=pos.device)
    attr = torch.empty((nodes, (lmax_attr + 1) ** 2), dtype=torch.float32, device=pos.device)
    with torch.cuda.device(pos.device):
        if lmax_attr == 1:
            check(lib.segnn_prep_fwd(_p(pos), _p(vel), batch_size, num_nodes, _p(x_in), _p(attr), _stream()),
                  "segnn_prep_fwd")
        else:
            check(lib.segnn_prep_fwd_lmax(_p(pos), _p(vel), batch_size, num_nodes, int(lmax_attr), _p(x_in), _p(attr),
                                          _stream()), "segnn_prep_fwd_lmax")
    _bump()
    return x_in, attr


# fp16 operand copies of the node features next to the fp32 ones (compute_mode 'fp16p', tensor-core node GEMMs): the
# producers store the rounding the GEMM loader would apply, the GEMMs read half the bytes (segnn_node_gemm_tc_x16).
# False (or SEGNN_X16_FEATURES=0): the GEMMs convert fp32 features on load; results are bit-identical either way.
X16_FEATURES = os.environ.get("SEGNN_X16_FEATURES", "1") != "0"


def embed(x_in, node_attr, w_embed, bias, n: int, want16: bool = False):
    """want16: returns (h, h16) with h16 the fp16 copy of h (segnn_embed_fwd_x16)."""
    nodes = x_in.shape[0]
    h = torch.empty((nodes, 4, n), dtype=torch.float32, device=x_in.device)
    if want16:
        h16 = torch.empty((nodes, 4, n), dtype=torch.float16, device=x_in.device)
        with torch.cuda.device(x_in.device):
            check(lib.segnn_embed_fwd_x16(_p(x_in), _p(node_attr), _p(w_embed), _p(bias), nodes, n, _p(h), _p(h16),
                                          _stream()), "segnn_embed_fwd_x16")
        _bump()
        return h, h16
    with torch.cuda.device(x_in.device):
        check(lib.segnn_embed_fwd(_p(x_in), _p(node_attr), _p(w_embed), _p(bias), nodes, n, _p(h), _stream()),
              "segnn_embed_fwd")
    _bump()
    return h


# fp32 node-level products on tcgen05 (segnn_node_gemm_tf32x3); False (or SEGNN_NODE_GEMM_TF32X3=0): the FFMA kernel
NODE_GEMM_TF32X3 = os.environ.get("SEGNN_NODE_GEMM_TF32X3", "1") != "0"
NODE_GEMM_TF32X3_MIN_NODES = 4096


def node_gemm(x0, x1, w, n_out: int, bias=None, n_bias: int = 0, split: int = 0, tc=False):
    """w: dict with fp32 'w_s','w_v' [K][n_out] and, for the tensor-core kernel, 16-bit 'wt_s','wt_v' [n_out][K] packed
    with the operand format w['operand']. ``tc``: False (FFMA) or True (tcgen05). With split > 0 returns
    (y0 [nodes,4,split], y1 [nodes,4,n_out-split]), else one tensor [nodes,4,n_out]."""
    nodes, _, n_in = x0.shape
    dev = x0.device
    if split:
        y0 = torch.empty((nodes, 4, split), dtype=torch.float32, device=dev)
        y1 = torch.empty((nodes, 4, n_out - split), dtype=torch.float32, device=dev)
    else:
        y0, y1 = torch.empty((nodes, 4, n_out), dtype=torch.float32, device=dev), None
    with torch.cuda.device(dev):
        if tc:
            check(lib.segnn_node_gemm_tc(_p(x0), _p(x1), nodes, n_in, _p(w["wt_s"]), _p(w["wt_v"]), _p(bias), n_bias,
                                         n_out, _p(y0), _p(y1), split, int(w.get("operand", OPERAND_BF16)), _stream()),
                  "segnn_node_gemm_tc")
        elif NODE_GEMM_TF32X3 and nodes >= NODE_GEMM_TF32X3_MIN_NODES:
            # fp32-accurate on the tensor cores (3xTF32): 2 - 3.7x the FFMA kernel on the 102,400-node products of the
            # fp32-mode rollout (1.84 -> 0.96 ms, 1.94 -> 0.52 ms, 0.62 -> 0.28 ms); on a few hundred nodes its three
            # launches are no faster than the one FFMA launch (README training step unchanged), hence the floor
            K = (2 if x1 is not None else 1) * n_in
            ws = torch.empty(max(4, int(lib.segnn_node_gemm_tf32x3_workspace(K, n_out)) // 4), dtype=torch.float32,
                             device=dev)
            check(lib.segnn_node_gemm_tf32x3(_p(x0), _p(x1), nodes, n_in, _p(w["w_s"]), _p(w["w_v"]), _p(bias), n_bias,
                                             n_out, _p(y0), _p(y1), split, _p(ws), _stream()),
                  "segnn_node_gemm_tf32x3")
            _bump(2)
        else:
            check(lib.segnn_node_gemm(_p(x0), _p(x1), nodes, n_in, _p(w["w_s"]), _p(w["w_v"]), _p(bias), n_bias,
                                      n_out, _p(y0), _p(y1), split, _stream()), "segnn_node_gemm")
    _bump()
    return (y0, y1) if split else y0


def node_gemm_out16(x0, x1, w, n_out: int):
    """Tensor-core node GEMM with plain fp16 rows [nodes, 4, n_out] (segnn_node_gemm_tc_out16): for outputs that only
    feed tp_combine, which reads them through segnn_tp_combine_y16."""
    nodes, _, n_in = x0.shape
    y = torch.empty((nodes, 4, n_out), dtype=torch.float16, device=x0.device)
    if x0.dtype == torch.float16:  # fp16 feature copies: no conversion on load (segnn_node_gemm_tc_x16)
        if int(w.get("operand", OPERAND_BF16)) != OPERAND_FP16 or (x1 is not None and x1.dtype != torch.float16):
            raise ValueError("fp16 feature rows need fp16 operand weights and an fp16 second input")
        with torch.cuda.device(x0.device):
            check(lib.segnn_node_gemm_tc_x16(_p(x0), _p(x1), nodes, n_in, _p(w["wt_s"]), _p(w["wt_v"]), None, 0, n_out,
                                             _p(y), None, n_out, OPERAND_FP16, 2, _stream()), "segnn_node_gemm_tc_x16")
        _bump()
        return y
    with torch.cuda.device(x0.device):
        check(lib.segnn_node_gemm_tc_out16(_p(x0), _p(x1), nodes, n_in, _p(w["wt_s"]), _p(w["wt_v"]), n_out, _p(y),
                                           int(w.get("operand", OPERAND_BF16)), _stream()), "segnn_node_gemm_tc_out16")
    _bump()
    return y


def node_gemm_pair16(x0, w, n_out: int, bias, n_bias: int, split: int):
    """Tensor-core node GEMM with fp16 output, nodes interleaved in pairs (segnn_node_gemm_tc_pair16): returns
    (y0 [nodes/2, 4, split, 2], y1 [nodes/2, 4, n_out - split, 2]) float16."""
    nodes, _, n_in = x0.shape
    dev = x0.device
    if nodes % 2:
        raise ValueError("pair-interleaved projections need an even node count")
    y0 = torch.empty((nodes // 2, 4, split, 2), dtype=torch.float16, device=dev)
    y1 = torch.empty((nodes // 2, 4, n_out - split, 2), dtype=torch.float16, device=dev)
    if x0.dtype == torch.float16:  # fp16 feature copy (segnn_node_gemm_tc_x16)
        if int(w.get("operand", OPERAND_FP16)) != OPERAND_FP16:
            raise ValueError("fp16 feature rows need fp16 operand weights")
        with torch.cuda.device(dev):
            check(lib.segnn_node_gemm_tc_x16(_p(x0), None, nodes, n_in, _p(w["wt_s"]), _p(w["wt_v"]), _p(bias), n_bias,
                                             n_out, _p(y0), _p(y1), split, OPERAND_FP16, 1, _stream()),
                  "segnn_node_gemm_tc_x16")
        _bump()
        return y0, y1
    with torch.cuda.device(dev):
        check(lib.segnn_node_gemm_tc_pair16(_p(x0), None, nodes, n_in, _p(w["wt_s"]), _p(w["wt_v"]), _p(bias), n_bias,
                                            n_out, _p(y0), _p(y1), split, int(w.get("operand", OPERAND_FP16)),
                                            _stream()), "segnn_node_gemm_tc_pair16")
    _bump()
    return y0, y1


def pack_node_weight_tc(w: torch.Tensor, operand: int = OPERAND_BF16) -> torch.Tensor:
    """fp32 [K][n_out] -> bf16 / fp16 [n_out][K] (segnn_pack_node_weight_tc)."""
    K, n_out = w.shape
    out = torch.empty((n_out, K), dtype=torch.float16 if operand == OPERAND_FP16 else torch.bfloat16, device=w.device)
    with torch.cuda.device(w.device):
        check(lib.segnn_pack_node_weight_tc(_p(w), K, n_out, int(operand), _p(out), _stream()),
              "segnn_pack_node_weight_tc")
    _bump()
    return out


def tp_combine(y, node_attr, n: int, gate: bool, bias=None, residual=None, bn_mul=None, bn_add=None,
               out: Optional[torch.Tensor] = None, out16: Optional[str] = None):
    """out16 (fp16 GEMM rows only): 'both' returns (out fp32, fp16 copy), 'only' returns the fp16 copy alone
    (segnn_tp_combine_y16_x16)."""
    nodes = y.shape[0]
    if out16 is not None:
        if y.dtype != torch.float16 or out16 not in ("both", "only"):
            raise ValueError("out16 ('both' / 'only') needs fp16 GEMM rows")
        o = None if out16 == "only" else torch.empty((nodes, 4, n), dtype=torch.float32, device=y.device)
        o16 = torch.empty((nodes, 4, n), dtype=torch.float16, device=y.device)
        with torch.cuda.device(y.device):
            check(lib.segnn_tp_combine_y16_x16(_p(y), _p(node_attr), nodes, n, int(gate), _p(bias), _p(residual),
                                               _p(bn_mul), _p(bn_add), _p(o), _p(o16), _stream()),
                  "segnn_tp_combine_y16_x16")
        _bump()
        return o16 if o is None else (o, o16)
    o = out if out is not None else torch.empty((nodes, 4, n), dtype=torch.float32, device=y.device)
    fn = lib.segnn_tp_combine_y16 if y.dtype == torch.float16 else lib.segnn_tp_combine
    with torch.cuda.device(y.device):
        check(fn(_p(y), _p(node_attr), nodes, n, int(gate), _p(bias), _p(residual), _p(bn_mul), _p(bn_add), _p(o),
                 _stream()), "segnn_tp_combine")
    _bump()
    return o


def edge_layer(mode: int, pos, mass, batch_size: int, num_nodes: int, n: int, p, q, w_edge1, w2, bn_mul=None,
               bn_add=None, want_moments: bool = False, out16: bool = False):
    """w2: dict with fp32 blocks 'ss','vs','sv','vv','b' and (tensor-core mode) 'tc' image.  out16 (MODE_FP16_PACKED):
    the aggregate comes back as fp16 rows (segnn_edge_layer_fwd_out16)."""
    nodes = batch_size * num_nodes
    if out16:
        if mode != MODE_FP16_PACKED or want_moments:
            raise ValueError("fp16 aggregate rows are an output of the packed-half mode without moments")
        agg16 = torch.empty((nodes, 4, n), dtype=torch.float16, device=pos.device)
        with torch.cuda.device(pos.device):
            check(lib.segnn_edge_layer_fwd_out16(_p(pos), _p(mass), batch_size, num_nodes, n, _p(p), _p(q), _p(w_edge1),
                                                 _p(w2["b"]), _p(w2.get("tc")), _p(bn_mul), _p(bn_add), _p(agg16),
                                                 _stream()), "segnn_edge_layer_fwd_out16")
        _bump()
        return agg16
    if mode == MODE_FP32 and _use_gemm_form(batch_size, num_nodes, n):
        return edge_layer_gemm_fwd(pos, mass, batch_size, num_nodes, n, p, q, w_edge1, w2, bn_mul, bn_add, want_moments)
    agg = torch.empty((nodes, 4, n), dtype=torch.float32, device=pos.device)
    mom = torch.empty((nodes, 2 * n), dtype=torch.float32, device=pos.device) if want_moments else None
    with torch.cuda.device(pos.device):
        check(lib.segnn_edge_layer_fwd(mode, _p(pos), _p(mass), batch_size, num_nodes, n, _p(p), _p(q), _p(w_edge1),
                                       _p(w2.get("ss")), _p(w2.get("vs")), _p(w2.get("sv")), _p(w2.get("vv")),
                                       _p(w2["b"]), _p(w2.get("tc")), _p(bn_mul), _p(bn_add), _p(agg), _p(mom),
                                       _stream()), "segnn_edge_layer_fwd")
    _bump()
    return (agg, mom) if want_moments else agg


def head(h, node_attr, w_head, n: int):
    nodes = h.shape[0]
    pred = torch.empty((nodes, 6), dtype=torch.float32, device=h.device)
    with torch.cuda.device(h.device):
        check(lib.segnn_head_fwd(_p(h), _p(node_attr), _p(w_head), nodes, n, _p(pred), _stream()), "segnn_head_fwd")
    _bump()
    return pred


def integrate(pred, pos, vel, traj_pos=None, traj_vel=None, frame=None):
    nodes = pos.shape[0]
    max_frames = int(traj_pos.shape[0]) if traj_pos is not None else 0
    with torch.cuda.device(pos.device):
        check(lib.segnn_integrate(_p(pred), _p(pos), _p(vel), nodes, _p(traj_pos), _p(traj_vel), _p(frame),
                                  max_frames, _stream()), "segnn_integrate")
    _bump()


def counter_add(counter: torch.Tensor, delta: int):
    with torch.cuda.device(counter.device):
        check(lib.segnn_counter_add(_p(counter), int(delta), _stream()), "segnn_counter_add")
    _bump()


TC_MULTIPLICITIES = (32, 64, 96)


PACK_MSG1, PACK_MSG2, PACK_UPDATE1, PACK_UPDATE2, PACK_POOL1, PACK_EMBED, PACK_HEAD = range(7)


def pack_weights(kind: int, n: int, tp_weight: torch.Tensor, biases: Optional[torch.Tensor]) -> dict:
    """segnn_pack_weights: the reference's flat ``tp.weight`` / ``biases`` -> the operand blocks of the kernels (views of
    one buffer, same names and layouts as packing.pack_*)."""
    w = _f32(tp_weight.detach(), "tp.weight")
    b = None if biases is None else _f32(biases.detach(), "biases")
    total = int(lib.segnn_pack_weights_size(kind, n))
    out = torch.empty(total, dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        check(lib.segnn_pack_weights(kind, n, _p(w), _p(b), _p(out), _stream()), "segnn_pack_weights")
    _bump()
    if kind == PACK_MSG1:
        k = 6 * n * n
        return dict(w_s=out[:k].view(n, 6 * n), w_v=out[k:2 * k].view(n, 6 * n), bias=out[2 * k:2 * k + 2 * n],
                    w_edge=out[2 * k + 2 * n:])
    if kind == PACK_MSG2:
        a, c = 2 * n * n, n * n
        return dict(ss=out[:a].view(n, 2 * n), vs=out[a:2 * a].view(n, 2 * n), sv=out[2 * a:2 * a + c].view(n, n),
                    vv=out[2 * a + c:2 * a + 2 * c].view(n, n), b=out[2 * a + 2 * c:])
    if kind in (PACK_UPDATE1, PACK_UPDATE2, PACK_POOL1):
        rows = 2 * n if kind == PACK_UPDATE1 else n
        cols = 2 * n if kind == PACK_UPDATE2 else 3 * n
        k = rows * cols
        return dict(w_s=out[:k].view(rows, cols), w_v=out[k:2 * k].view(rows, cols), bias=out[2 * k:])
    if kind == PACK_EMBED:
        return dict(w=out[:6 * n].view(6, n), bias=out[6 * n:])
    return out.view(2, n, 2)


def fold_batchnorm(weight, bias, running_mean, running_var, n: int, eps: float, degree: float):
    """segnn_fold_batchnorm -> (mul [2n], add [n])."""
    # (keep the fp32 copies referenced until the launch: a temporary freed early could be handed out again)
    w32, b32, rm32, rv32 = [_f32(t.detach(), "batchnorm") for t in (weight, bias, running_mean, running_var)]
    mul = torch.empty(2 * n, dtype=torch.float32, device=weight.device)
    add = torch.empty(n, dtype=torch.float32, device=weight.device)
    with torch.cuda.device(weight.device):
        check(lib.segnn_fold_batchnorm(_p(w32), _p(b32), _p(rm32), _p(rv32), n,
                                       float(eps), float(degree), _p(mul), _p(add), _stream()), "segnn_fold_batchnorm")
    _bump()
    return mul, add


def pack_w2_tc(w2: dict, n: int, operand: int = OPERAND_BF16) -> torch.Tensor:
    """message_layer_2 weight image for the tcgen05 kernel: [128 lanes][3n] 16-bit pairs (segnn_pack_w2_tc)."""
    nbytes = lib.segnn_pack_w2_tc(None, None, None, None, n, int(operand), None, None)
    if nbytes <= 0:
        check(int(nbytes), "segnn_pack_w2_tc")
    out = torch.empty(nbytes // 4, dtype=torch.int32, device=w2["ss"].device)
    with torch.cuda.device(out.device):
        rc = lib.segnn_pack_w2_tc(_p(w2["ss"]), _p(w2["vs"]), _p(w2["sv"]), _p(w2["vv"]), n, int(operand), _p(out),
                                  _stream())
    if rc < 0:
        check(int(rc), "segnn_pack_w2_tc")
    _bump()
    return out


def gemm_tf32x3(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """C = A @ B in fp32 accuracy on the tensor cores (segnn_gemm_tf32x3: tf32 hi/lo split, three MMAs per product).
    ``a`` [M, K] and ``b`` [K, N] may be row-strided views (leading dimension = stride(0)); ``out`` likewise."""
    if not (a.is_cuda and b.is_cuda):
        raise RuntimeError("gemm_tf32x3 needs CUDA tensors (no CPU fallback)")
    assert a.dtype == b.dtype == torch.float32 and a.dim() == b.dim() == 2 and a.shape[1] == b.shape[0]
    assert a.stride(1) == 1 and b.stride(1) == 1, "row-major operands"
    M, K = a.shape
    N = b.shape[1]
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    assert out.shape == (M, N) and out.stride(1) == 1
    ws = torch.empty(max(4, lib.segnn_gemm_tf32x3_workspace(K, N) // 4), dtype=torch.float32, device=a.device)
    lda = a.stride(0) if M > 1 else max(K, a.stride(0))
    with torch.cuda.device(a.device):
        check(lib.segnn_gemm_tf32x3(_p(a), lda, _p(b), b.stride(0) if K > 1 else N, M, K, N, _p(out),
                                    out.stride(0) if M > 1 else N, _p(ws), _stream()), "segnn_gemm_tf32x3")
    _bump(2)
    return out


def gemm_tn_tf32x3(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None,
                   accumulate: bool = False, groups: int = 1) -> torch.Tensor:
    """C (+)= A^T @ B in fp32 accuracy on the tensor cores (segnn_gemm_tn_tf32x3): ``a`` [K, M] and ``b`` [K, N]
    row-major (row-strided views allowed), the contraction runs over their rows with a fixed-order split-K reduction."""
    if not (a.is_cuda and b.is_cuda):
        raise RuntimeError("gemm_tn_tf32x3 needs CUDA tensors (no CPU fallback)")
    assert a.dtype == b.dtype == torch.float32 and a.dim() == b.dim() == 2 and a.shape[0] == b.shape[0]
    assert a.stride(1) == 1 and b.stride(1) == 1, "row-major operands"
    K, M = a.shape
    N = b.shape[1]
    if groups > 1:  # `groups` blocks side by side in every row: sum_g a[:, g]^T b[:, g]
        assert M % groups == 0 and N % groups == 0
        M, N = M // groups, N // groups
    if out is None:
        assert not accumulate
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    assert out.shape == (M, N) and out.stride(1) == 1
    nbytes = int(lib.segnn_gemm_tn_tf32x3_workspace(K, M, N))
    if nbytes < 0:
        raise RuntimeError(f"gemm_tn_tf32x3: unsupported sizes K={K} M={M} N={N}")
    ws = torch.empty(max(4, nbytes // 4), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        check(lib.segnn_gemm_tn_grouped_tf32x3(_p(a), a.stride(0) if K > 1 else max(groups * M, a.stride(0)), _p(b),
                                               b.stride(0) if K > 1 else max(groups * N, b.stride(0)), K, M, N, groups,
                                               _p(out), out.stride(0) if M > 1 else N, int(accumulate), _p(ws),
                                               _stream()), "segnn_gemm_tn_grouped_tf32x3")
    _bump(2)
    return out


# Graphs with many nodes (BASELINE configuration 4: N = 1000) run the edge layer in GEMM form over the edge rows
# (csrc/segnn_edge_gemm.cu): from this many rows = B * N * N on, and with at most this much workspace per call.
GEMM_FORM_MIN_ROWS = 1 << 18
# Training (forward that keeps its rows, and every backward) runs in GEMM form at ANY size: the README training step
# (64 graphs x 5 bodies) takes 2.4 ms against 3.1 ms with the fused fp32 kernels, whose backward keeps one thread group
# per stationary node busy with a serial recompute.
GEMM_FORM_MIN_ROWS_TRAINING = 0
GEMM_FORM_BUDGET_BYTES = 6 << 30
# A training forward may leave the edge rows of a layer (11 n floats per row) in HBM for the backward call instead of
# having them recomputed, when they fit this many bytes per layer (180 GB of HBM3e: 6 layers of BASELINE configuration 4
# keep 17 GB).  0 disables it (recompute, nothing per edge stored between forward and backward).
GEMM_FORM_KEEP_BYTES_PER_LAYER = 4 << 30


GEMM_FORM_MAX_GRAPH_BYTES = 48 << 30  # a chunk holds whole graphs: one graph must fit (N ~ 2800 at n = 64)


def _use_gemm_form(batch_size: int, num_nodes: int, n: int, training: bool = False) -> bool:
    floor = GEMM_FORM_MIN_ROWS_TRAINING if training else GEMM_FORM_MIN_ROWS
    if not (batch_size * num_nodes * num_nodes >= floor and n % 4 == 0 and n <= 96 and num_nodes >= 2):
        return False
    # 16 n floats per edge row of ONE graph (backward): graphs too large for that stay on the fused kernels
    return 4 * 16 * n * num_nodes * num_nodes <= GEMM_FORM_MAX_GRAPH_BYTES


def _edge_gemm_workspace(batch_size: int, num_nodes: int, n: int, backward: bool, device):
    nbytes = int(lib.segnn_edge_layer_gemm_workspace(batch_size, num_nodes, n, int(backward), GEMM_FORM_BUDGET_BYTES))
    if nbytes < 0:
        raise RuntimeError("segnn_edge_layer_gemm_workspace: unsupported sizes")
    ws = torch.empty(nbytes // 4 + 64, dtype=torch.float32, device=device)
    assert ws.data_ptr() % 256 == 0
    return ws, nbytes


def gemm_form_keeps_rows(batch_size: int, num_nodes: int, n: int) -> bool:
    """True when a training forward of this size runs in GEMM form and keeps its edge rows for the backward call."""
    if not _use_gemm_form(batch_size, num_nodes, n, training=True):
        return False
    need = int(lib.segnn_edge_layer_gemm_workspace(batch_size, num_nodes, n, 0, 0))
    return 0 < need <= GEMM_FORM_KEEP_BYTES_PER_LAYER


def edge_layer_gemm_fwd(pos, mass, batch_size: int, num_nodes: int, n: int, p, q, w_edge1, w2, bn_mul=None,
                        bn_add=None, want_moments: bool = False, keep_rows: bool = False):
    """segnn_edge_layer_gemm_fwd: the fp32-mode edge layer of graphs with many nodes as 3xTF32 GEMMs over edge rows.
    keep_rows: also return (workspace, bytes) holding the rows of all graphs, for edge_layer_gemm_bwd(rows=...)."""
    nodes = batch_size * num_nodes
    agg = torch.empty((nodes, 4, n), dtype=torch.float32, device=pos.device)
    mom = torch.empty((nodes, 2 * n), dtype=torch.float32, device=pos.device) if want_moments else None
    if keep_rows:  # one chunk holding every graph
        nbytes = int(lib.segnn_edge_layer_gemm_workspace(batch_size, num_nodes, n, 0, 0))
        ws = torch.empty(nbytes // 4 + 64, dtype=torch.float32, device=pos.device)
    else:
        ws, nbytes = _edge_gemm_workspace(batch_size, num_nodes, n, False, pos.device)
    with torch.cuda.device(pos.device):
        check(lib.segnn_edge_layer_gemm_fwd(_p(pos), _p(mass), batch_size, num_nodes, n, _p(p), _p(q), _p(w_edge1),
                                            _p(w2["ss"]), _p(w2["vs"]), _p(w2["sv"]), _p(w2["vv"]), _p(w2["b"]),
                                            _p(bn_mul), _p(bn_add), _p(agg), _p(mom), _p(ws), nbytes, _stream()),
              "segnn_edge_layer_gemm_fwd")
    _bump(8)
    out = (agg, mom) if want_moments else (agg,)
    if keep_rows:
        return out + ((ws, nbytes),)
    return out if want_moments else agg


def edge_layer_gemm_bwd(pos, mass, batch_size: int, num_nodes: int, n: int, p, q, w_edge1, w2, bn_a, bn_b, bn_c, dagg,
                        dP=None, dQ=None, gz=None, rows=None, side=None):
    """segnn_edge_layer_gemm_bwd: same results as edge_layer_bwd (dP, dQ, message_layer_2 gradient blocks, dw_edge1).
    rows: what edge_layer_gemm_fwd(keep_rows=True) returned for the same layer (skips the recompute).
    side (an object with .run(fn), training._SideWork): when every graph fits one chunk the message_layer_2 weight
    gradients (split-K TN GEMMs, off the critical path) are launched through it beside the data-gradient chain; the
    returned gradient blocks are complete once the caller has joined it."""
    nodes = batch_size * num_nodes
    dev = pos.device
    f = dict(dtype=torch.float32, device=dev)
    if dP is None:
        dP, dQ = torch.empty((nodes, 4, 3 * n), **f), torch.empty((nodes, 4, 3 * n), **f)
        gz = torch.empty(6 * n * n + 2 * n, **f)
    g = dict(ss=gz[: 2 * n * n].view(n, 2 * n), vs=gz[2 * n * n: 4 * n * n].view(n, 2 * n),
             sv=gz[4 * n * n: 5 * n * n].view(n, n), vv=gz[5 * n * n: 6 * n * n].view(n, n), b=gz[6 * n * n:])
    dwe_partial = torch.empty((nodes, 6 * n), **f)
    w2t = {k: (w2[k + "_t"] if k + "_t" in w2 else w2[k].t().contiguous()) for k in ("ss", "vs", "sv", "vv")}
    dagg = dagg.contiguous()
    ws, nbytes = _edge_gemm_workspace(batch_size, num_nodes, n, True, dev)
    def launch(phases: int):
        with torch.cuda.device(dev):
            check(lib.segnn_edge_layer_gemm_bwd_phases(
                _p(pos), _p(mass), batch_size, num_nodes, n, _p(p), _p(q), _p(w_edge1), _p(w2["ss"]), _p(w2["vs"]),
                _p(w2["sv"]), _p(w2["vv"]), _p(w2["b"]), _p(w2t["ss"]), _p(w2t["vs"]), _p(w2t["sv"]), _p(w2t["vv"]),
                _p(bn_a), _p(bn_b), _p(bn_c), _p(dagg), _p(dP), _p(dQ), _p(g["ss"]), _p(g["vs"]), _p(g["sv"]),
                _p(g["vv"]), _p(g["b"]), _p(dwe_partial), _p(ws), nbytes, _p(rows[0]) if rows is not None else None,
                rows[1] if rows is not None else 0, phases, _stream()), "segnn_edge_layer_gemm_bwd_phases")
    one_chunk = nbytes >= int(lib.segnn_edge_layer_gemm_workspace(batch_size, num_nodes, n, 1, 0))
    if side is not None and one_chunk and SIDE_STREAM_W2_GRADS:
        launch(1)
        side.run(lambda: launch(2))  # the closure keeps ws / rows / g alive until the caller joins the side stream
        launch(4)
    else:
        launch(7)
    _bump(24 if rows is None else 18)
    return dP, dQ, g, colsum(dwe_partial)


def tc_available() -> bool:
    """True when the tcgen05 (SEGNN_MODE_BF16_TC) edge kernel is compiled into the library."""
    return lib.segnn_pack_w2_tc(None, None, None, None, 96, OPERAND_BF16, None, None) > 0


# ---- training-side wrappers (fp32) --------------------------------------------------------------------------------
def colsum(x: torch.Tensor, y: Optional[torch.Tensor] = None, mode: int = 0) -> torch.Tensor:
    """Deterministic column reduction of a dense [rows, cols] view: sum x (0), sum x^2 (1), sum x*y (2)."""
    rows, cols = x.shape
    assert x.is_contiguous() and (y is None or (y.is_contiguous() and y.shape == x.shape))
    out = torch.empty(cols, dtype=torch.float32, device=x.device)
    ws = torch.empty(max(1, lib.segnn_colsum_workspace(rows, cols) // 8), dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.segnn_colsum(_p(x), _p(y), rows, cols, mode, _p(ws), _p(out), _stream()), "segnn_colsum")
    _bump(1 if rows <= 2048 else 2)  # up to 2048 rows: both reduction stages in one launch
    return out


def colsum2(xa, ya, mode_a: int, xb, yb, mode_b: int):
    """Two deterministic column reductions (see colsum) in one launch up to 2048 rows each (segnn_colsum2)."""
    (ra, ca), (rb, cb) = xa.shape, xb.shape
    assert xa.is_contiguous() and xb.is_contiguous() and (ya is None or ya.is_contiguous()) \
        and (yb is None or yb.is_contiguous())
    out_a = torch.empty(ca, dtype=torch.float32, device=xa.device)
    out_b = torch.empty(cb, dtype=torch.float32, device=xa.device)
    need = max(lib.segnn_colsum_workspace(ra, ca), lib.segnn_colsum_workspace(rb, cb))
    ws = torch.empty(max(1, need // 8), dtype=torch.float64, device=xa.device)
    with torch.cuda.device(xa.device):
        check(lib.segnn_colsum2(_p(xa), _p(ya), ra, ca, mode_a, _p(out_a), _p(xb), _p(yb), rb, cb, mode_b, _p(out_b),
                                _p(ws), _stream()), "segnn_colsum2")
    _bump(1 if max(ra, rb) <= 2048 else 4)
    return out_a, out_b


def lincomb(dy, x, A, B=None, C=None):
    """out[r][c] = A[c]*dy[r][c] + B[c]*x[r][c] + C[c] on dense [rows, cols] views."""
    rows, cols = dy.shape
    assert dy.is_contiguous() and A.numel() == cols
    out = torch.empty_like(dy)
    with torch.cuda.device(dy.device):
        check(lib.segnn_lincomb(_p(dy), _p(x), _p(A), _p(B), _p(C), rows, cols, _p(out), _stream()), "segnn_lincomb")
    _bump()
    return out


def add3(a, b, c=None):
    out = torch.empty_like(a)
    with torch.cuda.device(a.device):
        check(lib.segnn_add3(_p(a), _p(b), _p(c), a.numel(), _p(out), _stream()), "segnn_add3")
    _bump()
    return out


def tp_combine_bwd(y, node_attr, n: int, gate: bool, bias, dout):
    nodes = y.shape[0]
    n0 = 2 * n if gate else n
    dy = torch.empty((nodes, 4, n0 + n), dtype=torch.float32, device=y.device)
    dz0 = torch.empty((nodes, n0), dtype=torch.float32, device=y.device)
    with torch.cuda.device(y.device):
        check(lib.segnn_tp_combine_bwd(_p(y), _p(node_attr), nodes, n, int(gate), _p(bias), _p(dout.contiguous()),
                                       _p(dy), _p(dz0), _stream()), "segnn_tp_combine_bwd")
    _bump()
    return dy, dz0


def node_gemm_wgrad(x0, x1, dy0, dy1, split: int):
    """dw_s, dw_v [K][n_out] of node_gemm(x0 | x1) -> (dy0 | dy1)."""
    nodes, _, n_in = x0.shape
    K = n_in * (2 if x1 is not None else 1)
    n_out = dy0.shape[2] + (dy1.shape[2] if dy1 is not None else 0)
    dev = x0.device
    dw_s = torch.empty((K, n_out), dtype=torch.float32, device=dev)
    dw_v = torch.empty((K, n_out), dtype=torch.float32, device=dev)
    ws = torch.empty(max(1, lib.segnn_node_gemm_wgrad_workspace(nodes, K, n_out) // 4), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.segnn_node_gemm_wgrad(_p(x0), _p(x1), _p(dy0), _p(dy1), split, nodes, n_in, n_out, _p(ws),
                                        _p(dw_s), _p(dw_v), _stream()), "segnn_node_gemm_wgrad")
    _bump(4)
    return dw_s, dw_v


_SIDE_STREAMS = {}
# GEMM-form edge backward: message_layer_2 weight gradients on the side stream (SEGNN_SIDE_W2_GRADS=0: in line)
SIDE_STREAM_W2_GRADS = os.environ.get("SEGNN_SIDE_W2_GRADS", "1") != "0"


def side_stream(device, k: int = 0) -> "torch.cuda.Stream":
    """One of a few cached side streams per device: independent kernels of the backward pass (the dQ pass of the edge
    backward, the weight-gradient GEMMs) run there so that they overlap the critical dgrad chain -- every one of them
    fills only a fraction of the 148 SMs at training sizes.  Fork / join are stream waits, which a CUDA-graph capture
    records as parallel branches."""
    key = (torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device(), k)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


_SPLIT_WGRAD_MAX_EDGES = 1 << 16  # below this the edge backward runs dP and the weight gradients as separate launches


def edge_layer_bwd(pos, mass, batch_size: int, num_nodes: int, n: int, p, q, w_edge1, w2, bn_a, bn_b, bn_c, dagg,
                   side=None):
    """Backward of the fused edge layer with recompute. Returns dP, dQ [nodes,4,3n], the message_layer_2 gradient
    blocks {ss, vs, sv, vv, b} and dw_edge1 [6n]."""
    nodes = batch_size * num_nodes
    dev = pos.device
    f = dict(dtype=torch.float32, device=dev)
    dP, dQ = torch.empty((nodes, 4, 3 * n), **f), torch.empty((nodes, 4, 3 * n), **f)
    gz = torch.empty(6 * n * n + 2 * n, **f)  # written by the fixed-order slab reduction (no atomics)
    if _use_gemm_form(batch_size, num_nodes, n, training=True):
        return edge_layer_gemm_bwd(pos, mass, batch_size, num_nodes, n, p, q, w_edge1, w2, bn_a, bn_b, bn_c, dagg,
                                   dP, dQ, gz, side=side)
    ws = torch.empty(max(1, int(lib.segnn_edge_layer_bwd_workspace(batch_size, num_nodes, n)) // 4), **f)
    g = dict(ss=gz[: 2 * n * n].view(n, 2 * n), vs=gz[2 * n * n: 4 * n * n].view(n, 2 * n),
             sv=gz[4 * n * n: 5 * n * n].view(n, n), vv=gz[5 * n * n: 6 * n * n].view(n, n), b=gz[6 * n * n:])
    dwe_partial = torch.empty((nodes, 6 * n), **f)
    w2t = {k: (w2[k + "_t"] if k + "_t" in w2 else w2[k].t().contiguous()) for k in ("ss", "vs", "sv", "vv")}
    dagg = dagg.contiguous()
    with torch.cuda.device(dev):
        main, side = torch.cuda.current_stream(dev), side_stream(dev, 0)
        side.wait_stream(main)  # every input (and the zero-fill above) is ready

        def launch(pas, dout):
            check(lib.segnn_edge_layer_bwd(pas, _p(pos), _p(mass), batch_size, num_nodes, n, _p(p), _p(q), _p(w_edge1),
                                           _p(w2["ss"]), _p(w2["vs"]), _p(w2["sv"]), _p(w2["vv"]), _p(w2["b"]),
                                           _p(w2t["ss"]), _p(w2t["vs"]), _p(w2t["sv"]), _p(w2t["vv"]), _p(bn_a),
                                           _p(bn_b), _p(bn_c), _p(dagg), _p(dout), _p(g["ss"]),
                                           _p(g["vs"]), _p(g["sv"]), _p(g["vv"]), _p(g["b"]), _p(dwe_partial),
                                           _p(ws), _stream()), "segnn_edge_layer_bwd")
            _bump(2 if pas in (0, 3) else 1)
        with torch.cuda.stream(side):  # pass 1 (dQ) is independent of pass 0 (dP, weight gradients): overlap them
            launch(1, dQ)
        if nodes * (num_nodes - 1) <= _SPLIT_WGRAD_MAX_EDGES:
            # training-size graphs: the GPU is far from full, so the weight gradients (slab read-modify-writes) get
            # their own launch on a second side stream and leave the critical dP -> node-backward chain
            side2 = side_stream(dev, 2)
            side2.wait_stream(main)
            with torch.cuda.stream(side2):
                launch(3, None)
            launch(2, dP)
            main.wait_stream(side2)
        else:
            launch(0, dP)
        main.wait_stream(side)
    return dP, dQ, g, colsum(dwe_partial)


def embed_bwd(x_in, node_attr, dh, n: int):
    nodes = x_in.shape[0]
    contrib = torch.empty((nodes, 7 * n), dtype=torch.float32, device=x_in.device)
    with torch.cuda.device(x_in.device):
        check(lib.segnn_embed_bwd(_p(x_in), _p(node_attr), _p(dh.contiguous()), nodes, n, _p(contrib), _stream()),
              "segnn_embed_bwd")
    _bump()
    s = colsum(contrib)
    return s[: 6 * n].view(6, n), s[6 * n:]


def head_bwd(h, node_attr, w_head, dpred, n: int):
    """dh [nodes,4,n] and dw_head [2][n][2]."""
    nodes = h.shape[0]
    dh = torch.empty((nodes, 4, n), dtype=torch.float32, device=h.device)
    contrib = torch.empty((nodes, 4 * n), dtype=torch.float32, device=h.device)
    with torch.cuda.device(h.device):
        check(lib.segnn_head_bwd(_p(h), _p(node_attr), _p(w_head), _p(dpred), nodes, n, _p(dh), _p(contrib), _stream()),
              "segnn_head_bwd")
    _bump()
    s = colsum(contrib).view(2, 2, n)  # (ws0, ws1), (wv0, wv1)
    return dh, s.permute(0, 2, 1).contiguous()


def bn_forward_coeffs(bn, n: int, rows: float, deg: float, sums, sq, v_planes: int, training: bool, update: bool):
    """e3nn BatchNorm statistics -> folded affine (one launch). bn: dict(weight, bias, running_mean, running_var, eps,
    momentum). Returns dict(mulcols [4n], addcols [4n], stats [5, n])."""
    dev = bn["weight"].device
    cols = torch.empty((2, 4 * n), dtype=torch.float32, device=dev)
    stats = torch.empty((5, n), dtype=torch.float32, device=dev)
    rm, rv = bn["running_mean"], bn["running_var"]
    native = rm.dtype == torch.float32 and rv.dtype == torch.float32 and rm.is_contiguous() and rv.is_contiguous()
    rm32, rv32 = (rm, rv) if native else (rm.to(torch.float32).contiguous(), rv.to(torch.float32).contiguous())
    with torch.cuda.device(dev):
        check(lib.segnn_bn_coeffs_fwd(_p(sums), _p(sq), v_planes, n, float(rows), float(deg), _p(bn["weight"].contiguous()),
                                      _p(bn["bias"].contiguous()), _p(rm32), _p(rv32), float(bn["eps"]),
                                      float(bn["momentum"]), int(training), int(update), _p(cols), _p(stats), _stream()),
              "segnn_bn_coeffs_fwd")
    _bump()
    if training and update:
        if native:  # the kernel wrote through raw pointers: tell torch (SEGNN.packed() keys its eval-mode
            #         folded-BatchNorm cache on the buffers' version counters)
            torch.autograd.graph.increment_version(rm)
            torch.autograd.graph.increment_version(rv)
        else:  # e.g. a .double() module: write the updated statistics back
            rm.copy_(rm32)
            rv.copy_(rv32)
    return dict(mulcols=cols[0], addcols=cols[1], stats=stats)


def bn_backward_coeffs(bn, st, n: int, rows: float, deg: float, sum_g, sum_gx, training: bool):
    """Returns dict(A4, B4, C4 [4n] planar-column coefficients; edge = (bn_a [2n], bn_b [2n], bn_c [n]); dweight [2n],
    dbias [n])."""
    dev = sum_g.device
    cols = torch.empty((3, 4 * n), dtype=torch.float32, device=dev)
    edge = torch.empty(5 * n, dtype=torch.float32, device=dev)
    dparam = torch.empty(3 * n, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.segnn_bn_coeffs_bwd(_p(sum_g), _p(sum_gx), n, float(rows), float(deg), _p(bn["weight"].contiguous()),
                                      _p(st["stats"]), int(training), _p(cols), _p(edge), _p(dparam), _stream()),
              "segnn_bn_coeffs_bwd")
    _bump()
    return dict(A4=cols[0], B4=cols[1], C4=cols[2], bn_a=edge[: 2 * n], bn_b=edge[2 * n: 4 * n], bn_c=edge[4 * n:],
                dweight=dparam[: 2 * n], dbias=dparam[2 * n:])
