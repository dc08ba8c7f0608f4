"""PyTorch custom operators (``torch.ops.segnn_b200.*``) over the C ABI, registered with ``torch.library`` and, where a
backward kernel exists, ``register_autograd`` -- the operator boundary BASELINE.json's north_star names and SURVEY.md
section 8(b) describes (the reference has no FFI: its operators are the e3nn / PyG calls inside
``models/segnn/segnn.py:264-304`` and ``models/segnn/o3_building_blocks.py:150-162``).

Every operator takes and returns plain tensors (no dicts), has a fake (meta) implementation so that it traces under
``torch.compile`` / ``torch.export`` / FakeTensorMode, and launches on the current CUDA stream (CUDA-graph capturable).
The arithmetic is the same kernels ``ops.py`` launches; nothing here computes on the host and there is no fallback:
CPU tensors raise.

Operators (shapes: nodes = B * N, hidden features planar ``[nodes, 4, n]``):

==========================  ================================================================  =====================
operator                    reference code it replaces                                         backward kernel
==========================  ================================================================  =====================
``prep``                    ``O3Transform`` node part (``o3_building_blocks.py:230-278``)      -- (inputs only)
``embed``                   ``embedding_layer`` (``segnn.py:62-66``)                           ``segnn_embed_bwd``
``node_gemm``               weight contraction of a node-level ``O3TensorProduct`` (:150-162)  dgrad = same op with
                                                                                               transposed blocks,
                                                                                               ``segnn_node_gemm_wgrad``
``tp_combine``              attribute coupling + bias (+ e3nn ``Gate``) of the same product    ``segnn_tp_combine_bwd``
``edge_layer``              ``SEGNNLayer.message`` + aggregation (``segnn.py:264-284,205``)    ``segnn_edge_layer_bwd``
``edge_layer_tc``           same on tcgen05 (bf16 / fp16 / packed-fp16 modes), eval only       --
``head``                    ``pre_pool2`` on the gated features (``segnn.py:96-100``)          ``segnn_head_bwd``
``prep_lmax``               ``O3Transform`` node part for ``lmax_attr`` in 0..2                -- (inputs only)
``edge_attr_list``          ``O3Transform`` edge part on an explicit (kNN) edge list           -- (inputs only)
``message_input_list``      ``cat(x_i, x_j, additional_message_features)`` (``segnn.py:264``)  -- (inference)
``segment_reduce``          PyG ``aggr="add"`` / ``scatter(reduce="mean")`` over the targets   -- (inference)
==========================  ================================================================  =====================
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import Tensor

from . import ops

NS = "segnn_b200"
__all__ = ["prep", "embed", "node_gemm", "tp_combine", "edge_layer", "edge_layer_tc", "head", "prep_lmax",
           "edge_attr_list", "message_input_list", "segment_reduce", "NS"]


def _cuda(*ts: Optional[Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("segnn_b200 operators run on CUDA tensors only (no CPU fallback)")


# ---- prep --------------------------------------------------------------------------------------------------------------
@torch.library.custom_op(f"{NS}::prep", mutates_args=())
def prep(pos: Tensor, vel: Tensor, batch_size: int, num_nodes: int) -> Tuple[Tensor, Tensor]:
    """(x_in [nodes, 7], node_attr [nodes, 4]) of the fully connected graphs: node part of O3Transform."""
    _cuda(pos, vel)
    return ops.prep(pos, vel, batch_size, num_nodes)


@prep.register_fake
def _(pos, vel, batch_size, num_nodes):
    nodes = batch_size * num_nodes
    return pos.new_empty((nodes, 7), dtype=torch.float32), pos.new_empty((nodes, 4), dtype=torch.float32)


# ---- embed -------------------------------------------------------------------------------------------------------------
@torch.library.custom_op(f"{NS}::embed", mutates_args=())
def embed(x_in: Tensor, node_attr: Tensor, w: Tensor, bias: Tensor, n: int) -> Tensor:
    """Embedding tensor product 7 -> n x 0e + n x 1o, planar output [nodes, 4, n]; w [6, n], bias [n]."""
    _cuda(x_in, node_attr, w, bias)
    return ops.embed(x_in, node_attr, w, bias, n)


@embed.register_fake
def _(x_in, node_attr, w, bias, n):
    return x_in.new_empty((x_in.shape[0], 4, n), dtype=torch.float32)


@torch.library.custom_op(f"{NS}::embed_bwd", mutates_args=())
def embed_bwd(x_in: Tensor, node_attr: Tensor, dh: Tensor, n: int) -> Tuple[Tensor, Tensor]:
    dw, db = ops.embed_bwd(x_in, node_attr, dh, n)
    return dw.clone(), db.clone()  # views of one reduction buffer: custom-op outputs must not alias each other


@embed_bwd.register_fake
def _(x_in, node_attr, dh, n):
    return x_in.new_empty((6, n), dtype=torch.float32), x_in.new_empty((n,), dtype=torch.float32)


def _embed_setup(ctx, inputs, output):
    x_in, node_attr, w, bias, n = inputs
    ctx.save_for_backward(x_in, node_attr)
    ctx.n = n


def _embed_backward(ctx, dh):
    x_in, node_attr = ctx.saved_tensors
    dw, db = embed_bwd(x_in, node_attr, dh.contiguous(), ctx.n)
    return None, None, dw, db, None


embed.register_autograd(_embed_backward, setup_context=_embed_setup)


# ---- node_gemm ---------------------------------------------------------------------------------------------------------
@torch.library.custom_op(f"{NS}::node_gemm", mutates_args=())
def node_gemm(x0: Tensor, x1: Optional[Tensor], w_s: Tensor, w_v: Tensor, bias: Optional[Tensor], n_bias: int) -> Tensor:
    """y[c] = (x0 | x1)[c] @ W_c per plane c (scalar plane: w_s, vector planes: w_v; both [K, n_out], K = n_in or
    2 n_in), bias [n_bias] added to the first n_bias columns of the scalar plane. fp32 FFMA kernel."""
    _cuda(x0, x1, w_s, w_v, bias)
    return ops.node_gemm(x0, x1, dict(w_s=w_s, w_v=w_v), int(w_s.shape[1]), bias=bias, n_bias=n_bias)


@node_gemm.register_fake
def _(x0, x1, w_s, w_v, bias, n_bias):
    return x0.new_empty((x0.shape[0], 4, w_s.shape[1]), dtype=torch.float32)


@torch.library.custom_op(f"{NS}::node_gemm_wgrad", mutates_args=())
def node_gemm_wgrad(x0: Tensor, x1: Optional[Tensor], dy: Tensor) -> Tuple[Tensor, Tensor]:
    return ops.node_gemm_wgrad(x0, x1, dy, None, 0)


@node_gemm_wgrad.register_fake
def _(x0, x1, dy):
    K = x0.shape[2] * (2 if x1 is not None else 1)
    return (x0.new_empty((K, dy.shape[2]), dtype=torch.float32), x0.new_empty((K, dy.shape[2]), dtype=torch.float32))


def _node_gemm_setup(ctx, inputs, output):
    x0, x1, w_s, w_v, bias, n_bias = inputs
    ctx.save_for_backward(x0, x1, w_s, w_v)
    ctx.n_bias, ctx.has_bias = n_bias, bias is not None


def _node_gemm_backward(ctx, dy):
    x0, x1, w_s, w_v = ctx.saved_tensors
    dy = dy.contiguous()
    dw_s, dw_v = node_gemm_wgrad(x0, x1, dy)
    dx = node_gemm(dy, None, w_s.t().contiguous(), w_v.t().contiguous(), None, 0)  # dgrad: same kernel, W^T
    n_in = x0.shape[2]
    dx0 = dx if x1 is None else dx[:, :, :n_in].contiguous()
    dx1 = None if x1 is None else dx[:, :, n_in:].contiguous()
    dbias = ops.colsum(dy[:, 0, :ctx.n_bias].contiguous()) if ctx.has_bias else None
    return dx0, dx1, dw_s, dw_v, dbias, None


node_gemm.register_autograd(_node_gemm_backward, setup_context=_node_gemm_setup)


# ---- tp_combine --------------------------------------------------------------------------------------------------------
@torch.library.custom_op(f"{NS}::tp_combine", mutates_args=())
def tp_combine(y: Tensor, node_attr: Tensor, n: int, gate: bool, bias: Optional[Tensor]) -> Tensor:
    """Attribute coupling of a node-level tensor product on the GEMM rows y [nodes, 4, n0 + n] (n0 = 2n with the gate,
    n without), + bias [n0] on the scalar outputs, + e3nn Gate (SiLU scalars, sigmoid gates) when ``gate``."""
    _cuda(y, node_attr, bias)
    return ops.tp_combine(y, node_attr, n, gate, bias=bias)


@tp_combine.register_fake
def _(y, node_attr, n, gate, bias):
    return y.new_empty((y.shape[0], 4, n), dtype=torch.float32)


@torch.library.custom_op(f"{NS}::tp_combine_bwd", mutates_args=())
def tp_combine_bwd(y: Tensor, node_attr: Tensor, n: int, gate: bool, bias: Optional[Tensor], dout: Tensor) -> Tuple[Tensor, Tensor]:
    return ops.tp_combine_bwd(y, node_attr, n, gate, bias, dout)


@tp_combine_bwd.register_fake
def _(y, node_attr, n, gate, bias, dout):
    n0 = 2 * n if gate else n
    return y.new_empty((y.shape[0], 4, n0 + n), dtype=torch.float32), y.new_empty((y.shape[0], n0), dtype=torch.float32)


def _tp_combine_setup(ctx, inputs, output):
    y, node_attr, n, gate, bias = inputs
    ctx.save_for_backward(y, node_attr, bias)
    ctx.n, ctx.gate = n, gate


def _tp_combine_backward(ctx, dout):
    y, node_attr, bias = ctx.saved_tensors
    dy, dz0 = tp_combine_bwd(y, node_attr, ctx.n, ctx.gate, bias, dout.contiguous())
    dbias = ops.colsum(dz0) if bias is not None else None
    return dy, None, None, None, dbias


tp_combine.register_autograd(_tp_combine_backward, setup_context=_tp_combine_setup)


# ---- fused edge layer --------------------------------------------------------------------------------------------------
@torch.library.custom_op(f"{NS}::edge_layer", mutates_args=())
def edge_layer(pos: Tensor, mass: Tensor, batch_size: int, num_nodes: int, n: int, p: Tensor, q: Tensor, w_edge1: Tensor,
               w2_ss: Tensor, w2_vs: Tensor, w2_sv: Tensor, w2_vv: Tensor, b2: Tensor) -> Tensor:
    """agg_i = sum_j message(i, j): hoisted message_layer_1 combine + gate + message_layer_2 + gate + sum over the
    senders of the fully connected graph, fp32 FFMA kernel (differentiable)."""
    _cuda(pos, mass, p, q)
    return ops.edge_layer(ops.MODE_FP32, pos, mass, batch_size, num_nodes, n, p, q, w_edge1,
                          dict(ss=w2_ss, vs=w2_vs, sv=w2_sv, vv=w2_vv, b=b2))


@edge_layer.register_fake
def _(pos, mass, batch_size, num_nodes, n, p, q, w_edge1, w2_ss, w2_vs, w2_sv, w2_vv, b2):
    return pos.new_empty((batch_size * num_nodes, 4, n), dtype=torch.float32)


@torch.library.custom_op(f"{NS}::edge_layer_bwd", mutates_args=())
def edge_layer_bwd(pos: Tensor, mass: Tensor, batch_size: int, num_nodes: int, n: int, p: Tensor, q: Tensor,
                   w_edge1: Tensor, w2_ss: Tensor, w2_vs: Tensor, w2_sv: Tensor, w2_vv: Tensor, b2: Tensor,
                   dagg: Tensor) -> List[Tensor]:
    """[dP, dQ, dw_edge1, dss, dvs, dsv, dvv, db2]; fixed-order reductions (bit-identical run to run)."""
    # gradient reaching a message = A * dagg_i + B * m + C per channel: (1, 0, 0) is the plain sum over the senders
    one = torch.ones(2 * n, dtype=torch.float32, device=pos.device)
    zero, zero_c = torch.zeros_like(one), torch.zeros(n, dtype=torch.float32, device=pos.device)
    dP, dQ, g, dwe = ops.edge_layer_bwd(pos, mass, batch_size, num_nodes, n, p, q, w_edge1,
                                        dict(ss=w2_ss, vs=w2_vs, sv=w2_sv, vv=w2_vv, b=b2), one, zero, zero_c, dagg)
    return [dP, dQ, dwe, g["ss"].clone(), g["vs"].clone(), g["sv"].clone(), g["vv"].clone(), g["b"].clone()]


@edge_layer_bwd.register_fake
def _(pos, mass, batch_size, num_nodes, n, p, q, w_edge1, w2_ss, w2_vs, w2_sv, w2_vv, b2, dagg):
    e = lambda *s: pos.new_empty(s, dtype=torch.float32)
    return [torch.empty_like(p), torch.empty_like(q), e(6 * n), e(n, 2 * n), e(n, 2 * n), e(n, n), e(n, n), e(2 * n)]


def _edge_layer_setup(ctx, inputs, output):
    pos, mass, B, N, n, p, q, w_edge1, ss, vs, sv, vv, b2 = inputs
    ctx.save_for_backward(pos, mass, p, q, w_edge1, ss, vs, sv, vv, b2)
    ctx.dims = (B, N, n)


def _edge_layer_backward(ctx, dagg):
    pos, mass, p, q, w_edge1, ss, vs, sv, vv, b2 = ctx.saved_tensors
    B, N, n = ctx.dims
    dP, dQ, dwe, dss, dvs, dsv, dvv, db = edge_layer_bwd(pos, mass, B, N, n, p, q, w_edge1, ss, vs, sv, vv, b2,
                                                        dagg.contiguous())
    return None, None, None, None, None, dP, dQ, dwe.reshape(w_edge1.shape), dss, dvs, dsv, dvv, db


edge_layer.register_autograd(_edge_layer_backward, setup_context=_edge_layer_setup)


@torch.library.custom_op(f"{NS}::edge_layer_tc", mutates_args=())
def edge_layer_tc(mode: int, pos: Tensor, mass: Tensor, batch_size: int, num_nodes: int, n: int, p: Tensor, q: Tensor,
                  w_edge1: Tensor, b2: Tensor, w2_tc: Tensor, bn_mul: Optional[Tensor], bn_add: Optional[Tensor]) -> Tensor:
    """The same layer on tcgen05 (mode = ops.MODE_BF16_TC / MODE_FP16_TC / MODE_FP16_PACKED), eval-mode BatchNorm folded
    into (bn_mul, bn_add); w2_tc is the operand image of segnn_pack_w2_tc. Inference only (no autograd)."""
    _cuda(pos, mass, p, q, w2_tc)
    return ops.edge_layer(mode, pos, mass, batch_size, num_nodes, n, p, q, w_edge1, dict(b=b2, tc=w2_tc), bn_mul, bn_add)


@edge_layer_tc.register_fake
def _(mode, pos, mass, batch_size, num_nodes, n, p, q, w_edge1, b2, w2_tc, bn_mul, bn_add):
    return pos.new_empty((batch_size * num_nodes, 4, n), dtype=torch.float32)


# ---- head --------------------------------------------------------------------------------------------------------------
@torch.library.custom_op(f"{NS}::head", mutates_args=())
def head(h: Tensor, node_attr: Tensor, w_head: Tensor, n: int) -> Tensor:
    """pre_pool2: gated features [nodes, 4, n] -> prediction [nodes, 6] (two vectors); w_head [2, n, 2]."""
    _cuda(h, node_attr, w_head)
    return ops.head(h, node_attr, w_head, n)


@head.register_fake
def _(h, node_attr, w_head, n):
    return h.new_empty((h.shape[0], 6), dtype=torch.float32)


@torch.library.custom_op(f"{NS}::head_bwd", mutates_args=())
def head_bwd(h: Tensor, node_attr: Tensor, w_head: Tensor, dpred: Tensor, n: int) -> Tuple[Tensor, Tensor]:
    return ops.head_bwd(h, node_attr, w_head, dpred, n)


@head_bwd.register_fake
def _(h, node_attr, w_head, dpred, n):
    return torch.empty_like(h), torch.empty_like(w_head)


def _head_setup(ctx, inputs, output):
    h, node_attr, w_head, n = inputs
    ctx.save_for_backward(h, node_attr, w_head)
    ctx.n = n


def _head_backward(ctx, dpred):
    h, node_attr, w_head = ctx.saved_tensors
    dh, dw = head_bwd(h, node_attr, w_head, dpred.contiguous(), ctx.n)
    return dh, None, dw, None


head.register_autograd(_head_backward, setup_context=_head_setup)


# ---- generic-irreps path: lmax_attr <= 2 geometry and explicit edge lists (kNN graphs), inference -------------------------
@torch.library.custom_op(f"{NS}::prep_lmax", mutates_args=())
def prep_lmax(pos: Tensor, vel: Tensor, batch_size: int, num_nodes: int, lmax_attr: int) -> Tuple[Tensor, Tensor]:
    """(x_in [nodes, 7], node_attr [nodes, (lmax_attr + 1)^2]) of the fully connected graphs, lmax_attr in 0..2."""
    _cuda(pos, vel)
    return ops.prep(pos, vel, batch_size, num_nodes, lmax_attr)


@prep_lmax.register_fake
def _(pos, vel, batch_size, num_nodes, lmax_attr):
    nodes = batch_size * num_nodes
    return (pos.new_empty((nodes, 7), dtype=torch.float32),
            pos.new_empty((nodes, (lmax_attr + 1) ** 2), dtype=torch.float32))


@torch.library.custom_op(f"{NS}::edge_attr_list", mutates_args=())
def edge_attr_list(pos: Tensor, mass: Tensor, edge_index: Tensor, lmax_attr: int) -> Tuple[Tensor, Tensor]:
    """(edge_attr [E, (lmax_attr + 1)^2], additional_message_features [E, 2]) of an int64 edge list [2, E]."""
    _cuda(pos, mass, edge_index)
    return ops.edge_attr_list(pos, mass, edge_index, lmax_attr)


@edge_attr_list.register_fake
def _(pos, mass, edge_index, lmax_attr):
    E = edge_index.shape[1]
    return pos.new_empty((E, (lmax_attr + 1) ** 2), dtype=torch.float32), pos.new_empty((E, 2), dtype=torch.float32)


@torch.library.custom_op(f"{NS}::message_input_list", mutates_args=())
def message_input_list(x: Tensor, add: Tensor, edge_index: Tensor) -> Tensor:
    """[E, 2 D + d_add] = cat(x[target], x[source], add) for an int64 edge list [2, E] (row 0 = source)."""
    _cuda(x, add, edge_index)
    from ._lib import check, lib
    x, add, edge_index = x.float().contiguous(), add.float().contiguous(), edge_index.contiguous()
    E, D, d_add = edge_index.shape[1], x.shape[1], add.shape[1]
    out = torch.empty((E, 2 * D + d_add), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.segnn_generic_message_input_list(ops._p(x), ops._p(add), ops._p(edge_index), E, D, d_add, ops._p(out),
                                                   ops._stream()), "segnn_generic_message_input_list")
    ops._bump()
    return out


@message_input_list.register_fake
def _(x, add, edge_index):
    return x.new_empty((edge_index.shape[1], 2 * x.shape[1] + add.shape[1]), dtype=torch.float32)


@torch.library.custom_op(f"{NS}::segment_reduce", mutates_args=())
def segment_reduce(values: Tensor, order: Tensor, ptr: Tensor, mean: bool) -> Tensor:
    """out[node] = sum (mean) of values[order[ptr[node]:ptr[node + 1]]] in that order: deterministic scatter."""
    _cuda(values, order, ptr)
    return ops.segment_reduce(values, order, ptr, mean)


@segment_reduce.register_fake
def _(values, order, ptr, mean):
    return values.new_empty((ptr.shape[0] - 1, values.shape[1]), dtype=torch.float32)
