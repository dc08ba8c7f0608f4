"""Host-side mirror of models/segnn/segnn.py: same classes, constructor signature, parameter names and
``forward(graph) -> [nodes, 6]``; the arithmetic is the sm_100a kernel sequence K1..K6 behind the C ABI."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from . import ops, packing, training
from .graph import infer_graph_shape
from .irreps import Irreps, weight_balanced_irreps
from .o3_building_blocks import BatchNorm, InstanceNorm, O3TensorProduct, O3TensorProductSwishGate

def _pack_tp(kind: int, mod, n: int):
    """Operand blocks of one tensor product: from the C ABI (segnn_pack_weights) for device parameters; host-resident
    parameters (the CPU emulation tests of the kernel algebra) go through packing.py, the same layouts in torch."""
    w, b = mod.tp.weight, mod.biases
    if w.is_cuda:
        return ops.pack_weights(kind, n, w, b)
    f = lambda t: t.detach().to(torch.float32)
    if kind == ops.PACK_MSG1:
        return packing.pack_msg1(f(w), f(b), n)
    if kind == ops.PACK_MSG2:
        return packing.pack_msg2(f(w), f(b), n)
    if kind == ops.PACK_UPDATE1:
        return packing.pack_node_tp(f(w), f(b), 2, n, 2 * n)
    if kind == ops.PACK_UPDATE2:
        return packing.pack_node_tp(f(w), f(b), 1, n, n)
    if kind == ops.PACK_POOL1:
        return packing.pack_node_tp(f(w), f(b), 1, n, 2 * n)
    if kind == ops.PACK_EMBED:
        return packing.pack_embedding(f(w), f(b), n)
    return packing.pack_head(f(w), n)


def _fold_bn(bn, n: int, degree: float):
    if bn.weight.is_cuda:
        return ops.fold_batchnorm(bn.weight, bn.bias, bn.running_mean, bn.running_var, n, bn.eps, degree)
    f = lambda t: t.detach().to(torch.float32)
    return packing.fold_batchnorm(f(bn.weight), f(bn.bias), f(bn.running_mean), f(bn.running_var), n, bn.eps, degree)


_MODES = {"fp32": ops.MODE_FP32, "bf16": ops.MODE_BF16_TC, "fp16": ops.MODE_FP16_TC, "fp16p": ops.MODE_FP16_PACKED}
_TC_MODES = (ops.MODE_BF16_TC, ops.MODE_FP16_TC, ops.MODE_FP16_PACKED)
SMALL_BATCH_NODES = 2048  # below this the update / pool GEMMs of the tensor-core modes run on the FFMA kernel
_FP16_OPERAND_MODES = ("fp16", "fp16p")


class SEGNNLayer(nn.Module):
    """models/segnn/segnn.py:192-304 (PyG MessagePassing, aggr='add'). Holds the four tensor products and the two
    BatchNorms under the reference's names; ``forward`` runs node GEMM -> fused edge kernel -> node update."""

    def __init__(self, input_irreps, hidden_irreps, output_irreps, edge_attr_irreps, node_attr_irreps, norm=None,
                 additional_message_irreps=None):
        super().__init__()
        input_irreps, hidden_irreps = Irreps(str(input_irreps)), Irreps(str(hidden_irreps))
        self.hidden_irreps = hidden_irreps
        edge_attr_irreps, node_attr_irreps = Irreps(str(edge_attr_irreps)), Irreps(str(node_attr_irreps))
        add = Irreps(str(additional_message_irreps)) if additional_message_irreps is not None else Irreps()
        message_input_irreps = (2 * input_irreps + add).simplify()
        update_input_irreps = (input_irreps + hidden_irreps).simplify()
        self.message_layer_1 = O3TensorProductSwishGate(message_input_irreps, hidden_irreps, edge_attr_irreps)
        self.message_layer_2 = O3TensorProductSwishGate(hidden_irreps, hidden_irreps, edge_attr_irreps)
        self.update_layer_1 = O3TensorProductSwishGate(update_input_irreps, hidden_irreps, node_attr_irreps)
        self.update_layer_2 = O3TensorProduct(hidden_irreps, hidden_irreps, node_attr_irreps)
        self.norm = norm
        self.feature_norm = None
        self.message_norm = None
        if norm == "batch":
            self.feature_norm = BatchNorm(hidden_irreps)
            self.message_norm = BatchNorm(hidden_irreps)
        elif norm == "instance":  # segnn.py:236-237: per-graph normalisation of the node features, no message norm
            self.feature_norm = InstanceNorm(hidden_irreps)
        if list(add) != [(2, 0, 1)] or list(input_irreps) != list(hidden_irreps):
            raise NotImplementedError(f"SEGNN layers take hidden -> hidden irreps with 2x0e message features; got "
                                      f"input={input_irreps}, hidden={hidden_irreps}, additional={add}")
        # the fused kernels are specialised for n x0e + n x1o (lmax_h = 1); anything else runs the generic path
        # with the steering attributes 1x0e + 1x1o (lmax_attr = 1)
        self.fused = list(hidden_irreps) == [(hidden_irreps[0][0], 0, 1), (hidden_irreps[0][0], 1, -1)] and \
            list(edge_attr_irreps) == [(1, 0, 1), (1, 1, -1)] and list(node_attr_irreps) == [(1, 0, 1), (1, 1, -1)] \
            and norm != "instance"  # InstanceNorm depends on the features: it cannot be folded into the kernels
        self.n = hidden_irreps[0][0]

    # -- weight packing -----------------------------------------------------------------------------------------
    def pack(self, degree: int, eval_bn: bool = True, operand: int = 0) -> Dict[str, object]:
        n = self.n
        f = lambda t: t.detach().to(torch.float32)
        # inference: the operand blocks come from the C ABI itself (segnn_pack_weights); packing.py is its
        # differentiable twin, used by the training path
        pk_ = lambda kind, mod: _pack_tp(kind, mod, n)
        out = dict(msg1=pk_(ops.PACK_MSG1, self.message_layer_1), msg2=pk_(ops.PACK_MSG2, self.message_layer_2),
                   upd1=pk_(ops.PACK_UPDATE1, self.update_layer_1), upd2=pk_(ops.PACK_UPDATE2, self.update_layer_2),
                   bn_msg=(None, None), bn_feat=(None, None))
        if out["msg2"]["ss"].is_cuda and n in ops.TC_MULTIPLICITIES and ops.tc_available():
            out["msg2"]["tc"] = ops.pack_w2_tc(out["msg2"], n, operand)
            for key in ("upd1", "upd2"):
                out[key]["wt_s"] = ops.pack_node_weight_tc(out[key]["w_s"], operand)
                out[key]["wt_v"] = ops.pack_node_weight_tc(out[key]["w_v"], operand)
                out[key]["operand"] = operand
            # message_layer_1 projections for the tensor-core edge kernel: columns of each 3n block reordered to
            # (scalar part [n] | (gate, vector) pairs [n][2]) so the kernel reads both parts with one 64-bit load
            m1 = out["msg1"]
            j = torch.arange(2 * n, device=m1["w_s"].device)
            perm = torch.cat([torch.arange(n, device=j.device), n + (j % 2) * n + j // 2])
            perm6 = torch.cat([perm, 3 * n + perm])
            m1["wt_s"] = ops.pack_node_weight_tc(m1["w_s"][:, perm6].contiguous(), operand)
            m1["wt_v"] = ops.pack_node_weight_tc(m1["w_v"][:, perm6].contiguous(), operand)
            m1["operand"] = operand
            bias3 = torch.cat([m1["bias"], m1["bias"].new_zeros(n)])  # bias on the (0s, 0g) parts of P only
            m1["bias_tc"] = bias3[perm].contiguous()
            if operand == 1:
                # packed-half edge kernel (compute_mode 'fp16p'): the factor 1/2 of the (0s, 0g) pre-activations is
                # folded into the projections (the fp32-math kernels apply it on load)
                half = torch.ones(3 * n, device=j.device)
                half[:n] = 0.5
                half[n::2] = 0.5
                half6 = torch.cat([half, half])
                m1["wt_s_h"] = ops.pack_node_weight_tc((m1["w_s"][:, perm6] * half6).contiguous(), operand)
                m1["wt_v_h"] = ops.pack_node_weight_tc((m1["w_v"][:, perm6] * half6).contiguous(), operand)
                m1["bias_tc_h"] = (m1["bias_tc"] * half).contiguous()
        if eval_bn and self.message_norm is not None:
            out["bn_msg"] = _fold_bn(self.message_norm, n, float(degree))
            out["bn_feat"] = _fold_bn(self.feature_norm, n, 1.0)
        return out

    def run_x16(self, w, h, h16, pos, mass, node_attr, batch_size: int, num_nodes: int):
        """`run` in the packed-half mode with fp16 operand copies of the features: every tensor-core node GEMM reads
        16-bit rows its producer rounded (embed / combine / edge kernel), the fp32 features survive only as the
        residual.  Bit-identical to `run` (same roundings, applied by the producer instead of the GEMM loader).
        Returns (h_out fp32, h_out fp16)."""
        n = self.n
        m1, u1, u2 = w["msg1"], w["upd1"], w["upd2"]
        p, q = ops.node_gemm_pair16(h16, dict(wt_s=m1["wt_s_h"], wt_v=m1["wt_v_h"], operand=1), 6 * n,
                                    m1["bias_tc_h"], 3 * n, 3 * n)
        agg16 = ops.edge_layer(ops.MODE_FP16_PACKED, pos, mass, batch_size, num_nodes, n, p, q, m1["w_edge"],
                               w["msg2"], w["bn_msg"][0], w["bn_msg"][1], out16=True)
        g1_16 = ops.tp_combine(ops.node_gemm_out16(h16, agg16, u1, 3 * n), node_attr, n, True, bias=u1["bias"],
                               out16="only")
        return ops.tp_combine(ops.node_gemm_out16(g1_16, None, u2, 2 * n), node_attr, n, False, bias=u2["bias"],
                              residual=h, bn_mul=w["bn_feat"][0], bn_add=w["bn_feat"][1], out16="both")

    def run(self, w, mode: int, h, pos, mass, node_attr, batch_size: int, num_nodes: int):
        """One layer on planar features h [nodes,4,n] (eval-mode BatchNorm)."""
        n = self.n
        tc = mode in _TC_MODES
        m1 = w["msg1"]
        if mode == ops.MODE_FP16_PACKED:  # fp16 projections, nodes interleaved in pairs
            if num_nodes % 2:
                raise RuntimeError("compute_mode='fp16p' needs an even number of bodies per graph (sender pairs); "
                                   "use 'fp16' or 'bf16'")
            p, q = ops.node_gemm_pair16(h, dict(wt_s=m1["wt_s_h"], wt_v=m1["wt_v_h"], operand=1), 6 * n,
                                        m1["bias_tc_h"], 3 * n, 3 * n)
        elif tc:  # permuted column order + bias (see pack)
            p, q = ops.node_gemm(h, None, m1, 6 * n, bias=m1["bias_tc"], n_bias=3 * n, split=3 * n, tc=True)
        else:
            p, q = ops.node_gemm(h, None, m1, 6 * n, bias=m1["bias"], n_bias=2 * n, split=3 * n, tc=False)
        agg = ops.edge_layer(mode, pos, mass, batch_size, num_nodes, n, p, q, m1["w_edge"], w["msg2"],
                             w["bn_msg"][0], w["bn_msg"][1])
        u1 = w["upd1"]
        # tensor-core modes: the GEMM outputs that only feed an attribute-combine pass travel as fp16 rows; on a few
        # hundred nodes (BASELINE configuration 1: 500) the persistent tcgen05 kernel is all fixed cost (8 us against
        # ~4 us for the FFMA kernel), so small batches keep their update GEMMs on the fp32 kernel
        tc_upd = tc and h.shape[0] >= SMALL_BATCH_NODES
        y1 = ops.node_gemm_out16(h, agg, u1, 3 * n) if tc_upd else ops.node_gemm(h, agg, u1, 3 * n)
        g1 = ops.tp_combine(y1, node_attr, n, True, bias=u1["bias"])
        u2 = w["upd2"]
        y2 = ops.node_gemm_out16(g1, None, u2, 2 * n) if tc_upd else ops.node_gemm(g1, None, u2, 2 * n)
        return ops.tp_combine(y2, node_attr, n, False, bias=u2["bias"], residual=h, bn_mul=w["bn_feat"][0],
                              bn_add=w["bn_feat"][1])

    def run_batch_statistics(self, w, h, pos, mass, node_attr, batch_size: int, num_nodes: int,
                             update_running_stats: bool = True):
        """The layer with TRAIN-mode BatchNorm (batch statistics) on the tensor-core kernels, inference only: the
        reference evaluates its rollouts with the module in training mode (trainer.py:373, 929-942), so this is what a
        faithful `run_self_feed` runs. K3 (packed-half mode) returns the raw sums and the per-receiver moments of the
        messages; statistics are deterministic float64 column sums, the affine is one `segnn_lincomb` pass."""
        n = self.n
        nodes, deg = batch_size * num_nodes, num_nodes - 1
        m1 = w["msg1"]
        p, q = ops.node_gemm_pair16(h, dict(wt_s=m1["wt_s_h"], wt_v=m1["wt_v_h"], operand=1), 6 * n, m1["bias_tc_h"],
                                    3 * n, 3 * n)
        agg, mom = ops.edge_layer(ops.MODE_FP16_PACKED, pos, mass, batch_size, num_nodes, n, p, q, m1["w_edge"],
                                  w["msg2"], None, None, want_moments=True)

        def bn_dict(bn):
            return dict(weight=bn.weight.detach().float(), bias=bn.bias.detach().float(), running_mean=bn.running_mean,
                        running_var=bn.running_var, eps=bn.eps, momentum=bn.momentum)
        if self.message_norm is not None:
            flat = agg.view(nodes, 4 * n)
            st = ops.bn_forward_coeffs(bn_dict(self.message_norm), n, float(nodes * deg), float(deg), ops.colsum(flat),
                                       ops.colsum(mom), 1, True, update_running_stats)
            agg = ops.lincomb(flat, None, st["mulcols"], None, st["addcols"]).view(nodes, 4, n)
        u1, u2 = w["upd1"], w["upd2"]
        g1 = ops.tp_combine(ops.node_gemm_out16(h, agg, u1, 3 * n), node_attr, n, True, bias=u1["bias"])
        pre = ops.tp_combine(ops.node_gemm_out16(g1, None, u2, 2 * n), node_attr, n, False, bias=u2["bias"], residual=h)
        if self.feature_norm is None:
            return pre
        flat = pre.view(nodes, 4 * n)
        st = ops.bn_forward_coeffs(bn_dict(self.feature_norm), n, float(nodes), 1.0, ops.colsum(flat),
                                   ops.colsum(flat, None, 1), 3, True, update_running_stats)
        return ops.lincomb(flat, None, st["mulcols"], None, st["addcols"]).view(nodes, 4, n)

    def forward(self, x, edge_index, edge_attr, node_attr, batch, additional_message_features=None, *, pos=None,
                mass=None, num_graphs=None, n_nodes=None, mode="fp32"):
        """Reference signature (segnn.py:239-247) on e3nn-layout features. The graph is implicit: ``edge_index`` /
        ``edge_attr`` / ``additional_message_features`` are ignored and recomputed from ``pos`` / ``mass``.

        Called the way the reference calls it (no ``pos`` / ``mass``: explicit ``edge_index``, ``edge_attr`` and
        ``additional_message_features``), the layer runs the reference's own formulation on the generic-irreps kernels:
        any hidden irreps with l <= 2, any edge list, the attributes as given (inference)."""
        if self.training and self.norm == "batch":
            raise NotImplementedError("train-mode BatchNorm for the standalone layer: use SEGNN.forward")
        if pos is None or mass is None or num_graphs is None or n_nodes is None:
            if edge_index is None or edge_attr is None:
                raise ValueError("SEGNNLayer.forward needs either pos, mass, num_graphs, n_nodes (implicit complete "
                                 "graph) or the explicit edge_index, edge_attr (+ additional_message_features)")
            return self._forward_edge_list(x, edge_index, edge_attr, node_attr, batch, additional_message_features)
        if not self.fused:
            raise NotImplementedError("the implicit-graph form of the standalone SEGNNLayer.forward is built for "
                                      "lmax_h = 1; pass edge_index / edge_attr explicitly, or use SEGNN.forward")
        w = self.pack(n_nodes - 1, operand=1 if mode in _FP16_OPERAND_MODES else 0)
        h = packing.to_planar(x.float(), self.n)
        out = self.run(w, _MODES[mode], h, pos.float().contiguous(), mass.float().reshape(-1).contiguous(),
                       node_attr.float().contiguous(), num_graphs, n_nodes)
        return packing.from_planar(out).to(x.dtype)


    @torch.no_grad()
    def _forward_edge_list(self, x, edge_index, edge_attr, node_attr, batch, add):
        """segnn.py:239-304 as written there: gather (x_i = x[target], x_j = x[source]), message_layer_1 / 2 with their
        gates, eval BatchNorm per edge, deterministic sum over the incoming edges, update, residual, feature norm."""
        from .generic import GatePlan, TensorProductPlan, _bn_eval_columns
        from ._lib import check, lib
        if not x.is_cuda:
            raise RuntimeError("SEGNNLayer (B200) needs CUDA tensors: there is no CPU fallback")
        dev, dtype = x.device, x.dtype
        plans = self.__dict__.get("_plans")
        if plans is None or plans["dev"] != dev:
            tp = lambda m: TensorProductPlan(m, dev)
            plans = dict(dev=dev, msg1=tp(self.message_layer_1), g_msg1=GatePlan(self.message_layer_1, dev),
                         msg2=tp(self.message_layer_2), g_msg2=GatePlan(self.message_layer_2, dev),
                         upd1=tp(self.update_layer_1), g_upd1=GatePlan(self.update_layer_1, dev),
                         upd2=tp(self.update_layer_2))
            self.__dict__["_plans"] = plans
        f32 = lambda t: t.to(torch.float32).contiguous()
        x, ea, attr = f32(x), f32(edge_attr), f32(node_attr)
        ei = edge_index.to(device=dev, dtype=torch.int64).contiguous()
        E, D = ei.shape[1], x.shape[1]
        d_add = 0 if add is None else add.shape[1]
        if 2 * D + d_add != self.message_layer_1.irreps_in1.dim:
            raise ValueError(f"message input has {2 * D + d_add} columns, message_layer_1 takes "
                             f"{self.message_layer_1.irreps_in1.dim}")
        addf = f32(add) if add is not None else None
        inp = torch.empty((E, 2 * D + d_add), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.segnn_generic_message_input_list(ops._p(x), ops._p(addf), ops._p(ei), E, D, d_add, ops._p(inp),
                                                       ops._stream()), "segnn_generic_message_input_list")
        ops._bump()
        m = plans["g_msg1"].run(plans["msg1"].run(inp, ea))
        m = plans["g_msg2"].run(plans["msg2"].run(m, ea))
        if self.message_norm is not None:
            mul, addc = _bn_eval_columns(self.message_norm, self.hidden_irreps)
            m = ops.lincomb(m, None, mul, None, addc)
        order, ptr = ops.edge_list_csr(ei, x.shape[0])
        agg = ops.segment_reduce(m, order, ptr)
        u = plans["g_upd1"].run(plans["upd1"].run(torch.cat([x, agg], dim=1).contiguous(), attr))
        u = plans["upd2"].run(u, attr)
        out = ops.add3(x, u)
        if self.feature_norm is not None:
            if self.norm == "instance":
                out = self.feature_norm(out, batch)
            else:
                mul, addc = _bn_eval_columns(self.feature_norm, self.hidden_irreps)
                out = ops.lincomb(out, None, mul, None, addc)
        return out.to(dtype)


class SEGNN(nn.Module):
    """Steerable E(3) equivariant message passing network -- models/segnn/segnn.py:14-189 (task='node').

    ``compute_mode``: 'fp32' (FFMA kernels, the 1e-5 parity mode), 'bf16' (tcgen05 tensor-core kernels, the mode the
    benchmark reports), 'fp16' (the same kernels with fp16 operands: 8x smaller operand rounding at the same speed) or
    'fp16p' (fp16 operands and the message_layer_1 combine in packed fp16 on fp16 projections; even N only)."""

    def __init__(self, input_irreps="2x1o + 1x0e", hidden_features=64, lmax_h=1, lmax_attr=1, num_layers=4,
                 output_irreps="2x1o", norm="batch", pool="avg", task="node", additional_message_irreps="2x0e",
                 training_args=None, compute_mode="fp32"):
        super().__init__()
        if task != "node":
            raise NotImplementedError("only task='node' is on the N-body path")
        if int(lmax_attr) not in (1, 2) or int(lmax_h) not in (1, 2):
            raise NotImplementedError(f"built for lmax_attr in (1, 2) and lmax_h in (1, 2) (got {lmax_h}, {lmax_attr})")
        if Irreps(str(input_irreps)) != Irreps("2x1o+1x0e") or Irreps(str(output_irreps)) != Irreps("2x1o"):
            raise NotImplementedError("kernels are built for the N-body irreps: input 2x1o+1x0e, output 2x1o")
        self.hidden_features, self.lmax_h, self.lmax_attr, self.num_layers = hidden_features, lmax_h, lmax_attr, num_layers
        self.node_attr_irreps = Irreps.spherical_harmonics(lmax_attr)
        self.edge_attr_irreps = Irreps.spherical_harmonics(lmax_attr)
        self.hidden_irreps = weight_balanced_irreps(hidden_features, self.node_attr_irreps, lmax=lmax_h)
        self.task, self.norm, self.pool, self.training_args = task, norm, pool, training_args
        self.additional_message_irreps = Irreps(str(additional_message_irreps))
        self.compute_mode = compute_mode
        self._pack_map = None  # (key, gather tables) of the parameter -> operand-block re-layout (training path)
        h = self.hidden_irreps
        self.n = h[0][0]
        self.embedding_layer = O3TensorProduct(Irreps(str(input_irreps)), h, self.node_attr_irreps)
        self.layers = nn.ModuleList([
            SEGNNLayer(h, h, h, self.edge_attr_irreps, self.node_attr_irreps, norm=norm,
                       additional_message_irreps=self.additional_message_irreps) for _ in range(num_layers)])
        self.pre_pool1 = O3TensorProductSwishGate(h, h, self.node_attr_irreps)
        self.pre_pool2 = O3TensorProduct(h, Irreps(str(output_irreps)), self.node_attr_irreps)
        self._pack_key = None
        self._packed = None
        # lmax_h = 1 with lmax_attr = 1 runs the fused kernels (compute_mode 'fp32' / 'bf16'); other hidden irreps
        # (lmax_h = 2, BASELINE config 3) and lmax_attr = 2 run the generic-irreps fp32 path (generic.py), which is also
        # selectable with compute_mode='generic'
        self.fused = all(layer.fused for layer in self.layers)
        self._generic = None

    def get_model_size(self):
        return self.hidden_features

    def get_serializable_attributes(self):
        return {
            "hidden_features": self.hidden_features, "lmax_h": self.lmax_h, "lmax_attr": self.lmax_attr,
            "node_attr_irreps": str(self.node_attr_irreps), "num_layers": self.num_layers,
            "input_irreps": str(self.embedding_layer.irreps_in1), "hidden_irreps": str(self.hidden_irreps),
            "output_irreps": str(self.pre_pool2.irreps_out), "edge_attr_irreps": str(self.edge_attr_irreps),
            "norm": self.norm, "pool": None, "task": self.task,
            "additional_message_irreps": str(self.additional_message_irreps), "training_args": self.training_args,
            "num_params": sum(p.numel() for p in self.parameters()),
        }

    def load_state_dict(self, state_dict, strict=True, assign=False):
        """Accepts reference checkpoints: e3nn-internal buffers (tp.output_mask, compiled w3j constants, empty
        gate.mul.weight, ...) are filtered out; the learnable tensors keep the reference's names."""
        own = set(self.state_dict().keys())
        filtered = {k: v for k, v in state_dict.items() if k in own}
        return super().load_state_dict(filtered, strict=strict, assign=assign)

    # -- packing cache ------------------------------------------------------------------------------------------
    def packed(self, degree: int):
        operand = 1 if self.compute_mode in _FP16_OPERAND_MODES else 0  # 16-bit operand format of the tensor-core images
        tensors = list(self.parameters()) + list(self.buffers())
        key = (degree, operand, tuple((t.data_ptr(), t._version, t.device, t.dtype) for t in tensors))
        if key != self._pack_key:
            f = lambda t: t.detach().to(torch.float32)
            n = self.n
            self._packed = dict(
                embed=_pack_tp(ops.PACK_EMBED, self.embedding_layer, n),
                layers=[layer.pack(degree, operand=operand) for layer in self.layers],
                pool1=_pack_tp(ops.PACK_POOL1, self.pre_pool1, n),
                head=_pack_tp(ops.PACK_HEAD, self.pre_pool2, n))
            if "tc" in self._packed["layers"][0]["msg2"]:
                for name in ("w_s", "w_v"):  # (must not reuse `key`: it is stored as the cache key below)
                    self._packed["pool1"]["wt" + name[1:]] = ops.pack_node_weight_tc(self._packed["pool1"][name], operand)
                self._packed["pool1"]["operand"] = operand
            self._pack_key = key
        return self._packed

    # -- forward ------------------------------------------------------------------------------------------------
    def forward_state(self, pos, vel, mass, batch_size: int, num_nodes: int, x_in=None, node_attr=None,
                      return_layers: bool = False):
        """pos, vel [nodes,3] fp32 CUDA, mass [nodes] -> pred [nodes,6] fp32 (eval-mode BatchNorm)."""
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        bn_training = self.training and self.norm == "batch"
        if not self.fused or self.compute_mode == "generic":
            if needs_grad or bn_training:
                raise NotImplementedError("the generic-irreps path (lmax_h != 1 or compute_mode='generic') is inference "
                                          "only: call model.eval() under torch.no_grad()")
            from .generic import GenericRunner
            if self._generic is None or self._generic.embed.instr.device != pos.device:
                self._generic = GenericRunner(self, pos.device)
            return self._generic.forward(pos, vel, mass, batch_size, num_nodes, return_layers, x_in, node_attr)
        mode = _MODES[self.compute_mode]
        batch_stats_tc = (bn_training and not needs_grad and mode == ops.MODE_FP16_PACKED and num_nodes % 2 == 0
                          and self.n in ops.TC_MULTIPLICITIES)
        if (needs_grad or bn_training) and not batch_stats_tc:
            return self._forward_train(pos, vel, mass, batch_size, num_nodes, bn_training, needs_grad, return_layers)
        w = self.packed(num_nodes - 1)
        n = self.n
        if mode in _TC_MODES and "tc" not in w["layers"][0]["msg2"]:
            raise RuntimeError(f"compute_mode='bf16'/'fp16' needs hidden multiplicity in {ops.TC_MULTIPLICITIES} "
                               f"(hidden_features 64/128/192); this model has n={n}")
        if x_in is None or node_attr is None:
            x_in, node_attr = ops.prep(pos, vel, batch_size, num_nodes)
        # packed-half mode on batches large enough for the tensor-core update GEMMs: fp16 operand copies of the features
        x16 = (ops.X16_FEATURES and mode == ops.MODE_FP16_PACKED and not batch_stats_tc and num_nodes % 2 == 0
               and n % 2 == 0 and x_in.shape[0] >= SMALL_BATCH_NODES)
        h16 = None
        if x16:
            h, h16 = ops.embed(x_in, node_attr, w["embed"]["w"], w["embed"]["bias"], n, want16=True)
        else:
            h = ops.embed(x_in, node_attr, w["embed"]["w"], w["embed"]["bias"], n)
        per_layer = [h]
        for layer, lw in zip(self.layers, w["layers"]):
            if batch_stats_tc:  # train-mode BatchNorm without gradients: tensor-core kernels + batch statistics
                h = layer.run_batch_statistics(lw, h, pos, mass, node_attr, batch_size, num_nodes)
            elif x16:
                h, h16 = layer.run_x16(lw, h, h16, pos, mass, node_attr, batch_size, num_nodes)
            else:
                h = layer.run(lw, mode, h, pos, mass, node_attr, batch_size, num_nodes)
            per_layer.append(h)
        p1 = w["pool1"]
        y = ops.node_gemm_out16(h16 if x16 else h, None, p1, 3 * n) \
            if mode in _TC_MODES and h.shape[0] >= SMALL_BATCH_NODES else ops.node_gemm(h, None, p1, 3 * n)
        hp = ops.tp_combine(y, node_attr, n, True, bias=p1["bias"])
        pred = ops.head(hp, node_attr, w["head"], n)
        if return_layers:
            return pred, per_layer
        return pred

    # -- training path: train-mode BatchNorm + hand-written backward (fp32 kernels) -----------------------------
    def packed_train(self, dtype=torch.float32, transposed: bool = False):
        """Differentiable re-layout of the parameters into kernel operand blocks (no caching: the parameters
        change every optimizer step) + the BatchNorm buffers that ride along outside autograd. ``transposed`` adds
        the transposed weight blocks the backward kernels read (keys ``*_t``: derived leaves without gradient)."""
        n = self.n
        f = lambda t: t.to(dtype)

        def with_t(d, keys):
            if transposed:
                for k in keys:
                    d[k + "_t"] = d[k].t().contiguous()
            return d
        layers, bufs = [], []
        for layer in self.layers:
            lw = dict(
                msg1=with_t(packing.pack_msg1(f(layer.message_layer_1.tp.weight), f(layer.message_layer_1.biases), n),
                            ("w_s", "w_v")),
                msg2=with_t(packing.pack_msg2(f(layer.message_layer_2.tp.weight), f(layer.message_layer_2.biases), n),
                            ("ss", "vs", "sv", "vv")),
                upd1=with_t(packing.pack_node_tp(f(layer.update_layer_1.tp.weight), f(layer.update_layer_1.biases), 2,
                                                 n, 2 * n), ("w_s", "w_v")),
                upd2=with_t(packing.pack_node_tp(f(layer.update_layer_2.tp.weight), f(layer.update_layer_2.biases), 1,
                                                 n, n), ("w_s", "w_v")),
                bn_msg=None, bn_feat=None)
            lb = dict(bn_msg=None, bn_feat=None)
            for key, bn in (("bn_msg", layer.message_norm), ("bn_feat", layer.feature_norm)):
                if bn is not None:
                    lw[key] = dict(weight=f(bn.weight), bias=f(bn.bias))
                    lb[key] = dict(running_mean=bn.running_mean, running_var=bn.running_var, eps=bn.eps,
                                   momentum=bn.momentum)
            layers.append(lw)
            bufs.append(lb)
        tree = dict(embed=packing.pack_embedding(f(self.embedding_layer.tp.weight), f(self.embedding_layer.biases), n),
                    layers=layers,
                    pool1=with_t(packing.pack_node_tp(f(self.pre_pool1.tp.weight), f(self.pre_pool1.biases), 1, n,
                                                      2 * n), ("w_s", "w_v")),
                    head=packing.pack_head(f(self.pre_pool2.tp.weight), n))
        return tree, bufs

    # -- flat parameter / gradient storage (training plumbing) ----------------------------------------------------
    def use_flat_storage(self):
        """Re-points every parameter at a slice of ONE flat fp32 buffer and every ``.grad`` at the matching slice of a
        second one.  The training function then reads the flat buffer directly (no per-step ``cat`` of ~100 tensors),
        writes the whole gradient with one scatter-add into the gradient buffer (no per-parameter clone by autograd's
        AccumulateGrad) and a data-parallel step all-reduces that buffer in place (no flatten / unflatten).  Opt-in
        (``TrainStep`` does it): gradients are OVERWRITTEN by every backward, not accumulated.  Call after the module
        sits on its device in float32; ``.to()`` / ``.float()`` afterwards silently drops back to the generic path."""
        params = list(self.parameters())
        if any(p.dtype != torch.float32 for p in params) or len({p.device for p in params}) != 1:
            raise RuntimeError("flat storage needs float32 parameters on one device")
        total = sum(p.numel() for p in params)
        flat = torch.empty(total, dtype=torch.float32, device=params[0].device)
        sink = torch.zeros(total, dtype=torch.float32, device=params[0].device)
        off = 0
        with torch.no_grad():
            for p in params:
                k = p.numel()
                flat[off:off + k].copy_(p.data.reshape(-1))
                p.data = flat[off:off + k].view(p.shape)
                p.grad = sink[off:off + k].view(p.shape)
                off += k
        self._flat = (flat, sink)
        return flat, sink

    def _flat_storage(self, params):
        """(flat parameters, gradient sink) when every parameter / gradient still aliases its slice, else (None, None)."""
        flat_sink = getattr(self, "_flat", None)
        if flat_sink is None:
            return None, None
        flat, sink = flat_sink
        off, p0, g0 = 0, flat.data_ptr(), sink.data_ptr()
        for p in params:
            if p.data_ptr() != p0 + 4 * off or p.grad is None or p.grad.data_ptr() != g0 + 4 * off \
                    or p.dtype != torch.float32:
                return None, None
            off += p.numel()
        return flat, sink

    def bn_buffers(self):
        """The BatchNorm running statistics as they are NOW (never cached: dtype casts and load_state_dict(assign=True)
        replace the buffer objects)."""
        out = []
        for layer in self.layers:
            lb = dict(bn_msg=None, bn_feat=None)
            for key, bn in (("bn_msg", layer.message_norm), ("bn_feat", layer.feature_norm)):
                if bn is not None:
                    lb[key] = dict(running_mean=bn.running_mean, running_var=bn.running_var, eps=bn.eps,
                                   momentum=bn.momentum)
            out.append(lb)
        return out

    def _forward_train(self, pos, vel, mass, batch_size, num_nodes, bn_training, needs_grad, return_layers=False,
                       backend=None, dtype=torch.float32):
        if needs_grad and not return_layers:
            # parameters go in as they are: packing is one gather inside the function (training.build_pack_map)
            params = list(self.parameters())
            key = tuple((tuple(p.shape), p.device) for p in params)  # the map depends on shapes only
            if self._pack_map is None or self._pack_map[0] != key:
                self._pack_map = (key, training.build_pack_map(self))
            flat, sink = self._flat_storage(params) if dtype == torch.float32 else (None, None)
            cfg = dict(pack_map=self._pack_map[1], dtype=dtype, n=self.n, B=batch_size, N=num_nodes,
                       bn_training=bn_training, backend=backend, bn_buffers=self.bn_buffers(), flat_params=flat,
                       grad_sink=sink)
            return training.SegnnTrainFunctionFlat.apply(cfg, pos, vel, mass, *params)
        tree, bufs = self.packed_train(dtype)
        leaves, spec = training.flatten_packed(tree)
        with torch.no_grad():
            W = training.unflatten_packed([t.detach() for t in leaves], spec)
            training.attach_bn_buffers(W, bufs)
            pred, saved = training.forward_train(W, self.n, pos, vel, mass, batch_size, num_nodes, bn_training,
                                                 backend=backend)
        if return_layers:
            return pred, [rec["h"] for rec in saved["layers"]] + [saved["h_last"]]
        return pred

    def forward_edge_list(self, pos, vel, mass, edge_index, batch_size: int, num_nodes: int,
                          return_layers: bool = False, x_in=None, node_attr=None):
        """pos, vel [nodes,3], mass [nodes], edge_index int64 [2,E] (row 0 = source, row 1 = target), batch_size graphs
        of num_nodes consecutive nodes -> pred [nodes,6]; inference only (eval-mode BatchNorm, no gradients), fp32, any
        configuration the generic kernels cover."""
        if (torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())) or \
                (self.training and self.norm == "batch"):
            raise NotImplementedError("message passing on explicit edge lists (kNN graphs) is inference only: call "
                                      "model.eval() under torch.no_grad(); training runs on the complete graph "
                                      "(num_neighbors = N - 1)")
        from .generic import GenericRunner
        if self._generic is None or self._generic.embed.instr.device != pos.device:
            self._generic = GenericRunner(self, pos.device)
        return self._generic.forward_edge_list(pos, vel, mass.reshape(-1), edge_index, int(batch_size), int(num_nodes),
                                               return_layers, x_in, node_attr)

    def forward(self, graph, return_layers: bool = False):
        """SEGNN forward pass on a batched implicit graph (reference: segnn.py:150-189)."""
        b, n_nodes = infer_graph_shape(graph)
        dtype = graph.pos.dtype
        dev = graph.pos.device
        if dev.type != "cuda":
            raise RuntimeError("SEGNN (B200) needs CUDA tensors: there is no CPU fallback")
        explicit = getattr(graph, "__dict__", {}).get("edge_index")
        f32 = lambda t: t.to(torch.float32).contiguous()
        pos, vel, mass = f32(graph.pos), f32(graph.vel), f32(graph.mass).reshape(-1)
        x_in = node_attr = None
        if getattr(graph, "x", None) is not None and getattr(graph, "node_attr", None) is not None:
            x_in, node_attr = f32(graph.x), f32(graph.node_attr).clone()
            node_attr[:, 0] = 1.0  # catch_isolated_nodes, segnn.py:148
        if torch.is_tensor(explicit) and explicit.shape[1] != b * n_nodes * (n_nodes - 1):
            # an explicit edge list that is not the complete graph (kNN, num_neighbors < N - 1): generic-irreps kernels
            # with gathers through edge_index; the edge geometry is recomputed from pos / mass
            out = self.forward_edge_list(pos, vel, mass, explicit, b, n_nodes, return_layers, x_in, node_attr)
            if return_layers:
                return out[0].to(dtype), [h.to(dtype) for h in out[1]]
            return out.to(dtype)
        out = self.forward_state(pos, vel, mass, b, n_nodes, x_in, node_attr, return_layers)
        if return_layers:
            unpack = packing.from_planar if (self.fused and self.compute_mode != "generic") else (lambda h: h)
            return out[0].to(dtype), [unpack(h).to(dtype) for h in out[1]]
        return out.to(dtype)
