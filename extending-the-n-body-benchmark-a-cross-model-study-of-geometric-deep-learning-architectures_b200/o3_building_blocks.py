"""Host-side mirror of models/segnn/o3_building_blocks.py: same class names, constructor arguments, parameter
names (``tp.weight``, ``biases``) and initialisation, with the arithmetic running in the sm_100a kernels."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import ops, packing
from .graph import infer_graph_shape
from .irreps import Irreps, tp_instructions


class _TensorProductWeights(nn.Module):
    """Stands where e3nn's FullyConnectedTensorProduct sits in the reference (attribute ``tp``): owns the flat
    ``weight`` parameter in e3nn's instruction order."""

    def __init__(self, irreps_in1, irreps_in2, irreps_out):
        super().__init__()
        self.irreps_in1, self.irreps_in2, self.irreps_out = Irreps(irreps_in1), Irreps(irreps_in2), Irreps(irreps_out)
        self.instructions, self.weight_numel = tp_instructions(self.irreps_in1, self.irreps_in2, self.irreps_out)
        self.weight = nn.Parameter(torch.empty(self.weight_numel))

    def weight_views(self):
        return [self.weight[i["offset"]: i["offset"] + math.prod(i["shape"])].view(i["shape"])
                for i in self.instructions]


class O3TensorProduct(nn.Module):
    """models/segnn/o3_building_blocks.py:10-167. Parameters: ``tp.weight`` (flat), ``biases`` (one per l=0 output).
    Initialisation follows tensor_product_init (:79-116): U(-1/sqrt(fan_in), 1/sqrt(fan_in)) per output slice."""

    def __init__(self, irreps_in1, irreps_out, irreps_in2=None, tp_rescale=True) -> None:
        super().__init__()
        if not tp_rescale:
            raise NotImplementedError("tp_rescale=False is never used by the SEGNN path")
        self.irreps_in1, self.irreps_out = Irreps(str(irreps_in1)), Irreps(str(irreps_out))
        self.irreps_in2_provided = irreps_in2 is not None
        self.irreps_in2 = Irreps(str(irreps_in2)) if irreps_in2 is not None else Irreps("1x0e")
        self.tp_rescale = tp_rescale
        self.tp = _TensorProductWeights(self.irreps_in1, self.irreps_in2, self.irreps_out)
        fan = {}
        for ins in self.tp.instructions:
            fan[ins["io"]] = fan.get(ins["io"], 0) + ins["shape"][0] * ins["shape"][1]
        self.slices_fan_in = fan
        with torch.no_grad():
            for ins, view in zip(self.tp.instructions, self.tp.weight_views()):
                k = 1.0 / math.sqrt(fan[ins["io"]])
                view.uniform_(-k, k)
            biases = []
            for io, (mul, l, _) in enumerate(self.irreps_out):
                if l == 0:
                    k = 1.0 / math.sqrt(fan[io])
                    biases.append(torch.empty(mul).uniform_(-k, k))
        self.biases = nn.Parameter(torch.cat(biases)) if biases else None

    # -- module-level forward: used for API parity, not by SEGNN.forward --------------------------------------
    def _is_hidden_shaped(self):
        n_out = self.irreps_out[-1][0]
        n_in = self.irreps_in1[0][0]
        n_blocks = len(self.irreps_in1) // 2
        hid = [(n_in, 0, 1), (n_in, 1, -1)] * n_blocks
        n0 = self.irreps_out[0][0]
        ok = (len(self.irreps_out) == 2 and list(self.irreps_in1) == hid
              and list(self.irreps_out) == [(n0, 0, 1), (n_out, 1, -1)]
              and list(self.irreps_in2) == [(1, 0, 1), (1, 1, -1)] and n_in == n_out and 1 <= n_blocks <= 2
              and self.biases is not None)
        return ok, n_blocks, n_in, n0, n_out

    def _node_forward(self, data_in1, data_in2, gate: bool):
        """o3_building_blocks.py:150-167 (+ the Gate of :197-203 when ``gate``).  Hidden -> hidden products on node
        rows run the fused node kernels (GEMM + attribute combine); every other combination of irreps (l <= 2, steering
        attribute 1x0e or 1x0e+1x1o) runs the generic tensor-product kernels of generic.py."""
        if not data_in1.is_cuda:
            raise RuntimeError("O3TensorProduct (B200) needs CUDA tensors: there is no CPU fallback")
        ok, n_blocks, n_in, n0, n_out = self._is_hidden_shaped()
        if ok:
            w = packing.pack_node_tp(self.tp.weight.detach().float(), self.biases.detach().float(), n_blocks, n_in, n0)
            x = data_in1.float()
            xs = [packing.to_planar(x[:, b * 4 * n_in:(b + 1) * 4 * n_in], n_in) for b in range(n_blocks)]
            attr = data_in2.float().contiguous()
            y = ops.node_gemm(xs[0], xs[1] if n_blocks == 2 else None, w, n0 + n_out)
            out = ops.tp_combine(y, attr, n_out, gate, bias=w["bias"])
            return packing.from_planar(out).to(data_in1.dtype)
        from .generic import GatePlan, TensorProductPlan
        dev = data_in1.device
        if getattr(self, "_plan_device", None) != dev:
            self._plan = TensorProductPlan(self, dev)
            self._gate_plan = GatePlan(self, dev) if gate else None
            self._plan_device = dev
        x1 = data_in1.float().contiguous()
        x2 = torch.ones_like(x1[:, 0:1]) if data_in2 is None else data_in2.float().contiguous()  # :151-152
        out = self._plan.run(x1, x2)
        if gate:
            out = self._gate_plan.run(out)
            if self._gate_plan.n_g == 0:  # no gated irreps: the reference applies a plain nn.SiLU (:194-195), i.e.
                out = out * (1.0 / ops.C_SILU)  # without e3nn's normalize2mom constant the gate kernel folds in
        return out.to(data_in1.dtype)

    def forward(self, data_in1, data_in2=None) -> torch.Tensor:
        return self._node_forward(data_in1, data_in2, gate=False)


class O3TensorProductSwishGate(O3TensorProduct):
    """models/segnn/o3_building_blocks.py:170-203: the TP output carries one extra scalar per gated irrep; e3nn Gate
    (SiLU on scalars, sigmoid gates, both normalize2mom-scaled) is fused into the kernels' epilogues."""

    def __init__(self, irreps_in1, irreps_out, irreps_in2=None) -> None:
        irreps_out = Irreps(str(irreps_out))
        scalars = Irreps([irreps_out[0]])
        gated = irreps_out[1:]
        gates = Irreps([(gated.num_irreps, 0, 1)])
        irreps_g = (scalars + gates + gated).simplify()
        super().__init__(irreps_in1, irreps_g, irreps_in2)
        self.irreps_gated_out = irreps_out

    def forward(self, data_in1, data_in2=None) -> torch.Tensor:
        return self._node_forward(data_in1, data_in2, gate=True)


class BatchNorm(nn.Module):
    """Parameter/buffer holder with e3nn.nn.BatchNorm's names and shapes (models/segnn/segnn.py:233-235):
    weight [num_irreps], bias [n_scalar], running_mean [n_scalar], running_var [num_irreps]."""

    def __init__(self, irreps, eps=1e-5, momentum=0.1):
        super().__init__()
        self.irreps = Irreps(str(irreps))
        self.eps, self.momentum = eps, momentum
        n_scalar = sum(m for m, l, p in self.irreps if l == 0 and p == 1)
        n_feat = self.irreps.num_irreps
        self.register_buffer("running_mean", torch.zeros(n_scalar))
        self.register_buffer("running_var", torch.ones(n_feat))
        self.weight = nn.Parameter(torch.ones(n_feat))
        self.bias = nn.Parameter(torch.zeros(n_scalar))


class O3Transform:
    """models/segnn/o3_building_blocks.py:225-278. Runs K1 (node_attr, x) and leaves the per-edge quantities
    implicit; they are recomputed in registers inside the fused edge kernel."""

    def __init__(self, lmax_attr, use_force_input=False):
        if int(lmax_attr) not in (0, 1, 2):
            raise NotImplementedError("O3Transform is built for lmax_attr <= 2")
        self.attr_irreps = Irreps.spherical_harmonics(int(lmax_attr))
        self.use_force_input = use_force_input

    def __call__(self, graph):
        b, n = infer_graph_shape(graph)
        graph.num_graphs, graph.n_nodes = b, n
        graph.lmax_attr = lmax = self.attr_irreps.lmax  # degree of the lazily materialised edge attributes
        explicit = getattr(graph, "__dict__", {}).get("edge_index")
        if torch.is_tensor(explicit) and explicit.shape[1] != b * n * (n - 1):
            # kNN edge list (num_neighbors < N - 1): everything is materialised, as in the reference
            ei = explicit.to(torch.int64).contiguous()
            ea, add = ops.edge_attr_list(graph.pos, graph.mass, ei, lmax)
            order, ptr = ops.edge_list_csr(ei, graph.pos.shape[0])
            x, attr = ops.prep_list(graph.pos, graph.vel, ea, order, ptr, lmax)
            graph.edge_attr = ea.to(graph.pos.dtype)
            graph.additional_message_features = add.to(graph.pos.dtype)
        else:
            x, attr = ops.prep(graph.pos, graph.vel, b, n, lmax)
        if self.use_force_input:  # :267-271: node_attr += Y(force)
            attr = ops.add_vector_harmonics(attr, graph.force, lmax)
        graph.x = x.to(graph.pos.dtype)
        graph.node_attr = attr.to(graph.pos.dtype)
        return graph


class InstanceNorm(nn.Module):
    """models/segnn/instance_norm.py:8-129: per-graph normalisation of every irrep channel (mean removed for l = 0,
    'component' norm, mean over the graph's nodes), affine weight per irrep and bias per scalar."""

    def __init__(self, irreps, eps=1e-5, affine=True, reduce="mean", normalization="component"):
        super().__init__()
        if reduce != "mean" or normalization != "component":
            raise NotImplementedError("InstanceNorm is built for reduce='mean', normalization='component' (the defaults)")
        self.irreps, self.eps, self.affine = Irreps(str(irreps)), eps, affine
        self.reduce, self.normalization = reduce, normalization
        num_scalar = sum(m for m, l, _ in self.irreps if l == 0)
        if affine:
            self.weight = nn.Parameter(torch.ones(self.irreps.num_irreps))
            self.bias = nn.Parameter(torch.zeros(num_scalar))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)
        blocks, off, iw, ib = [], 0, 0, 0
        for m, l, _ in self.irreps:
            d = 2 * l + 1
            blocks.append([off, m, d, l, iw, ib])
            off, iw = off + m * d, iw + m
            if d == 1:
                ib += m
        self.register_buffer("_blocks", torch.tensor(blocks, dtype=torch.int32), persistent=False)

    def __repr__(self):
        return f"{self.__class__.__name__} ({self.irreps}, eps={self.eps})"

    def forward_equal_graphs(self, input, num_graphs: int, n_nodes: int):
        """The same normalisation for ``num_graphs`` graphs of ``n_nodes`` consecutive rows each: no host
        synchronisation (the form SEGNN's generic path uses, CUDA-graph capturable)."""
        ptr = torch.arange(num_graphs + 1, dtype=torch.int64, device=input.device) * n_nodes
        w = self.weight.detach().float().contiguous() if self.affine else None
        b = self.bias.detach().float().contiguous() if self.affine else None
        return ops.instance_norm(input.float().contiguous(), ptr, self._blocks.to(input.device), w, b, self.eps)

    def forward(self, input, batch):
        if not input.is_cuda:
            raise RuntimeError("InstanceNorm (B200) needs CUDA tensors: there is no CPU fallback")
        batch = batch.to(input.device)
        if batch.numel() > 1 and bool((batch[1:] < batch[:-1]).any()):
            raise ValueError("InstanceNorm needs the rows sorted by graph (a PyG batch vector)")
        graphs = int(batch.max().item()) + 1 if batch.numel() else 0
        ptr = torch.zeros(graphs + 1, dtype=torch.int64, device=input.device)
        ptr[1:] = torch.bincount(batch, minlength=graphs).cumsum(0)
        w = self.weight.detach().float().contiguous() if self.affine else None
        b = self.bias.detach().float().contiguous() if self.affine else None
        out = ops.instance_norm(input.float().contiguous(), ptr, self._blocks.to(input.device), w, b, self.eps)
        return out.to(input.dtype)
