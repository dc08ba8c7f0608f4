"""``torch_geometric.nn.MessagePassing`` (flow source_to_target, aggr add/mean/max) and the global pools -- TEST
INFRASTRUCTURE, see ../../README.md.

``propagate(edge_index, **kwargs)`` follows PyG's collect / message / aggregate / update contract: an argument of
``message`` named ``<name>_j`` receives ``kwargs[name].index_select(node_dim, edge_index[0])`` (source / sender), one
named ``<name>_i`` receives ``... edge_index[1]`` (target / receiver), any other argument is passed through;
messages are reduced onto ``edge_index[1]`` with ``dim_size`` = number of nodes; ``update`` receives the aggregate plus
the keyword arguments its signature names."""
import inspect

import torch
import torch.nn as nn

from torch_scatter import scatter


class MessagePassing(nn.Module):
    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2, **kwargs):
        super().__init__()
        assert flow in ("source_to_target", "target_to_source")
        self.aggr, self.flow, self.node_dim = aggr, flow, node_dim

    def propagate(self, edge_index, size=None, **kwargs):
        i, j = (1, 0) if self.flow == "source_to_target" else (0, 1)
        msg_params = [p for p in inspect.signature(self.message).parameters]
        num_nodes = None
        msg_kwargs = {}
        for name in msg_params:
            if name.endswith("_i") or name.endswith("_j"):
                src = kwargs[name[:-2]]
                if src is None:
                    msg_kwargs[name] = None
                    continue
                num_nodes = src.size(self.node_dim)
                msg_kwargs[name] = src.index_select(self.node_dim, edge_index[i if name.endswith("_i") else j])
            else:
                msg_kwargs[name] = kwargs.get(name)
        if num_nodes is None:
            num_nodes = int(edge_index.max()) + 1
        out = self.message(**msg_kwargs)
        out = self.aggregate(out, edge_index[i], dim_size=num_nodes)
        upd_params = list(inspect.signature(self.update).parameters)[1:]
        return self.update(out, **{k: kwargs[k] for k in upd_params if k in kwargs})

    def aggregate(self, inputs, index, dim_size=None):
        reduce = {"add": "sum", "sum": "sum", "mean": "mean", "max": "max"}[self.aggr]
        return scatter(inputs, index, dim=self.node_dim, dim_size=dim_size, reduce=reduce)

    def message(self, x_j):
        return x_j

    def update(self, inputs):
        return inputs


def _pool(x, batch, reduce, size=None):
    if batch is None:
        return getattr(x, {"sum": "sum", "mean": "mean"}.get(reduce, "amax"))(dim=0, keepdim=True)
    size = int(batch.max()) + 1 if size is None else size
    return scatter(x, batch, dim=0, dim_size=size, reduce=reduce)


def global_add_pool(x, batch, size=None):
    return _pool(x, batch, "sum", size)


def global_mean_pool(x, batch, size=None):
    return _pool(x, batch, "mean", size)


def global_max_pool(x, batch, size=None):
    return _pool(x, batch, "max", size)


def knn_graph(*a, **k):  # imported by modules the SEGNN path never calls
    raise NotImplementedError("shim: torch_geometric.nn.knn_graph is not on the SEGNN path")


def radius_graph(*a, **k):
    raise NotImplementedError("shim: torch_geometric.nn.radius_graph is not on the SEGNN path")
