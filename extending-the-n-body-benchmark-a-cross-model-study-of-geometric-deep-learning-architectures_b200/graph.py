"""Implicit fully-connected batched graph (replaces torch_geometric Data/Batch + the edge list of
utils/build_fully_connected_graph.py on the hot path)."""
from __future__ import annotations

import torch

from . import ops


class GraphBatch:
    """Attribute bag with the field names the reference's graphs carry (pos, vel, force, mass, y, batch, x,
    node_attr, ...), plus ``num_graphs`` / ``n_nodes`` describing the implicit complete graph. ``edge_index``,
    ``edge_attr`` and ``additional_message_features`` are materialised lazily and only when someone asks."""

    _LAZY = ("edge_index", "edge_attr", "additional_message_features")

    def __init__(self, **fields):
        self.__dict__.update(fields)

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                self.__dict__[k] = v.to(device)
        return self

    @property
    def num_nodes(self):
        return self.pos.shape[0]

    def has_isolated_nodes(self):
        return self.n_nodes < 2

    def __getattr__(self, name):  # only called when normal lookup fails
        if name in GraphBatch._LAZY and "pos" in self.__dict__:
            if name == "edge_index":
                val = ops.edge_index(self.num_graphs, self.n_nodes, self.pos.device)
            else:
                ea, add = ops.edge_attr(self.pos, self.mass, self.num_graphs, self.n_nodes,
                                        int(self.__dict__.get("lmax_attr", 1)))
                self.__dict__["edge_attr"] = ea.to(self.pos.dtype)
                self.__dict__["additional_message_features"] = add.to(self.pos.dtype)
                return self.__dict__[name]
            self.__dict__[name] = val
            return val
        raise AttributeError(name)


def _build_fully_connected_edge_index(batch_size, num_nodes, device):
    """utils/build_fully_connected_graph.py:4-20, enumerated on the device by one kernel (no nonzero/host sync)."""
    return ops.edge_index(int(batch_size), int(num_nodes), device)


def build_graph_with_knn(loc, batch_size, num_nodes, device, num_neighbors):
    """utils/build_fully_connected_graph.py:23-80 (complete graph fast path :39-40, kNN branch :42-80)."""
    num_nodes = int(num_nodes)
    num_neighbors = int(num_neighbors) if num_neighbors is not None else num_nodes - 1
    if num_neighbors >= num_nodes:
        raise ValueError("Graph cannot have more neighbors than there are nodes in simulation - 1")
    if num_neighbors != num_nodes - 1:
        # :42-80: k nearest neighbours per node.  The fused SEGNN kernels are specialised for the complete graph (the
        # configured path, num_neighbors = N - 1); a model run on this edge list takes the generic-irreps kernels
        # (SEGNN.forward_edge_list, inference).
        return ops.knn_edge_index(loc, int(batch_size), num_nodes, num_neighbors, device)
    return _build_fully_connected_edge_index(batch_size, num_nodes, device)


def infer_graph_shape(graph):
    """(num_graphs, nodes_per_graph) of a batched graph of equal-size complete graphs."""
    if hasattr(graph, "num_graphs") and hasattr(graph, "n_nodes"):
        return int(graph.num_graphs), int(graph.n_nodes)
    nodes = graph.pos.shape[0]
    batch = getattr(graph, "batch", None)
    if batch is None:
        return 1, nodes
    b = int(batch.max().item()) + 1
    if nodes % b != 0:
        raise ValueError("the accelerated SEGNN path needs equal-size graphs")
    return b, nodes // b
