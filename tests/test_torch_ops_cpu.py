"""torch.library registration of the operator boundary, checked without a GPU: every operator exists with a schema,
traces under FakeTensorMode (shape inference through register_fake, no kernel launched) and refuses CPU tensors."""
import pytest
import torch

import segnn_b200  # noqa: F401  (registers torch.ops.segnn_b200.*)

EXPECTED = ["prep", "embed", "embed_bwd", "node_gemm", "node_gemm_wgrad", "tp_combine", "tp_combine_bwd", "edge_layer",
            "edge_layer_bwd", "edge_layer_tc", "head", "head_bwd", "prep_lmax", "edge_attr_list", "message_input_list",
            "segment_reduce"]


def test_every_operator_is_registered_with_a_schema():
    for name in EXPECTED:
        op = getattr(torch.ops.segnn_b200, name)
        assert str(op.default._schema).startswith(f"segnn_b200::{name}("), name


def test_differentiable_operators_have_autograd_registered():
    from segnn_b200 import torch_ops as TO
    for op in (TO.embed, TO.node_gemm, TO.tp_combine, TO.edge_layer, TO.head):
        assert op._backward_fn is not None, op


def test_fake_tensor_tracing_of_one_layer():
    from torch._subclasses.fake_tensor import FakeTensorMode
    T = torch.ops.segnn_b200
    B, N, n = 3, 10, 16
    nodes = B * N
    with FakeTensorMode():
        e = lambda *s: torch.empty(*s, device="cuda")
        pos, vel, mass = e(nodes, 3), e(nodes, 3), e(nodes)
        x_in, attr = T.prep(pos, vel, B, N)
        h = T.embed(x_in, attr, e(6, n), e(n), n)
        pq = T.node_gemm(h, None, e(n, 6 * n), e(n, 6 * n), e(2 * n), 2 * n)
        p, q = pq[:, :, :3 * n].contiguous(), pq[:, :, 3 * n:].contiguous()
        agg = T.edge_layer(pos, mass, B, N, n, p, q, e(6 * n), e(n, 2 * n), e(n, 2 * n), e(n, n), e(n, n), e(2 * n))
        agg16 = T.edge_layer_tc(3, pos, mass, B, N, n, p, q, e(6 * n), e(2 * n), e(128, 3 * n), None, None)
        g1 = T.tp_combine(T.node_gemm(h, agg, e(2 * n, 3 * n), e(2 * n, 3 * n), None, 0), attr, n, True, e(2 * n))
        pred = T.head(g1, attr, e(2, n, 2), n)
        assert tuple(x_in.shape) == (nodes, 7) and tuple(attr.shape) == (nodes, 4)
        assert tuple(agg.shape) == tuple(agg16.shape) == tuple(g1.shape) == (nodes, 4, n)
        assert tuple(pred.shape) == (nodes, 6)


def test_cpu_tensors_are_refused():
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        torch.ops.segnn_b200.prep(torch.zeros(12, 3), torch.zeros(12, 3), 2, 6)


def test_fake_tensor_tracing_of_the_edge_list_operators():
    """lmax_attr = 2 geometry + kNN edge list -> message input -> deterministic segment sum, shapes only."""
    from torch._subclasses.fake_tensor import FakeTensorMode
    T = torch.ops.segnn_b200
    B, N, k, D = 2, 9, 3, 40
    nodes, E = B * N, B * N * 3
    with FakeTensorMode():
        e = lambda *s: torch.empty(*s, device="cuda")
        pos, vel, mass = e(nodes, 3), e(nodes, 3), e(nodes)
        ei = torch.empty(2, E, dtype=torch.int64, device="cuda")
        x_in, attr = T.prep_lmax(pos, vel, B, N, 2)
        ea, add = T.edge_attr_list(pos, mass, ei, 2)
        inp = T.message_input_list(e(nodes, D), add, ei)
        order, ptr = torch.empty(E, dtype=torch.int64, device="cuda"), torch.empty(nodes + 1, dtype=torch.int64, device="cuda")
        agg = T.segment_reduce(e(E, D), order, ptr, False)
        assert tuple(attr.shape) == (nodes, 9) and tuple(ea.shape) == (E, 9) and tuple(add.shape) == (E, 2)
        assert tuple(inp.shape) == (E, 2 * D + 2) and tuple(agg.shape) == (nodes, D)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        T.segment_reduce(torch.zeros(4, 3), torch.zeros(4, dtype=torch.int64), torch.zeros(3, dtype=torch.int64), False)
