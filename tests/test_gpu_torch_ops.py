"""torch.library operators (torch.ops.segnn_b200.*): schema / fake-tensor / autograd-registration checks with
torch.library.opcheck, and the registered backward formulas against central differences of the composed network
(embed -> P/Q projection -> fused edge layer -> update_layer_1 (gate) -> update_layer_2 (+ residual) -> pre_pool1 ->
head), i.e. one SEGNN layer without BatchNorm written with nothing but the custom operators and torch autograd."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _inputs(B=2, N=6, n=8, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).cuda()
    nodes = B * N
    d = dict(pos=r(nodes, 3), vel=r(nodes, 3, sc=0.3), mass=torch.sign(r(nodes)) + 0.0)
    k = 1.0 / n ** 0.5
    d["w"] = dict(we=r(6, n, sc=0.5), be=r(n, sc=0.1), m_s=r(n, 6 * n, sc=k), m_v=r(n, 6 * n, sc=k), m_b=r(2 * n, sc=0.1),
                  wedge=r(6 * n, sc=0.3), ss=r(n, 2 * n, sc=k), vs=r(n, 2 * n, sc=k), sv=r(n, n, sc=k), vv=r(n, n, sc=k),
                  b2=r(2 * n, sc=0.1), u_s=r(2 * n, 3 * n, sc=k), u_v=r(2 * n, 3 * n, sc=k), u_b=r(2 * n, sc=0.1),
                  v_s=r(n, 2 * n, sc=k), v_v=r(n, 2 * n, sc=k), v_b=r(n, sc=0.1), p_s=r(n, 3 * n, sc=k),
                  p_v=r(n, 3 * n, sc=k), p_b=r(2 * n, sc=0.1), head=r(2, n, 2, sc=k))
    d["target"] = r(nodes, 6)
    return d, B, N, n


def _network(d, B, N, n, w):
    T = torch.ops.segnn_b200
    x_in, attr = T.prep(d["pos"], d["vel"], B, N)
    h = T.embed(x_in, attr, w["we"], w["be"], n)
    pq = T.node_gemm(h, None, w["m_s"], w["m_v"], w["m_b"], 2 * n)
    p, q = pq[:, :, :3 * n].contiguous(), pq[:, :, 3 * n:].contiguous()
    agg = T.edge_layer(d["pos"], d["mass"], B, N, n, p, q, w["wedge"], w["ss"], w["vs"], w["sv"], w["vv"], w["b2"])
    g1 = T.tp_combine(T.node_gemm(h, agg, w["u_s"], w["u_v"], None, 0), attr, n, True, w["u_b"])
    h2 = T.tp_combine(T.node_gemm(g1, None, w["v_s"], w["v_v"], None, 0), attr, n, False, w["v_b"]) + h
    hp = T.tp_combine(T.node_gemm(h2, None, w["p_s"], w["p_v"], None, 0), attr, n, True, w["p_b"])
    pred = T.head(hp, attr, w["head"], n)
    return (pred * d["target"]).sum() / pred.shape[0]


def test_operators_are_registered_with_schema_fake_and_autograd():
    import segnn_b200  # noqa: F401
    from segnn_b200 import torch_ops as TO
    d, B, N, n = _inputs()
    w = d["w"]
    T = torch.ops.segnn_b200
    x_in, attr = T.prep(d["pos"], d["vel"], B, N)
    h = T.embed(x_in, attr, w["we"], w["be"], n)
    pq = T.node_gemm(h, None, w["m_s"], w["m_v"], w["m_b"], 2 * n)
    p, q = pq[:, :, :3 * n].contiguous(), pq[:, :, 3 * n:].contiguous()
    y1 = T.node_gemm(h, h, w["u_s"], w["u_v"], None, 0)
    checks = ("test_schema", "test_faketensor", "test_autograd_registration")
    rg = lambda t: t.clone().requires_grad_(True)
    torch.library.opcheck(TO.prep, (d["pos"], d["vel"], B, N), test_utils=checks)
    torch.library.opcheck(TO.embed, (x_in, attr, rg(w["we"]), rg(w["be"]), n), test_utils=checks)
    torch.library.opcheck(TO.node_gemm, (rg(h), None, rg(w["m_s"]), rg(w["m_v"]), rg(w["m_b"]), 2 * n), test_utils=checks)
    torch.library.opcheck(TO.tp_combine, (rg(y1), attr, n, True, rg(w["u_b"])), test_utils=checks)
    torch.library.opcheck(TO.edge_layer, (d["pos"], d["mass"], B, N, n, rg(p), rg(q), rg(w["wedge"]), rg(w["ss"]),
                                          rg(w["vs"]), rg(w["sv"]), rg(w["vv"]), rg(w["b2"])), test_utils=checks)
    torch.library.opcheck(TO.head, (rg(h), attr, rg(w["head"]), n), test_utils=checks)
    # fake tensors: shapes without touching the GPU kernels
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode(allow_non_fake_inputs=False) as mode:
        fp, fq = mode.from_tensor(p), mode.from_tensor(q)
        out = T.edge_layer(mode.from_tensor(d["pos"]), mode.from_tensor(d["mass"]), B, N, n, fp, fq,
                           mode.from_tensor(w["wedge"]), mode.from_tensor(w["ss"]), mode.from_tensor(w["vs"]),
                           mode.from_tensor(w["sv"]), mode.from_tensor(w["vv"]), mode.from_tensor(w["b2"]))
        assert tuple(out.shape) == (B * N, 4, n)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        T.prep(d["pos"].cpu(), d["vel"].cpu(), B, N)


def test_registered_backward_matches_central_differences():
    """Directional derivatives of the composed network: autograd through register_autograd vs (f(w + e v) - f(w - e v))
    / 2e in fp32 with e = 2e-2 (truncation ~ e^2, rounding ~ 1e-7 / e); 2% of the directional derivative's scale."""
    import segnn_b200  # noqa: F401
    d, B, N, n = _inputs(seed=3)
    w = {k: v.clone().requires_grad_(True) for k, v in d["w"].items()}
    loss = _network(d, B, N, n, w)
    grads = torch.autograd.grad(loss, list(w.values()))
    g = dict(zip(w.keys(), grads))
    gen = torch.Generator().manual_seed(7)
    worst = 0.0
    for k in w:
        assert g[k] is not None and g[k].shape == w[k].shape and torch.isfinite(g[k]).all(), k
        v = torch.randn(w[k].shape, generator=gen).cuda()
        v = v / v.norm()
        eps = 2e-2
        with torch.no_grad():
            wp = {kk: (vv.detach() + eps * v if kk == k else vv.detach()) for kk, vv in w.items()}
            wm = {kk: (vv.detach() - eps * v if kk == k else vv.detach()) for kk, vv in w.items()}
            fd = float(_network(d, B, N, n, wp) - _network(d, B, N, n, wm)) / (2 * eps)
        an = float((g[k] * v).sum())
        scale = float(g[k].norm()) + 1e-6
        worst = max(worst, abs(fd - an) / scale)
        assert abs(fd - an) <= 2e-2 * scale + 2e-4, f"{k}: autograd {an:.6f} vs central difference {fd:.6f} (|g| {scale:.4f})"
    print(f"worst directional-derivative mismatch / |grad|: {worst:.2e}")
    # run-to-run bit identity of the registered backward (fixed-order reductions, no atomics)
    grads2 = torch.autograd.grad(_network(d, B, N, n, w), list(w.values()))
    for a, b in zip(grads, grads2):
        assert torch.equal(a, b)


def test_edge_list_operators_match_the_module_path():
    """torch.ops.segnn_b200.{prep_lmax, edge_attr_list, message_input_list, segment_reduce}: opcheck + the values the
    package's own edge-list path computes (kNN graph, lmax_attr = 2)."""
    import segnn_b200 as S
    torch.manual_seed(3)
    T = torch.ops.segnn_b200
    B, N, k, D = 2, 9, 3, 24
    nodes = B * N
    pos, vel = torch.randn(nodes, 3, device="cuda"), torch.randn(nodes, 3, device="cuda")
    mass = torch.randn(nodes, device="cuda")
    ei = S.build_graph_with_knn(pos, B, N, "cuda", k)
    order, ptr = S.ops.edge_list_csr(ei, nodes)
    ea, add = T.edge_attr_list(pos, mass, ei, 2)
    ea2, add2 = S.ops.edge_attr_list(pos, mass, ei, 2)
    assert torch.equal(ea, ea2) and torch.equal(add, add2)
    x_in, attr = T.prep_lmax(pos, vel, B, N, 2)
    # K1 stages the senders in tiles, the lmax kernel sums them in index order: same values to rounding
    assert attr.shape == (nodes, 9) and float((attr[:, :4] - S.ops.prep(pos, vel, B, N)[1]).abs().max()) < 1e-6
    x = torch.randn(nodes, D, device="cuda")
    inp = T.message_input_list(x, add, ei)
    assert torch.equal(inp, torch.cat([x[ei[1]], x[ei[0]], add], dim=1))
    vals = torch.randn(ei.shape[1], D, device="cuda")
    agg = T.segment_reduce(vals, order, ptr, False)
    ref = torch.zeros(nodes, D, device="cuda", dtype=torch.float64).index_add_(0, ei[1], vals.double())
    assert float((agg.double() - ref).abs().max()) < 1e-5
    mean = T.segment_reduce(vals, order, ptr, True)
    cnt = torch.bincount(ei[1], minlength=nodes).clamp_min(1).unsqueeze(1)
    assert float((mean.double() - ref / cnt).abs().max()) < 1e-5
    for op, args in ((T.edge_attr_list, (pos, mass, ei, 2)), (T.prep_lmax, (pos, vel, B, N, 2)),
                     (T.message_input_list, (x, add, ei)), (T.segment_reduce, (vals, order, ptr, False))):
        torch.library.opcheck(op, args, test_utils=("test_schema", "test_faketensor"))
