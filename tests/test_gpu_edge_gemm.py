"""Large-graph (GEMM-form) edge layer, csrc/segnn_edge_gemm.cu: the "TN" 3xTF32 GEMM against float64, the forward and
backward entry points against the fused fp32 kernels on the same inputs (which the other tests hold to the oracle),
chunking, bit-identical repeats, and the whole training step against float64 autograd through the oracle with the GEMM
form forced on small graphs."""
import pytest
import torch

import segnn_b200 as S
from oracle import segnn_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("K,M,N", [(1, 4, 4), (7, 8, 12), (32, 128, 32), (1000, 128, 192), (4099, 64, 64),
                                   (5000, 192, 288), (20000, 128, 192), (333, 100, 36), (70000, 64, 64)])
def test_gemm_tn_matches_float64(K, M, N):
    g = torch.Generator().manual_seed(K + M + N)
    a = torch.randn(K, M, generator=g).cuda()
    b = torch.randn(K, N, generator=g).cuda()
    ref = a.double().t() @ b.double()
    out = S.ops.gemm_tn_tf32x3(a, b)
    err = rel(out, ref)
    print(f"K={K} M={M} N={N}: max rel err {err:.2e}")
    assert err < 4e-6
    assert torch.equal(S.ops.gemm_tn_tf32x3(a, b), out), "fixed-order split-K: bit-identical repeats"
    # accumulate on top of an existing matrix, row-strided operands (columns of wider buffers)
    wide_a = torch.full((K, M + 8), float("nan")).cuda()
    wide_b = torch.full((K, N + 4), float("nan")).cuda()
    wide_a[:, 4:4 + M] = a
    wide_b[:, :N] = b
    acc = torch.ones(M, N).cuda()
    S.ops.gemm_tn_tf32x3(wide_a[:, 4:4 + M], wide_b[:, :N], out=acc, accumulate=True)
    assert rel(acc, ref + 1.0) < 4e-6


@pytest.mark.parametrize("K,M,N,G", [(1000, 64, 64, 3), (5000, 96, 96, 3), (77, 32, 32, 3), (3000, 128, 64, 2),
                                     (20000, 64, 64, 3)])
def test_gemm_tn_grouped_matches_float64(K, M, N, G):
    """sum_g A_g^T B_g with the G blocks side by side in every row (the W_vv weight gradient over the three vector
    components)."""
    g = torch.Generator().manual_seed(K + M + N + G)
    a = torch.randn(K, G * M, generator=g).cuda()
    b = torch.randn(K, G * N, generator=g).cuda()
    ref = sum(a[:, i * M:(i + 1) * M].double().t() @ b[:, i * N:(i + 1) * N].double() for i in range(G))
    out = S.ops.gemm_tn_tf32x3(a, b, groups=G)
    err = rel(out, ref)
    print(f"K={K} M={M} N={N} G={G}: max rel err {err:.2e}")
    assert out.shape == (M, N) and err < 4e-6
    assert torch.equal(S.ops.gemm_tn_tf32x3(a, b, groups=G), out)
    # the same numbers as the stacked form [G K][M] x [G K][N]
    stacked = S.ops.gemm_tn_tf32x3(a.reshape(K * G, M), b.reshape(K * G, N))
    assert rel(out, stacked) < 2e-6


def _edge_inputs(B, N, n, seed=0, bn=False):
    g = torch.Generator().manual_seed(seed)
    nodes = B * N
    r = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).cuda()
    pos, mass = r(nodes, 3), (torch.rand(nodes, generator=g) + 0.5).cuda()
    p, q = r(nodes, 4, 3 * n, scale=0.5), r(nodes, 4, 3 * n, scale=0.5)
    w_edge1 = r(6 * n, scale=0.3)
    s = 1.0 / (2 * n) ** 0.5
    w2 = dict(ss=r(n, 2 * n, scale=s), vs=r(n, 2 * n, scale=s), sv=r(n, n, scale=s), vv=r(n, n, scale=s),
              b=r(2 * n, scale=0.1))
    bn_mul = (torch.rand(2 * n, generator=g) + 0.5).cuda() if bn else None
    bn_add = r(n, scale=0.1) if bn else None
    return pos, mass, p, q, w_edge1, w2, bn_mul, bn_add


CASES = [(3, 7, 32), (1, 50, 64), (2, 12, 96), (5, 5, 64), (1, 300, 64), (2, 33, 48), (2, 9, 20), (3, 2, 4)]


@pytest.mark.parametrize("B,N,n", CASES)
@pytest.mark.parametrize("bn", [False, True])
def test_gemm_form_forward_matches_fused_fp32(B, N, n, bn):
    pos, mass, p, q, w_edge1, w2, bn_mul, bn_add = _edge_inputs(B, N, n, seed=B + N + n, bn=bn)
    assert not S.ops._use_gemm_form(B, N, n), "the comparison target must be the fused kernel"
    agg0, mom0 = S.ops.edge_layer(S.ops.MODE_FP32, pos, mass, B, N, n, p, q, w_edge1, w2, bn_mul, bn_add,
                                  want_moments=True)
    agg1, mom1 = S.ops.edge_layer_gemm_fwd(pos, mass, B, N, n, p, q, w_edge1, w2, bn_mul, bn_add, want_moments=True)
    print(f"B={B} N={N} n={n}: agg {rel(agg1, agg0):.2e} moments {rel(mom1, mom0):.2e}")
    assert rel(agg1, agg0) < 5e-6 and rel(mom1, mom0) < 5e-6
    assert torch.equal(S.ops.edge_layer_gemm_fwd(pos, mass, B, N, n, p, q, w_edge1, w2, bn_mul, bn_add), agg1)


@pytest.mark.parametrize("B,N,n", CASES)
@pytest.mark.parametrize("chunked", [False, True])
def test_gemm_form_backward_matches_fused_fp32(B, N, n, chunked, monkeypatch):
    pos, mass, p, q, w_edge1, w2, _, _ = _edge_inputs(B, N, n, seed=2 * B + N + n)
    g = torch.Generator().manual_seed(99)
    r = lambda *s: torch.randn(*s, generator=g).cuda()
    bn_a, bn_b, bn_c = r(2 * n) * 0.5 + 1.0, r(2 * n) * 0.01, r(n) * 0.01
    dagg = r(B * N, 4, n)
    monkeypatch.setattr(S.ops, "GEMM_FORM_MIN_ROWS_TRAINING", 1 << 62)  # the comparison target is the fused kernel
    dP0, dQ0, g0, dwe0 = S.ops.edge_layer_bwd(pos, mass, B, N, n, p, q, w_edge1, w2, bn_a, bn_b, bn_c, dagg)
    if chunked:  # room for one graph per chunk: B chunks, accumulated weight gradients
        one = int(S.ops.lib.segnn_edge_layer_gemm_workspace(1, N, n, 1, 0))
        monkeypatch.setattr(S.ops, "GEMM_FORM_BUDGET_BYTES", one)
    dP1, dQ1, g1, dwe1 = S.ops.edge_layer_gemm_bwd(pos, mass, B, N, n, p, q, w_edge1, w2, bn_a, bn_b, bn_c, dagg)
    errs = dict(dP=rel(dP1, dP0), dQ=rel(dQ1, dQ0), dwe=rel(dwe1, dwe0), **{k: rel(g1[k], g0[k]) for k in g0})
    print(f"B={B} N={N} n={n} chunked={chunked}: " + " ".join(f"{k} {v:.1e}" for k, v in errs.items()))
    assert max(errs.values()) < 2e-5, errs
    dP2, dQ2, g2, dwe2 = S.ops.edge_layer_gemm_bwd(pos, mass, B, N, n, p, q, w_edge1, w2, bn_a, bn_b, bn_c, dagg)
    assert torch.equal(dP1, dP2) and torch.equal(dQ1, dQ2) and torch.equal(dwe1, dwe2)
    assert all(torch.equal(g1[k], g2[k]) for k in g1), "weight gradients are bit-identical from run to run"
    # rows kept by the forward call instead of the recompute: same bits, and the kept rows are not modified
    agg, rows = S.ops.edge_layer_gemm_fwd(pos, mass, B, N, n, p, q, w_edge1, w2, keep_rows=True)
    snapshot = rows[0].clone()
    for _ in range(2):
        dP3, dQ3, g3, dwe3 = S.ops.edge_layer_gemm_bwd(pos, mass, B, N, n, p, q, w_edge1, w2, bn_a, bn_b, bn_c, dagg,
                                                       rows=rows)
        assert torch.equal(dP1, dP3) and torch.equal(dQ1, dQ3) and torch.equal(dwe1, dwe3)
        assert all(torch.equal(g1[k], g3[k]) for k in g1)
    assert torch.equal(snapshot, rows[0])


@pytest.mark.parametrize("keep", [True, False])
@pytest.mark.parametrize("H,L,B,N,bn_train", [(64, 2, 4, 5, True), (192, 1, 2, 20, True), (128, 2, 1, 33, False),
                                                (128, 1, 1, 150, True)])
def test_training_step_in_gemm_form_matches_oracle(H, L, B, N, bn_train, keep, monkeypatch):
    """Whole model, forward + backward, with every edge layer forced onto the GEMM form (edge rows kept between forward
    and backward, or recomputed): prediction 1e-5, gradients within the tolerance of
    tests/test_gpu_parity.py::test_training_gradients_match_oracle."""
    monkeypatch.setattr(S.ops, "GEMM_FORM_MIN_ROWS", 0)
    monkeypatch.setattr(S.ops, "GEMM_FORM_MIN_ROWS_TRAINING", 0)
    if not keep:
        monkeypatch.setattr(S.ops, "GEMM_FORM_KEEP_BYTES_PER_LAYER", 0)
    assert S.ops.gemm_form_keeps_rows(B, N, H // 2) == keep
    torch.manual_seed(0)
    om = O.SEGNN(hidden_features=H, num_layers=L)
    O.perturb_bn_buffers(om, seed=1)
    om.train(bn_train)
    m = S.SEGNN(hidden_features=H, num_layers=L)
    m.load_state_dict(om.state_dict())
    m = m.float().cuda().train(bn_train)
    pos, vel, mass = O.synthetic_system(B, N, seed=2)
    y = torch.randn(B * N, 6, dtype=torch.float64)
    ref = om(O.make_graph(pos.reshape(-1, 3), vel.reshape(-1, 3), mass.reshape(-1, 1), B, N))
    O.target_common_loss(ref, y).backward()
    g = S.GraphBatch(pos=pos.reshape(-1, 3).float().cuda(), vel=vel.reshape(-1, 3).float().cuda(),
                     mass=mass.reshape(-1, 1).float().cuda(), num_graphs=B, n_nodes=N)
    before = S.ops.launch_count()
    pred = m(g)
    O.target_common_loss(pred, y.float().cuda()).backward()
    assert S.ops.launch_count() - before > 30 * L
    assert rel(pred.detach().cpu(), ref.detach()) < 1e-5
    top = max(float(a.grad.abs().max()) for a in om.parameters())
    worst = 0.0
    for (k, a), (k2, b) in zip(om.named_parameters(), m.named_parameters()):
        assert k == k2 and b.grad is not None, k
        scale = float(a.grad.abs().max())
        err = float((a.grad - b.grad.double().cpu()).abs().max())
        worst = max(worst, err / max(scale, 1e-30)) if scale > 1e-9 else worst
        assert err <= 1e-4 * scale + 1e-5 * top, f"{k}: {err} vs scale {scale} (top {top})"
    print(f"H={H} N={N}: worst gradient rel err {worst:.2e}")


@pytest.mark.parametrize("keep", [True, False])
def test_side_stream_weight_gradients_are_bit_identical_to_the_inline_order(keep, monkeypatch):
    """GEMM-form edge backward with the message_layer_2 weight gradients launched through the side stream
    (segnn_edge_layer_gemm_bwd_phases 1 | 2 | 4) against the single in-line call (phases = 7): every gradient of a
    README-shaped training step is bit-identical, kept rows and recomputed rows."""
    monkeypatch.setattr(S.ops, "GEMM_FORM_MIN_ROWS", 0)
    monkeypatch.setattr(S.ops, "GEMM_FORM_MIN_ROWS_TRAINING", 0)
    if not keep:
        monkeypatch.setattr(S.ops, "GEMM_FORM_KEEP_BYTES_PER_LAYER", 0)
    B, N, H, L = 16, 5, 192, 2
    torch.manual_seed(3)
    m = S.SEGNN(hidden_features=H, num_layers=L).float().cuda().train()
    pos, vel, mass = O.synthetic_system(B, N, seed=4)
    y = torch.randn(B * N, 6).cuda()
    g = S.GraphBatch(pos=pos.reshape(-1, 3).float().cuda(), vel=vel.reshape(-1, 3).float().cuda(),
                     mass=mass.reshape(-1, 1).float().cuda(), num_graphs=B, n_nodes=N)
    state = {k: v.clone() for k, v in m.state_dict().items()}
    grads = {}
    for side in (False, True, True):
        monkeypatch.setattr(S.ops, "SIDE_STREAM_W2_GRADS", side)
        m.load_state_dict(state)  # the same running statistics before every pass
        m.zero_grad(set_to_none=True)
        O.target_common_loss(m(g), y).backward()
        torch.cuda.synchronize()
        cur = {k: p.grad.clone() for k, p in m.named_parameters()}
        if side in grads:
            assert all(torch.equal(cur[k], grads[side][k]) for k in cur)
        grads[side] = cur
    for k in grads[True]:
        assert torch.equal(grads[True][k], grads[False][k]), k
