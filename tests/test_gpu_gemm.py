"""segnn_gemm_tf32x3: the fp32-accurate tcgen05 GEMM of the generic-irreps tensor products, against float64 torch on
ragged shapes (M not a multiple of 128, K not a multiple of 32 or 4, N below / above one 128-column block)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,K,N", [(1, 1, 1), (5, 7, 3), (128, 32, 16), (300, 146, 73), (1000, 219, 73), (515, 146, 73),
                                   (4097, 292, 146), (2000, 64, 200), (777, 500, 129), (20000, 219, 80)])
def test_gemm_tf32x3_matches_float64(M, K, N):
    import segnn_b200 as S
    g = torch.Generator().manual_seed(M + K + N)
    a = torch.randn(M, K, generator=g).cuda()
    b = torch.randn(K, N, generator=g).cuda()
    ref = a.double() @ b.double()
    out = S.ops.gemm_tf32x3(a, b)
    err = float((out.double() - ref).abs().max() / ref.abs().max())
    err32 = float(((a @ b).double() - ref).abs().max() / ref.abs().max())  # the library fp32 SGEMM, for scale
    print(f"M={M} K={K} N={N}: max rel err {err:.2e} (library fp32 GEMM {err32:.2e})")
    assert err < 4e-6
    # strided views (padded leading dimensions, as the generic path passes them) and bit-identical repeats
    lda, ldy = (K + 3) & ~3, (N + 3) & ~3
    ap = torch.full((M, lda), float("nan")).cuda()
    ap[:, :K] = a
    yp = torch.zeros((M, ldy)).cuda()
    S.ops.gemm_tf32x3(ap[:, :K], b, out=yp[:, :N])
    assert torch.equal(yp[:, :N], out), "padding columns must not be read as data; result independent of strides"
    assert float(yp[:, N:].abs().max() if ldy > N else 0.0) == 0.0, "nothing is written past N"
    assert torch.equal(S.ops.gemm_tf32x3(a, b), out)


def test_gemm_tf32x3_is_much_more_accurate_than_plain_tf32():
    """The error compensation is what keeps the generic path inside the 1e-5 budget: a single tf32 product is ~1e-3."""
    import segnn_b200 as S
    g = torch.Generator().manual_seed(0)
    a, b = torch.randn(512, 256, generator=g).cuda(), torch.randn(256, 96, generator=g).cuda()
    ref = a.double() @ b.double()
    out = S.ops.gemm_tf32x3(a, b)
    trunc = lambda t: (t.view(torch.int32) & ~0x1FFF).view(torch.float32)
    plain = trunc(a).double() @ trunc(b).double()
    e3 = float((out.double() - ref).abs().max() / ref.abs().max())
    e1 = float((plain - ref).abs().max() / ref.abs().max())
    print(f"3xTF32 {e3:.2e} vs plain tf32 {e1:.2e}")
    assert e3 < 4e-6 and e1 > 50 * e3


@pytest.mark.parametrize("nodes,n_in,n_out,two,split,n_bias", [(320, 96, 576, False, 288, 192), (320, 96, 288, True, 0, 0),
                                                                (7, 32, 64, False, 0, 64), (1000, 64, 384, False, 192, 128),
                                                                (33, 25, 50, True, 0, 25), (5000, 96, 192, False, 0, 96)])
def test_node_gemm_tf32x3_matches_the_ffma_kernel(nodes, n_in, n_out, two, split, n_bias, monkeypatch):
    """segnn_node_gemm_tf32x3 (tcgen05, both row classes in one launch, concatenated K, bias, split outputs) against
    segnn_node_gemm (FFMA) and a float64 product."""
    import segnn_b200 as S
    g = torch.Generator().manual_seed(nodes + n_in)
    r = lambda *s: torch.randn(*s, generator=g).cuda()
    x0, x1 = r(nodes, 4, n_in), (r(nodes, 4, n_in) if two else None)
    K = n_in * (2 if two else 1)
    w = dict(w_s=r(K, n_out) / K ** 0.5, w_v=r(K, n_out) / K ** 0.5)
    bias = r(n_bias) if n_bias else None
    outs = {}
    monkeypatch.setattr(S.ops, "NODE_GEMM_TF32X3_MIN_NODES", 0)
    for flag in (True, False):
        monkeypatch.setattr(S.ops, "NODE_GEMM_TF32X3", flag)
        y = S.ops.node_gemm(x0, x1, w, n_out, bias=bias, n_bias=n_bias, split=split)
        outs[flag] = torch.cat(y, dim=2) if split else y
    xin = (torch.cat([x0, x1], dim=2) if two else x0).double()
    ref = torch.empty(nodes, 4, n_out, dtype=torch.float64, device="cuda")
    ref[:, 0] = xin[:, 0] @ w["w_s"].double()
    ref[:, 1:] = xin[:, 1:] @ w["w_v"].double()
    if bias is not None:
        ref[:, 0, :n_bias] += bias.double()
    e_tc = float((outs[True].double() - ref).abs().max() / ref.abs().max())
    e_ff = float((outs[False].double() - ref).abs().max() / ref.abs().max())
    print(f"nodes={nodes} K={K} n_out={n_out}: tcgen05 3xTF32 {e_tc:.2e}, FFMA {e_ff:.2e}")
    assert e_tc < 3e-6 and e_ff < 3e-6


@pytest.mark.parametrize("nodes,n_in,two,n_out,split", [(320, 96, False, 576, 288), (320, 96, True, 288, 288),
                                                        (37, 96, False, 192, 192), (1000, 288, False, 192, 96),
                                                        (500, 32, True, 96, 96), (321, 64, False, 100, 100),
                                                        (4096, 96, True, 96, 96)])
def test_small_batch_node_gemm_is_bit_identical_to_the_tile_kernel(nodes, n_in, two, n_out, split):
    """segnn_node_gemm below 4096 nodes runs node_gemm_small_kernel (32 x 64 tiles, 4-stage cp.async ring); the same
    rows inside a batch above the threshold run node_gemm_kernel (64 x 64 tiles).  Same accumulation order: the outputs
    of the shared rows are bit-identical -- ragged row and column tiles, one and two inputs, split outputs, bias."""
    import segnn_b200 as S
    ops = S.ops
    old = ops.NODE_GEMM_TF32X3
    ops.NODE_GEMM_TF32X3 = False  # keep the large batch on the FFMA tile kernel
    try:
        gen = torch.Generator(device="cpu").manual_seed(nodes + n_out)
        big = 4100 + nodes
        K = n_in * (2 if two else 1)
        x0 = torch.randn(big, 4, n_in, generator=gen).cuda()
        x1 = torch.randn(big, 4, n_in, generator=gen).cuda() if two else None
        w = dict(w_s=(torch.randn(K, n_out, generator=gen) / K ** 0.5).cuda(),
                 w_v=(torch.randn(K, n_out, generator=gen) / K ** 0.5).cuda())
        n_bias = min(n_out, 64)
        bias = torch.randn(n_bias, generator=gen).cuda()
        kw = dict(bias=bias, n_bias=n_bias)
        if split < n_out:
            r0, r1 = ops.node_gemm(x0, x1, w, n_out, split=split, tc=False, **kw)
            s0, s1 = ops.node_gemm(x0[:nodes].contiguous(), None if x1 is None else x1[:nodes].contiguous(), w, n_out,
                                   split=split, tc=False, **kw)
            assert torch.equal(s0, r0[:nodes]) and torch.equal(s1, r1[:nodes])
        else:
            r = ops.node_gemm(x0, x1, w, n_out, tc=False, **kw)
            s = ops.node_gemm(x0[:nodes].contiguous(), None if x1 is None else x1[:nodes].contiguous(), w, n_out,
                              tc=False, **kw)
            assert torch.equal(s, r[:nodes])
            ref = torch.cat([x0[:nodes], x1[:nodes]], 2) if two else x0[:nodes]
            ref = torch.stack([ref[:, 0].double() @ w["w_s"].double()] +
                              [ref[:, c].double() @ w["w_v"].double() for c in (1, 2, 3)], 1)
            ref[:, 0, :n_bias] += bias.double()
            assert float((s.double() - ref).abs().max()) < 1e-4
    finally:
        ops.NODE_GEMM_TF32X3 = old
