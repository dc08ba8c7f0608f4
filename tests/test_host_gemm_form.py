"""Host-side logic of the GEMM-form edge layers (no GPU): workspace sizing / chunking of segnn_edge_layer_gemm_*, the
dispatch thresholds in ops.py, and the irreps test that selects the lmax_h = 2 GEMM form."""
import pytest

import segnn_b200 as S
from segnn_b200 import ops
from segnn_b200._lib import lib


def test_workspace_grows_with_graphs_and_respects_the_budget():
    N, n = 100, 96
    one = lib.segnn_edge_layer_gemm_workspace(1, N, n, 1, 0)
    all64 = lib.segnn_edge_layer_gemm_workspace(64, N, n, 1, 0)
    assert 0 < one < all64
    per_graph = (all64 - one) / 63
    assert per_graph > 4 * N * N * 16 * n, "16 n floats per edge row backward, plus partial rows"
    # a budget of ~10 graphs gives a chunk of <= 10 graphs; a budget below one graph still gives one graph
    ten = lib.segnn_edge_layer_gemm_workspace(64, N, n, 1, int(one + 9.5 * per_graph))
    assert one < ten <= one + 10 * per_graph
    tiny = lib.segnn_edge_layer_gemm_workspace(64, N, n, 1, 1024)  # the fixed part grows a little with B (per-chunk rows)
    assert one <= tiny < one + (1 << 20)
    # forward keeps 11 n floats per row: smaller than backward
    assert lib.segnn_edge_layer_gemm_workspace(64, N, n, 0, 0) < all64
    # unsupported multiplicities are refused
    assert lib.segnn_edge_layer_gemm_workspace(1, N, 25, 1, 0) == -1
    assert lib.segnn_edge_layer_gemm_workspace(1, N, 100, 1, 0) == -1


def test_tn_gemm_workspace_covers_the_split_k_partials():
    for K, M, N in [(1, 4, 4), (1000, 128, 192), (10 ** 7, 192, 288), (3 * 10 ** 6, 64, 64)]:
        b = lib.segnn_gemm_tn_tf32x3_workspace(K, M, N)
        mblocks, NP = (M + 127) // 128, (N + 31) // 32 * 32
        assert b >= mblocks * 128 * NP * 4 and b % (mblocks * 128 * NP * 4) == 0
        splits = b // (mblocks * 128 * NP * 4)
        assert 1 <= splits <= 296 and splits <= (K + 63) // 64 or splits == 1
    assert lib.segnn_gemm_tn_tf32x3_workspace(100, 64, 513) == -1


def test_dispatch_thresholds(monkeypatch):
    # training: GEMM form at every size when n is a multiple of 4; no-grad fp32 forward: from 2^18 rows
    assert ops._use_gemm_form(64, 5, 96, training=True)
    assert not ops._use_gemm_form(64, 5, 96)
    assert ops._use_gemm_form(1, 1000, 64) and ops._use_gemm_form(1024, 100, 96)
    assert not ops._use_gemm_form(1, 1000, 25, training=True), "n = 25 (hidden 50) stays on the fused kernels"
    assert not ops._use_gemm_form(4, 1, 64, training=True)
    assert not ops._use_gemm_form(1, 5000, 64, training=True), "one graph must fit a chunk"
    monkeypatch.setattr(ops, "GEMM_FORM_MIN_ROWS_TRAINING", 1 << 62)
    assert not ops._use_gemm_form(1, 1000, 64, training=True)


def test_keep_rows_budget(monkeypatch):
    assert ops.gemm_form_keeps_rows(1, 1000, 64), "configuration 4: 2.8 GB per layer"
    assert not ops.gemm_form_keeps_rows(1024, 100, 96), "a cfg5-size training batch does not fit 4 GB per layer"
    monkeypatch.setattr(ops, "GEMM_FORM_KEEP_BYTES_PER_LAYER", 0)
    assert not ops.gemm_form_keeps_rows(1, 1000, 64)


@pytest.mark.parametrize("H,lmax_h,expect", [(192, 2, True), (32, 2, True), (64, 1, False), (192, 1, False)])
def test_lmax2_gemm_form_is_selected_by_the_irreps(H, lmax_h, expect):
    from segnn_b200.generic import Lmax2EdgePlan
    m = S.SEGNN(hidden_features=H, num_layers=1, lmax_h=lmax_h)
    assert Lmax2EdgePlan.supported(m.layers[0], m.hidden_irreps) == expect
