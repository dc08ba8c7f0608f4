"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs and weights.
Tolerances: bit-exact for the integer graph enumeration; 1e-5 relative (max-norm) per layer in fp32 mode and 2e-2 in
bf16 tensor-core mode (BASELINE.json north_star, check (b)); the fp16-operand modes ('fp16', and 'fp16p' with
packed-half producers on fp16 projections) are held to 2.5e-3."""
import os

import pytest
import torch

import segnn_b200 as S
from oracle import segnn_oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "segnn_small.pt")
TOL = {"fp32": 1e-5, "bf16": 2e-2, "fp16": 2.5e-3, "fp16p": 2.5e-3}


def rel(a, b):
    return float((a.double().cpu() - b).abs().max() / b.abs().max().clamp_min(1e-30))


def modes(n=32, num_nodes=1):
    """Compute modes that apply: tensor-core modes need n in {32, 64, 96}; the packed-half mode also an even N."""
    tc = S.ops.tc_available() and n in S.ops.TC_MULTIPLICITIES
    return ["fp32"] + (["bf16", "fp16"] if tc else []) + (["fp16p"] if tc and num_nodes % 2 == 0 else [])


def make_pair(H, L, seed=0, dtype=torch.float32):
    torch.manual_seed(seed)
    om = O.SEGNN(hidden_features=H, num_layers=L).eval()
    O.perturb_bn_buffers(om, seed=seed + 1)
    m = S.SEGNN(hidden_features=H, num_layers=L)
    m.load_state_dict(om.state_dict())
    return om, m.to(dtype).cuda().eval()


def gpu_graph(pos, vel, mass, B, N, dtype=torch.float32):
    return S.GraphBatch(pos=pos.reshape(-1, 3).to(dtype).cuda(), vel=vel.reshape(-1, 3).to(dtype).cuda(),
                        mass=mass.reshape(-1, 1).to(dtype).cuda(), num_graphs=B, n_nodes=N)


@pytest.mark.parametrize("B,N", [(1, 2), (2, 3), (3, 5), (7, 100), (1, 1000)])
def test_edge_enumeration_bit_exact(B, N):
    got = S.build_graph_with_knn(None, B, N, "cuda", N - 1).cpu()
    assert got.dtype == torch.int64
    assert torch.equal(got, O.fully_connected_edge_index(B, N))


def test_graph_builder_errors():
    with pytest.raises(ValueError):
        S.build_graph_with_knn(None, 2, 5, "cuda", 5)
    with pytest.raises(ValueError, match="positions"):  # the kNN branch needs the node positions
        S.build_graph_with_knn(None, 2, 5, "cuda", 3)
    knn = S.build_graph_with_knn(torch.randn(10, 3), 2, 5, "cuda", 3)
    assert knn.shape == (2, 30) and int(knn[0].max()) == 9
    assert S.build_graph_with_knn(None, 0, 5, "cuda", None).shape == (2, 0)


@pytest.mark.parametrize("B,N", [(3, 5), (2, 100), (1, 300)])
def test_o3_transform(B, N):
    pos, vel, mass = O.synthetic_system(B, N, seed=11)
    ref = O.make_graph(pos.reshape(-1, 3), vel.reshape(-1, 3), mass.reshape(-1, 1), B, N)
    g = S.O3Transform(1)(gpu_graph(pos, vel, mass, B, N))
    ref_attr = ref.node_attr.clone()
    ref_attr[:, 0] = 1.0  # K1 applies segnn.py:148 directly
    assert float((g.node_attr.double().cpu() - ref_attr).abs().max()) < 2e-6
    assert float((g.x.double().cpu() - ref.x).abs().max()) < 2e-6 * float(ref.x.abs().max())
    assert torch.equal(g.edge_index.cpu(), ref.edge_index)
    assert float((g.edge_attr.double().cpu() - ref.edge_attr).abs().max()) < 2e-6
    add = g.additional_message_features.double().cpu()
    assert float((add - ref.additional_message_features).abs().max()) < 1e-5


@pytest.mark.parametrize("H,L,B,N", [(64, 4, 100, 5), (192, 6, 64, 5), (192, 6, 2, 100), (128, 2, 1, 37),
                                       (50, 2, 3, 9), (64, 1, 1, 2), (192, 6, 3, 6), (128, 3, 2, 38), (64, 2, 5, 12),
                                       (64, 2, 1, 300), (192, 2, 1, 300)])
def test_segnn_per_layer_parity(H, L, B, N):
    om, m = make_pair(H, L, seed=H + N)
    pos, vel, mass = O.synthetic_system(B, N, seed=5)
    with torch.no_grad():
        ref, ref_layers = om(O.make_graph(pos.reshape(-1, 3), vel.reshape(-1, 3), mass.reshape(-1, 1), B, N),
                             return_layers=True)
        for mode in modes(m.n, N):
            m.compute_mode = mode
            out, layers = m(gpu_graph(pos, vel, mass, B, N), return_layers=True)
            print(f"[{mode}] H={H} N={N}: per-layer rel err", [f"{rel(a, b):.2e}" for a, b in zip(layers, ref_layers)],
                  f"out {rel(out, ref):.2e}")
            # 16-bit operand rounding is coherent across the N - 1 messages of a receiver, so its error grows like
            # sqrt(N): the per-mode figures are for N <= 100; at N = 300 the fp16 modes measure
            # 3.2e-3 .. 4.2e-3 (held to 6e-3, inside north_star's 2e-2) and bf16 1.3e-2 .. 2.5e-2 (held to 4e-2, which
            # is why graphs with N > 100 run the fp16 modes)
            tol = TOL[mode] if N <= 100 or mode == "fp32" else (4e-2 if mode == "bf16" else 6e-3)
            for i, (a, b) in enumerate(zip(layers, ref_layers)):
                assert rel(a, b) < tol, f"{mode} layer {i}: {rel(a, b)}"
            assert rel(out, ref) < tol, f"{mode} output: {rel(out, ref)}"


def test_packed_half_mode_refuses_odd_graph_size():
    if not S.ops.tc_available():
        pytest.skip("tensor-core kernels not built")
    _, m = make_pair(64, 1)
    m.compute_mode = "fp16p"
    pos, vel, mass = O.synthetic_system(2, 5, seed=1)
    with torch.no_grad(), pytest.raises(RuntimeError, match="even number of bodies"):
        m(gpu_graph(pos, vel, mass, 2, 5))


@pytest.mark.parametrize("nodes,n_in,n_out", [(18, 96, 576), (500, 96, 288), (1000, 32, 192), (262, 64, 384)])
def test_node_gemm_16bit_output_modes_match_fp32_output(nodes, n_in, n_out):
    """segnn_node_gemm_tc_pair16 (fp16, node pairs interleaved) and segnn_node_gemm_tc_out16 (fp16 rows) against the
    fp32-output tensor-core GEMM on the same fp16 operands: equal up to the fp16 rounding of the stored result."""
    if not S.ops.tc_available():
        pytest.skip("tensor-core kernels not built")
    ops = S.ops
    gen = torch.Generator(device="cpu").manual_seed(nodes + n_out)
    x = torch.randn(nodes, 4, n_in, generator=gen).cuda()
    w = dict(w_s=(torch.randn(n_in, n_out, generator=gen) / n_in ** 0.5).cuda(),
             w_v=(torch.randn(n_in, n_out, generator=gen) / n_in ** 0.5).cuda(), operand=1)
    w["wt_s"], w["wt_v"] = ops.pack_node_weight_tc(w["w_s"], 1), ops.pack_node_weight_tc(w["w_v"], 1)
    split = (n_out // 64) * 32  # split and n_bias must be multiples of 32
    bias = torch.randn(split, generator=gen).cuda()
    ref0, ref1 = ops.node_gemm(x, None, w, n_out, bias=bias, n_bias=split, split=split, tc=True)
    y0, y1 = ops.node_gemm_pair16(x, w, n_out, bias, split, split)
    for y, ref in ((y0, ref0), (y1, ref1)):  # [nodes/2, 4, cols, 2] -> [nodes, 4, cols]
        got = y.permute(0, 3, 1, 2).reshape(nodes, 4, -1).float()
        assert float((got - ref).abs().max()) <= 1e-3 * float(ref.abs().max()) + 1e-6
    full = ops.node_gemm(x, None, w, n_out, tc=True)
    rows16 = ops.node_gemm_out16(x, None, w, n_out)
    assert rows16.dtype == torch.float16 and rows16.shape == full.shape
    assert float((rows16.float() - full).abs().max()) <= 1e-3 * float(full.abs().max()) + 1e-6
    # the combine pass reads the fp16 rows through its own entry point
    n = n_out // 3 if n_out % 3 == 0 else None
    if n is not None and n % 2 == 0:
        attr = torch.randn(nodes, 4, generator=gen).cuda()
        a = ops.tp_combine(full, attr, n, True)
        b = ops.tp_combine(rows16, attr, n, True)
        assert float((a - b).abs().max()) <= 2e-3 * float(a.abs().max()) + 1e-6


def test_double_precision_module_and_precomputed_attributes():
    """The reference default is float64 (.double() model, float64 graph): accept it, compute in fp32, return float64."""
    om, m = make_pair(64, 2, dtype=torch.float64)
    B, N = 4, 5
    pos, vel, mass = O.synthetic_system(B, N, seed=2)
    g = S.O3Transform(1)(gpu_graph(pos, vel, mass, B, N, dtype=torch.float64))
    with torch.no_grad():
        out = m(g)
        ref = om(O.make_graph(pos.reshape(-1, 3), vel.reshape(-1, 3), mass.reshape(-1, 1), B, N))
    assert out.dtype == torch.float64 and rel(out, ref) < 1e-5


def test_golden_fixture():
    gold = torch.load(GOLDEN)
    cfg = gold["config"]
    m = S.SEGNN(hidden_features=cfg["hidden_features"], num_layers=cfg["num_layers"])
    m.load_state_dict(gold["state_dict"])
    m = m.float().cuda().eval()
    B, N = cfg["B"], cfg["N"]
    with torch.no_grad():
        for mode in modes(m.n):
            m.compute_mode = mode
            out, layers = m(gpu_graph(gold["pos"], gold["vel"], gold["mass"], B, N), return_layers=True)
            assert rel(out, gold["pred"]) < TOL[mode]
            for a, b in zip(layers, gold["layers"]):
                assert rel(a, b) < TOL[mode]
    assert torch.equal(S.build_graph_with_knn(None, B, N, "cuda", None).cpu(), gold["edge_index"])


def test_module_level_tensor_products():
    torch.manual_seed(0)
    n, rows = 32, 50
    for cls_s, cls_o, in1 in [(S.O3TensorProductSwishGate, O.O3TensorProductSwishGate, f"{n}x0e+{n}x1o"),
                              (S.O3TensorProduct, O.O3TensorProduct, f"{n}x0e+{n}x1o"),
                              (S.O3TensorProductSwishGate, O.O3TensorProductSwishGate,
                               f"{n}x0e+{n}x1o+{n}x0e+{n}x1o")]:
        ot = cls_o(in1, f"{n}x0e+{n}x1o", "1x0e+1x1o")
        st = cls_s(in1, f"{n}x0e+{n}x1o", "1x0e+1x1o")
        st.load_state_dict(ot.state_dict())
        st = st.cuda()
        x = torch.randn(rows, O.Irreps(in1).dim, dtype=torch.float64)
        a = torch.randn(rows, 4, dtype=torch.float64)
        with torch.no_grad():
            assert rel(st(x.float().cuda(), a.float().cuda()), ot(x, a)) < 1e-5


def test_standalone_layer_matches_oracle():
    torch.manual_seed(0)
    n, B, N = 32, 3, 6
    h = f"{n}x0e+{n}x1o"
    ol = O.SEGNNLayer(h, h, h, "1x0e+1x1o", "1x0e+1x1o", norm="batch", additional_message_irreps="2x0e").eval()
    O.perturb_bn_buffers(ol)
    sl = S.SEGNNLayer(h, h, h, "1x0e+1x1o", "1x0e+1x1o", norm="batch", additional_message_irreps="2x0e")
    sl.load_state_dict(ol.state_dict())
    sl = sl.cuda().eval()
    pos, vel, mass = O.synthetic_system(B, N, seed=8)
    g = O.make_graph(pos.reshape(-1, 3), vel.reshape(-1, 3), mass.reshape(-1, 1), B, N)
    g.node_attr[:, 0] = 1.0
    x = torch.randn(B * N, 4 * n, dtype=torch.float64)
    with torch.no_grad():
        ref = ol(x, g.edge_index, g.edge_attr, g.node_attr, None, g.additional_message_features)
        out = sl(x.float().cuda(), None, None, g.node_attr.float().cuda(), None, None,
                 pos=pos.reshape(-1, 3).float().cuda(), mass=mass.reshape(-1).float().cuda(), num_graphs=B, n_nodes=N)
    assert rel(out, ref) < 1e-5


@pytest.mark.parametrize("h,attr,norm", [("24x0e+24x1o", "1x0e+1x1o", "batch"), ("10x0e+10x1o+10x2e", "1x0e+1x1o", "batch"),
                                         ("12x0e+12x1o+12x2e", "1x0e+1x1o+1x2e", "instance"),
                                         ("16x0e+16x1o", "1x0e+1x1o+1x2e", None)])
def test_standalone_layer_on_explicit_edge_list_matches_oracle(h, attr, norm):
    """SEGNNLayer.forward called the reference's way (segnn.py:239-304: explicit edge_index, edge_attr, node_attr, batch,
    additional_message_features; here a kNN edge list and ARBITRARY attribute values) runs the generic kernels."""
    torch.manual_seed(1)
    B, N, k = 3, 7, 3
    ol = O.SEGNNLayer(h, h, h, attr, attr, norm=norm, additional_message_irreps="2x0e").eval()
    if norm == "batch":
        O.perturb_bn_buffers(ol)
    sl = S.SEGNNLayer(h, h, h, attr, attr, norm=norm, additional_message_irreps="2x0e")
    sl.load_state_dict(ol.state_dict())
    sl = sl.cuda().eval()
    D, d = O.Irreps(h).dim, O.Irreps(attr).dim
    ei = O.knn_edge_index(torch.randn(B * N, 3, dtype=torch.float64), B, N, k)
    E = ei.shape[1]
    x = torch.randn(B * N, D, dtype=torch.float64)
    ea, na = torch.randn(E, d, dtype=torch.float64), torch.randn(B * N, d, dtype=torch.float64)
    add = torch.randn(E, 2, dtype=torch.float64)
    batch = torch.arange(B).repeat_interleave(N)
    with torch.no_grad():
        ref = ol(x, ei, ea, na, batch, add)
        out = sl(x.float().cuda(), ei.cuda(), ea.float().cuda(), na.float().cuda(), batch.cuda(), add.float().cuda())
        out2 = sl(x.float().cuda(), ei.cuda(), ea.float().cuda(), na.float().cuda(), batch.cuda(), add.float().cuda())
    assert out.shape == ref.shape and rel(out, ref) < 1e-5, rel(out, ref)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("use_graph", [False, True])
def test_rollout_matches_oracle(use_graph):
    om, m = make_pair(64, 4, seed=3)
    B, N, steps = 6, 5, 8
    pos, vel, mass = O.synthetic_system(B, N, seed=9)
    ref_loc, ref_vel = O.rollout(om, pos, vel, mass, steps)
    roll = S.SelfFeedRollout(m, B, N, "cuda", max_frames=steps + 1, use_cuda_graph=use_graph)
    roll.reset(pos, vel, mass)
    tp, tv = roll.run(steps)
    got_loc = tp.reshape(steps + 1, B, N, 3).permute(1, 0, 2, 3)
    got_vel = tv.reshape(steps + 1, B, N, 3).permute(1, 0, 2, 3)
    assert rel(got_loc, ref_loc) < 5e-5 and rel(got_vel, ref_vel) < 5e-5
    assert int(roll.frame.item()) == steps + 1


def test_run_inference_api(tmp_path):
    om, m = make_pair(64, 2, seed=4)
    B, N, T = 3, 5, 6
    pos, vel, mass = O.synthetic_system(B, N, seed=10)
    truth_loc, truth_vel = O.rollout(om, pos, vel, mass, T - 1)  # any [B,T,N,3] ground truth will do
    d, loc, velc = S.run_inference("segnn", None, model=m, save_dir=str(tmp_path), print_step=False,
                                   ground_truth=(truth_loc, truth_vel, mass))
    assert loc.shape == (2, B, T, N, 3) and velc.shape == (2, B, T, N, 3)
    assert sorted(os.listdir(d))[0] == "loc_actual_sim_0.npy" and len(os.listdir(d)) == 4 * B
    assert float(abs(loc[1] - loc[0]).max()) < 1e-4 * float(abs(loc[0]).max())
    with pytest.raises(ValueError):
        S.run_inference("ponita", None, model=m, ground_truth=(truth_loc, truth_vel, mass))


def _rotation(seed):
    g = torch.Generator().manual_seed(seed)
    q, r = torch.linalg.qr(torch.randn(3, 3, generator=g, dtype=torch.float64))
    q = q * torch.sign(torch.diagonal(r))
    if torch.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q


def test_equivariance_and_permutation_at_scale():
    """Size-independent properties on a workload too big for the oracle: rotating the system rotates the
    prediction (x_in is passed rotated, which removes the reference's pos.mean(1) quirk from the test path), and
    permuting the bodies of every graph permutes the prediction."""
    _, m = make_pair(192, 6, seed=6)
    B, N = 8, 100
    pos, vel, mass = O.synthetic_system(B, N, seed=12)
    R = _rotation(1).float().cuda()
    p, v, ms = pos.reshape(-1, 3).float().cuda(), vel.reshape(-1, 3).float().cuda(), mass.reshape(-1).float().cuda()
    with torch.no_grad():
        x0, a0 = S.ops.prep(p, v, B, N)
        x0[:, :3] = p  # centred input without the quirk
        base = m.forward_state(p, v, ms, B, N, x_in=x0, node_attr=a0)
        pr, vr = (p @ R.T).contiguous(), (v @ R.T).contiguous()
        x1, a1 = S.ops.prep(pr, vr, B, N)
        x1[:, :3] = pr
        rot = m.forward_state(pr, vr, ms, B, N, x_in=x1, node_attr=a1)
        expect = torch.cat([base[:, :3] @ R.T, base[:, 3:] @ R.T], dim=1)
        assert float((rot - expect).abs().max() / expect.abs().max()) < 2e-4
        perm = torch.stack([torch.randperm(N) + b * N for b in range(B)]).reshape(-1).cuda()
        out_p = m.forward_state(p[perm].contiguous(), v[perm].contiguous(), ms[perm].contiguous(), B, N)
        out = m.forward_state(p, v, ms, B, N)
        assert float((out_p - out[perm]).abs().max() / out.abs().max()) < 2e-4


# ---- training: train-mode BatchNorm forward + hand-written backward vs autograd through the oracle ----------------
def _grad_case(H, L, B, N, bn_train, seed=0):
    torch.manual_seed(seed)
    om = O.SEGNN(hidden_features=H, num_layers=L)
    O.perturb_bn_buffers(om, seed=seed + 1)
    om.train(bn_train)
    m = S.SEGNN(hidden_features=H, num_layers=L)
    m.load_state_dict(om.state_dict())
    m = m.float().cuda().train(bn_train)
    pos, vel, mass = O.synthetic_system(B, N, seed=seed + 2)
    y = torch.randn(B * N, 6, dtype=torch.float64)
    ref = om(O.make_graph(pos.reshape(-1, 3), vel.reshape(-1, 3), mass.reshape(-1, 1), B, N))
    loss_ref = O.target_common_loss(ref, y)
    loss_ref.backward()
    pred = m(gpu_graph(pos, vel, mass, B, N))
    loss = O.target_common_loss(pred, y.float().cuda())
    loss.backward()
    return om, m, ref.detach(), pred.detach(), float(loss_ref), float(loss)


@pytest.mark.parametrize("edge_form", ["fused", "default"])
@pytest.mark.parametrize("H,L,B,N,bn_train", [(64, 2, 4, 5, True), (64, 2, 4, 5, False), (192, 1, 2, 20, True),
                                                (128, 2, 1, 33, True), (50, 1, 3, 6, True), (64, 1, 1, 200, True)])
def test_training_gradients_match_oracle(H, L, B, N, bn_train, edge_form, monkeypatch):
    """fp32 kernels vs float64 autograd: prediction 1e-5 (north_star fp32 tolerance); gradients within 1e-4 of each
    parameter's max-norm plus 1e-5 of the largest gradient in the model. The second term covers parameters that sit
    upstream of a train-mode BatchNorm (message_layer_2 / update_layer_2 biases): their gradient is a sum over all
    rows of terms that cancel almost exactly (the BatchNorm backward removes the mean), so its fp32 rounding error is
    set by the size of the terms, not of the result."""
    if edge_form == "fused":  # the fused fp32 edge kernels (training otherwise runs the edge layers in GEMM form)
        monkeypatch.setattr(S.ops, "GEMM_FORM_MIN_ROWS_TRAINING", 1 << 62)
    om, m, ref, pred, loss_ref, loss = _grad_case(H, L, B, N, bn_train)
    assert rel(pred, ref) < 1e-5
    assert abs(loss - loss_ref) < 1e-5 * abs(loss_ref)
    worst = 0.0
    top = max(float(a.grad.abs().max()) for a in om.parameters())
    for (k, a), (k2, b) in zip(om.named_parameters(), m.named_parameters()):
        assert k == k2 and b.grad is not None, k
        scale = float(a.grad.abs().max())
        err = float((a.grad - b.grad.double().cpu()).abs().max())
        if scale > 1e-9:
            worst = max(worst, err / scale)
        assert err <= 1e-4 * scale + 1e-5 * top, f"{k}: {err} vs scale {scale} (top {top})"
    print(f"H={H} N={N} bn_train={bn_train}: worst gradient rel err {worst:.2e}")
    sd = om.state_dict()
    for k, b in m.state_dict().items():
        if "running" in k:
            assert float((sd[k] - b.double().cpu()).abs().max()) < 1e-5 * (1 + float(sd[k].abs().max())), k


def test_train_mode_forward_without_grad_matches_oracle():
    """trainer.py never calls model.eval() before its in-training rollout (SURVEY note 6): train-mode BatchNorm must
    also work under no_grad, and must update the running statistics."""
    torch.manual_seed(1)
    om = O.SEGNN(hidden_features=64, num_layers=2).train()
    m = S.SEGNN(hidden_features=64, num_layers=2)
    m.load_state_dict(om.state_dict())
    m = m.float().cuda().train()
    B, N = 5, 5
    pos, vel, mass = O.synthetic_system(B, N, seed=4)
    with torch.no_grad():
        ref = om(O.make_graph(pos.reshape(-1, 3), vel.reshape(-1, 3), mass.reshape(-1, 1), B, N))
        out = m(gpu_graph(pos, vel, mass, B, N))
    assert rel(out, ref) < 1e-5
    assert rel(m.layers[0].message_norm.running_var, om.layers[0].message_norm.running_var) < 1e-5


@pytest.mark.parametrize("H,L,B,N", [(64, 2, 6, 6), (192, 3, 4, 20), (128, 2, 2, 38)])
def test_train_mode_batchnorm_forward_on_the_tensor_core_kernels(H, L, B, N):
    """The reference evaluates its rollouts with train-mode (batch-statistic) BatchNorm (trainer.py:373, 929-942). In
    the packed-half mode that forward runs on the tcgen05 kernels: K3 returns raw sums + per-receiver moments, the
    statistics are float64 column sums. Running statistics must move like
    e3nn's, and a following eval-mode forward must see them (the folded-BatchNorm cache is keyed on buffer versions)."""
    if not S.ops.tc_available():
        pytest.skip("tensor-core mode not built")
    torch.manual_seed(H + N)
    om = O.SEGNN(hidden_features=H, num_layers=L).train()
    m = S.SEGNN(hidden_features=H, num_layers=L, compute_mode="fp16p")
    m.load_state_dict(om.state_dict())
    m = m.float().cuda().train()
    pos, vel, mass = O.synthetic_system(B, N, seed=N)
    g = gpu_graph(pos, vel, mass, B, N)
    og = O.make_graph(pos.reshape(-1, 3), vel.reshape(-1, 3), mass.reshape(-1, 1), B, N)
    with torch.no_grad():
        m.eval()
        stale = m(g)  # fills the eval-mode operand cache with the initial running statistics
        m.train()
        ref, out = om(og), m(g)
        print(f"train-mode BatchNorm forward, fp16p H={H} N={N}: rel err {rel(out, ref):.2e}")
        # batch statistics of a few hundred rows divide by standard deviations that the fp16 rounding has moved:
        # measured 1.1e-3 .. 3.2e-3 on these shapes, held to 5e-3 (north_star's budget for this mode class is 2e-2)
        assert rel(out, ref) < 2 * TOL["fp16p"]
        for k, a in om.state_dict().items():
            if "running" in k:
                b = m.state_dict()[k].double().cpu()
                assert float((a - b).abs().max()) < 2e-3 * (1 + float(a.abs().max())), k
        om.eval()
        m.eval()
        ref_e, out_e = om(og), m(g)
        assert rel(out_e, ref_e) < 2 * TOL["fp16p"], "eval forward after a train-mode forward uses the updated statistics"
        # ... and it differs from the forward with the initial statistics
        assert float((out_e - stale).abs().max() / stale.abs().max()) > 1e-3


@pytest.mark.parametrize("use_graph", [False, True])
def test_train_step_loss_trajectory_matches_oracle(use_graph):
    """Five optimisation steps (AdamW + Noam schedule, trainer.py:170-195) on a fixed batch: the loss trajectory of the
    CUDA path must follow float64 training of the oracle. (Parameters themselves are not compared: biases in front of
    a train-mode BatchNorm have exactly-zero true gradients, and Adam turns their fp32 rounding noise into +-lr steps
    that do not change the function.)"""
    torch.manual_seed(3)
    H, L, B, N = 64, 2, 8, 5
    om = O.SEGNN(hidden_features=H, num_layers=L).train()
    m = S.SEGNN(hidden_features=H, num_layers=L)
    m.load_state_dict(om.state_dict())
    m = m.float().cuda().train()
    pos, vel, mass = O.synthetic_system(B, N, seed=7)
    y = torch.randn(B * N, 6, dtype=torch.float64)
    hp = dict(learning_rate=1.0, learning_rate_factor=4000.0, learning_rate_warmup_steps=4000)
    ts = S.TrainStep(m, B, N, use_cuda_graph=use_graph, clip_gradients_norm=10.0, **hp)
    opt = torch.optim.AdamW(om.parameters(), weight_decay=1e-8, lr=1.0, betas=(0.9, 0.98), eps=1e-9)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: S.noam_rate(s, H, 4000.0, 4000))
    g = O.make_graph(pos.reshape(-1, 3), vel.reshape(-1, 3), mass.reshape(-1, 1), B, N)
    ref_losses, losses = [], []
    for _ in range(5):
        opt.zero_grad()
        loss = O.target_common_loss(om(g), y)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(om.parameters(), 10.0)
        opt.step()
        sched.step()
        ref_losses.append(float(loss))
        losses.append(float(ts.step(pos, vel, mass, y)))
    print("loss trajectory", losses, ref_losses)
    assert ref_losses[-1] < ref_losses[0]
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) < 2e-4 * abs(b)
    assert rel(m.layers[0].feature_norm.running_var, om.layers[0].feature_norm.running_var) < 1e-4


# ---- rollout macros ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,N,T", [(3, 5, 7), (2, 100, 4), (1, 300, 2)])
def test_macro_kernel_matches_reference_definitions(B, N, T):
    import numpy as np
    gen = torch.Generator().manual_seed(B * N + T)
    loc = torch.randn(B, T, N, 3, generator=gen, dtype=torch.float64) * 2.0
    vel = torch.randn(B, T, N, 3, generator=gen, dtype=torch.float64)
    G, soft = 2.0, 0.2
    kin_ref, pot_ref, series = O.nbody_energies(loc.numpy(), vel.numpy(), G, soft)
    mom_ref = O.momentum_magnitude(vel.numpy())
    tp = loc.permute(1, 0, 2, 3).reshape(T, B * N, 3).float().cuda()
    tv = vel.permute(1, 0, 2, 3).reshape(T, B * N, 3).float().cuda()
    kin, pot, mom = S.macros.energy_momentum(tp, tv, B, N, G, soft)
    assert np.abs(kin.cpu().numpy().T - kin_ref).max() < 1e-5 * np.abs(kin_ref).max()
    assert np.abs(pot.cpu().numpy().T - pot_ref).max() < 1e-5 * np.abs(pot_ref).max()
    assert np.abs(mom.cpu().numpy().T - mom_ref).max() < 1e-5 * max(np.abs(mom_ref).max(), 1.0)
    got = S.macros.nbody_energies(tp, tv, B, N, G, soft)
    for k in ("potential", "kinetic", "total"):
        assert np.abs(got[k] - series[k]).max() < 1e-5 * np.abs(series[k]).max()


@pytest.mark.parametrize("mode,N", [("fp32", 5), ("bf16", 5), ("fp16p", 6)])
def test_rollout_macros_within_reference_statistical_tolerance(mode, N):
    """BASELINE north_star check (c): energy / momentum macros of the CUDA rollout against the oracle rollout on the same
    initial conditions and weights: two-sample KS p >= 0.05 per macro and Fisher-combined (the reference's acceptance
    threshold, figures/combined_pvalues_summary.csv `time_to_p_ge_0.05`), and the total-energy ratio stays within
    [1/2.5, 2.5] for every step (trainer.py:27,692-701)."""
    import numpy as np
    if mode != "fp32" and not (S.ops.tc_available()):
        pytest.skip("tensor-core mode not built")
    om, m = make_pair(64, 3, seed=21)
    m.compute_mode = mode
    B, steps = 24, 30
    pos, vel, mass = O.synthetic_system(B, N, seed=31, charged=False)
    ref_loc, ref_vel = O.rollout(om, pos, vel, mass, steps)
    roll = S.SelfFeedRollout(m, B, N, "cuda", max_frames=steps + 1)
    roll.reset(pos, vel, mass)
    tp, tv = roll.run(steps)
    G, soft = 2.0, 0.2
    kin, pot, mom = S.macros.energy_momentum(tp, tv, B, N, G, soft)
    kin_ref, pot_ref, series_ref = O.nbody_energies(ref_loc.numpy(), ref_vel.numpy(), G, soft)
    mom_ref = O.momentum_magnitude(ref_vel.numpy())
    ps = [S.macros.ks_p(kin.cpu().numpy().T[:, -1], kin_ref[:, -1]),
          S.macros.ks_p(pot.cpu().numpy().T[:, -1], pot_ref[:, -1]),
          S.macros.ks_p(mom.cpu().numpy().T[:, -1], mom_ref[:, -1]),
          S.macros.ks_p(kin.cpu().numpy().T.ravel(), kin_ref.ravel()),
          S.macros.ks_p(pot.cpu().numpy().T.ravel(), pot_ref.ravel())]
    combined = S.macros.combine_pvalues_fisher(ps)
    print(f"[{mode}] KS p-values {ps} combined {combined}")
    assert min(ps) >= 0.05 and combined >= 0.05
    got = S.macros.nbody_energies(tp, tv, B, N, G, soft)
    assert S.macros.energy_ratio_steps(got["total"], series_ref["total"]) == steps + 1


# ---- generic-irreps path (lmax_h = 2, BASELINE config 3) -------------------------------------------------------------
@pytest.mark.parametrize("l2_rows", [True, False])
@pytest.mark.parametrize("H,lmax_h,L,B,N", [(32, 2, 2, 2, 6), (192, 2, 1, 1, 10), (64, 2, 3, 3, 5), (64, 1, 2, 3, 5),
                                             (32, 2, 2, 8, 17), (192, 2, 1, 1, 50), (192, 2, 2, 3, 33)])
def test_generic_irreps_path_matches_oracle(H, lmax_h, L, B, N, l2_rows, monkeypatch):
    """Per-layer parity of the generic fp32 kernels (any hidden irreps; lmax_h = 2 is BASELINE config 3) at the fp32
    tolerance 1e-5; for lmax_h = 1 the same path is selected with compute_mode='generic' and must also agree with the
    fused kernels."""
    # l2_rows: the edge part of an lmax_h = 2 layer in GEMM form (csrc/segnn_l2_rows.cu, the default) or through the
    # table-driven generic kernels
    import segnn_b200.generic as G
    monkeypatch.setattr(G, "USE_L2_ROWS", l2_rows)
    torch.manual_seed(H + lmax_h)
    om = O.SEGNN(hidden_features=H, num_layers=L, lmax_h=lmax_h).eval()
    O.perturb_bn_buffers(om, seed=5)
    m = S.SEGNN(hidden_features=H, num_layers=L, lmax_h=lmax_h, compute_mode="generic")
    m.load_state_dict(om.state_dict())
    m = m.float().cuda().eval()
    pos, vel, mass = O.synthetic_system(B, N, seed=6)
    with torch.no_grad():
        ref, ref_layers = om(O.make_graph(pos.reshape(-1, 3), vel.reshape(-1, 3), mass.reshape(-1, 1), B, N),
                             return_layers=True)
        out, layers = m(gpu_graph(pos, vel, mass, B, N), return_layers=True)
        assert m._generic.use_l2_rows == (l2_rows and lmax_h == 2)
        print(f"generic H={H} lmax_h={lmax_h} l2_rows={l2_rows}: per-layer", [f"{rel(a, b):.2e}" for a, b in zip(layers, ref_layers)],
              f"out {rel(out, ref):.2e}")
        for a, b in zip(layers, ref_layers):
            assert rel(a, b) < 1e-5
        assert rel(out, ref) < 1e-5
        if lmax_h == 1:
            m.compute_mode = "fp32"
            fused = m(gpu_graph(pos, vel, mass, B, N))
            assert float((fused - out).abs().max() / out.abs().max()) < 1e-5


@pytest.mark.parametrize("B,N", [(3, 5), (2, 40), (1, 2)])
def test_o3_transform_lmax_attr2_matches_oracle(B, N):
    """O3Transform with lmax_attr = 2 (o3_building_blocks.py:230-278): l <= 2 'integral' harmonics of the edge vectors,
    node attributes = mean over the senders + harmonics of the velocity (segnn_edge_attr_lmax / segnn_prep_fwd_lmax)."""
    pos, vel, mass = O.synthetic_system(B, N, seed=11)
    vel[0, 0] = 0.0  # a body at rest: F.normalize gives the zero vector, every l >= 1 harmonic is 0
    ref = O.make_graph(pos.reshape(-1, 3), vel.reshape(-1, 3), mass.reshape(-1, 1), B, N, 2)
    g = S.O3Transform(2)(gpu_graph(pos, vel, mass, B, N))
    assert g.node_attr.shape == (B * N, 9) and g.edge_attr.shape == (B * N * (N - 1), 9)
    assert torch.equal(g.edge_index.cpu(), ref.edge_index)
    assert rel(g.x, ref.x) < 2e-6 and rel(g.edge_attr, ref.edge_attr) < 2e-6
    assert rel(g.additional_message_features, ref.additional_message_features) < 2e-6
    assert float((g.node_attr.double().cpu()[:, 1:] - ref.node_attr[:, 1:]).abs().max()) < 2e-6
    assert torch.all(g.node_attr[:, 0] == 1.0) or rel(g.node_attr[:, 0], ref.node_attr[:, 0]) < 2e-6


@pytest.mark.parametrize("lmax_attr,k,mode", [(1, None, "fp32"), (1, None, "generic"), (2, None, "fp32"), (1, 3, "fp32")])
def test_use_force_input_matches_oracle(lmax_attr, k, mode):
    """O3Transform(use_force_input=True) (o3_building_blocks.py:267-271: node_attr += Y(force)) and a model forward on
    the graph's own x / node_attr: fused kernels, generic kernels, lmax_attr = 2, and a kNN edge list."""
    from types import SimpleNamespace
    torch.manual_seed(21)
    B, N, H, L = 3, 8, 32, 2
    om = O.SEGNN(hidden_features=H, num_layers=L, lmax_attr=lmax_attr).eval()
    O.perturb_bn_buffers(om, seed=5)
    m = S.SEGNN(hidden_features=H, num_layers=L, lmax_attr=lmax_attr, compute_mode=mode)
    m.load_state_dict(om.state_dict())
    m = m.float().cuda().eval()
    pos, vel, mass = O.synthetic_system(B, N, seed=6)
    pos, vel, mass = pos.reshape(-1, 3), vel.reshape(-1, 3), mass.reshape(-1, 1)
    force = torch.randn(B * N, 3, dtype=torch.float64)
    og = SimpleNamespace(pos=pos, vel=vel, mass=mass, force=force, batch=torch.arange(B).repeat_interleave(N),
                         edge_index=O.build_graph_with_knn(pos, B, N, None, N - 1 if k is None else k))
    og = O.o3_transform(og, lmax_attr, use_force_input=True)
    g = gpu_graph(pos, vel, mass, B, N)
    g.force = force.float().cuda()
    if k is not None:
        g.edge_index = S.build_graph_with_knn(pos.cuda(), B, N, "cuda", k)
    g = S.O3Transform(lmax_attr, use_force_input=True)(g)
    assert float((g.node_attr.double().cpu()[:, 1:] - og.node_attr[:, 1:]).abs().max()) < 2e-6
    plain = S.O3Transform(lmax_attr)(gpu_graph(pos, vel, mass, B, N)).node_attr
    assert k is not None or float((g.node_attr - plain).abs().max()) > 1e-2  # the force term is there
    with torch.no_grad():
        ref = om(og)
        out = m(g)
    assert rel(out, ref) < 1e-5, rel(out, ref)


@pytest.mark.parametrize("H,lmax_h,L,B,N", [(32, 1, 2, 3, 5), (64, 2, 2, 2, 6), (128, 1, 2, 2, 33), (96, 2, 1, 1, 40)])
def test_lmax_attr2_matches_oracle(H, lmax_h, L, B, N):
    """lmax_attr = 2 (steering attributes 1x0e + 1x1o + 1x2e, models/segnn/segnn.py:22,36,47): every tensor product of the
    model through the table-driven generic kernels with [5][5][5] couplings; per-layer parity at the fp32 tolerance."""
    torch.manual_seed(H + lmax_h)
    om = O.SEGNN(hidden_features=H, num_layers=L, lmax_h=lmax_h, lmax_attr=2).eval()
    O.perturb_bn_buffers(om, seed=5)
    m = S.SEGNN(hidden_features=H, num_layers=L, lmax_h=lmax_h, lmax_attr=2)
    m.load_state_dict(om.state_dict())
    m = m.float().cuda().eval()
    assert not m.fused
    pos, vel, mass = O.synthetic_system(B, N, seed=6)
    with torch.no_grad():
        ref, ref_layers = om(O.make_graph(pos.reshape(-1, 3), vel.reshape(-1, 3), mass.reshape(-1, 1), B, N, 2),
                             return_layers=True)
        out, layers = m(gpu_graph(pos, vel, mass, B, N), return_layers=True)
        print(f"lmax_attr=2 H={H} lmax_h={lmax_h}: per-layer", [f"{rel(a, b):.2e}" for a, b in zip(layers, ref_layers)],
              f"out {rel(out, ref):.2e}")
        for a, b in zip(layers, ref_layers):
            assert rel(a, b) < 1e-5
        assert rel(out, ref) < 1e-5
        # the reference's formulation (gathered message input) agrees with the hoisted one
        m._generic.hoist_message_layer_1 = False
        out2 = m(gpu_graph(pos, vel, mass, B, N))
        assert rel(out2, ref) < 1e-5
    with pytest.raises(NotImplementedError):  # inference only, like every generic-irreps configuration
        m.train()(gpu_graph(pos, vel, mass, B, N))


def test_lmax_attr2_rollout_matches_oracle():
    torch.manual_seed(10)
    om = O.SEGNN(hidden_features=32, num_layers=2, lmax_h=1, lmax_attr=2).eval()
    O.perturb_bn_buffers(om, seed=3)
    m = S.SEGNN(hidden_features=32, num_layers=2, lmax_h=1, lmax_attr=2)
    m.load_state_dict(om.state_dict())
    m = m.float().cuda().eval()
    B, N, steps = 3, 5, 6
    pos, vel, mass = O.synthetic_system(B, N, seed=4)
    ref_loc, ref_vel = O.rollout(om, pos, vel, mass, steps)
    roll = S.SelfFeedRollout(m, B, N, "cuda", max_frames=steps + 1)
    roll.reset(pos, vel, mass)
    tp, tv = roll.run(steps)
    got_loc = tp.reshape(steps + 1, B, N, 3).permute(1, 0, 2, 3)
    assert rel(got_loc, ref_loc) < 5e-5


@pytest.mark.parametrize("H,lmax_h,lmax_attr,L,B,N,k", [(64, 1, 1, 3, 3, 10, 4), (32, 2, 1, 2, 2, 7, 1), (48, 1, 2, 2, 2, 9, 7),
                                                       (192, 1, 1, 2, 4, 50, 6)])
def test_knn_graph_forward_matches_oracle(H, lmax_h, lmax_attr, L, B, N, k):
    """SEGNN on the kNN graphs of build_graph_with_knn (num_neighbors < N - 1; variable in-degree, possibly isolated
    targets): edge list bit-exact against the oracle's, per-layer parity at the fp32 tolerance, run-to-run identical."""
    torch.manual_seed(H + k)
    om = O.SEGNN(hidden_features=H, num_layers=L, lmax_h=lmax_h, lmax_attr=lmax_attr).eval()
    O.perturb_bn_buffers(om, seed=5)
    m = S.SEGNN(hidden_features=H, num_layers=L, lmax_h=lmax_h, lmax_attr=lmax_attr)
    m.load_state_dict(om.state_dict())
    m = m.float().cuda().eval()
    pos, vel, mass = O.synthetic_system(B, N, seed=8)
    og = O.make_graph(pos.reshape(-1, 3), vel.reshape(-1, 3), mass.reshape(-1, 1), B, N, lmax_attr, k)
    g = gpu_graph(pos, vel, mass, B, N)
    # the oracle's positions are float64; the edge list is built from the same values (float32 inputs would reorder ties)
    g.edge_index = S.build_graph_with_knn(pos.reshape(-1, 3).double().cuda(), B, N, "cuda", k)
    assert torch.equal(g.edge_index.cpu(), og.edge_index)
    with torch.no_grad():
        ref, ref_layers = om(og, return_layers=True)
        out, layers = m(g, return_layers=True)
        out2 = m(g)
    print(f"kNN H={H} lmax_h={lmax_h} lmax_attr={lmax_attr} k={k}: per-layer",
          [f"{rel(a, b):.2e}" for a, b in zip(layers, ref_layers)], f"out {rel(out, ref):.2e}")
    for a, b in zip(layers, ref_layers):
        assert rel(a, b) < 1e-5
    assert rel(out, ref) < 1e-5
    assert torch.equal(out, out2)


@pytest.mark.parametrize("use_graph", [False, True])
def test_knn_rollout_matches_oracle(use_graph):
    """Self-feed rollout with the kNN graph rebuilt from the predicted positions every step
    (helper_scripts/infer_self_feed.py:175), eager and as a replayed CUDA graph."""
    torch.manual_seed(12)
    om = O.SEGNN(hidden_features=32, num_layers=2, lmax_h=1).eval()
    O.perturb_bn_buffers(om, seed=3)
    m = S.SEGNN(hidden_features=32, num_layers=2, lmax_h=1)
    m.load_state_dict(om.state_dict())
    m = m.float().cuda().eval()
    B, N, k, steps = 2, 7, 3, 5
    pos, vel, mass = O.synthetic_system(B, N, seed=4)
    loc, v = [pos.double()], [vel.double()]
    with torch.no_grad():
        for _ in range(steps):
            g = O.make_graph(loc[-1].reshape(-1, 3), v[-1].reshape(-1, 3), mass.reshape(-1, 1).double(), B, N, 1, k)
            pred = om(g).reshape(B, N, 6)
            loc.append(loc[-1] + pred[..., :3])
            v.append(pred[..., 3:])
    roll = S.SelfFeedRollout(m, B, N, "cuda", max_frames=steps + 1, use_cuda_graph=use_graph, num_neighbors=k)
    roll.reset(pos, vel, mass)
    tp, tv = roll.run(steps)
    got = tp.reshape(steps + 1, B, N, 3).permute(1, 0, 2, 3)
    assert rel(got, torch.stack(loc, dim=1)) < 5e-5
    assert rel(tv.reshape(steps + 1, B, N, 3).permute(1, 0, 2, 3), torch.stack(v, dim=1)) < 5e-5


@pytest.mark.parametrize("H,lmax_h,L,B,N", [(64, 1, 3, 4, 9), (48, 2, 2, 3, 6), (192, 1, 2, 2, 40)])
def test_instance_norm_model_matches_oracle(H, lmax_h, L, B, N):
    """norm='instance' (segnn.py:226-237, 257-261; models/segnn/instance_norm.py on the node features of every layer, no
    message norm): per-layer parity at the fp32 tolerance, and a replayed CUDA-graph rollout step (no host sync)."""
    torch.manual_seed(H + lmax_h)
    om = O.SEGNN(hidden_features=H, num_layers=L, lmax_h=lmax_h, norm="instance").eval()
    with torch.no_grad():
        for layer in om.layers:
            layer.feature_norm.weight.uniform_(0.5, 1.5)
            layer.feature_norm.bias.uniform_(-0.3, 0.3)
    m = S.SEGNN(hidden_features=H, num_layers=L, lmax_h=lmax_h, norm="instance")
    m.load_state_dict(om.state_dict())
    m = m.float().cuda().eval()
    assert not m.fused and m.layers[0].message_norm is None
    pos, vel, mass = O.synthetic_system(B, N, seed=6)
    with torch.no_grad():
        ref, ref_layers = om(O.make_graph(pos.reshape(-1, 3), vel.reshape(-1, 3), mass.reshape(-1, 1), B, N),
                             return_layers=True)
        out, layers = m(gpu_graph(pos, vel, mass, B, N), return_layers=True)
    print(f"instance norm H={H} lmax_h={lmax_h}: per-layer", [f"{rel(a, b):.2e}" for a, b in zip(layers, ref_layers)],
          f"out {rel(out, ref):.2e}")
    for a, b in zip(layers, ref_layers):
        assert rel(a, b) < 1e-5
    assert rel(out, ref) < 1e-5
    steps = 3
    ref_loc, _ = O.rollout(om, pos, vel, mass, steps)
    roll = S.SelfFeedRollout(m, B, N, "cuda", max_frames=steps + 1, use_cuda_graph=True)
    roll.reset(pos, vel, mass)
    tp, _ = roll.run(steps)
    assert rel(tp.reshape(steps + 1, B, N, 3).permute(1, 0, 2, 3), ref_loc) < 5e-5


def test_generic_rollout_lmax2_matches_oracle():
    torch.manual_seed(9)
    om = O.SEGNN(hidden_features=32, num_layers=2, lmax_h=2).eval()
    O.perturb_bn_buffers(om, seed=2)
    m = S.SEGNN(hidden_features=32, num_layers=2, lmax_h=2)
    m.load_state_dict(om.state_dict())
    m = m.float().cuda().eval()
    B, N, steps = 3, 5, 6
    pos, vel, mass = O.synthetic_system(B, N, seed=4)
    ref_loc, ref_vel = O.rollout(om, pos, vel, mass, steps)
    roll = S.SelfFeedRollout(m, B, N, "cuda", max_frames=steps + 1)
    roll.reset(pos, vel, mass)
    tp, tv = roll.run(steps)
    got_loc = tp.reshape(steps + 1, B, N, 3).permute(1, 0, 2, 3)
    assert rel(got_loc, ref_loc) < 5e-5


@pytest.mark.parametrize("B,N,T", [(4, 5, 40), (2, 30, 25)])
def test_event_counter_macros_bit_exact(B, N, T):
    """Collision / sticking / leaving / sharp-turn counters (visualization_utils.py:1093-1222) are integer statistics:
    bit-exact against the reference's loop structure on float32-representable trajectories built to exercise every
    branch (contacts of different run lengths, bodies that leave and return, reversals, a zero velocity)."""
    import numpy as np
    gen = torch.Generator().manual_seed(B + N + T)
    vel = torch.randn(B, T, N, 3, generator=gen) * 0.4
    vel[:, ::7] = -vel[:, ::7]                       # sharp turns
    vel[0, 3, 0] = 0.0                               # zero velocity -> NaN angle, counts as no turn
    loc = torch.cumsum(vel * 0.3, dim=1) + torch.randn(B, 1, N, 3, generator=gen) * 0.6
    loc[:, T // 2:, 1] += 40.0                       # body 1 leaves for the second half (run > 10 when T/2 > 10)
    loc[:, 5:9, 2] = loc[:, 5:9, 3] + 0.01           # a 4-step contact: collision upgraded to sticking
    loc[:, 12:14, 2] = loc[:, 12:14, 3] + 0.01       # a 2-step contact: plain collision
    ref = O.event_counters(loc.double().numpy(), vel.double().numpy())
    tp = loc.permute(1, 0, 2, 3).reshape(T, B * N, 3).contiguous().cuda()
    tv = vel.permute(1, 0, 2, 3).reshape(T, B * N, 3).contiguous().cuda()
    got = S.macros.event_counters(tp, tv, B, N)
    for k in ("stickings", "collisions", "bodies_left", "sharp_turns"):
        assert np.array_equal(got[k], ref[k].astype(np.int64)), (k, got[k], ref[k])
    assert ref["stickings"].sum() > 0 and ref["collisions"].sum() > 0 and ref["sharp_turns"].sum() > 0
    assert np.abs(got["max_com_distance"] - ref["max_com_distance"]).max() < 1e-5 * max(ref["max_com_distance"].max(), 1)


# ---- device ground-truth simulator ---------------------------------------------------------------------------------
@pytest.mark.parametrize("B,N,T,freq", [(3, 5, 300, 10), (2, 40, 100, 5)])
def test_gravity_simulator_matches_reference_integrator(B, N, T, freq):
    """GravitySim (synthetic_sim.py:305-420) on the device in float64 vs the NumPy restatement: same leapfrog, same
    softened force, frame k = state after k * sample_freq steps; agreement to 1e-9 over hundreds of steps."""
    import numpy as np
    sim = S.simulator.GravitySim(n_balls=N, interaction_strength=2.0, dt=0.01, softening=0.2)
    pos, vel, mass = sim.initial_conditions(B, seed=3, device="cuda")
    loc, v, f, m = sim.sample_trajectories(T=T, sample_freq=freq, initial_state=(pos, vel, mass))
    assert loc.shape == (B, T // freq, N, 3) and loc.dtype == torch.float64
    for b in range(B):
        rl, rv, rf = O.gravity_trajectory(pos[b].cpu().numpy(), vel[b].cpu().numpy(), mass[b].cpu().numpy(), 2.0, 0.2,
                                          0.01, T, freq)
        assert np.abs(loc[b].cpu().numpy() - rl).max() < 1e-9 * max(np.abs(rl).max(), 1.0)
        assert np.abs(v[b].cpu().numpy() - rv).max() < 1e-9 * max(np.abs(rv).max(), 1.0)
        assert np.abs(f[b].cpu().numpy() - rf).max() < 1e-9 * max(np.abs(rf).max(), 1.0)
    assert torch.equal(loc[:, 0], pos) and torch.equal(v[:, 0], vel)
    # the simulator feeds the macros: total energy of the ground truth is conserved by the symplectic integrator
    tp = loc.permute(1, 0, 2, 3).reshape(T // freq, B * N, 3).float().contiguous()
    tv = v.permute(1, 0, 2, 3).reshape(T // freq, B * N, 3).float().contiguous()
    e = S.macros.nbody_energies(tp, tv, B, N, 2.0, 0.2)["total"]
    assert np.abs(e - e[0]).max() < 2e-2 * abs(e[0])


# ---- dataloader (segnn_nbody) on the device ------------------------------------------------------------------------
def test_device_dataloader_contract_and_training_loop(tmp_path):
    from types import SimpleNamespace
    args = SimpleNamespace(batch_size=16, num_atoms=5, num_neighbors=4, lmax_attr=1, dataset_name="nbody_small",
                           target="pos_dt+vel", sample_freq=10, precision_mode="single", center_of_mass=False,
                           sim_length=300, seed=1)
    dl = S.SegnnNBodyDataLoader(args)
    assert dl.get_num_nodes() == 5 and dl.dataset.num_steps == 30
    seen = set()
    loc_all, vel_all = dl.dataset.data[0], dl.dataset.data[1]
    for _ in range(29):  # every frame index is used exactly once before new trajectories are generated
        (batch,), _ = dl.get_batch()
        f0 = int((loc_all[0, :, 0] == batch.pos[0]).all(dim=1).nonzero()[0])
        assert f0 not in seen
        seen.add(f0)
        want = torch.cat([loc_all[:, f0 + 1] - loc_all[:, f0], vel_all[:, f0 + 1]], dim=2).reshape(-1, 6)
        assert torch.equal(batch.y, want) and batch.pos.shape == (80, 3) and batch.mass.shape == (80, 1)
    assert seen == set(range(29))
    g = dl.preprocess_batch(batch, "cuda")
    assert g.x.shape == (80, 7) and g.node_attr.shape == (80, 4) and g.edge_index.shape == (2, 16 * 20)
    args.num_neighbors = 5
    with pytest.raises(ValueError):
        dl.preprocess_batch(batch, "cuda")
    args.num_neighbors = 4
    # a short training run driven by the dataloader: the loss must go down
    torch.manual_seed(0)
    model = S.SEGNN(hidden_features=64, num_layers=2).cuda().train()
    step = S.TrainStep(model, 16, 5, learning_rate=1.0, learning_rate_factor=2000.0, clip_gradients_norm=10.0)
    losses = []
    for _ in range(60):
        (batch,), _ = dl.get_batch()
        losses.append(float(step.step(batch.pos, batch.vel, batch.mass, batch.y)))
    assert sum(losses[-10:]) < 0.7 * sum(losses[:10]), losses
    # rollout against simulator ground truth through the reference's run_inference signature
    model.eval()
    d, loc, vel = S.run_inference("segnn", dl, model=model, save_dir=str(tmp_path), print_step=False)
    assert loc.shape == (2, 16, 30, 5, 3) and np.isfinite(loc).all()
    # num_neighbors < N - 1: preprocess_batch materialises the kNN edge list, the model runs on it (generic kernels),
    # run_inference rebuilds the neighbourhood every step; the first predicted frame equals a forward on the batch
    args.num_neighbors = 2
    (batch,), _ = dl.get_batch()
    g = dl.preprocess_batch(batch, "cuda")
    assert g.edge_index.shape == (2, 16 * 5 * 2) and g.edge_attr.shape == (160, 4) and g.node_attr.shape == (80, 4)
    with torch.no_grad():
        pred = model(g)
    assert pred.shape == (80, 6) and bool(torch.isfinite(pred).all())
    d, loc2, vel2 = S.run_inference("segnn", dl, model=model, save_dir=str(tmp_path / "knn"), print_step=False,
                                    num_neighbors=2, max_rollout_steps=6)
    assert loc2.shape == (2, 16, 6, 5, 3) and np.isfinite(loc2).all()
    assert np.abs(loc2[1, :, 1] - loc[1, :, 1]).max() > 1e-6  # not the complete-graph rollout


import numpy as np  # noqa: E402


def test_large_graph_tensor_core_mode_agrees_with_fp32_mode():
    """BASELINE config 4 shape (one N = 1000 graph, hidden 128): too large for the float64 oracle (1 M edges x 514
    message-input columns), so the two independent CUDA paths check each other: tcgen05/bf16 vs FFMA/fp32, plus
    permutation equivariance of the tensor-core path. Tolerance 6e-2 here, not the 2e-2 of the N <= 100 configurations:
    the aggregate of N - 1 messages has a coherent part that grows like N and is removed again by the mean-subtracting
    BatchNorm, while the signal grows like sqrt(N), so bf16 operand rounding (which is coherent across edges: the same
    rounded weights and projections enter every message) is amplified by ~sqrt(N): measured 1e-2 at N = 100 and 4.6e-2
    at N = 1000. Configuration 4 is a training configuration and trains with the fp32 kernels. The BatchNorm running
    statistics are first
    calibrated on this input with train-mode forwards: with random statistics an untrained network sums 999 messages
    per node unnormalised, its features grow to ~600 and the comparison would measure that ill-conditioning (the mode
    difference grows smoothly from 1e-2 at N = 100 to 0.3 at N = 1000 in that case), not the kernels."""
    if not S.ops.tc_available():
        pytest.skip("tensor-core mode not built")
    torch.manual_seed(11)
    m = S.SEGNN(hidden_features=128, num_layers=2).cuda()
    B, N = 1, 1000
    pos, vel, mass = O.synthetic_system(B, N, seed=13)
    p, v, ms = pos.reshape(-1, 3).float().cuda(), vel.reshape(-1, 3).float().cuda(), mass.reshape(-1).float().cuda()
    with torch.no_grad():
        m.train()
        for _ in range(40):  # momentum 0.1: the running statistics converge to this input's
            m.forward_state(p, v, ms, B, N)
        m.eval()
        m.compute_mode = "fp32"
        ref = m.forward_state(p, v, ms, B, N)
        m.compute_mode = "fp16"  # same kernels, fp16 operands: 8x smaller rounding keeps N = 1000 inside the 2e-2 budget
        out16 = m.forward_state(p, v, ms, B, N)
        err16 = float((out16 - ref).abs().max() / ref.abs().max())
        m.compute_mode = "fp16p"  # packed-half producers on fp16 projections
        out16p = m.forward_state(p, v, ms, B, N)
        err16p = float((out16p - ref).abs().max() / ref.abs().max())
        m.compute_mode = "bf16"
        out = m.forward_state(p, v, ms, B, N)
        err = float((out - ref).abs().max() / ref.abs().max())
        print(f"N=1000 vs fp32 mode: bf16 {err:.2e}, fp16 {err16:.2e} (output max {float(ref.abs().max()):.2f})")
        print(f"N=1000 fp16p vs fp32 mode: {err16p:.2e}")
        assert err < 6e-2 and err16 < 2e-2 and err16p < 2e-2
        perm = torch.randperm(N).cuda()
        out_p = m.forward_state(p[perm].contiguous(), v[perm].contiguous(), ms[perm].contiguous(), B, N)
        assert float((out_p - out[perm]).abs().max() / out.abs().max()) < 2e-2  # bf16 operands, different tile grouping


@pytest.mark.parametrize("n", [8, 32, 96])
def test_c_abi_weight_packing_matches_python_packing(n):
    """segnn_pack_weights / segnn_fold_batchnorm (what a non-Python consumer of include/segnn_b200.h calls) against
    packing.py, the differentiable re-layout the training path uses: same blocks, bit for bit."""
    P, ops = S.packing, S.ops
    gen = torch.Generator().manual_seed(n)
    sizes = {ops.PACK_MSG1: (2, 2 * n, 2), ops.PACK_MSG2: (1, 2 * n, 0), ops.PACK_UPDATE1: (2, 2 * n, 0),
             ops.PACK_UPDATE2: (1, n, 0), ops.PACK_POOL1: (1, 2 * n, 0)}
    for kind, (nb, n0, extra) in sizes.items():
        numel = nb * (2 * n * n0 + 2 * n * n) + extra * (n0 + n)
        w = torch.randn(numel, generator=gen).cuda()
        b = torch.randn(n0, generator=gen).cuda()
        got = ops.pack_weights(kind, n, w, b)
        ref = {ops.PACK_MSG1: lambda: P.pack_msg1(w, b, n), ops.PACK_MSG2: lambda: P.pack_msg2(w, b, n),
               ops.PACK_UPDATE1: lambda: P.pack_node_tp(w, b, 2, n, 2 * n),
               ops.PACK_UPDATE2: lambda: P.pack_node_tp(w, b, 1, n, n),
               ops.PACK_POOL1: lambda: P.pack_node_tp(w, b, 1, n, 2 * n)}[kind]()
        assert set(got) == set(ref)
        for k in ref:
            assert got[k].shape == ref[k].shape and torch.equal(got[k], ref[k]), (kind, k)
    w, b = torch.randn(6 * n, generator=gen).cuda(), torch.randn(n, generator=gen).cuda()
    got, ref = ops.pack_weights(ops.PACK_EMBED, n, w, b), P.pack_embedding(w, b, n)
    assert torch.equal(got["w"], ref["w"]) and torch.equal(got["bias"], ref["bias"])
    w = torch.randn(4 * n, generator=gen).cuda()
    assert torch.equal(ops.pack_weights(ops.PACK_HEAD, n, w, None), P.pack_head(w, n))
    bn = [torch.rand(2 * n, generator=gen).cuda() + 0.5, torch.randn(n, generator=gen).cuda(),
          torch.randn(n, generator=gen).cuda(), torch.rand(2 * n, generator=gen).cuda() + 0.5]
    mul, add = ops.fold_batchnorm(*bn, n, 1e-5, 7.0)
    rmul, radd = P.fold_batchnorm(*bn, n, 1e-5, 7.0)
    assert torch.allclose(mul, rmul, rtol=2e-6) and torch.allclose(add, radd, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("nodes,n_in,two_inputs,n_out", [(18, 96, False, 576), (2050, 96, True, 288), (1000, 32, True, 192),
                                                         (262, 64, False, 384), (4098, 96, False, 192),
                                                         (3000, 48, True, 96), (700, 80, True, 160)])
def test_node_gemm_fp16_rows_bit_identical_to_fp32_rows(nodes, n_in, two_inputs, n_out):
    """segnn_node_gemm_tc_x16 (fp16 feature copies, no conversion on load, loads of two batches in flight across tile
    boundaries) against the fp32-input entry points on the same values: the rounding moved from the GEMM loader to the
    producer of the copy, so both output modes are bit-identical -- ragged last tiles, one and two inputs, every K the
    launcher accepts in one and two batches per tile."""
    if not S.ops.tc_available():
        pytest.skip("tensor-core kernels not built")
    ops = S.ops
    gen = torch.Generator(device="cpu").manual_seed(nodes + n_out)
    K = n_in * (2 if two_inputs else 1)
    x0 = torch.randn(nodes, 4, n_in, generator=gen).cuda()
    x1 = torch.randn(nodes, 4, n_in, generator=gen).cuda() if two_inputs else None
    w = dict(w_s=(torch.randn(K, n_out, generator=gen) / K ** 0.5).cuda(),
             w_v=(torch.randn(K, n_out, generator=gen) / K ** 0.5).cuda(), operand=1)
    w["wt_s"], w["wt_v"] = ops.pack_node_weight_tc(w["w_s"], 1), ops.pack_node_weight_tc(w["w_v"], 1)
    h0, h1 = x0.half(), (x1.half() if two_inputs else None)
    ref = ops.node_gemm_out16(x0, x1, w, n_out)
    got = ops.node_gemm_out16(h0, h1, w, n_out)
    assert got.dtype == torch.float16 and torch.equal(got, ref)
    if not two_inputs and n_out % 64 == 0:
        split = n_out // 2
        bias = torch.randn(split, generator=gen).cuda()
        r0, r1 = ops.node_gemm_pair16(x0, w, n_out, bias, split, split)
        g0, g1 = ops.node_gemm_pair16(h0, w, n_out, bias, split, split)
        assert torch.equal(g0, r0) and torch.equal(g1, r1)


@pytest.mark.parametrize("H,L,B,N", [(192, 2, 24, 100), (64, 2, 350, 6), (128, 1, 3, 700)])
def test_fp16_feature_copies_leave_the_packed_half_forward_bit_identical(H, L, B, N):
    """compute_mode 'fp16p' with the fp16 operand copies of the node features (embed / combine / edge kernel write them,
    the tensor-core node GEMMs read them: SEGNNLayer.run_x16) against the same mode converting fp32 features on load:
    every layer output and the prediction are bit-identical, so the oracle parity of the mode carries over."""
    if not S.ops.tc_available():
        pytest.skip("tensor-core kernels not built")
    om, m = make_pair(H, L, seed=5)
    m.compute_mode = "fp16p"
    pos, vel, mass = O.synthetic_system(B, N, seed=7)
    g = gpu_graph(pos, vel, mass, B, N)
    old = S.ops.X16_FEATURES
    try:
        with torch.no_grad():
            S.ops.X16_FEATURES = False
            ref, ref_layers = m(g, return_layers=True)
            S.ops.X16_FEATURES = True
            out, layers = m(g, return_layers=True)
    finally:
        S.ops.X16_FEATURES = old
    assert B * N >= 2048, "the fp16-copy path is taken from 2048 nodes on"
    for a, b in zip(layers, ref_layers):
        assert torch.equal(a, b)
    assert torch.equal(out, ref)
