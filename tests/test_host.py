"""CPU tests of the host side: C-ABI exports, packing + kernel algebra (emulated) against the oracle, irreps
sizing, checkpoint key compatibility, error behaviour without a GPU, and 2-rank gloo sharding."""
import ctypes
import os
import re

import pytest
import torch

import emulate as E
import segnn_b200 as S
from oracle import segnn_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "segnn_b200.h")).read()
    declared = set(re.findall(r"\b(segnn_[a-z0-9_]+)\s*\(", header))
    declared -= {"segnn_stream_t"}
    assert len(declared) >= 12
    lib = ctypes.CDLL(S._lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/segnn_b200.h but not exported"
    assert declared == set(S._lib.PROTOTYPES.keys())
    assert lib.segnn_version() >= 100


def test_cabi_argument_errors_without_gpu():
    lib = S._lib.lib
    assert lib.segnn_prep_fwd(None, None, 2, 5, None, None, None) == -1
    assert b"null" in lib.segnn_last_error()
    assert lib.segnn_edge_index(0, 5, None, None) == 0  # empty batch is a no-op
    assert lib.segnn_edge_layer_fwd(7, None, None, 0, 5, 8, None, None, None, None, None, None, None, None, None,
                                    None, None, None, None, None) == 0


def test_untaken_branch_exports_argument_errors_without_gpu():
    """lmax_attr <= 2 geometry and the edge-list (kNN) entry points: argument checks run before any launch."""
    lib = S._lib.lib
    assert lib.segnn_prep_fwd_lmax(None, None, 2, 5, 3, None, None, None) == -1 and b"lmax_attr" in lib.segnn_last_error()
    assert lib.segnn_edge_attr_lmax(None, None, 2, 5, 2, None, None, None) == -1 and b"null" in lib.segnn_last_error()
    assert lib.segnn_edge_attr_lmax(None, None, 0, 5, 2, None, None, None) == 0
    assert lib.segnn_edge_attr_list(None, None, None, 0, 1, None, None, None) == 0  # empty edge list is a no-op
    assert lib.segnn_edge_attr_list(None, None, None, 4, 1, None, None, None) == -1
    assert lib.segnn_generic_message_input_list(None, None, None, 4, 0, 2, None, None) == -1
    assert lib.segnn_segment_reduce(None, None, None, 0, 8, 0, None, None) == 0
    assert lib.segnn_segment_reduce(None, None, None, 3, 8, 0, None, None) == -1
    assert lib.segnn_prep_fwd_list(None, None, None, 3, 1, None, None, None) == -1


def test_edge_list_csr_groups_edges_by_target_in_edge_order():
    """Index plumbing of segnn_segment_reduce: stable sort of the targets + searchsorted; isolated targets get an
    empty segment; the order inside a segment is the edge order (what makes the sum reproducible)."""
    ei = torch.tensor([[0, 1, 2, 3, 4, 0, 2], [3, 3, 0, 1, 3, 1, 0]])
    order, ptr = S.ops.edge_list_csr(ei, 6)
    assert ptr.tolist() == [0, 2, 4, 4, 7, 7, 7]  # nodes 2, 4, 5 receive nothing
    assert order.tolist() == [2, 6, 3, 5, 0, 1, 4]
    o2 = O.knn_edge_index(torch.randn(2 * 7, 3, dtype=torch.float64), 2, 7, 3)
    order, ptr = S.ops.edge_list_csr(o2, 14)
    assert int(ptr[-1]) == o2.shape[1] and torch.equal(o2[1][order], torch.sort(o2[1]).values)
    for node in range(14):
        seg = order[ptr[node]: ptr[node + 1]]
        assert bool((o2[1][seg] == node).all()) and seg.tolist() == sorted(seg.tolist())


def test_configuration_routing_of_the_untaken_branches():
    """lmax_attr = 2, norm = 'instance' and string irreps: which configurations take the fused kernels."""
    from segnn_b200.generic import GenericRunner
    m = S.SEGNN(hidden_features=32, lmax_h=2, lmax_attr=2, num_layers=1)
    om = O.SEGNN(hidden_features=32, lmax_h=2, lmax_attr=2, num_layers=1)
    assert not m.fused and str(m.hidden_irreps).replace(" ", "") == str(om.hidden_irreps).replace(" ", "")
    assert set(m.state_dict()) == set(om.state_dict())
    r = GenericRunner(m, "cpu")  # plans are host-side tables
    assert not r.use_l2_rows and r.layers[0]["msg1h"].n_pairs == 11 and r.layers[0]["msg1h"].n_adds == 3
    assert r.embed.cg.shape[1:] == (5, 5, 5)
    mi = S.SEGNN(hidden_features=32, num_layers=1, norm="instance")
    assert not mi.fused and mi.layers[0].message_norm is None
    assert set(mi.state_dict()) == set(O.SEGNN(hidden_features=32, num_layers=1, norm="instance").state_dict())
    assert S.SEGNN(hidden_features=32, num_layers=1, norm=None).fused
    h = "16x0e+16x1o"
    assert S.SEGNNLayer(h, h, h, "1x0e+1x1o", "1x0e+1x1o", norm="batch", additional_message_irreps="2x0e").fused
    assert not S.SEGNNLayer(h, h, h, "1x0e+1x1o+1x2e", "1x0e+1x1o+1x2e", norm="batch",
                            additional_message_irreps="2x0e").fused
    with pytest.raises(NotImplementedError):
        S.SEGNN(hidden_features=32, lmax_attr=3)
    with pytest.raises(ValueError):
        S.SelfFeedRollout(S.SEGNN(hidden_features=16, num_layers=1).eval(), 1, 5, "cpu", max_frames=2, num_neighbors=5)


def test_no_cpu_fallback():
    model = S.SEGNN(hidden_features=16, num_layers=1).eval()
    g = S.GraphBatch(pos=torch.zeros(4, 3), vel=torch.zeros(4, 3), mass=torch.ones(4, 1), num_graphs=1, n_nodes=4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(g)


def test_irreps_and_sizing():
    assert str(S.weight_balanced_irreps(192, S.Irreps("1x0e+1x1o"), 1)) == "96x0e+96x1o"
    assert str(S.weight_balanced_irreps(64, S.Irreps("1x0e+1x1o"), 1)) == "32x0e+32x1o"
    assert str(S.weight_balanced_irreps(192, S.Irreps("1x0e+1x1o"), 2)) == "73x0e+73x1o+73x2e"
    h = S.Irreps("96x0e+96x1o")
    assert str((2 * h + S.Irreps("2x0e")).simplify()) == "96x0e+96x1o+96x0e+96x1o+2x0e"
    assert (h + h).simplify().dim == 768
    with pytest.raises(NotImplementedError):
        S.SEGNN(hidden_features=64, lmax_h=3)
    m2 = S.SEGNN(hidden_features=192, lmax_h=2, num_layers=6)  # BASELINE config 3: generic-irreps path
    assert str(m2.hidden_irreps) == "73x0e+73x1o+73x2e" and not m2.fused
    assert sum(p.numel() for p in m2.parameters()) == 2053709


def test_state_dict_interchange_with_oracle():
    om = O.SEGNN(hidden_features=64, num_layers=4)
    m = S.SEGNN(hidden_features=64, num_layers=4)
    assert list(m.state_dict().keys()) == list(om.state_dict().keys())
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(om.state_dict()[k].shape), k
    sd = dict(om.state_dict())
    sd["layers.0.message_layer_1.tp.output_mask"] = torch.ones(3)  # e3nn-internal buffer in reference checkpoints
    m.load_state_dict(sd)
    assert m.get_model_size() == 64
    assert m.get_serializable_attributes()["num_params"] == 148256
    assert m.get_serializable_attributes()["hidden_irreps"] == "32x0e+32x1o"


def test_init_matches_reference_distribution():
    torch.manual_seed(0)
    tp = S.O3TensorProductSwishGate("96x0e+96x1o+96x0e+96x1o+2x0e", "96x0e+96x1o", "1x0e+1x1o")
    assert tp.tp.weight.numel() == 111168 and tp.biases.numel() == 192
    bound = 1.0 / (386 ** 0.5)
    assert float(tp.tp.weight.abs().max()) <= bound and float(tp.tp.weight.abs().max()) > 0.95 * bound


@pytest.mark.parametrize("H,L,B,N", [(64, 2, 3, 5), (10, 2, 2, 7)])
def test_packing_and_kernel_algebra_vs_oracle(H, L, B, N):
    torch.manual_seed(0)
    om = O.SEGNN(hidden_features=H, num_layers=L).eval()
    O.perturb_bn_buffers(om)
    m = S.SEGNN(hidden_features=H, num_layers=L).double().eval()
    m.load_state_dict(om.state_dict())
    pos, vel, mass = O.synthetic_system(B, N, seed=3)
    g = O.make_graph(pos.reshape(-1, 3), vel.reshape(-1, 3), mass.reshape(-1, 1), B, N)
    with torch.no_grad():
        ref, ref_layers = om(g, return_layers=True)
        packed = m.packed(N - 1)  # fp32 operands, exactly what the kernels receive
        out, layers = E.model_forward(packed, m.n, pos.reshape(-1, 3), vel.reshape(-1, 3), mass.reshape(-1), B, N,
                                      return_layers=True)
    for a, b in zip(layers, ref_layers):
        assert float((S.packing.from_planar(a) - b).abs().max() / b.abs().max()) < 2e-6
    assert float((out - ref).abs().max() / ref.abs().max()) < 2e-6


def test_planar_layout_roundtrip():
    x = torch.randn(5, 4 * 6)
    assert torch.equal(S.packing.from_planar(S.packing.to_planar(x, 6)), x)


def test_shard_simulations_partition():
    for total, world in [(8192, 8), (10, 4), (3, 8), (0, 2)]:
        spans = [S.shard_simulations(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == total
        for (s0, c0), (s1, _) in zip(spans, spans[1:]):
            assert s0 + c0 == s1
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def _gloo_worker(rank, world, port, total, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, count = S.shard_simulations(total, rank, world)
    owned = torch.zeros(total, dtype=torch.int64)
    owned[start:start + count] = 1
    dist.all_reduce(owned)  # every simulation must be owned by exactly one rank
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        out.put((owned.tolist(), float(t)))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, 11, q)) for r in range(2)]
    for p in procs:
        p.start()
    owned, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert owned == [1] * 11 and tmax == 2.0


@pytest.mark.parametrize("bn_train", [True, False])
def test_training_orchestration_and_batchnorm_algebra_vs_oracle(bn_train):
    """training.forward_train/backward_train driven by the torch stand-in for the kernels (float64, CPU): predictions,
    every parameter gradient and the BatchNorm running statistics must match autograd through the oracle."""
    torch.manual_seed(0)
    H, L, B, N = 16, 2, 3, 5
    om = O.SEGNN(hidden_features=H, num_layers=L)
    O.perturb_bn_buffers(om)
    om.train(bn_train)
    m = S.SEGNN(hidden_features=H, num_layers=L).double()
    m.load_state_dict(om.state_dict())
    m.train(bn_train)
    pos, vel, mass = O.synthetic_system(B, N, seed=3)
    y = torch.randn(B * N, 6, dtype=torch.float64)
    ref = om(O.make_graph(pos.reshape(-1, 3), vel.reshape(-1, 3), mass.reshape(-1, 1), B, N))
    O.target_common_loss(ref, y).backward()
    pred = m._forward_train(pos.reshape(-1, 3), vel.reshape(-1, 3), mass.reshape(-1), B, N, bn_train, True,
                            backend=E.TorchBackend, dtype=torch.float64)
    O.target_common_loss(pred, y).backward()
    assert float((pred.detach() - ref.detach()).abs().max() / ref.detach().abs().max()) < 1e-11
    for (k, a), (k2, b) in zip(om.named_parameters(), m.named_parameters()):
        assert k == k2 and b.grad is not None, k
        scale = float(a.grad.abs().max())
        assert float((a.grad - b.grad).abs().max()) <= 1e-8 * scale + 1e-14, k
    sd = om.state_dict()
    for k, b in m.state_dict().items():
        if "running" in k:
            assert float((sd[k] - b).abs().max()) < 1e-10, k


def test_real_wigner_3j_matches_oracle_and_closed_forms():
    import numpy as np
    from segnn_b200 import cg
    for ls in [(0, 0, 0), (0, 1, 1), (1, 0, 1), (1, 1, 0), (1, 1, 2), (2, 0, 2), (2, 1, 1)]:
        assert np.abs(cg.real_wigner_3j(*ls) - O.wigner_3j(*ls).numpy()).max() < 1e-14
        assert abs(np.linalg.norm(cg.real_wigner_3j(*ls)) - 1.0) < 1e-14
    c110 = cg.real_wigner_3j(1, 1, 0)[:, :, 0]
    assert np.abs(c110 - np.eye(3) / np.sqrt(3)).max() < 1e-14          # (1,1,0) = x.y / sqrt(3)
    c112 = cg.real_wigner_3j(1, 1, 2)
    assert abs(c112[0, 2, 0] - 1 / np.sqrt(10)) < 1e-14 and abs(c112[1, 1, 2] - 2 / np.sqrt(30)) < 1e-14


def test_noam_rate_matches_reference_formula():
    # trainer.py:188-195: factor * size^-0.5 * min(step^-0.5, step * warmup^-1.5); step 0 is treated as step 1
    assert S.noam_rate(0, 192.0, 1.0, 4000) == S.noam_rate(1, 192.0, 1.0, 4000)
    assert abs(S.noam_rate(4000, 192.0, 2.0, 4000) - 2.0 * 192 ** -0.5 * 4000 ** -0.5) < 1e-15
    assert abs(S.noam_rate(10, 64.0, 1.0, 4000) - 64 ** -0.5 * 10 * 4000 ** -1.5) < 1e-18
    assert S.noam_rate(8000, 64.0) < S.noam_rate(4000, 64.0)


def _ddp_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(rank)  # ranks start from different weights and buffers
    model = S.SEGNN(hidden_features=16, num_layers=1)
    S.trainer.broadcast_module_state(model, 0)
    ref = [p.detach().clone() for p in model.parameters()]
    for i, p in enumerate(model.parameters()):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1)) if i != 2 else None  # one parameter without a gradient
    S.allreduce_gradients(model.parameters())
    grads = [float(p.grad.reshape(-1)[0]) for p in model.parameters()]
    # flat storage: the gradients are slices of one buffer, averaged in place by one collective
    flat, sink = model.use_flat_storage()
    assert all(torch.equal(a, b) for a, b in zip(ref, model.parameters()))  # values preserved by the re-pointing
    sink.fill_(float(rank + 1))
    S.trainer.allreduce_flat(sink)
    grads.append(float(next(iter(model.parameters())).grad.reshape(-1)[0]))
    gathered = [torch.zeros_like(ref[0]) for _ in range(world)]
    dist.all_gather(gathered, ref[0])
    same = all(torch.equal(g, gathered[0]) for g in gathered)
    if rank == 0:
        out.put((grads, same))
    dist.destroy_process_group()


def test_two_rank_gloo_gradient_allreduce():
    """world_size-2 data-parallel plumbing on CPU: parameters are broadcast from rank 0, gradients are averaged with one
    flat-bucket all-reduce, a parameter without gradient contributes zeros."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    grads, same = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same
    assert grads.pop() == 1.5  # allreduce_flat: mean of (1, 2), seen through a parameter's .grad view
    for i, g in enumerate(grads):
        assert g == (0.0 if i == 2 else 1.5 * (i + 1))


def test_flat_parameter_and_gradient_storage_matches_per_tensor_path():
    """SEGNN.use_flat_storage: same prediction and same gradients as the per-tensor path (emulated kernels, float32
    on CPU), gradients land in the flat buffer through the .grad views, and a dtype cast drops back to the generic path."""
    torch.manual_seed(4)
    H, L, B, N = 16, 2, 3, 5
    a = S.SEGNN(hidden_features=H, num_layers=L).train()
    b = S.SEGNN(hidden_features=H, num_layers=L).train()
    b.load_state_dict(a.state_dict())
    flat, sink = b.use_flat_storage()
    params = list(b.parameters())
    assert b._flat_storage(params)[0] is flat and flat.numel() == sum(p.numel() for p in params)
    pos, vel, mass = O.synthetic_system(B, N, seed=3)
    pos, vel, mass = pos.reshape(-1, 3).float(), vel.reshape(-1, 3).float(), mass.reshape(-1).float()
    y = torch.randn(B * N, 6)
    outs = []
    for m in (a, b):
        pred = m._forward_train(pos, vel, mass, B, N, True, True, backend=E.TorchBackend, dtype=torch.float32)
        O.target_common_loss(pred, y).backward()
        outs.append(pred.detach())
    assert torch.allclose(outs[0], outs[1], atol=1e-6)
    off = 0
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert pb.grad.data_ptr() == sink.data_ptr() + 4 * off  # still the view, not a clone
        assert torch.allclose(pa.grad, pb.grad, atol=1e-5 * float(pa.grad.abs().max()) + 1e-9)
        off += pb.numel()
    with torch.no_grad():  # an optimizer-style in-place update is visible through the flat buffer
        params[0].add_(1.0)
    assert float(flat[0]) == float(params[0].reshape(-1)[0])
    b.double()
    assert b._flat_storage(list(b.parameters())) == (None, None)


def test_fisher_and_ks_match_reference_libraries():
    """macros.combine_pvalues_fisher (closed form, log space) and macros.ks_p against the oracle's restatement of
    utils/ks_utils.py (scipy + mpmath), including the 1e-300 floor and the NaN / non-positive filtering."""
    import numpy as np
    M = S.macros
    rng = np.random.default_rng(0)
    for ps in ([0.5], [0.01, 0.2, 0.9], [1e-5, 1e-7, 0.3, float("nan"), 0.0], list(rng.uniform(1e-6, 1, 40)),
               [1e-120] * 5, [1.0, 1.0]):
        ref, got = O.combine_pvalues_fisher(ps), M.combine_pvalues_fisher(ps)
        assert abs(got - ref) <= 1e-9 * ref + 1e-300, (ps, ref, got)
    assert M.combine_pvalues_fisher([float("nan"), -1.0]) != M.combine_pvalues_fisher([float("nan"), -1.0])  # NaN
    a, b = rng.normal(size=300), rng.normal(0.3, 1.0, size=250)
    assert abs(M.ks_p(a, b) - O.ks_p(a, b)) < 1e-12
    assert abs(M.ks_statistic(a, b) - __import__("scipy").stats.ks_2samp(a, b)[0]) < 1e-12
    assert M.ks_p([], b) != M.ks_p([], b)
    assert M.energy_ratio_steps(np.array([1.0, 1.2, 3.0, 1.0]), np.ones(4)) == 2
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        M.energy_momentum(torch.zeros(2, 10, 3), torch.zeros(2, 10, 3), 2, 5)


# ---- checkpoints / run directory in the reference's on-disk layout ------------------------------
def _ref_run_dir():
    import os
    return os.path.join(os.path.dirname(__file__), "golden", "ref_run", "2025-01-02_03-04-05")


def test_reference_written_checkpoint_loads_and_round_trips(tmp_path):
    """tests/golden/ref_run/... was written by the reference's own Trainer.save_model / save_model_params /
    save_dataset_attributes (tests/golden/make_reference_golden.py::make_checkpoint_fixture)."""
    import json
    import os
    import torch
    import segnn_b200 as S
    from oracle import segnn_oracle as O
    run = _ref_run_dir()
    ck = torch.load(os.path.join(run, "model.pth"), map_location="cpu", weights_only=False)
    assert set(ck) == {"model_state_dict", "optimizer_state_dict", "step_count", "best_metrics", "scheduler_state_dict"}
    model = S.SEGNN(hidden_features=16, num_layers=1).double()
    ts_params = [p for p in model.parameters()]
    opt = torch.optim.AdamW(ts_params, weight_decay=1e-8, lr=1.0, betas=(0.9, 0.98), eps=1e-9)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: S.noam_rate(s, 16, 1.0, 100))
    S.load_checkpoint(os.path.join(run, "model.pth"), "cpu", model=model, optimizer=opt, scheduler=sched)
    assert S.load_checkpoint.last == {"step_count": 2, "best_metrics": {"valid_loss": 0.25}}
    for k, v in model.state_dict().items():
        assert torch.equal(v, ck["model_state_dict"][k]), k
    assert len(opt.state_dict()["state"]) == len(ts_params) and sched.last_epoch == 2
    assert opt.state_dict()["state"][0]["exp_avg"].shape == ts_params[0].shape
    # what we write carries exactly the keys the reference wrote (incl. e3nn's tp.output_mask buffers)
    assert set(S.checkpoint.reference_state_dict(model)) == set(ck["model_state_dict"])
    path = S.save_model(model, opt, sched, step_count=3, best_metrics={"valid_loss": 0.2}, save_path=str(tmp_path))
    ck2 = torch.load(path, map_location="cpu", weights_only=False)
    assert set(ck2) == set(ck) and ck2["step_count"] == 3
    assert ck2["optimizer_state_dict"]["param_groups"][0]["betas"] == ck["optimizer_state_dict"]["param_groups"][0]["betas"]
    om = O.SEGNN(hidden_features=16, num_layers=1)
    om.load_state_dict({k: v for k, v in ck2["model_state_dict"].items() if "output_mask" not in k})
    # model_params.json: same keys and values as the reference wrote for the same architecture
    ref_params = json.load(open(os.path.join(run, "model_params.json")))
    mine = model.get_serializable_attributes()
    assert set(mine) == set(ref_params)
    for k in ("hidden_features", "lmax_h", "lmax_attr", "num_layers", "norm", "task", "num_params"):
        assert mine[k] == ref_params[k], k
    for k in ("node_attr_irreps", "input_irreps", "hidden_irreps", "output_irreps", "edge_attr_irreps",
              "additional_message_irreps"):
        assert mine[k].replace(" ", "") == ref_params[k].replace(" ", ""), k
    assert S.checkpoint.get_dataset_metadata_path(os.path.join(run, "model.pth")) == \
        os.path.join(run, "nbody_small_dataset", "metadata.json")
    with pytest.raises(FileNotFoundError):
        S.checkpoint.get_dataset_metadata_path("/tmp")
