"""GPU parity against vectors produced by executing the reference's own code (tests/golden/ref_*.pt; generator
tests/golden/make_reference_golden.py, provider recorded in each fixture's 'kind').  Everything goes through the C ABI.
Tolerances: bit-exact for integer work (edge order, event / group-collision counters); per layer 1e-5 (fp32 mode),
2e-2 (bf16), 2.5e-3 (fp16 modes) max-norm relative, as stated in BASELINE.json north_star check (b)."""
import os

import numpy as np
import pytest
import torch

import segnn_b200 as S
from golden.golden_weights import golden_state

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
TOL = {"fp32": 1e-5, "bf16": 2e-2, "fp16": 2.5e-3, "fp16p": 2.5e-3, "generic": 1e-5}


def load(name):
    return torch.load(os.path.join(G, name), weights_only=False)


def rel(a, b):
    return float((a.double().cpu() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def cuda_model(fx, train=False):
    c = fx["config"]
    m = S.SEGNN(hidden_features=c["hidden_features"], lmax_h=c["lmax_h"], lmax_attr=c.get("lmax_attr", 1),
                num_layers=c["num_layers"], norm=c.get("norm", "batch"))
    m.load_state_dict(golden_state(fx["shapes"], fx["ranges"], fx["weight_seed"]))
    return m.float().cuda().train(train)


def graph(fx):
    c = fx["config"]
    return S.GraphBatch(pos=fx["pos"].float().cuda(), vel=fx["vel"].float().cuda(), mass=fx["mass"].float().cuda(),
                        num_graphs=c["B"], n_nodes=c["N"])


def modes_for(m, N):
    if not m.fused:  # lmax_h = 2, lmax_attr = 2, norm = "instance"
        return ["generic"]
    tc = S.ops.tc_available() and m.n in S.ops.TC_MULTIPLICITIES
    return ["fp32"] + (["bf16", "fp16"] if tc else []) + (["fp16p"] if tc and N % 2 == 0 else [])


def test_edge_enumeration_matches_reference_bit_exact():
    fx = load("ref_graph.pt")
    for (B, N), ref in fx["full"].items():
        got = S.build_graph_with_knn(None, B, N, "cuda", N - 1).cpu()
        assert got.dtype == torch.int64 and torch.equal(got, ref), (B, N)
    for case in fx["knn"]:
        got = S.build_graph_with_knn(case["loc"].cuda(), case["B"], case["N"], "cuda", case["k"]).cpu()
        assert torch.equal(got, case["edge_index"]), (case["B"], case["N"], case["k"])


@pytest.mark.parametrize("case", ["h64_n5", "h192_n8", "h128_n12", "h32_l2_n6", "h32_a2_n6", "h32_l2_a2_n5",
                                  "h32_inorm_n6", "h32_nonorm_n5"])
def test_segnn_eval_per_layer_matches_reference(case):
    fx = load(f"ref_segnn_{case}.pt")
    c = fx["config"]
    m = cuda_model(fx)
    assert str(m.hidden_irreps).replace(" ", "") == fx["hidden_irreps"]
    assert sum(p.numel() for p in m.parameters()) == fx["num_params"]
    ref = fx["eval"]
    g = S.O3Transform(c.get("lmax_attr", 1))(graph(fx))
    tr = ref["transform"]
    assert torch.equal(g.edge_index.cpu(), tr["edge_index"])
    assert rel(g.x, tr["x"]) < 2e-6 and rel(g.edge_attr, tr["edge_attr"]) < 2e-6
    assert rel(g.additional_message_features, tr["additional_message_features"]) < 2e-6
    assert float((g.node_attr.double().cpu()[:, 1:] - tr["node_attr"][:, 1:]).abs().max()) < 2e-6
    with torch.no_grad():
        for mode in modes_for(m, c["N"]):
            m.compute_mode = mode
            out, layers = m(graph(fx), return_layers=True)
            errs = [rel(a, b) for a, b in zip(layers, ref["layers"])]
            print(f"[{case} {mode}] per-layer rel err vs reference", [f"{e:.2e}" for e in errs], f"out {rel(out, ref['out']):.2e}")
            assert max(errs) < TOL[mode] and rel(out, ref["out"]) < TOL[mode], (mode, errs)


@pytest.mark.parametrize("case", ["h32_knn3_n8", "h32_l2_knn2_n6", "h32_l2_inorm_knn3_n7"])
def test_segnn_on_knn_graph_matches_reference(case):
    """num_neighbors < N - 1 (utils/build_fully_connected_graph.py:42-80): the kNN edge list bit-exact, O3Transform on it,
    per-layer outputs of the generic kernels with gathers through edge_index against the reference run."""
    fx = load(f"ref_segnn_{case}.pt")
    c = fx["config"]
    m = cuda_model(fx)
    ref, tr = fx["eval"], fx["eval"]["transform"]
    g = graph(fx)
    g.edge_index = S.build_graph_with_knn(g.pos, c["B"], c["N"], "cuda", c["num_neighbors"])
    assert torch.equal(g.edge_index.cpu(), tr["edge_index"])
    g = S.O3Transform(1)(g)
    assert rel(g.x, tr["x"]) < 2e-6 and rel(g.edge_attr, tr["edge_attr"]) < 2e-6
    assert rel(g.additional_message_features, tr["additional_message_features"]) < 2e-6
    assert float((g.node_attr.double().cpu()[:, 1:] - tr["node_attr"][:, 1:]).abs().max()) < 2e-6
    with torch.no_grad():
        out, layers = m(g, return_layers=True)
    errs = [rel(a, b) for a, b in zip(layers, ref["layers"])]
    print(f"[{case}] per-layer rel err vs reference", [f"{e:.2e}" for e in errs], f"out {rel(out, ref['out']):.2e}")
    assert max(errs) < 1e-5 and rel(out, ref["out"]) < 1e-5
    with pytest.raises(NotImplementedError):  # inference only
        m.train()(g)


@pytest.mark.parametrize("case", ["h64_n5", "h192_n8", "h128_n12"])
def test_segnn_training_step_matches_reference(case):
    """Train-mode BatchNorm forward, loss, hand-written backward and running statistics against the reference's autograd."""
    fx = load(f"ref_segnn_{case}.pt")
    ref = fx["train"]
    with torch.no_grad():  # per-layer outputs come from the non-differentiable entry point, on their own model copy
        _, layers = cuda_model(fx, train=True)(graph(fx), return_layers=True)
    for i, (a, b) in enumerate(zip(layers, ref["layers"])):
        assert rel(a, b) < 1e-5, (case, i)
    m = cuda_model(fx, train=True)
    out = m(graph(fx))
    loss = S.target_common_loss(out, fx["y"].float().cuda())
    assert abs(float(loss) - ref["loss"]) < 1e-5 * abs(ref["loss"])
    loss.backward()
    top = max(ref["grad_norms"].values())
    worst = 0.0
    for k, p in m.named_parameters():
        stride, vals = ref["grads"][k]
        got = p.grad.double().cpu().reshape(-1)[::stride]
        scale = float(vals.abs().max())
        err = float((got - vals).abs().max())
        worst = max(worst, err / max(scale, 1e-6 * top))
        assert err <= 1e-4 * scale + 1e-5 * top, (case, k, err, scale)
    print(f"[{case}] worst gradient rel err vs reference {worst:.2e}")
    sd = m.state_dict()
    for k, v in ref["running"].items():
        assert rel(sd[k], v) < 1e-5, k


@pytest.mark.parametrize("use_graph", [False, True])
def test_rollout_matches_reference_run_inference(use_graph, tmp_path):
    """The reference's run_inference on its own simulator ground truth vs SelfFeedRollout / S.run_inference."""
    ro = load("ref_sim_macros.pt")["rollout"]
    fx = load("ref_segnn_h64_n5.pt")
    loc, vel = ro["combined_locations"], ro["combined_velocities"]
    B, T, N = loc.shape[1:4]
    m = cuda_model(fx)
    mass = torch.ones(B, N, 1)
    roll = S.SelfFeedRollout(m, B, N, "cuda", max_frames=T, use_cuda_graph=use_graph)
    roll.reset(loc[0, :, 0], vel[0, :, 0], mass)
    tp, tv = roll.run(T - 1)
    got_loc = tp.reshape(T, B, N, 3).permute(1, 0, 2, 3)
    got_vel = tv.reshape(T, B, N, 3).permute(1, 0, 2, 3)
    assert rel(got_loc, loc[1]) < 5e-5 and rel(got_vel, vel[1]) < 5e-5
    d, cl, cv = S.run_inference("segnn", None, model=m, save_dir=str(tmp_path), print_step=False,
                                ground_truth=(loc[0], vel[0], mass))
    assert sorted(os.listdir(d)) == ro["files"]
    assert np.abs(cl[1] - loc[1].numpy()).max() < 5e-5 * float(loc[1].abs().max())
    assert np.array_equal(cl[0], loc[0].numpy())


@pytest.mark.parametrize("key", ["gravity", "gravity7"])
def test_gravity_simulator_matches_reference(key):
    fx = load("ref_sim_macros.pt")[key]
    p = fx["params"]
    N = fx["loc"].shape[1]
    sim = S.simulator.GravitySim(n_balls=N, interaction_strength=p["G"], dt=p["dt"], softening=p["softening"])
    init = (fx["loc"][:1], fx["vel"][:1], fx["mass"].reshape(1, N, 1))
    loc, vel, force, _ = sim.sample_trajectories(T=p["T"], sample_freq=p["sample_freq"], initial_state=init)
    for got, ref in ((loc, fx["loc"]), (vel, fx["vel"]), (force, fx["force"])):
        assert float((got[0].cpu() - ref).abs().max()) < 1e-9 * max(1.0, float(ref.abs().max()))


def test_charged_simulator_matches_reference():
    fx = load("ref_sim_macros.pt")["charged"]
    p = fx["params"]
    N = fx["x0"].shape[0]
    sim = S.simulator.ChargedSim(n_balls=N, interaction_strength=p["interaction_strength"], delta_t=p["delta_t"])
    assert abs(sim.max_force - p["max_F"]) < 1e-12
    X, V = sim.simulate(fx["x0"].reshape(1, N, 3), fx["v0"].reshape(1, N, 3), fx["charges"].reshape(1, N, 1),
                        steps=fx["X"].shape[0])
    assert float((X[0].cpu() - fx["X"]).abs().max()) < 1e-9 and float((V[0].cpu() - fx["V"]).abs().max()) < 1e-9


def test_macro_kernels_match_reference():
    mc = load("ref_sim_macros.pt")["macros"]
    loc, vel = mc["loc"], mc["vel"]
    S_, T, N = loc.shape[:3]
    tp = loc.permute(1, 0, 2, 3).reshape(T, S_ * N, 3).float().contiguous().cuda()
    tv = vel.permute(1, 0, 2, 3).reshape(T, S_ * N, 3).float().contiguous().cuda()
    # counters are integer statistics of float32 trajectories: compare with the reference run on the float64 ones; the
    # fixture keeps every distance / angle away from its threshold by more than float32 resolution (checked below)
    d = (loc[:, :, :, None] - loc[:, :, None]).norm(dim=-1)
    for thr in (0.5, 1.0, 2.0, 1.5):
        assert float((d - thr).abs().min()) > 1e-5
    ev = S.macros.event_counters(tp, tv, S_, N)
    assert np.array_equal(ev["stickings"], mc["stickings"].numpy().astype(np.int64))
    assert np.array_equal(ev["collisions"], mc["collisions"].numpy().astype(np.int64))
    assert np.array_equal(ev["bodies_left"], mc["bodies_left_d15"].numpy().astype(np.int64))
    assert np.array_equal(ev["sharp_turns"], mc["sharp_turns_30"].numpy().astype(np.int64))
    assert np.abs(ev["max_com_distance"] - mc["max_com_distance"].numpy()).max() < 1e-5
    ev2 = S.macros.event_counters(tp, tv, S_, N, time_threshold=2, contact_distance=1.0, leave_distance=1.4,
                                  turn_angle_degrees=90.0)
    assert np.array_equal(ev2["stickings"], mc["stickings_t2_d1"].numpy().astype(np.int64))
    assert np.array_equal(ev2["collisions"], mc["collisions_t2_d1"].numpy().astype(np.int64))
    assert np.array_equal(ev2["bodies_left"], mc["bodies_left_d2"].numpy().astype(np.int64))
    assert np.array_equal(ev2["sharp_turns"], mc["sharp_turns_90"].numpy().astype(np.int64))
    for tt, dd in ((2, 2.0), (3, 1.5)):
        got = S.macros.group_collisions(tp, S_, N, time_threshold=tt, distance_threshold=dd)
        key = f"group_collisions_t{tt}_d{2 if dd == 2.0 else dd}"
        assert np.array_equal(got, mc[key].numpy().astype(np.int64)), (got, mc[key])
    en = S.macros.nbody_energies(tp, tv, S_, N, 2.0, 0.2)
    for k in ("potential", "kinetic", "total"):
        assert np.abs(en[k] - mc["energies"][k].numpy()).max() < 1e-5 * np.abs(mc["energies"][k].numpy()).max()
    assert np.abs(S.macros.momentum_statistics(tp, tv, S_, N) - mc["momentum_mean_over_time"].numpy()).max() < 1e-5


def test_ks_fisher_noam_match_reference():
    fx = load("ref_sim_macros.pt")
    ks = fx["ks"]
    a, b, a_nan = ks["a"].numpy(), ks["b"].numpy(), ks["a_nan"].numpy()
    got = [S.macros.ks_p(a, b), S.macros.ks_p(a, a), S.macros.ks_p(a_nan, b), S.macros.ks_p(np.array([]), b)]
    for g, r in zip(got, ks["p"]):
        assert (g != g and r != r) or abs(g - r) <= 1e-12 * abs(r)
    assert abs(S.macros.combine_pvalues_fisher(ks["fisher_inputs"]) - ks["fisher"]) <= 1e-9 * ks["fisher"]
    assert S.macros.combine_pvalues_fisher([1e-200, 1e-250, 1e-100]) == ks["fisher_tiny"]
    nm = fx["noam"]
    for s, r in zip((0, 1, 10, 2999, 3000, 3001, 100000), nm["rates"]):
        assert abs(S.noam_rate(s, nm["hidden"], nm["factor"], nm["warmup"]) - r) <= 1e-14 * r


def test_module_level_tensor_products_match_reference():
    """O3TensorProduct[SwishGate] as free-standing modules on arbitrary irreps (o3_building_blocks.py:10-203) and
    InstanceNorm (instance_norm.py), with the reference's parameters."""
    fx = load("ref_tensor_products.pt")
    for case in fx["cases"]:
        cls = getattr(S, case["cls"])
        mod = cls(case["in1"], case["out"], case["in2"])
        mod.load_state_dict(case["state"])
        mod = mod.float().cuda()
        x2 = None if case["x2"] is None else case["x2"].float().cuda()
        with torch.no_grad():
            got = mod(case["x1"].float().cuda(), x2)
        assert rel(got, case["y"]) < 1e-5, (case["cls"], case["in1"], case["out"])
    inorm = fx["instance_norm"]
    mod = S.InstanceNorm(inorm["irreps"])
    with torch.no_grad():
        mod.weight.copy_(inorm["weight"])
        mod.bias.copy_(inorm["bias"])
    mod = mod.float().cuda()
    assert rel(mod(inorm["x"].float().cuda(), inorm["batch"].cuda()), inorm["y"]) < 1e-5
