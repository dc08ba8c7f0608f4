"""The oracle against vectors produced by EXECUTING THE REFERENCE'S OWN CODE (tests/golden/ref_*.pt, generated in the
build container by tests/golden/make_reference_golden.py from /root/reference), and against the e3nn constants the
reference vendors (Jd.pt + wigner_D of models/equiformer_v2/architecture/wigner.py).

These run on CPU without /root/reference (the fixtures travel, the reference does not).  The GPU counterparts, which
compare the CUDA path with the same fixtures, are in tests/test_gpu_reference.py."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import segnn_oracle as O
from golden.golden_weights import golden_state

G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return torch.load(os.path.join(G, name), weights_only=False)


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-300))


# ---------------------------------------------------------------------------------------------------------------------
# graph enumeration (utils/build_fully_connected_graph.py)
# ---------------------------------------------------------------------------------------------------------------------
def test_edge_enumeration_matches_reference_bit_exact():
    fx = load("ref_graph.pt")
    for (B, N), ref in fx["full"].items():
        got = O.fully_connected_edge_index(B, N)
        assert got.dtype == ref.dtype == torch.int64 and torch.equal(got, ref), (B, N)
    for case in fx["knn"]:
        got = O.build_graph_with_knn(case["loc"], case["B"], case["N"], None, case["k"])
        assert torch.equal(got, case["edge_index"])
    name, msg = fx["too_many_neighbors"]
    with pytest.raises(ValueError, match="more neighbors"):
        O.build_graph_with_knn(torch.zeros(4, 3), 1, 4, None, 4)
    assert name == "ValueError" and "more neighbors" in msg


# ---------------------------------------------------------------------------------------------------------------------
# e3nn basis conventions, pinned by the reference-held e3nn constants (Jd.pt) and the vendored wigner_D
# ---------------------------------------------------------------------------------------------------------------------
def _z_rot(angle, l):
    m = torch.zeros(2 * l + 1, 2 * l + 1, dtype=torch.float64)
    inds, rev = torch.arange(2 * l + 1), torch.arange(2 * l, -1, -1)
    freq = torch.arange(l, -l - 1, -1, dtype=torch.float64)
    m[inds, rev] = torch.sin(freq * angle)
    m[inds, inds] = torch.cos(freq * angle)
    return m


def _wigner_D(Jd, l, a, b, c):
    """models/equiformer_v2/architecture/wigner.py:16-43 (borrowed there from e3nn 0.4.0 _wigner.py)."""
    return _z_rot(a, l) @ Jd[l] @ _z_rot(b, l) @ Jd[l] @ _z_rot(c, l)


def _rot_y(t):
    c, s = math.cos(t), math.sin(t)
    return torch.tensor([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=torch.float64)


def _rot_x(t):
    c, s = math.cos(t), math.sin(t)
    return torch.tensor([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=torch.float64)


def test_wigner_D_restatement_matches_reference_output():
    fx = load("ref_wigner.pt")
    for l in range(3):
        for k, (a, b, c) in enumerate(fx["angles"].tolist()):
            assert torch.allclose(_wigner_D(fx["Jd"], l, a, b, c), fx["D"][l][k], atol=1e-14)


def test_l1_basis_is_xyz_with_y_polar_axis():
    """D^1(alpha, beta, gamma) of the reference-held e3nn constants is the Cartesian rotation R_y(alpha) R_x(beta)
    R_y(gamma) acting on (x, y, z): e3nn's l = 1 components are (x, y, z) and the polar axis is y."""
    fx = load("ref_wigner.pt")
    for k, (a, b, c) in enumerate(fx["angles"].tolist()):
        assert torch.allclose(fx["D"][1][k], _rot_y(a) @ _rot_x(b) @ _rot_y(c), atol=1e-14)


def test_oracle_spherical_harmonics_transform_with_reference_wigner_D():
    """Y^l(R x) = D^l(R) Y^l(x) with the reference's D matrices: pins the oracle's real basis for l <= 2 (component
    order and relative signs) to e3nn's, up to one global factor per l, which the SymPy tests and the closed forms fix."""
    fx = load("ref_wigner.pt")
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(64, 3, generator=gen, dtype=torch.float64)
    for k, (a, b, c) in enumerate(fx["angles"].tolist()):
        R = fx["D"][1][k]
        y, yr = O.spherical_harmonics(2, x), O.spherical_harmonics(2, x @ R.T)
        for l in range(3):
            sl = slice(l * l, (l + 1) * (l + 1))
            assert torch.allclose(yr[:, sl], y[:, sl] @ fx["D"][l][k].T, atol=1e-13), (l, k)


@pytest.mark.parametrize("l1,l2,l3", [(0, 0, 0), (0, 1, 1), (1, 0, 1), (1, 1, 0), (1, 1, 2), (2, 0, 2), (2, 1, 1),
                                       (0, 2, 2), (2, 2, 0), (1, 1, 1), (2, 1, 2), (1, 2, 1), (2, 2, 2)])
def test_coupling_tensors_are_invariant_under_reference_wigner_D(l1, l2, l3):
    """C_{ijk} D1_{ii'} D2_{jj'} D3_{kk'} = C_{i'j'k'}: the space of invariant tensors is one-dimensional, so this
    fixes every coupling up to a scalar; unit Frobenius norm fixes the magnitude."""
    import segnn_b200.cg as cg
    fx = load("ref_wigner.pt")
    for name, table in (("oracle", O.wigner_3j(l1, l2, l3)), ("package", cg.real_wigner_3j(l1, l2, l3))):
        table = torch.as_tensor(table, dtype=torch.float64)
        assert abs(float(table.norm()) - 1.0) < 1e-12
        for k in range(fx["angles"].shape[0]):
            D1, D2, D3 = fx["D"][l1][k], fx["D"][l2][k], fx["D"][l3][k]
            rot = torch.einsum("ijk,ia,jb,kc->abc", table, D1, D2, D3)
            assert torch.allclose(rot, table, atol=1e-13), (name, k)


def test_gaunt_quadrature_couplings_equal_su2_formula():
    """Two derivations of e3nn's wigner_3j: SU(2) Clebsch-Gordan + change of basis (oracle) and normalised Gaunt
    integrals of the harmonics by exact quadrature (oracle/ref_shims/e3nn/o3) agree in sign and value."""
    import sys
    shims = os.path.join(os.path.dirname(os.path.dirname(__file__)), "oracle", "ref_shims")
    sys.path.insert(0, shims)
    try:
        for m in [m for m in sys.modules if m.split(".")[0] == "e3nn"]:
            del sys.modules[m]
        from e3nn import o3
        for key in [(0, 0, 0), (0, 1, 1), (1, 0, 1), (1, 1, 0), (1, 1, 2), (2, 0, 2), (2, 1, 1), (0, 2, 2), (2, 2, 2)]:
            assert torch.allclose(o3.wigner_3j(*key), O.wigner_3j(*key), atol=1e-13), key
        x = torch.randn(10, 3, dtype=torch.float64)
        assert torch.allclose(o3.spherical_harmonics(o3.Irreps.spherical_harmonics(2), x, True, "integral"),
                              O.spherical_harmonics(2, x), atol=1e-14)
    finally:
        sys.path.remove(shims)
        for m in [m for m in sys.modules if m.split(".")[0] == "e3nn"]:
            del sys.modules[m]


# ---------------------------------------------------------------------------------------------------------------------
# the reference's own module code: tensor products, SEGNN forward / backward, rollout
# ---------------------------------------------------------------------------------------------------------------------
def test_module_level_tensor_products_match_reference():
    fx = load("ref_tensor_products.pt")
    for case in fx["cases"]:
        cls = getattr(O, case["cls"])
        mod = cls(case["in1"], case["out"], case["in2"])
        mod.load_state_dict(case["state"])
        got = mod(case["x1"], case["x2"])
        assert rel(got, case["y"]) < 1e-12, case
        ref_instr = case["instructions"]
        assert [(a, b, c) for a, b, c, _ in ref_instr] == list(mod.tp.instructions)
        assert np.allclose([w for *_, w in ref_instr], mod.tp.path_weights, rtol=1e-14)
        assert torch.allclose(mod.sqrt_k_correction, case["sqrt_k_correction"].double())
    inorm = fx["instance_norm"]
    mod = O.InstanceNorm(inorm["irreps"])
    with torch.no_grad():
        mod.weight.copy_(inorm["weight"])
        mod.bias.copy_(inorm["bias"])
    assert rel(mod(inorm["x"], inorm["batch"]), inorm["y"]) < 1e-12


CASES = ["h64_n5", "h192_n8", "h128_n12", "h32_l2_n6", "h32_a2_n6", "h32_l2_a2_n5",  # a2: lmax_attr = 2
         "h32_knn3_n8", "h32_l2_knn2_n6",  # knn: num_neighbors < N - 1
         "h32_inorm_n6", "h32_l2_inorm_knn3_n7", "h32_nonorm_n5"]  # norm = "instance" / None


def oracle_model(fx):
    c = fx["config"]
    m = O.SEGNN(hidden_features=c["hidden_features"], lmax_h=c["lmax_h"], lmax_attr=c.get("lmax_attr", 1),
                num_layers=c["num_layers"], norm=c.get("norm", "batch"))
    m.load_state_dict(golden_state(fx["shapes"], fx["ranges"], fx["weight_seed"]))
    return m


@pytest.mark.parametrize("case", CASES)
def test_segnn_forward_backward_matches_reference(case):
    fx = load(f"ref_segnn_{case}.pt")
    c = fx["config"]
    B, N = c["B"], c["N"]
    m = oracle_model(fx)
    assert str(m.hidden_irreps) == fx["hidden_irreps"]
    assert sum(p.numel() for p in m.parameters()) == fx["num_params"]
    assert set(m.state_dict().keys()) == set(fx["state_keys"]) - set(fx["extra_keys"])
    # O3Transform (models/segnn/o3_building_blocks.py:230-278) + the edge order
    g = O.make_graph(fx["pos"], fx["vel"], fx["mass"], B, N, c.get("lmax_attr", 1), c.get("num_neighbors"))
    tr = fx["eval"]["transform"]
    assert torch.equal(g.edge_index, tr["edge_index"])
    for k in ("x", "edge_attr", "node_attr", "additional_message_features"):  # captured before SEGNN.forward
        assert rel(getattr(g, k), tr[k]) < 1e-13, k
    # eval-mode BatchNorm
    m.eval()
    with torch.no_grad():
        out, layers = m(g, return_layers=True)
    ref = fx["eval"]
    assert len(ref["layers"]) == len(layers) + 1  # the reference list also holds pre_pool1's output
    for i, (a, b) in enumerate(zip(layers, ref["layers"])):
        assert rel(a, b) < 1e-10, (case, "eval layer", i, rel(a, b))
    assert rel(out, ref["out"]) < 1e-10
    # train-mode BatchNorm, loss, gradients, running statistics
    m = oracle_model(fx).train()
    out, layers = m(O.make_graph(fx["pos"], fx["vel"], fx["mass"], B, N, c.get("lmax_attr", 1), c.get("num_neighbors")), return_layers=True)
    ref = fx["train"]
    for i, (a, b) in enumerate(zip(layers, ref["layers"])):
        assert rel(a, b) < 1e-10, (case, "train layer", i, rel(a, b))
    loss = O.target_common_loss(out, fx["y"])
    assert abs(float(loss) - ref["loss"]) < 1e-12 * abs(ref["loss"])
    loss.backward()
    floor = 1e-13 * max(ref["grad_norms"].values())  # biases in front of a train-mode BatchNorm have zero gradient
    for k, p in m.named_parameters():
        stride, vals = ref["grads"][k]
        got = p.grad.reshape(-1)[::stride]
        scale = ref["grad_norms"][k]
        assert float((got - vals).abs().max()) < 1e-9 * scale + floor, (case, k)
        assert abs(float(p.grad.norm()) - ref["grad_norms"][k]) < 1e-9 * scale + floor
    for k, v in ref["running"].items():
        assert rel(m.state_dict()[k], v) < 1e-12, k


def test_rollout_matches_reference_run_inference():
    """helper_scripts/infer_self_feed.py run_inference, executed on the real GravityDatasetOtf ground truth."""
    sm = load("ref_sim_macros.pt")
    fx = load("ref_segnn_h64_n5.pt")
    ro = sm["rollout"]
    loc, vel = ro["combined_locations"], ro["combined_velocities"]  # [2, B, T, N, 3]
    m = oracle_model(fx).eval()
    B, T, N = loc.shape[1:4]
    p_loc, p_vel = O.rollout(m, loc[0, :, 0], vel[0, :, 0], torch.ones(B, N, 1, dtype=torch.float64), T - 1)
    assert rel(p_loc, loc[1]) < 1e-9 and rel(p_vel, vel[1]) < 1e-9
    assert torch.equal(ro["loc_pred_sim_1"], loc[1, 1])
    assert ro["files"] == sorted(f"{q}_{w}_sim_{i}.npy" for q in ("loc", "vel") for w in ("actual", "pred")
                                 for i in range(B))
    # the ground truth half is the reference simulator's output: reproduce it from its own first frame
    meta = ro["metadata"]
    for b in range(B):
        ps, vs, _ = O.gravity_trajectory(loc[0, b, 0].numpy(), vel[0, b, 0].numpy(), np.ones((N, 1)),
                                         meta["interaction_strength"], meta["softening"], meta["dt"],
                                         meta["sim_length"], meta["sample_freq"])
        assert np.abs(ps - loc[0, b].numpy()).max() < 1e-10 and np.abs(vs - vel[0, b].numpy()).max() < 1e-10


# ---------------------------------------------------------------------------------------------------------------------
# simulators, macros, statistics
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("key", ["gravity", "gravity7"])
def test_gravity_simulator_matches_reference(key):
    fx = load("ref_sim_macros.pt")[key]
    p = fx["params"]
    ps, vs, fs = O.gravity_trajectory(fx["loc"][0].numpy(), fx["vel"][0].numpy(), fx["mass"].numpy(), p["G"],
                                      p["softening"], p["dt"], p["T"], p["sample_freq"])
    for got, ref in ((ps, fx["loc"]), (vs, fx["vel"]), (fs, fx["force"])):
        assert np.abs(got - ref.numpy()).max() < 1e-11 * max(1.0, float(ref.abs().max()))


def test_charged_simulator_matches_reference():
    fx = load("ref_sim_macros.pt")["charged"]
    X, V = O.charged_trajectory(fx["x0"].numpy(), fx["v0"].numpy(), fx["charges"].numpy(),
                                fx["params"]["interaction_strength"], fx["params"]["delta_t"], fx["X"].shape[0],
                                fx["params"]["max_F"])
    assert np.abs(X - fx["X"].numpy()).max() < 1e-11 and np.abs(V - fx["V"].numpy()).max() < 1e-11


def test_macro_counters_match_reference():
    mc = load("ref_sim_macros.pt")["macros"]
    loc, vel = mc["loc"].numpy(), mc["vel"].numpy()
    ev = O.event_counters(loc, vel)
    assert np.array_equal(ev["stickings"], mc["stickings"].numpy())
    assert np.array_equal(ev["collisions"], mc["collisions"].numpy())
    assert np.array_equal(ev["bodies_left"], mc["bodies_left_d15"].numpy())
    assert np.array_equal(ev["sharp_turns"], mc["sharp_turns_30"].numpy())
    assert np.allclose(ev["max_com_distance"], mc["max_com_distance"].numpy(), rtol=1e-13)
    ev2 = O.event_counters(loc, vel, time_threshold=2, contact_distance=1.0, leave_distance=1.4, turn_angle=90.0)
    assert np.array_equal(ev2["stickings"], mc["stickings_t2_d1"].numpy())
    assert np.array_equal(ev2["collisions"], mc["collisions_t2_d1"].numpy())
    assert np.array_equal(ev2["bodies_left"], mc["bodies_left_d2"].numpy())
    assert np.array_equal(ev2["sharp_turns"], mc["sharp_turns_90"].numpy())
    assert mc["stickings"].sum() > 0 and mc["collisions"].sum() > 0 and mc["bodies_left_d2"].sum() > 0
    for tt, dd in ((2, 2), (3, 1.5)):
        ref = mc[f"group_collisions_t{tt}_d{dd}"].numpy()
        assert np.array_equal(O.group_collision_counts(loc, tt, dd), ref)
        assert ref.sum() > 0
    kin, pot, series = O.nbody_energies(loc, vel, 2.0, 0.2)
    for k in ("potential", "kinetic", "total"):
        assert np.allclose(series[k], mc["energies"][k].numpy(), rtol=1e-13)
    assert np.allclose(O.momentum_magnitude(vel).mean(axis=1), mc["momentum_mean_over_time"].numpy(), rtol=1e-13)


def test_ks_fisher_and_noam_match_reference():
    fx = load("ref_sim_macros.pt")
    ks = fx["ks"]
    a, b, a_nan = ks["a"].numpy(), ks["b"].numpy(), ks["a_nan"].numpy()
    got = [O.ks_p(a, b), O.ks_p(a, a), O.ks_p(a_nan, b), O.ks_p(np.array([]), b)]
    for g, r in zip(got, ks["p"]):
        assert (g != g and r != r) or abs(g - r) <= 1e-15 * abs(r)
    assert abs(O.combine_pvalues_fisher(ks["fisher_inputs"]) - ks["fisher"]) <= 1e-12 * ks["fisher"]
    assert O.combine_pvalues_fisher([1e-200, 1e-250, 1e-100]) == ks["fisher_tiny"] == 1e-300
    assert O.combine_pvalues_fisher([float("nan")]) != O.combine_pvalues_fisher([float("nan")])  # nan
    assert ks["fisher_empty"] != ks["fisher_empty"]
    nm = fx["noam"]
    for s, r in zip((0, 1, 10, 2999, 3000, 3001, 100000), nm["rates"]):
        assert abs(O.noam_rate(s, nm["hidden"], nm["factor"], nm["warmup"]) - r) <= 1e-15 * r
