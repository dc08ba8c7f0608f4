"""Generates the reference-run golden vectors under tests/golden/ref_*.pt.

Run HERE (build container), where /root/reference exists:

    python tests/golden/make_reference_golden.py

Every vector is produced by importing and executing the reference's OWN code from /root/reference through
``oracle/ref_loader.py``: ``models.segnn.segnn.SEGNN`` / ``SEGNNLayer``, ``O3TensorProduct[SwishGate]``, ``O3Transform``,
``WeightBalancedIrreps``, ``InstanceNorm``, ``build_graph_with_knn``, ``run_inference``, ``GravityDatasetOtf`` /
``GravitySim``, the charged ``System``, the macro counters of ``visualization_utils``, ``Trainer._compute_nbody_energies``
/ ``Trainer._rate``, ``utils.ks_utils``, ``training.losses.TargetCommonLoss`` and the vendored e3nn ``wigner_D`` +
``Jd.pt``.  The third-party packages e3nn / torch_geometric / torch_scatter are NOT installable here; the fixture
records which provider ran (``kind``): 'reference' = real packages, 'reference+shims' = the stand-ins of
``oracle/ref_shims`` (see its README for what that does and does not prove).

The fixtures travel to the GPU box (the reference does not); tests/test_reference_golden.py consumes them.
"""
import json
import os
import sys
import tempfile
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import ref_loader  # noqa: E402
from golden_weights import golden_state, key_range, sample_grad  # noqa: E402

KIND = ref_loader.setup()
torch.set_default_dtype(torch.float64)

from models.segnn.segnn import SEGNN  # noqa: E402
from models.segnn.o3_building_blocks import O3Transform, O3TensorProduct, O3TensorProductSwishGate  # noqa: E402
from models.segnn.instance_norm import InstanceNorm  # noqa: E402
from utils.build_fully_connected_graph import build_graph_with_knn  # noqa: E402
from torch_geometric.data import Data  # noqa: E402
from e3nn.o3 import Irreps  # noqa: E402
from training.losses import TargetCommonLoss  # noqa: E402


def save(name, obj):
    path = os.path.join(HERE, name)
    torch.save(obj, path)
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.1f} KiB")


def synthetic(B, N, seed, charged):
    gen = torch.Generator().manual_seed(seed)
    pos = torch.randn(B * N, 3, generator=gen) * (N / 5.0) ** (1 / 3)
    vel = torch.randn(B, N, 3, generator=gen)
    vel = (vel - vel.mean(1, keepdim=True)).reshape(B * N, 3)
    mass = (torch.randint(0, 2, (B * N, 1), generator=gen).double() * 2 - 1) if charged else torch.ones(B * N, 1)
    y = torch.randn(B * N, 6, generator=gen)
    return pos, vel, mass, y


def ref_graph(pos, vel, mass, B, N, lmax_attr=1, num_neighbors=None):
    g = Data(pos=pos.clone(), vel=vel.clone(), force=torch.zeros_like(pos), mass=mass.clone())
    g.batch = torch.arange(B).repeat_interleave(N)
    g.edge_index = build_graph_with_knn(g.pos, B, N, torch.device("cpu"), N - 1 if num_neighbors is None else num_neighbors)
    return O3Transform(lmax_attr)(g)


# ------------------------------------------------------------------------------------------------------------------
def make_graph_fixture():
    out = {"kind": KIND, "full": {}, "knn": []}
    for B, N in [(1, 2), (2, 3), (3, 5), (2, 8), (1, 33)]:
        out["full"][(B, N)] = build_graph_with_knn(torch.zeros(B * N, 3), B, N, torch.device("cpu"), N - 1)
    gen = torch.Generator().manual_seed(5)
    for B, N, k in [(2, 6, 2), (3, 7, 3), (1, 12, 5)]:
        loc = torch.randn(B * N, 3, generator=gen)
        out["knn"].append({"B": B, "N": N, "k": k, "loc": loc,
                           "edge_index": build_graph_with_knn(loc, B, N, torch.device("cpu"), k)})
    try:
        build_graph_with_knn(torch.zeros(4, 3), 1, 4, torch.device("cpu"), 4)
        out["too_many_neighbors"] = None
    except Exception as e:  # noqa: BLE001
        out["too_many_neighbors"] = (type(e).__name__, str(e))
    save("ref_graph.pt", out)


def make_wigner_fixture():
    w = ref_loader.load_wigner()
    gen = torch.Generator().manual_seed(11)
    angles = (torch.rand(8, 3, generator=gen) * 2 - 1) * torch.tensor([3.1, 1.5, 3.1])
    out = {"source": "models/equiformer_v2/architecture/{wigner.py,Jd.pt} (e3nn 0.4.0 code + constants)",
           "Jd": [w._Jd[l].double().clone() for l in range(3)], "angles": angles,
           "D": [w.wigner_D(l, angles[:, 0], angles[:, 1], angles[:, 2]).clone() for l in range(3)]}
    save("ref_wigner.pt", out)


def run_model_case(name, H, lmax_h, L, B, N, seed, charged=True, lmax_attr=1, num_neighbors=None, norm="batch"):
    torch.manual_seed(seed)
    model = SEGNN(hidden_features=H, lmax_h=lmax_h, lmax_attr=lmax_attr, num_layers=L, norm=norm).double()
    sd = model.state_dict()
    extra_keys = sorted(k for k in sd if "output_mask" in k or k.endswith("num_batches_tracked"))
    shapes = {k: tuple(v.shape) for k, v in sd.items() if k not in extra_keys}
    ranges = {k: key_range(k, float(sd[k].abs().max())) for k in shapes}
    state = golden_state(shapes, ranges, seed)
    model.load_state_dict({**{k: sd[k] for k in extra_keys}, **state})
    pos, vel, mass, y = synthetic(B, N, seed + 1, charged)

    def forward(train: bool):
        model.train(train)
        model.load_state_dict({**{k: sd[k] for k in extra_keys}, **state})
        model.zero_grad(set_to_none=True)
        layers, hooks = [], []
        mods = [model.embedding_layer] + list(model.layers) + [model.pre_pool1]
        for m in mods:
            hooks.append(m.register_forward_hook(lambda _m, _i, o: layers.append(o.detach().clone())))
        g = ref_graph(pos, vel, mass, B, N, lmax_attr, num_neighbors)
        g.y = y
        transform = {k: getattr(g, k).detach().clone() for k in ("x", "node_attr", "edge_attr",
                                                                 "additional_message_features", "edge_index")}
        out = model(g)
        for h in hooks:
            h.remove()
        res = {"layers": layers, "out": out.detach().clone(), "transform": transform}
        if train:
            args = SimpleNamespace(target="pos_dt+vel", position_loss_weight=1.0, velocity_loss_weight=1.0,
                                   force_loss_weight=1.0)
            loss = TargetCommonLoss(args)(out, g)
            loss.backward()
            res["loss"] = float(loss)
            res["grads"] = {k: sample_grad(p.grad) for k, p in model.named_parameters()}
            res["grad_norms"] = {k: float(p.grad.norm()) for k, p in model.named_parameters()}
            res["running"] = {k: v.detach().clone() for k, v in model.state_dict().items() if "running_" in k}
        return res

    with torch.no_grad():
        ev = forward(False)
    tr = forward(True)
    config = dict(hidden_features=H, lmax_h=lmax_h, num_layers=L, B=B, N=N, charged=charged)
    if lmax_attr != 1:  # the key is absent from the lmax_attr = 1 fixtures written before the option existed
        config["lmax_attr"] = lmax_attr
    if num_neighbors is not None:  # kNN graph (utils/build_fully_connected_graph.py:42-80)
        config["num_neighbors"] = num_neighbors
    if norm != "batch":  # segnn.py:226-237
        config["norm"] = norm
    fx = {"kind": KIND, "config": config,
          "hidden_irreps": str(model.hidden_irreps), "num_params": sum(p.numel() for p in model.parameters()),
          "serializable": {k: v for k, v in model.get_serializable_attributes().items()},
          "state_keys": list(sd.keys()), "extra_keys": extra_keys,
          "weight_seed": seed, "shapes": shapes, "ranges": ranges,
          "pos": pos, "vel": vel, "mass": mass, "y": y, "eval": ev, "train": tr}
    save(f"ref_segnn_{name}.pt", fx)
    model.load_state_dict({**{k: sd[k] for k in extra_keys}, **state})  # undo the running-stat update of the train pass
    return model


def make_tp_fixture():
    """Module-level tensor products on their own (o3_building_blocks.py:10-203), incl. non-hidden irreps."""
    cases = []
    gen = torch.Generator().manual_seed(3)
    specs = [("O3TensorProduct", "2x1o+1x0e", "8x0e+8x1o", "1x0e+1x1o"),
             ("O3TensorProductSwishGate", "8x0e+8x1o+8x0e+8x1o+2x0e", "8x0e+8x1o", "1x0e+1x1o"),
             ("O3TensorProductSwishGate", "6x0e+6x1o+6x2e", "6x0e+6x1o+6x2e", "1x0e+1x1o"),
             ("O3TensorProduct", "8x0e+8x1o", "2x1o", "1x0e+1x1o"),
             ("O3TensorProduct", "5x0e+3x1o", "4x0e", None),
             ("O3TensorProductSwishGate", "5x0e", "7x0e", None)]
    for cls_name, in1, out, in2 in specs:
        cls = {"O3TensorProduct": O3TensorProduct, "O3TensorProductSwishGate": O3TensorProductSwishGate}[cls_name]
        torch.manual_seed(17)
        mod = cls(Irreps(in1), Irreps(out), Irreps(in2) if in2 else None).double()
        sd = {k: v.clone() for k, v in mod.state_dict().items() if "output_mask" not in k}
        x1 = torch.randn(9, Irreps(in1).dim, generator=gen)
        x2 = torch.randn(9, Irreps(in2).dim, generator=gen) if in2 else None
        y = mod(x1, x2)
        cases.append({"cls": cls_name, "in1": in1, "out": out, "in2": in2, "state": sd, "x1": x1, "x2": x2,
                      "y": y.detach().clone(), "tp_irreps_out": str(mod.tp.irreps_out),
                      "sqrt_k_correction": mod.sqrt_k_correction.clone(),
                      "instructions": [(i.i_in1, i.i_in2, i.i_out, float(i.path_weight)) for i in mod.tp.instructions]})
    # InstanceNorm (models/segnn/instance_norm.py) on a ragged batch
    inorm = InstanceNorm(Irreps("4x0e+3x1o")).double()
    with torch.no_grad():
        inorm.weight.copy_(torch.rand(7, generator=gen) + 0.5)
        inorm.bias.copy_(torch.randn(4, generator=gen) * 0.1)
    xin = torch.randn(11, 13, generator=gen)
    batch = torch.tensor([0] * 4 + [1] * 5 + [2] * 2)
    save("ref_tensor_products.pt", {"kind": KIND, "cases": cases,
                                    "instance_norm": {"irreps": "4x0e+3x1o", "weight": inorm.weight.detach().clone(),
                                                      "bias": inorm.bias.detach().clone(), "x": xin, "batch": batch,
                                                      "y": inorm(xin, batch).detach().clone()}})


def make_rollout_fixture(model):
    """The real run_inference (helper_scripts/infer_self_feed.py:21-254) with the real GravityDatasetOtf."""
    from helper_scripts.infer_self_feed import run_inference
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            run_dir = os.path.join(tmp, "runs", "segnn", "2025-01-01_00-00-00")
            os.makedirs(os.path.join(run_dir, "nbody_small_dataset"))
            meta = {"dataset_name": "nbody_small", "target": "pos_dt+vel", "path": os.path.join(tmp, "data"),
                    "batch_size": 3, "sim_length": 120, "sample_freq": 10, "noise_var": 0, "n_balls": 5,
                    "vel_norm": 1e-16, "interaction_strength": 2, "dt": 0.01, "softening": 0.2,
                    "double_precision": True, "center_of_mass": False}
            with open(os.path.join(run_dir, "nbody_small_dataset", "metadata.json"), "w") as f:
                json.dump(meta, f)
            model.eval()
            d, loc, vel = run_inference("segnn", None, model_path=os.path.join(run_dir, "model.pth"), model=model,
                                        save_dir=os.path.join(tmp, "out"), print_step=False, n_bodies=5,
                                        device=torch.device("cpu"))
            files = sorted(os.listdir(d))
            sample = np.load(os.path.join(d, "loc_pred_sim_1.npy"))
        finally:
            os.chdir(cwd)
    return {"metadata": meta, "combined_locations": torch.tensor(np.asarray(loc)),
            "combined_velocities": torch.tensor(np.asarray(vel)), "files": files,
            "loc_pred_sim_1": torch.tensor(sample)}


def make_sim_and_macro_fixture(rollout):
    from datasets.nbody.dataset.synthetic_sim import GravitySim
    from datasets.nbody import visualization_utils as vu
    import trainer as ref_trainer
    from utils.ks_utils import _combine_pvalues_fisher, _ks_p
    import system as charged_system

    out = {"kind": KIND, "rollout": rollout}
    # --- GravitySim (datasets/nbody/dataset/synthetic_sim.py:305-420), the dataset's parameters
    sim = GravitySim(noise_var=0, n_balls=5, vel_norm=1e-16, interaction_strength=2, dt=0.01, softening=0.2)
    loc, vel, force, mass = sim.sample_trajectory(T=400, sample_freq=10, random_seed=7)
    out["gravity"] = {"params": dict(G=2, softening=0.2, dt=0.01, T=400, sample_freq=10),
                      "loc": torch.tensor(loc), "vel": torch.tensor(vel), "force": torch.tensor(force),
                      "mass": torch.tensor(mass)}
    sim7 = GravitySim(noise_var=0, n_balls=7, vel_norm=1e-16, interaction_strength=2, dt=0.01, softening=0.2)
    loc7, vel7, force7, mass7 = sim7.sample_trajectory(T=100, sample_freq=5, random_seed=3)
    out["gravity7"] = {"params": dict(G=2, softening=0.2, dt=0.01, T=100, sample_freq=5),
                       "loc": torch.tensor(loc7), "vel": torch.tensor(vel7), "force": torch.tensor(force7),
                       "mass": torch.tensor(mass7)}
    # --- charged System (datasets/nbody_offline/datagen/system.py:6-123), isolated bodies only
    np.random.seed(10)
    sys_ = charged_system.System(n_isolated=6, n_stick=0, n_hinge=0, delta_t=0.001, box_size=None, loc_std=1.0,
                                 vel_norm=0.5, interaction_strength=1.0)
    x0, v0 = sys_.X.copy(), sys_.V.copy()
    xs, vs = [], []
    for _ in range(300):
        sys_.simulate_one_step()
        xs.append(sys_.X.copy())
        vs.append(sys_.V.copy())
    out["charged"] = {"params": dict(delta_t=0.001, interaction_strength=1.0, max_F=sys_._max_F),
                      "charges": torch.tensor(sys_.charges), "x0": torch.tensor(x0), "v0": torch.tensor(v0),
                      "order": [o.node_idx[0] for o in sys_.physical_objects],
                      "X": torch.tensor(np.stack(xs)), "V": torch.tensor(np.stack(vs))}
    # --- macros on synthetic crowded trajectories
    rng = np.random.default_rng(21)
    S, T, N = 5, 48, 6
    steps = rng.normal(size=(S, T, N, 3)) * 0.12
    loc_m = rng.normal(size=(S, 1, N, 3)) * 0.8 + np.cumsum(steps, axis=1)
    vel_m = rng.normal(size=(S, T, N, 3))
    vel_m[:, 1::3] = vel_m[:, 0:-1:3] * (1 + 0.1 * rng.normal(size=vel_m[:, 0:-1:3].shape))  # some small turns
    stick, coll = vu.count_stickings_and_collisions(loc_m)
    stick2, coll2 = vu.count_stickings_and_collisions(loc_m, time_threshold=2, distance_threshold=1.0)
    macros = {"loc": torch.tensor(loc_m), "vel": torch.tensor(vel_m),
              "stickings": torch.tensor(stick), "collisions": torch.tensor(coll),
              "stickings_t2_d1": torch.tensor(stick2), "collisions_t2_d1": torch.tensor(coll2),
              "bodies_left_d15": torch.tensor(vu.count_balls_leaving_defined_area(loc_m)),
              "bodies_left_d2": torch.tensor(vu.count_balls_leaving_defined_area(loc_m, distance_threshold=1.4)),
              "max_com_distance": torch.tensor(vu.get_max_distance_of_com_from_starting_position(loc_m)),
              "sharp_turns_30": torch.tensor(vu.count_sharp_turns(vel_m)),
              "sharp_turns_90": torch.tensor(vu.count_sharp_turns(vel_m, angle_threshold=90))}
    with tempfile.TemporaryDirectory() as tmp:
        for (tt, dd) in [(2, 2), (3, 1.5)]:
            vu.plot_group_collision_distribution_multiplot(np.stack([loc_m, loc_m[::-1]]), time_threshold=tt,
                                                           distance_threshold=dd, save_dir=tmp,
                                                           title_suffixes=["a", "b"])
            with open(os.path.join(tmp, "group_collision_distribution.json")) as f:
                js = json.load(f)
            macros[f"group_collisions_t{tt}_d{dd}"] = torch.tensor(js["a"]["group_collision_count"])
        vu.plot_momentum_statistics(np.stack([vel_m, vel_m]), save_dir=tmp, title_suffixes=["a", "b"])
        with open(os.path.join(tmp, "momentum_statistics.json")) as f:
            macros["momentum_mean_over_time"] = torch.tensor(json.load(f)["a"]["momentum_statistics"])
    en = ref_trainer.Trainer._compute_nbody_energies(None, loc_m, vel_m, 2.0, 0.2)
    macros["energies"] = {k: torch.tensor(v) for k, v in en.items()}
    out["macros"] = macros
    # --- KS / Fisher (utils/ks_utils.py)
    a, b = rng.normal(size=200), rng.normal(loc=0.15, size=180)
    a_nan = a.copy()
    a_nan[::7] = np.nan
    ps = [_ks_p(a, b), _ks_p(a, a), _ks_p(a_nan, b), _ks_p(np.array([]), b), 1e-30, float("nan"), 0.0]
    out["ks"] = {"a": torch.tensor(a), "b": torch.tensor(b), "a_nan": torch.tensor(a_nan),
                 "p": [float(p) for p in ps[:4]], "fisher_inputs": ps,
                 "fisher": float(_combine_pvalues_fisher(ps)),
                 "fisher_tiny": float(_combine_pvalues_fisher([1e-200, 1e-250, 1e-100])),
                 "fisher_empty": float(_combine_pvalues_fisher([float("nan")]))}
    # --- Noam schedule (trainer.py:179-195) with the README model size
    fake = SimpleNamespace(model=SimpleNamespace(get_model_size=lambda: 192))
    out["noam"] = {"hidden": 192, "factor": 1.0, "warmup": 3000,
                   "rates": [ref_trainer.Trainer._rate(fake, s, 1.0, 3000) for s in (0, 1, 10, 2999, 3000, 3001, 100000)]}
    save("ref_sim_macros.pt", out)


def make_checkpoint_fixture():
    """A checkpoint written by the reference's own Trainer.save_model / save_model_params / save_dataset_attributes
    (trainer.py:526-537,599-612) after two real AdamW steps on a tiny SEGNN."""
    import shutil
    import trainer as ref_trainer
    from datasets.nbody.dataset_gravity_otf import GravityDatasetOtf
    torch.manual_seed(5)
    model = SEGNN(hidden_features=16, num_layers=1).double()
    fake = SimpleNamespace(model=model)
    opt = ref_trainer.Trainer.create_optimizer(SimpleNamespace(model=model, args=SimpleNamespace(learning_rate=1.0)))
    fake.optimizer = opt
    fake.args = SimpleNamespace(learning_rate_factor=1.0, learning_rate_warmup_steps=100, dataset_name="nbody_small")
    fake._rate = lambda step, factor, warmup: ref_trainer.Trainer._rate(fake, step, factor, warmup)
    fake.lr_scheduler = ref_trainer.Trainer.create_lr_scheduler(fake)
    pos, vel, mass, y = synthetic(2, 5, 9, False)
    for _ in range(2):
        opt.zero_grad()
        g = ref_graph(pos, vel, mass, 2, 5)
        (model(g) - y).pow(2).mean().backward()
        opt.step()
        fake.lr_scheduler.step()
    fake.step_count, fake.best_metrics = 2, {"valid_loss": 0.25}
    out_dir = os.path.join(HERE, "ref_run", "2025-01-02_03-04-05")
    shutil.rmtree(os.path.join(HERE, "ref_run"), ignore_errors=True)
    fake.save_dir_path = out_dir
    ref_trainer.Trainer.save_model(fake)
    ref_trainer.Trainer.save_model_params(fake)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            ds = GravityDatasetOtf(batch_size=2, sim_length=40, num_nodes=5, double_precision=True, cache_data=False,
                                   path=os.path.join(tmp, "data"))
        finally:
            os.chdir(cwd)
    fake.dataloader = SimpleNamespace(dataset=ds)
    ref_trainer.Trainer.save_dataset_attributes(fake)
    meta_path = os.path.join(out_dir, "nbody_small_dataset", "metadata.json")
    meta = json.load(open(meta_path))
    meta["path"] = "datasets/nbody/dataset/gravity"  # the absolute path of this container means nothing elsewhere
    json.dump(meta, open(meta_path, "w"), indent=4)
    model.eval()
    with torch.no_grad():
        pred = model(ref_graph(pos, vel, mass, 2, 5))
    save("ref_checkpoint_io.pt", {"kind": KIND, "pos": pos, "vel": vel, "mass": mass, "pred_eval": pred,
                                  "lr": opt.param_groups[0]["lr"], "run_dir": "ref_run/2025-01-02_03-04-05"})
    print("wrote", out_dir, os.listdir(out_dir))


def make_lmax_attr2_fixtures():
    """lmax_attr = 2 (l <= 2 steering attributes, models/segnn/segnn.py:22,36,47): the branch no BASELINE configuration
    takes; written separately so the fixtures above stay byte-identical."""
    run_model_case("h32_a2_n6", H=32, lmax_h=1, L=2, B=2, N=6, seed=5, lmax_attr=2)
    run_model_case("h32_l2_a2_n5", H=32, lmax_h=2, L=2, B=2, N=5, seed=6, lmax_attr=2)


def make_knn_fixtures():
    """SEGNN on kNN graphs (num_neighbors < N - 1, utils/build_fully_connected_graph.py:42-80; variable in-degree)."""
    run_model_case("h32_knn3_n8", H=32, lmax_h=1, L=2, B=2, N=8, seed=7, num_neighbors=3)
    run_model_case("h32_l2_knn2_n6", H=32, lmax_h=2, L=2, B=3, N=6, seed=8, num_neighbors=2)


def make_instance_norm_fixtures():
    """norm='instance' (models/segnn/instance_norm.py on the node features, no message norm) and norm=None."""
    run_model_case("h32_inorm_n6", H=32, lmax_h=1, L=2, B=3, N=6, seed=9, norm="instance")
    run_model_case("h32_l2_inorm_knn3_n7", H=32, lmax_h=2, L=2, B=2, N=7, seed=10, norm="instance", num_neighbors=3)
    run_model_case("h32_nonorm_n5", H=32, lmax_h=1, L=2, B=2, N=5, seed=11, norm=None)


if __name__ == "__main__":
    print("third-party provider:", KIND)
    if sys.argv[1:] == ["knn"]:
        make_knn_fixtures()
        sys.exit(0)
    if sys.argv[1:] == ["instance_norm"]:
        make_instance_norm_fixtures()
        sys.exit(0)
    if sys.argv[1:] == ["lmax_attr2"]:
        make_lmax_attr2_fixtures()
        sys.exit(0)
    make_graph_fixture()
    make_wigner_fixture()
    make_tp_fixture()
    model_a = run_model_case("h64_n5", H=64, lmax_h=1, L=2, B=3, N=5, seed=1)
    run_model_case("h192_n8", H=192, lmax_h=1, L=6, B=2, N=8, seed=2)
    run_model_case("h128_n12", H=128, lmax_h=1, L=2, B=1, N=12, seed=3, charged=False)
    run_model_case("h32_l2_n6", H=32, lmax_h=2, L=2, B=2, N=6, seed=4)
    rollout = make_rollout_fixture(model_a)
    make_sim_and_macro_fixture(rollout)
    make_checkpoint_fixture()
    make_lmax_attr2_fixtures()
    make_knn_fixtures()
    make_instance_norm_fixtures()
