"""Offline charged N-body loader (dataloaders/segnn_nbody_offline_dataloader.py, datasets/nbody_offline/dataset.py)
against what the reference's own loader produced on the same files (tests/golden/ref_offline_loader.pt, written by
tests/golden/make_offline_golden.py).  CPU part: data set, batch order (PyG DataLoader semantics from the same torch
RNG state), test-partition rotation (same `random` stream), targets.  GPU part: the O3 attributes of preprocess_batch and
a model step through the loader."""
import os
import random
from types import SimpleNamespace

import pytest
import torch

import segnn_b200 as S

HERE = os.path.dirname(os.path.abspath(__file__))
FX = torch.load(os.path.join(HERE, "golden", "ref_offline_loader.pt"), weights_only=False)
DATA = os.path.join(HERE, "golden", "offline_small")


def _args(device, target="pos_dt+vel"):
    a = dict(FX["args"])
    a.update(data_directory=DATA, device=device, target=target)
    return SimpleNamespace(**a)


@pytest.mark.parametrize("part", ["train", "test"])
def test_batches_match_the_reference_loader(part):
    random.seed(11)
    torch.manual_seed(5)
    dl = S.SegnnNbodyOfflineDataloader(_args("cpu"), partition=part)
    ref = FX[part]
    assert len(dl) == ref["len"] and len(dl.dataset) == 6  # max_samples
    for rb in ref["batches"]:
        (b,), none = dl.get_batch()
        assert none is None
        for k, v in rb["raw"].items():
            got = getattr(b, k)
            assert got.shape == v.shape, k
            assert torch.equal(got, v) if k == "batch" else float((got - v).abs().max()) < 1e-6, k
        assert b.num_graphs * b.n_nodes * (b.n_nodes - 1) == rb["n_edges"]
    item, ritem = dl.dataset[2], ref["item2"]
    for k in ("loc_0", "vel_0", "node_feat", "node_attr", "loc_mean"):
        assert float((getattr(item, k) - ritem[k]).abs().max()) < 1e-6, k
    ei = dl.dataset.cutoff_edge(item.loc_0)
    assert torch.equal(ei, ritem["edge_index"])
    assert dl.dataset.get_serializable_attributes()["frame_T"] == 8


def test_incomplete_graphs_are_refused():
    a = _args("cpu")
    a.cutoff_rate = 0.3
    with pytest.raises(NotImplementedError):
        S.SegnnNbodyOfflineDataloader(a, partition="train")


@pytest.mark.gpu
@pytest.mark.parametrize("part", ["train", "test"])
def test_preprocess_matches_the_reference_loader(part):
    random.seed(11)
    torch.manual_seed(5)
    dl = S.SegnnNbodyOfflineDataloader(_args("cuda"), partition=part)
    for rb in FX[part]["batches"]:
        (b,), _ = dl.get_batch()
        g = dl.preprocess_batch(b, "cuda")
        for k in ("pos", "vel", "y", "mass", "x", "node_attr"):
            got, v = getattr(g, k).cpu(), rb[k]
            assert got.shape == v.shape, k
            if k == "node_attr":
                # K1 already applies SEGNN.catch_isolated_nodes (segnn.py:148: the l = 0 slot is overwritten with 1.0
                # at the top of forward); the reference's transform leaves 2 Y_0 there until then
                assert bool((got[:, 0] == 1.0).all())
                got, v = got[:, 1:], v[:, 1:]
            assert float((got - v).abs().max()) < 2e-6 * (1 + float(v.abs().max())), k


@pytest.mark.gpu
@pytest.mark.parametrize("target", ["pos", "pos_dt", "pos+vel"])
def test_targets(target):
    random.seed(11)
    torch.manual_seed(5)
    dl = S.SegnnNbodyOfflineDataloader(_args("cuda", target), partition="valid")
    (b,), _ = dl.get_batch()
    y = dl.preprocess_batch(b, "cuda").y.cpu()
    assert y.shape == FX[f"y_{target}"].shape and float((y - FX[f"y_{target}"]).abs().max()) < 1e-6


@pytest.mark.gpu
def test_training_step_through_the_offline_loader():
    torch.manual_seed(0)
    dl = S.SegnnNbodyOfflineDataloader(_args("cuda"), partition="train")
    model = S.SEGNN(hidden_features=64, num_layers=2).cuda().train()
    (b,), _ = dl.get_batch()
    g = dl.preprocess_batch(b, "cuda")
    pred = dl.postprocess_batch(model(g), "cuda")
    assert pred.shape == g.y.shape
    S.target_common_loss(pred, g.y).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
