"""Golden vectors for the offline charged N-body loader: writes a tiny data set in the reference's file layout
(tests/golden/offline_small/*.npy, seeded NumPy numbers -- the loader does not care where trajectories come from) and
runs the reference's own SegnnNbodyOfflineDataloader (dataloaders/segnn_nbody_offline_dataloader.py over
datasets/nbody_offline/dataset.py, imported from /root/reference through oracle/ref_loader.py) on it.  Run in the build
container:  python tests/golden/make_offline_golden.py  ->  tests/golden/ref_offline_loader.pt"""
import os
import pickle
import random
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
DATA = os.path.join(HERE, "offline_small")
NAME, SYSTEMS, FRAMES, N = "5_tiny", 7, 12, 5


def write_files():
    os.makedirs(DATA, exist_ok=True)
    for k, part in enumerate(("train", "valid", "test")):
        rng = np.random.default_rng(100 + k)
        suffix = f"{part}_charged{NAME}"
        np.save(f"{DATA}/loc_{suffix}.npy", rng.normal(size=(SYSTEMS, FRAMES, N, 3)))
        np.save(f"{DATA}/vel_{suffix}.npy", rng.normal(size=(SYSTEMS, FRAMES, N, 3)))
        np.save(f"{DATA}/charges_{suffix}.npy", rng.choice([-1.0, 1.0], size=(SYSTEMS, N, 1)))
        np.save(f"{DATA}/edges_{suffix}.npy", np.zeros((SYSTEMS, N, N)))
        with open(f"{DATA}/cfg_{suffix}.pkl", "wb") as f:
            pickle.dump({"n_balls": N}, f)


def loader_args(target="pos_dt+vel"):
    return SimpleNamespace(dataset_name=NAME, data_directory=DATA, virtual_channels=2, max_samples=6, frame_0=3,
                           frame_T=8, cutoff_rate=0.0, batch_size=4, lmax_attr=1, target=target, gpu_id=-1,
                           device="cpu")


def main():
    from oracle import ref_loader
    kind = ref_loader.setup()
    from dataloaders.segnn_nbody_offline_dataloader import SegnnNbodyOfflineDataloader
    write_files()
    out = dict(kind=kind, args=vars(loader_args()))
    for part in ("train", "test"):
        random.seed(11)
        torch.manual_seed(5)
        dl = SegnnNbodyOfflineDataloader(loader_args(), partition=part)
        rec = dict(len=len(dl), batches=[])
        for _ in range(3):  # crosses the end of a pass
            (b,), _ = dl.get_batch()
            raw = {k: getattr(b, k).clone() for k in ("loc_0", "loc_t", "vel_0", "vel_t", "node_attr", "batch")}
            g = dl.preprocess_batch(b, "cpu")
            rec["batches"].append(dict(raw=raw, pos=g.pos.clone(), vel=g.vel.clone(), y=g.y.clone(),
                                       mass=g.mass.clone(), x=g.x.clone(), node_attr=g.node_attr.clone(),
                                       n_edges=int(g.edge_index.shape[1])))
        item = dl.dataset[2]
        rec["item2"] = {k: getattr(item, k).clone() for k in ("loc_0", "vel_0", "node_feat", "node_attr", "loc_mean",
                                                               "edge_index", "edge_attr")}
        out[part] = rec
    for target in ("pos", "pos_dt", "pos+vel"):
        random.seed(11)
        torch.manual_seed(5)
        dl = SegnnNbodyOfflineDataloader(loader_args(target), partition="valid")
        (b,), _ = dl.get_batch()
        out[f"y_{target}"] = dl.preprocess_batch(b, "cpu").y.clone()
    torch.save(out, os.path.join(HERE, "ref_offline_loader.pt"))
    print("wrote ref_offline_loader.pt, provider:", kind)


if __name__ == "__main__":
    main()
