"""Generates tests/golden/segnn_small.pt from the CPU oracle (oracle/segnn_oracle.py).

PARITY UNPINNED: the reference ships no golden vectors for SEGNN and its dependencies (e3nn, torch_geometric,
torch_scatter) cannot be imported in the build container, so these vectors pin the ORACLE (regression guard and a
fixture that travels to the GPU box), not the real e3nn stack.  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import segnn_oracle as O  # noqa: E402


def main():
    torch.manual_seed(1234)
    H, L, B, N = 16, 2, 2, 4
    model = O.SEGNN(hidden_features=H, num_layers=L).eval()
    O.perturb_bn_buffers(model, seed=5)
    pos, vel, mass = O.synthetic_system(B, N, seed=7, charged=True)
    g = O.make_graph(pos.reshape(-1, 3), vel.reshape(-1, 3), mass.reshape(-1, 1), B, N)
    node_attr_transform = g.node_attr.clone()  # O3Transform output, before SEGNN.forward forces [:,0] = 1
    with torch.no_grad():
        pred, layers = model(g, return_layers=True)
        loc, velr = O.rollout(model, pos, vel, mass, steps=3)
    out = dict(
        config=dict(hidden_features=H, num_layers=L, B=B, N=N, lmax_h=1),
        state_dict={k: v.clone() for k, v in model.state_dict().items()},
        pos=pos, vel=vel, mass=mass,
        edge_index=g.edge_index, edge_attr=g.edge_attr, node_attr=node_attr_transform, x=g.x,
        additional_message_features=g.additional_message_features,
        pred=pred, layers=layers, rollout_loc=loc, rollout_vel=velr,
    )
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "segnn_small.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
