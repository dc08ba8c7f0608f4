"""Runs the UNMODIFIED reference from /root/reference live (through oracle/ref_loader.py) next to the oracle on fresh
random inputs.  Only where the reference checkout exists (the build container); skipped on the GPU box, where the
committed fixtures of tests/golden/ stand in for it (tests/test_reference_golden.py)."""
import subprocess
import sys
import os

import pytest

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_SCRIPT = r'''
import sys, torch
sys.path.insert(0, %(root)r)
from oracle import ref_loader
kind = ref_loader.setup()
torch.set_default_dtype(torch.float64)
from models.segnn.segnn import SEGNN
from models.segnn.o3_building_blocks import O3Transform
from utils.build_fully_connected_graph import build_graph_with_knn
from torch_geometric.data import Data
from oracle import segnn_oracle as O
worst = 0.0
for H, lmax_h, L, B, N, seed in [(64, 1, 3, 4, 5, 0), (48, 2, 2, 2, 7, 1), (96, 1, 2, 1, 16, 2)]:
    torch.manual_seed(seed)
    ref = SEGNN(hidden_features=H, lmax_h=lmax_h, num_layers=L).double()
    om = O.SEGNN(hidden_features=H, lmax_h=lmax_h, num_layers=L)
    missing = om.load_state_dict({k: v for k, v in ref.state_dict().items() if "output_mask" not in k})
    O.perturb_bn_buffers(om, seed)
    ref.load_state_dict(om.state_dict(), strict=False)
    pos, vel, mass = O.synthetic_system(B, N, seed=seed + 3)
    pos, vel, mass = pos.reshape(-1, 3), vel.reshape(-1, 3), mass.reshape(-1, 1)
    for train in (False, True):
        ref.train(train); om.train(train)
        g = Data(pos=pos.clone(), vel=vel.clone(), force=torch.zeros_like(pos), mass=mass.clone())
        g.batch = torch.arange(B).repeat_interleave(N)
        g.edge_index = build_graph_with_knn(g.pos, B, N, torch.device("cpu"), N - 1)
        a = ref(O3Transform(1)(g))
        b = om(O.make_graph(pos, vel, mass, B, N))
        worst = max(worst, float((a - b).abs().max() / b.abs().max()))
        if train:
            ga = torch.autograd.grad(a.pow(2).sum(), ref.layers[0].message_layer_1.tp.weight)[0]
            gb = torch.autograd.grad(b.pow(2).sum(), om.layers[0].message_layer_1.tp.weight)[0]
            worst = max(worst, float((ga - gb).abs().max() / gb.abs().max()))
# O3Transform(lmax_attr = 2, use_force_input = True) (o3_building_blocks.py:267-271), on a kNN graph
from types import SimpleNamespace
torch.manual_seed(9)
B, N, k = 2, 7, 3
pos, vel, mass = O.synthetic_system(B, N, seed=4)
pos, vel, mass = pos.reshape(-1, 3), vel.reshape(-1, 3), mass.reshape(-1, 1)
force = torch.randn_like(pos)
g = Data(pos=pos.clone(), vel=vel.clone(), force=force.clone(), mass=mass.clone())
g.edge_index = build_graph_with_knn(g.pos, B, N, torch.device("cpu"), k)
g = O3Transform(2, use_force_input=True)(g)
og = SimpleNamespace(pos=pos, vel=vel, mass=mass, force=force, edge_index=O.build_graph_with_knn(pos, B, N, None, k))
og = O.o3_transform(og, 2, use_force_input=True)
assert torch.equal(g.edge_index, og.edge_index)
for key in ("node_attr", "edge_attr", "x", "additional_message_features"):
    worst = max(worst, float((getattr(g, key) - getattr(og, key)).abs().max()))
print("KIND", kind, "WORST", worst)
assert worst < 1e-10, worst
'''


def test_live_reference_segnn_equals_oracle():
    """Own process: ref_loader rewires sys.modules (models, utils, matplotlib ...), which must not leak into pytest."""
    out = subprocess.run([sys.executable, "-c", _SCRIPT % {"root": ROOT}], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "WORST" in out.stdout
