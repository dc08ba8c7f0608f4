"""CPU oracle for the SEGNN hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import this file.  The product package must
never route through it.

HOW THIS ORACLE IS PINNED (tests/test_reference_golden.py, tests/test_reference_live.py, tests/test_oracle.py):

1. Against the reference's OWN code, executed unmodified from /root/reference in the build container
   (tests/golden/make_reference_golden.py -> tests/golden/ref_*.pt): models/segnn/segnn.py (SEGNN, SEGNNLayer),
   o3_building_blocks.py (O3TensorProduct[SwishGate], O3Transform), balanced_irreps.py, instance_norm.py,
   utils/build_fully_connected_graph.py (both branches), helper_scripts/infer_self_feed.py::run_inference on the real
   GravityDatasetOtf, GravitySim, the charged System, the macro counters incl. group collisions, Trainer energies and
   Noam rate, utils/ks_utils.py, training/losses.py.  Per-layer outputs (eval and train BatchNorm), loss, gradients,
   running statistics, rollouts and macros agree to <= 1e-10.
2. The arithmetic of that path lives in un-vendored third-party packages (e3nn==0.5.1, torch_geometric==2.6.1,
   torch_scatter==2.1.2 -- /root/reference/requirements.txt:9,21,22) that cannot be installed here or on the GPU box
   (no network, not in /opt/wheelhouse: profiles/r2_ref_env_attempt.log).  In (1) they are provided by the stand-ins of
   oracle/ref_shims (fixtures record kind = 'reference+shims'); with the real packages present the same script
   records 'reference'.  The e3nn CONVENTIONS are pinned by material that is not ours: the e3nn constants the
   reference vendors (models/equiformer_v2/architecture/Jd.pt + wigner.py: the l = 1 basis is (x, y, z), the
   spherical harmonics transform with e3nn's Wigner D, every coupling tensor is D-invariant with unit norm), SymPy
   (real harmonics Znm, real Gaunt coefficients: basis order, signs, values) and a second derivation of the couplings
   by exact quadrature.  What remains 'recalled, not executed' is e3nn's normalisation choices (path weights,
   normalize2mom constants, BatchNorm conventions), restated twice independently (here and in the shims).

What is restated, with the reference call site each piece follows:

* ``Irreps``                      e3nn.o3.Irreps as used in models/segnn/segnn.py:37-45,209-210
* ``weight_balanced_irreps``      models/balanced_irreps.py:51-85
* ``wigner_3j``                   e3nn.o3.wigner_3j (real basis, unit Frobenius norm)
* ``FullyConnectedTensorProduct`` e3nn.o3.FullyConnectedTensorProduct, called at
                                  models/segnn/o3_building_blocks.py:43-49
* ``O3TensorProduct``             models/segnn/o3_building_blocks.py:10-167
* ``Gate`` / ``normalize2mom``    e3nn.nn.Gate, called at o3_building_blocks.py:186-193
* ``O3TensorProductSwishGate``    models/segnn/o3_building_blocks.py:170-203
* ``BatchNorm``                   e3nn.nn.BatchNorm, called at models/segnn/segnn.py:233-235
* ``spherical_harmonics``         e3nn.o3.spherical_harmonics, called at o3_building_blocks.py:243-268
* ``O3Transform``                 models/segnn/o3_building_blocks.py:225-278
* ``fully_connected_edge_index``  utils/build_fully_connected_graph.py:4-40
* ``SEGNNLayer`` / ``SEGNN``      models/segnn/segnn.py:14-304
* ``rollout``                     helper_scripts/infer_self_feed.py:99-211 (segnn branch)
* ``target_common_loss``          training/losses.py:22-45
"""

from __future__ import annotations

import math
import re
from types import SimpleNamespace
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

# ----------------------------------------------------------------------------------------------
# Irreps (e3nn.o3.Irreps subset: only what the SEGNN call sites need)
# ----------------------------------------------------------------------------------------------


class Irreps:
    """Ordered list of (mul, l, parity) blocks; feature layout is mul-major inside a block."""

    def __init__(self, spec=None):
        if spec is None:
            self.blocks: List[Tuple[int, int, int]] = []
        elif isinstance(spec, Irreps):
            self.blocks = list(spec.blocks)
        elif isinstance(spec, (list, tuple)):
            self.blocks = [(int(m), int(l), int(p)) for (m, l, p) in spec]
        else:
            self.blocks = []
            text = str(spec).strip()
            if text:
                for tok in text.split("+"):
                    tok = tok.strip()
                    m = re.fullmatch(r"(?:(\d+)x)?(\d+)([eo])", tok)
                    if m is None:
                        raise ValueError(f"cannot parse irrep {tok!r}")
                    mul = int(m.group(1)) if m.group(1) else 1
                    self.blocks.append((mul, int(m.group(2)), 1 if m.group(3) == "e" else -1))

    @staticmethod
    def spherical_harmonics(lmax: int) -> "Irreps":
        return Irreps([(1, l, (-1) ** l) for l in range(lmax + 1)])

    @property
    def dim(self) -> int:
        return sum(m * (2 * l + 1) for m, l, _ in self.blocks)

    @property
    def num_irreps(self) -> int:
        return sum(m for m, _, _ in self.blocks)

    @property
    def lmax(self) -> int:
        return max(l for _, l, _ in self.blocks)

    def slices(self) -> List[slice]:
        out, start = [], 0
        for m, l, _ in self.blocks:
            out.append(slice(start, start + m * (2 * l + 1)))
            start += m * (2 * l + 1)
        return out

    def simplify(self) -> "Irreps":
        """Merge ADJACENT equal irreps only (e3nn semantics; segnn.py:209 relies on it)."""
        out: List[Tuple[int, int, int]] = []
        for m, l, p in self.blocks:
            if out and out[-1][1] == l and out[-1][2] == p:
                out[-1] = (out[-1][0] + m, l, p)
            elif m > 0:
                out.append((m, l, p))
        return Irreps(out)

    def sort(self) -> "Irreps":
        return Irreps(sorted(self.blocks, key=lambda b: (b[1], -b[2])))

    def __add__(self, other) -> "Irreps":
        return Irreps(self.blocks + Irreps(other).blocks)

    def __mul__(self, k: int) -> "Irreps":
        return Irreps(self.blocks * int(k))

    __rmul__ = __mul__

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            return Irreps(self.blocks[idx])
        return self.blocks[idx]

    def __len__(self):
        return len(self.blocks)

    def __iter__(self):
        return iter(self.blocks)

    def __eq__(self, other):
        return self.blocks == Irreps(other).blocks

    def __repr__(self):
        return "+".join(f"{m}x{l}{'e' if p == 1 else 'o'}" for m, l, p in self.blocks)


# ----------------------------------------------------------------------------------------------
# Wigner 3j in e3nn's real basis (e3nn/o3/_wigner.py: _su2_clebsch_gordan, change_basis_real_to_complex)
# ----------------------------------------------------------------------------------------------


def _su2_cg_coeff(j1, m1, j2, m2, j3, m3) -> float:
    if m3 != m1 + m2:
        return 0.0
    f = math.factorial
    vmin = int(max(-j1 + j2 + m3, -j1 + m1, 0))
    vmax = int(min(j2 + j3 + m1, j3 - j1 + j2, j3 + m3))
    c = math.sqrt(
        (2.0 * j3 + 1.0)
        * f(j3 + j1 - j2) * f(j3 - j1 + j2) * f(j1 + j2 - j3) / f(j1 + j2 + j3 + 1)
        * f(j3 + m3) * f(j3 - m3)
        / (f(j1 - m1) * f(j1 + m1) * f(j2 - m2) * f(j2 + m2))
    )
    s = 0.0
    for v in range(vmin, vmax + 1):
        s += (-1.0) ** (v + j2 + m2) / f(v) * f(j2 + j3 + m1 - v) * f(j1 - m1 + v) \
            / f(j3 - j1 + j2 - v) / f(j3 + m3 - v) / f(v + j1 - j2 - m3)
    return c * s


def _real_to_complex(l: int) -> torch.Tensor:
    q = torch.zeros((2 * l + 1, 2 * l + 1), dtype=torch.complex128)
    for m in range(-l, 0):
        q[l + m, l + abs(m)] = 1 / math.sqrt(2)
        q[l + m, l - abs(m)] = -1j / math.sqrt(2)
    q[l, l] = 1
    for m in range(1, l + 1):
        q[l + m, l + abs(m)] = (-1) ** m / math.sqrt(2)
        q[l + m, l - abs(m)] = 1j * (-1) ** m / math.sqrt(2)
    return (-1j) ** l * q


_W3J_CACHE = {}


def wigner_3j(l1: int, l2: int, l3: int) -> torch.Tensor:
    """Real-basis Wigner 3j with unit Frobenius norm, shape [2l1+1, 2l2+1, 2l3+1], float64."""
    key = (l1, l2, l3)
    if key not in _W3J_CACHE:
        assert abs(l1 - l2) <= l3 <= l1 + l2
        c = torch.zeros((2 * l1 + 1, 2 * l2 + 1, 2 * l3 + 1), dtype=torch.complex128)
        for m1 in range(-l1, l1 + 1):
            for m2 in range(-l2, l2 + 1):
                if abs(m1 + m2) <= l3:
                    c[l1 + m1, l2 + m2, l3 + m1 + m2] = _su2_cg_coeff(l1, m1, l2, m2, l3, m1 + m2)
        q1, q2, q3 = _real_to_complex(l1), _real_to_complex(l2), _real_to_complex(l3)
        c = torch.einsum("ij,kl,mn,ikn->jlm", q1, q2, torch.conj(q3.T), c)
        assert float(c.imag.abs().max()) < 1e-9
        c = c.real.clone()
        _W3J_CACHE[key] = c / c.norm()
    return _W3J_CACHE[key]


# ----------------------------------------------------------------------------------------------
# Spherical harmonics, l <= 2 (e3nn polynomial basis; l=1 components are (x, y, z))
# ----------------------------------------------------------------------------------------------


def spherical_harmonics(lmax: int, vec: torch.Tensor, normalize: bool = True,
                        normalization: str = "integral") -> torch.Tensor:
    """[..., 3] -> [..., (lmax+1)^2] for lmax <= 2; 'integral' normalisation as at o3_building_blocks.py:243."""
    assert lmax <= 2 and normalization == "integral"
    if normalize:
        vec = vec / vec.norm(dim=-1, keepdim=True).clamp_min(1e-12)  # torch.nn.functional.normalize
    x, y, z = vec[..., 0], vec[..., 1], vec[..., 2]
    out = [torch.ones_like(x)]
    if lmax >= 1:
        s3 = math.sqrt(3.0)
        out += [s3 * x, s3 * y, s3 * z]
    if lmax >= 2:
        s5, s15 = math.sqrt(5.0), math.sqrt(15.0)
        out += [s15 * x * z, s15 * x * y, s5 * (y * y - 0.5 * (x * x + z * z)),
                s15 * y * z, 0.5 * s15 * (z * z - x * x)]
    return torch.stack(out, dim=-1) / math.sqrt(4.0 * math.pi)


# ----------------------------------------------------------------------------------------------
# FullyConnectedTensorProduct (uvw, shared weights, 'component' irrep norm, 'element' path norm)
# ----------------------------------------------------------------------------------------------


class FullyConnectedTensorProduct(nn.Module):
    """e3nn.o3.FullyConnectedTensorProduct restated. ``weight`` is the flat e3nn parameter:
    instruction views [mul1, mul2, mul_out] row-major, concatenated in instruction order."""

    def __init__(self, irreps_in1, irreps_in2, irreps_out, dtype=torch.float64):
        super().__init__()
        self.irreps_in1, self.irreps_in2, self.irreps_out = Irreps(irreps_in1), Irreps(irreps_in2), Irreps(irreps_out)
        self.instructions: List[Tuple[int, int, int]] = []
        for i1, (_, l1, p1) in enumerate(self.irreps_in1):
            for i2, (_, l2, p2) in enumerate(self.irreps_in2):
                for io, (_, lo, po) in enumerate(self.irreps_out):
                    if abs(l1 - l2) <= lo <= l1 + l2 and po == p1 * p2:
                        self.instructions.append((i1, i2, io))
        self.weight_shapes = [
            (self.irreps_in1[i1][0], self.irreps_in2[i2][0], self.irreps_out[io][0])
            for i1, i2, io in self.instructions
        ]
        self.weight_numel = sum(a * b * c for a, b, c in self.weight_shapes)
        # path weight = sqrt(dim(ir_out) / sum over instructions into the same output slice of mul1*mul2)
        fan = {}
        for (i1, i2, io), (m1, m2, _) in zip(self.instructions, self.weight_shapes):
            fan[io] = fan.get(io, 0) + m1 * m2
        self.path_weights = [
            math.sqrt((2 * self.irreps_out[io][1] + 1) / fan[io]) for (_, _, io) in self.instructions
        ]
        self.weight = nn.Parameter(torch.randn(self.weight_numel, dtype=dtype))

    def weight_views(self) -> List[torch.Tensor]:
        out, off = [], 0
        for shp in self.weight_shapes:
            k = shp[0] * shp[1] * shp[2]
            out.append(self.weight[off:off + k].view(shp))
            off += k
        return out

    def forward(self, x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
        s1, s2, so = self.irreps_in1.slices(), self.irreps_in2.slices(), self.irreps_out.slices()
        rows = x1.shape[0]
        out = x1.new_zeros(rows, self.irreps_out.dim)
        for (i1, i2, io), w, pw in zip(self.instructions, self.weight_views(), self.path_weights):
            m1, l1, _ = self.irreps_in1[i1]
            m2, l2, _ = self.irreps_in2[i2]
            mo, lo, _ = self.irreps_out[io]
            a = x1[:, s1[i1]].reshape(rows, m1, 2 * l1 + 1)
            b = x2[:, s2[i2]].reshape(rows, m2, 2 * l2 + 1)
            c3 = wigner_3j(l1, l2, lo).to(x1.dtype)
            # couple with the Wigner 3j first (cheap: mul2 = 1 everywhere on this path), then mix multiplicities
            t = torch.einsum("zui,zvj,ijk->zuvk", a, b, c3)
            y = torch.einsum("zuvk,uvw->zwk", t, w.to(x1.dtype))
            out[:, so[io]] = out[:, so[io]] + pw * y.reshape(rows, mo * (2 * lo + 1))
        return out


# ----------------------------------------------------------------------------------------------
# O3TensorProduct / Gate / O3TensorProductSwishGate
# ----------------------------------------------------------------------------------------------


class O3TensorProduct(nn.Module):
    """models/segnn/o3_building_blocks.py:10-167 (tp_rescale=True path)."""

    def __init__(self, irreps_in1, irreps_out, irreps_in2=None, dtype=torch.float64):
        super().__init__()
        self.irreps_in1, self.irreps_out = Irreps(irreps_in1), Irreps(irreps_out)
        self.irreps_in2 = Irreps("1x0e") if irreps_in2 is None else Irreps(irreps_in2)
        self.irreps_in2_provided = irreps_in2 is not None
        self.tp = FullyConnectedTensorProduct(self.irreps_in1, self.irreps_in2, self.irreps_out, dtype=dtype)
        fan = {}
        for (_, _, io), (m1, m2, _) in zip(self.tp.instructions, self.tp.weight_shapes):
            fan[io] = fan.get(io, 0) + m1 * m2
        self.fan_in = fan
        out_slices = self.irreps_out.slices()
        sqrt_k = torch.zeros(self.irreps_out.dim, dtype=dtype)
        with torch.no_grad():
            for (_, _, io), w in zip(self.tp.instructions, self.tp.weight_views()):
                w.uniform_(-1.0 / math.sqrt(fan[io]), 1.0 / math.sqrt(fan[io]))
                sqrt_k[out_slices[io]] = 1.0 / math.sqrt(fan[io])
        bias_idx, biases = [], []
        for io, (mul, l, _) in enumerate(self.irreps_out):
            if l == 0:
                bias_idx.append(torch.arange(out_slices[io].start, out_slices[io].stop))
                k = 1.0 / math.sqrt(fan[io])
                biases.append(torch.empty(mul, dtype=dtype).uniform_(-k, k))
        if biases:
            self.biases = nn.Parameter(torch.cat(biases))
            self.register_buffer("bias_idx", torch.cat(bias_idx), persistent=False)
        else:
            self.biases = None
        self.register_buffer("sqrt_k_correction", sqrt_k, persistent=False)

    def forward_tp_rescale_bias(self, x1, x2=None):
        if x2 is None:
            x2 = torch.ones_like(x1[:, 0:1])
        out = self.tp(x1, x2) / self.sqrt_k_correction.to(x1.dtype)
        if self.biases is not None:
            out[:, self.bias_idx] = out[:, self.bias_idx] + self.biases.to(x1.dtype)
        return out

    def forward(self, x1, x2=None):
        return self.forward_tp_rescale_bias(x1, x2)


_NORM2MOM_CACHE = {}


def normalize2mom_const(name: str) -> float:
    """e3nn.math.normalize2mom constant: (E f(z)^2)^-1/2 over 1e6 float64 normal samples, seed 0."""
    if name not in _NORM2MOM_CACHE:
        gen = torch.Generator(device="cpu").manual_seed(0)
        z = torch.randn(1_000_000, generator=gen, dtype=torch.float64)
        f = torch.nn.functional.silu if name == "silu" else torch.sigmoid
        _NORM2MOM_CACHE[name] = float(f(z).pow(2).mean().pow(-0.5))
    return _NORM2MOM_CACHE[name]


class Gate(nn.Module):
    """e3nn.nn.Gate(scalars, [SiLU], gates, [sigmoid], gated): input [scalars | gates | gated]."""

    def __init__(self, n_scalars: int, irreps_gated: Irreps):
        super().__init__()
        self.n_scalars, self.irreps_gated = n_scalars, Irreps(irreps_gated)
        self.n_gates = self.irreps_gated.num_irreps
        self.c_silu, self.c_sig = normalize2mom_const("silu"), normalize2mom_const("sigmoid")

    def forward(self, x):
        s = x[:, : self.n_scalars]
        g = x[:, self.n_scalars: self.n_scalars + self.n_gates]
        v = x[:, self.n_scalars + self.n_gates:]
        s = self.c_silu * torch.nn.functional.silu(s)
        g = self.c_sig * torch.sigmoid(g)
        outs, off_v, off_g = [s], 0, 0
        for mul, l, _ in self.irreps_gated:
            d = 2 * l + 1
            blk = v[:, off_v: off_v + mul * d].reshape(-1, mul, d)
            outs.append((blk * g[:, off_g: off_g + mul, None]).reshape(-1, mul * d))
            off_v += mul * d
            off_g += mul
        return torch.cat(outs, dim=-1)


class O3TensorProductSwishGate(O3TensorProduct):
    """models/segnn/o3_building_blocks.py:170-203."""

    def __init__(self, irreps_in1, irreps_out, irreps_in2=None, dtype=torch.float64):
        irreps_out = Irreps(irreps_out)
        scalars = Irreps([irreps_out[0]])
        gated = irreps_out[1:]
        gates = Irreps([(gated.num_irreps, 0, 1)])
        irreps_g = (scalars + gates + gated).simplify()
        super().__init__(irreps_in1, irreps_g, irreps_in2, dtype=dtype)
        self.irreps_final = irreps_out
        self.gate = Gate(scalars.num_irreps, gated) if gated.num_irreps > 0 else None

    def forward(self, x1, x2=None):
        out = self.forward_tp_rescale_bias(x1, x2)
        if self.gate is None:
            return torch.nn.functional.silu(out)
        return self.gate(out)


# ----------------------------------------------------------------------------------------------
# e3nn BatchNorm (reduce='mean', normalization='component', affine, not instance)
# ----------------------------------------------------------------------------------------------


class BatchNorm(nn.Module):
    def __init__(self, irreps, eps: float = 1e-5, momentum: float = 0.1, dtype=torch.float64):
        super().__init__()
        self.irreps, self.eps, self.momentum = Irreps(irreps), eps, momentum
        n_scalar = sum(m for m, l, p in self.irreps if l == 0 and p == 1)
        n_feat = self.irreps.num_irreps
        self.register_buffer("running_mean", torch.zeros(n_scalar, dtype=dtype))
        self.register_buffer("running_var", torch.ones(n_feat, dtype=dtype))
        self.weight = nn.Parameter(torch.ones(n_feat, dtype=dtype))
        self.bias = nn.Parameter(torch.zeros(n_scalar, dtype=dtype))

    def forward(self, x):
        rows = x.shape[0]
        fields, ix, irm, irv, ib = [], 0, 0, 0, 0
        new_means, new_vars = [], []
        for mul, l, p in self.irreps:
            d = 2 * l + 1
            f = x[:, ix: ix + mul * d].reshape(rows, mul, d)
            ix += mul * d
            scalar = l == 0 and p == 1
            if scalar:
                if self.training:
                    mean = f.mean(dim=(0,)).reshape(mul)
                    new_means.append((1 - self.momentum) * self.running_mean[irm: irm + mul]
                                     + self.momentum * mean.detach())
                else:
                    mean = self.running_mean[irm: irm + mul]
                irm += mul
                f = f - mean.reshape(1, mul, 1)
            if self.training:
                norm = f.pow(2).mean(dim=2).mean(dim=0)
                new_vars.append((1 - self.momentum) * self.running_var[irv: irv + mul]
                                + self.momentum * norm.detach())
            else:
                norm = self.running_var[irv: irv + mul]
            scale = (norm + self.eps).pow(-0.5) * self.weight[irv: irv + mul]
            irv += mul
            f = f * scale.reshape(1, mul, 1)
            if scalar:
                f = f + self.bias[ib: ib + mul].reshape(1, mul, 1)
                ib += mul
            fields.append(f.reshape(rows, mul * d))
        if self.training:
            with torch.no_grad():
                if new_means:
                    self.running_mean.copy_(torch.cat(new_means))
                self.running_var.copy_(torch.cat(new_vars))
        return torch.cat(fields, dim=-1)


# ----------------------------------------------------------------------------------------------
# Hidden irreps sizing, graph, O3Transform
# ----------------------------------------------------------------------------------------------


def weight_balanced_irreps(hidden_features: int, irreps_attr: Irreps, lmax: Optional[int] = None) -> Irreps:
    """models/balanced_irreps.py:51-85 with sh=True: smallest n with numel(h x attr -> h) >= H^2."""
    irreps_attr = Irreps(irreps_attr)
    lmax = irreps_attr.lmax if lmax is None else lmax
    target = hidden_features * hidden_features

    def numel(n):
        h = Irreps([(n, l, (-1) ** l) for l in range(lmax + 1)])
        cnt = 0
        for (m1, l1, p1) in h:
            for (m2, l2, p2) in irreps_attr:
                for (mo, lo, po) in h:
                    if abs(l1 - l2) <= lo <= l1 + l2 and po == p1 * p2:
                        cnt += m1 * m2 * mo
        return cnt

    n = 1
    while numel(n) < target:
        n += 1
    return Irreps([(n, l, (-1) ** l) for l in range(lmax + 1)])


def fully_connected_edge_index(batch_size: int, num_nodes: int) -> torch.Tensor:
    """utils/build_fully_connected_graph.py:4-20: row-major nonzero(~eye(N)), graph offset g*N.
    Row 0 = source/sender, row 1 = target/receiver (PyG flow source_to_target)."""
    mask = ~torch.eye(num_nodes, dtype=torch.bool)
    row, col = torch.nonzero(mask, as_tuple=True)
    per = row.numel()
    offs = torch.arange(0, batch_size * num_nodes, num_nodes).repeat_interleave(per)
    return torch.stack([row.repeat(batch_size) + offs, col.repeat(batch_size) + offs], dim=0)


def build_graph_with_knn(loc, batch_size, num_nodes, device=None, num_neighbors=None):
    """utils/build_fully_connected_graph.py:23-40 (fully-connected fast path only)."""
    num_nodes = int(num_nodes)
    num_neighbors = num_nodes - 1 if num_neighbors is None else int(num_neighbors)
    if num_neighbors >= num_nodes:
        raise ValueError("Graph cannot have more neighbors than there are nodes in simulation - 1")
    if num_neighbors != num_nodes - 1:
        return knn_edge_index(loc, batch_size, num_nodes, num_neighbors)
    return fully_connected_edge_index(batch_size, num_nodes)


def o3_transform(graph, lmax_attr: int = 1, use_force_input: bool = False):
    """models/segnn/o3_building_blocks.py:230-278: attaches edge_attr, node_attr, x, additional_message_features."""
    pos, vel, mass = graph.pos, graph.vel, graph.mass
    src, tgt = graph.edge_index[0], graph.edge_index[1]
    prod_mass = mass[src] * mass[tgt]
    rel_pos = pos[src] - pos[tgt]
    edge_dist = rel_pos.pow(2).sum(1, keepdim=True).sqrt()
    graph.edge_attr = spherical_harmonics(lmax_attr, rel_pos)
    vel_emb = spherical_harmonics(lmax_attr, vel)
    num_nodes = pos.shape[0]
    summed = torch.zeros(num_nodes, graph.edge_attr.shape[1], dtype=pos.dtype).index_add_(0, tgt, graph.edge_attr)
    count = torch.zeros(num_nodes, dtype=pos.dtype).index_add_(0, tgt, torch.ones_like(tgt, dtype=pos.dtype))
    graph.node_attr = summed / count.clamp_min(1).unsqueeze(1) + vel_emb
    if use_force_input:  # :267-271
        graph.node_attr = graph.node_attr + spherical_harmonics(lmax_attr, graph.force)
    vel_abs = vel.pow(2).sum(1, keepdim=True).sqrt()
    mean_pos = pos.mean(1, keepdim=True)  # reference quirk: mean over xyz of each node
    graph.x = torch.cat((pos - mean_pos, vel, vel_abs), 1)
    graph.additional_message_features = torch.cat((edge_dist, prod_mass), dim=-1)
    return graph


# ----------------------------------------------------------------------------------------------
# SEGNNLayer / SEGNN
# ----------------------------------------------------------------------------------------------


class SEGNNLayer(nn.Module):
    """models/segnn/segnn.py:192-304 with PyG propagate written out (gather, message, scatter-add, update)."""

    def __init__(self, input_irreps, hidden_irreps, output_irreps, edge_attr_irreps, node_attr_irreps,
                 norm="batch", additional_message_irreps=None, dtype=torch.float64):
        super().__init__()
        input_irreps, hidden_irreps = Irreps(input_irreps), Irreps(hidden_irreps)
        msg_in = (2 * input_irreps + Irreps(additional_message_irreps)).simplify()
        upd_in = (input_irreps + hidden_irreps).simplify()
        self.message_layer_1 = O3TensorProductSwishGate(msg_in, hidden_irreps, edge_attr_irreps, dtype=dtype)
        self.message_layer_2 = O3TensorProductSwishGate(hidden_irreps, hidden_irreps, edge_attr_irreps, dtype=dtype)
        self.update_layer_1 = O3TensorProductSwishGate(upd_in, hidden_irreps, node_attr_irreps, dtype=dtype)
        self.update_layer_2 = O3TensorProduct(hidden_irreps, hidden_irreps, node_attr_irreps, dtype=dtype)
        self.norm = norm
        self.feature_norm = BatchNorm(hidden_irreps, dtype=dtype) if norm == "batch" else None
        self.message_norm = BatchNorm(hidden_irreps, dtype=dtype) if norm == "batch" else None
        if norm == "instance":  # segnn.py:236-237: per-graph normalisation of the node features only
            self.feature_norm = InstanceNorm(hidden_irreps, dtype=dtype)

    def message(self, x_i, x_j, edge_attr, add):
        inp = torch.cat((x_i, x_j) if add is None else (x_i, x_j, add), dim=-1)
        m = self.message_layer_1(inp, edge_attr)
        m = self.message_layer_2(m, edge_attr)
        if self.message_norm is not None:
            m = self.message_norm(m)
        return m

    def forward(self, x, edge_index, edge_attr, node_attr, batch=None, additional_message_features=None):
        src, tgt = edge_index[0], edge_index[1]
        m = self.message(x[tgt], x[src], edge_attr, additional_message_features)
        agg = torch.zeros_like(x).index_add_(0, tgt, m)
        upd = self.update_layer_1(torch.cat((x, agg), dim=-1), node_attr)
        upd = self.update_layer_2(upd, node_attr)
        x = x + upd
        if self.feature_norm is not None:
            x = self.feature_norm(x, batch) if self.norm == "instance" else self.feature_norm(x)  # segnn.py:257-261
        return x


class SEGNN(nn.Module):
    """models/segnn/segnn.py:14-189, task='node'. state_dict keys match the reference's parameter names."""

    def __init__(self, input_irreps="2x1o+1x0e", hidden_features=64, lmax_h=1, lmax_attr=1, num_layers=4,
                 output_irreps="2x1o", norm="batch", pool="avg", task="node",
                 additional_message_irreps="2x0e", training_args=None, dtype=torch.float64):
        super().__init__()
        assert task == "node"
        self.hidden_features, self.lmax_h, self.lmax_attr, self.num_layers = hidden_features, lmax_h, lmax_attr, num_layers
        attr = Irreps.spherical_harmonics(lmax_attr)
        self.node_attr_irreps = self.edge_attr_irreps = attr
        self.hidden_irreps = weight_balanced_irreps(hidden_features, attr, lmax=lmax_h)
        h = self.hidden_irreps
        self.embedding_layer = O3TensorProduct(input_irreps, h, attr, dtype=dtype)
        self.layers = nn.ModuleList([
            SEGNNLayer(h, h, h, attr, attr, norm=norm, additional_message_irreps=additional_message_irreps, dtype=dtype)
            for _ in range(num_layers)
        ])
        self.pre_pool1 = O3TensorProductSwishGate(h, h, attr, dtype=dtype)
        self.pre_pool2 = O3TensorProduct(h, output_irreps, attr, dtype=dtype)

    def get_model_size(self):
        return self.hidden_features

    def forward(self, graph, return_layers: bool = False):
        graph.node_attr = graph.node_attr.clone()
        graph.node_attr[:, 0] = 1.0  # segnn.py:148
        x = self.embedding_layer(graph.x, graph.node_attr)
        per_layer = [x]
        add = getattr(graph, "additional_message_features", None)
        for layer in self.layers:
            x = layer(x, graph.edge_index, graph.edge_attr, graph.node_attr, getattr(graph, "batch", None), add)
            per_layer.append(x)
        x = self.pre_pool1(x, graph.node_attr)
        x = self.pre_pool2(x, graph.node_attr)
        if return_layers:
            return x, per_layer
        return x


# ----------------------------------------------------------------------------------------------
# Whole-step helpers: graph build + transform + forward, rollout, loss
# ----------------------------------------------------------------------------------------------


def make_graph(pos, vel, mass, batch_size: int, num_nodes: int, lmax_attr: int = 1, num_neighbors=None):
    """dataloaders/segnn_n_body_dataloader.py:9-33 without PyG: attribute bag with the same field names."""
    g = SimpleNamespace(pos=pos, vel=vel, mass=mass.reshape(-1, 1), force=torch.zeros_like(pos))
    g.batch = torch.arange(batch_size).repeat_interleave(num_nodes)
    g.edge_index = build_graph_with_knn(pos, batch_size, num_nodes, None,
                                        num_nodes - 1 if num_neighbors is None else num_neighbors)
    return o3_transform(g, lmax_attr)


@torch.no_grad()
def rollout(model: SEGNN, pos0, vel0, mass, steps: int, target: str = "pos_dt+vel"):
    """helper_scripts/infer_self_feed.py:99-211, segnn branch. pos0/vel0 [B,N,3], mass [B,N,1].
    Returns loc [B,steps+1,N,3], vel [B,steps+1,N,3]."""
    b, n, _ = pos0.shape
    locs, vels = [pos0], [vel0]
    for _ in range(steps):
        g = make_graph(locs[-1].reshape(b * n, 3), vels[-1].reshape(b * n, 3), mass.reshape(b * n, 1), b, n,
                       model.lmax_attr)
        pred = model(g)
        p_loc, p_vel = pred[:, :3].reshape(b, n, 3), pred[:, 3:].reshape(b, n, 3)
        if target == "pos_dt+vel":
            p_loc = locs[-1] + p_loc
        locs.append(p_loc)
        vels.append(p_vel)
    return torch.stack(locs, dim=1), torch.stack(vels, dim=1)


def target_common_loss(pred, y):
    """training/losses.py:22-45 for target 'pos_dt+vel' with unit weights."""
    mse = torch.nn.functional.mse_loss
    return mse(pred[..., 0:3], y[..., 0:3]) + mse(pred[..., 3:6], y[..., 3:6])


def synthetic_system(batch_size: int, num_nodes: int, seed: int = 0, charged: bool = True, dtype=torch.float64):
    """SURVEY 8(d) synthetic inputs (density rule of synthetic_sim.py:375-381)."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    pos = torch.randn(batch_size, num_nodes, 3, generator=gen, dtype=torch.float64) * (num_nodes / 5.0) ** (1.0 / 3.0)
    vel = torch.randn(batch_size, num_nodes, 3, generator=gen, dtype=torch.float64)
    vel = vel - vel.mean(dim=1, keepdim=True)
    if charged:
        mass = (torch.randint(0, 2, (batch_size, num_nodes, 1), generator=gen).to(torch.float64) * 2.0 - 1.0)
    else:
        mass = torch.ones(batch_size, num_nodes, 1, dtype=torch.float64)
    return pos.to(dtype), vel.to(dtype), mass.to(dtype)


def perturb_bn_buffers(model: nn.Module, seed: int = 1):
    """Make eval-mode BN non-trivial: running_mean ~ 0.1 randn, running_var ~ U(0.5,1.5),
    and non-trivial affine weight/bias."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    with torch.no_grad():
        for mod in model.modules():
            if hasattr(mod, "running_mean") and hasattr(mod, "running_var"):
                mod.running_mean.copy_(0.1 * torch.randn(mod.running_mean.shape, generator=gen, dtype=torch.float64))
                mod.running_var.copy_(0.5 + torch.rand(mod.running_var.shape, generator=gen, dtype=torch.float64))
                mod.weight.copy_(0.75 + 0.5 * torch.rand(mod.weight.shape, generator=gen, dtype=torch.float64))
                mod.bias.copy_(0.1 * torch.randn(mod.bias.shape, generator=gen, dtype=torch.float64))


# ----------------------------------------------------------------------------------------------
# Rollout macros (acceptance statistics) -- restated with the reference's own loops / libraries
# ----------------------------------------------------------------------------------------------
def nbody_energies(loc, vel, G: float, softening: float):
    """trainer.py:888-927 `_compute_nbody_energies`. loc, vel [B, T, N, 3] -> per-simulation arrays [B, T] and the
    batch-averaged series dict (unit masses)."""
    import numpy as np
    loc, vel = np.asarray(loc, dtype=np.float64), np.asarray(vel, dtype=np.float64)
    batch, steps, n, _ = loc.shape
    kinetic, potential = np.zeros((batch, steps)), np.zeros((batch, steps))
    iu = np.triu_indices(n, 1)
    for b in range(batch):
        L, V = loc[b], vel[b]
        kinetic[b] = 0.5 * np.sum(V * V, axis=(1, 2))
        d = L[:, None, :, :] - L[:, :, None, :]
        inv_r = np.sqrt((d * d).sum(-1) + softening * softening)
        inv_r[inv_r > 0] = 1.0 / inv_r[inv_r > 0]
        potential[b] = -G * np.sum(inv_r[:, iu[0], iu[1]], axis=1)
    series = {"potential": potential.mean(0), "kinetic": kinetic.mean(0)}
    series["total"] = series["potential"] + series["kinetic"]
    return kinetic, potential, series


def momentum_magnitude(vel):
    """datasets/nbody/visualization_utils.py:959-960: |sum_i v_i| per (simulation, frame). vel [B, T, N, 3]."""
    import numpy as np
    v = np.asarray(vel, dtype=np.float64).sum(axis=2)
    return np.sqrt((v ** 2).sum(axis=-1))


def ks_p(a, b) -> float:
    """utils/ks_utils.py:7-19."""
    import numpy as np
    from scipy import stats
    a, b = np.asarray(a).ravel(), np.asarray(b).ravel()
    if a.size == 0 or b.size == 0 or np.all(np.isnan(a)) or np.all(np.isnan(b)):
        return float("nan")
    a, b = a[~np.isnan(a)], b[~np.isnan(b)]
    if a.size == 0 or b.size == 0:
        return float("nan")
    return float(stats.ks_2samp(a, b)[1])


def combine_pvalues_fisher(p_values) -> float:
    """utils/ks_utils.py:22-29 (mpmath 200-digit sum + scipy chi2.sf, floored at 1e-300)."""
    from mpmath import log, mp
    from scipy.stats import chi2
    vals = [p for p in p_values if p == p and p > 0.0]
    if not vals:
        return float("nan")
    mp.dps = 200
    chi_stat = float(-2 * mp.fsum([log(mp.mpf(p)) for p in vals]))
    return float(max(chi2.sf(chi_stat, 2 * len(vals)), 1e-300))


def event_counters(loc, vel, time_threshold=3, contact_distance=0.5, leave_distance=15.0, turn_angle=30.0):
    """datasets/nbody/visualization_utils.py:1093-1124 (stickings / collisions), :1145-1167 (bodies leaving),
    :1170-1187 (max centre-of-mass drift), :1201-1222 (sharp turns), with the reference's own loop structure.
    loc, vel [B, T, N, 3] -> dict of per-simulation arrays."""
    import numpy as np
    loc, vel = np.asarray(loc), np.asarray(vel)
    sims, steps, n = loc.shape[:3]
    stick, coll, left, turns, drift = (np.zeros(sims) for _ in range(5))
    for s in range(sims):
        ongoing = np.zeros((n, n))
        outside = np.zeros(n)
        com0 = loc[s, 0].mean(axis=0)
        for t in range(1, steps):
            for i in range(n):
                for j in range(i + 1, n):
                    if np.linalg.norm(loc[s, t, i] - loc[s, t, j]) <= contact_distance:
                        ongoing[i, j] += 1
                        if ongoing[i, j] == 1:
                            coll[s] += 1
                        if ongoing[i, j] == time_threshold:
                            stick[s] += 1
                            coll[s] -= 1
                    else:
                        ongoing[i, j] = 0
            com = loc[s, t].mean(axis=0)
            drift[s] = max(drift[s], np.sqrt(((com - com0) ** 2).sum()))
            for b in range(n):
                if np.sqrt(((loc[s, t, b] - com) ** 2).sum()) > leave_distance:
                    outside[b] += 1
                else:
                    outside[b] = 0
                a, c = vel[s, t, b], vel[s, t - 1, b]
                with np.errstate(invalid="ignore", divide="ignore"):
                    cosang = np.clip(np.dot(a, c) / (np.linalg.norm(a) * np.linalg.norm(c)), -1, 1)
                    if np.degrees(np.arccos(cosang)) > turn_angle:
                        turns[s] += 1
        left[s] = len([i for i in outside if i > 10])
    return {"stickings": stick, "collisions": coll, "bodies_left": left, "sharp_turns": turns,
            "max_com_distance": drift}


def gravity_trajectory(pos, vel, mass, G, softening, dt, T, sample_freq):
    """datasets/nbody/dataset/synthetic_sim.py:319-358,360-418 (GravitySim) for given initial conditions, NumPy
    float64, one simulation: pos, vel [N,3], mass [N,1] -> loc, vel, force [T / sample_freq, N, 3]."""
    import numpy as np

    def acceleration(p):
        x, y, z = p[:, 0:1], p[:, 1:2], p[:, 2:3]
        dx, dy, dz = x.T - x, y.T - y, z.T - z
        inv_r3 = dx ** 2 + dy ** 2 + dz ** 2 + softening ** 2
        inv_r3[inv_r3 > 0] = inv_r3[inv_r3 > 0] ** (-1.5)
        return np.hstack((G * (dx * inv_r3) @ mass, G * (dy * inv_r3) @ mass, G * (dz * inv_r3) @ mass))

    pos, vel, mass = np.array(pos, dtype=np.float64), np.array(vel, dtype=np.float64), np.array(mass, dtype=np.float64)
    frames = T // sample_freq
    ps, vs, fs = (np.zeros((frames,) + pos.shape) for _ in range(3))
    acc = acceleration(pos)
    k = 0
    for i in range(T):
        if i % sample_freq == 0:
            ps[k], vs[k], fs[k] = pos, vel, acc * mass
            k += 1
        vel = vel + acc * dt / 2.0
        pos = pos + vel * dt
        acc = acceleration(pos)
        vel = vel + acc * dt / 2.0
    return ps, vs, fs


def charged_trajectory(x0, v0, charges, interaction_strength, delta_t, steps, max_force=None):
    """datasets/nbody_offline/datagen/system.py:78-123 (``System.compute_F`` / ``simulate_one_step``) with isolated
    bodies only (physical_objects.py:49-57: v += F dt, x += v dt). x0, v0 [N,3], charges [N,1] -> X, V [steps, N, 3]."""
    import numpy as np
    x, v = np.array(x0, dtype=np.float64), np.array(v0, dtype=np.float64)
    q = np.array(charges, dtype=np.float64).reshape(-1, 1)
    edges = q @ q.T
    max_force = 0.1 / delta_t if max_force is None else max_force
    xs, vs = [], []
    for _ in range(steps):
        with np.errstate(divide="ignore", invalid="ignore"):
            sq = (x ** 2).sum(1)
            dist2 = sq[:, None] + sq[None, :] - 2 * x @ x.T
            size = interaction_strength * edges / np.power(dist2, 1.5)
            np.fill_diagonal(size, 0)
            f = (size[:, :, None] * (x[:, None, :] - x[None, :, :])).sum(axis=1)
        f = np.clip(f, -max_force, max_force)
        v = v + f * delta_t
        x = x + v * delta_t
        xs.append(x.copy())
        vs.append(v.copy())
    return np.stack(xs), np.stack(vs)


def group_collision_counts(loc, time_threshold=2, distance_threshold=2.0):
    """datasets/nbody/visualization_utils.py:1455-1610 (the counting part of
    ``plot_group_collision_distribution_multiplot``): a pair and a disjoint triplet of bodies, each 'stuck' (all mutual
    distances <= threshold for >= time_threshold consecutive steps), count one group collision per (pair interval,
    triplet interval) with overlapping lifetimes if any body of the pair comes within the threshold of any body of the
    triplet at some step t >= the start of the overlap.  loc [S, T, N, 3] -> counts [S]."""
    import numpy as np
    from itertools import combinations
    loc = np.asarray(loc)
    sims, steps, n = loc.shape[:3]
    out = np.zeros(sims)
    for s in range(sims):
        d = np.linalg.norm(loc[s, :, :, None, :] - loc[s, :, None, :, :], axis=-1)  # [T, N, N]
        close = d <= distance_threshold

        def intervals(flag):  # maximal runs of True with length >= time_threshold -> [start, end]
            res, run = [], 0
            for t in range(steps):
                if flag[t]:
                    run += 1
                    if run == time_threshold:
                        res.append([t - time_threshold + 1, None])
                else:
                    if run >= time_threshold:
                        res[-1][1] = t - 1
                    run = 0
            if res and res[-1][1] is None:
                res[-1][1] = steps - 1
            return res

        pairs = {(i, j): intervals(close[:, i, j]) for i in range(n) for j in range(i + 1, n)}
        trips = {(i, j, k): intervals(close[:, i, j] & close[:, i, k] & close[:, j, k])
                 for i, j, k in combinations(range(n), 3)}
        count = 0
        for pair, p_int in pairs.items():
            for trip, t_int in trips.items():
                if set(pair).isdisjoint(trip) and p_int and t_int:
                    touch = np.zeros(steps, dtype=bool)
                    for i in pair:
                        for j in trip:
                            touch |= close[:, i, j]
                    for ps, pe in p_int:
                        for ts, te in t_int:
                            lo = max(ps, ts)
                            if lo <= min(pe, te) and touch[lo:].any():
                                count += 1
        out[s] = count
    return out


def noam_rate(step: int, model_size: int, factor: float, warmup: int) -> float:
    """trainer.py:189-195 ``Trainer._rate``."""
    step = max(int(step), 1)
    return factor * (model_size ** (-0.5) * min(step ** (-0.5), step * warmup ** (-1.5)))


class InstanceNorm(nn.Module):
    """models/segnn/instance_norm.py:8-129 (reduce='mean', normalization='component', affine)."""

    def __init__(self, irreps, eps: float = 1e-5, dtype=torch.float64):
        super().__init__()
        self.irreps, self.eps = Irreps(irreps), eps
        self.weight = nn.Parameter(torch.ones(self.irreps.num_irreps, dtype=dtype))
        self.bias = nn.Parameter(torch.zeros(sum(m for m, l, _ in self.irreps if l == 0), dtype=dtype))

    def forward(self, x, batch):
        graphs = int(batch.max()) + 1
        count = torch.zeros(graphs, dtype=x.dtype).index_add_(0, batch, torch.ones_like(batch, dtype=x.dtype))

        def gmean(t):
            return torch.zeros((graphs,) + t.shape[1:], dtype=t.dtype).index_add_(0, batch, t) \
                / count.reshape(-1, *([1] * (t.dim() - 1)))

        outs, ix, iw, ib = [], 0, 0, 0
        for mul, l, _ in self.irreps:
            d = 2 * l + 1
            f = x[:, ix: ix + mul * d].reshape(-1, mul, d)
            ix += mul * d
            if l == 0:
                f = f - gmean(f)[batch]
            norm = gmean(f.pow(2).mean(-1))
            scale = (norm + self.eps).pow(-0.5) * self.weight[None, iw: iw + mul]
            iw += mul
            f = f * scale[batch].reshape(-1, mul, 1)
            if d == 1:
                f = f + self.bias[ib: ib + mul].reshape(mul, 1)
                ib += mul
            outs.append(f.reshape(-1, mul * d))
        return torch.cat(outs, dim=-1)


def knn_edge_index(loc, batch_size: int, num_nodes: int, num_neighbors: int) -> torch.Tensor:
    """utils/build_fully_connected_graph.py:42-80: row 0 = the node itself, row 1 = its k nearest neighbours by
    ``cdist`` + ``topk(largest=False)`` with the node itself (distance 0, first) dropped."""
    pts = loc.reshape(batch_size, num_nodes, -1)
    k = min(num_neighbors + 1, num_nodes)
    idx = torch.topk(torch.cdist(pts, pts), k=k, largest=False).indices[:, :, 1:]
    rows = torch.arange(num_nodes).view(1, -1, 1).expand(batch_size, num_nodes, num_neighbors)
    offs = torch.arange(0, batch_size * num_nodes, num_nodes).view(batch_size, 1, 1)
    return torch.stack([(rows + offs).flatten(), (idx + offs).flatten()], dim=0)
