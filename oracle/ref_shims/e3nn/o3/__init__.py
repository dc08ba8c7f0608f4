"""Stand-in for ``e3nn.o3`` (e3nn 0.5.1) -- TEST INFRASTRUCTURE, see ../../README.md.

Only the API the reference's SEGNN path touches: ``Irrep``, ``Irreps``, ``FullyConnectedTensorProduct`` (mode
``uvw``, shared weights, 'component' irrep normalisation, 'element' path normalisation), ``spherical_harmonics``
(l <= 2, the generated-polynomial form), ``wigner_3j``.  Written from e3nn's published algorithm, independently of the
classes in ``oracle/segnn_oracle.py``:

* the coupling tensors are obtained here as normalised Gaunt integrals of this file's own spherical harmonics by exact
  quadrature (every path the SEGNN models use has l1 + l2 + l3 even); ``oracle.segnn_oracle.wigner_3j`` follows e3nn's
  SU(2) Clebsch-Gordan + change-of-basis code instead.  ``tests/test_reference_golden.py`` asserts the two agree.
* the tensor product contracts ``einsum('uvw,ijk,zuvij->zwk')`` like e3nn's generated code.
"""
from __future__ import annotations

import collections
import math
from typing import List

import torch
import torch.nn as nn

__all__ = ["Irrep", "Irreps", "FullyConnectedTensorProduct", "TensorProduct", "Linear", "spherical_harmonics",
           "wigner_3j"]


class Irrep(tuple):
    def __new__(cls, l, p=None):
        if p is None:
            if isinstance(l, Irrep):
                return l
            if isinstance(l, str):
                name = l.strip()
                l, p = int(name[:-1]), {"e": 1, "o": -1, "y": None}[name[-1]]
                if p is None:
                    p = (-1) ** l
            elif isinstance(l, tuple):
                l, p = l
        assert isinstance(l, int) and l >= 0 and p in (-1, 1), (l, p)
        return super().__new__(cls, (l, p))

    @property
    def l(self) -> int:  # noqa: E743
        return self[0]

    @property
    def p(self) -> int:
        return self[1]

    @property
    def dim(self) -> int:
        return 2 * self.l + 1

    def __repr__(self):
        return f"{self.l}{'e' if self.p == 1 else 'o'}"

    def __mul__(self, other):
        other = Irrep(other)
        p = self.p * other.p
        for l in range(abs(self.l - other.l), self.l + other.l + 1):
            yield Irrep(l, p)


class _MulIr(tuple):
    def __new__(cls, mul, ir=None):
        if ir is None:
            mul, ir = mul
        return super().__new__(cls, (int(mul), Irrep(ir)))

    @property
    def mul(self) -> int:
        return self[0]

    @property
    def ir(self) -> Irrep:
        return self[1]

    @property
    def dim(self) -> int:
        return self.mul * self.ir.dim

    def __repr__(self):
        return f"{self.mul}x{self.ir}"


class Irreps(tuple):
    def __new__(cls, irreps=None):
        if isinstance(irreps, Irreps):
            return super().__new__(cls, irreps)
        out: List[_MulIr] = []
        if isinstance(irreps, Irrep):
            out.append(_MulIr(1, irreps))
        elif isinstance(irreps, _MulIr):
            out.append(irreps)
        elif isinstance(irreps, str):
            if irreps.strip() != "":
                for tok in irreps.split("+"):
                    tok = tok.strip()
                    if "x" in tok:
                        mul, ir = tok.split("x")
                        out.append(_MulIr(int(mul), Irrep(ir)))
                    else:
                        out.append(_MulIr(1, Irrep(tok)))
        elif irreps is None:
            pass
        else:
            for item in irreps:
                if isinstance(item, _MulIr):
                    out.append(item)
                elif isinstance(item, (str, Irrep)):
                    out.append(_MulIr(1, Irrep(item)))
                else:
                    mul, ir = item
                    out.append(_MulIr(mul, Irrep(ir)))
        return super().__new__(cls, out)

    @staticmethod
    def spherical_harmonics(lmax: int, p: int = -1) -> "Irreps":
        return Irreps([(1, (l, p ** l)) for l in range(lmax + 1)])

    def slices(self):
        out, i = [], 0
        for mul_ir in self:
            out.append(slice(i, i + mul_ir.dim))
            i += mul_ir.dim
        return out

    @property
    def dim(self) -> int:
        return sum(mul_ir.dim for mul_ir in self)

    @property
    def num_irreps(self) -> int:
        return sum(mul for mul, _ in self)

    @property
    def ls(self) -> List[int]:
        return [ir.l for mul, ir in self for _ in range(mul)]

    @property
    def lmax(self) -> int:
        if len(self) == 0:
            raise ValueError("Cannot get lmax of empty Irreps")
        return max(self.ls)

    def simplify(self) -> "Irreps":
        out = []
        for mul, ir in self:
            if out and out[-1][1] == ir:
                out[-1] = (out[-1][0] + mul, ir)
            elif mul > 0:
                out.append((mul, ir))
        return Irreps(out)

    def remove_zero_multiplicities(self) -> "Irreps":
        return Irreps([(mul, ir) for mul, ir in self if mul > 0])

    def sort(self):
        Ret = collections.namedtuple("sort", ["irreps", "p", "inv"])
        out = sorted((ir, i, mul) for i, (mul, ir) in enumerate(self))
        inv = tuple(i for _, i, _ in out)
        p = [0] * len(inv)
        for i, j in enumerate(inv):
            p[j] = i
        return Ret(Irreps([(mul, ir) for ir, _, mul in out]), tuple(p), inv)

    def count(self, ir) -> int:
        ir = Irrep(ir)
        return sum(mul for mul, ir_ in self if ir_ == ir)

    def __getitem__(self, i):
        x = super().__getitem__(i)
        if isinstance(i, slice):
            return Irreps(x)
        return x

    def __contains__(self, ir) -> bool:
        ir = Irrep(ir)
        return any(ir == ir_ for _, ir_ in self)

    def __add__(self, other):
        return Irreps(super().__add__(Irreps(other)))

    def __radd__(self, other):
        return Irreps(other) + self

    def __mul__(self, k):
        if isinstance(k, Irreps):
            raise NotImplementedError
        return Irreps(super().__mul__(int(k)))

    def __rmul__(self, k):
        return Irreps(super().__rmul__(int(k)))

    def __eq__(self, other):
        try:
            return tuple(self) == tuple(Irreps(other))
        except Exception:
            return False

    def __ne__(self, other):
        return not self == other

    __hash__ = tuple.__hash__

    def __repr__(self):
        return "+".join(f"{mul_ir}" for mul_ir in self)


# ---------------------------------------------------------------------------------------------------------------------
# spherical harmonics: e3nn's generated polynomials ('component' normalised: |Y^l|^2 = 2l + 1 on the unit sphere)
# ---------------------------------------------------------------------------------------------------------------------
def _sh_component(lmax: int, x, y, z):
    sh = [torch.ones_like(x)]
    if lmax >= 1:
        sh += [math.sqrt(3) * x, math.sqrt(3) * y, math.sqrt(3) * z]
    if lmax >= 2:
        x2, y2, z2 = x * x, y * y, z * z
        sh += [math.sqrt(15) * x * z, math.sqrt(15) * x * y, math.sqrt(5) * (y2 - 0.5 * (x2 + z2)),
               math.sqrt(15) * y * z, 0.5 * math.sqrt(15) * (z2 - x2)]
    if lmax >= 3:
        raise NotImplementedError("shim: spherical harmonics up to l = 2")
    return sh


def spherical_harmonics(l, x, normalize: bool, normalization: str = "integral"):
    if isinstance(l, (Irreps, str)):
        ls = [ir.l for _, ir in Irreps(l)]
    elif isinstance(l, int):
        ls = [l]
    else:
        ls = list(l)
    if normalize:
        x = torch.nn.functional.normalize(x, dim=-1)  # x / max(|x|, 1e-12)
    comps = _sh_component(max(ls), x[..., 0], x[..., 1], x[..., 2])
    out = []
    for li in ls:
        blk = torch.stack(comps[li * li:(li + 1) * (li + 1)], dim=-1)
        if normalization == "integral":
            blk = blk / math.sqrt(4 * math.pi)
        elif normalization == "norm":
            blk = blk / math.sqrt(2 * li + 1)
        elif normalization != "component":
            raise ValueError(normalization)
        out.append(blk)
    return torch.cat(out, dim=-1)


# ---------------------------------------------------------------------------------------------------------------------
# Wigner 3j as normalised Gaunt integrals (exact product quadrature: Gauss-Legendre in cos(theta) x uniform in phi)
# ---------------------------------------------------------------------------------------------------------------------
_W3J = {}


def _sphere_quadrature(n_theta: int = 12, n_phi: int = 24):
    import numpy as np
    t, w = np.polynomial.legendre.leggauss(n_theta)
    phi = (np.arange(n_phi) + 0.5) * (2 * np.pi / n_phi)
    ct, ph = np.meshgrid(t, phi, indexing="ij")
    st = np.sqrt(1 - ct * ct)
    # e3nn puts the polar axis on y
    pts = np.stack([st * np.sin(ph), ct, st * np.cos(ph)], axis=-1).reshape(-1, 3)
    wts = (w[:, None] * np.full((1, n_phi), 2 * np.pi / n_phi)).reshape(-1)
    return torch.tensor(pts, dtype=torch.float64), torch.tensor(wts, dtype=torch.float64)


def wigner_3j(l1: int, l2: int, l3: int, dtype=torch.float64, device=None):
    key = (l1, l2, l3)
    if key not in _W3J:
        assert abs(l1 - l2) <= l3 <= l1 + l2 and max(l1, l2, l3) <= 2
        if (l1 + l2 + l3) % 2 == 0:
            pts, wts = _sphere_quadrature()
            y = _sh_component(2, pts[:, 0], pts[:, 1], pts[:, 2])
            blk = lambda l: torch.stack(y[l * l:(l + 1) * (l + 1)], dim=-1)  # noqa: E731
            c = torch.einsum("z,zi,zj,zk->ijk", wts, blk(l1), blk(l2), blk(l3))
            c[c.abs() < 1e-13] = 0.0
            _W3J[key] = c / c.norm()
        else:  # parity-odd couplings (1 x 1 -> 1 ...) are not Gaunt integrals; never used by the SEGNN models
            from oracle.segnn_oracle import wigner_3j as _formula
            _W3J[key] = _formula(l1, l2, l3).clone()
    return _W3J[key].to(dtype=dtype, device=device)


# ---------------------------------------------------------------------------------------------------------------------
# TensorProduct (uvw, shared weights) / FullyConnectedTensorProduct
# ---------------------------------------------------------------------------------------------------------------------
Instruction = collections.namedtuple("Instruction", "i_in1 i_in2 i_out connection_mode has_weight path_weight path_shape")


class TensorProduct(nn.Module):
    def __init__(self, irreps_in1, irreps_in2, irreps_out, instructions, irrep_normalization="component",
                 path_normalization="element", shared_weights=True, internal_weights=None):
        super().__init__()
        self.irreps_in1, self.irreps_in2, self.irreps_out = Irreps(irreps_in1), Irreps(irreps_in2), Irreps(irreps_out)
        assert shared_weights and irrep_normalization == "component" and path_normalization == "element"
        self.shared_weights, self.internal_weights = True, True
        ins = []
        for i1, i2, io, mode, has_weight, pw in instructions:
            assert mode == "uvw" and has_weight
            shape = (self.irreps_in1[i1].mul, self.irreps_in2[i2].mul, self.irreps_out[io].mul)
            ins.append(Instruction(i1, i2, io, mode, True, pw, shape))

        def alpha(i):  # e3nn TensorProduct.__init__: 'component' x 'element' normalisation
            a = self.irreps_out[i.i_out].ir.dim
            x = sum(self.irreps_in1[j.i_in1].mul * self.irreps_in2[j.i_in2].mul for j in ins if j.i_out == i.i_out)
            return math.sqrt(a * i.path_weight / x) if x > 0 else 0.0

        self.instructions = [Instruction(i.i_in1, i.i_in2, i.i_out, i.connection_mode, i.has_weight, alpha(i),
                                         i.path_shape) for i in ins]
        self.weight_numel = sum(math.prod(i.path_shape) for i in self.instructions)
        self.weight = nn.Parameter(torch.randn(self.weight_numel))
        out_mask = torch.cat([
            torch.ones(mul * ir.dim) if any(i.i_out == k and i.path_weight != 0 for i in self.instructions)
            else torch.zeros(mul * ir.dim) for k, (mul, ir) in enumerate(self.irreps_out)]) \
            if self.irreps_out.dim > 0 else torch.ones(0)
        self.register_buffer("output_mask", out_mask)  # e3nn keeps this buffer in the state_dict

    def weight_views(self, weight=None, yield_instruction: bool = False):
        weight = self.weight if weight is None else weight
        offset = 0
        for k, ins in enumerate(self.instructions):
            flat = math.prod(ins.path_shape)
            view = weight.narrow(-1, offset, flat).view(ins.path_shape)
            offset += flat
            yield (k, ins, view) if yield_instruction else view

    def forward(self, x, y, weight=None):
        assert x.shape[-1] == self.irreps_in1.dim and y.shape[-1] == self.irreps_in2.dim
        lead = x.shape[:-1]
        x = x.reshape(-1, self.irreps_in1.dim)
        y = y.reshape(-1, self.irreps_in2.dim)
        s1, s2 = self.irreps_in1.slices(), self.irreps_in2.slices()
        outs = [None] * len(self.irreps_out)
        for ins, w in zip(self.instructions, self.weight_views(weight)):
            (m1, ir1), (m2, ir2), (mo, iro) = self.irreps_in1[ins.i_in1], self.irreps_in2[ins.i_in2], \
                self.irreps_out[ins.i_out]
            x1 = x[:, s1[ins.i_in1]].reshape(-1, m1, ir1.dim)
            x2 = y[:, s2[ins.i_in2]].reshape(-1, m2, ir2.dim)
            w3j = wigner_3j(ir1.l, ir2.l, iro.l, dtype=x.dtype, device=x.device)
            xx = torch.einsum("zui,zvj->zuvij", x1, x2)
            res = ins.path_weight * torch.einsum("uvw,ijk,zuvij->zwk", w.to(x.dtype), w3j, xx)
            res = res.reshape(-1, mo * iro.dim)
            outs[ins.i_out] = res if outs[ins.i_out] is None else outs[ins.i_out] + res
        cols = [o if o is not None else x.new_zeros(x.shape[0], self.irreps_out[k].dim)
                for k, o in enumerate(outs)]
        return torch.cat(cols, dim=-1).reshape(*lead, self.irreps_out.dim)


class FullyConnectedTensorProduct(TensorProduct):
    def __init__(self, irreps_in1, irreps_in2, irreps_out, irrep_normalization=None, path_normalization=None,
                 normalization=None, **kwargs):
        irreps_in1, irreps_in2, irreps_out = Irreps(irreps_in1), Irreps(irreps_in2), Irreps(irreps_out)
        if normalization is not None:  # deprecated alias used at o3_building_blocks.py:48
            irrep_normalization = normalization
        instr = [(i1, i2, io, "uvw", True, 1.0)
                 for i1, (_, ir1) in enumerate(irreps_in1)
                 for i2, (_, ir2) in enumerate(irreps_in2)
                 for io, (_, iro) in enumerate(irreps_out)
                 if iro in ir1 * ir2]
        super().__init__(irreps_in1, irreps_in2, irreps_out, instr,
                         irrep_normalization=irrep_normalization or "component",
                         path_normalization=path_normalization or "element", **kwargs)


class Linear(nn.Module):  # imported (never instantiated) by models/balanced_irreps.py:2
    def __init__(self, *a, **k):
        raise NotImplementedError("shim: e3nn.o3.Linear is not on the SEGNN path")
