"""Minimal irreps algebra for the host side (replaces e3nn.o3.Irreps at the call sites of
models/segnn/segnn.py:37-45,209-210 and models/balanced_irreps.py:51-85). Pure integer/string logic."""
from __future__ import annotations

import re
from typing import List, Tuple


class Irreps:
    """Ordered (mul, l, parity) blocks; e3nn string syntax ('96x0e+96x1o'); layout is mul-major per block."""

    def __init__(self, spec=None):
        self.blocks: List[Tuple[int, int, int]] = []
        if spec is None:
            return
        if isinstance(spec, Irreps):
            self.blocks = list(spec.blocks)
        elif isinstance(spec, (list, tuple)):
            self.blocks = [(int(m), int(l), int(p)) for m, l, p in spec]
        else:
            for tok in str(spec).split("+"):
                tok = tok.strip()
                if not tok:
                    continue
                m = re.fullmatch(r"(?:(\d+)x)?(\d+)([eo])", tok)
                if m is None:
                    raise ValueError(f"cannot parse irrep {tok!r}")
                self.blocks.append((int(m.group(1)) if m.group(1) else 1, int(m.group(2)),
                                    1 if m.group(3) == "e" else -1))

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

    def simplify(self) -> "Irreps":
        out: List[Tuple[int, int, int]] = []
        for m, l, p in self.blocks:  # adjacent-only merge, like e3nn
            if out and out[-1][1:] == (l, p):
                out[-1] = (out[-1][0] + m, l, p)
            elif m > 0:
                out.append((m, l, p))
        return Irreps(out)

    def __add__(self, other):
        return Irreps(self.blocks + Irreps(other).blocks)

    def __mul__(self, k):
        return Irreps(self.blocks * int(k))

    __rmul__ = __mul__

    def __iter__(self):
        return iter(self.blocks)

    def __len__(self):
        return len(self.blocks)

    def __getitem__(self, i):
        return Irreps(self.blocks[i]) if isinstance(i, slice) else self.blocks[i]

    def __eq__(self, other):
        return self.blocks == Irreps(other).blocks

    def __repr__(self):
        return "+".join(f"{m}x{l}{'e' if p == 1 else 'o'}" for m, l, p in self.blocks)


def tp_instructions(irreps_in1: Irreps, irreps_in2: Irreps, irreps_out: Irreps):
    """FullyConnectedTensorProduct instruction enumeration (e3nn order: in1, then in2, then out) with the flat
    weight offset of every [mul1, mul2, mul_out] view. Returns (list of dicts, weight_numel)."""
    out, off = [], 0
    for i1, (m1, l1, p1) in enumerate(irreps_in1):
        for i2, (m2, l2, p2) in enumerate(irreps_in2):
            for io, (mo, lo, po) in enumerate(irreps_out):
                if abs(l1 - l2) <= lo <= l1 + l2 and po == p1 * p2:
                    out.append(dict(i1=i1, i2=i2, io=io, shape=(m1, m2, mo), offset=off, ls=(l1, l2, lo)))
                    off += m1 * m2 * mo
    return out, off


def weight_balanced_irreps(hidden_features: int, irreps_attr: Irreps, lmax=None) -> Irreps:
    """models/balanced_irreps.py:51-85 (sh=True): smallest n such that the tensor product
    (n x SH(lmax)) x attr -> (n x SH(lmax)) has at least hidden_features^2 weights."""
    irreps_attr = Irreps(irreps_attr)
    lmax = irreps_attr.lmax if lmax is None else int(lmax)
    n = 1
    while True:
        h = Irreps([(n, l, (-1) ** l) for l in range(lmax + 1)])
        if tp_instructions(h, irreps_attr, h)[1] >= hidden_features * hidden_features:
            return h
        n += 1
