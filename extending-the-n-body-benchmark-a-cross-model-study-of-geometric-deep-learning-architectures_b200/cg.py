"""Real-basis Wigner 3j symbols (unit Frobenius norm) for the generic tensor-product path.

e3nn 0.5.1 (un-vendored dependency of the reference, requirements.txt:9) builds them from the SU(2) Clebsch-Gordan
coefficients and its real <-> complex change of basis Q_l (with the (-i)^l phase); this is an independent numpy
restatement of that published construction. The tests compare it against the oracle's torch restatement and against
the closed forms of SURVEY appendix B.
"""
from __future__ import annotations

import math
from functools import lru_cache

import numpy as np


def _su2_cg(j1, m1, j2, m2, j3, m3) -> float:
    """<j1 m1 j2 m2 | j3 m3> (Racah's formula)."""
    if m3 != m1 + m2:
        return 0.0
    f = math.factorial
    pref = math.sqrt((2 * j3 + 1) * f(j3 + j1 - j2) * f(j3 - j1 + j2) * f(j1 + j2 - j3) / f(j1 + j2 + j3 + 1)
                     * f(j3 + m3) * f(j3 - m3) / (f(j1 - m1) * f(j1 + m1) * f(j2 - m2) * f(j2 + m2)))
    lo = max(-j1 + j2 + m3, -j1 + m1, 0)
    hi = min(j2 + j3 + m1, j3 - j1 + j2, j3 + m3)
    total = 0.0
    for v in range(int(lo), int(hi) + 1):
        total += ((-1.0) ** (v + j2 + m2) * f(j2 + j3 + m1 - v) * f(j1 - m1 + v)
                  / (f(v) * f(j3 - j1 + j2 - v) * f(j3 + m3 - v) * f(v + j1 - j2 - m3)))
    return pref * total


def _q(l: int) -> np.ndarray:
    """Change of basis real -> complex spherical harmonics of degree l, e3nn phase convention."""
    q = np.zeros((2 * l + 1, 2 * l + 1), dtype=np.complex128)
    s = 1.0 / math.sqrt(2.0)
    for m in range(-l, 0):
        q[l + m, l - m] = s
        q[l + m, l + m] = -1j * s
    q[l, l] = 1.0
    for m in range(1, l + 1):
        q[l + m, l + m] = (-1) ** m * s
        q[l + m, l - m] = 1j * (-1) ** m * s
    return ((-1j) ** l) * q


@lru_cache(maxsize=None)
def real_wigner_3j(l1: int, l2: int, l3: int) -> np.ndarray:
    """[2l1+1, 2l2+1, 2l3+1] float64, Frobenius norm 1."""
    if not abs(l1 - l2) <= l3 <= l1 + l2:
        raise ValueError("triangle inequality violated")
    c = np.zeros((2 * l1 + 1, 2 * l2 + 1, 2 * l3 + 1), dtype=np.complex128)
    for m1 in range(-l1, l1 + 1):
        for m2 in range(-l2, l2 + 1):
            if abs(m1 + m2) <= l3:
                c[l1 + m1, l2 + m2, l3 + m1 + m2] = _su2_cg(l1, m1, l2, m2, l3, m1 + m2)
    out = np.einsum("ij,kl,mn,ikn->jlm", _q(l1), _q(l2), np.conj(_q(l3).T), c)
    if np.abs(out.imag).max() > 1e-9:
        raise AssertionError("real-basis 3j symbol has an imaginary part")
    out = out.real
    return out / np.linalg.norm(out)
