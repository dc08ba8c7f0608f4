"""Deterministic model states for the reference-generated fixtures (shared by ``make_reference_golden.py`` and the
tests): every tensor of a ``state_dict`` is ``lo + (hi - lo) * U[0, 1)`` from a generator seeded per key, so a fixture
only has to carry the (lo, hi) range of each key instead of megabytes of weights."""
import torch


def key_range(key: str, init_bound: float):
    if key.endswith("running_var"):
        return (0.5, 1.5)
    if key.endswith("running_mean"):
        return (-0.2, 0.2)
    if key.endswith("_norm.weight"):
        return (0.75, 1.25)
    if key.endswith("_norm.bias"):
        return (-0.1, 0.1)
    return (-float(init_bound), float(init_bound))


def golden_state(shapes: dict, ranges: dict, seed: int, dtype=torch.float64) -> dict:
    out = {}
    for i, key in enumerate(sorted(shapes)):
        gen = torch.Generator(device="cpu").manual_seed(int(seed) * 100003 + i)
        lo, hi = ranges[key]
        u = torch.rand(tuple(shapes[key]), generator=gen, dtype=torch.float64)
        out[key] = (lo + (hi - lo) * u).to(dtype)
    return out


def sample_grad(g: torch.Tensor, limit: int = 4096):
    """(stride, values): the whole flattened gradient when small, else a strided sample."""
    flat = g.detach().reshape(-1).double()
    stride = 1 if flat.numel() <= limit else flat.numel() // 2048
    return stride, flat[::stride].clone()
