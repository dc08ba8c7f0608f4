"""Rollout macros: the acceptance statistics the reference computes on rollouts.

* energies / momentum per (frame, simulation) on the device, straight from the rollout's trajectory buffers
  (trainer.py:888-927 `_compute_nbody_energies`; datasets/nbody/visualization_utils.py:959-960);
* two-sample Kolmogorov-Smirnov p-value and Fisher's combination (utils/ks_utils.py:7-29), host side: they consume a few
  thousand scalars, not trajectories. The Fisher combination uses the closed form of the chi-square survival function
  for even degrees of freedom evaluated in log space, which replaces the reference's 200-digit mpmath sum.
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, Iterable

import numpy as np
import torch

from ._lib import check, lib
from . import ops

__all__ = ["group_collisions", "event_counters", "energy_momentum", "nbody_energies", "momentum_statistics", "ks_statistic", "ks_p",
           "combine_pvalues_fisher", "energy_ratio_steps"]


def energy_momentum(traj_pos: torch.Tensor, traj_vel: torch.Tensor, batch_size: int, num_nodes: int, G: float = 1.0,
                    softening: float = 0.0):
    """traj_pos, traj_vel [frames, B*N, 3] (CUDA, fp32) -> kinetic, potential, momentum, each [frames, B]."""
    if not traj_pos.is_cuda:
        raise RuntimeError("macros run on the device trajectory buffers: there is no CPU fallback")
    traj_pos = traj_pos.to(torch.float32).contiguous()
    traj_vel = traj_vel.to(torch.float32).contiguous()
    frames = traj_pos.shape[0]
    assert traj_pos.shape == (frames, batch_size * num_nodes, 3) and traj_vel.shape == traj_pos.shape
    out = torch.empty((frames, batch_size, 3), dtype=torch.float32, device=traj_pos.device)
    with torch.cuda.device(traj_pos.device):
        check(lib.segnn_macros_energy_momentum(ctypes.c_void_p(traj_pos.data_ptr()),
                                               ctypes.c_void_p(traj_vel.data_ptr()), frames, batch_size, num_nodes,
                                               float(G), float(softening), ctypes.c_void_p(out.data_ptr()),
                                               ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
              "segnn_macros_energy_momentum")
    ops._bump()
    return out[..., 0], out[..., 1], out[..., 2]


def event_counters(traj_pos: torch.Tensor, traj_vel: torch.Tensor, batch_size: int, num_nodes: int,
                   time_threshold: int = 3, contact_distance: float = 0.5, leave_distance: float = 15.0,
                   turn_angle_degrees: float = 30.0) -> Dict[str, np.ndarray]:
    """visualization_utils.py:1093-1222 per simulation: {'stickings', 'collisions', 'bodies_left', 'sharp_turns'}
    (int64 [B]) and 'max_com_distance' (float64 [B]), from the device trajectory buffers [frames, B*N, 3]."""
    if not traj_pos.is_cuda:
        raise RuntimeError("macros run on the device trajectory buffers: there is no CPU fallback")
    traj_pos = traj_pos.to(torch.float32).contiguous()
    traj_vel = traj_vel.to(torch.float32).contiguous()
    frames = traj_pos.shape[0]
    assert traj_pos.shape == (frames, batch_size * num_nodes, 3) and traj_vel.shape == traj_pos.shape
    counts = torch.empty((batch_size, 4), dtype=torch.int32, device=traj_pos.device)
    com = torch.empty((batch_size,), dtype=torch.float32, device=traj_pos.device)
    with torch.cuda.device(traj_pos.device):
        check(lib.segnn_macros_counters(ctypes.c_void_p(traj_pos.data_ptr()), ctypes.c_void_p(traj_vel.data_ptr()),
                                        frames, batch_size, num_nodes, int(time_threshold), float(contact_distance),
                                        float(leave_distance), float(turn_angle_degrees),
                                        ctypes.c_void_p(counts.data_ptr()), ctypes.c_void_p(com.data_ptr()),
                                        ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
              "segnn_macros_counters")
    ops._bump()
    c = counts.cpu().numpy().astype(np.int64)
    return {"stickings": c[:, 0], "collisions": c[:, 1], "bodies_left": c[:, 2], "sharp_turns": c[:, 3],
            "max_com_distance": com.double().cpu().numpy()}


def group_collisions(traj_pos: torch.Tensor, batch_size: int, num_nodes: int, time_threshold: int = 2,
                     distance_threshold: float = 2.0) -> np.ndarray:
    """visualization_utils.py:1455-1610 (the counting part of plot_group_collision_distribution_multiplot): per
    simulation the number of collisions between a stuck pair and a disjoint stuck triplet -> int64 [B]."""
    if not traj_pos.is_cuda:
        raise RuntimeError("macros run on the device trajectory buffers: there is no CPU fallback")
    traj_pos = traj_pos.to(torch.float32).contiguous()
    frames = traj_pos.shape[0]
    assert traj_pos.shape == (frames, batch_size * num_nodes, 3)
    nbytes = int(lib.segnn_macros_group_collisions_workspace(frames, batch_size, num_nodes))
    ws = torch.empty((max(nbytes, 1),), dtype=torch.uint8, device=traj_pos.device)
    counts = torch.empty((batch_size,), dtype=torch.int32, device=traj_pos.device)
    with torch.cuda.device(traj_pos.device):
        check(lib.segnn_macros_group_collisions(ctypes.c_void_p(traj_pos.data_ptr()), frames, batch_size, num_nodes,
                                                int(time_threshold), float(distance_threshold),
                                                ctypes.c_void_p(ws.data_ptr()), ctypes.c_void_p(counts.data_ptr()),
                                                ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
              "segnn_macros_group_collisions")
    ops._bump(4)
    return counts.cpu().numpy().astype(np.int64)


def nbody_energies(traj_pos, traj_vel, batch_size: int, num_nodes: int, G: float, softening: float) -> Dict[str, np.ndarray]:
    """trainer.py:888-927: per-step series averaged across the batch: {'potential', 'kinetic', 'total'}."""
    kin, pot, _ = energy_momentum(traj_pos, traj_vel, batch_size, num_nodes, G, softening)
    pot_series = pot.double().mean(dim=1).cpu().numpy()
    kin_series = kin.double().mean(dim=1).cpu().numpy()
    return {"potential": pot_series, "kinetic": kin_series, "total": pot_series + kin_series}


def momentum_statistics(traj_pos, traj_vel, batch_size: int, num_nodes: int) -> np.ndarray:
    """visualization_utils.py:959-993: mean over time of |sum_i v_i| per simulation -> [B]."""
    _, _, mom = energy_momentum(traj_pos, traj_vel, batch_size, num_nodes)
    return mom.double().mean(dim=0).cpu().numpy()


def energy_ratio_steps(total_pred: np.ndarray, total_true: np.ndarray, bound: float = 2.5) -> int:
    """trainer.py:27,692-701: number of leading steps whose predicted/actual total-energy ratio stays within
    [1/bound, bound]."""
    ratio = np.asarray(total_pred, dtype=np.float64) / np.asarray(total_true, dtype=np.float64)
    ok = (ratio >= 1.0 / bound) & (ratio <= bound)
    bad = np.nonzero(~ok)[0]
    return int(bad[0]) if bad.size else int(ok.size)


# ---- Kolmogorov-Smirnov + Fisher (utils/ks_utils.py) ---------------------------------------------------------------
def ks_statistic(a: np.ndarray, b: np.ndarray) -> float:
    """Two-sample KS statistic D = sup |F_a - F_b|."""
    a, b = np.sort(np.asarray(a, dtype=np.float64).ravel()), np.sort(np.asarray(b, dtype=np.float64).ravel())
    allv = np.concatenate([a, b])
    cdf_a = np.searchsorted(a, allv, side="right") / a.size
    cdf_b = np.searchsorted(b, allv, side="right") / b.size
    return float(np.max(np.abs(cdf_a - cdf_b)))


def _kolmogorov_sf(x: float) -> float:
    """Survival function of the Kolmogorov distribution, Q(x) = 2 sum_{k>=1} (-1)^{k-1} exp(-2 k^2 x^2)."""
    if x <= 0.0:
        return 1.0
    if x < 0.2:  # the alternating series converges slowly here; Q is 1 to double precision
        return 1.0
    s = 0.0
    for k in range(1, 101):
        term = math.exp(-2.0 * k * k * x * x)
        s += term if k % 2 == 1 else -term
        if term < 1e-18:
            break
    return float(min(max(2.0 * s, 0.0), 1.0))


def ks_p(a, b) -> float:
    """utils/ks_utils.py:7-19 `_ks_p`: NaNs dropped, NaN result for empty samples; p-value from scipy.stats.ks_2samp
    when scipy is importable (what the reference calls), else the asymptotic Kolmogorov distribution."""
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    if a.size == 0 or b.size == 0 or np.all(np.isnan(a)) or np.all(np.isnan(b)):
        return float("nan")
    a, b = a[~np.isnan(a)], b[~np.isnan(b)]
    if a.size == 0 or b.size == 0:
        return float("nan")
    try:
        from scipy import stats
        return float(stats.ks_2samp(a, b)[1])
    except ImportError:
        d = ks_statistic(a, b)
        en = math.sqrt(a.size * b.size / (a.size + b.size))
        return _kolmogorov_sf((en + 0.12 + 0.11 / en) * d)


def combine_pvalues_fisher(p_values: Iterable[float]) -> float:
    """utils/ks_utils.py:22-29: chi2.sf(-2 sum log p, 2k), floored at 1e-300; NaN / non-positive p dropped.
    For even degrees of freedom sf(x; 2k) = exp(-x/2) sum_{i<k} (x/2)^i / i!, summed in log space."""
    vals = [float(p) for p in p_values if p == p and p > 0.0]
    if not vals:
        return float("nan")
    half = -sum(math.log(p) for p in vals)  # x / 2
    k = len(vals)
    if half <= 0.0:
        return 1.0
    log_terms = [i * math.log(half) - math.lgamma(i + 1) for i in range(k)]
    m = max(log_terms)
    log_sf = -half + m + math.log(sum(math.exp(t - m) for t in log_terms))
    return float(max(math.exp(log_sf) if log_sf > -745.0 else 0.0, 1e-300))
