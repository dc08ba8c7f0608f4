"""Device-side ground-truth generator: the reference's `GravitySim`
(datasets/nbody/dataset/synthetic_sim.py:305-420) as used by the on-the-fly dataset
(datasets/nbody/dataset_gravity_otf.py:38-45,91-107: unit masses, G = interaction_strength, softening, dt = 0.01,
10,000 steps sampled every 10). One launch integrates a whole batch of trajectories in float64."""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import ops
from ._lib import check, lib


class GravitySim:
    """Same constructor arguments as the reference class (synthetic_sim.py:306-316)."""

    def __init__(self, n_balls=100, loc_std=1, vel_norm=0.5, interaction_strength=1, noise_var=0, dt=0.001,
                 softening=0.1, dim=3):
        if dim != 3:
            raise NotImplementedError("the device simulator is three-dimensional")
        self.n_balls, self.loc_std, self.vel_norm = n_balls, loc_std, vel_norm
        self.interaction_strength, self.noise_var, self.dt, self.softening, self.dim = (interaction_strength, noise_var,
                                                                                       dt, softening, dim)

    def initial_conditions(self, batch_size: int, seed: int, device) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """synthetic_sim.py:372-381: pos ~ N(0, cbrt(N/5)^2), vel ~ N(0, 1) in the centre-of-mass frame, unit masses
        (torch generator instead of numpy's global RNG: the stream differs, the distribution does not)."""
        gen = torch.Generator(device="cpu").manual_seed(int(seed))
        n = self.n_balls
        pos = torch.randn(batch_size, n, 3, generator=gen, dtype=torch.float64) * (n / 5.0) ** (1.0 / 3.0)
        vel = torch.randn(batch_size, n, 3, generator=gen, dtype=torch.float64)
        vel = vel - vel.mean(dim=1, keepdim=True)
        mass = torch.ones(batch_size, n, 1, dtype=torch.float64)
        return pos.to(device), vel.to(device), mass.to(device)

    def sample_trajectories(self, batch_size: Optional[int] = None, T: int = 10000, sample_freq: int = 10,
                            random_seed: int = 0, device="cuda", initial_state=None, with_force: bool = True):
        """Batched `sample_trajectory` (:360-418). Returns loc, vel, force [B, T / sample_freq, N, 3] and mass [B, N, 1]
        (float64, on the device). ``initial_state`` = (pos [B,N,3], vel [B,N,3], mass [B,N,1]) overrides the random
        initial conditions."""
        if T % sample_freq != 0:
            raise AssertionError("T % sample_freq == 0")
        if initial_state is None:
            pos, vel, mass = self.initial_conditions(int(batch_size), random_seed, device)
        else:
            pos, vel, mass = [torch.as_tensor(t, dtype=torch.float64).to(device).clone() for t in initial_state]
        if pos.device.type != "cuda":
            raise RuntimeError("the simulator runs on the device: there is no CPU fallback")
        B, N = pos.shape[0], pos.shape[1]
        frames = T // sample_freq
        pos, vel, mass = pos.reshape(B * N, 3).contiguous(), vel.reshape(B * N, 3).contiguous(), mass.reshape(B * N).contiguous()
        tp = torch.empty((frames, B * N, 3), dtype=torch.float64, device=pos.device)
        tv = torch.empty_like(tp)
        tf = torch.empty_like(tp) if with_force else None
        p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
        with torch.cuda.device(pos.device):
            check(lib.segnn_sim_gravity(p(pos), p(vel), p(mass), B, N, float(self.interaction_strength),
                                        float(self.softening), float(self.dt), int(T), int(sample_freq), p(tp), p(tv),
                                        p(tf), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
                  "segnn_sim_gravity")
        ops._bump()
        if self.noise_var:
            gen = torch.Generator(device=pos.device).manual_seed(int(random_seed) + 1)
            for t in (tp, tv) + ((tf,) if tf is not None else ()):
                t.add_(torch.randn(t.shape, generator=gen, dtype=t.dtype, device=t.device) * self.noise_var)
        shape = lambda t: None if t is None else t.reshape(frames, B, N, 3).permute(1, 0, 2, 3)
        return shape(tp), shape(tv), shape(tf), mass.reshape(B, N, 1)
