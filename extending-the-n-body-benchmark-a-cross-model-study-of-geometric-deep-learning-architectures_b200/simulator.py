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


class ChargedSim:
    """Charged-particle ground truth with isolated bodies: the reference's offline generator ``System``
    (datasets/nbody_offline/datagen/system.py:6-123 with n_stick = n_hinge = 0; ``Isolated.update``,
    physical_objects.py:49-57).  The same force law and integrator drive ``ChargedParticlesSim``
    (datasets/nbody/dataset/synthetic_sim.py:155-300).  Charges ride in the ``mass`` slot of the SEGNN graphs
    (dataloaders/segnn_nbody_offline_dataloader.py:78-84)."""

    def __init__(self, n_balls=5, delta_t=0.001, loc_std=1.0, vel_norm=0.5, interaction_strength=1.0,
                 charge_types=(1.0, -1.0)):
        self.n_balls, self.delta_t, self.vel_norm = int(n_balls), float(delta_t), float(vel_norm)
        self.interaction_strength = float(interaction_strength)
        self.max_force = 0.1 / self.delta_t  # system.py:13
        self.loc_std = loc_std * (float(self.n_balls) / 5.0) ** (1 / 3) + 0.1  # system.py:21
        self.charge_types = tuple(charge_types)

    def initial_conditions(self, batch_size: int, seed: int, device):
        """system.py:30-40: charges uniform over charge_types, X ~ N(0, loc_std^2), V of norm vel_norm."""
        gen = torch.Generator(device="cpu").manual_seed(int(seed))
        n = self.n_balls
        types = torch.tensor(self.charge_types, dtype=torch.float64)
        charges = types[torch.randint(0, len(types), (batch_size, n, 1), generator=gen)]
        pos = torch.randn(batch_size, n, 3, generator=gen, dtype=torch.float64) * self.loc_std
        vel = torch.randn(batch_size, n, 3, generator=gen, dtype=torch.float64)
        vel = vel / vel.norm(dim=-1, keepdim=True) * self.vel_norm
        return pos.to(device), vel.to(device), charges.to(device)

    def simulate(self, pos, vel, charges, steps: int, sample_freq: int = 1, device="cuda"):
        """pos, vel [B,N,3], charges [B,N,1] -> X, V [B, steps / sample_freq, N, 3] (float64, device): frame f is the
        state after (f + 1) * sample_freq calls of ``simulate_one_step``."""
        pos, vel, charges = [torch.as_tensor(t, dtype=torch.float64).to(device).clone() for t in (pos, vel, charges)]
        if pos.device.type != "cuda":
            raise RuntimeError("the simulator runs on the device: there is no CPU fallback")
        if steps % sample_freq != 0:
            raise AssertionError("steps % sample_freq == 0")
        B, N = pos.shape[0], pos.shape[1]
        frames = steps // sample_freq
        pos, vel = pos.reshape(B * N, 3).contiguous(), vel.reshape(B * N, 3).contiguous()
        q = charges.reshape(B * N).contiguous()
        tp = torch.empty((frames, B * N, 3), dtype=torch.float64, device=pos.device)
        tv = torch.empty_like(tp)
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        with torch.cuda.device(pos.device):
            check(lib.segnn_sim_charged(p(pos), p(vel), p(q), B, N, self.interaction_strength, self.delta_t,
                                        self.max_force, int(steps), int(sample_freq), p(tp), p(tv),
                                        ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "segnn_sim_charged")
        ops._bump()
        shape = lambda t: t.reshape(frames, B, N, 3).permute(1, 0, 2, 3)
        return shape(tp), shape(tv)
