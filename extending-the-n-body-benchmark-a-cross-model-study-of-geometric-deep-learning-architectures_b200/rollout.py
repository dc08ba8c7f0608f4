"""Self-feed rollout (helper_scripts/infer_self_feed.py:21-254, segnn branch) with the state resident in HBM:
one autoregressive step = K1 prep -> K2 embed -> L x (node GEMM, fused edge kernel, node update) -> head ->
integrate, captured once as a CUDA graph and replayed; the trajectory is written by the integrate kernel into a
device buffer at a device-side frame cursor. Simulations are independent (eval-mode BatchNorm), so multi-GPU
rollouts shard simulations across ranks with no data-path collective."""
from __future__ import annotations

import os
from datetime import datetime
from typing import Optional, Tuple

import numpy as np
import torch

from . import ops


def shard_simulations(total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [start, start+count) of the ``total`` simulations owned by ``rank``; blocks differ by at
    most one simulation and cover the batch exactly once."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(total, world_size)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


class SelfFeedRollout:
    """Device-resident autoregressive rollout of a SEGNN over B independent N-body systems."""

    def __init__(self, model, batch_size: int, num_nodes: int, device, max_frames: int, use_cuda_graph: bool = True,
                 target: str = "pos_dt+vel", allow_train_mode: bool = False, num_neighbors: Optional[int] = None):
        if target != "pos_dt+vel":
            raise NotImplementedError("only the reference's default target 'pos_dt+vel' is built")
        self.model, self.B, self.N = model, int(batch_size), int(num_nodes)
        # num_neighbors < N - 1: the kNN graph is rebuilt from the predicted positions every step, as
        # infer_self_feed.py:175 does (generic-irreps kernels on the edge list); None / N - 1: the complete graph
        self.num_neighbors = None if num_neighbors is None or int(num_neighbors) == self.N - 1 else int(num_neighbors)
        if self.num_neighbors is not None and not 1 <= self.num_neighbors < self.N:
            raise ValueError("Graph cannot have more neighbors than there are nodes in simulation - 1")
        self.allow_train_mode = bool(allow_train_mode)
        self.device = torch.device(device)
        self.nodes = self.B * self.N
        self.max_frames = int(max_frames)
        f32 = dict(dtype=torch.float32, device=self.device)
        self.pos = torch.zeros((self.nodes, 3), **f32)
        self.vel = torch.zeros((self.nodes, 3), **f32)
        self.mass = torch.ones((self.nodes,), **f32)
        self.traj_pos = torch.zeros((self.max_frames, self.nodes, 3), **f32)
        self.traj_vel = torch.zeros((self.max_frames, self.nodes, 3), **f32)
        self.frame = torch.zeros((1,), dtype=torch.int32, device=self.device)  # next frame slot to write
        self.frames_written = 0  # host mirror of the device cursor (set by reset, advanced by step)
        self.use_cuda_graph = use_cuda_graph
        self._graph = None
        self.launches_per_step = None

    def _check_mode(self):
        # eval-mode BatchNorm is what makes the simulations independent (sharding) and keeps a replayed
        # graph from mutating the running statistics; the in-training rollout of trainer.py:929-942 (train-mode
        # statistics) is available as SelfFeedRollout(..., allow_train_mode=True)
        if self.model.training and not self.allow_train_mode:
            raise RuntimeError("SelfFeedRollout needs model.eval() (eval-mode BatchNorm); pass allow_train_mode=True "
                               "for the reference's in-training rollout with batch statistics")

    @torch.no_grad()
    def reset(self, pos0, vel0, mass):
        """pos0, vel0 [B,N,3]; mass [B,N,1] or [B,N] (host or device, any float dtype). Frame 0 = initial state."""
        self.pos.copy_(torch.as_tensor(pos0).reshape(self.nodes, 3), non_blocking=True)
        self.vel.copy_(torch.as_tensor(vel0).reshape(self.nodes, 3), non_blocking=True)
        self.mass.copy_(torch.as_tensor(mass).reshape(self.nodes), non_blocking=True)
        self.traj_pos[0].copy_(self.pos)
        self.traj_vel[0].copy_(self.vel)
        self.frame.fill_(1)
        self.frames_written = 1

    def _step_eager(self):
        if self.num_neighbors is not None:
            ei = ops.knn_edge_index(self.pos, self.B, self.N, self.num_neighbors, self.device)
            pred = self.model.forward_edge_list(self.pos, self.vel, self.mass, ei, self.B, self.N)
        else:
            pred = self.model.forward_state(self.pos, self.vel, self.mass, self.B, self.N)
        ops.integrate(pred, self.pos, self.vel, self.traj_pos, self.traj_vel, self.frame)
        ops.counter_add(self.frame, 1)

    @torch.no_grad()
    def capture(self):
        """Warm up (packs weights, sets kernel attributes), then capture one step as a CUDA graph."""
        self._check_mode()
        if getattr(self.model, "fused", True) and self.model.compute_mode != "generic" and self.num_neighbors is None:
            self.model.packed(self.N - 1)
        state = (self.pos.clone(), self.vel.clone(), self.frame.clone())
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            before = ops.launch_count()
            self._step_eager()
            self.launches_per_step = ops.launch_count() - before
        torch.cuda.current_stream(self.device).wait_stream(side)
        self.pos.copy_(state[0]); self.vel.copy_(state[1]); self.frame.copy_(state[2])
        if self.use_cuda_graph:
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._step_eager()
            self._graph = graph
            # the capture pass does not execute, state is untouched
        return self

    @torch.no_grad()
    def step(self):
        self._check_mode()
        if self.frames_written >= self.max_frames:
            raise RuntimeError(f"trajectory buffer is full ({self.max_frames} frames): call reset() or allocate a "
                               f"larger max_frames")
        if self.launches_per_step is None:
            self.capture()
        self.frames_written += 1
        if self._graph is not None:
            self._graph.replay()
        else:
            self._step_eager()

    @torch.no_grad()
    def run(self, steps: int):
        """``steps`` more steps from the current cursor; returns every frame written since reset()."""
        if self.frames_written + steps > self.max_frames:
            raise ValueError(f"trajectory buffer holds {self.max_frames} frames, {self.frames_written} are written, "
                             f"cannot add {steps}")
        for _ in range(steps):
            self.step()
        return self.traj_pos[: self.frames_written], self.traj_vel[: self.frames_written]


@torch.no_grad()
def run_inference(model_type, dataloader, model_path=None, model=None, save_dir=None, print_step=True, n_bodies=None,
                  plot_macros=False, num_neighbors=None, device=None, max_rollout_steps=None, ground_truth=None):
    """helper_scripts/infer_self_feed.py:21-254 for model_type='segnn'.

    ``ground_truth`` = (loc_actual [B,T,N,3], vel_actual [B,T,N,3], mass [B,N,1]) replaces the reference's call to
    the CPU simulator (dataset.get_ground_truth_trajectories, :51) when given; otherwise ``dataloader.dataset`` must
    provide ``get_ground_truth_trajectories``. Returns (trajectories_dir, combined_locations [2,B,T,N,3],
    combined_velocities [2,B,T,N,3]) and writes the same four .npy files per simulation (:232-248)."""
    if model_type != "segnn":
        raise ValueError(f"only model_type='segnn' is accelerated, got {model_type!r}")
    if model is None:
        raise ValueError("pass the SEGNN module (loading from model_path is the trainer's job)")
    if plot_macros:
        raise NotImplementedError("macro plotting is outside the accelerated path")
    torch.manual_seed(42)
    if device is None:
        device = next(model.parameters()).device
    if ground_truth is None:
        dataset = dataloader.dataset
        batch_data, _ = dataset.get_ground_truth_trajectories(batch_size=dataset.batch_size)
        loc_actual, vel_actual, _force, mass_actual = [torch.from_numpy(np.array(d)) for d in zip(*batch_data)]
    else:
        loc_actual, vel_actual, mass_actual = [torch.as_tensor(t) for t in ground_truth]
    B, T, N, _ = loc_actual.shape
    if n_bodies is not None and int(n_bodies) != N:
        raise ValueError("n_bodies does not match the ground-truth trajectories")
    if num_neighbors is not None and int(num_neighbors) >= N:
        raise ValueError("Graph cannot have more neighbors than there are nodes in simulation - 1")
    if max_rollout_steps is not None and int(max_rollout_steps) > 0:
        T = min(T, int(max_rollout_steps))
        loc_actual, vel_actual = loc_actual[:, :T], vel_actual[:, :T]
    if print_step:
        print(f"Number of steps to generate: {T}")
    was_training = model.training
    model.eval()
    roll = SelfFeedRollout(model, B, N, device, max_frames=T, num_neighbors=num_neighbors)
    roll.reset(loc_actual[:, 0], vel_actual[:, 0], mass_actual.reshape(B, N))
    tp, tv = roll.run(T - 1)
    model.train(was_training)
    out_dtype = loc_actual.dtype
    loc_pred = tp.reshape(T, B, N, 3).permute(1, 0, 2, 3).to(out_dtype).cpu().numpy()
    vel_pred = tv.reshape(T, B, N, 3).permute(1, 0, 2, 3).to(out_dtype).cpu().numpy()
    loc_act, vel_act = loc_actual.numpy(), vel_actual.numpy()
    combined_locations = np.stack([loc_act, loc_pred], axis=0)
    combined_velocities = np.stack([vel_act, vel_pred], axis=0)
    if not save_dir:
        base = os.path.dirname(model_path) if model_path else "."
        save_dir = f"{base}/generated_trajectories/{datetime.now().strftime('%Y-%m-%d_%H-%M-%S')}"
    trajectories_save_dir = os.path.join(save_dir, "trajectories_data")
    os.makedirs(trajectories_save_dir, exist_ok=True)
    for i in range(B):
        np.save(os.path.join(trajectories_save_dir, f"loc_actual_sim_{i}.npy"), loc_act[i])
        np.save(os.path.join(trajectories_save_dir, f"loc_pred_sim_{i}.npy"), loc_pred[i])
        np.save(os.path.join(trajectories_save_dir, f"vel_actual_sim_{i}.npy"), vel_act[i])
        np.save(os.path.join(trajectories_save_dir, f"vel_pred_sim_{i}.npy"), vel_pred[i])
    return trajectories_save_dir, combined_locations, combined_velocities
