"""Training step of the SEGNN hot path (trainer.py:170-358 `Trainer.create_optimizer / create_lr_scheduler /
train_one_step`, training/losses.py:22-45 `TargetCommonLoss`) with data-parallel gradient averaging.

The forward (train-mode BatchNorm) and the backward are the hand-written kernel sequences of ``training.py``; what
lives here is the step around them: loss, gradient clipping, AdamW with the reference's hyper-parameters, the
Noam learning-rate schedule, and -- for multi-GPU training -- one all-reduce of a flat gradient bucket over
NCCL / NVLink (each rank draws its own batch; BatchNorm statistics are per rank, buffers are broadcast from rank 0
at start: torch DDP semantics; the reference itself is single-process).

Forward + backward of a fixed batch shape can be captured once as a CUDA graph and replayed: at the README training
size (64 graphs x 5 bodies) the step is launch-latency bound.
"""
from __future__ import annotations

from typing import Optional

import torch

__all__ = ["noam_rate", "target_common_loss", "flatten_gradients", "unflatten_gradients", "allreduce_gradients",
           "allreduce_flat",
           "broadcast_module_state", "TrainStep"]


def noam_rate(step: int, model_size: float, factor: float = 1.0, warmup: int = 4000) -> float:
    """trainer.py:188-195 `_rate`: factor * model_size^-0.5 * min(step^-0.5, step * warmup^-1.5)."""
    if step == 0:
        step = 1
    return factor * (model_size ** (-0.5) * min(step ** (-0.5), step * warmup ** (-1.5)))


def target_common_loss(pred: torch.Tensor, y: torch.Tensor, position_loss_weight: float = 1.0,
                       velocity_loss_weight: float = 1.0) -> torch.Tensor:
    """training/losses.py:22-45 for targets 'pos_dt+vel': weighted MSE of the two 3-vectors."""
    mse = torch.nn.functional.mse_loss
    return position_loss_weight * mse(pred[..., 0:3], y[..., 0:3]) + velocity_loss_weight * mse(pred[..., 3:6], y[..., 3:6])


# ---- data-parallel plumbing (host logic; runs on gloo/CPU tensors in the tests, NCCL on the GPUs) -----------------
def flatten_gradients(params, bucket: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Copies every parameter gradient into one flat fp32 bucket (zeros where a parameter has no gradient)."""
    params = [p for p in params if p.requires_grad]
    total = sum(p.numel() for p in params)
    if bucket is None:
        ref = params[0]
        bucket = torch.empty(total, dtype=torch.float32, device=ref.device)
    off = 0
    for p in params:
        n = p.numel()
        if p.grad is None:
            bucket[off:off + n].zero_()
        else:
            bucket[off:off + n].copy_(p.grad.reshape(-1))
        off += n
    return bucket


def unflatten_gradients(params, bucket: torch.Tensor) -> None:
    off = 0
    for p in params:
        if not p.requires_grad:
            continue
        n = p.numel()
        g = bucket[off:off + n].view_as(p).to(p.dtype)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n


def allreduce_flat(sink: torch.Tensor, group=None) -> torch.Tensor:
    """The gradients already live in one flat buffer (SEGNN.use_flat_storage): ONE in-place all-reduce, averaged by
    the collective itself where the backend can (NCCL ``AVG``), no flatten / unflatten / divide launches."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(sink, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(sink, op=dist.ReduceOp.SUM, group=group)
        sink.div_(world)
    return sink


def allreduce_gradients(params, group=None, bucket: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Average the gradients over the ranks of ``group`` with ONE all-reduce of a flat bucket (1.95 M fp32 = 7.8 MB
    for the README model: latency-bound over NVLink, so a single bucket beats per-tensor collectives)."""
    import torch.distributed as dist
    params = list(params)
    bucket = flatten_gradients(params, bucket)
    world = dist.get_world_size(group)
    dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group)
    bucket.div_(world)
    unflatten_gradients(params, bucket)
    return bucket


def broadcast_module_state(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Parameters and buffers of rank ``src`` to every rank (what DDP does at construction)."""
    import torch.distributed as dist
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)


class TrainStep:
    """One optimisation step of a SEGNN on a fixed batch shape (B graphs x N bodies).

    Hyper-parameters default to the reference's (trainer.py:170-195, config.yaml trainer section): AdamW(lr,
    betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-8), LambdaLR(Noam). ``clip_gradients_value`` /
    ``clip_gradients_norm`` as in `_limit_gradients` (:197-205). ``process_group``: torch.distributed group for
    gradient averaging (None = single process).
    """

    def __init__(self, model, batch_size: int, num_nodes: int, learning_rate: float = 1.0,
                 learning_rate_factor: float = 1.0, learning_rate_warmup_steps: int = 4000,
                 clip_gradients_value: Optional[float] = None, clip_gradients_norm: Optional[float] = None,
                 position_loss_weight: float = 1.0, velocity_loss_weight: float = 1.0, process_group=None,
                 distributed: bool = False, use_cuda_graph: bool = True, fused_optimizer: Optional[bool] = None,
                 flat_storage: bool = True):
        self.model, self.B, self.N = model, int(batch_size), int(num_nodes)
        self.nodes = self.B * self.N
        self.params = [p for p in model.parameters() if p.requires_grad]
        dev = self.params[0].device
        self.device = dev
        # flat parameter / gradient storage: see SEGNN.use_flat_storage (skipped for non-fp32 modules)
        self.flat_params = self.grad_sink = None
        if flat_storage and hasattr(model, "use_flat_storage") and len(self.params) == len(list(model.parameters())) \
                and all(p.dtype == torch.float32 for p in self.params):
            self.flat_params, self.grad_sink = model.use_flat_storage()
        if fused_optimizer is None:
            fused_optimizer = dev.type == "cuda"
        self.optimizer = torch.optim.AdamW(self.params, weight_decay=1e-8, lr=learning_rate, betas=(0.9, 0.98),
                                           eps=1e-9, fused=fused_optimizer)
        size = float(model.get_model_size())
        self.lr_scheduler = torch.optim.lr_scheduler.LambdaLR(
            self.optimizer, lr_lambda=lambda s: noam_rate(s, size, learning_rate_factor, learning_rate_warmup_steps))
        self.clip_value, self.clip_norm = clip_gradients_value, clip_gradients_norm
        self.w_pos, self.w_vel = position_loss_weight, velocity_loss_weight
        self.group, self.distributed = process_group, distributed
        self.use_cuda_graph = use_cuda_graph and dev.type == "cuda"
        f32 = dict(dtype=torch.float32, device=dev)
        # static input buffers (the CUDA graph reads these addresses)
        self.pos = torch.zeros((self.nodes, 3), **f32)
        self.vel = torch.zeros((self.nodes, 3), **f32)
        self.mass = torch.ones((self.nodes,), **f32)
        self.y = torch.zeros((self.nodes, 6), **f32)
        self.loss = torch.zeros((), **f32)
        self._graph = None
        self._bucket = None
        self.step_count = 0
        if distributed:
            broadcast_module_state(model, 0, process_group)

    def _zero_grad(self):
        """With flat storage every backward overwrites the whole gradient buffer and the ``.grad`` views must stay."""
        if self.grad_sink is None:
            self.optimizer.zero_grad(set_to_none=True)

    # -- forward + loss + backward on the static buffers ----------------------------------------------------------
    def _forward_backward(self):
        pred = self.model.forward_state(self.pos, self.vel, self.mass, self.B, self.N)
        loss = target_common_loss(pred, self.y, self.w_pos, self.w_vel)
        loss.backward()
        self.loss.copy_(loss.detach())

    def capture(self):
        """Warm up on a side stream (kernel attributes, allocator), then capture forward + backward as one graph."""
        self.model.train()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        bn_state = [b.clone() for b in self.model.buffers()]
        with torch.cuda.stream(side):
            for _ in range(2):
                self._zero_grad()
                self._forward_backward()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        for b, saved in zip(self.model.buffers(), bn_state):  # warm-up must not move the running statistics
            b.copy_(saved)
        graph = torch.cuda.CUDAGraph()
        self._zero_grad()
        with torch.cuda.graph(graph):
            self._forward_backward()
        self._graph = graph
        return self

    def load_batch(self, pos, vel, mass, y):
        """Host (ideally pinned) or device tensors -> the static buffers."""
        self.pos.copy_(torch.as_tensor(pos).reshape(self.nodes, 3), non_blocking=True)
        self.vel.copy_(torch.as_tensor(vel).reshape(self.nodes, 3), non_blocking=True)
        self.mass.copy_(torch.as_tensor(mass).reshape(self.nodes), non_blocking=True)
        self.y.copy_(torch.as_tensor(y).reshape(self.nodes, 6), non_blocking=True)

    def step(self, pos=None, vel=None, mass=None, y=None) -> torch.Tensor:
        """trainer.py:233-326: zero_grad -> forward -> loss -> backward -> (all-reduce) -> clip -> AdamW -> LR."""
        if pos is not None:
            self.load_batch(pos, vel, mass, y)
        self.model.train()
        if self.use_cuda_graph:
            if self._graph is None:
                self.capture()
            self._graph.replay()  # gradients are rewritten in place by the replay
        else:
            self._zero_grad()
            self._forward_backward()
        flat = self.grad_sink is not None and self.params[0].grad is not None \
            and self.params[0].grad.data_ptr() == self.grad_sink.data_ptr()
        if self.distributed:
            if flat:
                allreduce_flat(self.grad_sink, self.group)
            else:
                self._bucket = allreduce_gradients(self.params, self.group, self._bucket)
        if self.clip_value is not None:
            if flat:
                self.grad_sink.clamp_(-self.clip_value, self.clip_value)
            else:
                torch.nn.utils.clip_grad_value_(self.params, clip_value=self.clip_value)
        if self.clip_norm is not None:
            if flat:  # torch.nn.utils.clip_grad_norm_ on the flat buffer: three launches instead of one per tensor
                norm = torch.linalg.vector_norm(self.grad_sink)
                self.grad_sink.mul_(torch.clamp(self.clip_norm / (norm + 1e-6), max=1.0))
            else:
                torch.nn.utils.clip_grad_norm_(self.params, max_norm=self.clip_norm)
        self.optimizer.step()
        self.lr_scheduler.step()
        self.step_count += 1
        return self.loss
