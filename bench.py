#!/usr/bin/env python
"""Benchmark of the SEGNN self-feed hot path (BASELINE.json metric: particle-steps/s, plus fused edge-msgs/s).

    python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...     # the reference's CPU formulation (oracle port)

Workload (config.workload): the per-GPU shard of BASELINE config 5 -- 1024 independent N=100 charged systems per
GPU (8192 over 8 GPUs), SEGNN 6 layers / hidden 192 / lmax_h 1, eval-mode BatchNorm, autoregressive self-feed
rollout. One "step" advances every simulation of the shard by one model step. Weak scaling: simulations are
sharded across ranks with no data-path collective. Synthetic inputs, random-init weights, seed 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HIDDEN, LAYERS, NBODY = 192, 6, 100
FLOP_PER_EDGE_MSG2 = 20 * (HIDDEN // 2) ** 2       # 184,320: irreducible per-edge contraction
FLOP_PER_EDGE_REFERENCE = 554_880                   # msg1 + msg2 in the reference's formulation
K3_DRAM_BYTES_PER_LAUNCH = 945_620_224 + 151_324_416  # measured with ncu on this workload (1024 sims x 100 bodies)
K3_DRAM_BYTES_PER_LAUNCH_PACKED = 474_707_712 + 71_970_816  # packed-half mode: fp16 projections in, fp16 aggregate rows out


def synthetic_system(batch, n, seed):
    """pos ~ randn * cbrt(N/5), vel ~ randn with the per-system mean removed, charges +-1."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    pos = torch.randn(batch, n, 3, generator=gen) * (n / 5.0) ** (1.0 / 3.0)
    vel = torch.randn(batch, n, 3, generator=gen)
    vel = vel - vel.mean(dim=1, keepdim=True)
    charge = torch.randint(0, 2, (batch, n, 1), generator=gen).float() * 2.0 - 1.0
    return pos, vel, charge


def perturb_batchnorm(model, seed=1):
    gen = torch.Generator(device="cpu").manual_seed(seed)
    with torch.no_grad():
        for mod in model.modules():
            if hasattr(mod, "running_mean") and hasattr(mod, "running_var"):
                mod.running_mean.copy_(0.1 * torch.randn(mod.running_mean.shape, generator=gen))
                mod.running_var.copy_(0.5 + torch.rand(mod.running_var.shape, generator=gen))


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons of one GPU while the timed region runs: through NVML (pynvml, every 5 ms)
    when it is importable, else through nvidia-smi (one sample per ~100 ms)."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._halt = index, [], threading.Event()
        self.max_mhz, self.source = None, "nvidia-smi"
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # LOCAL_RANK indexes the visible devices: resolve through CUDA_VISIBLE_DEVICES when it is a list of indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [v for v in vis.split(",") if v.strip().isdigit()]
            phys = int(ids[index]) if index < len(ids) else index
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM))
            self._nvml, self.source = pynvml, "nvml"
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        nv = self._nvml
        mhz = int(nv.nvmlDeviceGetClockInfo(self._handle, nv.NVML_CLOCK_SM))
        try:
            mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._handle))
        except Exception:
            mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._handle))
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        self.samples.append([str(mhz), str(self.max_mhz)] +
                            ["Active" if mask & bits[n] else "Not Active" for n in self.NAMES])

    def run(self):
        while not self._halt.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                    parts = [p.strip() for p in out.strip().split(",")]
                    if len(parts) >= 6:
                        self.samples.append(parts)
            except Exception:
                pass
            self._halt.wait(0.005 if self._nvml is not None else 0.2)

    def finish(self):
        self._halt.set()
        self.join(timeout=5)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        mhz = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        reasons = [n for i, n in enumerate(self.NAMES)
                   if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None,
                "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples), "source": self.source}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d.get("bf16_tflops_sustained", 1394.7), d.get("hbm_gbs", 6550.1), "measured (MEASURED_PEAKS.json)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def reference_provider():
    """BASELINE.md 3.1 probe: the reference's own SEGNN + rollout code when a checkout is present AND its third-party
    imports resolve (real e3nn / torch_geometric / torch_scatter -> kind 'reference'; the stand-ins of oracle/ref_shims
    -> still the reference's module code, reported as kind 'port' with provider 'reference+shims'); otherwise the
    oracle restatement (kind 'port', provider 'oracle').  The GPU boxes carry no reference checkout."""
    try:
        from oracle import ref_loader
        if ref_loader.available():
            kind = ref_loader.setup()
            return ("reference" if kind == "reference" else "port"), kind
    except Exception as exc:  # noqa: BLE001
        print(f"[bench] reference import failed, using the oracle port: {exc!r}", file=sys.stderr)
    return "port", "oracle"


def cpu_reference_steps(sims, steps, warmup, dtype=torch.float32, provider="oracle", hidden=HIDDEN, layers=LAYERS,
                        nbody=NBODY, lmax_h=1):
    """The reference's CPU formulation (explicit edge_index, gathers, per-path tensor products, scatter-sum) on a
    bounded sample of the workload; returns the list of seconds per timed step."""
    from oracle import segnn_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    pos, vel, mass = synthetic_system(sims, nbody, seed=0)
    pos, vel, mass = pos.to(dtype), vel.to(dtype), mass.to(dtype)
    if provider != "oracle":  # the reference's own classes (helper_scripts/infer_self_feed.py:99-194 loop body)
        from models.segnn.segnn import SEGNN as RefSEGNN
        from models.segnn.o3_building_blocks import O3Transform
        from torch_geometric.data import Data
        from utils.build_fully_connected_graph import build_graph_with_knn
        model = RefSEGNN(hidden_features=hidden, num_layers=layers, lmax_h=lmax_h).to(dtype).eval()

        def one_step(p, v):
            g = Data(pos=p.reshape(-1, 3), vel=v.reshape(-1, 3), force=torch.zeros(sims * nbody, 3, dtype=dtype),
                     mass=mass.reshape(-1, 1))
            g.batch = torch.arange(sims).repeat_interleave(nbody)
            g.edge_index = build_graph_with_knn(g.pos, sims, nbody, torch.device("cpu"), nbody - 1)
            pred = model(O3Transform(1)(g))
            return p + pred[:, :3].reshape(sims, nbody, 3), pred[:, 3:].reshape(sims, nbody, 3)
    else:
        model = O.SEGNN(hidden_features=hidden, num_layers=layers, lmax_h=lmax_h, dtype=dtype).eval()

        def one_step(p, v):
            loc, vv = O.rollout(model, p, v, mass, 1)
            return loc[:, -1], vv[:, -1]
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            pos, vel = one_step(pos, vel)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return times


def median(xs):
    xs = sorted(xs)
    return xs[len(xs) // 2] if len(xs) % 2 else 0.5 * (xs[len(xs) // 2 - 1] + xs[len(xs) // 2])


def cpu_baseline_record(sims, repeats=3, dtype=torch.float32, **kw):
    """1 warm-up + >= 3 timed repeats, median."""
    kind, provider = reference_provider()
    times = cpu_reference_steps(sims, max(3, repeats), 1, dtype=dtype, provider=provider, **kw)
    cores = os.cpu_count() or 1
    nbody = kw.get("nbody", NBODY)
    return {"value": sims * nbody / median(times), "unit": "particle-steps/s", "cores": cores,
            "threads": torch.get_num_threads(), "kind": kind, "provider": provider,
            "dtype": "f64" if dtype == torch.float64 else "f32", "seconds_per_step": [round(t, 4) for t in times],
            "sample": f"{sims} sims x N={nbody} x 1 rollout step; 1 warm-up + {len(times)} timed repeats, median; "
                      f"{cores} host cores"}


def run_reference(args, rank):
    if rank != 0:
        return
    sims = args.cpu_sims
    kind, provider = reference_provider()
    times = cpu_reference_steps(sims, args.steps, max(args.warmup, 1), provider=provider)
    sec = sum(times) / len(times)
    value = sims * NBODY / sec
    cores = os.cpu_count() or 1
    sample = (f"{sims} sims x N={NBODY} x 1 rollout step per bench step, float32, {cores} threads, provider {provider}; "
              f"normalised per particle-step (the GPU arm runs {args.sims_per_gpu} sims per GPU)")
    line = {
        "impl": "reference", "metric": "SEGNN self-feed particle-steps/s", "value": value, "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, sims_per_gpu=sims),
        "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": cores, "kind": kind, "provider": provider,
                         "sample": sample},
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, sims_per_gpu):
    return {"workload": f"cfg5 shard: {sims_per_gpu} independent N={NBODY} charged systems per GPU, SEGNN "
                        f"{LAYERS} layers hidden {HIDDEN} lmax_h 1, eval-BN self-feed rollout",
            "sims_per_gpu": sims_per_gpu, "n_bodies": NBODY, "layers": LAYERS, "hidden": HIDDEN, "lmax_h": 1,
            "parallelism": f"sims sharded x{args.gpus}, no data-path collective",
            "l2_policy": "per-step working set (>1 GB of node projections per layer) exceeds the 126 MB L2"}


def measure_training(S, dev, world, rank, dist, cpu_baseline):
    """Secondary numbers (not the headline metric): BASELINE config 2 -- the README training configuration (SEGNN 6
    layers / hidden 192 / lmax_h 1, 64 graphs x 5 bodies, fp32 kernels, train-mode BatchNorm, hand-written backward,
    AdamW + Noam schedule; forward + backward replayed as a CUDA graph; one flat-bucket NCCL all-reduce per step when
    N > 1) -- and BASELINE config 4 -- one N=1000 graph, hidden 128, forward + backward."""
    out = {}
    torch.manual_seed(0)
    B, N = 64, 5
    model = S.SEGNN(hidden_features=HIDDEN, num_layers=LAYERS, lmax_h=1).to(dev).train()
    ts = S.TrainStep(model, B, N, learning_rate_factor=1.0, process_group=None, distributed=world > 1,
                     use_cuda_graph=True)
    pos, vel, charge = synthetic_system(B, N, seed=77 + rank)
    y = torch.randn(B * N, 6)
    hp, hv, hm, hy = [t.pin_memory() for t in (pos, vel, charge, y)]
    for _ in range(5):
        ts.step(hp, hv, hm, hy)
    steps = 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = S.ops.launch_count()
    e0.record()
    for _ in range(steps):
        loss = ts.step(hp, hv, hm, hy)       # H2D of the batch from pinned memory every step
    loss_host = float(loss)                  # D2H of the loss
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.cpu()) / steps
    out["cfg2_readme_training"] = {
        "config": "SEGNN 6 layers hidden 192 lmax_h 1, batch 64 x N=5 per GPU, fp32 kernels, train-mode BatchNorm, "
                  "AdamW + Noam, fwd+bwd as CUDA graph" + (", DDP flat-bucket all-reduce (NCCL)" if world > 1 else ""),
        "ms_per_step": ms, "train_steps_per_s": 1e3 / ms, "particle_steps_per_s": world * B * N * 1e3 / ms,
        "n_gpus": world, "steps": steps, "loss": loss_host,
        "reference_published": "<= 7.6 train steps/s (BASELINE.md, unrecorded hardware, fp64)"}
    if rank == 0 and world == 1:
        torch.manual_seed(0)
        N4, H4 = 1000, 128
        m4 = S.SEGNN(hidden_features=H4, num_layers=LAYERS, lmax_h=1).to(dev).train()
        p4, v4, c4 = synthetic_system(1, N4, seed=5)
        g = S.GraphBatch(pos=p4.reshape(-1, 3).to(dev), vel=v4.reshape(-1, 3).to(dev), mass=c4.reshape(-1, 1).to(dev),
                         num_graphs=1, n_nodes=N4)
        y4 = torch.randn(N4, 6, device=dev)
        def cfg4_run(keep_bytes):
            saved = S.ops.GEMM_FORM_KEEP_BYTES_PER_LAYER
            S.ops.GEMM_FORM_KEEP_BYTES_PER_LAYER = keep_bytes
            try:
                torch.cuda.empty_cache()
                torch.cuda.reset_peak_memory_stats(dev)  # peak of THIS configuration, not of the workload before it
                times = []
                for i in range(4):
                    m4.zero_grad(set_to_none=True)
                    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a0.record()
                    S.target_common_loss(m4(g), y4).backward()
                    a1.record()
                    torch.cuda.synchronize()
                    times.append(a0.elapsed_time(a1))
                return min(times[1:]), torch.cuda.max_memory_allocated(dev) / 2 ** 30
            finally:
                S.ops.GEMM_FORM_KEEP_BYTES_PER_LAYER = saved
        ms_keep, mem_keep = cfg4_run(S.ops.GEMM_FORM_KEEP_BYTES_PER_LAYER)
        ms_rec, mem_rec = cfg4_run(0)
        # 2 * (8 n^2 MACs forward + 8 n^2 data gradient + 8 n^2 weight gradient) per edge and layer, n = 64
        flop4 = 999000 * LAYERS * 3 * 16 * 64 * 64
        out["cfg4_n1000_fwd_bwd"] = {
            "config": "SEGNN 6 layers hidden 128 lmax_h 1, one N=1000 fully-connected graph (999,000 edges), forward + "
                      "backward, fp32-accurate: edge layers in GEMM form (3xTF32 tcgen05 GEMMs over the edge rows of "
                      "the graph, csrc/segnn_edge_gemm.cu), the forward leaves each layer's edge rows in HBM for its "
                      "backward call",
            "ms_fwd_bwd": ms_keep, "edge_msgs_fwd_bwd_per_s": 999000 * LAYERS / (ms_keep * 1e-3),
            "peak_mem_gb": mem_keep,
            "useful_tflops": flop4 / (ms_keep * 1e-3) / 1e12,
            "recompute_variant": {"note": "GEMM_FORM_KEEP_BYTES_PER_LAYER = 0: nothing per edge kept between forward "
                                          "and backward, rows recomputed per layer", "ms_fwd_bwd": ms_rec,
                                  "peak_mem_gb": mem_rec},
            "round1_fused_fp32_kernels_ms": 207.7}
        if cpu_baseline:
            from oracle import segnn_oracle as O
            torch.set_num_threads(os.cpu_count() or 1)
            om = O.SEGNN(hidden_features=HIDDEN, num_layers=LAYERS, dtype=torch.float32).train()
            og = O.make_graph(pos.reshape(-1, 3), vel.reshape(-1, 3), charge.reshape(-1, 1), B, N)
            opt = torch.optim.AdamW(om.parameters(), weight_decay=1e-8, lr=1e-4, betas=(0.9, 0.98), eps=1e-9)
            ct = []
            for i in range(4):
                t0 = time.perf_counter()
                opt.zero_grad()
                O.target_common_loss(om(og), y).backward()
                opt.step()
                ct.append(time.perf_counter() - t0)
            out["cfg2_readme_training"]["cpu_baseline"] = {
                "train_steps_per_s": 1.0 / min(ct[1:]), "cores": os.cpu_count() or 1, "kind": "port",
                "sample": "3 timed steps after 1 warm-up, float32, oracle port (explicit edges, autograd)"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sims-per-gpu", type=int, default=1024)
    ap.add_argument("--cpu-sims", type=int, default=4, help="bounded CPU sample: simulations per CPU step")
    ap.add_argument("--mode", default="auto", choices=["auto", "fp32", "bf16", "fp16", "fp16p"],
                    help="auto = fp16p: tcgen05 with fp16 operands, fp32 accumulate, packed-half producers (the fastest "
                         "mode inside the 2e-2 budget; needs an even N)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-training", action="store_true", help="skip the secondary training measurements")
    ap.add_argument("--no-secondary", action="store_true", help="skip the other precision modes / BASELINE configs")
    ap.add_argument("--cfg3-sims", type=int, default=64, help="simulations of the lmax_h = 2 (cfg3) measurement")
    ap.add_argument("--cfg4-cpu-seconds", type=int, default=45, help="time limit of the cfg4 CPU attempt")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    import segnn_b200 as S

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # keep stdout for the one JSON line: NCCL's version banner (NCCL_DEBUG=VERSION on some boxes) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    mode = args.mode
    if mode == "auto":
        mode = ("fp16p" if NBODY % 2 == 0 else "bf16") if S.ops.tc_available() else "fp32"

    B, N = args.sims_per_gpu, NBODY
    torch.manual_seed(0)
    model = S.SEGNN(hidden_features=HIDDEN, num_layers=LAYERS, lmax_h=1, compute_mode=mode)
    perturb_batchnorm(model)
    model = model.to(dev).eval()
    start, _ = S.shard_simulations(B * world, rank, world)
    pos, vel, charge = synthetic_system(B, N, seed=1000 + start)

    def timed_rollout(mdl, steps, warmup, use_graph=True, allow_train_mode=False, batch=B, nbody=N, state=None):
        """ms per step of `steps` self-feed steps (CUDA events around the loop, synchronised on both sides)."""
        st = state if state is not None else (pos, vel, charge)
        r = S.SelfFeedRollout(mdl, batch, nbody, dev, max_frames=warmup + steps + 1, use_cuda_graph=use_graph,
                              allow_train_mode=allow_train_mode)
        r.reset(*st)
        r.capture()
        for _ in range(warmup):
            r.step()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(steps):
            r.step()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / steps, r.launches_per_step

    # ---- (1) the headline timed region: K replays of the captured CUDA graph of one self-feed step, state in HBM ----
    roll = S.SelfFeedRollout(model, B, N, dev, max_frames=args.warmup + args.steps + 1, use_cuda_graph=True)
    roll.reset(pos, vel, charge)
    roll.capture()
    for _ in range(args.warmup):
        roll.step()
    sampler = ClockSampler(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()  # samples cover the timed region only (not the wait at the barrier)
    ev0.record()
    for _ in range(args.steps):
        roll.step()
    ev1.record()
    torch.cuda.synchronize()
    clocks = sampler.finish()
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    launches = roll.launches_per_step * args.steps  # kernels of this repo inside the replayed graphs
    del roll

    # ---- (2) instrumented pass (NOT the timed region): the same K steps eagerly, CUDA events around every fused edge
    #          launch on the launching stream -> average K3 launch duration for the roofline ---------------------------
    k3_events = []
    orig_edge_layer = S.ops.edge_layer

    def timed_edge_layer(*a, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig_edge_layer(*a, **kw)
        e1.record()
        k3_events.append((e0, e1))
        return out

    roll_i = S.SelfFeedRollout(model, B, N, dev, max_frames=args.warmup + args.steps + 1, use_cuda_graph=False)
    roll_i.reset(pos, vel, charge)
    roll_i.capture()
    for _ in range(args.warmup):
        roll_i.step()
    S.ops.edge_layer = timed_edge_layer
    torch.cuda.synchronize()
    for _ in range(args.steps):
        roll_i.step()
    torch.cuda.synchronize()
    S.ops.edge_layer = orig_edge_layer
    k3_ms = sum(a.elapsed_time(b) for a, b in k3_events) / max(len(k3_events), 1)
    del roll_i

    # ---- (3) end-to-end region: host buffers in, host buffers out, every step (public API), args.steps steps --------
    h_pos, h_vel, h_mass = pos.clone().pin_memory(), vel.clone().pin_memory(), charge.clone().pin_memory()
    o_pos = torch.empty(B * N, 3).pin_memory()
    o_vel = torch.empty(B * N, 3).pin_memory()
    e2e_steps = args.steps
    roll2 = S.SelfFeedRollout(model, B, N, dev, max_frames=2, use_cuda_graph=True)
    roll2.reset(h_pos, h_vel, h_mass)
    roll2.capture()
    for _ in range(2):
        roll2.reset(h_pos, h_vel, h_mass)
        roll2.step()
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    g0.record()
    in_pos, in_vel = h_pos.reshape(B * N, 3), h_vel.reshape(B * N, 3)
    for _ in range(e2e_steps):
        roll2.reset(in_pos, in_vel, h_mass)        # H2D of this step's inputs (pinned)
        roll2.step()                               # one graph replay
        o_pos.copy_(roll2.pos, non_blocking=True)  # D2H of the step's result
        o_vel.copy_(roll2.vel, non_blocking=True)
        torch.cuda.synchronize()
        # the host now holds the result in (o_pos, o_vel): they are the next step's inputs (two pinned buffer pairs
        # used alternately, instead of a host-side copy back into the first pair)
        in_pos, o_pos = o_pos, in_pos
        in_vel, o_vel = o_vel, in_vel
    g1.record()
    torch.cuda.synchronize()
    e2e_ms = g0.elapsed_time(g1)
    del roll2

    t = torch.tensor([ms, e2e_ms, k3_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, k3_ms = [float(x) for x in t.cpu()]

    # ---- (4) secondary numbers, rank 0 of a single-GPU run only: the other precision modes, the other BASELINE
    #          configurations, the train-mode-BatchNorm rollout of trainer.py:929-942 -------------------------------
    modes, configs = None, None
    if world == 1 and not args.no_secondary:
        modes = {}
        for md in ("fp32", "bf16", "fp16", "fp16p"):
            if md != "fp32" and not S.ops.tc_available():
                continue
            model.compute_mode = md
            m_ms, _ = timed_rollout(model, 3, 3)
            modes[md] = {"ms_per_step": m_ms, "particle_steps_per_s": B * N / (m_ms * 1e-3)}
        model.compute_mode = mode
        configs = {}
        # the reference's in-training rollout never calls model.eval(): batch-statistic BatchNorm (SURVEY note 6)
        model.train()
        with torch.no_grad():
            tb_ms, _ = timed_rollout(model, 2, 1, use_graph=False, allow_train_mode=True)
        model.eval()
        configs["cfg5_train_mode_batchnorm_rollout"] = {
            "config": f"{B} x N={N}, 6 layers hidden 192, train-mode (batch-statistic) BatchNorm as in "
                      "trainer.run_self_feed; " + ("tcgen05 kernels (fp16p, K3 emits the message moments)" if
                                                   mode == "fp16p" else "fp32 kernels") +
                      " + deterministic float64 statistics",
            "ms_per_step": tb_ms, "particle_steps_per_s": B * N / (tb_ms * 1e-3)}
        # cfg1: 4 layers hidden 64, 100 x 5 bodies, 100-step rollout (the reference's CPU-runnable case)
        torch.manual_seed(0)
        m1 = S.SEGNN(hidden_features=64, num_layers=4, lmax_h=1,
                     compute_mode="bf16" if S.ops.tc_available() else "fp32")
        perturb_batchnorm(m1)
        m1 = m1.to(dev).eval()
        c1_ms, c1_launches = timed_rollout(m1, 100, 5, batch=100, nbody=5, state=synthetic_system(100, 5, seed=3))
        configs["cfg1_n5_b100_100step"] = {
            "config": "SEGNN 4 layers hidden 64 lmax_h 1, 100 x N=5 charged systems, 100-step self-feed rollout, "
                      "CUDA-graph replay (launch-latency bound: reported as latency, not roofline)",
            "ms_per_step": c1_ms, "rollout_ms": 100 * c1_ms, "particle_steps_per_s": 500 / (c1_ms * 1e-3),
            "launches_per_step": c1_launches}
        # cfg3: lmax_h = 2 (73x0e+73x1o+73x2e), N=100: generic-irreps fp32 path
        torch.manual_seed(0)
        m3 = S.SEGNN(hidden_features=HIDDEN, num_layers=LAYERS, lmax_h=2)
        perturb_batchnorm(m3)
        m3 = m3.to(dev).eval()
        b3 = args.cfg3_sims
        c3_ms, c3_launches = timed_rollout(m3, 3, 1, use_graph=False, batch=b3, nbody=N,
                                           state=synthetic_system(b3, N, seed=4))
        # the three message_layer_2 GEMMs of a layer: 2 * (146 * 219 + 3 * 219 * 73 + 5 * 146 * 73) flop per edge
        c3_flop = b3 * N * (N - 1) * LAYERS * 2 * (146 * 219 + 3 * 219 * 73 + 5 * 146 * 73)
        configs["cfg3_lmax_h2_n100"] = {
            "config": f"SEGNN 6 layers hidden 192 lmax_h 2 (73x0e+73x1o+73x2e), {b3} x N=100, self-feed step, fp32 "
                      "accurate (1e-5 parity): edge layers in GEMM form (csrc/segnn_l2_rows.cu: message_layer_1 coupling "
                      "+ gate + message_layer_2 coupling in one kernel, three 3xTF32 tcgen05 GEMMs, gate + sender sum "
                      "+ BatchNorm in one kernel), node-level products through the generic kernels; per particle-step, "
                      "the 1000-step rollout is this step repeated",
            "ms_per_step": c3_ms, "particle_steps_per_s": b3 * N / (c3_ms * 1e-3), "launches_per_step": c3_launches,
            "message_layer_2_tflops": c3_flop / (c3_ms * 1e-3) / 1e12,
            "round1_ms_per_step_16_sims": 35.7}
        del m1, m3
        torch.cuda.empty_cache()
    training = None
    if not args.no_training:
        training = measure_training(S, dev, world, rank, dist if world > 1 else None,
                                    cpu_baseline=(world == 1 and not args.no_cpu_baseline))

    if rank == 0:
        tensor_peak, hbm_peak, peak_src = measured_peaks()
        particle_steps = world * B * N * args.steps
        value = particle_steps / (ms * 1e-3)
        e2e_value = world * B * N * e2e_steps / (e2e_ms * 1e-3)
        edges = B * N * (N - 1)
        achieved = edges * FLOP_PER_EDGE_MSG2 / (k3_ms * 1e-3) / 1e12
        line = {
            "metric": "SEGNN self-feed particle-steps/s", "value": value, "unit": "particle-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"bf16": "bf16", "fp16": "f16", "fp16p": "f16"}.get(mode, "f32"), "data": "synthetic",
            "config": workload_config(args, B),
            "timed_region": "K replays of the CUDA graph of one self-feed step (state and trajectory in HBM)",
            "edge_msgs_per_s": world * edges * LAYERS * args.steps / (ms * 1e-3),
            "fused_edge_kernel_edge_msgs_per_s": world * edges / (k3_ms * 1e-3),
            "e2e": {"value": e2e_value, "unit": "particle-steps/s", "h2d_bytes_per_step": B * N * 7 * 4,
                    "d2h_bytes_per_step": B * N * 6 * 4, "steps": e2e_steps, "path": "SelfFeedRollout.reset/step "
                    "(CUDA-graph replay) with pinned host buffers"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"kernel": "edge_layer (K3, message_layer_1 combine + gate + message_layer_2 + gate + "
                                   "aggregation)", "bound": "tensor", "achieved": achieved, "peak": tensor_peak,
                         "unit": "TFLOP/s", "frac": achieved / tensor_peak,
                         "traffic": (K3_DRAM_BYTES_PER_LAUNCH_PACKED if mode == "fp16p" else K3_DRAM_BYTES_PER_LAUNCH)
                         if (mode in ("bf16", "fp16", "fp16p") and B == 1024) else None,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full "
                                           "(" + ("profiles/r2_k3_final2_ncu_summary.json" if mode == "fp16p" else
                                                  "profiles/r1_v15_ncu_full_summary.json")
                                           + "); algorithmic P + Q + agg bytes = "
                                           + ("5.5e8 (fp16 projections in, fp16 aggregate out)" if mode == "fp16p"
                                              else "1.10e9"),
                         "peak_source": peak_src + ", bf16 dense sustained",
                         "flop_per_edge": FLOP_PER_EDGE_MSG2, "edges_per_launch": edges,
                         "avg_launch_ms": k3_ms, "launches_timed": len(k3_events),
                         "timing": "CUDA events around every K3 launch in a separate eager pass over the same steps "
                                   "(the timed region itself replays a CUDA graph and carries no extra events)",
                         "share_of_step": k3_ms * LAYERS / (ms / args.steps),
                         "reference_equivalent_tflops": edges * FLOP_PER_EDGE_REFERENCE / (k3_ms * 1e-3) / 1e12,
                         "note": ("fp32 FFMA mode: the tensor pipe is idle, fraction shown against the bf16 tensor "
                                  "roofline for continuity" if mode not in ("bf16", "fp16", "fp16p") else f"{mode} tcgen05 mode")},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_record(args.cpu_sims, 3, torch.float32)
            line["cpu_baseline_f64"] = cpu_baseline_record(max(1, args.cpu_sims // 2), 3, torch.float64)
            if configs is not None:
                configs["cfg1_n5_b100_100step"]["cpu_baseline"] = cpu_baseline_record(
                    100, 3, torch.float64, hidden=64, layers=4, nbody=5)
                configs["cfg3_lmax_h2_n100"]["cpu_baseline"] = cpu_baseline_record(
                    2, 3, torch.float32, hidden=HIDDEN, layers=LAYERS, nbody=NBODY, lmax_h=2)
                configs["cfg4_n1000_fwd_bwd_cpu"] = cfg4_cpu_attempt(args.cfg4_cpu_seconds)
        if modes is not None:
            line["modes"] = modes
        if configs is not None:
            line["configs"] = configs
        if training is not None:
            line["training"] = training
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cfg4_cpu_attempt(seconds: int):
    """BASELINE config 4 on the host: the reference formulation's forward + backward of ONE N=1000 graph (hidden 128,
    6 layers) stores every [E, .] intermediate for autograd (E = 999,000; the message input alone is [E, 514]).  The
    attempt runs in a child process with an address-space limit and a time limit; whatever happens is reported."""
    import resource
    code = (
        "import os, sys, time, torch, resource\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "lim = 96 << 30\n"
        "resource.setrlimit(resource.RLIMIT_AS, (lim, lim))\n"
        "from oracle import segnn_oracle as O\n"
        "torch.set_num_threads(os.cpu_count() or 1)\n"
        "torch.manual_seed(0)\n"
        "m = O.SEGNN(hidden_features=128, num_layers=6, dtype=torch.float32).train()\n"
        "pos, vel, mass = O.synthetic_system(1, 1000, seed=5, dtype=torch.float32)\n"
        "t0 = time.perf_counter()\n"
        "g = O.make_graph(pos.reshape(-1, 3), vel.reshape(-1, 3), mass.reshape(-1, 1), 1, 1000)\n"
        "out = m(g)\n"
        "print('FWD', time.perf_counter() - t0, flush=True)\n"
        "out.pow(2).mean().backward()\n"
        "print('FWDBWD', time.perf_counter() - t0, resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 2**20, flush=True)\n")
    rec = {"config": "oracle port (reference formulation), float32, one N=1000 graph, hidden 128, 6 layers, "
                     f"forward + backward; limits: {seconds} s, 96 GiB address space", "cores": os.cpu_count() or 1}
    try:
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=seconds)
        txt = out.stdout
        if "FWDBWD" in txt:
            parts = txt.split("FWDBWD")[1].split()
            rec.update(outcome="completed", seconds_fwd_bwd=float(parts[0]), peak_rss_gib=float(parts[1]))
        else:
            tail = (out.stderr or "").strip().splitlines()[-1:] or ["?"]
            rec.update(outcome="reference infeasible at this size", detail=tail[0][:200],
                       forward_seconds=float(txt.split("FWD")[1].split()[0]) if "FWD" in txt else None)
    except subprocess.TimeoutExpired as exc:
        txt = exc.stdout.decode() if isinstance(exc.stdout, bytes) else (exc.stdout or "")
        rec.update(outcome=f"reference infeasible at this size within {seconds} s",
                   forward_seconds=float(txt.split("FWD")[1].split()[0]) if "FWD" in txt else None)
    return rec


if __name__ == "__main__":
    main()
