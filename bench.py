#!/usr/bin/env python
"""Benchmark of the yanerf hot path on B200 (contract: ONE JSON line on stdout from rank 0).

Headline (top-level keys): BASELINE.json configs[1], one step = one full 800x800 render of configs/nerf/lego.yml's
pipeline (64 coarse + 128 fine samples appended to the coarse ones, two 8x256 NeRF MLPs, random-init weights, synthetic
camera) through `NeRFPipeline.forward(EVALUATION)`, one image per rank per step (weak scaling: how the reference
shards evaluation, DistributedSampler over images, runners/utils.py:112-116).

  value         rays/s, inputs (pose, focal, ground-truth image) already resident in HBM, per-call profiler OFF
  e2e           same call with HOST inputs: pinned H2D of pose/focal/image and D2H of rgb/depth/alpha every step
  roofline      the fine-pass `mlp_fwd_kernel` launch (the dominant kernel): algorithmic FLOPs / CUDA-event time
  cpu_baseline  the oracle port (torch-CPU restatement of the reference, = what `--device cpu` runs) on a fixed
                4-chunk slice of the same image, all host threads (N=1 only)

Sub-records of the same line (the driver runs only `bench.py --gpus N`):
  train               configs[2]: lego.yml training step, 4096 rays/GPU, coarse+fine fwd/bwd + Adam as one CUDA graph,
                      gradient all-reduce for N>1; `cpu_baseline` = configs[0] (oracle step, 1024 rays, N=1 only)
  strong_render       ONE 800x800 image cut into ray slabs over the N ranks + all-gather
  fern                configs[3]: 378x504 LLFF-shaped camera, 64+64 samples: full render and 1024-ray training step
  microbench          configs[4]: 1 Mi rays, alpha-composite fwd/bwd (P=192) and sample_pdf+merge vs the HBM roofline (N=1)
  gpu_eager_baseline  the oracle port executed on the GPU (torch eager, fp32, TF32 off): the "stronger baseline" of
                      SURVEY 8(d) / BASELINE.md 3(c) (N=1)

`--impl reference` times only the CPU path (render headline + the configs[0] training step).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
for _p in (REPO, os.path.join(REPO, "yet-another-nerf_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

if __name__ == "__main__":  # `import bench` from helpers must see this module instance (EMIT, peaks, ...)
    sys.modules.setdefault("bench", sys.modules["__main__"])

N_COARSE = 64
FLOP_PER_POINT_FWD = 2 * 589_952          # the ALGORITHMIC work of the reference's layer list
# what the kernels execute: the linear intermediate layer (256 x 256) is multiplied into the colour hidden layer's weights at
# pack time, so 65 536 MAC/point of the algorithmic count are never issued (forward; the data gradient saves the same)
FLOP_PER_POINT_FWD_EXECUTED = 2 * (589_952 - 65_536)
FLOP_PER_RAY_FWD = 6_912
FLOP_PER_POINT_TRAIN = 3_475_200
CHUNK = 131072
CPU_RENDER_CHUNKS = 4   # fixed size of the CPU render sample (x 2045 rays), independent of --steps
CPU_TRAIN_RAYS = 1024   # BASELINE.json configs[0]

# the two scene shapes of BASELINE.json (SURVEY 8(d) synthetic inputs)
LEGO = dict(name="lego", H=800, W=800, n_fine=128, n_rays=4096, noise=0.2, min_depth=2.0, max_depth=6.0, focal=1111.111)
FERN = dict(name="fern", H=378, W=504, n_fine=64, n_rays=1024, noise=0.0, min_depth=1.2, max_depth=12.0, focal=407.6)
H, W, N_FINE = LEGO["H"], LEGO["W"], LEGO["n_fine"]


def load_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(tensor=float(d["bf16_tflops_sustained"]), tensor_burst=float(d["bf16_tflops"]),
                    hbm=float(d["hbm_gbs"]), source="measured")
    return dict(tensor=1400.0, tensor_burst=1590.0, hbm=6650.0, source="fallback")


def profile_traffic(name: str, launch: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch from a committed `ncu --set full` summary
    (profiles/<name>); None when the file or the launch is absent."""
    path = os.path.join(REPO, "profiles", name)
    if not os.path.exists(path):
        return None
    try:
        rec = json.load(open(path))["launches"][launch]

        def _bytes(txt):
            v, unit = txt.split()
            return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]

        return _bytes(rec["dram__bytes_read.sum"]) + _bytes(rec["dram__bytes_write.sum"])
    except Exception:
        return None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.rows.append(l) for l in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None
        return self

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def synthetic_inputs(rank: int, shape=LEGO):
    from tools import synthetic as syn

    poses = syn.synth_camera(1, seed=rank, jitter=0.0 if rank == 0 else 0.05)
    focal = torch.full((1, 1), shape["focal"])
    image = syn.synth_image(1, shape["H"], shape["W"], seed=1 + rank)
    return poses, focal, image


def build_pipeline_for(shape, device):
    from tools.testing import build_pipeline, load_synth_nets

    pipe = build_pipeline(shape["H"], shape["W"], shape["n_rays"], shape["n_fine"], shape["noise"], CHUNK,
                          min_depth=shape["min_depth"], max_depth=shape["max_depth"]).to(device)
    nets = load_synth_nets(pipe, seeds=(0, 1), gain=1.0)
    return pipe, nets


def build_lego_pipeline(device, n_rays=4096):
    return build_pipeline_for({**LEGO, "n_rays": n_rays}, device)


def oracle_spec_for(shape):
    from oracle import nerf_oracle as O

    return O.PipelineSpec(image_height=shape["H"], image_width=shape["W"], n_pts_fine=shape["n_fine"],
                          density_noise_std_train=shape["noise"], chunk_size_grid=CHUNK, min_depth=shape["min_depth"],
                          max_depth=shape["max_depth"])


# --------------------------------------------------------------------------- CPU / torch-eager baselines (oracle port)
def oracle_render_rate(device, n_chunks: int, shape=LEGO):
    """Oracle render of `n_chunks` consecutive reference-sized chunks of the synthetic image on `device`."""
    from oracle import nerf_oracle as O
    from tools import synthetic as syn

    spec = oracle_spec_for(shape)
    nets = [{k: v.to(device) for k, v in syn.synth_mlp_state(spec.mlp.param_shapes(), s, 1.0).items()} for s in (0, 1)]
    poses, focal, _ = synthetic_inputs(0, shape)
    poses, focal = poses.to(device), focal.to(device)
    _, per = O.chunk_plan(shape["H"] * shape["W"], N_COARSE, CHUNK)
    start = (shape["H"] // 2) * shape["W"]  # centre rows of the image
    sync = (lambda: torch.cuda.synchronize(device)) if torch.device(device).type == "cuda" else (lambda: None)
    with torch.no_grad():
        O.render_image(nets, spec, poses, focal, ray_slice=(start, start + 256))  # warm-up
        sync()
        t0 = time.perf_counter()
        O.render_image(nets, spec, poses, focal, ray_slice=(start, start + per * n_chunks))
        sync()
        dt = time.perf_counter() - t0
    return per * n_chunks / dt, per, dt


def cpu_render_baseline():
    """All host threads, a FIXED sample of CPU_RENDER_CHUNKS x 2045 rays (the same in every run and arm)."""
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    value, per, dt = oracle_render_rate("cpu", CPU_RENDER_CHUNKS)
    return dict(value=value, unit="rays/s", cores=threads, kind="port", sample_ms=dt * 1e3,
                sample=f"{CPU_RENDER_CHUNKS} x {per}-ray chunks of the 800x800 lego render ({per * CPU_RENDER_CHUNKS} rays, "
                       f"{dt:.1f} s), torch {torch.__version__} CPU fp32")


def oracle_train_rate(device, n_rays: int, shape=LEGO, steps: int = 2):
    """The reference's training iteration (runners/apis.py:81-89, scripts/run.py:159) restated by the oracle: forward with
    fresh draws, `objective.mean().backward()`, Adam on every tensor.  One warm-up step, then `steps` timed steps."""
    from oracle import nerf_oracle as O
    from tools import synthetic as syn

    spec = oracle_spec_for(shape)
    nets = [{k: v.to(device).requires_grad_(True) for k, v in syn.synth_mlp_state(spec.mlp.param_shapes(), s, 1.0).items()}
            for s in (0, 1)]
    moments = [{k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in net.items()} for net in nets]
    poses, focal, image = (t.to(device) for t in synthetic_inputs(0, shape))
    sync = (lambda: torch.cuda.synchronize(device)) if torch.device(device).type == "cuda" else (lambda: None)
    n_pix = shape["H"] * shape["W"]
    draws = [{k: v.to(device) for k, v in syn.synth_draws(1, n_rays, n_pix, N_COARSE, shape["n_fine"], seed=s).items()}
             for s in range(steps + 1)]  # generated up front: the timed region is forward + backward + Adam

    def one(step):
        out = O.train_forward(nets, spec, poses, focal, image, draws[step])
        for net in nets:
            for v in net.values():
                v.grad = None
        out["objective"].mean().backward()
        with torch.no_grad():
            for net, mom in zip(nets, moments):
                for k, v in net.items():
                    O.adam_step(v, v.grad, mom[k][0], mom[k][1], step + 1, 5e-4)

    one(0)
    sync()
    t0 = time.perf_counter()
    for s in range(1, steps + 1):
        one(s)
    sync()
    dt = (time.perf_counter() - t0) / steps
    return n_rays / dt, dt


def cpu_train_baseline():
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    value, dt = oracle_train_rate("cpu", CPU_TRAIN_RAYS)
    return dict(value=value, unit="rays/s", cores=threads, kind="port", ms_per_step=dt * 1e3,
                sample=f"BASELINE configs[0]: lego.yml shape, one coarse+fine train step (fwd + bwd + Adam), {CPU_TRAIN_RAYS} rays, "
                       f"64+128 samples, random init; 1 warm-up + 2 timed steps, torch {torch.__version__} CPU fp32")


def gpu_eager_baseline(dev):
    """The same oracle restatement executed by torch eager on the B200 in fp32 with TF32 off (= the reference's ATen
    launch sequence: 8+ GEMMs, ReLUs, cat, ~15 raymarcher ops, ~20 refiner ops per pass)."""
    saved = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        n_chunks = 32
        rate, per, dt = oracle_render_rate(dev, n_chunks)
        train_rate, train_dt = oracle_train_rate(dev, LEGO["n_rays"], steps=5)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = saved
    torch.cuda.empty_cache()
    return {"kind": "oracle port on cuda (torch eager, fp32, TF32 off)", "torch": torch.__version__,
            "render": {"value": round(rate, 1), "unit": "rays/s", "extrapolated": True,
                       "sample": f"{n_chunks} of the 313 chunks of the 800x800 render ({n_chunks * per} rays in {dt * 1e3:.0f} ms); "
                                 "the full image is this rate x 640 000 rays"},
            "train": {"value": round(train_rate, 1), "unit": "rays/s", "ms_per_step": round(train_dt * 1e3, 2),
                      "sample": "lego.yml step, 4096 rays, fwd + bwd + per-tensor Adam, 5 timed steps (draws uploaded beforehand)"}}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    vals, times = [], []
    base = None
    for i in range(args.warmup + steps):
        base = cpu_render_baseline()
        if i >= args.warmup:
            vals.append(base["value"])
            times.append(base["sample_ms"])
    value = sum(vals) / len(vals)
    base["value"] = value
    train = cpu_train_baseline()
    line = dict(metric="render rays/sec (lego.yml 800x800, 64+128 samples)", value=value, unit="rays/s",
                n_gpus=args.gpus, steps=steps, warmup=args.warmup, ms_per_step=round(sum(times) / len(times), 1),
                higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config={"workload": "lego.yml full 800x800 synthetic-camera render (inference, chunked, 64+128 samples)",
                        "note": "reference algorithm on host cores (oracle port); every step renders the same fixed "
                                f"{CPU_RENDER_CHUNKS}-chunk ray sample"},
                cpu_baseline=base,
                e2e={"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                train={"metric": "train rays/sec (lego.yml shape, CPU: 1024 rays)", "value": train["value"], "unit": "rays/s",
                       "ms_per_step": train["ms_per_step"], "cpu_baseline": train})
    emit(line)


class StdoutToStderr:
    """Everything written to fd 1 while the bench runs (NCCL banners, library chatter) goes to stderr, so that
    stdout carries exactly the one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def emit(line) -> None:
    EMIT.append(json.dumps(line))


EMIT = []


# --------------------------------------------------------------------------- GPU arm
class Ctx:
    """rank / world / device + the barrier-bracketed, max-over-ranks CUDA-event timer every number goes through."""

    def __init__(self, rank, world, dev):
        self.rank, self.world, self.dev = rank, world, dev

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps):
        """total ms of `steps` calls, max over ranks."""
        self.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        self.barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=self.dev)
        if self.world > 1:
            import torch.distributed as dist

            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())


def bench_render(ctx: Ctx, shape, steps: int, warmup: int, want_roofline: bool):
    """Full-image render of `shape`, one image per rank.  Returns (record, pipeline)."""
    from yanerf import ops
    from yanerf.pipelines.utils import EvaluationMode

    dev = ctx.dev
    Hs, Ws, n_fine = shape["H"], shape["W"], shape["n_fine"]
    pipe, _ = build_pipeline_for(shape, dev)
    poses, focal, image = synthetic_inputs(ctx.rank, shape)
    extra = {} if shape is LEGO else dict(min_depth=shape["min_depth"], max_depth=shape["max_depth"])
    poses_d, focal_d, image_d = poses.to(dev), focal.to(dev), image.to(dev)
    poses_h, focal_h, image_h = poses.pin_memory(), focal.pin_memory(), image.pin_memory()
    out_h = {k: torch.empty(1, Hs, Ws, c).pin_memory() for k, c in (("rendered_images", 3), ("rendered_depths", 1), ("rendered_alpha_masks", 1))}
    loss_h = torch.empty(1).pin_memory()
    h2d = sum(t.numel() * 4 for t in (poses_h, focal_h, image_h))
    d2h = sum(t.numel() * 4 for t in out_h.values()) + 4

    def step_resident():
        with torch.no_grad():
            return pipe(poses=poses_d, focal_lengths=focal_d, image_rgb=image_d, evaluation_mode=EvaluationMode.EVALUATION, **extra)

    def step_e2e():
        with torch.no_grad():
            p, f, im = (t.to(dev, non_blocking=True) for t in (poses_h, focal_h, image_h))
            preds = pipe(poses=p, focal_lengths=f, image_rgb=im, evaluation_mode=EvaluationMode.EVALUATION, **extra)
            for k, t in out_h.items():
                t.copy_(preds[k], non_blocking=True)
            loss_h.copy_(preds["objective"], non_blocking=True)
        return preds

    for _ in range(max(3, warmup)):
        step_resident()
    torch.cuda.synchronize()
    sampler = ClockSampler(dev.index or 0).start() if ctx.rank == 0 else None
    ops.Profiler.enabled = False  # the headline is timed without the per-call event brackets
    l0 = ops.Profiler.launches
    total_ms = ctx.timed(step_resident, steps)
    launches = (ops.Profiler.launches - l0) // max(1, steps)
    clocks = sampler.stop() if sampler else None
    for _ in range(2):
        step_e2e()
    e2e_ms = ctx.timed(step_e2e, steps)
    rays = Hs * Ws * ctx.world
    rec = dict(value=round(rays * steps / (total_ms * 1e-3), 1), unit="rays/s", ms_per_step=round(total_ms / steps, 3),
               clocks=clocks, gpu_launches=int(launches),
               e2e={"value": round(rays * steps / (e2e_ms * 1e-3), 1), "unit": "rays/s", "ms_per_step": round(e2e_ms / steps, 3),
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h})
    if want_roofline:
        # second pass with CUDA events around every C-ABI call: per-kernel times of the same step.  Profiler records both
        # yn_mlp_fwd launches of every step in order (coarse, fine, coarse, fine, ...)
        peaks = load_peaks()
        ops.Profiler.reset()
        ops.Profiler.enabled = True
        prof_ms = ctx.timed(step_resident, steps)
        ops.Profiler.enabled = False
        prof = ops.Profiler.summary()
        fwd = [s.elapsed_time(e) for name, s, e in ops.Profiler.records if name == "yn_mlp_fwd"]
        fine, coarse = fwd[1::2], fwd[0::2]
        fine_ms = sum(fine) / len(fine)
        flops_fine = Hs * Ws * ((N_COARSE + n_fine) * FLOP_PER_POINT_FWD + FLOP_PER_RAY_FWD)
        achieved = flops_fine / (fine_ms * 1e-3) / 1e12
        ratio = FLOP_PER_POINT_FWD_EXECUTED / FLOP_PER_POINT_FWD
        rec["roofline"] = {
            "bound": "tensor", "kernel": f"mlp_fwd_kernel (fine pass, {N_COARSE + n_fine} points/ray)", "achieved": round(achieved, 1),
            "peak": peaks["tensor"], "peak_source": f"{peaks['source']} bf16_tflops_sustained (kernel timed inside a long step)",
            "unit": "TFLOP/s", "frac": round(achieved / peaks["tensor"], 4),
            "traffic": profile_traffic("mlp_fwd_r02_ncu_full.json", 1) or profile_traffic("mlp_fwd_r01_ncu_full.json", 1),
            "traffic_note": "DRAM bytes of one fine-pass launch from the committed ncu --set full summary under profiles/; "
                            "algorithmic HBM bytes are 20 B/point = 2.46e9",
            "flops_per_launch": flops_fine,
            "executed": {"flops_per_launch": Hs * Ws * ((N_COARSE + n_fine) * FLOP_PER_POINT_FWD_EXECUTED + FLOP_PER_RAY_FWD),
                         "achieved": round(achieved * ratio, 1), "frac": round(achieved * ratio / peaks["tensor"], 4),
                         "note": "`achieved` counts SURVEY 8(d)'s algorithmic FLOPs; the kernel issues 11 % fewer because "
                                 "the linear intermediate layer is folded into the next layer's weights (exact algebra, "
                                 "redone at every weight pack); this is the tensor-pipe rate actually sustained"},
            "launch_ms": round(fine_ms, 3), "coarse_launch_ms": round(sum(coarse) / len(coarse), 3),
            "share_of_step": round((sum(fine) + sum(coarse)) / prof_ms, 4),
            "kernel_ms_per_step": {k: round(v[1] / steps, 3) for k, v in prof.items()},
            "profiled_pass_ms_per_step": round(prof_ms / steps, 3)}
    return rec, pipe


def bench_train(ctx: Ctx, shape, steps: int):
    """Training step of `shape` (n_rays per GPU): FusedTrainer, whole iteration captured as one CUDA graph, NCCL
    all-reduce + Adam behind it for N>1."""
    from yanerf import ops
    from yanerf.runners.engine import FusedTrainer

    dev, world = ctx.dev, ctx.world
    n_rays, n_fine = shape["n_rays"], shape["n_fine"]
    pipe, _ = build_pipeline_for(shape, dev)
    use_graph = os.environ.get("YANERF_TRAIN_GRAPH", "1") != "0"
    trainer = FusedTrainer(pipe, lr=5e-4 * world, use_cuda_graph=use_graph)  # linear LR scaling, scripts/run.py:152-156
    poses, focal, image = synthetic_inputs(ctx.rank, shape)
    # per-image depth bounds as host floats (what DeviceSceneFeed hands out); the reference's LLFF wrapper passes [B,1]
    # tensors that the sampler reduces with .mean().item() right away (ray_sampler.py:280-283)
    extra = {} if shape is LEGO else dict(min_depth=shape["min_depth"], max_depth=shape["max_depth"])
    batch_d = dict(poses=poses.to(dev), focal_lengths=focal.to(dev), image_rgb=image.to(dev), **extra)
    host = dict(poses=poses.pin_memory(), focal_lengths=focal.pin_memory(), image_rgb=image.pin_memory(), **extra)
    loss_h = torch.empty(1).pin_memory()
    h2d = sum(t.numel() * 4 for t in host.values() if torch.is_tensor(t))

    def step_resident():
        return trainer.train_step(batch_d)

    def step_e2e():
        # graph mode: pinned H2D straight into the graph's static inputs
        b = host if use_graph else {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in host.items()}
        preds = trainer.train_step(b)
        loss_h.copy_(preds["objective"], non_blocking=True)
        return preds

    for _ in range(8):  # includes the 3 eager steps before the graph is captured
        step_resident()
    torch.cuda.synchronize()
    sampler = ClockSampler(dev.index or 0).start() if ctx.rank == 0 else None
    total_ms = ctx.timed(step_resident, steps)
    clocks = sampler.stop() if sampler else None
    for _ in range(2):
        step_e2e()
    e2e_ms = ctx.timed(step_e2e, steps)
    # per-kernel times and the launch count come from the same step run eagerly (a graph replay makes the same launches
    # but bypasses the host-side event brackets)
    n_prof = min(steps, 20)
    ops.Profiler.reset()
    ops.Profiler.enabled = True
    l0 = ops.Profiler.launches
    ctx.timed(lambda: trainer.eager_step(batch_d), n_prof)
    ops.Profiler.enabled = False
    launches = (ops.Profiler.launches - l0) // n_prof
    prof = ops.Profiler.summary()
    trainer.finish()
    peaks = load_peaks()
    rays = n_rays * world
    mlp_ms = sum(prof.get(k, (0, 0.0))[1] for k in ("yn_mlp_fwd", "yn_mlp_bwd")) / n_prof
    flops = n_rays * (N_COARSE + N_COARSE + n_fine) * FLOP_PER_POINT_TRAIN
    achieved = flops / (mlp_ms * 1e-3) / 1e12
    traffic = None
    if shape is LEGO:
        parts = [profile_traffic("mlp_train_r02_ncu_full.json", i) for i in range(6)]  # fwd, dgrad, wgrad x (coarse, fine)
        if all(p is not None for p in parts):
            traffic = sum(parts)
    return dict(
        metric=f"train rays/sec ({shape['name']}.yml step, {n_rays} rays/GPU, 64+{N_COARSE + n_fine} points/ray)",
        value=round(rays * steps / (total_ms * 1e-3), 1), unit="rays/s", steps=steps, ms_per_step=round(total_ms / steps, 3),
        scaling="weak", dtype=os.environ.get("YANERF_MLP_TRAIN_DTYPE", "f16") + " operands (16-bit gradients carried times one "
                              "power of two per backward call), f32 accumulate / master weights",
        config={"workload": f"{shape['name']}.yml training step, {n_rays} rays/GPU, coarse+fine fwd/bwd + Adam, ray-sharded DDP",
                "cuda_graph": use_graph, "parallelism": f"dp{world}",
                "collective": "none (N=1)" if world == 1 else "one NCCL all-reduce (sum) of the flat 4.77 MB fp32 gradient per step, "
                                                               "1/world folded into the Adam kernel",
                "l2": "stash + gradient stash of one step exceed the 126 MB L2"},
        clocks=clocks, gpu_launches=int(launches), launch_note="C-ABI kernel launches of one eager step (the graph replays the same)",
        e2e={"value": round(rays * steps / (e2e_ms * 1e-3), 1), "unit": "rays/s", "ms_per_step": round(e2e_ms / steps, 3),
             "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
        roofline={"bound": "tensor", "kernel": "mlp fwd + dgrad + wgrad kernels (coarse + fine)", "achieved": round(achieved, 1),
                  "peak": peaks["tensor"], "peak_source": f"{peaks['source']} bf16_tflops_sustained", "unit": "TFLOP/s",
                  "frac": round(achieved / peaks["tensor"], 4), "traffic": traffic,
                  "traffic_note": "sum of dram bytes of the six MLP launches of one step (ncu --set full, profiles/)",
                  "flops_per_step": flops, "mlp_kernels_ms_per_step": round(mlp_ms, 3),
                  "executed_note": "algorithmic FLOPs of SURVEY 8(d) (3 475 200 per point); the kernels issue 3 213 056: the "
                                   "linear intermediate layer is folded into the colour hidden layer (forward and data "
                                   "gradient), its weight gradient comes from a 128x256x256 post-product",
                  "kernel_ms_per_step": {k: round(v[1] / n_prof, 3) for k, v in prof.items()}},
    )


def bench_strong_render(ctx: Ctx, pipe, steps: int):
    """ONE 800x800 image over all ranks: contiguous ray slabs + all-gather; every rank ends with the full
    image.  At N=1 this is the plain render."""
    from yanerf.pipelines.utils import EvaluationMode
    from yanerf.runners.apis import enable_ray_sharding

    dev = ctx.dev
    poses, focal, image = synthetic_inputs(0)  # the SAME image on every rank
    batch = dict(poses=poses.to(dev), focal_lengths=focal.to(dev), image_rgb=image.to(dev))
    sharded = enable_ray_sharding(pipe)

    def step():
        with torch.no_grad():
            return pipe(**batch, evaluation_mode=EvaluationMode.EVALUATION)

    for _ in range(3):
        step()
    total_ms = ctx.timed(step, steps)
    pipe.ray_shard = None
    return {"metric": "single-image render rays/sec (one 800x800 lego image in ray slabs over all ranks)",
            "value": round(H * W * steps / (total_ms * 1e-3), 1), "unit": "rays/s", "ms_per_image": round(total_ms / steps, 3),
            "scaling": "strong", "sharded": bool(sharded), "n_gpus": ctx.world,
            "collective": "all-gather of the slabs' [rgb, depth, alpha] of both stages" if sharded else "none (N=1)"}


def bench_microbench(dev):
    """configs[4]: 1 Mi rays; algorithmic bytes / CUDA-event time against the measured HBM copy peak."""
    from yanerf import ops

    R = 1 << 20
    hbm = load_peaks()["hbm"]
    out = {"rays": R, "peak_GBs": hbm}

    def timeit(fn, iters=10, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / iters

    def rec(bytes_per_ray, ms):
        gbs = bytes_per_ray * R / (ms * 1e-3) / 1e9
        return {"ms": round(ms, 4), "algorithmic_bytes_per_ray": bytes_per_ray, "achieved_GBs": round(gbs, 1), "frac": round(gbs / hbm, 4)}

    g = torch.Generator(device=dev).manual_seed(0)
    P = 192
    sig = 3 * torch.randn(R, P, device=dev, generator=g) + 0.5
    rgb = torch.rand(R, P, 3, device=dev, generator=g)
    z = torch.sort(2 + 4 * torch.rand(R, P, device=dev, generator=g), dim=-1)[0]
    d = torch.randn(R, 3, device=dev, generator=g)
    gf = torch.randn(R, 3, device=dev, generator=g)
    cfg = ops.march_cfg(1e10, 1e-6, 0.0, False, False, (0.0, 0.0, 0.0))
    with torch.no_grad():
        out["composite_fwd_P192"] = rec(24 * P + 32, timeit(lambda: ops.composite(sig, rgb, z, d, cfg)))
    sig.requires_grad_(True); rgb.requires_grad_(True)
    f = ops.composite(sig, rgb, z, d, cfg)[0]
    ops.Profiler.reset()
    ops.Profiler.enabled = True
    for _ in range(13):
        torch.autograd.grad(f, (sig, rgb), gf, retain_graph=True)
    torch.cuda.synchronize()
    ops.Profiler.enabled = False
    times = [s.elapsed_time(e) for name, s, e in ops.Profiler.records if name == "yn_composite_bwd"][3:]
    out["composite_bwd_P192"] = rec(40 * P + 32, sum(times) / len(times))
    del sig, rgb, z, d, f, gf
    for Pc, n in ((64, 128), (192, 128)):
        z = torch.sort(2 + 4 * torch.rand(R, Pc, device=dev, generator=g), dim=-1)[0]
        w = torch.rand(R, Pc, device=dev, generator=g) ** 4
        u = torch.rand(R, n, device=dev, generator=g)
        out[f"sample_pdf_merge_{Pc}+{n}_random"] = rec((Pc - 2) * 4 + Pc * 4 + n * 4 + (Pc + n) * 4, timeit(lambda: ops.sample_pdf_merge(z, w, n, u)))
        out[f"sample_pdf_merge_{Pc}+{n}_det"] = rec((Pc - 2) * 4 + Pc * 4 + (Pc + n) * 4, timeit(lambda: ops.sample_pdf_merge(z, w, n, None)))
        del z, w, u
    torch.cuda.empty_cache()
    return out


def main():
    with StdoutToStderr():
        _main()
    for text in EMIT:
        print(text, flush=True)


def _main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--train-steps", type=int, default=300, help="timed steps of the training sub-records (>= 1 s of work)")
    ap.add_argument("--only", default="", help="comma list of sub-records to run besides the headline: "
                                               "train,strong,fern,micro,eager,cpu (default: all)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", default="", choices=["", "train", "render", "micro"],
                    help="profiling aid (ncu launch lists / --set full): run ONLY this workload, eagerly, for --steps steps")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the yanerf hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ctx = Ctx(rank, world, dev)
    want = set(filter(None, args.only.split(","))) or {"train", "strong", "fern", "micro", "eager", "cpu"}
    if args.no_cpu_baseline:
        want.discard("cpu")
    if args.profile:
        if args.profile == "train":
            os.environ["YANERF_TRAIN_GRAPH"] = "0"  # eager launches: every kernel is visible to the profiler
            rec = bench_train(ctx, LEGO, max(1, args.steps))
        elif args.profile == "render":
            rec, _ = bench_render(ctx, LEGO, max(1, args.steps), args.warmup, want_roofline=False)
        else:
            rec = bench_microbench(dev)
        if rank == 0:
            emit({"profile": args.profile, **rec})
        return

    head, pipe = bench_render(ctx, LEGO, args.steps, args.warmup, want_roofline=True)
    line = dict(
        metric="render rays/sec (lego.yml 800x800, 64+128 samples)", value=head["value"], unit="rays/s",
        n_gpus=world, steps=args.steps, warmup=max(3, args.warmup), ms_per_step=head["ms_per_step"],
        higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f16 operands, f32 accumulate", data="synthetic",
        config={"workload": "lego.yml full 800x800 synthetic-camera render (inference, chunked, 64+128 samples)",
                "rays_per_step_per_gpu": H * W, "chunk_size_grid": CHUNK, "weights": "random init (seeded), 2 x 595844 params",
                "l2": "per-step working set (>4 GB of depths/densities/colours) exceeds the 126 MB L2; no flush needed",
                "parallelism": f"image-per-rank x{world}", "profiler": "off in the timed region; kernel times from a second pass"},
        clocks=head["clocks"], gpu_launches=head["gpu_launches"], e2e=head["e2e"], roofline=head["roofline"],
    )
    if "strong" in want:
        line["strong_render"] = bench_strong_render(ctx, pipe, args.steps)
    del pipe
    torch.cuda.empty_cache()
    if "train" in want:
        line["train"] = bench_train(ctx, LEGO, args.train_steps)
        torch.cuda.empty_cache()
    if "fern" in want:
        fr, fpipe = bench_render(ctx, FERN, max(args.steps, 10), args.warmup, want_roofline=False)
        del fpipe
        fr["metric"] = "render rays/sec (fern.yml shape 378x504, 64+64 samples)"
        line["fern"] = {"render": fr, "train": bench_train(ctx, FERN, args.train_steps)}
        torch.cuda.empty_cache()
    if world == 1 and rank == 0:
        if "micro" in want:
            line["microbench"] = bench_microbench(dev)
        if "eager" in want:
            line["gpu_eager_baseline"] = gpu_eager_baseline(dev)
        if "cpu" in want:
            line["cpu_baseline"] = cpu_render_baseline()
            if "train" in line:
                line["train"]["cpu_baseline"] = cpu_train_baseline()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(line)


if __name__ == "__main__":
    main()
