#!/usr/bin/env python
"""Benchmark of the yanerf hot path on B200 (contract: one JSON line on stdout from rank 0).

Workload at every N (BASELINE.json configs[1], weak scaling = one image per rank per step, which is how the
reference shards evaluation: DistributedSampler over images, runners/utils.py:112-116): one step = one full
800x800 render of configs/nerf/lego.yml's pipeline (64 coarse + 128 fine samples appended to the coarse ones,
two 8x256 NeRF MLPs, random-init weights, synthetic camera) through `NeRFPipeline.forward(EVALUATION)`.

  value   rays/s, inputs (pose, focal, ground-truth image) already resident in HBM
  e2e     same call with HOST inputs: pinned H2D of pose/focal/image and D2H of rgb/depth/alpha every step
  roofline  the fine-pass `mlp_fwd_kernel` launch (the dominant kernel): algorithmic FLOPs / CUDA-event time
  cpu_baseline  the oracle port (torch-CPU restatement of the reference, = what `--device cpu` runs) on a
                bounded ray slice of the same image, all host threads
`--impl reference` times only that CPU path.  `--workload train` reports the lego.yml training step instead
(4096 rays per GPU, coarse+fine forward+backward+Adam, gradient all-reduce for N>1).
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

H = W = 800
N_COARSE, N_FINE = 64, 128
FLOP_PER_POINT_FWD = 2 * 589_952          # the ALGORITHMIC work of the reference's layer list
# what the kernels execute: the linear intermediate layer (256 x 256) is multiplied into the colour hidden layer's weights at
# pack time, so 65 536 MAC/point of the algorithmic count are never issued (forward; the data gradient saves the same)
FLOP_PER_POINT_FWD_EXECUTED = 2 * (589_952 - 65_536)
FLOP_PER_RAY_FWD = 6_912
FLOP_PER_POINT_TRAIN = 3_475_200
CHUNK = 131072


def load_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(tensor=float(d["bf16_tflops_sustained"]), tensor_burst=float(d["bf16_tflops"]),
                    hbm=float(d["hbm_gbs"]), source="measured")
    return dict(tensor=1400.0, tensor_burst=1590.0, hbm=6650.0, source="fallback")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.rows.append(l) for l in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None

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


def synthetic_inputs(rank: int):
    from tools import synthetic as syn

    poses = syn.synth_camera(1, seed=rank, jitter=0.0 if rank == 0 else 0.05)
    focal = torch.full((1, 1), syn.LEGO_FOCAL)
    image = syn.synth_image(1, H, W, seed=1 + rank)
    return poses, focal, image


def build_lego_pipeline(device, n_rays=4096):
    from tools.testing import build_pipeline, load_synth_nets

    pipe = build_pipeline(H, W, n_rays, N_FINE, 0.2, CHUNK).to(device)
    nets = load_synth_nets(pipe, seeds=(0, 1), gain=1.0)
    return pipe, nets


# --------------------------------------------------------------------------- CPU baseline (oracle port)
def cpu_render_baseline(budget_s: float = 12.0):
    """Oracle render of consecutive reference-sized chunks (2045 rays) of the synthetic 800x800 image on all
    host threads; returns rays/s and a description of the sample."""
    from oracle import nerf_oracle as O
    from tools import synthetic as syn
    from tools.testing import LEGO_MLP  # noqa: F401

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    spec = O.PipelineSpec()
    nets = [syn.synth_mlp_state(spec.mlp.param_shapes(), s, 1.0) for s in (0, 1)]
    poses, focal, _ = synthetic_inputs(0)
    _, per = O.chunk_plan(H * W, N_COARSE, CHUNK)
    start = (H // 2) * W  # centre rows of the image
    with torch.no_grad():
        O.render_image(nets, spec, poses, focal, ray_slice=(start, start + 256))  # warm-up
        t0 = time.perf_counter()
        O.render_image(nets, spec, poses, focal, ray_slice=(start, start + per))
        t1 = time.perf_counter() - t0
        n_chunks = max(1, min(8, int(budget_s / max(t1, 1e-3))))
        t0 = time.perf_counter()
        O.render_image(nets, spec, poses, focal, ray_slice=(start, start + per * n_chunks))
        dt = time.perf_counter() - t0
    return dict(value=per * n_chunks / dt, unit="rays/s", cores=threads, kind="port", sample_ms=dt * 1e3,
                sample=f"{n_chunks} x {per}-ray chunks of the 800x800 lego render ({per * n_chunks} rays, {dt:.1f} s), "
                       f"torch {torch.__version__} CPU fp32")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    vals, times = [], []
    base = None
    for i in range(args.warmup + steps):
        base = cpu_render_baseline(budget_s=max(2.0, 60.0 / (args.warmup + steps)))
        if i >= args.warmup:
            vals.append(base["value"])
            times.append(base["sample_ms"])
    value = sum(vals) / len(vals)
    base["value"] = value
    line = dict(metric="render rays/sec (lego.yml 800x800, 64+128 samples)", value=value, unit="rays/s",
                n_gpus=args.gpus, steps=steps, warmup=args.warmup, ms_per_step=round(sum(times) / len(times), 1),
                higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config={"workload": "lego.yml full 800x800 synthetic-camera render (inference, chunked, 64+128 samples)",
                        "note": "reference algorithm on host cores (oracle port), bounded ray sample per step"},
                cpu_baseline=base,
                e2e={"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
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
    ap.add_argument("--workload", default="render", choices=["render", "train"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch.distributed as dist

    from yanerf import ops
    from yanerf.pipelines.utils import EvaluationMode

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
    if args.workload == "train":
        from yanerf.runners.bench_train import run_train_bench

        return run_train_bench(args, rank, world, dev)

    peaks = load_peaks()
    pipe, _ = build_lego_pipeline(dev)
    poses, focal, image = synthetic_inputs(rank)
    poses_d, focal_d, image_d = poses.to(dev), focal.to(dev), image.to(dev)
    poses_h, focal_h, image_h = poses.pin_memory(), focal.pin_memory(), image.pin_memory()
    out_h = {k: torch.empty(1, H, W, c).pin_memory() for k, c in (("rendered_images", 3), ("rendered_depths", 1), ("rendered_alpha_masks", 1))}
    loss_h = torch.empty(1).pin_memory()
    h2d = sum(t.numel() * 4 for t in (poses_h, focal_h, image_h))
    d2h = sum(t.numel() * 4 for t in out_h.values()) + 4

    def step_resident():
        with torch.no_grad():
            return pipe(poses=poses_d, focal_lengths=focal_d, image_rgb=image_d, evaluation_mode=EvaluationMode.EVALUATION)

    def step_e2e():
        with torch.no_grad():
            p, f, im = (t.to(dev, non_blocking=True) for t in (poses_h, focal_h, image_h))
            preds = pipe(poses=p, focal_lengths=f, image_rgb=im, evaluation_mode=EvaluationMode.EVALUATION)
            for k, t in out_h.items():
                t.copy_(preds[k], non_blocking=True)
            loss_h.copy_(preds["objective"], non_blocking=True)
        return preds

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(3, args.warmup)):
        step_resident()
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ops.Profiler.reset()
    ops.Profiler.enabled = True
    launches0 = ops.Profiler.launches
    total_ms = timed(step_resident, args.steps)
    ops.Profiler.enabled = False
    launches = (ops.Profiler.launches - launches0) // max(1, args.steps)
    prof = ops.Profiler.summary()
    clocks = sampler.stop() if rank == 0 else None

    for _ in range(2):
        step_e2e()
    e2e_ms = timed(step_e2e, args.steps)

    rays_per_step = H * W * world
    value = rays_per_step * args.steps / (total_ms * 1e-3)
    e2e_value = rays_per_step * args.steps / (e2e_ms * 1e-3)

    # roofline of the dominant kernel: the fine-pass MLP launch (192 points per ray); Profiler records both
    # yn_mlp_fwd launches of every step in order (coarse, fine, coarse, fine, ...)
    fwd = [(s.elapsed_time(e)) for name, s, e in ops.Profiler.records if name == "yn_mlp_fwd"]
    fine = fwd[1::2]
    coarse = fwd[0::2]
    fine_ms = sum(fine) / len(fine)
    flops_fine = H * W * ((N_COARSE + N_FINE) * FLOP_PER_POINT_FWD + FLOP_PER_RAY_FWD)
    achieved = flops_fine / (fine_ms * 1e-3) / 1e12
    step_ms = total_ms / args.steps
    kernel_ms = {k: round(v[1] / args.steps, 3) for k, v in prof.items()}
    traffic = None
    prof_path = os.path.join(REPO, "profiles", "mlp_fwd_r01_ncu_full.json")
    if os.path.exists(prof_path):  # dram__bytes_read.sum + dram__bytes_write.sum of the fine-pass launch (ncu --set full)
        try:
            fine_launch = json.load(open(prof_path))["launches"][1]

            def _bytes(txt):
                v, unit = txt.split()
                return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]

            traffic = _bytes(fine_launch["dram__bytes_read.sum"]) + _bytes(fine_launch["dram__bytes_write.sum"])
        except Exception:
            traffic = None
    roofline = {"bound": "tensor", "kernel": "mlp_fwd_kernel (fine pass, 192 points/ray)", "achieved": round(achieved, 1),
                "peak": peaks["tensor"], "peak_source": f"{peaks['source']} bf16_tflops_sustained (kernel timed inside a long step)",
                "unit": "TFLOP/s", "frac": round(achieved / peaks["tensor"], 4), "traffic": traffic,
                "traffic_note": "DRAM bytes of one fine-pass launch from profiles/mlp_fwd_r01_ncu_full.json; algorithmic "
                                "HBM bytes are 20 B/point = 2.46e9",
                "flops_per_launch": flops_fine,
                "executed": {"flops_per_launch": H * W * ((N_COARSE + N_FINE) * FLOP_PER_POINT_FWD_EXECUTED + FLOP_PER_RAY_FWD),
                             "achieved": round(achieved * FLOP_PER_POINT_FWD_EXECUTED / FLOP_PER_POINT_FWD, 1),
                             "frac": round(achieved * FLOP_PER_POINT_FWD_EXECUTED / FLOP_PER_POINT_FWD / peaks["tensor"], 4),
                             "note": "`achieved` counts SURVEY 8(d)'s algorithmic FLOPs; the kernel issues 11 % fewer because "
                                     "the linear intermediate layer is folded into the next layer's weights (exact algebra, "
                                     "redone at every weight pack); this is the tensor-pipe rate actually sustained"},
                "launch_ms": round(fine_ms, 3), "coarse_launch_ms": round(sum(coarse) / len(coarse), 3),
                "share_of_step": round((sum(fine) + sum(coarse)) / total_ms, 4), "kernel_ms_per_step": kernel_ms}

    line = None
    if rank == 0:
        line = dict(
            metric="render rays/sec (lego.yml 800x800, 64+128 samples)", value=round(value, 1), unit="rays/s",
            n_gpus=world, steps=args.steps, warmup=max(3, args.warmup), ms_per_step=round(step_ms, 3),
            higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f16 operands, f32 accumulate", data="synthetic",
            config={"workload": "lego.yml full 800x800 synthetic-camera render (inference, chunked, 64+128 samples)",
                    "rays_per_step_per_gpu": H * W, "chunk_size_grid": CHUNK, "weights": "random init (seeded), 2 x 595844 params",
                    "l2": "per-step working set (>4 GB of depths/densities/colours) exceeds the 126 MB L2; no flush needed",
                    "parallelism": f"image-per-rank x{world}"},
            clocks=clocks, gpu_launches=int(launches),
            e2e={"value": round(e2e_value, 1), "unit": "rays/s", "ms_per_step": round(e2e_ms / args.steps, 3),
                 "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            roofline=roofline,
        )
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_render_baseline()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        emit(line)


if __name__ == "__main__":
    main()
