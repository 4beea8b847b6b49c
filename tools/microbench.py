#!/usr/bin/env python
"""BASELINE.json configs[4]: standalone renderer microbench, 1 Mi rays: alpha-composite fwd/bwd (P=192) and
sample_pdf + merge (64 -> +128 random draws; 192 -> +128 stress shape) against the measured HBM roofline.
Prints one JSON line per kernel: algorithmic bytes / CUDA-event time."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "yet-another-nerf_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from yanerf import ops  # noqa: E402

dev = torch.device("cuda")
R = int(os.environ.get("MB_RAYS", 1 << 20))
peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
HBM = float(peaks["hbm_gbs"])
only = sys.argv[1] if len(sys.argv) > 1 else ""


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


def report(name, bytes_per_ray, ms):
    gbs = bytes_per_ray * R / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": name, "rays": R, "ms": round(ms, 4), "algorithmic_bytes_per_ray": bytes_per_ray,
                      "achieved_GBs": round(gbs, 1), "peak_GBs": HBM, "frac": round(gbs / HBM, 4),
                      "rays_per_s": round(R / (ms * 1e-3))}), flush=True)


torch.manual_seed(0)
if only in ("", "composite"):
    P = 192
    sig = 3 * torch.randn(R, P, device=dev) + 0.5
    rgb = torch.rand(R, P, 3, device=dev)
    z = torch.sort(2 + 4 * torch.rand(R, P, device=dev), dim=-1)[0]
    d = torch.randn(R, 3, device=dev)
    gf = torch.randn(R, 3, device=dev)
    cfg = ops.march_cfg(1e10, 1e-6, 0.0, False, False, (0.0, 0.0, 0.0))
    with torch.no_grad():
        ms = timeit(lambda: ops.composite(sig, rgb, z, d, cfg))
    report("composite_fwd P=192", 24 * P + 32, ms)
    sig.requires_grad_(True); rgb.requires_grad_(True)
    f, dep, op, w = ops.composite(sig, rgb, z, d, cfg)
    ms = timeit(lambda: torch.autograd.grad(f, (sig, rgb), gf, retain_graph=True))
    report("composite_bwd P=192 (torch.autograd.grad call: kernel + output allocation + autograd glue)", 40 * P + 32, ms)
    ops.Profiler.reset()
    ops.Profiler.enabled = True
    for _ in range(10):
        torch.autograd.grad(f, (sig, rgb), gf, retain_graph=True)
    torch.cuda.synchronize()
    ops.Profiler.enabled = False
    calls, total = ops.Profiler.summary()["yn_composite_bwd"]
    report("composite_bwd P=192 (kernel only, CUDA events around yn_composite_bwd)", 40 * P + 32, total / calls)
    del sig, rgb, z, d, f, dep, op, w
if only in ("", "pdf"):
    for P, n in ((64, 128), (192, 128)):
        z = torch.sort(2 + 4 * torch.rand(R, P, device=dev), dim=-1)[0]
        w = torch.rand(R, P, device=dev) ** 4
        u = torch.rand(R, n, device=dev)
        ms = timeit(lambda: ops.sample_pdf_merge(z, w, n, u))
        report(f"sample_pdf_merge {P}->+{n} random", (P - 2) * 4 + P * 4 + n * 4 + (P + n) * 4, ms)
        ms = timeit(lambda: ops.sample_pdf_merge(z, w, n, None))
        report(f"sample_pdf_merge {P}->+{n} det", (P - 2) * 4 + P * 4 + (P + n) * 4, ms)
