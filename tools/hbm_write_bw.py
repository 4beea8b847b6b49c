#!/usr/bin/env python
"""Write-only and read-only HBM bandwidth next to the copy figure of MEASURED_PEAKS.json (context for the training step,
whose forward and data-gradient kernels are dominated by stash WRITES)."""
import json
import torch

dev = torch.device("cuda")
n = 1 << 30  # 4 GiB of fp32
x = torch.empty(n, device=dev)
y = torch.empty(n, device=dev)


def t(fn, it=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / it


ms_w = t(lambda: x.fill_(1.0))
ms_c = t(lambda: y.copy_(x))
ms_r = t(lambda: x.sum())
print(json.dumps({"bytes": 4 * n, "write_only_GBs": round(4 * n / ms_w / 1e6, 1), "copy_GBs_read_plus_write": round(8 * n / ms_c / 1e6, 1),
                  "read_only_GBs": round(4 * n / ms_r / 1e6, 1)}))
