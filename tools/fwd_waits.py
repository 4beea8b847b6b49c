#!/usr/bin/env python
"""Debug aid (needs the instrumented library: `make -C yet-another-nerf_b200/csrc clean all INSTRUMENT=1`): where CTA 0 of
mlp_fwd_kernel waits.  YN_FWD_DEBUG=128: no event log (which perturbs the kernel by ~10 %); issuer 0, epilogue group 0 and
the producer sum the cycles of each kind of barrier wait."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "yet-another-nerf_b200"))
import torch  # noqa: E402

from yanerf.pipelines.models.nerf_mlp import NeRFMLP  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
net = NeRFMLP().to(dev).eval()
pairs_per_cta = 64
R, P = 148 * 2 * pairs_per_cta, 128
o = torch.randn(R, 1, 3, device=dev) * 0.1
d = torch.randn(R, 1, 3, device=dev)
z = torch.sort(2 + 4 * torch.rand(R, 1, P, device=dev), dim=-1)[0]
os.environ["YN_FWD_DEBUG"] = "128"
with torch.no_grad():
    for _ in range(3):
        net(o, d, z)
    torch.cuda.synchronize()
    path = os.path.join(REPO, "gpurun_out", "fwd_waits.bin")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    os.environ["YN_FWD_TRACE"] = path
    net(o, d, z)
    torch.cuda.synchronize()
raw = np.fromfile(path, dtype=np.int64).reshape(4, 2048, 2).reshape(4, -1)
iss, epi, prod = raw[0], raw[2], raw[3]
assert iss[0] == 0x7a11 and epi[0] == 0x7a11 and prod[0] == 0x7a11, "library built without INSTRUMENT=1?"
n_layers = pairs_per_cta * 9  # 8 trunk layers + colour hidden per pair
def show(name, total, parts):
    print(f"{name}: {total} cycles = {total / pairs_per_cta:.0f} per tile pair")
    for k, v in parts:
        print(f"   {k:32s} {v:10d}  {100.0 * v / total:5.1f} %   {v / n_layers:7.0f} per layer")
show("issuer 0", int(iss[1]), [("wait epi_done[0] / next_pair", int(iss[2])), ("wait epi_done[1]", int(iss[3])),
                               ("wait weights (ring full)", int(iss[4])),
                               ("issue + everything else", int(iss[1] - iss[2] - iss[3] - iss[4]))])
show("epilogue group 0 (warp 2, lane 0)", int(epi[1]),
     [("wait half_full[0]", int(epi[2])), ("wait blk01_free", int(epi[3])), ("wait half_full[1]", int(epi[4])),
      ("wait density / colour head", int(epi[5])), ("epilogue work", int(epi[1] - epi[2] - epi[3] - epi[4] - epi[5]))])
show("producer", int(prod[1]), [("wait ring slot empty", int(prod[2])), ("issue", int(prod[1] - prod[2]))])
