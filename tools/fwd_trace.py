#!/usr/bin/env python
"""Debug aid (needs the instrumented library: `make -C yet-another-nerf_b200/csrc clean all INSTRUMENT=1`):
timeline of CTA 0 of mlp_fwd_kernel (YN_FWD_TRACE): per layer, when each MMA issuer / epilogue group
waited and for how long.  Prints cycles relative to the start of the chosen tile pair."""
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
pairs_per_cta = 6
R, P = 148 * 2 * pairs_per_cta, 128
o = torch.randn(R, 1, 3, device=dev) * 0.1
d = torch.randn(R, 1, 3, device=dev)
z = torch.sort(2 + 4 * torch.rand(R, 1, P, device=dev), dim=-1)[0]
with torch.no_grad():
    for _ in range(3):
        net(o, d, z)
    torch.cuda.synchronize()
    path = os.path.join(REPO, "gpurun_out", "fwd_trace.bin")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    os.environ["YN_FWD_TRACE"] = path
    net(o, d, z)
    torch.cuda.synchronize()
    del os.environ["YN_FWD_TRACE"]
raw = np.fromfile(path, dtype=np.int64).reshape(4, 2048, 2)
L = 10
names = ["iss0", "iss1", "epi0", "epi1"]
ev = {}
for r in range(4):
    tags, clk = raw[r, :, 0], raw[r, :, 1]
    n = int((clk != 0).sum())
    ev[r] = (tags[:n], clk[:n])
    print(names[r], "events", n)
# split into pairs: issuer logs tag (l<<8|0) with l == 0 at every pair start
def split(tags, clk):
    starts = [i for i in range(len(tags)) if tags[i] == 0]
    return [(tags[a:b], clk[a:b]) for a, b in zip(starts, starts[1:] + [len(tags)])]
sel = int(os.environ.get("PAIR", 3))
t0 = None
rows = {}
for r in range(4):
    parts = split(*ev[r])
    tg, ck = parts[sel]
    if t0 is None:
        t0 = ck[0]
    rows[r] = (tg, ck - t0)
    print(names[r], "pair", sel, "span", int(ck[-1] - ck[0]), "cycles; pair period",
          int(parts[sel + 1][1][0] - ck[0]) if sel + 1 < len(parts) else -1)
print("\nissuer timeline (cycles since pair start): per layer: wait_epi0 [start->done], wait_epi1 [start->done], block issue times, hfull commits")
for r in (0, 1):
    tg, ck = rows[r]
    for l in range(L):
        m = (tg >> 8) == l
        items = []
        for t, c in zip(tg[m], ck[m]):
            code = t & 255
            nm = {0: "w0>", 1: "w0<", 2: "w1>", 3: "w1<", 4: "HF0", 5: "HF1"}.get(code, None)
            if nm is None:
                nm = f"b{(code >> 3) & 1}{code & 7}"
            items.append(f"{nm}@{int(c)}")
        print(names[r], "L", l, " ".join(items))
print("\nepilogue timeline: >hf0 wait start, hf0 done, b01 done, ep0 done, hf1 done, ep1 done")
for r in (2, 3):
    tg, ck = rows[r]
    for l in range(L):
        m = (tg >> 8) == l
        print(names[r], "L", l, " ".join(f"{int(t & 255)}@{int(c)}" for t, c in zip(tg[m], ck[m])))
    m = (tg >> 8) == 15
    print(names[r], "embedding", " ".join(f"{int(t & 255)}@{int(c)}" for t, c in zip(tg[m], ck[m])))
