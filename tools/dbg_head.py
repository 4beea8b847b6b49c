import os, sys
sys.path.insert(0, "/root/repo/yet-another-nerf_b200")
import torch
from yanerf.pipelines.models.nerf_mlp import NeRFMLP
dev = torch.device("cuda")
torch.manual_seed(0)
net = NeRFMLP().to(dev)
for R, P in ((70, 64), (4096, 64), (4096, 192), (1000, 128)):
    o = torch.randn(R, 1, 3, device=dev) * 0.1
    d = torch.randn(R, 1, 3, device=dev)
    z = torch.sort(2 + 4 * torch.rand(R, 1, P, device=dev), dim=-1)[0]
    net.train()
    out_t = net(o, d, z)
    with torch.no_grad():
        out_e = net(o, d, z)
    dr = (out_t["rays_features"] - out_e["rays_features"]).abs()
    dd = (out_t["rays_densities"] - out_e["rays_densities"]).abs()
    print(R, P, "rgb max diff", float(dr.max()), "mean", float(dr.mean()), "dens max diff", float(dd.max()),
          "bad rows", int((dr.reshape(-1, 3).max(-1)[0] > 0.02).sum()), "of", R * P)
    bad = (dr.reshape(-1, 3).max(-1)[0] > 0.02).nonzero().flatten()
    if bad.numel():
        print("  first bad points", bad[:10].tolist(), "tiles", (bad[:10] // 128).tolist(), "last", bad[-5:].tolist())
