"""Forward kernel with its epilogue work / weight copies switched off (YN_FWD_DEBUG bits 1, 2): what bounds a layer.
Needs the instrumented library: `make -C yet-another-nerf_b200/csrc clean all INSTRUMENT=1`."""
import os, sys, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "yet-another-nerf_b200")); sys.path.insert(0, REPO)
from yanerf.pipelines.models.nerf_mlp import NeRFMLP
dev = torch.device("cuda")
torch.manual_seed(0)
net = NeRFMLP().to(dev).eval()
R, P = 148 * 2 * 128, 128   # 37888 rays x 128 = 4.85 M points
o = torch.randn(R, 1, 3, device=dev) * 0.1
d = torch.randn(R, 1, 3, device=dev)
z = torch.sort(2 + 4 * torch.rand(R, 1, P, device=dev), dim=-1)[0]
def run(flag):
    os.environ["YN_FWD_DEBUG"] = str(flag)
    with torch.no_grad():
        for _ in range(3): net(o, d, z)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): net(o, d, z)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"debug={flag}: {ms:.3f} ms  ({R*P/ms/1e3:.1f} M points/s, {R*P*1.05e6/ms/1e9:.0f} TFLOP/s issued)")
for f in (0, 1, 2, 3, 0):
    run(f)
