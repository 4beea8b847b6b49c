"""cProfile of the HOST side of one evaluation render call (tiny image, so the kernels are negligible).

python tools/host_profile.py [sharded]
"""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    from yanerf.pipelines.utils import EvaluationMode

    pipe, _ = bench.build_lego_pipeline(dev)
    if "sharded" in sys.argv:
        pipe.ray_shard = (0, 1, None)
    poses, focal, image = bench.synthetic_inputs(0)
    batch = dict(poses=poses.to(dev), focal_lengths=focal.to(dev), image_rgb=image.to(dev))

    def step():
        with torch.no_grad():
            return pipe(**batch, evaluation_mode=EvaluationMode.EVALUATION, image_height=16, image_width=16)

    for _ in range(20):
        step()
    torch.cuda.synchronize()
    import time
    t = time.perf_counter()
    for _ in range(500):
        step()
    torch.cuda.synchronize()
    print(f"wall per call: {(time.perf_counter() - t) / 500 * 1e6:.1f} us")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(500):
        step()
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr)
    st.sort_stats("cumulative").print_stats(70)
    st.sort_stats("tottime").print_stats(40)


if __name__ == "__main__":
    main()
