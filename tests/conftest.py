import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_ROOT = os.path.join(REPO, "yet-another-nerf_b200")
for p in (REPO, PKG_ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name + ".npz")))

    return load


def oracle_spec(H, W, n_fine, noise_std, chunk, min_depth=2.0, max_depth=6.0):
    """The oracle's PipelineSpec matching `tools.testing.pipeline_cfg` (test infrastructure)."""
    from oracle import nerf_oracle as O

    return O.PipelineSpec(image_height=H, image_width=W, n_pts_fine=n_fine, density_noise_std_train=noise_std,
                          chunk_size_grid=chunk, min_depth=min_depth, max_depth=max_depth)
