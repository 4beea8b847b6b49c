"""Device-resident data feed for the training loop.

The reference loads one image per iteration through a DataLoader (`runners/utils.py:112-140`,
`dataset/blender_dataset.py:62-75`): host decode, collate, a 7.7 MB host->device copy per step for an 800x800 image.
All of a Blender / LLFF scene fits in HBM many times over (100 x 800 x 800 x 3 fp32 = 768 MB of 180 GB), so the whole
split is uploaded once and every batch is a VIEW of device memory: no worker processes, no per-step copy, and the
tensors have stable addresses, which is what a captured CUDA graph wants.

Sharding and order follow the reference: `torch.utils.data.DistributedSampler` semantics (a permutation drawn from
`seed + epoch`, padded to a multiple of the world size, rank r takes indices r, r + world, ...), one image per batch,
fields named like the dataset wrappers' NamedTuple fields (`poses`, `focal_lengths`, `image_rgb`, optional
`min_depth` / `max_depth` for LLFF), which are the pipeline's keyword names (`runners/apis.py:55-57`).
Reading image files stays with the caller; this class starts from tensors.
"""
from __future__ import annotations

from typing import Any, Dict, Iterator, Optional

import torch


class DeviceSceneFeed:
    def __init__(self, poses: torch.Tensor, focal_lengths: torch.Tensor, image_rgb: torch.Tensor, device,
                 min_depth: Optional[torch.Tensor] = None, max_depth: Optional[torch.Tensor] = None,
                 rank: int = 0, world_size: int = 1, shuffle: bool = True, seed: int = 0) -> None:
        n = image_rgb.shape[0]
        if poses.shape[0] != n or poses.shape[1:] not in ((3, 4), (4, 4)):
            raise ValueError(f"poses must be [{n},3,4] or [{n},4,4], got {tuple(poses.shape)}")
        if image_rgb.ndim != 4 or image_rgb.shape[-1] != 3:
            raise ValueError("image_rgb must be [N,H,W,3]")  # README.md:79 of the reference
        focal_lengths = torch.as_tensor(focal_lengths, dtype=torch.float32)
        if focal_lengths.ndim == 0:  # one focal length per scene (blender_dataset.py:55)
            focal_lengths = focal_lengths.expand(n)
        if not 0 <= rank < world_size:
            raise ValueError("rank outside [0, world_size)")
        to = lambda t: None if t is None else torch.as_tensor(t, dtype=torch.float32).to(device).contiguous()
        self.poses, self.image_rgb = to(poses), to(image_rgb)
        self.focal_lengths = to(focal_lengths.reshape(n, 1))
        self.min_depth = None if min_depth is None else to(torch.as_tensor(min_depth).reshape(n, 1))
        self.max_depth = None if max_depth is None else to(torch.as_tensor(max_depth).reshape(n, 1))
        self.rank, self.world_size, self.shuffle, self.seed = rank, world_size, shuffle, seed
        self.epoch = 0

    def __len__(self) -> int:
        """Batches per epoch on this rank (= DistributedSampler.num_samples)."""
        return -(-self.image_rgb.shape[0] // self.world_size)

    def set_epoch(self, epoch: int) -> None:
        self.epoch = epoch

    def indices(self) -> list:
        n = self.image_rgb.shape[0]
        if self.shuffle:
            g = torch.Generator()
            g.manual_seed(self.seed + self.epoch)
            order = torch.randperm(n, generator=g).tolist()
        else:
            order = list(range(n))
        total = len(self) * self.world_size
        order += order[: total - n]  # pad by wrapping around, like DistributedSampler
        return order[self.rank:total:self.world_size]

    def batch(self, i: int) -> Dict[str, Any]:
        out = dict(poses=self.poses[i:i + 1], focal_lengths=self.focal_lengths[i:i + 1], image_rgb=self.image_rgb[i:i + 1])
        if self.min_depth is not None:
            out["min_depth"] = self.min_depth[i:i + 1]
        if self.max_depth is not None:
            out["max_depth"] = self.max_depth[i:i + 1]
        return out

    def __iter__(self) -> Iterator[Dict[str, Any]]:
        for i in self.indices():
            yield self.batch(i)
