"""world_size-2 `gloo` tests (CPU) of the multi-GPU host logic: the flat-buffer gradient all-reduce + 1/world
averaging of FusedTrainer, the initial parameter broadcast and the eval all-gather.  The device kernels are not
involved: `ops.adam_step` is replaced by the oracle's Adam restatement (test infrastructure) so the step can run
on CPU."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Toy(torch.nn.Module):
    """Two parameter tensors and a quadratic objective: stands in for the pipeline (no CUDA on this box)."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.a = torch.nn.Parameter(torch.randn(5, 3))
        self.b = torch.nn.Parameter(torch.randn(7))

    def forward(self, *, target, evaluation_mode=None):
        return {"objective": ((self.a.sum(dim=0).mean() + self.b - target) ** 2).mean()[None]}


def _worker(rank, world, port, tmp):
    for p in (REPO, os.path.join(REPO, "yet-another-nerf_b200")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import nerf_oracle as O
    from yanerf import ops
    from yanerf.runners import FusedTrainer
    from yanerf.runners.apis import concat_all_gather

    def cpu_adam(params, grads, m, v, lr, step, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
        O.adam_step(params, grads * grad_scale, m, v, step, lr, beta1, beta2, eps)

    ops.adam_step = cpu_adam
    model = _Toy()
    if rank == 1:  # ranks start from different weights; the trainer must broadcast rank 0's
        with torch.no_grad():
            model.a.add_(1.0)
    trainer = FusedTrainer(model, lr=1e-2)
    for s in range(3):
        trainer.train_step({"target": torch.full((7,), float(rank + 1 + s))})
    gathered = concat_all_gather(torch.tensor([float(rank)]))
    torch.save({"flat": trainer.flat.clone(), "gathered": gathered}, os.path.join(tmp, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_flat_allreduce_matches_single_process(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (torch.load(tmp_path / f"r{i}.pt") for i in range(2))
    assert torch.equal(r0["flat"], r1["flat"]), "ranks diverged"
    assert r0["gathered"].tolist() == [0.0, 1.0]
    # single-process reference: Adam on the MEAN of the two ranks' gradients
    from oracle import nerf_oracle as O

    model = _Toy()
    params = [model.a, model.b]
    flat = torch.cat([p.detach().reshape(-1) for p in params])
    m, v = torch.zeros_like(flat), torch.zeros_like(flat)
    for s in range(3):
        grads = []
        for rank in range(2):
            model.zero_grad()
            with torch.no_grad():
                off = 0
                for p in params:
                    p.copy_(flat[off:off + p.numel()].view_as(p))
                    off += p.numel()
            model(target=torch.full((7,), float(rank + 1 + s)))["objective"].mean().backward()
            grads.append(torch.cat([p.grad.reshape(-1) for p in params]))
        O.adam_step(flat, (grads[0] + grads[1]) * 0.5, m, v, s + 1, 1e-2)
    torch.testing.assert_close(r0["flat"], flat, rtol=1e-6, atol=1e-7)


def _slab_worker(rank, world, port, tmp):
    for p in (REPO, os.path.join(REPO, "yet-another-nerf_b200")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from yanerf.pipelines.nerf_pipeline import gather_slabs, slab_bounds

    n_rays, B = 1000, 2  # not a multiple of the 128-ray alignment: the last slab is short
    full = torch.arange(B * n_rays * 5, dtype=torch.float32).reshape(B, n_rays, 5)
    s, e, per = slab_bounds(n_rays, world, rank)
    out = gather_slabs(full[:, s:e].clone(), n_rays, per)
    torch.save({"out": out, "bounds": (s, e, per)}, os.path.join(tmp, f"s{rank}.pt"))
    dist.destroy_process_group()


def test_ray_slab_partition_and_gather(tmp_path):
    """SURVEY 8(e) render sharding: contiguous, tile-aligned slabs that cover every ray once; the all-gather rebuilds
    the full ray list on every rank (world 2, gloo)."""
    for p in (REPO, os.path.join(REPO, "yet-another-nerf_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from yanerf.pipelines.nerf_pipeline import slab_bounds

    for n_rays, world in ((640000, 8), (190512, 4), (1000, 2), (100, 4), (129, 3)):
        covered = 0
        for r in range(world):
            s, e, per = slab_bounds(n_rays, world, r)
            assert s == min(r * per, n_rays) and s <= e <= n_rays and per % 128 == 0
            assert s == covered or s == n_rays
            covered = max(covered, e)
        assert covered == n_rays
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_slab_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    full = torch.arange(2 * 1000 * 5, dtype=torch.float32).reshape(2, 1000, 5)
    for r in range(2):
        got = torch.load(tmp_path / f"s{r}.pt")
        assert torch.equal(got["out"], full)
    assert torch.load(tmp_path / "s0.pt")["bounds"] == (0, 512, 512)
    assert torch.load(tmp_path / "s1.pt")["bounds"] == (512, 1000, 512)
