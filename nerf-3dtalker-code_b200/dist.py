"""Multi-GPU plumbing for the hot path: one process per GPU (torchrun), rays / batch items sharded across
ranks with NO data-path collective (every ray is independent, weights are replicated), and ONE NCCL
all-reduce per training step over a single flat fp32 gradient bucket.  The reference trains on
a single GPU (talker_trainer.py:704-712); this is the data-parallel extension the north star asks for."""
import os
from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str = None) -> Tuple[int, int, int]:
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun).  -> (rank, local_rank, world_size)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29512")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local, world


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of n_items for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rays(batch_xy: torch.Tensor, rank: int, world: int, multiple: int = 2) -> Tuple[torch.Tensor, int, int]:
    """Ray-shard [B,2,N_r] inside every item (used when B < world).  Slice bounds are rounded to `multiple`
    rays so every shard keeps whole 128-sample tiles.  -> (xy shard, lo, hi)."""
    n_r = batch_xy.shape[-1]
    lo, hi = shard_range(n_r // multiple, rank, world)
    lo, hi = lo * multiple, (hi * multiple if rank < world - 1 else n_r)
    return batch_xy[:, :, lo:hi].contiguous(), lo, hi


class GradBucket:
    """All parameter gradients live in ONE flat fp32 buffer (p.grad are views), so the training step needs a
    single memset and a single all-reduce — latency-bound on NVLink 5 (10.8–14 MB)."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce(self, average: bool = True):
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            if average:
                self.flat.div_(dist.get_world_size())
        return self.flat


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def max_over_ranks(value: float, device) -> float:
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
