"""Multi-GPU plumbing for the hot path: one process per GPU (torchrun), rays / batch items sharded across
ranks (every ray is independent, weights are replicated), and ONE NCCL all-reduce per training step over a
single flat fp32 gradient bucket (SURVEY.md §8e collective 1) - optionally as two ranges, the consumer's
gradients being reduced while the MLP backward still runs.  When rays are sharded INSIDE an item (B < world)
and the NeuralRenderer consumer needs whole feature maps, `gather_rays` is §8e's collective 2: an all-gather of
the [B, N_r/G, 257] slices (256 features + bg_alpha) forward, a reduce-scatter of their gradient backward.
The reference trains on a single GPU (talker_trainer.py:704-712); this is the data-parallel extension the
north star asks for."""
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


FLAT_ALIGN = 64          # floats: every tensor of a flat buffer starts on a 256-byte boundary (vector loads, bulk copies, library GEMMs)


def flat_layout(tensors, align: int = FLAT_ALIGN):
    """Offsets (in elements) of `tensors` laid end to end with every start rounded up to `align`.  -> (offsets, total)."""
    offs, off = [], 0
    for t in tensors:
        offs.append(off)
        off += t.numel()
        off += (-off) % align
    return offs, off


class GradBucket:
    """All parameter gradients live in ONE flat fp32 buffer (p.grad are views), so the training step needs a
    single memset and a single all-reduce — latency-bound on NVLink 5 (10.8–14 MB).
    `early`: parameters whose gradients are complete before the rest (the NeuralRenderer consumer's, ready when
    the backward pass reaches the feature map); they are laid out first so that `all_reduce_early()` can reduce
    that contiguous range asynchronously while the MLP backward still runs (`all_reduce()` then covers the rest)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], early: Iterable[torch.nn.Parameter] = ()):
        early = [p for p in early if p.requires_grad]
        ids = {id(p) for p in early}
        rest = [p for p in params if p.requires_grad and id(p) not in ids]
        self.params: List[torch.nn.Parameter] = early + rest
        self.offsets, n = flat_layout(self.params)
        self.n_early = self.offsets[len(early)] if (early and rest) else (n if early else 0)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)      # the alignment gaps stay zero
        for p, off in zip(self.params, self.offsets):
            p.grad = self.flat[off:off + p.numel()].view_as(p)
        self._early_work = None

    def compact(self):
        """The gradients without the alignment gaps, concatenated in bucket order (tests, logging)."""
        return torch.cat([self.flat[o:o + p.numel()] for p, o in zip(self.params, self.offsets)])

    def zero(self):
        self.flat.zero_()

    def all_reduce_early(self):
        """Start the all-reduce (sum) of the early range on the collective's own stream; returns at once."""
        if self.n_early and dist.is_initialized() and dist.get_world_size() > 1:
            self._early_work = dist.all_reduce(self.flat[:self.n_early], op=dist.ReduceOp.SUM, async_op=True)

    def all_reduce(self, average: bool = True):
        """Sum over ranks of whatever all_reduce_early() has not covered, then wait for both.  average=False leaves the sum
        (train.FusedAdam folds the 1/world into its kernel)."""
        if dist.is_initialized() and dist.get_world_size() > 1:
            if self._early_work is not None:
                dist.all_reduce(self.flat[self.n_early:], op=dist.ReduceOp.SUM)
                self._early_work.wait()
                self._early_work = None
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            if average:
                self.flat.div_(dist.get_world_size())
        return self.flat


class _GatherRays(torch.autograd.Function):
    """x [B, n_local, C] (this rank's rays of every item) -> [B, sum n_local, C] on every rank.  Backward: reduce-scatter (sum) of
    the full gradient - each rank keeps the sum over ranks of its own slice (with a replicated consumer and the usual 1/world
    averaging of the parameter all-reduce this is exactly the single-process gradient)."""

    @staticmethod
    def forward(ctx, x, sizes, rank, group):
        world = len(sizes)
        n_max = max(sizes)
        B, n_loc, Cc = x.shape
        assert n_loc == sizes[rank]
        send = x if n_loc == n_max else torch.cat([x, x.new_zeros(B, n_max - n_loc, Cc)], dim=1)
        recv = x.new_empty(world * B, n_max, Cc)                    # rank-major concatenation along dim 0
        dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
        recv = recv.view(world, B, n_max, Cc)
        ctx.sizes, ctx.rank, ctx.group = sizes, rank, group
        return torch.cat([recv[r, :, :sizes[r]] for r in range(world)], dim=1)

    @staticmethod
    def backward(ctx, g):
        sizes, rank, group = ctx.sizes, ctx.rank, ctx.group
        world, n_max = len(sizes), max(sizes)
        B, _, Cc = g.shape
        send = g.new_zeros(world, B, n_max, Cc)
        off = 0
        for r in range(world):
            send[r, :, :sizes[r]] = g[:, off:off + sizes[r]]
            off += sizes[r]
        if dist.get_backend(group) == "nccl":
            out = g.new_empty(B, n_max, Cc)
            dist.reduce_scatter_tensor(out, send.view(world * B, n_max, Cc), op=dist.ReduceOp.SUM, group=group)
        else:                                                    # gloo (CPU tests) has no reduce-scatter: all-reduce, keep the own slice
            dist.all_reduce(send, op=dist.ReduceOp.SUM, group=group)
            out = send[rank]
        return out[:, :sizes[rank]].contiguous(), None, None, None


def ray_shard_sizes(n_rays: int, world: int, multiple: int = 2) -> List[int]:
    """Rays per rank under shard_rays' partition."""
    sizes = []
    for r in range(world):
        lo, hi = shard_range(n_rays // multiple, r, world)
        sizes.append((hi * multiple if r < world - 1 else n_rays) - lo * multiple)
    return sizes


def gather_rays(Fm: torch.Tensor, bg: torch.Tensor, n_rays: int, rank: int, world: int, group=None):
    """§8e collective 2: this rank's F [B, n_local, 256] and bg_alpha [B, n_local] -> the full [B, n_rays, 256] / [B, n_rays] on every
    rank, as ONE all-gather of [B, n_local, 257] slices (differentiable: reduce-scatter backward)."""
    sizes = ray_shard_sizes(n_rays, world)
    both = _GatherRays.apply(torch.cat([Fm, bg.unsqueeze(-1)], dim=-1), sizes, rank, group)
    return both[..., :-1], both[..., -1]


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def max_over_ranks(value: float, device) -> float:
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
