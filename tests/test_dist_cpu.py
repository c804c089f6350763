"""Multi-process (world_size 2, gloo, CPU) tests of the data-parallel plumbing used at N>1 GPUs: contiguous ray / item
sharding and the single flat-bucket gradient all-reduce."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, ret):
    import importlib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    hn = importlib.import_module("nerf-3dtalker-code_b200")
    r, l, w = hn.dist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    lin = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3))
    bucket = hn.dist.GradBucket(lin.parameters())
    assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in lin.parameters())
    xs = torch.arange(8 * 7, dtype=torch.float32).view(8, 7) / 10.0
    lo, hi = hn.dist.shard_range(8, rank, world)
    bucket.zero()
    lin(xs[lo:hi]).pow(2).sum().backward()                 # accumulates INTO the bucket views
    bucket.all_reduce(average=False)
    flat = bucket.compact()
    # reference: single-process gradient over the full batch
    ref = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3))
    ref.load_state_dict(lin.state_dict())
    ref(xs).pow(2).sum().backward()
    ref_flat = torch.cat([p.grad.reshape(-1) for p in ref.parameters()])
    ok = torch.allclose(flat, ref_flat, rtol=1e-5, atol=1e-6)
    mx = hn.dist.max_over_ranks(float(rank + 1), torch.device("cpu"))
    xy = torch.arange(2 * 2 * 10, dtype=torch.float32).view(2, 2, 10)
    shard, a, b = hn.dist.shard_rays(xy, rank, world, multiple=2)
    ret[rank] = (ok, mx, a, b, shard.shape[-1])
    hn.dist.barrier()
    dist.destroy_process_group()


def test_flat_bucket_allreduce_and_sharding_world2():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret[0][0] and ret[1][0], "all-reduced flat gradient differs from the single-process gradient"
    assert ret[0][1] == 2.0 and ret[1][1] == 2.0
    assert (ret[0][2], ret[0][3], ret[1][2], ret[1][3]) == (0, 6, 6, 10)      # 10 rays = 5 ray pairs -> 3 + 2 pairs
    assert ret[0][4] + ret[1][4] == 10


def test_shard_range_is_a_partition():
    import importlib
    hn = importlib.import_module("nerf-3dtalker-code_b200")
    for n in (1, 2, 7, 8, 4096):
        for world in (1, 2, 3, 8):
            spans = [hn.dist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _gather_worker(rank, world, port, ret):
    import importlib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    hn = importlib.import_module("nerf-3dtalker-code_b200")
    hn.dist.init_from_env(backend="gloo")
    B, n_rays, Cc = 2, 10, 5                                # 10 rays = 5 pairs -> shards of 6 and 4 rays (ragged: padded gather)
    g = torch.Generator().manual_seed(0)
    F_full = torch.randn(B, n_rays, Cc, generator=g)
    bg_full = torch.randn(B, n_rays, generator=g)
    w = torch.randn(B, n_rays, Cc + 1, generator=g)         # a replicated "consumer": loss = sum(w * [F | bg])
    xy = torch.zeros(B, 2, n_rays)
    _, lo, hi = hn.dist.shard_rays(xy, rank, world)
    F_loc = F_full[:, lo:hi].clone().requires_grad_(True)
    bg_loc = bg_full[:, lo:hi].clone().requires_grad_(True)
    Fg, bgg = hn.dist.gather_rays(F_loc, bg_loc, n_rays, rank, world)
    ok_fwd = torch.equal(Fg, F_full) and torch.equal(bgg, bg_full)
    loss = (Fg * w[..., :Cc]).sum() + (bgg * w[..., Cc]).sum()
    loss.backward()
    # every rank ran the same consumer: reduce-scatter(sum) gives world x the single-process slice gradient; the 1/world of the
    # parameter all-reduce (GradBucket.all_reduce(average=True) / FusedAdam grad_scale) brings it back
    ok_bwd = torch.allclose(F_loc.grad / world, w[:, lo:hi, :Cc]) and torch.allclose(bg_loc.grad / world, w[:, lo:hi, Cc])
    ret[rank] = (ok_fwd, ok_bwd, lo, hi, hn.dist.ray_shard_sizes(n_rays, world))
    hn.dist.barrier()
    dist.destroy_process_group()


def test_gather_rays_forward_and_reduce_scatter_backward_world2():
    """SURVEY.md section 8e collective 2 on gloo: all-gather of ragged ray shards forward, reduce-scatter of the gradient backward."""
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_gather_worker, args=(world, port, ret), nprocs=world, join=True)
    for r in range(world):
        assert ret[r][0], "gathered feature map differs from the unsharded one"
        assert ret[r][1], "reduce-scattered gradient differs from the single-process slice gradient"
    assert (ret[0][2], ret[0][3], ret[1][2], ret[1][3]) == (0, 6, 6, 10) and ret[0][4] == [6, 4]


def test_grad_bucket_early_range_layout():
    import importlib
    hn = importlib.import_module("nerf-3dtalker-code_b200")
    a, b, c = (torch.nn.Parameter(torch.randn(n)) for n in (3, 5, 7))
    bucket = hn.dist.GradBucket([a, b, c], early=[c])
    assert bucket.params[0] is c and bucket.n_early == 64 and bucket.offsets == [0, 64, 128] and bucket.flat.numel() == 192
    assert c.grad.data_ptr() == bucket.flat.data_ptr() and bucket.compact().numel() == 15
    assert all((p.grad.data_ptr() - bucket.flat.data_ptr()) % 256 == 0 for p in (a, b, c))   # 256-byte boundaries inside the buffer (CUDA allocations are 512-byte aligned)
    bucket.all_reduce_early(); bucket.all_reduce()              # no process group: no-ops
