"""Worker of tests/test_gpu_multi.py (launched with torch.distributed.run, one process per GPU, NCCL): the CUDA path under the
two partitions of SURVEY.md section 8e against the same step in ONE process.
  rays : B = 1 < world - every rank renders a slice of the item's rays, dist.gather_rays all-gathers the composited features
         (reduce-scatter backward), the consumer runs replicated, the flat gradient bucket is all-reduced and averaged;
  items: B = world - rank r renders item r, same bucket all-reduce.
Rank 0 also runs the unsharded step and writes the comparison to the JSON file named on the command line."""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import headnerf_oracle as O          # noqa: E402  (input / weight factory)


def main(out_path):
    hn = importlib.import_module("nerf-3dtalker-code_b200")
    rank, local, world = hn.dist.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    fs, S = 16, 64
    opt = O.OracleOptions(featmap_size=fs, pred_img_size=S)
    sd = O.formula_state_dict(opt, "init")
    lu = hn.HeadNeRFLossUtils(use_vgg_loss=False, device=dev)
    gen = torch.Generator().manual_seed(0)
    gt_all = torch.rand(world, 3, S, S, generator=gen).to(dev)
    mask_all = (torch.rand(world, 1, S, S, generator=gen) > 0.5).float().to(dev)
    inp_all = {k: v.to(dev) for k, v in O.synthetic_inputs(opt, world, seed=9).items()}

    def make_net():
        net = hn.HeadNeRFNet(hn.BaseOptions({"featmap_size": fs, "featmap_nc": 256, "pred_img_size": S}), False, False)
        net.load_state_dict(sd, strict=True)
        net = net.to(dev).eval()
        net.precision = "fast"
        early = [p for n, p in net.neural_render.named_parameters() if n != "bg_featmap"]
        bucket = hn.dist.GradBucket(net.parameters(), early=early)
        return net, bucket

    def step(net, bucket, items, reduce):
        x = {k: v[items] for k, v in inp_all.items()}
        bucket.zero()
        net.on_consumer_grads_ready = bucket.all_reduce_early if reduce else None
        out = net("test", x["batch_xy"], None, x["audiostyle"], None, x["shape_code"], x["appea_code"],
                  x["batch_Rmats"], x["batch_Tvecs"], x["batch_inv_inmats"])
        loss = lu.calc_total_loss(None, None, out, gt_all[items], mask_all[items], None)["total_loss"]
        loss.backward()
        if reduce:
            bucket.all_reduce(average=True)
        net.check_faults()
        return out["coarse_dict"]["merge_img"].detach().clone(), bucket.flat.clone(), float(loss)

    res = {"world": world}
    # ---- rays sharded inside one item
    net, bucket = make_net()
    net.set_ray_sharding(rank, world)
    img_s, flat_s, loss_s = step(net, bucket, slice(0, 1), True)
    net.set_ray_sharding()
    net.on_consumer_grads_ready = None
    img_1, flat_1, loss_1 = step(net, bucket, slice(0, 1), False)
    cos = float(torch.dot(flat_s.double(), flat_1.double()) / (flat_s.double().norm() * flat_1.double().norm()))
    res["rays"] = {"image_equal": bool(torch.equal(img_s, img_1)), "loss_equal": loss_s == loss_1, "grad_cosine": cos,
                   "grad_max_rel": float((flat_s - flat_1).abs().max() / flat_1.abs().max())}
    # ---- batch items sharded
    net, bucket = make_net()
    img_s, flat_s, loss_s = step(net, bucket, slice(rank, rank + 1), True)
    # one process, all items: the mean-reduced loss terms make the full-batch gradient the average of the per-item gradients only
    # when every item has the same number of head pixels; compare with the average of per-item single-process steps instead
    acc = torch.zeros_like(flat_s)
    for i in range(world):
        _, f, _ = step(net, bucket, slice(i, i + 1), False)
        acc += f
    acc /= world
    cos = float(torch.dot(flat_s.double(), acc.double()) / (flat_s.double().norm() * acc.double().norm()))
    res["items"] = {"grad_cosine": cos, "grad_max_rel": float((flat_s - acc).abs().max() / acc.abs().max())}
    hn.dist.barrier()
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(res, f)
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
