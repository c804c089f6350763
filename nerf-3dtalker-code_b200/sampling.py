"""Hierarchical resampling — drop-in for the reference's NetWorks/utils.py:164-265 `FineSample`.

Same constructor (`FineSample(opt)`), same call (`forward(batch_weight, coarse_sample_dict, disturb)`), same result dictionary
({"pts" [B,3,N_r,N_p], "dirs", "zvals" [B,1,N_r,N_p], "z_dists"}, N_p = num_sample_coarse + num_sample_fine); the per-ray pdf / cdf,
inverse-CDF lookup and the sort-merge with the coarse depths run in one kernel of libheadnerf_b200.so (hn_fine_sample, a warp per
ray).  Like the reference, the coarse weights are detached; train-mode uniforms are drawn here with torch.rand in the reference's
shape ([B*N_r, N_f+1]) so the generator stream is consumed identically.

HeadNeRFNet(hier_sampling=True) itself stays unsupported: the reference's fine pass is dead code that cannot run (it calls
calc_color_with_code with the wrong argument count, HeadNeRFNet.py:182-185) and would need 192-sample rays, which the fused
MLP kernels are not specialised for (32 / 64 / 128)."""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib as L
from . import ops


def fine_sample(weights, zvals, ray_o, ray_d, ray_l, n_fine, uniform=None, want_pts=True):
    """weights, zvals [B,N_r,N_c]; ray_o, ray_d [B,3,N_r]; ray_l [B,N_r]; uniform [B*N_r, n_fine+1] or None (linspace)
    -> zvals, z_dists [B,N_r,N_c+n_fine], pts [B,N_r,N_c+n_fine,3] (or None)."""
    lib = L.load()
    weights, zvals = ops._dev_f32(weights.detach(), "batch_weight"), ops._dev_f32(zvals, "zvals", weights.shape)
    B, n_r, n_c = weights.shape
    ray_o, ray_d = ops._dev_f32(ray_o, "batch_ray_o", (B, 3, n_r)), ops._dev_f32(ray_d, "batch_ray_d", (B, 3, n_r))
    ray_l = ops._dev_f32(ray_l, "batch_ray_l", (B, n_r))
    if uniform is not None:
        uniform = ops._dev_f32(uniform, "uniform_sample", (B * n_r, n_fine + 1))
    n_p = n_c + n_fine
    dev = weights.device
    out_z, out_d = torch.empty(B, n_r, n_p, device=dev), torch.empty(B, n_r, n_p, device=dev)
    out_p = torch.empty(B, n_r, n_p, 3, device=dev) if want_pts else None
    a = L.FineSample()
    a.n_rays_total, a.n_rays, a.n_coarse, a.n_fine = B * n_r, n_r, n_c, n_fine
    a.weights, a.zvals, a.uniform = ops._ptr(weights), ops._ptr(zvals), ops._ptr(uniform)
    a.ray_o, a.ray_d, a.ray_l = ops._ptr(ray_o), ops._ptr(ray_d), ops._ptr(ray_l)
    a.out_zvals, a.out_zdists, a.out_pts = ops._ptr(out_z), ops._ptr(out_d), ops._ptr(out_p)
    ops._call("hn_fine_sample", lib.hn_fine_sample, C.byref(a), ops._stream())
    return out_z, out_d, out_p


class FineSample(nn.Module):
    def __init__(self, opt) -> None:
        super().__init__()
        self.n_sample = opt.num_sample_fine + 1

    def forward(self, batch_weight, coarse_sample_dict, disturb):
        coarse_zvals = coarse_sample_dict["zvals"]                       # [B,1,N_r,N_c]
        B, _, n_r, n_c = coarse_zvals.shape
        uniform = None
        if disturb:                                                      # utils.py:233-234
            uniform = torch.rand(B * n_r, self.n_sample, device=batch_weight.device, dtype=batch_weight.dtype)
        ray_o, ray_d, ray_l = (coarse_sample_dict[k] for k in ("batch_ray_o", "batch_ray_d", "batch_ray_l"))
        z, zd, pts = fine_sample(batch_weight.reshape(B, n_r, n_c), coarse_zvals.reshape(B, n_r, n_c), ray_o.reshape(B, 3, n_r).contiguous(),
                                 ray_d.reshape(B, 3, n_r).contiguous(), ray_l.reshape(B, n_r).contiguous(), self.n_sample - 1, uniform)
        n_p = z.shape[-1]
        return {"pts": pts.permute(0, 3, 1, 2), "dirs": ray_d.reshape(B, 3, n_r, 1).expand(-1, -1, -1, n_p),
                "zvals": z.unsqueeze(1), "z_dists": zd.unsqueeze(1)}
