"""GPU parity of hierarchical resampling: hn_fine_sample through the drop-in FineSample module against the
oracle's restatement of NetWorks/utils.py:164-265 (pinned bit-equal to the reference's own FineSample)."""
import pytest
import torch

from oracle import headnerf_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("disturb,n_coarse,n_fine,seed", [(False, 64, 128, 0), (True, 64, 128, 1), (True, 32, 64, 2), (False, 128, 96, 3)])
def test_fine_sample_matches_oracle(hn, disturb, n_coarse, n_fine, seed):
    opt = O.OracleOptions(featmap_size=8, pred_img_size=32, num_sample_coarse=n_coarse)
    B = 2
    inp = O.synthetic_inputs(opt, B, seed=seed, n_rays=77, jitter=True)
    ro, rd, rl = O.gen_rays(inp["batch_xy"], inp["batch_Rmats"], inp["batch_Tvecs"], inp["batch_inv_inmats"])
    coarse = O.sample_points(ro, rd, rl, opt, True, inp["t_rand"])
    g = torch.Generator().manual_seed(10 + seed)
    dens = torch.relu(torch.randn(B, 1, 77, n_coarse, generator=g) * 8 - 2)
    dens[:, :, ::5] = 0.0                                            # empty rays: pdf = 0 everywhere, the 1e-5 guards decide
    feat = torch.randn(B, 3, 77, n_coarse, generator=g)
    _, _, _, w = O.composite(feat, dens, coarse["z_dists"], coarse["zvals"])
    u = torch.rand(B * 77, n_fine + 1, generator=g) if disturb else None
    ref = O.fine_sample(w, coarse, n_fine, disturb, uniform=u)

    class Opt:
        num_sample_fine = n_fine
    fs = hn.FineSample(Opt())
    coarse_c = {k: v.to(DEV) for k, v in coarse.items()}
    if disturb:                                                      # the module draws its own uniforms; feed the oracle's through the operator
        z, zd, pts = hn.sampling.fine_sample(w.reshape(B, 77, n_coarse).to(DEV), coarse_c["zvals"].reshape(B, 77, n_coarse).contiguous(),
                                             ro.contiguous().to(DEV), rd.contiguous().to(DEV), rl.reshape(B, 77).contiguous().to(DEV), n_fine, u.to(DEV))
        out = {"zvals": z.unsqueeze(1), "z_dists": zd.unsqueeze(1), "pts": pts.permute(0, 3, 1, 2)}
        res = fs(w.to(DEV), coarse_c, True)                          # shape / key contract of the module in train mode
        assert set(res.keys()) == {"pts", "dirs", "zvals", "z_dists"} and res["zvals"].shape == (B, 1, 77, n_coarse + n_fine)
        assert bool((res["zvals"][..., 1:] >= res["zvals"][..., :-1]).all())
    else:
        out = fs(w.to(DEV), coarse_c, False)
        assert torch.equal(out["dirs"].cpu(), ref["dirs"])
    n_p = n_coarse + n_fine
    assert out["zvals"].shape == (B, 1, 77, n_p) and out["pts"].shape == (B, 3, 77, n_p)
    # a depth sits within the fp32 conditioning of the reference formula: t = (u - cdf[below]) / (cdf[above] - cdf[below]) divides by a
    # cdf step as small as 1e-5, so the ~1e-7 difference between two summation orders of the cdf (torch.cumsum vs the warp scan)
    # moves t by up to 1e-2 of a bin (0.1 deep): 1e-3 worst case, 6e-5 observed; a uniform on a bin boundary may interpolate in
    # the neighbouring bin - the inverse CDF is continuous there
    assert (out["zvals"].cpu() - ref["zvals"]).abs().max() < 3e-4
    assert (out["z_dists"].cpu() - ref["z_dists"]).abs().max() < 6e-4
    assert (out["pts"].cpu() - ref["pts"]).abs().max() < 6e-4
    assert (out["zvals"].cpu() - ref["zvals"]).abs().mean() < 2e-6                # ... and almost every depth agrees to rounding
    assert bool((out["zvals"][..., 1:] >= out["zvals"][..., :-1]).all())      # sortedness (size-independent property)


def test_fine_sample_full_size_properties(hn):
    """Reso64 batch 2 (8192 rays): sorted depths, the coarse depths are a subset, z_dists telescopes to the ray's depth range."""
    opt = O.OracleOptions(featmap_size=64, pred_img_size=512)
    B, n_r = 2, 4096
    inp = {k: v.to(DEV) for k, v in O.synthetic_inputs(opt, B, seed=0).items()}
    ro, rd, rl = O.gen_rays(inp["batch_xy"], inp["batch_Rmats"], inp["batch_Tvecs"], inp["batch_inv_inmats"])
    coarse = O.sample_points(ro, rd, rl, opt, False)
    w = torch.rand(B, 1, n_r, 64, device=DEV) ** 4

    class Opt:
        num_sample_fine = 128
    out = hn.FineSample(Opt())(w, coarse, False)
    z = out["zvals"][:, 0]
    assert bool((z[..., 1:] >= z[..., :-1]).all())
    zc = coarse["zvals"][:, 0]
    pos = torch.searchsorted(z.contiguous(), zc.contiguous())
    assert bool((torch.gather(z, -1, pos.clamp(max=z.shape[-1] - 1)) == zc)[..., :-1].all())     # every coarse depth survives the merge
    total = out["z_dists"][:, 0].sum(-1) / rl[:, 0]
    assert (total - (zc[..., -1] - zc[..., 0])).abs().max() < 1e-3


def test_hier_sampling_flag_is_refused_with_a_reason(hn):
    with pytest.raises(NotImplementedError, match="hier_sampling"):
        hn.HeadNeRFNet(hn.BaseOptions({"featmap_size": 8, "featmap_nc": 256, "pred_img_size": 32}), False, True)
