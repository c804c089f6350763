"""The oracle reproduces the committed outputs of the real reference (tests/golden/, made by
tests/golden/make_golden.py). Runs everywhere (no /root/reference, no GPU needed)."""
import pytest
import torch

from oracle import headnerf_oracle as O
from _util import GOLDEN_CASES, LEAVES, golden_loss, load_golden, probe_index


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_reproduces_reference_outputs(name):
    g = load_golden(name)
    opt, inp = g["opt"], g["inp"]
    sd = {k: v.requires_grad_(not k.endswith(".f")) for k, v in O.formula_state_dict(opt, g["variant"]).items()}
    x = {k: v.clone().requires_grad_(k in LEAVES) for k, v in inp.items()}
    res, r = O.headnerf_forward(sd, opt, g["mode"], x["batch_xy"], x["audiostyle"], x["shape_code"], x["appea_code"],
                                x["batch_Rmats"], x["batch_Tvecs"], x["batch_inv_inmats"], t_rand=x.get("t_rand"))
    # fp32 on possibly different host CPUs: rounding-level tolerance, far below the 1e-3 product gate
    for k in ("F", "bg_alpha", "depth"):
        assert torch.allclose(r[k], g["out"][k], rtol=1e-4, atol=2e-5), k
    for k in ("merge_img", "bg_img"):
        assert torch.allclose(res["coarse_dict"][k], g["out"][k], rtol=1e-4, atol=1e-5), k
    golden_loss(res["coarse_dict"]["merge_img"]).backward()
    for k in LEAVES:
        ref = g["grads"][k]
        assert torch.allclose(x[k].grad, ref, rtol=2e-3, atol=1e-5 * float(ref.abs().max())), k
    for k, nrm in g["pnorm"].items():
        got = sd[k].grad.reshape(-1)
        assert abs(float(got.double().norm()) - nrm) <= 2e-3 * nrm + 1e-12, k
        assert torch.allclose(got[probe_index(got.numel())], g["pprobe"][k], rtol=5e-3, atol=1e-4 * nrm / max(got.numel(), 1) ** 0.5 + 1e-12), k


def test_formula_weights_are_deterministic():
    opt = O.OracleOptions(featmap_size=8, pred_img_size=32)
    a, b = O.formula_state_dict(opt, "trained"), O.formula_state_dict(opt, "trained")
    assert list(a) == list(O.state_dict_shapes(opt))
    assert all(torch.equal(a[k], b[k]) for k in a)
    n = sum(v.numel() for k, v in a.items() if not k.endswith(".f"))
    assert n > 2_000_000
