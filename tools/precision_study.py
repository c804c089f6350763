"""CPU study (fp32 oracle + emulated fp16 rounding, straight-through) of where the camera-gradient deviation comes from.
Result (profiles/r01_precision_study.md): backward rounding contributes nothing; any forward-side rounding of weights or
activations to 11 significant bits - even with exact arithmetic afterwards - moves the Rmats/Tvecs gradient cosine to ~0.998."""
import sys, torch, torch.nn.functional as F
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import headnerf_oracle as O
from _util import LEAVES, cosine, golden_loss, load_golden
torch.set_num_threads(8)
def r16(t):  # straight-through fp16 rounding
    return t + (t.half().float() - t).detach()
def mlp_emul(sd, audio, vps, vds, fwd_round, bwd_round, pe63, R_PE=True, R_WPE=True, R_W=True, R_ACT=True):
    # vps = cat([pe(63), shape]); emulate: PE rounded, hidden activations rounded, weights rounded; backward grads rounded via hooks
    W = lambda n: (r16(sd[n + ".weight"]) if fwd_round else sd[n + ".weight"])
    def hook(t):
        if bwd_round and t.requires_grad:
            t.register_hook(lambda g: (g * 4096).half().float() / 4096)
        return t
    p = "fg_CD_predictor."
    pe, shp = vps[:, :63], vps[:, 63:]
    pe_r = r16(pe) if (fwd_round and R_PE) else pe
    # per-sample part uses rounded weights/acts; latent part stays fp32 (folded bias)
    def layer(name, xs_rounded, xs_exact):
        w = sd[name + ".weight"]; b = sd[name + ".bias"]
        cols = 0; out = 0
        for x, rounded in xs_rounded + xs_exact:
            c = x.shape[1]
            wpart = w[:, cols:cols + c]
            is_pe = (x is pe_r)
            do = rounded and fwd_round and (R_WPE if is_pe else R_W)
            out = out + F.conv2d(x, (r16(wpart) if do else wpart))
            cols += c
        return out + b.view(1, -1, 1, 1)
    x = F.relu(layer(p + "FeaExt_module_0", [(pe_r, True)], [(shp, False), (audio, False)]))
    x = hook(r16(x) if (fwd_round and R_ACT) else x)
    for i in range(1, 8):
        if i == 5:
            x = F.relu(layer(p + "FeaExt_module_5", [(pe_r, True)], [(shp, False), (x, True)]))
        else:
            x = F.relu(layer(p + f"FeaExt_module_{i}", [(x, True)], []))
        x = hook(r16(x) if (fwd_round and R_ACT) else x)
    density = F.relu(F.conv2d(x, sd[p + "density_module.weight"], sd[p + "density_module.bias"]))
    y = layer(p + "RGB_layer_0", [(x, True)], [])
    y = hook(r16(y) if (fwd_round and R_ACT) else y)
    y = F.relu(layer(p + "RGB_layer_1", [(y, True)], [(vds, False)]))
    y = hook(r16(y) if (fwd_round and R_ACT) else y)
    feat = layer(p + "RGB_layer_2", [(y, True)], [])
    return feat, density
for name in ("fs8_test_init", "fs16_test_trained"):
    g = load_golden(name); opt = g["opt"]
    def run(fwd_round, bwd_round, **kw):
        sd = O.formula_state_dict(opt, g["variant"])
        x = {k: v.clone().requires_grad_(k in LEAVES) for k, v in g["inp"].items()}
        B, _, n_r = x["batch_xy"].shape; ns = 64
        ro, rd, rl = O.gen_rays(x["batch_xy"], x["batch_Rmats"], x["batch_Tvecs"], x["batch_inv_inmats"])
        smp = O.sample_points(ro, rd, rl, opt, False)
        pe = O.positional_encoding(smp["pts"])
        ex = lambda c: c.unsqueeze(-1).unsqueeze(-1).expand(-1, -1, n_r, ns)
        vps = torch.cat([pe, ex(x["shape_code"])], 1)
        feat, dens = mlp_emul(sd, ex(x["audiostyle"]), vps, ex(x["appea_code"]), fwd_round, bwd_round, None, **kw)
        Fm, bga, _, _ = O.composite(feat, dens, smp["z_dists"], smp["zvals"])
        fs = opt.featmap_size
        merge = Fm.view(B, 256, fs, fs) + bga.view(B, 1, fs, fs) * sd["neural_render.bg_featmap"]
        golden_loss(O.neural_render(sd, merge, opt)).backward()
        return {k: x[k].grad for k in LEAVES}
    ref = run(False, False)
    for label, kw in (("PE operand only", dict(R_PE=True, R_WPE=False, R_W=False, R_ACT=False)),
                      ("PE-column weights only", dict(R_PE=False, R_WPE=True, R_W=False, R_ACT=False)),
                      ("other weights only", dict(R_PE=False, R_WPE=False, R_W=True, R_ACT=False)),
                      ("hidden activations only", dict(R_PE=False, R_WPE=False, R_W=False, R_ACT=True)),
                      ("all but PE operand+PE weights", dict(R_PE=False, R_WPE=False, R_W=True, R_ACT=True))):
        got = run(True, False, **kw)
        print(name, f"{label:32s}", " ".join(f"{k}={cosine(got[k], ref[k]):.5f}" for k in ("batch_Rmats", "batch_Tvecs")), flush=True)
