// hn_render.cu — the whole hot path behind ONE call per direction ("a fused hn_render_fwd / hn_render_bwd
// once the kernels are chained"): a C or C++ host needs no knowledge of the intermediate buffers' order of use.
//   hn_render_fwd : latent folding -> sampling + positional encoding + fg_CD_predictor -> alpha compositing
//   hn_render_bwd : loss scale -> compositing backward -> MLP data gradients -> MLP weight gradients -> folding backward
//                   (-> ray set-up backward when camera gradients are requested)
// Every buffer is caller-owned (sizes from the hn_*_bytes queries); all launches go to `stream` in order, nothing synchronises.
// NetWorks/HeadNeRFNet.py:123-160 (calc_color_with_code up to the composited features) and its autograd.
#include "hn_api.h"

extern "C" int hn_render_fwd(const hn_render_fwd_t* a, void* stream) {
    using namespace hn;
    if (!a || !a->bias_eff || !a->feat || !a->sigma || !a->delta || !a->F || !a->bg_alpha || !a->status)
        return set_error(HN_E_BADARG, "hn_render_fwd: null pointer");
    if (a->fold.B != a->cam.B) return set_error(HN_E_BADARG, "hn_render_fwd: fold.B != cam.B");
    hn_fold_t fold = a->fold;
    fold.r0_fused = 1;                                              // the fused chains skip RGB_layer_0 (hn_mlp_sched.h)
    if (int rc = hn_fold_bias(&fold, a->bias_eff, stream)) return rc;
    hn_mlp_fwd_t m{};
    m.cam = a->cam; m.bias = a->bias_eff; m.w_density = a->w_density; m.packed = a->packed;
    m.feat = a->feat; m.sigma = a->sigma; m.delta = a->delta; m.zvals = nullptr;
    m.act = a->act; m.masks = a->masks; m.status = a->status;
    if (int rc = hn_mlp_fwd(&m, stream)) return rc;
    hn_composite_fwd_t c{};
    c.n_rays_total = a->cam.B * a->cam.n_rays; c.n_samples = a->cam.n_samples; c.C = HN_FEAT;
    c.feat = a->feat; c.sigma = a->sigma; c.delta = a->delta; c.zvals = nullptr;
    c.F = a->F; c.bg_alpha = a->bg_alpha; c.depth = nullptr; c.weights = nullptr;
    return hn_composite_fwd(&c, stream);
}

extern "C" int hn_render_bwd(const hn_render_bwd_t* a, void* stream) {
    using namespace hn;
    if (!a || !a->feat || !a->sigma || !a->delta || !a->act || !a->masks || !a->gF || !a->g_bg || !a->dfeat_image || !a->dsigma ||
        !a->scale || !a->scale_scratch8 || !a->status)
        return set_error(HN_E_BADARG, "hn_render_bwd: null pointer");
    if (a->fold.B != a->cam.B) return set_error(HN_E_BADARG, "hn_render_bwd: fold.B != cam.B");
    const bool need_cam = a->dR || a->dT || a->dKinv;
    if (need_cam && (!a->g_ray_o || !a->g_ray_v || !a->g_ray_l || !a->ddelta))
        return set_error(HN_E_BADARG, "hn_render_bwd: camera gradients need g_ray_o / g_ray_v / g_ray_l / ddelta workspaces");
    bool need_w = false;
    for (int i = 0; i < 12; ++i) need_w = need_w || (a->dw[i] != nullptr);
    bool need_fold = a->fold_grads.dshape || a->fold_grads.daudio || a->fold_grads.dappea || a->fold_grads.dw0 || a->fold_grads.dw5 || a->fold_grads.dwr1;
    for (int i = 0; i < 12; ++i) need_fold = need_fold || (a->fold_grads.dbias[i] != nullptr);
    const bool save = need_w || need_fold;
    if (save && (!a->grads || !a->dbias_eff || !a->items_workspace))
        return set_error(HN_E_BADARG, "hn_render_bwd: parameter / code gradients need grads, dbias_eff and items_workspace");
    const int rays = a->cam.B * a->cam.n_rays;
    if (int rc = hn_loss_scale(a->gF, (int64_t)rays * HN_FEAT, a->grad_target > 0.f ? a->grad_target : 64.f, a->scale, a->scale_scratch8, stream)) return rc;
    hn_composite_bwd_t c{};
    c.n_rays_total = rays; c.n_samples = a->cam.n_samples; c.C = HN_FEAT;
    c.feat = a->feat; c.sigma = a->sigma; c.delta = a->delta; c.zvals = nullptr;
    c.gF = a->gF; c.g_bg = a->g_bg; c.g_depth = nullptr;
    c.dfeat = nullptr; c.dfeat_image = a->dfeat_image; c.grad_scale = a->scale; c.dsigma = a->dsigma; c.ddelta = need_cam ? a->ddelta : nullptr;
    if (int rc = hn_composite_bwd(&c, stream)) return rc;
    hn_mlp_bwd_data_t d{};
    d.cam = a->cam; d.packed = a->packed; d.w_density = a->w_density; d.dfeat_image = a->dfeat_image;
    d.dsigma = a->dsigma; d.ddelta = need_cam ? a->ddelta : nullptr; d.sigma = a->sigma; d.grad_scale = a->scale;
    d.masks = a->masks; d.act = a->act; d.grads = save ? a->grads : nullptr;
    d.g_ray_o = need_cam ? a->g_ray_o : nullptr; d.g_ray_v = need_cam ? a->g_ray_v : nullptr; d.g_ray_l = need_cam ? a->g_ray_l : nullptr;
    d.status = a->status;
    if (int rc = hn_mlp_bwd_data(&d, stream)) return rc;
    if (save) {
        hn_mlp_bwd_weights_t w{};
        w.B = a->cam.B; w.n_rays = a->cam.n_rays; w.n_samples = a->cam.n_samples;
        w.act = a->act; w.grads = a->grads; w.dfeat_image = a->dfeat_image; w.grad_scale = a->scale;
        for (int i = 0; i < 12; ++i) { w.dw[i] = a->dw[i]; w.ld[i] = a->ld[i]; }
        // RGB_layer_0 is folded into RGB_layer_1: their weight gradients come from dL/dW_f (hn_unfuse_r0r1 below)
        const bool want_r = a->dw[9] || a->dw[10];
        if (want_r && (!a->dwf || !a->fold.wr0))
            return set_error(HN_E_BADARG, "hn_render_bwd: RGB_layer_0 / _1 weight gradients need the dwf workspace and fold.wr0");
        w.r0_fused = 1; w.dwf = want_r ? a->dwf : nullptr;
        w.dw[9] = nullptr; w.dw[10] = want_r ? a->dwf : nullptr;     // (non-NULL marks "wanted"; the kernel writes to dwf)
        w.l5_hidden_col = a->l5_hidden_col; w.dbias = a->dbias_eff;
        w.items_workspace = a->items_workspace; w.items_workspace_bytes = a->items_workspace_bytes; w.status = a->status;
        for (int i = 0; i < 12; ++i)
            if (i != 0 && i != 5 && i != 10 && a->fold_grads.dbias[i]) w.want_all_bias = 1;   // beyond the latent-folded layers (FeaExt_module_0, _5, RGB_layer_1)
        if (int rc = hn_mlp_bwd_weights(&w, stream)) return rc;
        if (want_r || a->fold_grads.dbias[9]) {
            hn_unfuse_t u{};
            u.B = a->cam.B; u.wr0 = a->fold.wr0; u.ldr0 = a->fold.ldr0; u.wr1 = a->fold.wr1; u.ldr1 = a->fold.ldr1; u.b_r0 = a->fold.bias[9];
            u.dwf = a->dwf; u.dbias_eff = a->dbias_eff; u.dwr0 = a->dw[9]; u.dwr1 = a->dw[10];
            if (!u.wr0 || !u.dwf) return set_error(HN_E_BADARG, "hn_render_bwd: RGB_layer_0's bias gradient needs fold.wr0 and the dwf workspace");
            if (int rc = hn_unfuse_r0r1(&u, stream)) return rc;
        }
        if (need_fold) {
            hn_fold_t fold = a->fold;
            fold.r0_fused = 1;
            if (int rc = hn_fold_bias_bwd(&fold, a->dbias_eff, &a->fold_grads, stream)) return rc;
        }
    }
    if (need_cam)
        if (int rc = hn_camera_bwd(&a->cam, a->g_ray_o, a->g_ray_v, a->g_ray_l, a->dR, a->dT, a->dKinv, stream)) return rc;
    return HN_OK;
}
