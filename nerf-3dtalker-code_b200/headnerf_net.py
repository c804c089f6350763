"""HeadNeRFNet — drop-in for the reference's NetWorks/HeadNeRFNet.py:10-207.

Same constructor, same forward signature, same state-dict keys/shapes (`fg_CD_predictor.*`,
`neural_render.*`), same seeded initialisation; the rendering hot path (sampling -> positional encoding ->
fg_CD_predictor -> compositing) runs in libheadnerf_b200.so instead of eager torch ops.  The latent codes
(shape/expression(+gaze), audio style, appearance) never get broadcast to [B,C,N_r,N_s] (HeadNeRFNet.py:149-152):
their weight columns are folded into one effective bias row per batch item (three tiny matmuls that stay in
autograd, so code gradients and the latent weight-column gradients come from the kernel's bias gradient)."""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L
from . import ops
from .neural_renderer import NeuralRenderer
from .options import BaseOptions


class MLPforNeRF(nn.Module):
    """Parameter container with the reference's layer names and init (NetWorks/models.py:29-59).
    All layers are 1x1 convolutions = per-sample linear maps; weights stay fp32 [out,in,1,1]."""

    def __init__(self, vp_channels, vd_channels, n_layers=8, h_channel=256, res_nfeat=3):
        super().__init__()
        self.vp_channels, self.vd_channels, self.n_layers = vp_channels, vd_channels, n_layers
        self.h_channel, self.res_nfeat, self.skips = h_channel, res_nfeat, [n_layers // 2]

        def conv(cin, cout, xavier):
            m = nn.Conv2d(cin, cout, kernel_size=1, stride=1, padding=0)
            if xavier:
                nn.init.xavier_uniform_(m.weight.data)
            return m

        self.add_module("FeaExt_module_0", conv(vp_channels + 64, h_channel, False))
        for i in range(n_layers - 1):
            cin = h_channel + vp_channels if i in self.skips else h_channel
            self.add_module("FeaExt_module_%d" % (i + 1), conv(cin, h_channel, True))
        self.add_module("density_module", conv(h_channel, 1, True))
        self.density_module.bias.data[:] = 0.0
        self.add_module("RGB_layer_0", conv(h_channel, h_channel, True))
        self.add_module("RGB_layer_1", conv(h_channel + vd_channels, h_channel // 2, False))
        self.add_module("RGB_layer_2", conv(h_channel // 2, res_nfeat, False))

    def layers(self):
        names = ["FeaExt_module_%d" % i for i in range(8)] + ["density_module", "RGB_layer_0", "RGB_layer_1", "RGB_layer_2"]
        return [self._modules[n] for n in names]

    def forward(self, *a, **k):
        raise RuntimeError("fg_CD_predictor runs inside the fused CUDA render operator; call HeadNeRFNet.forward")


class HeadNeRFNet(nn.Module):
    def __init__(self, opt: BaseOptions, include_vd, hier_sampling, include_gaze=False, eye_gaze_dim=2) -> None:
        super().__init__()
        if include_vd:
            raise NotImplementedError("include_vd=True (view-direction embedding) is not part of the accelerated path; "
                                      "every caller in the reference passes False (talker_trainer.py:693,699)")
        if hier_sampling:
            raise NotImplementedError("hier_sampling=True is dead code in the reference (HeadNeRFNet.py:182-185 passes the "
                                      "wrong argument count); not supported")
        self.hier_sampling, self.include_vd = hier_sampling, include_vd
        self.include_gaze, self.eye_gaze_dim = include_gaze, eye_gaze_dim
        self.opt = opt
        self.num_sample_coarse = opt.num_sample_coarse
        self.mlp_h_channel = opt.mlp_hidden_nchannels
        self.featmap_size, self.featmap_nc, self.pred_img_size = opt.featmap_size, opt.featmap_nc, opt.pred_img_size
        if self.mlp_h_channel != L.HIDDEN or self.featmap_nc != L.FEAT:
            raise NotImplementedError(f"the sm_100a kernels are specialised for mlp_hidden_nchannels={L.HIDDEN}, "
                                      f"featmap_nc={L.FEAT} (all shipped HeadNeRF checkpoints)")
        self.vp_n_freqs = 10
        shape_dims = opt.iden_code_dims + opt.expr_code_dims + (eye_gaze_dim if include_gaze else 0)
        vp_channels = shape_dims + self.vp_n_freqs * 6 + 3
        vd_channels = opt.text_code_dims + opt.illu_code_dims
        self.shape_dims, self.appea_dims = shape_dims, vd_channels
        self.fg_CD_predictor = MLPforNeRF(vp_channels=vp_channels, vd_channels=vd_channels,
                                          h_channel=self.mlp_h_channel, res_nfeat=self.featmap_nc)
        self.neural_render = NeuralRenderer(bg_type=opt.bg_type, feat_nc=self.featmap_nc, out_dim=3, final_actvn=True,
                                            min_feat=32, featmap_size=self.featmap_size, img_size=self.pred_img_size)
        self._packed = None
        self._packed_key = None
        self._packed_hl = None
        self._packed_hl_key = None
        self.last_meta = None
        # "fast": fused single-pass half-precision-operand kernels (fp32 accumulate); "high": split-operand (hi+lo) tensor-core
        # GEMMs with fp32 activations, ~fp32 accuracy at ~3x the tensor work; "auto" (default): "high"
        # exactly when a camera input (batch_Rmats / batch_Tvecs / batch_inv_inmats) requires a gradient - the fitting loop,
        # whose ill-conditioned camera gradients need it - and "fast" otherwise (training, inference).  Not part of the state dict.
        self.precision = os.environ.get("HN_PRECISION", "auto")
        # "auto" also measures, on the caller's own inputs, whether the single-pass kernels keep the feature map inside
        # `auto_tolerance` of the split-operand ones (see _calibrate): checkpoints with a large feature / density scale run "high".
        self.auto_tolerance = float(os.environ.get("HN_AUTO_TOLERANCE", "5e-4"))
        self.calib_interval = int(os.environ.get("HN_CALIB_INTERVAL", "64"))     # training: re-measure every N weight versions
        self.calib_rays = 512
        self._calib = None                       # (weights key, calls since, decision "fast" | "high", measured error)
        self._fuse_grads = False
        # Weight gradients meet through atomic adds by default (summation order, hence the last bits, vary from run to run).
        # deterministic = True: every work item writes a private slice and a second kernel adds them in a fixed order (bit-identical
        # runs, one extra ~20 us kernel and a few hundred MB of workspace); None (default) follows PyTorch's own switches - the
        # reference trains with cudnn.deterministic = True (train.py:26-29).
        self.deterministic = None
        self._kernel_cache = {}
        # every derived cache (packed operand images, cached bg_img, blur taps, the precision decision) is dropped whenever the
        # parameters may have been replaced behind autograd's version counters: load_state_dict, .to()/.cuda()/.float(), and -
        # because the reference's own loader writes through `model.state_dict()[k].data.copy_(...)` (talker_trainer.py:557-567),
        # which bumps no version counter - every call of state_dict()
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_caches())
        self._register_state_dict_hook(lambda module, sd, prefix, local_metadata: module.invalidate_caches())

    def invalidate_caches(self):
        """Forget everything derived from the parameters (packed weight operands, the cached background image, blur taps, the
        "auto" precision decision).  Called automatically by load_state_dict(), state_dict() and _apply() (.to / .cuda / .half);
        call it yourself after writing parameters through `.data` (EMA, clamping, re-initialisation), which autograd's version
        counters - the cache keys - do not see."""
        self._packed_key = None
        self._packed_hl_key = None
        self._bg_cache_key = None
        self._bg_cache = None
        self._calib = None
        for m in self.neural_render.modules():
            if hasattr(m, "_taps_key"):
                m._taps_key = None

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        if hasattr(self, "_packed_key"):
            self.invalidate_caches()
            self._packed = None
            self._packed_hl = None
        return out

    # ------------------------------------------------------------------ weights -> kernel operands
    def _packed_weights(self):
        ws = [m.weight for m in self.fg_CD_predictor.layers()]
        key = tuple((w.data_ptr(), w._version) for w in ws)
        if self._packed is None or key != self._packed_key:
            self._packed = ops.pack_weights(ws, L.PE + self.shape_dims, out=self._packed)
            self._packed_key = key
        return ws, self._packed

    def _packed_weights_precise(self):
        ws = [m.weight for m in self.fg_CD_predictor.layers()]
        key = tuple((w.data_ptr(), w._version) for w in ws)
        if self._packed_hl is None or key != self._packed_hl_key:
            self._packed_hl = ops.pack_weights_precise(ws, L.PE + self.shape_dims, out=self._packed_hl)
            self._packed_hl_key = key
        return ws, self._packed_hl

    def fuse_grad_accumulation(self, enable=True):
        """Opt-in: the kernels accumulate fg_CD_predictor's and NeuralRenderer's weight and bias gradients straight into the
        parameters' existing `.grad` buffers (e.g. the views of dist.GradBucket's flat all-reduce buffer) instead of returning
        them through autograd's AccumulateGrad - no per-parameter zero-fill and add kernels.  Every parameter needs a `.grad`
        beforehand."""
        self._fuse_grads = bool(enable)
        self.neural_render.fuse_grad_accumulation(enable)
        return self

    def _grad_into(self):
        if not (self._fuse_grads and torch.is_grad_enabled()):
            return None
        lay = self.fg_CD_predictor.layers()
        if any(m.weight.requires_grad and m.weight.grad is None for m in lay) or any(m.bias.requires_grad and m.bias.grad is None for m in lay):
            raise RuntimeError("fuse_grad_accumulation: every fg_CD_predictor parameter needs an allocated .grad (e.g. dist.GradBucket)")
        return {"w": [m.weight.grad if m.weight.requires_grad else None for m in lay],
                "b": [m.bias.grad if m.bias.requires_grad else None for m in lay]}

    def _fold_biases_cuda(self, shape_code, appea_code, audiostyle, grad_into, r0_fused=False):
        """Effective bias row per batch item through the library (hn_fold_bias): [B, HN_BIAS_STRIDE].  r0_fused (the fast kernels, which
        run RGB_layer_0 multiplied into RGB_layer_1): RGB_layer_1's row also carries W_R1[:, :384] b_R0."""
        lay = self.fg_CD_predictor.layers()
        return ops.FoldBiasFunction.apply(shape_code, audiostyle, appea_code, lay[0].weight, lay[5].weight, lay[10].weight,
                                          *[m.bias for m in lay], {"grad_into": grad_into, "r0_fused": r0_fused})

    def _fold_biases(self, shape_code, appea_code, audiostyle):
        """Effective bias row per batch item: [B, HN_BIAS_STRIDE].  Plain-PyTorch statement of the folding
        algebra (CPU host-logic tests, any dtype); the render path uses _fold_biases_cuda."""
        fg = self.fg_CD_predictor
        B, S, H = shape_code.shape[0], self.shape_dims, L.HIDDEN
        lay = fg.layers()
        w0 = lay[0].weight.flatten(1)
        w5 = lay[5].weight.flatten(1)
        wr1 = lay[10].weight.flatten(1)
        rows = []
        for i in range(8):
            b = lay[i].bias.unsqueeze(0).expand(B, H)
            if i == 0:
                b = b + F.linear(shape_code, w0[:, L.PE:L.PE + S]) + F.linear(audiostyle, w0[:, L.PE + S:L.PE + S + 64])
            elif i == 5:
                b = b + F.linear(shape_code, w5[:, L.PE:L.PE + S])
            rows.append(b)
        rows.append(lay[9].bias.unsqueeze(0).expand(B, H))
        rows.append(lay[10].bias.unsqueeze(0) + F.linear(appea_code, wr1[:, H:]))
        rows.append(lay[11].bias.unsqueeze(0).expand(B, L.FEAT))
        rows.append(lay[8].bias.unsqueeze(0).expand(B, 1))
        rows.append(shape_code.new_zeros(B, L.BIAS_STRIDE - L.BIAS_OFF_DENSITY - 1))
        return torch.cat(rows, dim=1)

    # ------------------------------------------------------------------ hot path
    def render_rays(self, mode, batch_xy, audiostyle, shape_code, appea_code, batch_Rmats, batch_Tvecs,
                    batch_inv_inmats, t_rand=None, max_chunk_samples=1 << 24):
        """Arbitrary ray sets (no featmap_size constraint): -> F [B,N_r,256], bg_alpha [B,N_r].
        Without autograd, rays are processed in chunks of at most `max_chunk_samples` ray*samples so that the per-sample
        feature tensor ([M,256] fp32) stays bounded: 4M rays x 128 samples would otherwise need 0.5 TB."""
        assert mode in ["train", "test"]
        B, two, n_r = batch_xy.shape
        assert two == 2
        per_ray = B * self.num_sample_coarse
        if not torch.is_grad_enabled() and n_r * per_ray > max_chunk_samples and n_r > 2:
            step = max(2, (max_chunk_samples // per_ray) & ~1)
            outs = [self.render_rays(mode, batch_xy[:, :, i:i + step].contiguous(), audiostyle, shape_code, appea_code, batch_Rmats,
                                     batch_Tvecs, batch_inv_inmats, None if t_rand is None else t_rand[:, i:i + step].contiguous(),
                                     max_chunk_samples) for i in range(0, n_r, step)]
            return torch.cat([o[0] for o in outs], dim=1), torch.cat([o[1] for o in outs], dim=1)
        if shape_code.shape[1] != self.shape_dims or appea_code.shape[1] != self.appea_dims or audiostyle.shape[1] != 64:
            raise ValueError("latent code dimensions do not match the network")
        ns = self.num_sample_coarse
        pad = 0
        while ((n_r + pad) * ns) % L.TILE != 0:
            pad += 1
        if pad:
            batch_xy = torch.cat([batch_xy, batch_xy[:, :, -1:].expand(B, 2, pad)], dim=2)
        if mode == "train" and t_rand is None:
            # same draw as the reference's rand_like(zvals) (NetWorks/utils.py:77): shape [B,N_r,N_s+1]
            t_rand = torch.rand(B, n_r, ns + 1, device=batch_xy.device, dtype=torch.float32)
        if t_rand is not None and pad:
            t_rand = torch.cat([t_rand, t_rand[:, -1:, :].expand(B, pad, ns + 1)], dim=1)
        if self.precision not in ("auto", "fast", "high"):
            raise ValueError(f"precision must be 'auto', 'fast' or 'high', got {self.precision!r}")
        camera_grad = torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad
                                                      for t in (batch_Rmats, batch_Tvecs, batch_inv_inmats))
        high = self.precision == "high" or (self.precision == "auto" and camera_grad)
        if self.precision == "auto" and not high:
            high = self._calibrate(mode, batch_xy, audiostyle, shape_code, appea_code, batch_Rmats, batch_Tvecs, batch_inv_inmats, t_rand) == "high"
        grad_into = self._grad_into()
        bias = self._fold_biases_cuda(shape_code.float(), appea_code.float(), audiostyle.float(), grad_into, r0_fused=not high)
        meta = {"n_samples": ns, "world_z1": self.opt.world_z1, "world_z2": self.opt.world_z2,
                "l5_hidden_col": L.PE + self.shape_dims, "precision": "high" if high else "fast", "grad_into": None if high else grad_into,
                "grad_target": float(getattr(self, "grad_target", 1024.0 if high else 64.0)),
                # frozen weights with trainable biases: the weight pass must still visit the layers whose bias gradients no latent code needs
                "all_bias": any(m.bias.requires_grad for i, m in enumerate(self.fg_CD_predictor.layers()) if i not in (0, 5, 10)),
                "deterministic": self._deterministic(), "cache": self._kernel_cache,
                "r0_fused": not high, "b_r0": self.fg_CD_predictor.RGB_layer_0.bias.detach(),
                "need_b_r0": self.fg_CD_predictor.RGB_layer_0.bias.requires_grad}
        if high:
            ws, meta["packed_hl"] = self._packed_weights_precise()
        else:
            ws, meta["packed"] = self._packed_weights()
        fn = ops.RenderFunctionPrecise if high else ops.RenderFunction
        Fm, bg = fn.apply(batch_xy.float(), batch_Rmats.float(), batch_Tvecs.float(),
                          batch_inv_inmats.float(), t_rand, bias, *ws, meta)
        self.last_meta = meta
        Fm, bg = Fm.view(B, n_r + pad, L.FEAT), bg.view(B, n_r + pad)
        if pad:
            Fm, bg = Fm[:, :n_r], bg[:, :n_r]
        return Fm, bg

    def _deterministic(self):
        if self.deterministic is not None:
            return bool(self.deterministic)
        return bool(torch.backends.cudnn.deterministic or torch.are_deterministic_algorithms_enabled())

    def _calibrate(self, mode, batch_xy, audiostyle, shape_code, appea_code, batch_Rmats, batch_Tvecs, batch_inv_inmats, t_rand):
        """precision="auto": which kernel family keeps THIS checkpoint's feature map inside the parity gate?  Measured, not
        guessed: up to `calib_rays` rays, strided over the caller's own ray set, are rendered by both families (no autograd) and
        max|F_fast - F_high|, max|bg_fast - bg_high| is compared with `auto_tolerance` (half the 1e-3 gate).  The single-pass
        11-bit operands carry ~5e-4 of error relative to the feature / density scale: fine for random-init
        and moderately scaled weights, outside the absolute gate once max|F| exceeds ~2.  The decision is cached per weight
        version (inference) and re-measured every `calib_interval` versions while the weights train; one host sync per probe."""
        ws = [m.weight for m in self.fg_CD_predictor.layers()]
        key = tuple((w.data_ptr(), w._version) for w in ws)
        c = self._calib
        if c is not None and (c[0] == key or (torch.is_grad_enabled() and any(w.requires_grad for w in ws) and c[1] < self.calib_interval)):
            self._calib = (c[0], c[1] + (c[0] != key), c[2], c[3])
            return c[2]
        n_r = batch_xy.shape[2]
        step = max(1, n_r // self.calib_rays)
        idx = torch.arange(0, n_r, step, device=batch_xy.device)[: self.calib_rays]
        if idx.numel() % 2:
            idx = idx[:-1] if idx.numel() > 1 else idx
        xy = batch_xy.detach()[:, :, idx].contiguous()
        tr = None if t_rand is None else t_rand.detach()[:, idx].contiguous()
        saved = self.precision
        res = {}
        try:
            with torch.no_grad():
                for prec in ("fast", "high"):
                    self.precision = prec
                    res[prec] = self.render_rays("test" if tr is None else mode, xy, audiostyle.detach(), shape_code.detach(), appea_code.detach(),
                                                 batch_Rmats.detach(), batch_Tvecs.detach(), batch_inv_inmats.detach(), t_rand=tr)
        finally:
            self.precision = saved
        err = max(float((res["fast"][0] - res["high"][0]).abs().max()), float((res["fast"][1] - res["high"][1]).abs().max()))
        decision = "high" if (err > self.auto_tolerance or err != err) else "fast"
        self._calib = (key, 0, decision, err)
        return decision

    def _forward(self, for_train, batch_xy, batch_uv, audiostyle, bg_code, shape_code, appea_code,
                 batch_Rmats, batch_Tvecs, batch_inv_inmats, dist_expr):
        batch_size, tv, n_r = batch_xy.size()
        assert tv == 2
        assert bg_code is None
        fs, C = self.featmap_size, self.featmap_nc
        assert n_r == fs * fs, "HeadNeRFNet.forward renders a full featmap (HeadNeRFNet.py:103-106); use render_rays otherwise"
        shard = getattr(self, "_ray_shard", None)
        if shard is None:
            Fm, bg = self.render_rays("train" if for_train else "test", batch_xy, audiostyle, shape_code, appea_code,
                                      batch_Rmats, batch_Tvecs, batch_inv_inmats)
        else:
            # rays sharded inside every item: this rank renders its contiguous slice of each item's rays, the
            # slices are all-gathered (reduce-scatter backward) and the consumer below runs on whole maps on every rank
            from . import dist as hdist
            rank, world, group = shard
            xy_loc, lo, hi = hdist.shard_rays(batch_xy, rank, world)
            Fm, bg = self.render_rays("train" if for_train else "test", xy_loc, audiostyle, shape_code, appea_code,
                                      batch_Rmats, batch_Tvecs, batch_inv_inmats)
            Fm, bg = hdist.gather_rays(Fm, bg, n_r, rank, world, group)
            Fm, bg = Fm.contiguous(), bg.contiguous()
        bg_featmap = self.neural_render.get_bg_featmap()
        merge_featmap = ops.MergeFunction.apply(Fm, bg, bg_featmap)       # F^T + bg_alpha * bg_featmap (HeadNeRFNet.py:103-113)
        hook = getattr(self, "on_consumer_grads_ready", None)
        if hook is not None and merge_featmap.requires_grad:
            # the backward pass reaches the feature map: every NeuralRenderer gradient (except bg_featmap's merge term) is complete
            merge_featmap.register_hook(lambda g: (hook(), None)[1])
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.neural_render.parameters()):
            # the two renderer calls of the reference (HeadNeRFNet.py:109,113) as ONE pass over B + 1 feature maps: every
            # operator of the renderer acts per item, so the images are the same and the launch count halves
            imgs = self._render_both(torch.cat([merge_featmap, bg_featmap], dim=0))
            merge_img, bg_img = imgs[:batch_size], imgs[batch_size:]
        else:
            bg_img = self._bg_image(bg_featmap)
            merge_img = self.neural_render(merge_featmap)
        return {"coarse_dict": {"merge_img": merge_img, "bg_img": bg_img}}

    def capture_consumer_graph(self, batch_size):
        """Opt-in: capture NeuralRenderer's forward and backward over `batch_size` + 1 feature maps (the merged maps and the
        background map, see _forward) into CUDA graphs (torch.cuda.make_graphed_callables).  Useful for the module-by-module path
        (`neural_renderer.FUSED_NET = False`, deterministic mode: ~130 small launches per step, bound by the host's launch rate);
        the default one-call renderer (hn_nr_fwd / hn_nr_bwd, 45 kernels from two C calls) is GPU-bound without it.  Only calls with exactly the captured
        input shape, autograd enabled and the captured training mode replay the graphs; every other call (validation under
        no_grad, other batch sizes, the background map alone) runs the eager module.  The module tree, its state_dict keys and
        the parameters are untouched; `release_consumer_graph()` restores plain eager launches."""
        dev = self.neural_render.bg_featmap.device
        if dev.type != "cuda":
            raise RuntimeError("capture_consumer_graph needs the module on a CUDA device")
        self.release_consumer_graph()
        fs, C = self.featmap_size, self.featmap_nc
        shape = (batch_size + 1, C, fs, fs)
        sample = torch.randn(*shape, device=dev, requires_grad=True)
        for blk in self.neural_render.feat_upsample_list:          # host copies of the blur taps are read once, outside the capture
            blk.blur_layer.taps()
        self.neural_render.rgb_upsample[1].taps()
        nr = self.neural_render
        eager = nr.forward                                          # the class's bound method
        training = nr.training
        # make_graphed_callables patches `nr.forward` in place (an instance attribute) and hands the same module back: keep the
        # patched function, and put a dispatcher in its place that only replays it for the captured situation
        torch.cuda.make_graphed_callables(nr, (sample,), allow_unused_input=True)   # bg_featmap is a parameter the module itself never reads
        graphed = nr.__dict__["forward"]

        def dispatch(x):
            if torch.is_grad_enabled() and tuple(x.shape) == shape and nr.training == training and x.is_cuda:
                return graphed(x)
            return eager(x)

        nr.forward = dispatch
        object.__setattr__(self, "_consumer_graph_state", {"shape": shape, "graphed": graphed})     # not a submodule, not in state_dict
        return self

    def release_consumer_graph(self):
        """Back to eager launches: removes the dispatcher that capture_consumer_graph() put over NeuralRenderer.forward."""
        if getattr(self, "_consumer_graph_state", None) is not None:
            self.neural_render.__dict__.pop("forward", None)
            object.__setattr__(self, "_consumer_graph_state", None)
        return self

    def _render_both(self, both):
        return self.neural_render(both)

    def _bg_image(self, bg_featmap):
        """bg_img = neural_render(bg_featmap) (HeadNeRFNet.py:109) depends on parameters only: when none of them can receive a
        gradient (inference, the fitting loop with frozen weights) it is computed once per parameter version."""
        params = list(self.neural_render.parameters())
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return self.neural_render(bg_featmap)
        key = tuple((p.data_ptr(), p._version) for p in params)
        if getattr(self, "_bg_cache_key", None) != key:
            with torch.no_grad():
                self._bg_cache = self.neural_render(bg_featmap)
            self._bg_cache_key = key
        return self._bg_cache

    def set_ray_sharding(self, rank=None, world=None, group=None):
        """forward() renders only this rank's slice of every item's rays and all-gathers the composited features (dist.gather_rays);
        for batches smaller than the number of GPUs.  Call without arguments to switch it off."""
        object.__setattr__(self, "_ray_shard", None if (rank is None or world is None or world <= 1) else (int(rank), int(world), group))
        return self

    def check_faults(self):
        """Wait for the launches issued so far and raise if any kernel reported an on-chip pipeline fault (ops.FaultMonitor
        polls the same status words asynchronously at every later library call)."""
        ops.FAULTS.flush()

    def forward(self, mode, batch_xy, batch_uv, audiostyle, bg_code, shape_code, appea_code,
                batch_Rmats, batch_Tvecs, batch_inv_inmats, dist_expr=False, **kwargs):
        assert mode in ["train", "test"]
        return self._forward(mode == "train", batch_xy, batch_uv, audiostyle, bg_code, shape_code, appea_code,
                             batch_Rmats, batch_Tvecs, batch_inv_inmats, dist_expr)
