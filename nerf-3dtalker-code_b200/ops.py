"""Host-side operators over the C ABI (include/headnerf_b200.h): thin torch.autograd.Functions that own
the tensors (inputs, outputs, saved-for-backward) and pass raw device pointers + the current CUDA stream
to libheadnerf_b200.so.  PyTorch is plumbing here (memory, streams, autograd graph); every FLOP and byte of
the hot path is moved by the CUDA library.  No fallback: CPU tensors or a missing library raise."""
import ctypes as C
import os

import torch
import torch.nn.functional as F

from . import _lib as L

_DEBUG_SYNC = os.environ.get("HN_DEBUG_SYNC", "0") == "1"


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev_f32(t, name, shape=None, align=4):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise L.HeadNeRFLibraryError(f"{name} must be a CUDA tensor: this path has no CPU implementation")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
    t = t.contiguous()
    if t.data_ptr() % align:                       # a view at an odd offset of somebody's flat buffer: the kernels use vector loads
        t = t.clone()
    return t


class KernelTimer:
    """Optional CUDA-event timing of every library launch (used by bench.py): events are recorded on the
    launching stream around each C-ABI call; durations are read after the caller synchronises."""

    def __init__(self):
        self.enabled = False
        self.records = []          # (name, start_event, end_event)
        self.launches = 0

    def reset(self):
        self.records, self.launches = [], 0

    def summary(self):
        out = {}
        for name, e0, e1 in self.records:
            ms = e0.elapsed_time(e1)
            d = out.setdefault(name, {"n": 0, "ms_total": 0.0})
            d["n"] += 1
            d["ms_total"] += ms
        for d in out.values():
            d["ms_avg"] = d["ms_total"] / d["n"]
        return out


TIMER = KernelTimer()


def _call(name, fn, *args, kernels=1):
    TIMER.launches += kernels                  # kernels launched by this library call
    if TIMER.enabled:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        TIMER.records.append((name, e0, e1))
    else:
        rc = fn(*args)
    L.check(rc, name)


def check_status(status: torch.Tensor, what: str):
    """Read a kernel status word (synchronises the stream)."""
    code = int(status[0].item())
    if code != 0:
        raise L.HeadNeRFLibraryError(f"{what}: on-chip pipeline fault, status {code}")


class FaultMonitor:
    """Makes on-chip pipeline faults (a bounded spin-wait that timed out, a protocol abort: the kernel's status word) loud
    WITHOUT a per-step synchronisation: after every MLP launch the status word is copied to a pinned host slot behind the
    kernel (4 bytes, asynchronous) and an event is recorded; the NEXT library calls poll the finished events and raise for a
    non-zero word, so a fault surfaces at most a call or two later - at the latest in `flush()` (HeadNeRFNet.check_faults(),
    called by bench.py and the tests) - instead of silently returning partial tensors."""

    def __init__(self, slots=32):
        self.slots, self.pending, self.free = slots, [], []
        self.enabled = os.environ.get("HN_FAULT_MONITOR", "1") != "0"

    def _slot(self):
        if self.free:
            return self.free.pop()
        return torch.zeros(1, dtype=torch.int32).pin_memory(), torch.cuda.Event()

    def watch(self, status, what):
        if not self.enabled or torch.cuda.is_current_stream_capturing():
            return
        host, ev = self._slot()
        host.copy_(status[:1], non_blocking=True)
        ev.record()
        self.pending.append((ev, host, what))
        if len(self.pending) > self.slots:
            self.poll(block_oldest=True)

    def poll(self, block_oldest=False):
        while self.pending:
            ev, host, what = self.pending[0]
            if block_oldest:
                ev.synchronize()
                block_oldest = False
            elif not ev.query():
                break
            self.pending.pop(0)
            code = int(host[0])
            self.free.append((host, ev))
            if code != 0:
                raise L.HeadNeRFLibraryError(f"{what}: on-chip pipeline fault, status {code} (detected asynchronously; outputs of that call are invalid)")

    def flush(self):
        """Wait for every watched launch and raise if any of them faulted."""
        while self.pending:
            self.poll(block_oldest=True)


FAULTS = FaultMonitor()


def _camera(xy, R, T, Kinv, t_rand, n_samples, z1, z2):
    B, two, n_rays = xy.shape
    cam = L.Camera()
    cam.B, cam.n_rays, cam.n_samples = B, n_rays, n_samples
    cam.world_z1, cam.world_z2 = float(z1), float(z2)
    cam.xy, cam.Rmats, cam.Tvecs, cam.inv_inmats = _ptr(xy), _ptr(R), _ptr(T), _ptr(Kinv)
    cam.t_rand = _ptr(t_rand)
    return cam


def _check_camera(xy, R, T, Kinv, t_rand, n_samples):
    B, two, n_rays = xy.shape
    if two != 2:
        raise ValueError("batch_xy must be [B,2,N_r]")
    xy = _dev_f32(xy, "batch_xy")
    R = _dev_f32(R, "batch_Rmats", (B, 3, 3))
    T = _dev_f32(T.reshape(B, 3), "batch_Tvecs", (B, 3))
    Kinv = _dev_f32(Kinv, "batch_inv_inmats", (B, 3, 3))
    if t_rand is not None:
        t_rand = _dev_f32(t_rand, "t_rand", (B, n_rays, n_samples + 1))
    return xy, R, T, Kinv, t_rand


# ---------------------------------------------------------------------------------------------------
# a1/a2 standalone sampler (NetWorks/utils.py:64-161) — utility / tests; not differentiable
# ---------------------------------------------------------------------------------------------------
def sample_rays(xy, R, T, Kinv, t_rand=None, n_samples=64, world_z1=2.5, world_z2=-3.5):
    lib = L.load()
    xy, R, T, Kinv, t_rand = _check_camera(xy, R, T, Kinv, t_rand, n_samples)
    B, _, n_rays = xy.shape
    M = B * n_rays * n_samples
    dev = xy.device
    pts = torch.empty(M, 3, device=dev)
    zvals, z_dists = torch.empty(M, device=dev), torch.empty(M, device=dev)
    ray_d, ray_l = torch.empty(B * n_rays, 3, device=dev), torch.empty(B * n_rays, device=dev)
    cam = _camera(xy, R, T, Kinv, t_rand, n_samples, world_z1, world_z2)
    _call("hn_sample_rays", lib.hn_sample_rays, C.byref(cam), _ptr(pts), _ptr(zvals), _ptr(z_dists), _ptr(ray_d), _ptr(ray_l), _stream())
    return {"pts": pts, "zvals": zvals, "z_dists": z_dists, "ray_d": ray_d, "ray_l": ray_l}


# ---------------------------------------------------------------------------------------------------
# a6 alpha compositing (NetWorks/utils.py:273-309), standalone differentiable form
# ---------------------------------------------------------------------------------------------------
def _composite_fwd(feat, sigma, delta, zvals, n_samples, want_depth=False, want_weights=False):
    lib = L.load()
    M, Cc = feat.shape
    R_ = M // n_samples
    dev = feat.device
    Fm = torch.empty(R_, Cc, device=dev)
    bg = torch.empty(R_, device=dev)
    depth = torch.empty(R_, device=dev) if (want_depth and zvals is not None) else None
    w = torch.empty(M, device=dev) if want_weights else None
    a = L.CompositeFwd()
    a.n_rays_total, a.n_samples, a.C = R_, n_samples, Cc
    a.feat, a.sigma, a.delta, a.zvals = _ptr(feat), _ptr(sigma), _ptr(delta), _ptr(zvals)
    a.F, a.bg_alpha, a.depth, a.weights = _ptr(Fm), _ptr(bg), _ptr(depth), _ptr(w)
    _call("hn_composite_fwd", lib.hn_composite_fwd, C.byref(a), _stream())
    return Fm, bg, depth, w


def _composite_bwd(feat, sigma, delta, zvals, gF, g_bg, g_depth, n_samples, image=False, grad_scale=None, want_ddelta=True):
    lib = L.load()
    M, Cc = feat.shape
    dev = feat.device
    dfeat = None if image else torch.empty(M, Cc, device=dev)
    dimg = torch.empty(lib.hn_dfeat_image_bytes(M), dtype=torch.uint8, device=dev) if image else None
    dsigma = torch.empty(M, device=dev)
    ddelta = torch.empty(M, device=dev) if want_ddelta else None
    a = L.CompositeBwd()
    a.n_rays_total, a.n_samples, a.C = M // n_samples, n_samples, Cc
    a.feat, a.sigma, a.delta, a.zvals = _ptr(feat), _ptr(sigma), _ptr(delta), _ptr(zvals)
    a.gF, a.g_bg, a.g_depth = _ptr(gF), _ptr(g_bg), _ptr(g_depth)
    a.dfeat, a.dfeat_image, a.grad_scale, a.dsigma, a.ddelta = _ptr(dfeat), _ptr(dimg), _ptr(grad_scale), _ptr(dsigma), _ptr(ddelta)
    _call("hn_composite_bwd", lib.hn_composite_bwd, C.byref(a), _stream())
    return dfeat, dimg, dsigma, ddelta


class CompositeFunction(torch.autograd.Function):
    """(feat [M,C], sigma [M], delta [M], zvals [M]) -> (F [R,C], bg_alpha [R], depth [R])."""

    @staticmethod
    def forward(ctx, feat, sigma, delta, zvals, n_samples):
        feat, sigma, delta = _dev_f32(feat, "feat"), _dev_f32(sigma, "sigma"), _dev_f32(delta, "delta")
        zvals = _dev_f32(zvals, "zvals") if zvals is not None else None
        Fm, bg, depth, _ = _composite_fwd(feat, sigma, delta, zvals, n_samples, want_depth=True)
        ctx.save_for_backward(feat, sigma, delta, zvals)
        ctx.n_samples = n_samples
        if depth is None:
            depth = torch.zeros_like(bg)
        ctx.mark_non_differentiable(depth) if zvals is None else None
        return Fm, bg, depth

    @staticmethod
    def backward(ctx, gF, g_bg, g_depth):
        feat, sigma, delta, zvals = ctx.saved_tensors
        g_depth = g_depth.contiguous() if (g_depth is not None and zvals is not None) else None
        dfeat, _, dsigma, ddelta = _composite_bwd(feat, sigma, delta, zvals, gF.contiguous(), g_bg.contiguous(),
                                                  g_depth, ctx.n_samples)
        dz = None
        if zvals is not None and g_depth is not None and ctx.needs_input_grad[3]:
            _, _, _, w = _composite_fwd(feat, sigma, delta, zvals, ctx.n_samples, want_weights=True)
            dz = w * g_depth.repeat_interleave(ctx.n_samples)
        return dfeat, dsigma, ddelta, dz, None


def composite(feat, sigma, delta, zvals=None, n_samples=64):
    return CompositeFunction.apply(feat, sigma, delta, zvals, n_samples)


# ---------------------------------------------------------------------------------------------------
# weight packing
# ---------------------------------------------------------------------------------------------------
def pack_weights(weights12, l5_hidden_col, out=None):
    """weights12: the 12 fp32 CUDA weight tensors in header order (any [out,in,...] shape) -> packed uint8 buffer."""
    lib = L.load()
    ws = [_dev_f32(w.detach(), f"weight[{i}]") for i, w in enumerate(weights12)]
    n = lib.hn_packed_weights_bytes()
    if out is None or out.numel() != n or out.device != ws[0].device:
        out = torch.empty(n, dtype=torch.uint8, device=ws[0].device)
    a = L.Weights()
    for i, w in enumerate(ws):
        a.w[i] = w.data_ptr()
        a.ld[i] = w.numel() // w.shape[0]
    a.l5_hidden_col = l5_hidden_col
    _call("hn_pack_weights", lib.hn_pack_weights, C.byref(a), _ptr(out), _stream())
    return out


# ---------------------------------------------------------------------------------------------------
# the fused render operator: sampling + PE + MLP + compositing, forward and backward
# ---------------------------------------------------------------------------------------------------
def ray_params_torch(xy, R, T, Kinv):
    """Differentiable per-ray quantities (NetWorks/utils.py:147-158): origin o, v = d*l, l.  Used only in
    backward to chain the kernel's per-ray gradients to R / T / K^-1 (a few hundred floats per ray)."""
    B, _, n_rays = xy.shape
    xyz = F.pad(xy, [0, 0, 0, 1, 0, 0], mode="constant", value=1.0)
    d = R.bmm(Kinv.bmm(xyz))
    d = d / torch.norm(d, dim=1, keepdim=True)
    l = -1.0 / d[:, -1:, :]
    o = T.reshape(B, 3, 1).expand(B, 3, n_rays)
    return o, d * l, l


_SCALE_SCRATCH = {}


def loss_scale(gF, target):
    """Device scalar 2^floor(log2(target / max|gF|)) (hn_loss_scale): one launch, no host sync."""
    lib = L.load()
    key = (gF.device, torch.cuda.current_stream().cuda_stream)
    if key not in _SCALE_SCRATCH:
        _SCALE_SCRATCH[key] = torch.zeros(2, dtype=torch.int32, device=gF.device)
    scale = torch.empty(1, device=gF.device)
    _call("hn_loss_scale", lib.hn_loss_scale, _ptr(gF), gF.numel(), C.c_float(float(target)), _ptr(scale), _ptr(_SCALE_SCRATCH[key]), _stream())
    return scale


def _grad_dst(into, i, like):
    """Gradient destination of input i: the caller's accumulation buffer (fused accumulation, e.g. a view of the flat
    all-reduce bucket) or None."""
    if into is None or into[i] is None:
        return None
    t = into[i]
    if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != like.numel() or t.device != like.device:
        raise ValueError("fused gradient accumulation needs contiguous float32 buffers shaped like the parameters")
    return t


class FoldBiasFunction(torch.autograd.Function):
    """(shape_code [B,S], audiostyle [B,64], appea_code [B,A], W0, W5, WR1, 12 biases, meta) -> effective bias rows
    [B, HN_BIAS_STRIDE] (hn_fold_bias / hn_fold_bias_bwd).  meta["grad_into"] (optional): {"w": 12 weight-gradient
    buffers, "b": 12 bias-gradient buffers} to accumulate into directly instead of returning parameter gradients."""

    @staticmethod
    def _args(shape, audio, appea, w0, w5, wr1, biases, r0_fused=False, wr0=None):
        a = L.Fold()
        a.r0_fused = 1 if r0_fused else 0                           # the fast chains skip RGB_layer_0 (csrc/hn_mlp_sched.h)
        if wr0 is not None:
            a.wr0, a.ldr0 = wr0.data_ptr(), wr0.numel() // wr0.shape[0]
        a.B, a.shape_dims, a.appea_dims = shape.shape[0], shape.shape[1], appea.shape[1]
        a.w0, a.ld0 = w0.data_ptr(), w0.numel() // w0.shape[0]
        a.w5, a.ld5 = w5.data_ptr(), w5.numel() // w5.shape[0]
        a.wr1, a.ldr1 = wr1.data_ptr(), wr1.numel() // wr1.shape[0]
        for i, b in enumerate(biases):
            a.bias[i] = b.data_ptr()
        a.shape_code, a.audio, a.appea = shape.data_ptr(), audio.data_ptr(), appea.data_ptr()
        return a

    @staticmethod
    def forward(ctx, shape, audio, appea, w0, w5, wr1, *rest):
        biases, meta = rest[:12], rest[12]
        lib = L.load()
        shape, audio, appea = _dev_f32(shape, "shape_code"), _dev_f32(audio, "audiostyle"), _dev_f32(appea, "appea_code")
        if audio.shape[1] != 64 or audio.shape[0] != shape.shape[0] or appea.shape[0] != shape.shape[0]:
            raise ValueError("latent code shapes: shape [B,S], audiostyle [B,64], appea [B,A]")
        w0, w5, wr1 = (_dev_f32(w.detach(), "folded weight") for w in (w0, w5, wr1))
        bs = [_dev_f32(b.detach(), "bias") for b in biases]
        out = torch.empty(shape.shape[0], L.BIAS_STRIDE, device=shape.device)
        a = FoldBiasFunction._args(shape, audio, appea, w0, w5, wr1, bs, meta.get("r0_fused", False))
        _call("hn_fold_bias", lib.hn_fold_bias, C.byref(a), _ptr(out), _stream())
        ctx.save_for_backward(shape, audio, appea, w0, w5, wr1, *bs)
        ctx.meta = meta
        return out

    @staticmethod
    def backward(ctx, g):
        lib = L.load()
        shape, audio, appea, w0, w5, wr1 = ctx.saved_tensors[:6]
        bs = ctx.saved_tensors[6:]
        need = ctx.needs_input_grad
        into = ctx.meta.get("grad_into")
        g = g.contiguous().float()
        a = FoldBiasFunction._args(shape, audio, appea, w0, w5, wr1, bs, ctx.meta.get("r0_fused", False))
        out = L.FoldGrads()
        res = [None] * 19
        for i, (name, t) in enumerate((("dshape", shape), ("daudio", audio), ("dappea", appea))):
            if need[i]:
                res[i] = torch.empty_like(t)
                setattr(out, name, res[i].data_ptr())
        for i, (name, t, widx) in enumerate((("dw0", w0, 0), ("dw5", w5, 5), ("dwr1", wr1, 10))):
            if need[3 + i]:
                dst = _grad_dst(into["w"] if into else None, widx, t)
                if dst is None:
                    dst = res[3 + i] = torch.zeros_like(t)
                setattr(out, name, dst.data_ptr())
        for i, b in enumerate(bs):
            if need[6 + i]:
                dst = _grad_dst(into["b"] if into else None, i, b)
                if dst is None:
                    dst = res[6 + i] = torch.zeros_like(b)
                out.dbias[i] = dst.data_ptr()
        _call("hn_fold_bias_bwd", lib.hn_fold_bias_bwd, C.byref(a), _ptr(g), C.byref(out), _stream(), kernels=2)
        return tuple(res)


def _det_workspace(meta, B, dev):
    """Deterministic weight gradients (meta["deterministic"]): the private-slice workspace of hn_mlp_bwd_weights, kept on the module
    (meta["cache"]) between steps - a few hundred MB that need no zero-fill."""
    if not meta.get("deterministic"):
        return None
    cache = meta.setdefault("cache", {})
    n = L.load().hn_wgrad_det_workspace_bytes(B)
    ws = cache.get("det_ws")
    if ws is None or ws.numel() < n or ws.device != dev:
        ws = cache["det_ws"] = torch.empty(n, dtype=torch.uint8, device=dev)
    return ws


class RenderFunction(torch.autograd.Function):
    """inputs : xy [B,2,N_r], R [B,3,3], T [B,3,1], K^-1 [B,3,3], t_rand or None, bias_eff [B,3920],
                12 weights (header order; [8] is density_module.weight), then non-tensor meta
       outputs: F [B*N_r, 256], bg_alpha [B*N_r]"""

    @staticmethod
    def forward(ctx, xy, R, T, Kinv, t_rand, bias_eff, *rest):
        weights, meta = rest[:12], rest[12]
        w_density = weights[8].detach()
        lib = L.load()
        FAULTS.poll()
        ns = meta["n_samples"]
        xy_c, R_c, T_c, K_c, tr_c = _check_camera(xy, R, T, Kinv, t_rand, ns)
        B, _, n_rays = xy_c.shape
        bias_c = _dev_f32(bias_eff, "bias_eff", (B, L.BIAS_STRIDE), align=16)
        wd = _dev_f32(w_density.reshape(-1), "w_density", (L.HIDDEN,), align=16)
        M = B * n_rays * ns
        dev = xy_c.device
        need_bwd = any(ctx.needs_input_grad)
        feat = torch.empty(M, L.FEAT, device=dev)
        sigma, delta = torch.empty(M, device=dev), torch.empty(M, device=dev)
        act = torch.empty(lib.hn_act_bytes(M), dtype=torch.uint8, device=dev) if need_bwd else None
        masks = torch.empty(M * L.MASK_WORDS, dtype=torch.int32, device=dev) if need_bwd else None
        status = torch.zeros(64 if not os.environ.get("HN_TRACE") else 8192, dtype=torch.int32, device=dev)   # HN_TRACE: room for the debug timeline
        a = L.MlpFwd()
        a.cam = _camera(xy_c, R_c, T_c, K_c, tr_c, ns, meta["world_z1"], meta["world_z2"])
        a.bias, a.w_density, a.packed = _ptr(bias_c), _ptr(wd), _ptr(meta["packed"])
        a.feat, a.sigma, a.delta, a.zvals = _ptr(feat), _ptr(sigma), _ptr(delta), None
        a.act, a.masks, a.status = _ptr(act), _ptr(masks), _ptr(status)
        _call("hn_mlp_fwd", lib.hn_mlp_fwd, C.byref(a), _stream())
        Fm, bg, _, _ = _composite_fwd(feat, sigma, delta, None, ns)
        if _DEBUG_SYNC:
            check_status(status, "hn_mlp_fwd")
        FAULTS.watch(status, "hn_mlp_fwd")
        meta["last_status"] = status
        if need_bwd:
            ctx.save_for_backward(xy_c, R_c, T_c, K_c, tr_c, wd, feat, sigma, delta, act, masks, *weights)
            ctx.meta = meta
            ctx.T_shape = T.shape
        return Fm, bg

    @staticmethod
    def backward(ctx, gF, g_bg):
        lib = L.load()
        FAULTS.poll()
        xy, R, T, Kinv, t_rand, wd, feat, sigma, delta, act, masks = ctx.saved_tensors[:11]
        weights = ctx.saved_tensors[11:]
        meta = ctx.meta
        ns = meta["n_samples"]
        B, _, n_rays = xy.shape
        M = B * n_rays * ns
        dev = xy.device
        need = ctx.needs_input_grad
        need_cam = need[1] or need[2] or need[3]
        need_w = any(need[6:18])
        need_bias = need[5]
        gF = gF.contiguous().float()
        g_bg = g_bg.contiguous().float()
        # power-of-two loss scale so that half-precision gradient operands stay in range
        scale = loss_scale(gF, meta.get("grad_target", 64.0))
        _, dimg, dsigma, ddelta = _composite_bwd(feat, sigma, delta, None, gF, g_bg, None, ns, image=True,
                                                 grad_scale=scale, want_ddelta=need_cam)
        save_grads = need_w or need_bias
        grads = torch.empty(lib.hn_grads_bytes(M), dtype=torch.uint8, device=dev) if save_grads else None
        # every zero-initialised output of this pass is a view of ONE buffer: one memset instead of ~20
        into = meta.get("grad_into")
        into_w = into["w"] if into else None
        dst_w = [_grad_dst(into_w, i, weights[i]) if (need_w and need[6 + i]) else None for i in range(12)]
        # RGB_layer_0 is folded into RGB_layer_1 (csrc/hn_mlp_sched.h): the weight pass delivers dL/dW_f, hn_unfuse_r0r1 the two gradients
        need_r = need_w and (need[6 + 9] or need[6 + 10])
        run_unfuse = need_r or (need_bias and meta.get("need_b_r0", False))
        sizes = {"status": 64, "g_o": B * n_rays * 3 if need_cam else 0, "g_v": B * n_rays * 3 if need_cam else 0,
                 "g_l": B * n_rays if need_cam else 0, "dbias": B * L.BIAS_STRIDE if save_grads else 0,
                 "dwf": L.RGB1 * L.HIDDEN if run_unfuse else 0}
        for i in range(12):
            sizes[f"w{i}"] = weights[i].numel() if (need_w and need[6 + i] and dst_w[i] is None) else 0
        zbuf = torch.zeros(sum(sizes.values()), device=dev)
        views, off = {}, 0
        for k, n in sizes.items():
            views[k] = zbuf[off:off + n] if n else None
            off += n
        g_o = views["g_o"].view(B * n_rays, 3) if need_cam else None
        g_v = views["g_v"].view(B * n_rays, 3) if need_cam else None
        g_l = views["g_l"]
        status = views["status"].view(torch.int32)
        a = L.MlpBwdData()
        a.cam = _camera(xy, R, T, Kinv, t_rand, ns, meta["world_z1"], meta["world_z2"])
        a.packed, a.w_density, a.dfeat_image = _ptr(meta["packed"]), _ptr(wd), _ptr(dimg)
        a.dsigma, a.ddelta, a.sigma, a.grad_scale = _ptr(dsigma), _ptr(ddelta), _ptr(sigma), _ptr(scale)
        a.masks, a.act, a.grads = _ptr(masks), _ptr(act), _ptr(grads)
        a.g_ray_o, a.g_ray_v, a.g_ray_l, a.status = _ptr(g_o), _ptr(g_v), _ptr(g_l), _ptr(status)
        _call("hn_mlp_bwd_data", lib.hn_mlp_bwd_data, C.byref(a), _stream())

        dws = [None] * 12
        dbias = None
        if save_grads:
            dbias = views["dbias"].view(B, L.BIAS_STRIDE)
            w = L.MlpBwdWeights()
            w.B, w.n_rays, w.n_samples = B, n_rays, ns
            w.act, w.grads, w.dfeat_image, w.grad_scale = _ptr(act), _ptr(grads), _ptr(dimg), _ptr(scale)
            ws_bytes = lib.hn_wgrad_workspace_bytes(B)
            wksp = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            w.items_workspace, w.items_workspace_bytes = _ptr(wksp), ws_bytes
            for i, wt in enumerate(weights):
                if need_w and need[6 + i]:
                    if dst_w[i] is not None:                       # fused accumulation: straight into the caller's buffer
                        w.dw[i] = dst_w[i].data_ptr()
                    else:
                        dws[i] = views[f"w{i}"].view_as(wt)
                        w.dw[i] = dws[i].data_ptr()
                else:
                    w.dw[i] = None
                w.ld[i] = wt.numel() // wt.shape[0]
            dst_r0 = (dst_w[9] if dst_w[9] is not None else dws[9]) if (need_w and need[6 + 9]) else None
            dst_r1 = (dst_w[10] if dst_w[10] is not None else dws[10]) if (need_w and need[6 + 10]) else None
            w.r0_fused = 1
            w.dw[9] = None
            w.dw[10] = views["dwf"].data_ptr() if need_r else None   # (non-NULL = "RGB_layer_1's fused gradient is wanted": it goes to dwf)
            w.dwf = _ptr(views["dwf"]) if need_r else None
            w.l5_hidden_col = meta["l5_hidden_col"]
            w.dbias, w.status = _ptr(dbias), _ptr(status)
            w.want_all_bias = 1 if (need_bias and meta.get("all_bias", True)) else 0
            det_ws = _det_workspace(meta, B, dev)
            w.det_workspace, w.det_workspace_bytes = _ptr(det_ws), (det_ws.numel() if det_ws is not None else 0)
            # with weight gradients the library launches two kernels: 3-CTA clusters for the 384-wide layers, then the rest
            _call("hn_mlp_bwd_weights", lib.hn_mlp_bwd_weights, C.byref(w), _stream(), kernels=3 if need_w else 1)
            if run_unfuse:
                u = L.Unfuse()
                u.B = B
                u.wr0, u.ldr0 = _ptr(weights[9]), weights[9].numel() // weights[9].shape[0]
                u.wr1, u.ldr1 = _ptr(weights[10]), weights[10].numel() // weights[10].shape[0]
                u.b_r0, u.dwf, u.dbias_eff = _ptr(meta["b_r0"]), _ptr(views["dwf"]), _ptr(dbias)
                u.dwr0, u.dwr1 = _ptr(dst_r0), _ptr(dst_r1)
                _call("hn_unfuse_r0r1", lib.hn_unfuse_r0r1, C.byref(u), _stream())
        if _DEBUG_SYNC:
            check_status(status, "hn_mlp_bwd")
        FAULTS.watch(status, "hn_mlp_bwd_data / hn_mlp_bwd_weights")
        meta["last_status"] = status

        gR, gT, gK = _camera_chain(xy, R, T, Kinv, g_o, g_v, g_l, need, ctx.T_shape) if need_cam else (None, None, None)
        g_weights = [dws[i] if need[6 + i] else None for i in range(12)]
        return (None, gR, gT, gK, None, dbias if need_bias else None, *g_weights, None)


def _camera_chain(xy, R, T, Kinv, g_o, g_v, g_l, need, T_shape):
    """Per-ray gradients (origin, direction*length, length) -> dL/dR, dL/dT, dL/dK^-1 through NetWorks/utils.py:147-158
    (hn_camera_bwd: one launch)."""
    lib = L.load()
    B, _, n_rays = xy.shape
    out = torch.zeros(B, 21, device=xy.device)
    gR, gT, gK = out[:, :9], out[:, 9:12], out[:, 12:]
    cam = _camera(xy, R, T, Kinv, None, 1, 0.0, 0.0)
    zeros21 = out.data_ptr()
    _call("hn_camera_bwd", lib.hn_camera_bwd, C.byref(cam), _ptr(g_o), _ptr(g_v), _ptr(g_l),
          C.c_void_p(zeros21) if need[1] else None, C.c_void_p(zeros21 + 4 * 9 * B) if need[2] else None,
          C.c_void_p(zeros21 + 4 * 12 * B) if need[3] else None, _stream())
    flat = out.view(-1)                                                # laid out [B*9 | B*3 | B*9]
    gR = flat[:9 * B].view(B, 3, 3) if need[1] else None
    gT = flat[9 * B:12 * B].view(B, 3).reshape(T_shape) if need[2] else None
    gK = flat[12 * B:].view(B, 3, 3) if need[3] else None
    return gR, gT, gK


def camera_chain_torch(xy, R, T, Kinv, g_o, g_v, g_l):
    """Reference statement of the same chain with torch autograd (tests only)."""
    B, _, n_rays = xy.shape
    with torch.enable_grad():
        Rr, Tr, Kr = R.detach().requires_grad_(True), T.detach().requires_grad_(True), Kinv.detach().requires_grad_(True)
        o, v, l = ray_params_torch(xy, Rr, Tr, Kr)
        gouts = [g_o.view(B, n_rays, 3).permute(0, 2, 1), g_v.view(B, n_rays, 3).permute(0, 2, 1), g_l.view(B, 1, n_rays)]
        return torch.autograd.grad([o, v, l], [Rr, Tr, Kr], gouts)


def pack_weights_precise(weights12, l5_hidden_col, out=None):
    """hi|lo split weight units for the high-precision mode (csrc/hn_precise.cu)."""
    lib = L.load()
    ws = [_dev_f32(w.detach(), f"weight[{i}]") for i, w in enumerate(weights12)]
    n = lib.hn_precise_packed_bytes()
    if out is None or out.numel() != n or out.device != ws[0].device:
        out = torch.empty(n, dtype=torch.uint8, device=ws[0].device)
    a = L.Weights()
    for i, w in enumerate(ws):
        a.w[i] = w.data_ptr()
        a.ld[i] = w.numel() // w.shape[0]
    a.l5_hidden_col = l5_hidden_col
    _call("hn_pack_weights_precise", lib.hn_pack_weights_precise, C.byref(a), _ptr(out), _stream())
    return out


class RenderFunctionPrecise(torch.autograd.Function):
    """Same contract as RenderFunction, high-precision mode: split-operand (hi+lo) tensor-core GEMMs layer by layer with
    fp32 activations in HBM (include/headnerf_b200.h, hn_mlp_fwd_precise / hn_mlp_bwd_data_precise)."""

    @staticmethod
    def forward(ctx, xy, R, T, Kinv, t_rand, bias_eff, *rest):
        weights, meta = rest[:12], rest[12]
        lib = L.load()
        ns = meta["n_samples"]
        xy_c, R_c, T_c, K_c, tr_c = _check_camera(xy, R, T, Kinv, t_rand, ns)
        B, _, n_rays = xy_c.shape
        bias_c = _dev_f32(bias_eff, "bias_eff", (B, L.BIAS_STRIDE), align=16)
        wd = _dev_f32(weights[8].detach().reshape(-1), "w_density", (L.HIDDEN,), align=16)
        M = B * n_rays * ns
        dev = xy_c.device
        feat = torch.empty(M, L.FEAT, device=dev)
        sigma, delta = torch.empty(M, device=dev), torch.empty(M, device=dev)
        acts = torch.empty(lib.hn_precise_workspace_floats(M), device=dev)
        status = torch.zeros(64, dtype=torch.int32, device=dev)
        a = L.MlpFwdPrecise()
        a.cam = _camera(xy_c, R_c, T_c, K_c, tr_c, ns, meta["world_z1"], meta["world_z2"])
        a.bias, a.w_density, a.packed_hl = _ptr(bias_c), _ptr(wd), _ptr(meta["packed_hl"])
        a.feat, a.sigma, a.delta, a.zvals, a.acts, a.status = _ptr(feat), _ptr(sigma), _ptr(delta), None, _ptr(acts), _ptr(status)
        _call("hn_mlp_fwd_precise", lib.hn_mlp_fwd_precise, C.byref(a), _stream(), kernels=13)
        Fm, bg, _, _ = _composite_fwd(feat, sigma, delta, None, ns)
        if _DEBUG_SYNC:
            check_status(status, "hn_mlp_fwd_precise")
        FAULTS.watch(status, "hn_mlp_fwd_precise")
        meta["last_status"] = status
        if any(ctx.needs_input_grad):
            ctx.save_for_backward(xy_c, R_c, T_c, K_c, tr_c, wd, feat, sigma, delta, acts, *weights)
            ctx.meta = meta
            ctx.T_shape = T.shape
        return Fm, bg

    @staticmethod
    def backward(ctx, gF, g_bg):
        lib = L.load()
        xy, R, T, Kinv, t_rand, wd, feat, sigma, delta, acts = ctx.saved_tensors[:10]
        weights = ctx.saved_tensors[10:]
        meta = ctx.meta
        ns = meta["n_samples"]
        B, _, n_rays = xy.shape
        M = B * n_rays * ns
        dev = xy.device
        need = ctx.needs_input_grad
        need_cam = need[1] or need[2] or need[3]
        need_w = any(need[6:18])
        need_bias = need[5]
        gF = gF.contiguous().float()
        g_bg = g_bg.contiguous().float()
        scale = loss_scale(gF, meta.get("grad_target", 64.0))
        dfeat, _, dsigma, ddelta = _composite_bwd(feat, sigma, delta, None, gF, g_bg, None, ns, image=False, want_ddelta=need_cam)
        gz = torch.empty(lib.hn_precise_workspace_floats(M), device=dev)
        g_o = torch.zeros(B * n_rays, 3, device=dev) if need_cam else None
        g_v = torch.zeros(B * n_rays, 3, device=dev) if need_cam else None
        g_l = torch.zeros(B * n_rays, device=dev) if need_cam else None
        dbias = torch.zeros(B, L.BIAS_STRIDE, device=dev) if (need_bias or need_w) else None
        status = torch.zeros(64, dtype=torch.int32, device=dev)
        act_img = grads_img = dfeat_img = None
        if need_w:
            act_img = torch.empty(lib.hn_act_bytes(M), dtype=torch.uint8, device=dev)
            grads_img = torch.empty(lib.hn_grads_bytes(M), dtype=torch.uint8, device=dev)
            dfeat_img = torch.empty(lib.hn_dfeat_image_bytes(M), dtype=torch.uint8, device=dev)
        a = L.MlpBwdDataPrecise()
        a.cam = _camera(xy, R, T, Kinv, t_rand, ns, meta["world_z1"], meta["world_z2"])
        a.packed_hl, a.w_density, a.dfeat = _ptr(meta["packed_hl"]), _ptr(wd), _ptr(dfeat)
        a.dsigma, a.ddelta, a.sigma, a.grad_scale = _ptr(dsigma), _ptr(ddelta), _ptr(sigma), _ptr(scale)
        a.acts, a.gz = _ptr(acts), _ptr(gz)
        a.g_ray_o, a.g_ray_v, a.g_ray_l = _ptr(g_o), _ptr(g_v), _ptr(g_l)
        # with weight gradients the bias gradients come out of the weight pass (ones-operand MMA) instead
        a.dbias = _ptr(dbias) if (need_bias and not need_w) else None
        a.act_image, a.grads_image, a.dfeat_image, a.status = _ptr(act_img), _ptr(grads_img), _ptr(dfeat_img), _ptr(status)
        n_k = 10 + (2 if need_cam else 0) + (12 if (need_bias and not need_w) else 0) + (23 if need_w else 0)
        _call("hn_mlp_bwd_data_precise", lib.hn_mlp_bwd_data_precise, C.byref(a), _stream(), kernels=n_k)
        dws = [None] * 12
        if need_w:
            w = L.MlpBwdWeights()
            w.B, w.n_rays, w.n_samples = B, n_rays, ns
            w.act, w.grads, w.dfeat_image, w.grad_scale = _ptr(act_img), _ptr(grads_img), _ptr(dfeat_img), _ptr(scale)
            ws_bytes = lib.hn_wgrad_workspace_bytes(B)
            wksp = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            w.items_workspace, w.items_workspace_bytes = _ptr(wksp), ws_bytes
            for i, wt in enumerate(weights):
                dws[i] = torch.zeros_like(wt)
                w.dw[i] = dws[i].data_ptr()
                w.ld[i] = wt.numel() // wt.shape[0]
            w.l5_hidden_col = meta["l5_hidden_col"]
            w.dbias, w.status = _ptr(dbias), _ptr(status)
            det_ws = _det_workspace(meta, B, dev)
            w.det_workspace, w.det_workspace_bytes = _ptr(det_ws), (det_ws.numel() if det_ws is not None else 0)
            _call("hn_mlp_bwd_weights", lib.hn_mlp_bwd_weights, C.byref(w), _stream(), kernels=3)
        if _DEBUG_SYNC:
            check_status(status, "hn_mlp_bwd_precise")
        FAULTS.watch(status, "hn_mlp_bwd_data_precise / hn_mlp_bwd_weights")
        meta["last_status"] = status
        gR, gT, gK = _camera_chain(xy, R, T, Kinv, g_o, g_v, g_l, need, ctx.T_shape) if need_cam else (None, None, None)
        g_weights = [dws[i] if need[6 + i] else None for i in range(12)]
        return (None, gR, gT, gK, None, dbias if need_bias else None, *g_weights, None)


# ---------------------------------------------------------------------------------------------------
# a7 merge (NetWorks/HeadNeRFNet.py:103-113)
# ---------------------------------------------------------------------------------------------------
class MergeFunction(torch.autograd.Function):
    """(F [B,N_r,C] ray-major, bg_alpha [B,N_r], bg_featmap [1,C,fs,fs]) -> merge featmap [B,C,fs,fs] = F^T + bg_alpha * bg_featmap
    (hn_merge_fwd / hn_merge_bwd: one transposing kernel each way)."""

    @staticmethod
    def forward(ctx, Fm, bg, bgfeat):
        lib = L.load()
        Fm, bg, bgfeat = _dev_f32(Fm, "F"), _dev_f32(bg, "bg_alpha"), _dev_f32(bgfeat, "bg_featmap")
        B, n_r, Cc = Fm.shape
        fs = bgfeat.shape[-1]
        if tuple(bg.shape) != (B, n_r) or tuple(bgfeat.shape) != (1, Cc, fs, fs) or fs * fs != n_r:
            raise ValueError("merge: F [B,fs*fs,C], bg_alpha [B,fs*fs], bg_featmap [1,C,fs,fs]")
        out = torch.empty(B, Cc, fs, fs, device=Fm.device)
        _call("hn_merge_fwd", lib.hn_merge_fwd, _ptr(Fm), _ptr(bg), _ptr(bgfeat), _ptr(out), B, n_r, Cc, _stream())
        ctx.save_for_backward(bg, bgfeat)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = L.load()
        bg, bgfeat = ctx.saved_tensors
        B, n_r = bg.shape
        Cc = bgfeat.shape[1]
        g = g.contiguous().float()
        need = ctx.needs_input_grad
        gF = torch.empty(B, n_r, Cc, device=g.device) if need[0] else None
        zb = torch.zeros(B * n_r + Cc * n_r, device=g.device) if (need[1] or need[2]) else None
        g_bg = zb[:B * n_r].view(B, n_r) if need[1] else None
        g_feat = zb[B * n_r:].view(bgfeat.shape) if need[2] else None
        _call("hn_merge_bwd", lib.hn_merge_bwd, _ptr(g), _ptr(bg), _ptr(bgfeat), _ptr(gF), _ptr(g_bg), _ptr(g_feat), B, n_r, Cc, _stream())
        return gF, g_bg, g_feat


# ---------------------------------------------------------------------------------------------------
# consumer side (first pieces): fused tails of NeuralRenderer's up-sampling blocks
# ---------------------------------------------------------------------------------------------------
def _taps(f3):
    return (C.c_float * 3)(*[float(v) for v in f3])


class UpsampleTailFunction(torch.autograd.Function):
    """(z2 [B,4C,H,W], x [B,C,H,W], taps) -> blur(pixel_shuffle(leaky_relu(z2, 0.2) + repeat(x, 4), 2)) [B,C,2H,2W]
    (NetWorks/PixelShuffleUpsample.py:36-45) in one kernel each way."""

    @staticmethod
    def forward(ctx, z2, x, f3):
        lib = L.load()
        z2, x = _dev_f32(z2, "z2"), _dev_f32(x, "x")
        B, C4, H, W = z2.shape
        if C4 != 4 * x.shape[1] or x.shape[0] != B or tuple(x.shape[2:]) != (H, W):
            raise ValueError("upsample tail: z2 must be [B,4C,H,W] and x [B,C,H,W]")
        y = torch.empty(B, C4 // 4, 2 * H, 2 * W, device=z2.device)
        _call("hn_upsample_tail_fwd", lib.hn_upsample_tail_fwd, _ptr(z2), _ptr(x), _taps(f3), _ptr(y), B, C4 // 4, H, W, _stream())
        ctx.save_for_backward(z2)
        ctx.f3 = tuple(f3)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = L.load()
        (z2,) = ctx.saved_tensors
        B, C4, H, W = z2.shape
        dy = dy.contiguous().float()
        dz2 = torch.empty_like(z2) if ctx.needs_input_grad[0] else None
        dx = torch.zeros(B, C4 // 4, H, W, device=z2.device) if ctx.needs_input_grad[1] else None      # accumulated with atomics
        _call("hn_upsample_tail_bwd", lib.hn_upsample_tail_bwd, _ptr(dy), _ptr(z2), _taps(ctx.f3), _ptr(dz2), _ptr(dx), B, C4 // 4, H, W, _stream())
        return dz2, dx, None


class RgbUpsampleFunction(torch.autograd.Function):
    """x [B,K,H,W] -> blur(bilinear x2, align_corners=False) [B,K,2H,2W] (NetWorks/neural_renderer.py:47-50)."""

    @staticmethod
    def forward(ctx, x, f3):
        lib = L.load()
        x = _dev_f32(x, "rgb")
        B, K, H, W = x.shape
        y = torch.empty(B, K, 2 * H, 2 * W, device=x.device)
        _call("hn_rgb_upsample_fwd", lib.hn_rgb_upsample_fwd, _ptr(x), _taps(f3), _ptr(y), B * K, H, W, _stream())
        ctx.shape, ctx.f3 = (B, K, H, W), tuple(f3)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = L.load()
        B, K, H, W = ctx.shape
        dy = dy.contiguous().float()
        dx = torch.zeros(B, K, H, W, device=dy.device)
        _call("hn_rgb_upsample_bwd", lib.hn_rgb_upsample_bwd, _ptr(dy), _taps(ctx.f3), _ptr(dx), B * K, H, W, _stream())
        return dx, None


_NR_STATUS = {}


def _nr_status(dev):
    """One persistent status word per device for the consumer's kernels (non-zero only after a pipeline fault)."""
    st = _NR_STATUS.get(dev)
    if st is None:
        st = _NR_STATUS[dev] = torch.zeros(64, dtype=torch.int32, device=dev)
    return st


def _nr_args(x, ps, meta, saved, img, status):
    nb = meta["n_blocks"]
    a = L.NrFwd()
    a.B, a.n_blocks, a.feat_nc, a.min_feat, a.featmap_size = x.shape[0], nb, x.shape[1], meta["min_feat"], x.shape[2]
    a.final_actvn = 1 if meta["final_actvn"] else 0
    a.x = _ptr(x)
    for i in range(nb):
        a.w1[i], a.b1[i], a.w2[i], a.b2[i], a.wf[i], a.bf[i] = (p.data_ptr() for p in ps[6 * i:6 * i + 6])
        for k in range(3):
            a.tail_taps[i][k] = float(meta["tail_taps"][i][k])
    for j in range(nb + 1):
        a.wrgb[j], a.brgb[j] = ps[6 * nb + 2 * j].data_ptr(), ps[6 * nb + 2 * j + 1].data_ptr()
    for k in range(3):
        a.rgb_taps[k] = float(meta["rgb_taps"][k])
    a.saved, a.img, a.status = _ptr(saved), _ptr(img), _ptr(status)
    return a


class NeuralRenderFunction(torch.autograd.Function):
    """(x [B,feat_nc,fs,fs], meta, parameters) -> images [B,3,fs << n_blocks,fs << n_blocks]: the whole NeuralRenderer
    (NetWorks/neural_renderer.py:72-91, PixelShuffleUpsample.py:36-45) in one library call per direction (hn_nr_fwd / hn_nr_bwd,
    csrc/hn_nr.cu).  Parameters: per block layer_1.weight, layer_1.bias, layer_2.weight, layer_2.bias, feat_layers.weight,
    feat_layers.bias; then feat_2_rgb_list.j weight, bias.  meta: n_blocks, min_feat, final_actvn, tail_taps, rgb_taps and
    optionally grad_into (a list parallel to the parameters: existing .grad buffers the kernels accumulate into)."""

    @staticmethod
    def forward(ctx, x, meta, *params):
        lib = L.load()
        nb = meta["n_blocks"]
        if len(params) != 6 * nb + 2 * (nb + 1):
            raise ValueError("NeuralRenderFunction: wrong number of parameters")
        x = _dev_f32(x, "x", align=16)
        ps = [_dev_f32(p.detach(), "parameter", align=16) for p in params]
        B, C0, fs, fs2 = x.shape
        n_saved = lib.hn_nr_saved_floats(B, nb, C0, meta["min_feat"], fs) if fs == fs2 else -1
        if n_saved < 0:
            raise ValueError("NeuralRenderFunction: unsupported geometry")
        saved = torch.empty(n_saved, device=x.device)
        img = torch.empty(B, 3, fs << nb, fs << nb, device=x.device)
        status = _nr_status(x.device)
        a = _nr_args(x, ps, meta, saved, img, status)
        _call("hn_nr_fwd", lib.hn_nr_fwd, C.byref(a), _stream(), kernels=lib.hn_nr_launches(nb, 0))
        FAULTS.watch(status, "hn_nr_fwd")
        ctx.save_for_backward(x, saved, img, *ps)
        ctx.meta, ctx.param_shapes = meta, [tuple(p.shape) for p in params]
        return img

    @staticmethod
    def backward(ctx, g_img):
        lib = L.load()
        x, saved, img, *ps = ctx.saved_tensors
        meta = ctx.meta
        nb = meta["n_blocks"]
        need = ctx.needs_input_grad
        dev = x.device
        g_img = _dev_f32(g_img, "g_img", align=16)
        B, C0, fs, _ = x.shape
        scratch = torch.empty(lib.hn_nr_scratch_floats(B, nb, C0, meta["min_feat"], fs), device=dev)
        g_x = torch.empty_like(x) if need[0] else None
        status = _nr_status(dev)
        b = L.NrBwd()
        b.f = _nr_args(x, ps, meta, saved, img, status)
        b.g_img, b.scratch, b.g_x = _ptr(g_img), _ptr(scratch), _ptr(g_x)
        # gradient destinations: the caller's accumulation buffers where given, else views of ONE zeroed buffer
        into = meta.get("grad_into")
        n_p = len(ps)
        pair_need = [need[2 + 2 * k] or need[3 + 2 * k] for k in range(n_p // 2)]
        dst = [_grad_dst(into, k, ps[k]) if need[2 + k] else None for k in range(n_p)]
        own = [pair_need[k // 2] and dst[k] is None for k in range(n_p)]
        zbuf = torch.zeros(sum((ps[k].numel() + 3) // 4 * 4 for k in range(n_p) if own[k]), device=dev)
        off, buf = 0, list(dst)
        for k in range(n_p):
            if own[k]:
                buf[k] = zbuf[off:off + ps[k].numel()]
                off += (ps[k].numel() + 3) // 4 * 4
        for i in range(nb):
            if pair_need[3 * i]:
                b.dw1[i], b.db1[i] = buf[6 * i].data_ptr(), buf[6 * i + 1].data_ptr()
            if pair_need[3 * i + 1]:
                b.dw2[i], b.db2[i] = buf[6 * i + 2].data_ptr(), buf[6 * i + 3].data_ptr()
            if pair_need[3 * i + 2]:
                b.dwf[i], b.dbf[i] = buf[6 * i + 4].data_ptr(), buf[6 * i + 5].data_ptr()
        for j in range(nb + 1):
            if pair_need[3 * nb + j]:
                b.dwrgb[j], b.dbrgb[j] = buf[6 * nb + 2 * j].data_ptr(), buf[6 * nb + 2 * j + 1].data_ptr()
        _call("hn_nr_bwd", lib.hn_nr_bwd, C.byref(b), _stream(), kernels=lib.hn_nr_launches(nb, 1))
        FAULTS.watch(status, "hn_nr_bwd")
        grads = [buf[k].view(ctx.param_shapes[k]) if (need[2 + k] and dst[k] is None) else None for k in range(n_p)]
        return (g_x, None, *grads)


# ---------------------------------------------------------------------------------------------------
# debugging / test helpers: decode operand images (csrc/hn_tc.cuh layout) back to dense matrices
# ---------------------------------------------------------------------------------------------------
def _image_index(device):
    r = torch.arange(128, device=device).view(128, 1)
    c = torch.arange(64, device=device).view(1, 64)
    return ((r // 8) * 512 + (r % 8) * 64 + (((c // 8) ^ (r % 8)) * 8) + (c % 8)).reshape(-1)


def decode_image(buf: torch.Tensor, first_block: int, n_blocks: int, n_tiles: int) -> torch.Tensor:
    """uint8 image buffer [blocks, n_tiles, 16 KiB] -> float32 [n_tiles*128, 64*n_blocks]."""
    halves = buf.view(torch.float16).view(-1, n_tiles, 8192)
    idx = _image_index(buf.device)
    cols = [halves[first_block + k][:, idx].view(n_tiles * 128, 64) for k in range(n_blocks)]
    return torch.cat(cols, dim=1).float()


def decode_masks(masks: torch.Tensor, word0: int, n_cols: int, M: int) -> torch.Tensor:
    """int32 mask buffer [M, MASK_WORDS] -> bool [M, n_cols]."""
    w = masks.view(M, L.MASK_WORDS)[:, word0:word0 + (n_cols + 31) // 32].long() & 0xFFFFFFFF
    bits = (w.unsqueeze(-1) >> torch.arange(32, device=masks.device)) & 1
    return bits.reshape(M, -1)[:, :n_cols].bool()
