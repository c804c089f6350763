"""The training-step pieces either side of the rendering path, behind the reference's own
interfaces:

  HeadNeRFLossUtils   Utils/HeadNeRFLossUtils.py:67-236 - calc_data_loss / calc_total_loss (bg, head and non-head MSE terms) as
                      one reduction kernel + one gradient kernel of libheadnerf_b200.so (hn_photo_loss_fwd / _bwd)
  FusedAdam           talker_trainer.py:722-723,1062-1067 - torch.optim.Adam(model.parameters(), lr): every parameter, its
                      gradient (dist.GradBucket) and both moments live in flat fp32 buffers, one hn_adam_step launch per step;
                      state_dict() / load_state_dict() keep torch.optim.Adam's format, so the checkpoints' "optim_state" loads
  Audio2style         talker_trainer.py:408-473 / FittingSingleImage_new.py:146-190 - the producer of `audiostyle` (2-layer
                      bidirectional LSTM + three Linear/LeakyReLU/Dropout stages), same parameter names, library (cuDNN) LSTM
  save_checkpoint / load_checkpoint   talker_trainer.py:915-936, FittingSingleImage_new.py:640-656 - the {"para", "net",
                      "audio2style", "optim_state", ...} .pth layout

There is no CPU implementation of the two kernels: CPU tensors raise."""
import ctypes as C
import math

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .options import BaseOptions


# ---------------------------------------------------------------------------------------------------------------------
# photometric loss
# ---------------------------------------------------------------------------------------------------------------------
class PhotoLossFunction(torch.autograd.Function):
    """(merge_img [B,3,H,W], bg_img [B_bg,3,H,W], gt_rgb [B,3,H,W], mask [B,1,H,W] float, bg_value) -> terms [4] =
    (bg_loss, head_loss, nonhead_loss, total); every entry is differentiable (total is their sum: use either, not both)."""

    @staticmethod
    def forward(ctx, img, bg_img, gt, mask, bg_value):
        lib = L.load()
        img, bg_img = ops._dev_f32(img, "merge_img"), ops._dev_f32(bg_img, "bg_img")
        gt, mask = ops._dev_f32(gt, "gt_rgb", img.shape), ops._dev_f32(mask, "mask")
        B, three, H, W = img.shape
        if three != 3 or tuple(mask.shape) != (B, 1, H, W) or bg_img.shape[1:] != img.shape[1:]:
            raise ValueError("photometric loss: merge_img / gt [B,3,H,W], mask [B,1,H,W], bg_img [B_bg,3,H,W]")
        ws = torch.zeros(lib.hn_photo_loss_workspace_bytes() // 4, dtype=torch.float32, device=img.device)
        out = torch.empty(8, device=img.device)
        a = L.PhotoLoss()
        a.B, a.B_bg, a.HW, a.bg_value = B, bg_img.shape[0], H * W, float(bg_value)
        a.img, a.bg_img, a.gt, a.mask = img.data_ptr(), bg_img.data_ptr(), gt.data_ptr(), mask.data_ptr()
        a.partials, a.ticket, a.out = ws.data_ptr(), ws.data_ptr() + ws.numel() * 4 - 16, out.data_ptr()
        ops._call("hn_photo_loss_fwd", lib.hn_photo_loss_fwd, C.byref(a), ops._stream())
        ctx.save_for_backward(img, bg_img, gt, mask, out)
        ctx.bg_value = float(bg_value)
        return out[:4].clone()

    @staticmethod
    def backward(ctx, g_terms):
        lib = L.load()
        img, bg_img, gt, mask, out = ctx.saved_tensors
        B, _, H, W = img.shape
        a = L.PhotoLoss()
        a.B, a.B_bg, a.HW, a.bg_value = B, bg_img.shape[0], H * W, ctx.bg_value
        a.img, a.bg_img, a.gt, a.mask, a.out = img.data_ptr(), bg_img.data_ptr(), gt.data_ptr(), mask.data_ptr(), out.data_ptr()
        gout = g_terms.contiguous().float()                     # dL/d(bg, head, nonhead, total)
        d_img = torch.empty_like(img) if ctx.needs_input_grad[0] else None
        d_bg = torch.empty_like(bg_img) if ctx.needs_input_grad[1] else None
        ops._call("hn_photo_loss_bwd", lib.hn_photo_loss_bwd, C.byref(a), ops._ptr(gout), ops._ptr(d_img), ops._ptr(d_bg), ops._stream())
        return d_img, d_bg, None, None, None


class HeadNeRFLossUtils(object):
    """Drop-in for Utils/HeadNeRFLossUtils.py:67-236 (the data terms the trainer sums; the VGG term needs torchvision weights that
    cannot be downloaded here and the code / camera regularisers are commented out in the reference's calc_total_loss)."""

    def __init__(self, bg_type="white", use_vgg_loss=True, device=None) -> None:
        super().__init__()
        if bg_type == "white":
            self.bg_value = 1.0
        elif bg_type == "black":
            self.bg_value = 0.0
        else:
            raise ValueError("Error BG type.")
        if use_vgg_loss:
            raise NotImplementedError("use_vgg_loss=True needs torchvision's pretrained VGG16 weights (HeadNeRFLossUtils.py:24-30); "
                                      "construct with use_vgg_loss=False for the MSE terms")
        self.use_vgg_loss = False
        self.device = device

    @staticmethod
    def calc_cam_loss(delta_cam_info):                          # HeadNeRFLossUtils.py:88-96
        return {"delta_eular": torch.mean(delta_cam_info["delta_eulur"] * delta_cam_info["delta_eulur"]),
                "delta_tvec": torch.mean(delta_cam_info["delta_tvec"] * delta_cam_info["delta_tvec"])}

    def calc_data_loss(self, data_dict, gt_rgb, head_mask_c1b, nonhead_mask_c1b):
        """head / non-head masks are the boolean [B,1,H,W] tensors of calc_total_loss (complementary); one fused reduction."""
        mask = head_mask_c1b.to(torch.float32)
        if nonhead_mask_c1b is not None and nonhead_mask_c1b.dtype == torch.bool and head_mask_c1b.dtype == torch.bool:
            # pixels in neither set (NaN mask values upstream) must drop out of both terms: encode them as NaN
            neither = ~(head_mask_c1b | nonhead_mask_c1b)
            mask = torch.where(neither, torch.full_like(mask, float("nan")), mask)
        terms = PhotoLossFunction.apply(data_dict["merge_img"], data_dict["bg_img"], gt_rgb, mask, self.bg_value)
        return {"bg_loss": terms[0], "head_loss": terms[1], "nonhaed_loss": terms[2]}

    def calc_total_loss(self, delta_cam_info, opt_code_dict, pred_dict, gt_rgb, mask_tensor, disp_pred_dict=None, eye_mask_tensor=None):
        """HeadNeRFLossUtils.py:196-236: head = mask >= 0.5, non-head = mask < 0.5; total = the sum of the three data terms (every
        other term of the reference's total is commented out there).  "nonhaed_loss" is the reference's own key."""
        terms = PhotoLossFunction.apply(pred_dict["coarse_dict"]["merge_img"], pred_dict["coarse_dict"]["bg_img"], gt_rgb,
                                        mask_tensor.to(torch.float32), self.bg_value)
        return {"bg_loss": terms[0].detach(), "head_loss": terms[1].detach(), "nonhaed_loss": terms[2].detach(), "total_loss": terms[3]}


# ---------------------------------------------------------------------------------------------------------------------
# optimizer
# ---------------------------------------------------------------------------------------------------------------------
class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam(params, lr, betas, eps, weight_decay) with flat storage and one kernel per step.

    On construction every parameter's `.data` is re-pointed into ONE flat fp32 buffer (values preserved) and, unless the
    parameters' gradients already are views of a flat buffer (`bucket=` a dist.GradBucket), a flat gradient buffer is created
    the same way.  `step(grad_scale=...)` is a single hn_adam_step launch; `grad_scale = 1 / world_size` folds the
    data-parallel averaging of an all-reduced (summed) bucket into it.  One param group; amsgrad / maximize unsupported."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, bucket=None):
        params = [p for p in params if p.requires_grad]
        if not params:
            raise ValueError("FusedAdam: no trainable parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) != 1:
            raise ValueError("FusedAdam keeps one flat buffer: one param group")
        ps = self.param_groups[0]["params"]
        dev = ps[0].device
        if dev.type != "cuda" or any(p.dtype != torch.float32 or p.device != dev for p in ps):
            raise L.HeadNeRFLibraryError("FusedAdam needs float32 CUDA parameters on one device (there is no CPU implementation)")
        from .dist import flat_layout
        self._offsets, n = flat_layout(ps)                      # every tensor starts on a 256-byte boundary; the gaps hold zeros
        self._n = n
        pad = 0
        self.flat_params = torch.zeros(n, device=dev)
        for p, off in zip(ps, self._offsets):
            self.flat_params[off:off + p.numel()].copy_(p.data.reshape(-1))
            p.data = self.flat_params[off:off + p.numel()].view_as(p)
        if bucket is not None:
            if [id(p) for p in bucket.params] != [id(p) for p in ps] or list(bucket.offsets) != list(self._offsets):
                raise ValueError("FusedAdam: the gradient bucket must hold the same parameters in the same order")
            self.flat_grads = bucket.flat
        else:
            self.flat_grads = torch.zeros(n, device=dev)
            for p, off in zip(ps, self._offsets):
                p.grad = self.flat_grads[off:off + p.numel()].view_as(p)
        self.exp_avg = torch.zeros(n + pad, device=dev)
        self.exp_avg_sq = torch.zeros(n + pad, device=dev)
        self._step = 0
        self._publish_state()

    def _publish_state(self):
        """torch.optim.Adam's per-parameter state layout, as views of the flat moments (so state_dict() is interchangeable)."""
        for p, off in zip(self.param_groups[0]["params"], self._offsets):
            k = p.numel()
            self.state[p] = {"step": torch.tensor(float(self._step)), "exp_avg": self.exp_avg[off:off + k].view_as(p),
                             "exp_avg_sq": self.exp_avg_sq[off:off + k].view_as(p)}

    def zero_grad(self, set_to_none: bool = False):
        self.flat_grads.zero_()                                 # one memset; the views stay in place

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = L.load()
        g = self.param_groups[0]
        self._step += 1
        h = L.Adam()
        h.lr, h.beta1, h.beta2, h.eps = float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"])
        h.weight_decay, h.grad_scale, h.step = float(g["weight_decay"]), float(grad_scale), self._step
        ops._call("hn_adam_step", lib.hn_adam_step, ops._ptr(self.flat_params), ops._ptr(self.flat_grads), ops._ptr(self.exp_avg),
                  ops._ptr(self.exp_avg_sq), self._n, C.byref(h), ops._stream())
        for st in self.state.values():
            st["step"].fill_(float(self._step))
        # the kernel wrote the parameters behind autograd's back: bump their version counters, which is what every cache keyed on
        # (data_ptr, _version) - HeadNeRFNet's packed weight operands among them - watches
        torch.autograd.graph.increment_version(self.param_groups[0]["params"])
        return loss

    def load_state_dict(self, state_dict):
        """Accepts torch.optim.Adam's state dict (the checkpoints' "optim_state", talker_trainer.py:932) and this class's own."""
        super().load_state_dict(state_dict)                     # replaces self.state[p] tensors by loaded copies ...
        step = 0
        for p, off in zip(self.param_groups[0]["params"], self._offsets):
            st = self.state.get(p, {})
            k = p.numel()
            if "exp_avg" in st:                                 # ... which go back into the flat buffers
                self.exp_avg[off:off + k].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
                step = max(step, int(float(st["step"])))
            else:
                self.exp_avg[off:off + k].zero_()
                self.exp_avg_sq[off:off + k].zero_()
        self._step = step
        self._publish_state()


# ---------------------------------------------------------------------------------------------------------------------
# audio front-end
# ---------------------------------------------------------------------------------------------------------------------
class RNNModel(nn.Module):
    """talker_trainer.py:408-425: 2-layer bidirectional LSTM; `fc1` exists in the state dict but is never applied."""

    def __init__(self, input_size=256, hidden_size=256, num_layers=2, batch_first=True, bidirectional=True):
        super().__init__()
        self.nhid, self.nlayers = hidden_size, num_layers
        self.rnn = nn.LSTM(input_size, hidden_size, num_layers, batch_first=True, bidirectional=True)
        if bidirectional:
            self.fc1 = nn.Linear(hidden_size * 2, hidden_size)

    def forward(self, inputs):
        output, _ = self.rnn(inputs)
        return output


class Audio2style(nn.Module):
    """talker_trainer.py:428-473: mel window [T,80,16] -> audiostyle [T,64] (the 64-wide code of NetWorks/models.py:32).  The T
    frames form ONE sequence of the LSTM (`unsqueeze(0)`); Dropout(0.5) is active in train mode exactly as in the reference."""

    def __init__(self, hidden_size=128):
        super().__init__()
        self.flatten = nn.Flatten()
        self.rnn = RNNModel(80 * 16, 40 * 16)
        self.linear1 = nn.Sequential(nn.Linear(80 * 16, 40 * 16), nn.LeakyReLU(0.2, True), nn.Dropout(p=0.5))
        self.linear2 = nn.Sequential(nn.Linear(40 * 16, 20 * 16), nn.LeakyReLU(0.2, True), nn.Dropout(p=0.5))
        self.linear3 = nn.Sequential(nn.Linear(20 * 16, 64), nn.LeakyReLU(0.2, True), nn.Dropout(p=0.5))

    def forward(self, audio_inputs):
        x = self.flatten(audio_inputs)
        x = self.rnn(x.unsqueeze(0))
        x = self.linear1(x[0])
        x = self.linear2(x)
        return self.linear3(x)


# ---------------------------------------------------------------------------------------------------------------------
# checkpoints
# ---------------------------------------------------------------------------------------------------------------------
def save_checkpoint(path, net, epoch=0, audio2style=None, optimizer=None, optimizer_style=None, scheduler=None):
    """The dictionary talker_trainer.py:923-936 saves: {"epoch", "net", "para", ["audio2style", "optim_state", "optim_style",
    "scheule_state"]} ("scheule_state" is the reference's own spelling)."""
    ck = {"epoch": epoch, "net": net.state_dict(),
          "para": {"featmap_size": net.opt.featmap_size, "featmap_nc": net.opt.featmap_nc, "pred_img_size": net.opt.pred_img_size}}
    if audio2style is not None:
        ck["audio2style"] = audio2style.state_dict()
    if optimizer is not None:
        ck["optim_state"] = optimizer.state_dict()
    if optimizer_style is not None:
        ck["optim_style"] = optimizer_style.state_dict()
    if scheduler is not None:
        ck["scheule_state"] = scheduler.state_dict()
    torch.save(ck, path)
    return ck


def load_checkpoint(path, device="cpu", include_gaze=False, eye_gaze_dim=2):
    """FittingSingleImage_new.py:640-656 (build_info): para -> BaseOptions -> HeadNeRFNet(include_vd=False, hier_sampling=False),
    strict load of "net", Audio2style from "audio2style" when present.  -> (net, audio2style or None, the raw dictionary)."""
    from .headnerf_net import HeadNeRFNet
    ck = torch.load(path, map_location=torch.device("cpu"))
    opt = BaseOptions(ck["para"])
    net = HeadNeRFNet(opt, include_vd=False, hier_sampling=False, include_gaze=include_gaze, eye_gaze_dim=eye_gaze_dim)
    net.load_state_dict(ck["net"])
    a2s = None
    if "audio2style" in ck:
        a2s = Audio2style()
        a2s.load_state_dict(ck["audio2style"])
        a2s = a2s.to(device)
    return net.to(device).eval(), a2s, ck
