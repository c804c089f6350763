"""filter2d with kornia 0.6.12 semantics as documented: kernel [1,kh,kw]; `normalized` divides it by
its absolute sum; 'reflect' border; 'same' padding; depthwise correlation. Test infrastructure only."""
import torch
import torch.nn.functional as F


def filter2d(input, kernel, border_type="reflect", normalized=False, padding="same"):
    k = kernel
    if normalized:
        k = k / k.abs().sum(dim=(-2, -1), keepdim=True)
    kh, kw = k.shape[-2:]
    c = input.shape[1]
    x = F.pad(input, (kw // 2, kw // 2, kh // 2, kh // 2), mode=border_type)
    w = k.to(input).reshape(1, 1, kh, kw).expand(c, 1, kh, kw)
    return F.conv2d(x, w, groups=c)
