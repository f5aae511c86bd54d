"""The evaluation metrics of `metrics.py` (reference: mmsr/utils/metrics.py:34-66, 69-143, 146-168) and the
`tensor2img` quantisation in front of them (mmsr/utils/util.py:107-162), computed where the SR batch already is.

Same definitions and the same dtypes step by step as the host versions (float32 image values, float64 Y-channel dot
product, float64 PSNR / SSIM arithmetic); what differs is only the summation order inside the means and inside the
11x11 Gaussian window, i.e. ~1e-15 relative.  `tests/test_host_cpu.py::test_metrics_torch_matches_host_metrics`
pins the two against each other."""
import math

import torch
import torch.nn.functional as F

_Y_COEF = (24.966, 128.553, 65.481)        # applied to (B, G, R), metrics.py:160


def _gauss_window(device):
    # cv2.getGaussianKernel(11, 1.5): exp(-(i-5)^2 / (2 sigma^2)) normalised to sum 1, in float64
    g = torch.tensor([math.exp(-((i - 5) ** 2) / (2 * 1.5 ** 2)) for i in range(11)], dtype=torch.float64)
    g = g / g.sum()
    return torch.outer(g, g).to(device)[None, None]


def _div(x, c):
    """IEEE division by a constant: `tensor / python_scalar` is evaluated as a multiplication by the reciprocal on
    CUDA, which rounds differently from numpy's true division in the last bit."""
    return x / torch.full((), c, dtype=x.dtype, device=x.device)


def quantise(img, min_max=(0, 1)):
    """tensor2img's value path for a [3,H,W] RGB tensor: clamp, rescale, x255, round-half-even; stays float32 RGB CHW."""
    t = img.float().clamp(*min_max)
    t = _div(t - min_max[0], float(min_max[1] - min_max[0]))
    return (t * 255.0).round()


def _y_channel(q):
    """`bgr2ycbcr(img / 255., only_y=True) * 255` of the host path for a float32 RGB CHW image with values in
    [0,255]: float32 /255 and x255, float64 dot product, float32 result."""
    x = (_div(q, 255.0) * 255.0).double()
    y = _div((x[2] * _Y_COEF[0] + x[1] * _Y_COEF[1]) + x[0] * _Y_COEF[2], 255.0) + 16.0
    return _div(y, 255.0).float() * 255.0


def _psnr(a, b):
    mse = ((a.double() - b.double()) ** 2).mean()
    return torch.where(mse == 0, torch.full_like(mse, float('inf')), 20.0 * torch.log10(255.0 / torch.sqrt(mse)))


def _ssim_1ch(a, b, win):
    c1, c2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    a, b = a.double(), b.double()
    stack = torch.stack([a, b, a * a, b * b, a * b])[:, None]          # [5,1,H,W]
    # cv2.filter2D(...)[5:-5, 5:-5] never touches the border: a 'valid' correlation with the symmetric window
    mu1, mu2, e11, e22, e12 = F.conv2d(stack, win)[:, 0]
    s1, s2, s12 = e11 - mu1 ** 2, e22 - mu2 ** 2, e12 - mu1 * mu2
    return (((2 * mu1 * mu2 + c1) * (2 * s12 + c2)) / ((mu1 ** 2 + mu2 ** 2 + c1) * (s1 + s2 + c2))).mean()


def score_image(sr, gt, crop_border=0, valid_hw=None):
    """One (SR, GT) pair of [3,H,W] RGB tensors in [0,1] -> float64 tensor [psnr, psnr_y, ssim_y, finite] on their
    device (no host synchronisation).  `valid_hw` = the un-padded (h, w) when the pair was zero-padded."""
    finite = torch.isfinite(sr).all()
    a, b = quantise(sr), quantise(gt)
    if valid_hw is not None:
        a, b = a[:, :valid_hw[0], :valid_hw[1]], b[:, :valid_hw[0], :valid_hw[1]]
    ya, yb = _y_channel(a), _y_channel(b)
    if crop_border:
        c = crop_border
        a, b, ya, yb = a[:, c:-c, c:-c], b[:, c:-c, c:-c], ya[c:-c, c:-c], yb[c:-c, c:-c]
    win = _gauss_window(sr.device)
    return torch.stack([_psnr(a, b), _psnr(ya, yb), _ssim_1ch(ya, yb, win), finite.double()])
