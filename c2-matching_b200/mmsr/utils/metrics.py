"""PSNR / SSIM / Y-channel conversion with the reference's definitions
(mmsr/utils/metrics.py:34-66, 69-143, 146-168) — the metric the 0.01 dB parity bar is stated in."""
import numpy as np


def _hwc(img, input_order='HWC'):
    if input_order not in ('HWC', 'CHW'):
        raise ValueError(f'Wrong input_order {input_order}. Supported input_orders are "HWC" and "CHW"')
    if img.ndim == 2:
        return img[..., None]
    return img.transpose(1, 2, 0) if input_order == 'CHW' else img


def psnr(img1, img2, crop_border=0, input_order='HWC'):
    assert img1.shape == img2.shape, f'Image shapes are differnet: {img1.shape}, {img2.shape}.'
    a, b = _hwc(img1, input_order), _hwc(img2, input_order)
    if crop_border:
        a = a[crop_border:-crop_border, crop_border:-crop_border, None]
        b = b[crop_border:-crop_border, crop_border:-crop_border, None]
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return float('inf') if mse == 0 else 20.0 * np.log10(255.0 / np.sqrt(mse))


def _ssim_1ch(img1, img2):
    import cv2
    c1, c2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    img1, img2 = img1.astype(np.float64), img2.astype(np.float64)
    k = cv2.getGaussianKernel(11, 1.5)
    win = np.outer(k, k.transpose())
    f = lambda x: cv2.filter2D(x, -1, win)[5:-5, 5:-5]
    mu1, mu2 = f(img1), f(img2)
    s1, s2, s12 = f(img1 ** 2) - mu1 ** 2, f(img2 ** 2) - mu2 ** 2, f(img1 * img2) - mu1 * mu2
    return (((2 * mu1 * mu2 + c1) * (2 * s12 + c2)) / ((mu1 ** 2 + mu2 ** 2 + c1) * (s1 + s2 + c2))).mean()


def ssim(img1, img2, crop_border=0, input_order='HWC'):
    assert img1.shape == img2.shape
    a, b = _hwc(img1, input_order), _hwc(img2, input_order)
    if crop_border:
        a = a[crop_border:-crop_border, crop_border:-crop_border]
        b = b[crop_border:-crop_border, crop_border:-crop_border]
    return float(np.mean([_ssim_1ch(a[..., i], b[..., i]) for i in range(a.shape[2])]))


def bgr2ycbcr(img, only_y=True):
    """ITU-R BT.601 as MATLAB's rgb2ycbcr; uint8 in [0,255] or float in [0,1] (metrics.py:146-168)."""
    in_type = img.dtype
    # the reference keeps the caller's float dtype (its astype(float32) result is discarded, :154)
    x = img.astype(np.float64) if in_type == np.uint8 else img * 255.0
    if only_y:
        out = np.dot(x, [24.966, 128.553, 65.481]) / 255.0 + 16.0
    else:
        out = np.matmul(x, [[24.966, 112.0, -18.214], [128.553, -74.203, -93.786],
                            [65.481, -37.797, 112.0]]) / 255.0 + [16, 128, 128]
    return out.round().astype(in_type) if in_type == np.uint8 else (out / 255.0).astype(in_type)
