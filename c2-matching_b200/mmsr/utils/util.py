"""tensor2img as used by the PSNR check (reference: mmsr/utils/util.py:107-162)."""
import math

import numpy as np
import torch


def tensor2img(tensor, out_type=np.uint8, min_max=(0, 1)):
    """RGB [0,1] CHW tensor(s) -> BGR HWC array(s), rounded to [0,255] (NOT cast, as in the
    reference, whose `img_np.astype(out_type)` result is discarded, util.py:157)."""
    single = torch.is_tensor(tensor)
    if not single and not (isinstance(tensor, (list, tuple)) and all(torch.is_tensor(t) for t in tensor)):
        raise TypeError(f'tensor or list of tensors expected, got {type(tensor)}')
    outs = []
    for t in [tensor] if single else tensor:
        t = t.squeeze(0).float().detach().cpu().clamp(*min_max)
        t = (t - min_max[0]) / (min_max[1] - min_max[0])
        if t.dim() == 4:
            from torchvision.utils import make_grid
            t = make_grid(t, nrow=int(math.sqrt(t.size(0))), normalize=False)
        if t.dim() == 3:
            # RGB CHW -> BGR HWC and the x255 rounding in torch (contiguous, GIL-free; same float32 arithmetic and the
            # same round-half-to-even as numpy's fancy-index + strided multiply, which took 80 ms per 640x640 image)
            t = t.flip(0).permute(1, 2, 0).contiguous()
        elif t.dim() != 2:
            raise TypeError(f'Only support 4D, 3D or 2D tensor. But received with dimension: {t.dim()}')
        if out_type == np.uint8:
            t = (t * 255.0).round()
        outs.append(t.numpy())
    return outs[0] if len(outs) == 1 else outs
