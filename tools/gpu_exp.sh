#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -k "tie_flooded" 2>&1 | tail -12
C2M_CORR_DEBUG=1 python - <<'PY' 2>&1 | grep "c2m corr" | tail -2
import sys; sys.path[:0]=['.','c2-matching_b200','tests/golden']
import torch, torch.nn.functional as F
import c2m_b200 as c2m
g = torch.Generator().manual_seed(21)
def smooth(c,h,w):
    base=torch.rand(1,c,5,5,generator=g); return (F.interpolate(base,size=(h,w),mode='bicubic',align_corners=False)+0.002*torch.randn(1,c,h,w,generator=g))
c2m.corr_argmax(smooth(64,40,40).cuda(), smooth(64,44,44).cuda(), norm_input=True); torch.cuda.synchronize()
PY
