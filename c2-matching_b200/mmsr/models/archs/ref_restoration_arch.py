"""Restoration generator (mmsr/models/archs/ref_restoration_arch.py:8-187): LR content features,
then three scales of [offset convs -> dynamic aggregation (DCN over Ref features) -> head ->
16 ResBlocks -> upsample], plus a bilinear x4 skip.

Same constructor (`ngf, n_blocks, groups`) and state-dict keys as the reference, so
`c2_matching_restoration_*.pth` loads with strict=True.  Differences are execution only: each
dynamic-aggregation tail is one fused kernel (see DCNv2/dcn_v2.py) including its LeakyReLU, and
the pre-offsets are read straight from the index map when CorrespondenceGenerationArch
provides a PreOffsets handle."""
import torch
import torch.nn.functional as F
from torch import nn

from . import arch_util
from .DCNv2.dcn_v2 import DCN_sep_pre_multi_offset as DynAgg

_LEVELS = (('small', 'relu3_1', 256), ('medium', 'relu2_1', 128), ('large', 'relu1_1', 64))


class ContentExtractor(nn.Module):

    def __init__(self, in_nc=3, out_nc=3, nf=64, n_blocks=16):
        super().__init__()
        self.conv_first = nn.Conv2d(in_nc, nf, 3, 1, 1)
        self.body = arch_util.make_layer(arch_util.ResidualBlockNoBN, n_blocks, nf=nf)
        arch_util.default_init_weights([self.conv_first], 0.1)

    def forward(self, x):
        from c2m_b200 import ops
        if x.is_cuda and arch_util.psa_conv_ok(self.conv_first, x.shape[2], x.shape[3]) and \
                arch_util.psa_conv_ok(self.body[0].conv1, x.shape[2], x.shape[3]):
            hp = ops.conv3x3_psa(ops.psa_from_f32(x), self.conv_first.weight, self.conv_first.bias, act='lrelu')
            return ops.psa_to_f32(arch_util.resblocks_psa(self.body, hp))
        return self.body(F.leaky_relu(self.conv_first(x), 0.1))


class DynamicAggregationRestoration(nn.Module):

    def __init__(self, ngf=64, n_blocks=16, groups=8):
        super().__init__()
        for size, _, c in _LEVELS:
            setattr(self, f'{size}_offset_conv1', nn.Conv2d(ngf + c, c, 3, 1, 1, bias=True))
            setattr(self, f'{size}_offset_conv2', nn.Conv2d(c, c, 3, 1, 1, bias=True))
            setattr(self, f'{size}_dyn_agg', DynAgg(c, c, 3, stride=1, padding=1, dilation=1,
                                                    deformable_groups=groups, extra_offset_mask=True))
            setattr(self, f'head_{size}', nn.Sequential(nn.Conv2d(ngf + c, ngf, 3, 1, 1), nn.LeakyReLU(0.1, True)))
            setattr(self, f'body_{size}', arch_util.make_layer(arch_util.ResidualBlockNoBN, n_blocks, nf=ngf))
        self.tail_small = nn.Sequential(nn.Conv2d(ngf, ngf * 4, 3, 1, 1), nn.PixelShuffle(2), nn.LeakyReLU(0.1, True))
        self.tail_medium = nn.Sequential(nn.Conv2d(ngf, ngf * 4, 3, 1, 1), nn.PixelShuffle(2), nn.LeakyReLU(0.1, True))
        self.tail_large = nn.Sequential(nn.Conv2d(ngf, ngf // 2, 3, 1, 1), nn.LeakyReLU(0.1, True),
                                        nn.Conv2d(ngf // 2, 3, 3, 1, 1))

    def forward(self, x, pre_offset, img_ref_feat):
        for size, key, _ in _LEVELS:
            ref = img_ref_feat[key]
            off = torch.cat([x, ref], 1)
            off = F.leaky_relu_(getattr(self, f'{size}_offset_conv1')(off), 0.1)
            off = F.leaky_relu_(getattr(self, f'{size}_offset_conv2')(off), 0.1)
            pre = pre_offset.handle(key) if hasattr(pre_offset, 'handle') else pre_offset[key]
            swapped = getattr(self, f'{size}_dyn_agg')([ref, off], pre, lrelu_slope=0.1)   # lrelu fused
            h = getattr(self, f'head_{size}')(torch.cat([x, swapped], 1))
            h = arch_util.body_forward(getattr(self, f'body_{size}'), h, skip=x)
            x = self._tail(size, h)
        return x

    def _tail(self, size, h):
        tail = getattr(self, f'tail_{size}')
        if size == 'large' and h.is_cuda and arch_util.psa_conv_ok(tail[0], h.shape[2], h.shape[3]) and \
                arch_util.psa_conv_ok(tail[2], h.shape[2], h.shape[3]):
            from c2m_b200 import ops
            t = ops.conv3x3_psa(ops.psa_from_f32(h), tail[0].weight, tail[0].bias, act='lrelu')
            return ops.psa_to_f32(ops.conv3x3_psa(t, tail[2].weight, tail[2].bias))
        return tail(h)


class RestorationNet(nn.Module):

    def __init__(self, ngf=64, n_blocks=16, groups=8):
        super().__init__()
        self.content_extractor = ContentExtractor(in_nc=3, out_nc=3, nf=ngf, n_blocks=n_blocks)
        self.dyn_agg_restore = DynamicAggregationRestoration(ngf, n_blocks, groups)
        arch_util.srntt_init_weights(self, init_type='normal', init_gain=0.02)
        self.re_init_dcn_offset()

    def re_init_dcn_offset(self):
        for size, _, _ in _LEVELS:
            getattr(self.dyn_agg_restore, f'{size}_dyn_agg').init_offset()

    def forward(self, x, pre_offset, img_ref_feat):
        base = F.interpolate(x, None, 4, 'bilinear', False)
        content_feat = self.content_extractor(x)
        return self.dyn_agg_restore(content_feat, pre_offset, img_ref_feat) + base
