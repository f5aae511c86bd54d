"""ORACLE — test infrastructure, NOT product code.

CPU restatement (PyTorch fp32 on host cores) of the reference's restoration-forward hot
path, written as pure functions over state dicts.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / `--impl reference` legs may import this file; nothing under
c2-matching_b200/ does.

Parity status: PINNED against outputs of the unmodified reference run in the authoring
container (tests/golden/*.npz, minted by tests/golden/make_golden.py; the reference itself
ships no golden vectors — SURVEY.md §4).  tests/test_oracle.py checks every function here
against those fixtures.

Each function cites the reference lines (relative to /root/reference) it restates.  The
algorithm is kept as the reference has it — materialised patch tensor, chunked
`conv2d` + `max` with a strict-`>` merge, Python loop over the batch, unfused DCN — because
this file doubles as the "reference CPU path" timed beside the GPU kernels.

The reference's `_ext.dcn_v2_forward` has no CPU implementation (DCNv2/src/dcn_v2.h:38) and
no longer builds (THC); its arithmetic is restated literally in oracle/c2m_oracle.c
(`oracle_dcn_v2_forward`) and, for speed on large maps, by torchvision's CPU
`deform_conv2d`, which tests/test_oracle.py shows agrees with the literal restatement.
"""
import torch
import torch.nn.functional as F

try:  # torchvision is only the fast DCN stand-in; the C oracle is the literal one
    from torchvision.ops import deform_conv2d as _tv_deform_conv2d
except Exception:  # pragma: no cover
    _tv_deform_conv2d = None


# ----------------------------------------------------------------------------- a2 / a3
def sample_patches(x, patch_size=3, stride=1):
    """mmsr/models/archs/ref_map_util.py:4-23 — [C,h,w] -> [C,p,p,N], row-major patches."""
    c = x.shape[0]
    u = x.unfold(1, patch_size, stride).unfold(2, patch_size, stride)  # [C, nh, nw, p, p]
    return u.reshape(c, -1, patch_size, patch_size).permute(0, 2, 3, 1)


def feature_match_index(feat_input, feat_ref, patch_size=3, input_stride=1, ref_stride=1,
                        is_norm=True, norm_input=False):
    """mmsr/models/archs/ref_map_util.py:26-86.

    Returns (max_idx int64 [h',w'], max_val fp32 [h',w']); max_idx is the flat index into the
    Ref patch grid; ties resolve to the lowest Ref index (first-max inside a chunk, strict
    `>` across chunks, :69-76)."""
    patches_ref = sample_patches(feat_ref, patch_size, ref_stride)
    _, h, w = feat_input.shape
    chunk = int(1024.0 ** 2 * 512 / (h * w))                     # :56
    n = patches_ref.shape[-1]
    best_idx = best_val = None
    for start in range(0, n, chunk):                             # :60
        blk = patches_ref[..., start:start + chunk]
        if is_norm:
            blk = blk / (blk.norm(p=2, dim=(0, 1, 2)) + 1e-5)    # :62-63
        corr = F.conv2d(feat_input.unsqueeze(0), blk.permute(3, 0, 1, 2), stride=input_stride)
        v, i = corr.squeeze(0).max(dim=0)                        # :69
        if best_idx is None:
            best_idx, best_val = i, v
        else:
            upd = v > best_val                                   # :74
            best_val[upd] = v[upd]
            best_idx[upd] = i[upd] + start
    if norm_input:                                               # :78-84
        pin = sample_patches(feat_input, patch_size, input_stride)
        nrm = pin.norm(p=2, dim=(0, 1, 2)) + 1e-5
        nrm = nrm.view(int((h - patch_size) / input_stride + 1),
                       int((w - patch_size) / input_stride + 1))
        best_val = best_val / nrm
    return best_idx, best_val


# ----------------------------------------------------------------------------- a4 / a5
def index_to_flow(max_idx):
    """mmsr/models/archs/corres_generation_arch.py:29-46 — idx -> [1,h+2,w+2,2] (x,y) flow."""
    h, w = max_idx.shape
    fx = (max_idx % w).float()                                   # NB: INPUT grid width (:32-34)
    fy = (max_idx // w).float()
    gy, gx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing='ij')
    flow = torch.stack((fx - gx, fy - gy), dim=2).unsqueeze(0)
    return F.pad(flow, (0, 0, 0, 2, 0, 2))


def tensor_shift(x, shift):
    """mmsr/models/archs/arch_util.py:291-315 — shift [b,h,w,c] down/right, zero fill."""
    sh, sw = shift
    out = torch.zeros_like(x)
    h, w = x.shape[1], x.shape[2]
    out[:, sh:, sw:, :] = x[:, :h - sh, :w - sw, :]
    return out


def offset_pyramid(flow3):
    """corres_generation_arch.py:70-104 — the 9-shift x 3-scale pre-offset stack for one image."""
    out = {}
    for name, s in (('relu3_1', 1), ('relu2_1', 2), ('relu1_1', 4)):
        f = flow3
        if s > 1:
            f = torch.repeat_interleave(torch.repeat_interleave(flow3, s, 1), s, 2) * s
        out[name] = torch.cat([tensor_shift(f, (i * s, j * s)) for i in range(3) for j in range(3)], 0)
    return out


def correspondence(feat1, feat2, patch_size=3, stride=1, return_idx=False):
    """corres_generation_arch.py:48-114 (everything except the VGG19 call at :116)."""
    per_scale = {'relu3_1': [], 'relu2_1': [], 'relu1_1': []}
    idxs = []
    for b in range(feat1.shape[0]):                              # :52 Python batch loop
        fi, fr = feat1[b], feat2[b]
        c, h, w = fi.shape
        fi = F.normalize(fi.reshape(c, -1), dim=0).view(c, h, w)  # :56
        fr = F.normalize(fr.reshape(c, -1), dim=0).view(c, h, w)  # :57-58 (viewed with INPUT h,w)
        idx, _ = feature_match_index(fi, fr, patch_size, stride, stride, is_norm=True, norm_input=True)
        idxs.append(idx)
        pyr = offset_pyramid(index_to_flow(idx))
        for k in per_scale:
            per_scale[k].append(pyr[k])
    pre = {k: torch.stack(v, 0) for k, v in per_scale.items()}
    return (pre, torch.stack(idxs)) if return_idx else pre


# ----------------------------------------------------------------------------- a7 / a8
def dcn_v2_forward(x, weight, bias, offset, mask, kh=3, kw=3, sh=1, sw=1, ph=1, pw=1, dh=1, dw=1, dg=8):
    """`_ext.dcn_v2_forward` (DCNv2/src/cuda/dcn_v2_cuda.cu:42-172, im2col kernel
    dcn_v2_im2col_cuda.cu:125-195).  Fast stand-in; literal restatement: c2m_oracle.c."""
    if _tv_deform_conv2d is None:
        raise RuntimeError('torchvision missing: use oracle.c_oracle.dcn_v2_forward')
    return _tv_deform_conv2d(x, offset, weight, bias, stride=(sh, sw), padding=(ph, pw),
                             dilation=(dh, dw), mask=mask)


def dcn_sep_pre_multi_offset(sd, prefix, x, feat, pre_offset, dg):
    """DCNv2/dcn_v2.py:222-253 — offsets = conv_offset_mask(feat)[:2/3] + reordered pre_offset,
    mask = sigmoid(last 1/3), then the modulated deformable 3x3 conv."""
    out = F.conv2d(feat, sd[prefix + 'conv_offset_mask.weight'], sd[prefix + 'conv_offset_mask.bias'], 1, 1)
    o1, o2, m = torch.chunk(out, 3, dim=1)
    offset = torch.cat((o1, o2), dim=1)
    pre = pre_offset.repeat([1, dg, 1, 1, 1])                    # [b, 9*dg, h, w, 2] (x,y)
    reord = torch.zeros_like(offset)
    reord[:, 0::2] = pre[..., 1]                                 # even channels: y
    reord[:, 1::2] = pre[..., 0]                                 # odd channels:  x
    return dcn_v2_forward(x, sd[prefix + 'weight'], sd[prefix + 'bias'], offset + reord,
                          torch.sigmoid(m), dg=dg)


# ----------------------------------------------------------------------------- nets
def _conv(sd, name, x, pad=1):
    return F.conv2d(x, sd[name + '.weight'], sd[name + '.bias'], 1, pad)


def _resblocks(sd, prefix, x, n):
    """arch_util.py:80-136 ResidualBlockNoBN x n (res_scale 1)."""
    for i in range(n):
        y = _conv(sd, f'{prefix}.{i}.conv2', F.relu(_conv(sd, f'{prefix}.{i}.conv1', x)))
        x = x + y
    return x


_VGG16_TO_CONV3_1 = ['conv1_1', 'relu', 'conv1_2', 'relu', 'pool', 'conv2_1', 'relu', 'conv2_2', 'relu', 'pool', 'conv3_1']
_VGG19_TO_RELU3_1 = _VGG16_TO_CONV3_1 + ['relu']


def _vgg_trunk(sd, owner, seq, x, layers, taps=()):
    """`owner` holds the mean/std buffers, `owner.seq` the named conv stack."""
    x = (x - sd[owner + '.mean']) / sd[owner + '.std']
    got = {}
    last_conv = None
    for name in layers:
        if name.startswith('conv'):
            x = _conv(sd, f'{owner}.{seq}.{name}', x)
            last_conv = name
        elif name == 'relu':
            x = F.relu(x)
            rname = last_conv.replace('conv', 'relu')
            if rname in taps:
                got[rname] = x
        else:
            x = F.max_pool2d(x, 2, 2)
    return x, got


def contras_extractor(sd, img1, img2):
    """contras_extractor_arch.py:8-59 — two VGG16[:conv3_1] trunks (no ReLU after conv3_1)."""
    f1, _ = _vgg_trunk(sd, 'feature_extraction_image1', 'model', img1, _VGG16_TO_CONV3_1)
    f2, _ = _vgg_trunk(sd, 'feature_extraction_image2', 'model', img2, _VGG16_TO_CONV3_1)
    return f1, f2


def vgg19_ref_features(sd, img_ref):
    """vgg_arch.py:59-145 with layer_name_list relu1_1/relu2_1/relu3_1 (corres…:116)."""
    _, got = _vgg_trunk(sd, 'vgg', 'vgg_net', img_ref, _VGG19_TO_RELU3_1, taps=('relu1_1', 'relu2_1', 'relu3_1'))
    return got


def restoration_net(sd, x, pre_offset, ref_feat, n_blocks=16, groups=8):
    """ref_restoration_arch.py:51-65 + 147-187."""
    lrelu = lambda t: F.leaky_relu(t, 0.1)
    base = F.interpolate(x, None, 4, 'bilinear', False)
    f = lrelu(_conv(sd, 'content_extractor.conv_first', x))
    f = _resblocks(sd, 'content_extractor.body', f, n_blocks)
    p = 'dyn_agg_restore.'
    for size, key in (('small', 'relu3_1'), ('medium', 'relu2_1'), ('large', 'relu1_1')):
        off = torch.cat([f, ref_feat[key]], 1)
        off = lrelu(_conv(sd, f'{p}{size}_offset_conv1', off))
        off = lrelu(_conv(sd, f'{p}{size}_offset_conv2', off))
        swapped = lrelu(dcn_sep_pre_multi_offset(sd, f'{p}{size}_dyn_agg.', ref_feat[key], off,
                                                 pre_offset[key], groups))
        h = lrelu(_conv(sd, f'{p}head_{size}.0', torch.cat([f, swapped], 1)))
        h = _resblocks(sd, f'{p}body_{size}', h, n_blocks) + f
        if size != 'large':
            f = lrelu(F.pixel_shuffle(_conv(sd, f'{p}tail_{size}.0', h), 2))
        else:
            f = _conv(sd, f'{p}tail_large.2', lrelu(_conv(sd, f'{p}tail_large.0', h)))
    return f + base


def full_forward(sd_extractor, sd_map, sd_g, img_in_lq, img_in_up, img_ref, return_idx=False):
    """ref_restoration_model.py:271-279 `test()`: extractor -> net_map -> net_g, fp32, no grad."""
    with torch.no_grad():
        f1, f2 = contras_extractor(sd_extractor, img_in_up, img_ref)
        pre, idx = correspondence(f1, f2, return_idx=True)
        ref_feat = vgg19_ref_features(sd_map, img_ref)
        sr = restoration_net(sd_g, img_in_lq, pre, ref_feat)
    return (sr, idx) if return_idx else sr
