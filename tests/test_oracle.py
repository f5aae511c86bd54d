"""Pin the oracle (oracle/ref_path.py torch port + oracle/c2m_oracle.c literal C) against the
golden fixtures minted from the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import make_golden as mg
import seeding
from oracle import c_oracle, ref_path


@pytest.fixture(scope='module', autouse=True)
def _build_c_oracle():
    c_oracle.build()


@pytest.mark.parametrize('case', mg.CORR_CASES, ids=[c[0] for c in mg.CORR_CASES])
def test_corr_torch_port_matches_reference(case, golden):
    g = golden['corr']
    name = case[0]
    fin, fref = mg.corr_inputs(case)
    assert seeding.sha(fin) == str(g[name + '/sha_in'])       # inputs rebuilt bit-identically
    assert seeding.sha(fref) == str(g[name + '/sha_ref'])
    for ni in (True, False):
        idx, val = ref_path.feature_match_index(fin, fref, 3, 1, 1, True, ni)
        assert np.array_equal(idx.numpy(), g[f'{name}/ni{int(ni)}/idx'])
        np.testing.assert_allclose(val.numpy(), g[f'{name}/ni{int(ni)}/val'], rtol=1e-6, atol=1e-7)


SMALL = [c for c in mg.CORR_CASES if c[0] not in ('cfg3_c256',)]


@pytest.mark.parametrize('case', SMALL, ids=[c[0] for c in SMALL])
def test_corr_c_oracle_matches_reference(case, golden):
    g = golden['corr']
    name = case[0]
    fin, fref = mg.corr_inputs(case)
    idx, val, gap = c_oracle.corr_argmax(fin, fref, is_norm=True, norm_input=True, want_gap=True)
    ref_idx = g[f'{name}/ni1/idx']
    assert np.array_equal(idx.numpy(), ref_idx), f'min fp64 gap {gap.min()}'
    assert np.array_equal(idx.numpy(), g[name + '/idx64'])
    np.testing.assert_allclose(val.numpy(), g[f'{name}/ni1/val'], rtol=2e-6, atol=1e-6)


def test_corr_is_norm_false(golden):
    g = golden['corr']
    fin, fref = mg.corr_inputs(mg.CORR_CASES[1])
    idx, val = c_oracle.corr_argmax(fin, fref, is_norm=False, norm_input=False)
    assert np.array_equal(idx.numpy(), g['odd_c32/raw/idx'])
    np.testing.assert_allclose(val.numpy(), g['odd_c32/raw/val'], rtol=2e-6, atol=1e-6)


def test_dup_ref_ties_take_lowest_index(golden):
    case = [c for c in mg.CORR_CASES if c[0] == 'dup_ref'][0]
    fin, fref = mg.corr_inputs(case)
    idx, _ = c_oracle.corr_argmax(fin, fref, is_norm=True)
    # Ref is a 5x5 block tiled 2x2 -> patch (y,x) with y,x<3 repeats at (y+5,x+5) etc.; the
    # winner must always be the first occurrence, i.e. inside rows/cols 0..4 of the 8x8 grid.
    iy, ix = idx // 8, idx % 8
    assert int(iy.max()) <= 4 and int(ix.max()) <= 4


@pytest.mark.parametrize('tag,shape', [('a', (32, 12, 12)), ('b', (16, 9, 14))])
def test_offset_pyramid(tag, shape, golden):
    g = golden['offsets']
    c, h, w = shape
    f1 = seeding.randn(300 + ord(tag), (2, c, h, w))
    f2 = seeding.randn(400 + ord(tag), (2, c, h, w))
    pre, idx = ref_path.correspondence(f1, f2, return_idx=True)
    for k, s in (('relu3_1', 1), ('relu2_1', 2), ('relu1_1', 4)):
        assert np.array_equal(pre[k].numpy().astype(np.int16), g[f'{tag}/{k}']), k
        for b in range(2):
            lit = c_oracle.offset_pyramid(idx[b], s)
            assert np.array_equal(lit.numpy().astype(np.int16), g[f'{tag}/{k}'][b]), (k, b)


def _dcn_module_sd(case):
    name, b, c, cout, h, w, dg, seed, osc = case
    spec = {'weight': (cout, c, 3, 3), 'bias': (cout,), 'conv_offset_mask.weight': (dg * 27, c, 3, 3),
            'conv_offset_mask.bias': (dg * 27,)}
    sd = seeding.seeded_state_dict(spec, seed + 3)
    sd['conv_offset_mask.weight'] = sd['conv_offset_mask.weight'] * osc
    return sd


@pytest.mark.parametrize('case', mg.DCN_CASES, ids=[c[0] for c in mg.DCN_CASES])
def test_dcn_module_port(case, golden):
    name, b, c, cout, h, w, dg, seed, osc = case
    sd = _dcn_module_sd(case)
    x, feat, pre = mg.dcn_inputs(case)
    y = ref_path.dcn_sep_pre_multi_offset(sd, '', x, feat, pre, dg)
    np.testing.assert_allclose(y.numpy(), golden['dcn'][name + '/out'], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('case', mg.DCN_CASES, ids=[c[0] for c in mg.DCN_CASES])
def test_dcn_literal_c_restatement_equals_standin(case):
    """The literal restatement of the reference .cu == torchvision's deform_conv2d (the stand-in
    the goldens were minted with), incl. integer, fractional and out-of-image offsets."""
    name, b, c, cout, h, w, dg, seed, osc = case
    x = seeding.randn(seed, (b, c, h, w))
    wgt = seeding.randn(seed + 5, (cout, c, 3, 3), 0.1)
    bias = seeding.randn(seed + 6, (cout,))
    off = seeding.randn(seed + 7, (b, 2 * dg * 9, h, w), 2.0 * osc)
    off[:, :, ::2, ::3] = off[:, :, ::2, ::3].round()          # exact-integer sampling points too
    mask = torch.sigmoid(seeding.randn(seed + 8, (b, dg * 9, h, w)))
    lit = c_oracle.dcn_v2_forward(x, wgt, bias, off, mask, dg=dg, acc64=True)
    tv = ref_path.dcn_v2_forward(x.double(), wgt.double(), bias.double(), off.double(), mask.double(), dg=dg)
    np.testing.assert_allclose(lit.numpy(), tv.float().numpy(), rtol=2e-5, atol=2e-5)
    lit32 = c_oracle.dcn_v2_forward(x, wgt, bias, off, mask, dg=dg, acc64=False)
    np.testing.assert_allclose(lit32.numpy(), lit.numpy(), rtol=1e-4, atol=1e-4)


def test_dcn_strided_dilated_literal():
    x = seeding.randn(1, (1, 8, 11, 9))
    wgt = seeding.randn(2, (6, 8, 3, 3), 0.2)
    bias = seeding.randn(3, (6,))
    ho, wo = (11 + 2 * 2 - 5) // 2 + 1, (9 + 2 * 2 - 5) // 2 + 1
    off = seeding.randn(4, (1, 2 * 2 * 9, ho, wo), 1.5)
    mask = torch.sigmoid(seeding.randn(5, (1, 2 * 9, ho, wo)))
    lit = c_oracle.dcn_v2_forward(x, wgt, bias, off, mask, 3, 3, 2, 2, 2, 2, 2, 2, dg=2)
    tv = ref_path.dcn_v2_forward(x, wgt, bias, off, mask, 3, 3, 2, 2, 2, 2, 2, 2, dg=2)
    np.testing.assert_allclose(lit.numpy(), tv.numpy(), rtol=2e-5, atol=2e-5)


def test_channel_l2norm():
    x = seeding.randn(9, (32, 6, 7))
    x[:, 0, 0] = 0
    ref = F.normalize(x.reshape(32, -1), dim=0).view_as(x)
    np.testing.assert_allclose(c_oracle.channel_l2norm(x).numpy(), ref.numpy(), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize('tag', ['cfg1', 'b2'])
def test_full_forward_port(tag, golden):
    """BASELINE config 1 (and a B=2 variant): the torch port reproduces the reference's SR image
    and index maps."""
    sd_e = seeding.share_extractor_weights(seeding.seeded_state_dict(seeding.spec_extractor(), 11))
    sd_m = seeding.seeded_state_dict(seeding.spec_net_map(), 12)
    sd_g = seeding.seeded_state_dict(seeding.spec_restoration_net(), 13)
    hr, img_lq, img_up, img_ref = seeding.full_case_inputs(tag)
    torch.set_num_threads(8)
    sr, idx = ref_path.full_forward(sd_e, sd_m, sd_g, img_lq, img_up, img_ref, return_idx=True)
    g = golden['full']
    assert np.array_equal(idx.numpy(), g[tag + '/max_idx'])
    np.testing.assert_allclose(sr.numpy(), g[tag + '/sr'], rtol=1e-4, atol=1e-4)
    assert g[tag + '/gap64'].shape == idx.shape


@pytest.mark.parametrize('cfg', [(2, 8, 6, 9, 7, 2, 3, 1, 1, 1), (1, 12, 10, 8, 10, 4, 3, 2, 2, 2), (1, 6, 4, 7, 6, 1, 5, 1, 2, 1)])
def test_dcn_backward_oracle_equals_torch_autograd(cfg):
    """Literal restatement of the reference backward == autograd through torchvision's CPU
    deform_conv2d (fp64), incl. offsets that leave the image, stride/dilation and a 5x5 kernel.
    Offsets are kept away from integer positions, where the bilinear derivative is one-sided."""
    B, C, cout, H, W, dg, ks, st, pad, dil = cfg
    T = ks * ks
    ho, wo = (H + 2 * pad - (dil * (ks - 1) + 1)) // st + 1, (W + 2 * pad - (dil * (ks - 1) + 1)) // st + 1
    x = seeding.randn(1, (B, C, H, W))
    wgt = seeding.randn(2, (cout, C, ks, ks), 0.2)
    bias = seeding.randn(3, (cout,))
    off = seeding.randn(4, (B, 2 * dg * T, ho, wo), 2.5)
    off = torch.where((off - off.round()).abs() < 0.05, off + 0.11, off)
    mask = torch.sigmoid(seeding.randn(5, (B, dg * T, ho, wo)))
    gout = seeding.randn(6, (B, cout, ho, wo))
    leaves = [t.double().requires_grad_(True) for t in (x, wgt, bias, off, mask)]
    y = ref_path.dcn_v2_forward(leaves[0], leaves[1], leaves[2], leaves[3], leaves[4], ks, ks, st, st, pad, pad, dil, dil, dg)
    y.backward(gout.double())
    gx, goff, gm, gw, gb = c_oracle.dcn_v2_backward(x, wgt, bias, off, mask, gout, ks, ks, st, st, pad, pad, dil, dil, dg)
    for got, leaf, name in ((gx, leaves[0], 'input'), (gw, leaves[1], 'weight'), (gb, leaves[2], 'bias'),
                            (goff, leaves[3], 'offset'), (gm, leaves[4], 'mask')):
        want = leaf.grad.float()
        assert float((got - want).abs().max()) <= 2e-5 * max(1.0, float(want.abs().max())), name
