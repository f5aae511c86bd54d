"""Checker helpers shared by tests/test_gpu_fullsize.py and bench.py's `parity` block (cpu_baseline leg):
compare one pass of the GPU pipeline with the CPU oracle on the same pair.  Test infrastructure — imports
oracle/, never imported by the product."""
import torch
import torch.nn.functional as F

from oracle import ref_path


def exact_score64(fin, fref, q, r, gw, rw):
    """fp64 score of query patch q against Ref patch r on fp32 feature maps [C,h,w], with the reference's
    arithmetic (Ref patch normalised in fp32 first, ref_map_util.py:63; input patch norm :78-84)."""
    qy, qx = divmod(int(q), gw)
    ry, rx = divmod(int(r), rw)
    pq = fin[:, qy:qy + 3, qx:qx + 3]
    pr = fref[:, ry:ry + 3, rx:rx + 3]
    prn = (pr / (pr.norm() + 1e-5)).double()
    return float((pq.double() * prn).sum() / (pq.double().norm() + 1e-5))


def flips_with_gaps(idx, want_idx, fin, fref):
    """[(query, ours, reference, |fp64 score difference|)] for every query whose index differs."""
    gw = idx.shape[-1]
    rw = fref.shape[-1] - 2
    bad = (idx != want_idx).flatten().nonzero().flatten().tolist()
    out = []
    for q in bad:
        a, b = int(idx.flatten()[q]), int(want_idx.flatten()[q])
        out.append((q, a, b, abs(exact_score64(fin, fref, q, a, gw, rw) - exact_score64(fin, fref, q, b, gw, rw))))
    return out



def full_forward_parity(pipe, sds, img_lq, img_up, img_ref, want=None):
    """Shared by this test and bench.py's `parity` block: run one pair through the GPU pipeline and the CPU
    oracle, return the parity figures.  `want` = (sr, idx) of the oracle if already computed."""
    from mmsr.utils import metrics
    from mmsr.utils.util import tensor2img
    sd_e, sd_m, sd_g = sds
    dev = pipe.device
    sr, idx = pipe.forward(img_lq.to(dev), img_up.to(dev), img_ref.to(dev), return_idx=True)
    sr, idx = sr.cpu(), idx.cpu()
    if want is None:
        want = ref_path.full_forward(sd_e, sd_m, sd_g, img_lq, img_up, img_ref, return_idx=True)
    want_sr, want_idx = want
    with torch.no_grad():
        f1, f2 = ref_path.contras_extractor(sd_e, img_up, img_ref)
    gaps = []
    for b in range(idx.shape[0]):
        c, h, w = f1[b].shape
        fi = F.normalize(f1[b].reshape(c, -1), dim=0).view(c, h, w)
        fr = F.normalize(f2[b].reshape(c, -1), dim=0).view(c, h, w)
        gaps += [g for *_, g in flips_with_gaps(idx[b], want_idx[b], fi, fr)]
    # downstream exactness: the oracle restoration evaluated on THIS run's index map
    pre = {k: torch.stack([ref_path.offset_pyramid(ref_path.index_to_flow(idx[b]))[k] for b in range(idx.shape[0])])
           for k in ('relu3_1', 'relu2_1', 'relu1_1')}
    with torch.no_grad():
        sr_given_idx = ref_path.restoration_net(sd_g, img_lq, pre, ref_path.vgg19_ref_features(sd_m, img_ref))
    scale = float(sr_given_idx.abs().max())
    psnr = lambda a: metrics.psnr(tensor2img(a), tensor2img(img_up[0]), crop_border=4)
    return {'queries': int(idx.numel()), 'idx_flips': len(gaps), 'max_gap64_of_flips': max(gaps) if gaps else 0.0,
            'sr_max_rel_err': float((sr - sr_given_idx).abs().max()) / scale,
            'sr_max_rel_err_vs_oracle_own_idx': float((sr - want_sr).abs().max()) / float(want_sr.abs().max()),
            'psnr_delta_db': abs(psnr(sr[0]) - psnr(want_sr[0]))}
