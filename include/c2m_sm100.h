/*
 * libc2m_sm100 — C ABI of the B200-native C2-Matching restoration-forward hot path.
 *
 * Plain C: raw device pointers, sizes, a CUDA stream handle, int status.  No torch / ATen
 * types.  Every entry point is asynchronous on `stream`, re-entrant (no global scratch: the
 * caller owns outputs and the workspace) and takes the device from the current CUDA context,
 * which is what the reference's operator boundary assumes (it launches on the current stream
 * of the tensors' device under the GIL, DCNv2/src/cuda/dcn_v2_cuda.cu:107,139).
 *
 * Reference interfaces replaced (paths relative to the reference repo root):
 *   c2m_dcn_v2_forward_f32          <- `_ext.dcn_v2_forward`
 *                                      mmsr/models/archs/DCNv2/src/vision.cpp:3-8 (pybind export),
 *                                      src/dcn_v2.h:9-39 (dispatch), src/cuda/dcn_v2_cuda.cu:42-172
 *   c2m_dcn_v2_fused_forward_f32    <- DCN_sep_pre_multi_offset.forward (everything after the
 *                                      conv_offset_mask conv) mmsr/models/archs/DCNv2/dcn_v2.py:229-253
 *   c2m_corr_argmax_f32             <- feature_match_index  mmsr/models/archs/ref_map_util.py:26-86
 *                                      (+ optional fused per-pixel channel L2 normalisation,
 *                                      mmsr/models/archs/corres_generation_arch.py:56-58)
 *   c2m_offset_pyramid_f32          <- index_to_flow + 9-shift x scale pyramid
 *                                      mmsr/models/archs/corres_generation_arch.py:29-46,70-104
 *
 * The Python binding a reference maintainer would add is shown in INTEGRATION.md.
 */
#ifndef C2M_SM100_H
#define C2M_SM100_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* same object as cudaStream_t / CUstream */
typedef struct CUstream_st *c2m_stream_t;

/* status codes (0 = ok).  c2m_last_error() returns a thread-local description. */
enum {
    C2M_OK = 0,
    C2M_ERR_INVALID = 1,      /* bad shape / argument (reference: AT_ASSERTM -> RuntimeError) */
    C2M_ERR_WORKSPACE = 2,    /* workspace missing or too small */
    C2M_ERR_CUDA = 3,         /* CUDA runtime / launch failure (reference only printf'd these) */
    C2M_ERR_UNSUPPORTED = 4   /* no sm_100 device / driver entry point missing */
};

int c2m_abi_version(void);      /* 4 since c2m_dcn_tc_args.x_il / c2m_psa_interleave (3: .mask; 2: out_f32_octets / om_octets) */
const char *c2m_last_error(void);

/* --- correlation / index_map --------------------------------------------------------------
 * For b in [0,B):  idx[b,q] = argmax_r < P_in(b,q), P_ref(b,r) / (||P_ref(b,r)|| + 1e-5) >
 * over the Ref patch grid (lowest r on ties), val[b,q] = that maximum (divided by
 * ||P_in(b,q)|| + 1e-5 when norm_input).  fin [B,C,h,w], fref [B,C,hr,wr] fp32 NCHW;
 * idx int64 [B,h',w'], val fp32 [B,h',w'], h' = (h-patch)/s_in+1.  B = 1 is the reference call.
 *
 * l2norm != 0 first applies x[:,p] / max(||x[:,p]||, 1e-12) to both maps (the caller's
 * F.normalize in the reference).
 *
 * flags: bit0 = force the generic CUDA-core search instead of the tcgen05 kernel.
 * The tcgen05 search proposes candidates; the returned idx/val always come from the exact
 * (double-accumulated, float-rounded) rescoring pass, so they do not depend on tensor-core
 * rounding.
 */
size_t c2m_corr_workspace_bytes(int B, int C, int h, int w, int hr, int wr, int patch, int s_in, int s_ref);

int c2m_corr_argmax_f32(const float *fin, const float *fref, int B, int C, int h, int w, int hr, int wr,
                        int patch, int s_in, int s_ref, int is_norm, int norm_input, int l2norm,
                        unsigned flags, int64_t *idx, float *val, void *ws, size_t ws_bytes,
                        c2m_stream_t stream);

/* --- idx -> pre-offset pyramid ------------------------------------------------------------
 * idx [B,gh,gw] -> out [B,9,s*(gh+2),s*(gw+2),2] fp32, last dim (x,y);  tap k=3i+j is the flow
 * map scaled by s, nearest-upsampled by s and shifted down/right by (s*i, s*j), zero filled.
 * ref_gw: patch-grid width used to decode idx (the reference uses gw, the INPUT grid width).
 */
int c2m_offset_pyramid_f32(const int64_t *idx, int B, int gh, int gw, int ref_gw, int scale, float *out,
                           c2m_stream_t stream);

/* --- modulated deformable convolution v2, forward -----------------------------------------
 * out[b,o,p] = bias[o] + sum_{c,i,j} W[o,c,i,j] * mask[b,g(c)*kh*kw+i*kw+j,p] *
 *              bilinear(x[b,c], ho*sh-ph+i*dh+off_y, wo*sw-pw+j*dw+off_x)
 * off_y/off_x = offset[b, 2*(g*kh*kw+i*kw+j) (+1), p];  sample is 0 unless
 * -1 < y < H and -1 < x < W; g(c) = c / (C/dg).   All tensors fp32.
 * x  : [B,C,H,W]; element (b,c,y,x) at x[b*xs_b + c*xs_c + y*xs_y + x*xs_x]  (element strides,
 *      so NCHW-contiguous and channels-last storage are both accepted)
 * out: [B,Cout,Ho,Wo]; element strides os_b, os_c, os_y, os_x likewise.
 * offset [B,2*dg*kh*kw,Ho,Wo], mask [B,dg*kh*kw,Ho,Wo], weight [Cout,C,kh,kw], bias [Cout] or NULL:
 * contiguous.
 */
typedef struct {
    int B, C, H, W, Cout;
    int kh, kw, sh, sw, ph, pw, dh, dw, dg;
    long long xs_b, xs_c, xs_y, xs_x;
    long long os_b, os_c, os_y, os_x;
} c2m_dcn_shape;

int c2m_dcn_v2_forward_f32(const float *x, const float *offset, const float *mask, const float *weight,
                           const float *bias, float *out, const c2m_dcn_shape *shape, c2m_stream_t stream);

/* Fused DCN_sep_pre_multi_offset tail: `om` = conv_offset_mask(feat) [B,3*dg*kh*kw,Ho,Wo] raw;
 *   off_y = om[2j] + pre[...,1], off_x = om[2j+1] + pre[...,0], mask = sigmoid(om[2*dg*kh*kw + j]),
 *   j = g*kh*kw + k;  pre = pre_offset [B,kh*kw,Ho,Wo,2] (x,y) or NULL.
 * If pre == NULL and idx != NULL the pre-offsets are rebuilt on the fly from idx [B,gh,gw]
 * (pyramid scale `pre_scale`, decode width ref_gw) exactly as c2m_offset_pyramid_f32 would.
 * lrelu_slope: out = v > 0 ? v : slope*v applied when lrelu_slope != 1 (pass 1.f for none).
 */
int c2m_dcn_v2_fused_forward_f32(const float *x, const float *om, const float *pre, const int64_t *idx,
                                 int gh, int gw, int ref_gw, int pre_scale, const float *weight,
                                 const float *bias, float lrelu_slope, float *out,
                                 const c2m_dcn_shape *shape, c2m_stream_t stream);

/* --- plain 3x3 / stride 1 / pad 1 convolution on tcgen05, fp32-grade (split fp16 x 3) -------
 * (every plain convolution of RestorationNet / the VGG trunks; the reference
 * runs them through cuDNN, e.g. arch_util.py:80-136, ref_restoration_arch.py:147-187.)
 * Activations travel between these convolutions in the packed-split layout "PSA":
 *     hi, lo : fp16 [B][ceil(C/8)][H][W][8],   value = (hi + lo) * 2^-sa
 * c2m_psa_from_f32 / c2m_psa_to_f32 convert from / to strided fp32 ([B,C,H,W] with element
 * strides; to_f32 can add a strided fp32 tensor of the same strides).
 * Weights [Cout,Cin,3,3] fp32 are packed once (c2m_conv3x3_pack_weights_f32) into a blob of
 * c2m_conv3x3_packed_weight_bytes(Cin,Cout) bytes (Cin = total input channels).
 *     y = act(conv(cat[in, in2], W) + bias)          act: 0 none, 1 ReLU, 2 LeakyReLU(0.1)
 *     PSA output  (optional): y + res + res2, or PixelShuffle(2)(y) when pixel_shuffle == 2
 *     fp32 output (optional): y + add_f32, strided [B,Cout,H,W]; or, with out_f32_octets != 0, y in the
 *     octet-planar fp32 layout [B][ceil(Cout/8)][H][W][8] (strides and add_f32 unused; padding channels 0)
 *     that c2m_dcn_v2_fused_tc reads its `om` operand from (om_octets)
 * Any Cin / Cout; in2 (optional) is concatenated after in along channels (then Cin % 32 == 0).
 * Needs H >= 18 and W >= 10.
 */
typedef struct {
    const void *in_hi, *in_lo;   int Cin;
    const void *in2_hi, *in2_lo; int Cin2;
    int B, H, W, sa_in;
    const void *packed_w; const float *bias; int Cout; int act;
    const void *res_hi, *res_lo, *res2_hi, *res2_lo; int sa_res;
    void *out_hi, *out_lo; int sa_out; int pixel_shuffle;
    float *out_f32; const float *add_f32; long long os_b, os_c, os_y, os_x;
    int out_f32_octets;
} c2m_conv3x3_args;

size_t c2m_conv3x3_packed_weight_bytes(int Cin, int Cout);
int c2m_conv3x3_pack_weights_f32(const float *w, int Cin, int Cout, void *packed, c2m_stream_t stream);
int c2m_psa_from_f32(const float *x, int B, int C, int H, int W, long long xs_b, long long xs_c, long long xs_y,
                     long long xs_x, int sa, void *hi, void *lo, c2m_stream_t stream);
int c2m_psa_to_f32(const void *hi, const void *lo, int B, int C, int H, int W, int sa, const float *add, float *out,
                   long long os_b, long long os_c, long long os_y, long long os_x, c2m_stream_t stream);
int c2m_conv3x3(const c2m_conv3x3_args *args, c2m_stream_t stream);
/* hi / lo planes -> interleaved [B][ceil(C/8)][H][W][hi 8 | lo 8] (the DCN gather's preferred operand). */
int c2m_psa_interleave(const void *hi, const void *lo, int B, int C, int H, int W, void *out, c2m_stream_t stream);
/* MaxPool2d(2, 2) on a PSA tensor (floor semantics: odd trailing row / column dropped), PSA result. */
int c2m_psa_maxpool2(const void *hi, const void *lo, int B, int C, int H, int W, void *out_hi, void *out_lo,
                     c2m_stream_t stream);

/* --- DCNv2 forward on tcgen05 (3x3 / stride 1 / pad 1 / dilation 1, C/dg % 8 == 0, Cout <= 256) ---
 * Same contract as c2m_dcn_v2_fused_forward_f32 (raw conv_offset_mask output `om`, pre-offsets from
 * `pre` or rebuilt from `idx`), but the contraction runs as a split-fp16 tensor-core GEMM fed by
 * gather warps; the input is a PSA tensor (x_hi / x_lo, see c2m_psa_from_f32): its octet-planar layout
 * lets a warp's corner fetches share cache lines.  Outputs: PSA (out_hi/out_lo) and / or
 * strided fp32; `lrelu` != 0 applies LeakyReLU(0.1).  Weights are packed once with
 * c2m_dcn_tc_pack_weights_f32 into c2m_dcn_tc_packed_weight_bytes(C, Cout, dg) bytes.
 */
typedef struct {
    const void *x_hi, *x_lo;
    const float *om; const float *pre; const int64_t *idx;
    int gh, gw, ref_gw, pre_scale;
    int B, C, H, W, Cout, dg;
    const void *packed_w; const float *bias; int lrelu;
    void *out_hi, *out_lo; int sa_out;
    float *out_f32; long long os_b, os_c, os_y, os_x;
    int om_octets;               /* != 0: om is octet-planar fp32 [B][ceil(27*dg/8)][H][W][8] */
    const float *mask;           /* != NULL: the `_ext.dcn_v2_forward` contract (DCNv2/src/dcn_v2.h:9-22) — `om` is the
                                  * FINAL offset tensor [B,2*dg*9,H,W], `mask` the FINAL modulation [B,dg*9,H,W]
                                  * (no sigmoid applied); pre, idx and om_octets must be unset */
    const void *x_il;            /* != NULL: the input with the two halves of a (pixel, octet) adjacent,
                                  * fp16 [B][C/8][H][W][hi 8 | lo 8] (c2m_psa_interleave), 32 B aligned; x_hi / x_lo may
                                  * then be NULL.  A corner fetch is one 32 B sector instead of two. */
} c2m_dcn_tc_args;

int c2m_dcn_tc_supported(int C, int Cout, int dg);
size_t c2m_dcn_tc_packed_weight_bytes(int C, int Cout, int dg);
int c2m_dcn_tc_pack_weights_f32(const float *w, int C, int Cout, int dg, void *packed, c2m_stream_t stream);
int c2m_dcn_v2_fused_tc(const c2m_dcn_tc_args *args, c2m_stream_t stream);

/* --- DCNv2 backward building blocks (`_ext.dcn_v2_backward`, DCNv2/src/dcn_v2.h:41-72) -----------
 * All tensors fp32, NCHW-contiguous; shape->x* / o* strides are ignored here.
 *   columns [B, C*kh*kw, Ho*Wo] = mask * bilinear(x)                      (for grad_weight = gout x columns^T)
 *   gcol    [B, C*kh*kw, Ho*Wo] = W^T x grad_output                       (computed by the caller)
 *   grad_offset [B,2*dg*kh*kw,Ho,Wo], grad_mask [B,dg*kh*kw,Ho,Wo] written; grad_input [B,C,H,W] ACCUMULATED
 *   (caller zero-fills it first).
 */
int c2m_dcn_v2_im2col_f32(const float *x, const float *offset, const float *mask, const c2m_dcn_shape *shape,
                          float *columns, c2m_stream_t stream);
int c2m_dcn_v2_col2im_coord_f32(const float *gcol, const float *x, const float *offset, const float *mask,
                                const c2m_dcn_shape *shape, float *grad_offset, float *grad_mask,
                                c2m_stream_t stream);
int c2m_dcn_v2_col2im_f32(const float *gcol, const float *offset, const float *mask, const c2m_dcn_shape *shape,
                          float *grad_input, c2m_stream_t stream);

/* number of kernels this library has launched since load (bench.py's gpu_launches) */
unsigned long long c2m_launch_count(void);

/* Measurement hook (bench.py `roofline`): while enabled, the library brackets every launch of its
 * three heavy kernel classes with CUDA events on the caller's stream:
 *   0 = correlation candidate search (corr_umma / generic), 1 = conv3x3_umma, 2 = dcn_umma.
 * c2m_profile_collect synchronises the recorded events of one class, returns their summed device
 * time, launch count and summed ALGORITHMIC flops / bytes (SURVEY.md §8d formulas), and clears them. */
int c2m_profile_enable(int on);
int c2m_profile_collect(int kernel, float *ms_total, int *launches, double *flops, double *bytes);

#ifdef __cplusplus
}
#endif
#endif /* C2M_SM100_H */
