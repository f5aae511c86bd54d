// Modulated deformable convolution v2 forward on the sm_100a tensor cores (3x3, stride 1, pad 1,
// dilation 1 — every DCN of C2-Matching), fused with DCN_sep_pre_multi_offset's prologue.
//
// Replaces `_ext.dcn_v2_forward` (mmsr/models/archs/DCNv2/src/cuda/dcn_v2_cuda.cu:42-172: im2col
// into a 236-944 MB `columns` buffer + batched fp32 SGEMM) and the Python prologue
// (DCNv2/dcn_v2.py:233-250).  Sampling arithmetic follows dcn_v2_im2col_cuda.cu:25-54,125-195
// operation for operation; the contraction is a split-fp16 (hi*hi + hi*lo + lo*hi) tcgen05 GEMM
// with fp32 TMEM accumulation, i.e. fp32-grade.
//
// Structure = conv3x3_umma.cu with the TMA activation producer replaced by GATHER warps:
//   K is ordered (deformable group g, tap, channel-in-group); a K chunk = 4 channel octets.
//   8 gather warps build, per (pixel tile, chunk), the column tile [4 octets][128 px][8] directly in
//   the tcgen05 K-major no-swizzle layout (hi and lo halves): per (pixel, g, tap) the sampling point
//   is computed once (raw conv_offset_mask output + pre-offset rebuilt from the index map or read
//   from the pre_offset tensor, sigmoid mask, validity of the 4 corners), the 8 channels of an octet
//   are fetched as 16 B (hi) + 16 B (lo) per corner from the octet-planar PSA input, blended in fp32,
//   multiplied by the mask, split and stored with one 16 B st.shared per half;
//   fence.proxy.async + mbarrier hand the stage to the MMA-issuing thread.
//   Weights stream as 8 KB chunks (bulk copy, 4-deep ring), shared by a group of up to 8 pixel
//   tiles (512/N TMEM accumulators, rotated across groups); every CTA owns an evenly sized contiguous
//   range of the flat (slice, image, tile) list, so small problems (BASELINE config 4: 200 tiles)
//   spread over all SMs.  The epilogue (bias, LeakyReLU, PSA and / or fp32 store) is shared with
//   the plain convolution.  -DC2M_DCN_TRACE prints a %globaltimer trace of CTA 0's pipeline.
// The im2col matrix never exists outside shared memory.
#include <cstdlib>

#include "umma_conv_common.cuh"

namespace c2m {

namespace {
constexpr int T_R = 16, T_C = 8;
constexpr int KOCT = 4;
constexpr int A_OCT_B = 128 * 16;          // 2048 B: one octet of all 128 pixels (LBO of A)
constexpr int A_HALF = KOCT * A_OCT_B;     // 8192 B
constexpr int A_STAGE = 2 * A_HALF;        // 16384 B
constexpr int NSTAGE = 4;
constexpr int NBST = 4;
constexpr int MAXT = 8;
#ifndef C2M_DCN_NU
#define C2M_DCN_NU 1
#endif
constexpr int NU = C2M_DCN_NU;                       // K octets per gather thread and stage
constexpr int NGATHER_WARPS = 4 * KOCT / NU;         // thread = (pixel of the tile, NU octets of the chunk)
constexpr int NTHREADS = 256 + 32 * NGATHER_WARPS;
constexpr int W_HDR = 256;

struct DcnTc {
    const __half *x_hi, *x_lo; // input in the packed-split layout [B][C/8][H][W][8] (value = hi + lo)
    int C8;
    const float *om;           // [B, 3*dg*9, H, W], or octet-planar [B][om_c8][H][W][8] when om_c8 > 0
    int om_c8;
    const float *mask;         // non-null: FINAL modulation mask [B, dg*9, H, W] (no sigmoid) and `om` holds only the
                               // FINAL offsets [B, 2*dg*9, H, W] — the `_ext.dcn_v2_forward` contract
    const float *pre;          // [B, 9, H, W, 2] or null
    const long long *idx;      // [B, gh, gw] or null
    int gh, gw, ref_gw, pre_scale;
    int C, dg, cpg, opp;       // opp = octets per (g, tap) pair = cpg / 8
    int n_ko;                  // real K octets = C/8 * 9
    int opp_shift;             // log2(opp): octets per (g,tap) pair is 1, 2 or 4 for C/dg in {8,16,32}; -1 otherwise
    float inv_ref_gw, inv_scale;   // reciprocals for the exact float-assisted integer divisions
};
}  // namespace

// exact n / d for 0 <= n < 2^23 with a precomputed float reciprocal (one multiply + fix-up instead
// of the ~25-instruction integer division sequence; the gather loop is issue-bound)
__device__ __forceinline__ int fast_div(int n, int d, float inv_d) {
    int q = (int)((float)n * inv_d);
    const int r = n - q * d;
    q += (r >= d) - (r < 0);
    return q;
}

#ifdef C2M_DCN_TRACE
__device__ unsigned long long g_tr[32][4], g_tm[32][3], g_t0[2];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define TR(x) x
#else
#define TR(x)
#endif

__global__ void __launch_bounds__(NTHREADS, 1)
dcn_umma_kernel(const ConvPtrs q, const ConvParams p, const DcnTc d) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t w_half = (uint32_t)KOCT * p.N * 16;
    const uint32_t w_chunk = 2 * w_half;
    uint8_t *sW = smem;                                   // [NBST][hi | lo]
    uint8_t *sA = smem + NBST * w_chunk;                  // N multiple of 16 -> multiple of 128
    uint64_t *bars = reinterpret_cast<uint64_t *>(sA + NSTAGE * A_STAGE);
    uint64_t *full = bars, *empty = full + NSTAGE, *bfull = empty + NSTAGE, *bempty = bfull + NBST,
             *tfull = bempty + NBST, *tempty = tfull + MAXT;
    uint32_t *tmem_base_p = reinterpret_cast<uint32_t *>(tempty + MAXT);
    float *sbias = reinterpret_cast<float *>(tmem_base_p + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TR(if (blockIdx.x == 0 && threadIdx.x == 0) g_t0[0] = gtime();)
    const int tiles_img = p.tiles_x * p.tiles_y;
    // Work = flat list of (Cout slice, image, pixel tile); every CTA owns one contiguous, evenly sized range
    // of it and walks the range in groups of <= T tiles that share the streamed weight chunks.
    const int TT = p.B * tiles_img;
    const int n_work = p.nslice * TT;
    const int w_begin = (int)((long long)n_work * blockIdx.x / gridDim.x);
    const int w_end = (int)((long long)n_work * (blockIdx.x + 1) / gridDim.x);
    const int nacc = p.T;                                // TMEM accumulators, rotated across groups
    const uint32_t need_cols = (uint32_t)p.N * p.T;
    const uint32_t tmem_cols = need_cols <= 32 ? 32 : need_cols <= 64 ? 64 : need_cols <= 128 ? 128
                               : need_cols <= 256 ? 256 : 512;

    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(&full[i], NGATHER_WARPS); mbar_init(&empty[i], 1); }
        for (int i = 0; i < NBST; ++i) { mbar_init(&bfull[i], 1); mbar_init(&bempty[i], 1); }
        for (int i = 0; i < MAXT; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_base_p, tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_p;
    TR(if (blockIdx.x == 0 && threadIdx.x == 0) g_t0[1] = gtime();)
    TR(int trs = 0;)
    const int sw = *reinterpret_cast<const int *>(q.wblob);

    auto group = [&](int pos, int &slice, int &nt) {
        slice = pos / TT;
        nt = min(min(p.T, w_end - pos), (slice + 1) * TT - pos);
    };

    // Register budget per warpgroup.  The CTA's pool is what it was launched with (80 x 768 = 61440):
    // the control warpgroup gives 40 x 128 back and the epilogue (32 TMEM columns in flight) takes
    // them, 40 + 120 + 4 x 80 = 480 = 6 x 80.  Requests beyond the pool would block forever.
    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
        // ================================ weight producer ===================================
        if (lane == 0) {
            int bst = 0, bphase = 0;
            for (int pos = w_begin, nt = 0; pos < w_end; pos += nt) {
                int slice;
                group(pos, slice, nt);
                for (int kc = 0; kc < p.nkc; ++kc) {
                    mbar_wait(&bempty[bst], bphase ^ 1);
                    mbar_arrive_expect_tx(&bfull[bst], w_chunk);
                    const uint8_t *src = q.wblob + W_HDR + ((size_t)slice * p.nkc + kc) * w_chunk;
                    asm volatile(
                        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                        ::"r"(smem_u32(sW + bst * w_chunk)), "l"(src), "r"(w_chunk), "r"(smem_u32(&bfull[bst]))
                        : "memory");
                    if (++bst == NBST) { bst = 0; bphase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ========================================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f16(128, p.N, 0);
            const uint32_t b_lbo = p.N * 16;
            int stage = 0, phase = 0, bst = 0, bphase = 0;
            uint32_t tph = 0;
            int abase = 0;
            for (int pos = w_begin, nt = 0; pos < w_end; pos += nt) {
                int slice;
                group(pos, slice, nt);
                for (int kc = 0; kc < p.nkc; ++kc) {
                    mbar_wait(&bfull[bst], bphase);
                    tc_fence_after();
                    TR(if (blockIdx.x == 0 && trs < 32) g_tm[trs][0] = gtime();)
                    const uint32_t w_hi = smem_u32(sW + bst * w_chunk), w_lo = w_hi + w_half;
                    for (int t = 0; t < nt; ++t) {
                        int a = abase + t;
                        if (a >= nacc) a -= nacc;
                        if (kc == 0) {
                            mbar_wait(&tempty[a], ((tph >> a) & 1u) ^ 1u);
                            tc_fence_after();
                        }
                        const uint32_t dacc = tmem_base + a * p.N;
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        TR(if (blockIdx.x == 0 && trs < 32) g_tm[trs][1] = gtime();)
                        const uint32_t a_hi = smem_u32(sA + stage * A_STAGE), a_lo = a_hi + A_HALF;
#pragma unroll
                        for (int j = 0; j < KOCT / 2; ++j) {
                            const uint32_t ao = j * 2 * A_OCT_B, bo = j * 2 * b_lbo;
                            const uint64_t dah = umma_smem_desc(a_hi + ao, A_OCT_B, 128);
                            const uint64_t dal = umma_smem_desc(a_lo + ao, A_OCT_B, 128);
                            const uint64_t dbh = umma_smem_desc(w_hi + bo, b_lbo, 128);
                            const uint64_t dbl = umma_smem_desc(w_lo + bo, b_lbo, 128);
                            umma_f16(dacc, dah, dbh, idesc, (kc | j) != 0);
                            umma_f16(dacc, dah, dbl, idesc, 1);
                            umma_f16(dacc, dal, dbh, idesc, 1);
                        }
                        umma_commit(&empty[stage]);
                        TR(if (blockIdx.x == 0 && trs < 32) { g_tm[trs][2] = gtime(); ++trs; })
                        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                        if (kc == p.nkc - 1) { umma_commit(&tfull[a]); tph ^= 1u << a; }
                    }
                    umma_commit(&bempty[bst]);
                    if (++bst == NBST) { bst = 0; bphase ^= 1; }
                }
                abase += nt;
                if (abase >= nacc) abase -= nacc;
            }
        }
    }
    } else if (warp < 8) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 120;");
        // ================================ epilogue ==========================================
        const int e = threadIdx.x - 128;
        const int quarter = warp & 3;
        const float out_scale = ldexpf(1.f, -(p.sa_in + sw));
        const float res_scale = 1.f;
        const float so = ldexpf(1.f, p.sa_out);
        uint32_t tph = 0;
        int abase = 0, cur_slice = -1;
        for (int pos = w_begin, nt = 0; pos < w_end; pos += nt) {
            int slice;
            group(pos, slice, nt);
            const int o_base = slice * p.N;
            if (slice != cur_slice) {                      // uniform over the 4 epilogue warps
                cur_slice = slice;
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (int i = e; i < p.N; i += 128) sbias[i] = (q.bias && o_base + i < p.Cout) ? q.bias[o_base + i] : 0.f;
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            for (int t = 0; t < nt; ++t) {
                const int r = pos + t - slice * TT;
                const int b = r / tiles_img, tt = r - b * tiles_img;
                const int y = (tt / p.tiles_x) * T_R + e / T_C, x = (tt % p.tiles_x) * T_C + e % T_C;
                const bool ok = y < p.H && x < p.W;
                int a = abase + t;
                if (a >= nacc) a -= nacc;
                mbar_wait(&tfull[a], (tph >> a) & 1u);
                tph ^= 1u << a;
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + a * p.N;
                for (int c0 = 0; c0 < p.N; c0 += 32)
                    epilogue_store_block(q, p, taddr, c0, b, y, x, ok, o_base, sbias, out_scale, res_scale, so);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[a]);
            }
            abase += nt;
            if (abase >= nacc) abase -= nacc;
        }
    } else {
        // ================================ gather producers ==================================
        // Thread = (pixel m of the tile, octet pair oh of the chunk).  Per stage two dependent
        // memory round trips exist (offsets/mask/idx -> corner addresses -> corner values); the
        // first one is taken off the critical path by fetching the NEXT stage's metadata while the
        // current stage's corners are in flight.
        const int g_tid = threadIdx.x - 256;
        const int m = g_tid & 127;
        const int oh = g_tid >> 7;
        const int P = p.H * p.W;
        struct Meta { float off_h, off_w, mr; int v; };
        int stage = 0, phase = 0;
        const float inv_ti = 1.f / (float)tiles_img;
        const size_t x_img = (size_t)d.C8 * P * 8, om_img = (size_t)(d.mask ? 2 : 3) * d.dg * 9 * P;
        for (int pos = w_begin, nt = 0; pos < w_end; pos += nt) {
            int slice;
            group(pos, slice, nt);
            const int r0g = pos - slice * TT;               // (image, tile) index of the group's first tile
            const int n_steps = p.nkc * nt;

            // (kc, t) and the pixel of this thread in tile t are advanced incrementally; all divisions by
            // run-time values go through fast_div / shifts
            const int mrow = m / T_C, mcol = m % T_C;
            const float inv_tx = 1.f / (float)p.tiles_x;
            const int opp_shift = d.opp_shift;
            const int om_mask_base = 2 * d.dg * 9;

            // metadata fetch for (kc, t, u): raw offsets + mask logit + index-map entry
            auto fetch = [&](int kc, int t, bool in_range, int u, Meta &mt, int &pair_out, bool &live) {
                const int b = fast_div(r0g + t, tiles_img, inv_ti);
                const int tt = r0g + t - b * tiles_img;
                const float *omb = d.om + b * om_img;
                const int ty = fast_div(tt, p.tiles_x, inv_tx);
                const int y = ty * T_R + mrow, xx = (tt - ty * p.tiles_x) * T_C + mcol;
                const int ko = kc * KOCT + oh * NU + u;
                live = in_range && y < p.H && xx < p.W && ko < d.n_ko;
                pair_out = opp_shift >= 0 ? (ko >> opp_shift) : ko / d.opp;
                mt.off_h = mt.off_w = mt.mr = 0.f;
                mt.v = -1;
                if (!live) return;
                const int g = pair_out / 9, tap = pair_out - g * 9;
                const int jj = g * 9 + tap, pp = y * p.W + xx;
                if (d.om_c8 > 0) {     // 2jj is even: the (y, x) offset pair shares an octet
                    const float *ob = d.om + (size_t)b * d.om_c8 * P * 8;
                    const int c0 = 2 * jj, cm = om_mask_base + jj;
                    const float2 of = *reinterpret_cast<const float2 *>(ob + ((c0 >> 3) * P + pp) * 8 + (c0 & 7));
                    mt.off_h = of.x;
                    mt.off_w = of.y;
                    mt.mr = ob[((cm >> 3) * P + pp) * 8 + (cm & 7)];
                } else {
                    mt.off_h = omb[(2 * jj) * P + pp];
                    mt.off_w = omb[(2 * jj + 1) * P + pp];
                    mt.mr = d.mask ? d.mask[((size_t)b * d.dg * 9 + jj) * P + pp] : omb[(om_mask_base + jj) * P + pp];
                }
                if (d.pre) {
                    const float2 pq = *reinterpret_cast<const float2 *>(d.pre + (((size_t)b * 9 + tap) * P + pp) * 2);
                    mt.off_w += pq.x;
                    mt.off_h += pq.y;
                } else if (d.idx) {
                    const long long *idxb = d.idx + (long long)b * d.gh * d.gw;
                    const int sc = d.pre_scale, ki = tap / 3, kj = tap - ki * 3;
                    const int ys = y - sc * ki, xs = xx - sc * kj;
                    if (ys >= 0 && xs >= 0) {
                        const int yy = fast_div(ys, sc, d.inv_scale), xg = fast_div(xs, sc, d.inv_scale);
                        if (yy < d.gh && xg < d.gw) mt.v = (int)idxb[yy * d.gw + xg];
                    }
                }
            };

            Meta mt[NU], nx[NU];
            int pr[NU], npr[NU];
            bool lv[NU], nlv[NU];
#pragma unroll
            for (int u = 0; u < NU; ++u) fetch(0, 0, n_steps > 0, u, mt[u], pr[u], lv[u]);
            int kc = 0, t = 0;
            for (int step = 0; step < n_steps; ++step) {
                const int b = fast_div(r0g + t, tiles_img, inv_ti);
                const int tt = r0g + t - b * tiles_img;
                const __half *xh = d.x_hi + b * x_img;
                const __half *xl = d.x_lo + b * x_img;
                TR(const bool trc = blockIdx.x == 0 && threadIdx.x == 256 && trs < 32; if (trc) g_tr[trs][0] = gtime();)
                const int ty = fast_div(tt, p.tiles_x, inv_tx);
                const int y = ty * T_R + mrow, xx = (tt - ty * p.tiles_x) * T_C + mcol;
                // next (kc, t)
                int nkc_ = kc, nt_ = t + 1;
                if (nt_ == nt) { nt_ = 0; ++nkc_; }
                // ---- sampling points of the two octets
                int o[NU][4];
                float wq[NU][4], mk[NU];
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    o[u][0] = o[u][1] = o[u][2] = o[u][3] = 0;
                    wq[u][0] = wq[u][1] = wq[u][2] = wq[u][3] = 0.f;
                    mk[u] = 0.f;
                    if (lv[u]) {
                        const int g = pr[u] / 9, tap = pr[u] - g * 9, ki = tap / 3, kj = tap - ki * 3;
                        float off_h = mt[u].off_h, off_w = mt[u].off_w;
                        if (mt[u].v >= 0) {
                            const int sc = d.pre_scale;
                            const int yy = fast_div(y - sc * ki, sc, d.inv_scale), xg = fast_div(xx - sc * kj, sc, d.inv_scale);
                            const int vy = fast_div(mt[u].v, d.ref_gw, d.inv_ref_gw), vx = mt[u].v - vy * d.ref_gw;
                            off_w += (float)(sc * (vx - xg));
                            off_h += (float)(sc * (vy - yy));
                        }
                        const float h_im = (float)(y - 1 + ki) + off_h;
                        const float w_im = (float)(xx - 1 + kj) + off_w;
                        if (h_im > -1.f && w_im > -1.f && h_im < (float)p.H && w_im < (float)p.W) {
                            const int h_low = (int)floorf(h_im), w_low = (int)floorf(w_im);
                            const int h_high = h_low + 1, w_high = w_low + 1;
                            const float lh = h_im - h_low, lw = w_im - w_low;
                            const float hh = 1.f - lh, hw = 1.f - lw;
                            const bool tv = h_low >= 0, bv = h_high <= p.H - 1, lvv = w_low >= 0, rv = w_high <= p.W - 1;
                            // channel octet of this K octet, as an element offset of its [H][W][8] plane
                            const int ko = kc * KOCT + oh * NU + u;
                            const int oct_c = g * (d.cpg / 8) + (ko - (opp_shift >= 0 ? (pr[u] << opp_shift) : pr[u] * d.opp));
                            const int cbase = oct_c * P * 8;
                            const int r0 = (h_low * p.W + w_low) * 8 + cbase;
                            if (tv && lvv) { o[u][0] = r0; wq[u][0] = hh * hw; }
                            if (tv && rv) { o[u][1] = r0 + 8; wq[u][1] = hh * lw; }
                            if (bv && lvv) { o[u][2] = r0 + p.W * 8; wq[u][2] = lh * hw; }
                            if (bv && rv) { o[u][3] = r0 + p.W * 8 + 8; wq[u][3] = lh * lw; }
                            // sigmoid to ~2 ulp: the TC path is fp32-grade, not bit-exact
                            mk[u] = d.mask ? mt[u].mr : __fdividef(1.f, 1.f + __expf(-mt[u].mr));
                        }
                    }
                }
                // octet-planar operand: the 32 lanes of a warp (4 rows x 8 pixels) read 16 B each from runs of
                // adjacent pixels (8 lines per request instead of 32 with a channels-last fp32 input)
                uint4 ch[NU][4], cl[NU][4];
#pragma unroll
                for (int u = 0; u < NU; ++u)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        ch[u][c] = *reinterpret_cast<const uint4 *>(xh + o[u][c]);
                        cl[u][c] = *reinterpret_cast<const uint4 *>(xl + o[u][c]);
                    }
#pragma unroll
                for (int u = 0; u < NU; ++u) fetch(nkc_, nt_, step + 1 < n_steps, u, nx[u], npr[u], nlv[u]);
                // ---- blend, modulate, split.  value = hi + lo: the hi halves are blended in fp32, the lo halves
                // (|lo| <= 2^-11 |value|) in packed fp16 — their rounding lands at 2^-22 of the value; the
                // sigmoid mask is folded into the four corner weights
                uint4 h_out[NU], l_out[NU];
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    float wm[4];
                    __half2 wh[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        wm[c] = wq[u][c] * mk[u];
                        wh[c] = __float2half2_rn(wm[c]);
                    }
                    const __half2 *hp[4], *lp[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        hp[c] = reinterpret_cast<const __half2 *>(&ch[u][c]);
                        lp[c] = reinterpret_cast<const __half2 *>(&cl[u][c]);
                    }
                    __align__(16) __half2 h4[4];
                    __align__(16) __half2 l4[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        __half2 ls = __hmul2(wh[0], lp[0][j]);
                        ls = __hfma2(wh[1], lp[1][j], ls);
                        ls = __hfma2(wh[2], lp[2][j], ls);
                        ls = __hfma2(wh[3], lp[3][j], ls);
                        const float2 lf = __half22float2(ls);
                        const float2 a = __half22float2(hp[0][j]), bq = __half22float2(hp[1][j]);
                        const float2 cq = __half22float2(hp[2][j]), dq = __half22float2(hp[3][j]);
                        const float vx = fmaf(wm[0], a.x, fmaf(wm[1], bq.x, fmaf(wm[2], cq.x, fmaf(wm[3], dq.x, lf.x))));
                        const float vy = fmaf(wm[0], a.y, fmaf(wm[1], bq.y, fmaf(wm[2], cq.y, fmaf(wm[3], dq.y, lf.y))));
                        const __half2 hq = __floats2half2_rn(vx, vy);
                        const float2 hf = __half22float2(hq);
                        h4[j] = hq;
                        l4[j] = __floats2half2_rn(vx - hf.x, vy - hf.y);
                    }
                    h_out[u] = *reinterpret_cast<const uint4 *>(h4);
                    l_out[u] = *reinterpret_cast<const uint4 *>(l4);
                }
                TR(if (trc) g_tr[trs][1] = gtime();)
                mbar_wait(&empty[stage], phase ^ 1);
                TR(if (trc) g_tr[trs][2] = gtime();)
                uint8_t *sdst = sA + stage * A_STAGE;
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    const int oct = oh * NU + u;
                    *reinterpret_cast<uint4 *>(sdst + oct * A_OCT_B + m * 16) = h_out[u];
                    *reinterpret_cast<uint4 *>(sdst + A_HALF + oct * A_OCT_B + m * 16) = l_out[u];
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[stage]);
                TR(if (trc) { g_tr[trs][3] = gtime(); } ++trs;)
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
#pragma unroll
                for (int u = 0; u < NU; ++u) { mt[u] = nx[u]; pr[u] = npr[u]; lv[u] = nlv[u]; }
                kc = nkc_;
                t = nt_;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
    TR(if (blockIdx.x == 0 && threadIdx.x == 0) {
        const unsigned long long z = g_t0[0];
        printf("TRACE setup %llu end %llu\n", g_t0[1] - z, gtime() - z);
        for (int i = 0; i < 24; ++i)
            printf("  s%02d gather top %6llu loads %6llu wait %6llu arr %6llu | mma bfull %6llu full %6llu commit %6llu\n", i,
                   g_tr[i][0] - z, g_tr[i][1] - z, g_tr[i][2] - z, g_tr[i][3] - z, g_tm[i][0] - z, g_tm[i][1] - z, g_tm[i][2] - z);
    })
}

// blob: header | [slice][kc][hi|lo][octet 4][N][8], K ordered (g, tap, channel-in-group)
__global__ void dcn_wpack_kernel(const float *__restrict__ w, int C, int Cout, int dg, int N, int nkc, int nslice,
                                 uint8_t *__restrict__ blob) {
    __shared__ int s_sw;
    if (threadIdx.x == 0) {
        const float a = __uint_as_float(reinterpret_cast<unsigned *>(blob)[1]);
        int e = 0;
        if (a > 0.f && isfinite(a)) frexpf(a, &e);
        s_sw = 14 - e;
        if (blockIdx.x == 0) reinterpret_cast<int *>(blob)[0] = s_sw;
    }
    __syncthreads();
    const float S = ldexpf(1.f, s_sw);
    const int cpg = C / dg, opp = cpg / 8, n_ko = (C / 8) * 9;
    const long long total = (long long)nslice * nkc * 2 * KOCT * N * 8;
    __half *dst = reinterpret_cast<__half *>(blob + W_HDR);
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(e & 7);
        long long rest = e >> 3;
        const int ol = (int)(rest % N); rest /= N;
        const int oct = (int)(rest % KOCT); rest /= KOCT;
        const int half = (int)(rest % 2); rest /= 2;
        const int kc = (int)(rest % nkc);
        const int slice = (int)(rest / nkc);
        const int ko = kc * KOCT + oct, o = slice * N + ol;
        float v = 0.f;
        if (ko < n_ko && o < Cout) {
            const int pair = ko / opp, oc = ko % opp, g = pair / 9, tap = pair % 9;
            const int c = g * cpg + oc * 8 + j;
            v = w[((size_t)o * C + c) * 9 + tap] * S;
        }
        const __half hh = __float2half_rn(v);
        dst[e] = half ? __float2half_rn(v - __half2float(hh)) : hh;
    }
}

__global__ void dcn_wamax_kernel(const float *__restrict__ w, int n, unsigned *__restrict__ bits) {
    float m = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = fmaxf(m, fabsf(w[i]));
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(bits, __float_as_uint(m));
}

static inline int d_pad16(int c) { return (c + 15) / 16 * 16; }
// the gather is the expensive side, so one CTA owns ALL output channels (N up to 256) instead of
// re-gathering per 64-wide slice; tiles per item shrink so that T*N <= 512 TMEM columns
static inline int d_slice_n(int cout) { return d_pad16(cout); }
static inline int d_n_slice(int cout) { (void)cout; return 1; }
static inline int d_nkc(int C) { return ((C / 8) * 9 + KOCT - 1) / KOCT; }

}  // namespace c2m

using namespace c2m;

extern "C" int c2m_dcn_tc_supported(int C, int Cout, int dg) {
    return (C > 0 && Cout > 0 && Cout <= 256 && dg > 0 && C % dg == 0 && (C / dg) % 8 == 0) ? 1 : 0;
}

extern "C" size_t c2m_dcn_tc_packed_weight_bytes(int C, int Cout, int dg) {
    if (!c2m_dcn_tc_supported(C, Cout, dg)) return 0;
    return (size_t)W_HDR + (size_t)d_n_slice(Cout) * d_nkc(C) * 2 * KOCT * d_slice_n(Cout) * 16;
}

extern "C" int c2m_dcn_tc_pack_weights_f32(const float *w, int C, int Cout, int dg, void *packed, c2m_stream_t stream) {
    C2M_CHECK_ARG(w && packed, "dcn_tc_pack_weights: null pointer");
    C2M_CHECK_ARG(c2m_dcn_tc_supported(C, Cout, dg), "dcn_tc_pack_weights: C=%d dg=%d unsupported (C/dg must be a multiple of 8)", C, dg);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    C2M_CUDA(cudaMemsetAsync(packed, 0, W_HDR, st));
    const int n = Cout * C * 9;
    dcn_wamax_kernel<<<ceil_div(n, 1024) > 64 ? 64 : ceil_div(n, 1024), 256, 0, st>>>(w, n, reinterpret_cast<unsigned *>(packed) + 1);
    C2M_LAUNCH_CHECK("dcn_wamax_kernel");
    const int N = d_slice_n(Cout), nkc = d_nkc(C), ns = d_n_slice(Cout);
    const long long total = (long long)ns * nkc * 2 * KOCT * N * 8;
    const int blocks = (int)((total + 255) / 256 > 592 ? 592 : (total + 255) / 256);
    dcn_wpack_kernel<<<blocks, 256, 0, st>>>(w, C, Cout, dg, N, nkc, ns, reinterpret_cast<uint8_t *>(packed));
    C2M_LAUNCH_CHECK("dcn_wpack_kernel");
    return C2M_OK;
}

extern "C" int c2m_dcn_v2_fused_tc(const c2m_dcn_tc_args *a, c2m_stream_t stream) {
    C2M_CHECK_ARG(a && a->x_hi && a->x_lo && a->om && a->packed_w, "dcn_v2_fused_tc: null pointer");
    C2M_CHECK_ARG(a->B > 0 && a->H > 0 && a->W > 0, "dcn_v2_fused_tc: bad shape");
    C2M_CHECK_ARG(c2m_dcn_tc_supported(a->C, a->Cout, a->dg), "dcn_v2_fused_tc: C=%d dg=%d unsupported", a->C, a->dg);
    C2M_CHECK_ARG(!(a->pre == nullptr && a->idx != nullptr) || (a->gh > 0 && a->gw > 0 && a->ref_gw > 0 && a->pre_scale > 0),
                  "dcn_v2_fused_tc: idx given without a valid grid/scale");
    C2M_CHECK_ARG((a->out_hi == nullptr) == (a->out_lo == nullptr) && (a->out_hi || a->out_f32), "dcn_v2_fused_tc: bad outputs");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    ConvParams p = {};
    p.B = a->B; p.H = a->H; p.W = a->W;
    p.nkc = d_nkc(a->C); p.nkc_a = p.nkc;
    p.Cout = a->Cout; p.N = d_slice_n(a->Cout); p.nslice = d_n_slice(a->Cout);
    p.tiles_x = ceil_div(a->W, T_C); p.tiles_y = ceil_div(a->H, T_R);
    p.T = 512 / p.N < MAXT ? 512 / p.N : MAXT;
    p.n_st = ceil_div(p.tiles_x * p.tiles_y, p.T);
    p.act = a->lrelu ? 2 : 0;
    p.sa_in = 0; p.sa_res = 0; p.sa_out = a->sa_out;
    p.ps = 0;
    p.stacked = 0;
    p.dbg = 0;
    p.C8out = (a->Cout + 7) / 8; p.Hout = a->H; p.Wout = a->W;
    p.os_b = a->os_b; p.os_c = a->os_c; p.os_y = a->os_y; p.os_x = a->os_x;
    p.f32_mode = a->out_f32 ? f32_store_mode(a->out_f32, nullptr, a->os_b, a->os_c, a->os_y, a->os_x) : 0;
    ConvPtrs q = {};
    q.wblob = reinterpret_cast<const uint8_t *>(a->packed_w);
    q.bias = a->bias;
    q.out_hi = reinterpret_cast<__half *>(a->out_hi); q.out_lo = reinterpret_cast<__half *>(a->out_lo);
    q.out_f32 = a->out_f32;
    DcnTc d;
    d.x_hi = reinterpret_cast<const __half *>(a->x_hi); d.x_lo = reinterpret_cast<const __half *>(a->x_lo);
    d.C8 = a->C / 8;
    d.om_c8 = a->om_octets ? (27 * a->dg + 7) / 8 : 0;
    C2M_CHECK_ARG(!a->om_octets || (reinterpret_cast<uintptr_t>(a->om) & 7) == 0, "dcn_v2_fused_tc: om must be 8 B aligned");
    C2M_CHECK_ARG(!a->mask || (!a->om_octets && !a->pre && !a->idx),
                  "dcn_v2_fused_tc: a final mask excludes octet-planar om and pre-offsets");
    d.mask = a->mask;
    d.om = a->om; d.pre = a->pre; d.idx = reinterpret_cast<const long long *>(a->idx);
    d.gh = a->gh; d.gw = a->gw; d.ref_gw = a->ref_gw; d.pre_scale = a->pre_scale;
    d.C = a->C; d.dg = a->dg; d.cpg = a->C / a->dg; d.opp = d.cpg / 8; d.n_ko = (a->C / 8) * 9;
    d.opp_shift = d.opp == 1 ? 0 : d.opp == 2 ? 1 : d.opp == 4 ? 2 : d.opp == 8 ? 3 : -1;
    d.inv_ref_gw = a->ref_gw > 0 ? 1.f / (float)a->ref_gw : 0.f;
    d.inv_scale = a->pre_scale > 0 ? 1.f / (float)a->pre_scale : 1.f;
    C2M_CHECK_ARG((long long)a->H * a->W * (27 * a->dg + 7) < (1ll << 31) && (long long)a->H * a->W * a->C < (1ll << 31),
                  "dcn_v2_fused_tc: map too large for 32-bit in-kernel indexing");
    C2M_CHECK_ARG(a->idx == nullptr || (long long)a->gh * a->gw < (1 << 23), "dcn_v2_fused_tc: index map too large");
    const size_t smem = (size_t)NBST * 2 * KOCT * p.N * 16 + NSTAGE * A_STAGE + 4096;
    C2M_CUDA(cudaFuncSetAttribute(dcn_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 0;
    C2M_CUDA(cudaGetDevice(&dev));
    C2M_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long n_work = (long long)a->B * p.tiles_x * p.tiles_y * p.nslice;
    C2M_CHECK_ARG(n_work < (1 << 23), "dcn_v2_fused_tc: too many pixel tiles");
    // flops = 2*B*Cout*C*9*Ho*Wo; bytes = 4*B*(C*H*W + 3*dg*9*H*W + Cout*H*W) + weights
    const double px = (double)a->B * a->H * a->W;
    const double flops = 2.0 * a->Cout * a->C * 9.0 * px;
    const double bytes = 4.0 * px * (a->C + 27.0 * a->dg + a->Cout) + 4.0 * a->Cout * (a->C * 9.0 + 1.0);
    void *ph = prof_begin(PROF_DCN, flops, bytes, st);
    dcn_umma_kernel<<<n_work < sms ? (int)n_work : sms, NTHREADS, smem, st>>>(q, p, d);
    C2M_LAUNCH_CHECK("dcn_umma_kernel");
    prof_end(ph, st);
    return C2M_OK;
}
