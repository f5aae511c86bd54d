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
//   K is ordered (deformable group g, tap, channel-in-group); a stage = (128-pixel tile, 4 K octets).
//   16 gather warps build the column tile [4 octets][128 px][8] directly in the tcgen05 K-major
//   no-swizzle layout (hi and lo halves).  A gather thread owns (pixel, run of L K-octets that share one
//   (g, tap) pair), L = min(C/dg/8, 4): the sampling point — raw conv_offset_mask output + pre-offset, sigmoid
//   mask, the 4 corner offsets and mask-folded bilinear weights — is computed ONCE per (pixel, g, tap) and reused
//   for the run's octets; for L > 1 the 4 thread groups work on L consecutive stages concurrently.  Per octet:
//   16 B (hi) + 16 B (lo) per corner from the octet-planar PSA input (L2 evict-last: the input map is re-read by all
//   nine taps and must stay resident, while offsets / outputs stream through with evict-first), blended (hi halves
//   in fp32, lo halves in packed fp16), split and stored with one 16 B st.shared per half;
//   fence.proxy.async + mbarrier hand the stage to the MMA-issuing thread.
//   Pre-offsets: per tile group the index-map cells a tile can touch ((16/s + 2) x (8/s + 2) per tile) are
//   decoded once into a shared-memory flow table, so the per-(pixel, tap) pre-offset is one ld.shared instead
//   of an 8-byte global load, two integer divisions and a modulo per (pixel, g, tap).
//   Weights stream as 8 KB chunks (bulk copy ring), shared by a group of up to 8 pixel
//   tiles (512/N TMEM accumulators, rotated across groups); every CTA owns an evenly sized contiguous
//   range of the flat (slice, image, tile) list, so small problems (BASELINE config 4: 200 tiles)
//   spread over all SMs.  The epilogue (bias, LeakyReLU, PSA and / or fp32 store) is shared with
//   the plain convolution.
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
constexpr int MAX_NSTAGE = 8;
constexpr int MAX_NBST = 4;
constexpr int MAXT = 8;
constexpr int NGATHER_WARPS = 16;
constexpr int NTHREADS = 256 + 32 * NGATHER_WARPS;
constexpr int W_HDR = 256;

struct DcnTc {
    const __half *x_hi, *x_lo; // input in the packed-split layout [B][C/8][H][W][8] (value = hi + lo)
    const __half *x_il;        // or (non-null) the same values interleaved [B][C/8][H][W][hi 8 | lo 8]: one sector per corner
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
    int L, lshift;             // K octets per gather thread and stage (1, 2 or 4), log2
    int opp_shift;             // log2(opp) when opp is a power of two, else -1 (then L == 1)
    int nstage, nbst;          // ring depths (activation stages, weight chunks); powers of two
    int nstage_shift;          // log2(nstage)
    int tg;                    // pixel tiles per work item (<= T)
    int policy;                // L2 policies: 1 = input map evict-last, 2 = offsets / mask evict-first, 4 = streaming output stores
    int sc_shift;              // log2(pre_scale) when the flow table is used (idx given, scale in {1,2,4,8}), else -1
    int tab_h, tab_w;          // flow-table cells per tile: 16/s + 2, 8/s + 2
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

// ---- cache-policy helpers: the gathered input map is re-read by all nine taps (keep it in L2), the offset /
// mask stream and the outputs are touched once (do not let them evict it)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_normal() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint4 ldg_keep_v4(const void *ptr, uint64_t pol) {
    uint4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.b32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(ptr), "l"(pol));
    return v;
}
// one (pixel, octet) of the interleaved operand: 32 B = one sector, one 256-bit load
__device__ __forceinline__ void ldg_keep_v8(const void *ptr, uint64_t pol, uint4 &h, uint4 &l) {
    asm volatile("ld.global.nc.L2::cache_hint.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8], %9;"
                 : "=r"(h.x), "=r"(h.y), "=r"(h.z), "=r"(h.w), "=r"(l.x), "=r"(l.y), "=r"(l.z), "=r"(l.w)
                 : "l"(ptr), "l"(pol));
}
__device__ __forceinline__ float2 ldg_stream_f2(const float *ptr, uint64_t pol) {
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(ptr), "l"(pol));
    return v;
}
__device__ __forceinline__ float ldg_stream_f1(const float *ptr, uint64_t pol) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(ptr), "l"(pol));
    return v;
}

__global__ void __launch_bounds__(NTHREADS, 1)
dcn_umma_kernel(const ConvPtrs q, const ConvParams p, const DcnTc d) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t w_half = (uint32_t)KOCT * p.N * 16;
    const uint32_t w_chunk = 2 * w_half;
    uint8_t *sW = smem;                                   // [nbst][hi | lo]
    uint8_t *sA = smem + d.nbst * w_chunk;                // N multiple of 16 -> multiple of 128
    uint64_t *bars = reinterpret_cast<uint64_t *>(sA + d.nstage * A_STAGE);
    uint64_t *full = bars, *empty = full + MAX_NSTAGE, *bfull = empty + MAX_NSTAGE, *bempty = bfull + MAX_NBST,
             *tfull = bempty + MAX_NBST, *tempty = tfull + MAXT;
    uint32_t *tmem_base_p = reinterpret_cast<uint32_t *>(tempty + MAXT);
    float *sbias = reinterpret_cast<float *>(tmem_base_p + 4);                    // [N <= 256], 16 B aligned
    int4 *tile_info = reinterpret_cast<int4 *>(sbias + 256);                      // [2][MAXT] (b, y0, x0, image pixel base)
    float2 *flow_tab = reinterpret_cast<float2 *>(tile_info + 2 * MAXT);          // [2][T][tab_h * tab_w]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_img = p.tiles_x * p.tiles_y;
    // Work = flat list of (image, pixel tile), cut into items of Tg consecutive tiles (they share the streamed weight
    // chunks); item i runs on CTA i % grid.  At any moment the CTAs therefore work on ~grid * Tg CONSECUTIVE tiles —
    // a third of one image at the largest layer — so the gathered input map of ONE image (105 MB) is what has to
    // stay in L2, not the whole batch (contiguous per-CTA ranges spread the CTAs over all images at once: ncu showed
    // 7.3 GB of DRAM reads for 2.3 GB of algorithmic traffic).  Tg shrinks for small problems so all SMs get work.
    const int TT = p.B * tiles_img;
    const int n_work = TT;
    const int Tg = d.tg;
    const int n_items = (n_work + Tg - 1) / Tg;
    const int nacc = p.T;                                // TMEM accumulators, rotated across groups
    const uint32_t need_cols = (uint32_t)p.N * p.T;
    const uint32_t tmem_cols = need_cols <= 32 ? 32 : need_cols <= 64 ? 64 : need_cols <= 128 ? 128
                               : need_cols <= 256 ? 256 : 512;
    const int nstage_mask = d.nstage - 1;                // ring depths are powers of two

    if (warp == 1 && lane == 0) {
        for (int i = 0; i < d.nstage; ++i) { mbar_init(&full[i], NGATHER_WARPS >> d.lshift); mbar_init(&empty[i], 1); }
        for (int i = 0; i < d.nbst; ++i) { mbar_init(&bfull[i], 1); mbar_init(&bempty[i], 1); }
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
    const int sw = *reinterpret_cast<const int *>(q.wblob);

    auto group = [&](int item, int &pos, int &nt) {
        pos = item * Tg;
        nt = min(Tg, n_work - pos);
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
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                for (int kc = 0; kc < p.nkc; ++kc) {
                    mbar_wait(&bempty[bst], bphase ^ 1);
                    mbar_arrive_expect_tx(&bfull[bst], w_chunk);
                    const uint8_t *src = q.wblob + W_HDR + (size_t)kc * w_chunk;
                    asm volatile(
                        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                        ::"r"(smem_u32(sW + bst * w_chunk)), "l"(src), "r"(w_chunk), "r"(smem_u32(&bfull[bst]))
                        : "memory");
                    if (++bst == d.nbst) { bst = 0; bphase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ========================================
        {   // whole warp, uniform control flow; one elected lane issues (see umma_f16_w)
            const uint32_t idesc = umma_idesc_f16(128, p.N, 0);
            const uint32_t b_lbo = p.N * 16;
            int stage = 0, phase = 0, bst = 0, bphase = 0;
            uint32_t tph = 0;
            int abase = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                int pos, nt;
                group(item, pos, nt);
                for (int kc = 0; kc < p.nkc; ++kc) {
                    mbar_wait(&bfull[bst], bphase);
                    tc_fence_after();
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
                        const uint32_t a_hi = smem_u32(sA + stage * A_STAGE), a_lo = a_hi + A_HALF;
#pragma unroll
                        for (int j = 0; j < KOCT / 2; ++j) {
                            const uint32_t ao = j * 2 * A_OCT_B, bo = j * 2 * b_lbo;
                            const uint64_t dah = umma_smem_desc(a_hi + ao, A_OCT_B, 128);
                            const uint64_t dal = umma_smem_desc(a_lo + ao, A_OCT_B, 128);
                            const uint64_t dbh = umma_smem_desc(w_hi + bo, b_lbo, 128);
                            const uint64_t dbl = umma_smem_desc(w_lo + bo, b_lbo, 128);
                            umma_f16_w(dacc, dah, dbh, idesc, (kc | j) != 0);
                            umma_f16_w(dacc, dah, dbl, idesc, 1);
                            umma_f16_w(dacc, dal, dbh, idesc, 1);
                        }
                        umma_commit_w(&empty[stage]);
                        if (++stage == d.nstage) { stage = 0; phase ^= 1; }
                        if (kc == p.nkc - 1) { umma_commit_w(&tfull[a]); tph ^= 1u << a; }
                    }
                    umma_commit_w(&bempty[bst]);
                    if (++bst == d.nbst) { bst = 0; bphase ^= 1; }
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
        const bool fast = epilogue_fast_ok(q, p) && p.sa_out == 0;     // (this kernel does not pre-scale bias / factors)
        uint32_t tph = 0;
        int abase = 0;
        for (int i = e; i < p.N; i += 128) sbias[i] = (q.bias && i < p.Cout) ? q.bias[i] : 0.f;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int o_base = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            int pos, nt;
            group(item, pos, nt);
            for (int t = 0; t < nt; ++t) {
                const int r = pos + t;
                const int b = r / tiles_img, tt = r - b * tiles_img;
                const int y = (tt / p.tiles_x) * T_R + e / T_C, x = (tt % p.tiles_x) * T_C + e % T_C;
                const bool ok = y < p.H && x < p.W;
                int a = abase + t;
                if (a >= nacc) a -= nacc;
                mbar_wait(&tfull[a], (tph >> a) & 1u);
                tph ^= 1u << a;
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + a * p.N;
                for (int c0 = 0; c0 < p.N; c0 += 32) {
                    if (fast && c0 + 32 <= p.N && c0 + 32 <= p.Cout)
                        epilogue_fast_dispatch<false>(q, p, taddr, c0, b, y, x, ok, o_base, sbias, out_scale, res_scale, nullptr);
                    else
                        epilogue_store_block(q, p, taddr, c0, b, y, x, ok, o_base, sbias, out_scale, res_scale, so);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[a]);
            }
            abase += nt;
            if (abase >= nacc) abase -= nacc;
        }
    } else {
        // ================================ gather producers ==================================
        // 4 thread groups of 128 (thread = pixel m of the tile).  With L K-octets per thread a stage has NS = 4 / L
        // thread slots, so the 4 groups split into NG = L stage groups that work on consecutive stages concurrently:
        // group grp -> (stage residue sg = grp / NS, slot = grp % NS); slot covers octets [slot*L, slot*L + L).
        // All index arithmetic is 32-bit (the host checks the map sizes); everything that depends only on the
        // K chunk (group, tap, operand plane offsets) is recomputed when kc changes, not per pixel.
        const int g_tid = threadIdx.x - 256;
        const int m = g_tid & 127;
        const int grp = g_tid >> 7;
        const int L = d.L, NG = L, NS = KOCT >> d.lshift;
        const int sg = grp / NS, slot = grp - sg * NS;
        const int P = p.H * p.W, W8 = p.W * 8;
        const int mrow = m / T_C, mcol = m % T_C;
        const uint64_t pol_keep = (d.policy & 1) ? l2_policy_evict_last() : l2_policy_evict_normal();
        const uint64_t pol_stream = (d.policy & 2) ? l2_policy_evict_first() : l2_policy_evict_normal();
        const int x_img = d.C8 * P * 8;                       // elements per image of one operand half (< 2^31, host-checked)
        const int om_img = (d.om_c8 > 0 ? d.om_c8 * 8 : (d.mask ? 2 : 3) * d.dg * 9) * P;
        const int om_mask_base = 2 * d.dg * 9;
        const int tab_n = d.tab_h * d.tab_w;
        // cell of this thread's pixel inside a tile's flow table, before the tap shift (+2 halo cells)
        const int cell0 = d.sc_shift >= 0 ? ((mrow >> d.sc_shift) + 2) * d.tab_w + (mcol >> d.sc_shift) + 2 : 0;
        const float Hf = (float)p.H, Wf = (float)p.W;
        struct Meta { float off_h, off_w, mr; };
        int ring = 0;                                       // stages issued by this CTA so far (all groups agree)
        int gi = 0;                                         // tile-group counter (table double buffer)
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++gi) {
            int r0g, nt;                                    // (image, tile) index of the group's first tile, tiles
            group(item, r0g, nt);
            const int n_steps = p.nkc * nt;
            int4 *tinfo = tile_info + (gi & 1) * MAXT;
            float2 *tab = flow_tab + (gi & 1) * p.T * tab_n;

            // ---- per tile group: tile origins and the decoded flow of every index-map cell a tile can touch
            if (g_tid < nt) {
                const int r = r0g + g_tid;
                const int b = r / tiles_img, tt = r - b * tiles_img;
                const int ty = tt / p.tiles_x;
                tinfo[g_tid] = make_int4(b, ty * T_R, (tt - ty * p.tiles_x) * T_C, 0);
            }
            if (d.sc_shift >= 0) {
                for (int i = g_tid; i < nt * tab_n; i += 32 * NGATHER_WARPS) {
                    const int t = i / tab_n, c = i - t * tab_n;
                    const int r = r0g + t;
                    const int b = r / tiles_img, tt = r - b * tiles_img;
                    const int ty = tt / p.tiles_x, tx = tt - ty * p.tiles_x;
                    const int cy = c / d.tab_w, cx = c - cy * d.tab_w;
                    const int acy = ((ty * T_R) >> d.sc_shift) - 2 + cy, acx = ((tx * T_C) >> d.sc_shift) - 2 + cx;
                    float2 fl = make_float2(0.f, 0.f);
                    if (acy >= 0 && acx >= 0 && acy < d.gh && acx < d.gw) {
                        const int v = (int)d.idx[((long long)b * d.gh + acy) * d.gw + acx];
                        const int vy = fast_div(v, d.ref_gw, d.inv_ref_gw), vx = v - vy * d.ref_gw;
                        fl = make_float2((float)((vx - acx) << d.sc_shift), (float)((vy - acy) << d.sc_shift));   // (x, y)
                    }
                    tab[i] = fl;
                }
            }
            asm volatile("bar.sync 2, 512;" ::: "memory");

            // ---- constants of K chunk kc for this thread's octet run
            int c_kc = -1, c_ki = 0, c_kj = 0, c_tap = 0, c_tapcell = 0, c_off = 0, c_moff = 0, c_xoff = 0;
            bool c_valid = false;
            float c_fy = 0.f, c_fx = 0.f;
            auto set_kc = [&](int kc) {
                c_kc = kc;
                const int ko = kc * KOCT + slot * L;
                c_valid = ko < d.n_ko;
                const int pair = d.opp_shift >= 0 ? (ko >> d.opp_shift) : ko / d.opp;
                const int g = pair / 9;
                c_tap = pair - g * 9;
                c_ki = c_tap / 3;
                c_kj = c_tap - c_ki * 3;
                c_fy = (float)(c_ki - 1);
                c_fx = (float)(c_kj - 1);
                c_tapcell = c_ki * d.tab_w + c_kj;
                if (d.om_c8 > 0) {     // octet-planar: channel c of pixel pp lives at ((c >> 3) * P + pp) * 8 + (c & 7)
                    const int c0 = 2 * pair, cm = om_mask_base + pair;
                    c_off = (c0 >> 3) * P * 8 + (c0 & 7);
                    c_moff = (cm >> 3) * P * 8 + (cm & 7);
                } else {
                    c_off = 2 * pair * P;
                    c_moff = (d.mask ? pair : om_mask_base + pair) * P;
                }
                // first channel octet of this run: group g, octets [ko - pair * opp, +L) of the group
                c_xoff = (g * d.opp + (ko - (d.opp_shift >= 0 ? (pair << d.opp_shift) : pair * d.opp))) * P * 8;
            };

            // ---- metadata of unit (kc, t): raw offsets (+ pre-offset) and mask logit of this thread's pixel
            auto fetch = [&](int kc, int t, bool in_range, Meta &mt, int &y_out, int &x_out, int &b_out, bool &live) {
                mt.off_h = mt.off_w = mt.mr = 0.f;
                y_out = x_out = b_out = 0;
                live = false;
                if (!in_range) return;
                if (kc != c_kc) set_kc(kc);
                const int4 ti = tinfo[t];
                const int b = ti.x, y = ti.y + mrow, xx = ti.z + mcol;
                live = c_valid && y < p.H && xx < p.W;
                y_out = y; x_out = xx; b_out = b;
                if (!live) return;
                const int pp = y * p.W + xx;
                const float *omb = d.om + (size_t)b * om_img;
                if (d.om_c8 > 0) {     // the (y, x) offset pair shares an octet: one 8-byte load
                    const float2 of = ldg_stream_f2(omb + c_off + pp * 8, pol_stream);
                    mt.off_h = of.x;
                    mt.off_w = of.y;
                    mt.mr = ldg_stream_f1(omb + c_moff + pp * 8, pol_stream);
                } else {
                    mt.off_h = ldg_stream_f1(omb + c_off + pp, pol_stream);
                    mt.off_w = ldg_stream_f1(omb + c_off + P + pp, pol_stream);
                    mt.mr = d.mask ? ldg_stream_f1(d.mask + (size_t)b * (d.dg * 9 * P) + c_moff + pp, pol_stream)
                                   : ldg_stream_f1(omb + c_moff + pp, pol_stream);
                }
                if (d.pre) {
                    const float2 pq = *reinterpret_cast<const float2 *>(d.pre + ((size_t)(b * 9 + c_tap) * P + pp) * 2);
                    mt.off_w += pq.x;
                    mt.off_h += pq.y;
                } else if (d.sc_shift >= 0) {
                    const float2 fl = tab[t * tab_n + cell0 - c_tapcell];
                    mt.off_w += fl.x;
                    mt.off_h += fl.y;
                } else if (d.idx) {                       // scales the table does not cover: decode in place
                    const int sc = d.pre_scale;
                    const int ys = y - sc * c_ki, xs = xx - sc * c_kj;
                    if (ys >= 0 && xs >= 0) {
                        const int yy = fast_div(ys, sc, d.inv_scale), xg = fast_div(xs, sc, d.inv_scale);
                        if (yy < d.gh && xg < d.gw) {
                            const int v = (int)d.idx[((long long)b * d.gh + yy) * d.gw + xg];
                            const int vy = fast_div(v, d.ref_gw, d.inv_ref_gw), vx = v - vy * d.ref_gw;
                            mt.off_w += (float)(sc * (vx - xg));
                            mt.off_h += (float)(sc * (vy - yy));
                        }
                    }
                }
            };

            // my units of this group: steps sg, sg + NG, ...; (kc, t) advanced incrementally
            int kc = 0, t = sg;
            while (t >= nt) { t -= nt; ++kc; }
            Meta mt, nx;
            int y, ny, xx, nxx, b, nb;
            bool lv, nlv;
            fetch(kc, t, sg < n_steps, mt, y, xx, b, lv);
            for (int step = sg; step < n_steps; step += NG) {
                int nkc_ = kc, nt_ = t + NG;
                while (nt_ >= nt) { nt_ -= nt; ++nkc_; }
                // ---- sampling point of the unit's (pixel, g, tap): dcn_v2_im2col_cuda.cu:25-54,172-190
                int o0 = 0, o1 = 0, o2 = 0, o3 = 0;
                float w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f;
                const int xoff = c_xoff;                    // belongs to kc (set_kc(kc) ran in this unit's fetch)
                if (lv) {
                    const float h_im = ((float)y + c_fy) + mt.off_h;
                    const float w_im = ((float)xx + c_fx) + mt.off_w;
                    if (h_im > -1.f && w_im > -1.f && h_im < Hf && w_im < Wf) {
                        const float hf = floorf(h_im), wf = floorf(w_im);
                        const int h_low = (int)hf, w_low = (int)wf;
                        const float lh = h_im - hf, lw = w_im - wf;
                        const float hh = 1.f - lh, hw = 1.f - lw;
                        const bool tv = h_low >= 0, bv = h_low + 1 <= p.H - 1, lvv = w_low >= 0, rv = w_low + 1 <= p.W - 1;
                        // sigmoid to ~2 ulp: the TC path is fp32-grade, not bit-exact; the mask is folded into the weights
                        const float mk = d.mask ? mt.mr : __fdividef(1.f, 1.f + __expf(-mt.mr));
                        const int r0 = h_low * W8 + w_low * 8;
                        const float mh = hh * mk, ml = lh * mk;
                        if (tv && lvv) { o0 = r0; w0 = mh * hw; }
                        if (tv && rv) { o1 = r0 + 8; w1 = mh * lw; }
                        if (bv && lvv) { o2 = r0 + W8; w2 = ml * hw; }
                        if (bv && rv) { o3 = r0 + W8 + 8; w3 = ml * lw; }
                    }
                }
                const __half *xh = d.x_hi + (size_t)b * x_img + (lv ? xoff : 0);
                const __half *xl = d.x_lo + (size_t)b * x_img + (lv ? xoff : 0);
                const __half2 wh0 = __float2half2_rn(w0), wh1 = __float2half2_rn(w1), wh2 = __float2half2_rn(w2),
                              wh3 = __float2half2_rn(w3);
                const int stage_idx = ring + step;
                const int rs = stage_idx & nstage_mask;
                const uint32_t rphase = (uint32_t)(stage_idx >> d.nstage_shift) & 1u;
                uint8_t *sdst = sA + rs * A_STAGE + (slot * L) * A_OCT_B + m * 16;
                // next unit's metadata goes out before this unit's corner fetches come back
                fetch(nkc_, nt_, step + NG < n_steps, nx, ny, nxx, nb, nlv);
#pragma unroll 1
                for (int u = 0; u < L; ++u) {
                    // octet-planar operand: the 32 lanes of a warp (4 rows x 8 pixels) read 16 B each from runs of
                    // adjacent pixels (8 lines per request instead of 32 with a channels-last fp32 input)
                    const int po = u * P * 8;
                    uint4 ch0, ch1, ch2, ch3, cl0, cl1, cl2, cl3;
                    if (d.x_il) {                          // element offsets double in the interleaved operand
                        const __half *xi = d.x_il + 2 * ((size_t)b * x_img + (lv ? xoff : 0) + po);
                        ldg_keep_v8(xi + 2 * o0, pol_keep, ch0, cl0);
                        ldg_keep_v8(xi + 2 * o1, pol_keep, ch1, cl1);
                        ldg_keep_v8(xi + 2 * o2, pol_keep, ch2, cl2);
                        ldg_keep_v8(xi + 2 * o3, pol_keep, ch3, cl3);
                    } else {
                        ch0 = ldg_keep_v4(xh + po + o0, pol_keep); ch1 = ldg_keep_v4(xh + po + o1, pol_keep);
                        ch2 = ldg_keep_v4(xh + po + o2, pol_keep); ch3 = ldg_keep_v4(xh + po + o3, pol_keep);
                        cl0 = ldg_keep_v4(xl + po + o0, pol_keep); cl1 = ldg_keep_v4(xl + po + o1, pol_keep);
                        cl2 = ldg_keep_v4(xl + po + o2, pol_keep); cl3 = ldg_keep_v4(xl + po + o3, pol_keep);
                    }
                    // ---- blend, modulate, split.  value = hi + lo: the hi halves are blended in fp32, the lo halves
                    // (|lo| <= 2^-11 |value|) in packed fp16 — their rounding lands at 2^-22 of the value
                    const __half2 *hp0 = reinterpret_cast<const __half2 *>(&ch0), *hp1 = reinterpret_cast<const __half2 *>(&ch1);
                    const __half2 *hp2 = reinterpret_cast<const __half2 *>(&ch2), *hp3 = reinterpret_cast<const __half2 *>(&ch3);
                    const __half2 *lp0 = reinterpret_cast<const __half2 *>(&cl0), *lp1 = reinterpret_cast<const __half2 *>(&cl1);
                    const __half2 *lp2 = reinterpret_cast<const __half2 *>(&cl2), *lp3 = reinterpret_cast<const __half2 *>(&cl3);
                    __align__(16) __half2 h4[4];
                    __align__(16) __half2 l4[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        __half2 ls = __hmul2(wh0, lp0[j]);
                        ls = __hfma2(wh1, lp1[j], ls);
                        ls = __hfma2(wh2, lp2[j], ls);
                        ls = __hfma2(wh3, lp3[j], ls);
                        const float2 lf = __half22float2(ls);
                        const float2 a = __half22float2(hp0[j]), bq = __half22float2(hp1[j]);
                        const float2 cq = __half22float2(hp2[j]), dq = __half22float2(hp3[j]);
                        const float vx = fmaf(w0, a.x, fmaf(w1, bq.x, fmaf(w2, cq.x, fmaf(w3, dq.x, lf.x))));
                        const float vy = fmaf(w0, a.y, fmaf(w1, bq.y, fmaf(w2, cq.y, fmaf(w3, dq.y, lf.y))));
                        const __half2 hq = __floats2half2_rn(vx, vy);
                        const float2 hf2 = __half22float2(hq);
                        h4[j] = hq;
                        l4[j] = __floats2half2_rn(vx - hf2.x, vy - hf2.y);
                    }
                    if (u == 0) mbar_wait(&empty[rs], rphase ^ 1u);
                    *reinterpret_cast<uint4 *>(sdst + u * A_OCT_B) = *reinterpret_cast<const uint4 *>(h4);
                    *reinterpret_cast<uint4 *>(sdst + A_HALF + u * A_OCT_B) = *reinterpret_cast<const uint4 *>(l4);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[rs]);
                mt = nx; lv = nlv; y = ny; xx = nxx; b = nb;
                kc = nkc_;
                t = nt_;
            }
            ring += n_steps;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
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
    C2M_CHECK_ARG(a && ((a->x_hi && a->x_lo) || a->x_il) && a->om && a->packed_w, "dcn_v2_fused_tc: null pointer");
    C2M_CHECK_ARG(!a->x_il || (reinterpret_cast<uintptr_t>(a->x_il) & 31) == 0, "dcn_v2_fused_tc: x_il must be 32 B aligned");
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
    p.cs = 1;
    p.C8out = (a->Cout + 7) / 8; p.Hout = a->H; p.Wout = a->W;
    p.os_b = a->os_b; p.os_c = a->os_c; p.os_y = a->os_y; p.os_x = a->os_x;
    p.f32_mode = a->out_f32 ? f32_store_mode(a->out_f32, nullptr, a->os_b, a->os_c, a->os_y, a->os_x) : 0;
    ConvPtrs q = {};
    q.wblob = reinterpret_cast<const uint8_t *>(a->packed_w);
    q.bias = a->bias;
    q.out_hi = reinterpret_cast<__half *>(a->out_hi); q.out_lo = reinterpret_cast<__half *>(a->out_lo);
    q.out_f32 = a->out_f32;
    DcnTc d;
    d.policy = 5;      // measured (profiles/r02_dcn_experiments.md): hints change nothing beyond noise; 2 (offsets evict-first) costs 3 %
    if (const char *ev = getenv("C2M_DCN_POLICY")) d.policy = atoi(ev);     // tuning experiments
    p.cs = (d.policy & 4) ? 1 : 0;
    if (const char *ev = getenv("C2M_L2_PERSIST_MB")) {                      // experiment: L2 set-aside for evict-last lines
        static int done = 0;
        if (!done) { cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)atoi(ev) << 20); done = 1; }
    }
    d.x_hi = reinterpret_cast<const __half *>(a->x_hi); d.x_lo = reinterpret_cast<const __half *>(a->x_lo);
    d.x_il = reinterpret_cast<const __half *>(a->x_il);
    if (!d.x_hi) { d.x_hi = d.x_il; d.x_lo = d.x_il; }     // never dereferenced when x_il is set; keep pointers valid
    d.C8 = a->C / 8;
    d.om_c8 = a->om_octets ? (27 * a->dg + 7) / 8 : 0;
    C2M_CHECK_ARG(!a->om_octets || (reinterpret_cast<uintptr_t>(a->om) & 7) == 0, "dcn_v2_fused_tc: om must be 8 B aligned");
    C2M_CHECK_ARG(!a->mask || (!a->om_octets && !a->pre && !a->idx),
                  "dcn_v2_fused_tc: a final mask excludes octet-planar om and pre-offsets");
    d.mask = a->mask;
    d.om = a->om; d.pre = a->pre; d.idx = reinterpret_cast<const long long *>(a->idx);
    d.gh = a->gh; d.gw = a->gw; d.ref_gw = a->ref_gw; d.pre_scale = a->pre_scale;
    d.C = a->C; d.dg = a->dg; d.cpg = a->C / a->dg; d.opp = d.cpg / 8; d.n_ko = (a->C / 8) * 9;
    d.opp_shift = -1;
    for (int sft = 0; sft < 8; ++sft)
        if (d.opp == (1 << sft)) d.opp_shift = sft;
    // K octets per gather thread: the (g, tap) run inside a 4-octet stage when C/dg/8 is a power of two
    d.L = d.opp_shift < 0 ? 1 : (d.opp >= KOCT ? KOCT : d.opp);
    d.lshift = d.L == 4 ? 2 : d.L == 2 ? 1 : 0;
    d.inv_ref_gw = a->ref_gw > 0 ? 1.f / (float)a->ref_gw : 0.f;
    d.inv_scale = a->pre_scale > 0 ? 1.f / (float)a->pre_scale : 1.f;
    // flow table: index-map cells a 16x8 tile (+ the 2-cell tap shift) can touch, when the scale divides the tile
    d.sc_shift = -1;
    d.tab_h = d.tab_w = 0;
    if (!a->pre && a->idx)
        for (int sft = 0; sft < 4; ++sft)
            if (a->pre_scale == (1 << sft)) {
                d.sc_shift = sft;
                d.tab_h = (T_R >> sft) + 2;
                d.tab_w = (T_C >> sft) + 2;
            }
    C2M_CHECK_ARG((long long)a->H * a->W * (27 * a->dg + 7) < (1ll << 31) && (long long)a->H * a->W * a->C < (1ll << 31),
                  "dcn_v2_fused_tc: map too large for 32-bit in-kernel indexing");
    C2M_CHECK_ARG(a->idx == nullptr || (long long)a->gh * a->gw < (1 << 23), "dcn_v2_fused_tc: index map too large");
    C2M_CHECK_ARG((long long)a->B * a->H * a->W * a->C < (1ll << 31), "dcn_v2_fused_tc: batch too large for 32-bit in-kernel indexing");
    // ring depths (powers of two): with L octets per thread, L stage groups fill consecutive stages concurrently
    const size_t w_chunk = (size_t)2 * KOCT * p.N * 16;
    d.nstage = d.L == 1 ? 4 : 8;
    d.nstage_shift = d.L == 1 ? 2 : 3;
    const size_t fixed = (size_t)d.nstage * A_STAGE + (2 * MAX_NSTAGE + 2 * MAX_NBST + 2 * MAXT) * 8 + 16 + 256 * 4 +
                         2 * MAXT * 16 + (size_t)2 * p.T * d.tab_h * d.tab_w * 8 + 1024 + 128;
    d.nbst = fixed + 4 * w_chunk <= 227 * 1024 ? 4 : 2;
    const size_t smem = fixed + d.nbst * w_chunk;
    C2M_CHECK_ARG(smem <= 227 * 1024, "dcn_v2_fused_tc: %zu bytes of shared memory needed", smem);
    C2M_CUDA(cudaFuncSetAttribute(dcn_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 0;
    C2M_CUDA(cudaGetDevice(&dev));
    C2M_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long n_work = (long long)a->B * p.tiles_x * p.tiles_y;
    C2M_CHECK_ARG(n_work < (1 << 23), "dcn_v2_fused_tc: too many pixel tiles");
    // tiles per work item: all T accumulators when there is plenty of work, fewer so that every SM gets an item otherwise
    d.tg = (int)((n_work + sms - 1) / sms);
    if (d.tg > p.T) d.tg = p.T;
    if (d.tg < 1) d.tg = 1;
    const int n_items = (int)((n_work + d.tg - 1) / d.tg);
    // flops = 2*B*Cout*C*9*Ho*Wo; bytes = 4*B*(C*H*W + 3*dg*9*H*W + Cout*H*W) + weights
    const double px = (double)a->B * a->H * a->W;
    const double flops = 2.0 * a->Cout * a->C * 9.0 * px;
    const double bytes = 4.0 * px * (a->C + 27.0 * a->dg + a->Cout) + 4.0 * a->Cout * (a->C * 9.0 + 1.0);
    void *ph = prof_begin(PROF_DCN, flops, bytes, st);
    dcn_umma_kernel<<<n_items < sms ? n_items : sms, NTHREADS, smem, st>>>(q, p, d);
    C2M_LAUNCH_CHECK("dcn_umma_kernel");
    prof_end(ph, st);
    return C2M_OK;
}
