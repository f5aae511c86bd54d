// Fused all-pairs patch correlation + running top-2 on the sm_100a tensor cores.
//
// Replaces the reference's chunked `F.conv2d(feat_in, ref_patches)` + `max`
// (mmsr/models/archs/ref_map_util.py:54-76), which materialises a [n_ref, h', w'] score tensor
// per chunk (2 x 2.09 GB per image at 160x160 maps).  Here no score ever reaches HBM.
//
// Formulation.  score(q,r) = sum_{tap in 3x3} sum_c in[q+tap][c] * ref[r+tap][c] is a GEMM with
// K = 9*C whose A/B rows for tap (dy,dx) are the SAME pixel rows shifted by (dy,dx).  Each CTA
// therefore stages, per 32-channel slice, one halo'd pixel block per side —
//   A: (16+2) x (8+2) input pixels   B: (32+2) x (8+2) Ref pixels
// — with one TMA box each into the tcgen05 K-major *no-swizzle* layout [octet][pixel][8 halfs]
// (rows 16 B apart), and issues the 9 taps as 9 MMAs whose shared-memory descriptors differ only
// in their start address (+ (dy*10+dx)*16 B).  8-row core-matrix groups are tile rows, so the
// stride-byte-offset is the block row pitch (10 px * 16 B); the two K core matrices of a
// UMMA_K=16 step are `octet pitch` apart (leading-byte-offset).  Operand traffic from L2 is 9x
// lower than an im2col GEMM and nothing is re-laid-out in shared memory.
//
// Precision.  Operands are split fp16 pairs of x*2^sexp (hi + lo, 22 mantissa bits); each K step
// issues hi*hi + hi*lo + lo*hi into the same fp32 TMEM accumulator (M=128, N=256).  The scores
// only RANK candidates: the epilogue keeps a top-2 per query and the exact rescoring kernel
// (corr_aux.cu) decides, so tensor-core rounding cannot leak into the index map.
//
// CTA = 256 threads: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7
// epilogue (one query row each, TMEM lane quarter = warp % 4).  3-stage smem ring (66.5 KB per
// stage), 2 x 256-column TMEM accumulators so the epilogue of Ref tile n overlaps the MMAs of
// tile n+1.  Persistent: grid = #SMs, work item = (image, query tile, Ref chunk).
#include "corr_internal.cuh"

namespace c2m {

namespace {
constexpr int PATCH = 3;
constexpr int TQ_R = 16, TQ_C = 8;                 // query tile -> UMMA M = 128
constexpr int TR_R = 32, TR_C = 8;                 // Ref tile   -> UMMA N = 256
constexpr int AB_C = TQ_C + PATCH - 1;             // 10 block columns (both sides)
constexpr int A_R = TQ_R + PATCH - 1;              // 18
constexpr int B_R = TR_R + PATCH - 1;              // 34
constexpr int KOCT = 4;                            // channel octets per stage (32 channels)
constexpr int ROW_B = AB_C * 16;                   // 160 B block row pitch  (= SBO)
constexpr int A_OCT_B = A_R * ROW_B;               // 2880 B octet pitch of A (= LBO)
constexpr int B_OCT_B = B_R * ROW_B;               // 5440 B octet pitch of B (= LBO)
constexpr int A_BYTES = KOCT * A_OCT_B;            // 11520
constexpr int B_BYTES = KOCT * B_OCT_B;            // 21760
constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);   // hi+lo of both sides = 66560
constexpr int NSTAGE = 3;
constexpr int UM = TQ_R * TQ_C, UN = TR_R * TR_C;  // 128, 256
constexpr int TMEM_COLS = 512;
constexpr int SMEM_AUX = 2 * UN * 4 + 2 * UN * 4 + 128;   // rinv + ridx double buffers + barriers
constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + SMEM_AUX + 1024;
static_assert(A_BYTES % 128 == 0 && B_BYTES % 128 == 0, "TMA destinations must stay 128B aligned");

struct UmmaParams {
    int B, C8;
    int gh, gw, rh, rw;          // patch grids
    int qt_x, qt_y, rt_x, rt_y;  // tile counts
    int nchunk, rt_per_chunk;
    int NQ, NR;
};
}  // namespace

__global__ void __launch_bounds__(256, 1)
corr_umma_kernel(const __grid_constant__ CUtensorMap tm_in_hi, const __grid_constant__ CUtensorMap tm_in_lo,
                 const __grid_constant__ CUtensorMap tm_ref_hi, const __grid_constant__ CUtensorMap tm_ref_lo,
                 const float *__restrict__ rinv, const int *__restrict__ sexp, Candidate *__restrict__ part,
                 const UmmaParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *aux = smem + NSTAGE * STAGE_BYTES;
    float *rs = reinterpret_cast<float *>(aux);                     // [2][UN]
    int *ri = reinterpret_cast<int *>(aux + 2 * UN * 4);            // [2][UN]
    uint64_t *bars = reinterpret_cast<uint64_t *>(aux + 4 * UN * 4);
    uint64_t *full = bars, *empty = bars + NSTAGE, *tfull = bars + 2 * NSTAGE, *tempty = bars + 2 * NSTAGE + 2;
    uint32_t *tmem_base_p = reinterpret_cast<uint32_t *>(bars + 2 * NSTAGE + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_in_hi);
        tma_prefetch_desc(&tm_in_lo);
        tma_prefetch_desc(&tm_ref_hi);
        tma_prefetch_desc(&tm_ref_lo);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_base_p, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_p;

    const int n_qt = p.qt_x * p.qt_y, n_rt = p.rt_x * p.rt_y;
    const int n_items = p.B * n_qt * p.nchunk;
    const int n_kc = (p.C8 + KOCT - 1) / KOCT;

    if (warp == 0) {
        // ================================ TMA producer =====================================
        if (lane == 0) {
            int stage = 0, phase = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int chunk = item % p.nchunk, qt = (item / p.nchunk) % n_qt, b = item / (p.nchunk * n_qt);
                const int qy0 = (qt / p.qt_x) * TQ_R, qx0 = (qt % p.qt_x) * TQ_C;
                const int rt_b = chunk * p.rt_per_chunk, rt_e = min(n_rt, rt_b + p.rt_per_chunk);
                for (int rt = rt_b; rt < rt_e; ++rt) {
                    const int ry0 = (rt / p.rt_x) * TR_R, rx0 = (rt % p.rt_x) * TR_C;
                    for (int kc = 0; kc < n_kc; ++kc) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        uint8_t *s = smem + stage * STAGE_BYTES;
                        mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
                        tma_load_4d(s, &tm_in_hi, &full[stage], qx0 * 8, qy0, kc * KOCT, b);
                        tma_load_4d(s + A_BYTES, &tm_in_lo, &full[stage], qx0 * 8, qy0, kc * KOCT, b);
                        tma_load_4d(s + 2 * A_BYTES, &tm_ref_hi, &full[stage], rx0 * 8, ry0, kc * KOCT, b);
                        tma_load_4d(s + 2 * A_BYTES + B_BYTES, &tm_ref_lo, &full[stage], rx0 * 8, ry0, kc * KOCT, b);
                        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer =======================================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(UM, UN, 0);
            int stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int chunk = item % p.nchunk;
                const int rt_b = chunk * p.rt_per_chunk, rt_e = min(n_rt, rt_b + p.rt_per_chunk);
                for (int rt = rt_b; rt < rt_e; ++rt) {
                    mbar_wait(&tempty[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d = tmem_base + acc * UN;
                    for (int kc = 0; kc < n_kc; ++kc) {
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                        const uint32_t a_hi = sa, a_lo = sa + A_BYTES, b_hi = sa + 2 * A_BYTES,
                                       b_lo = sa + 2 * A_BYTES + B_BYTES;
#pragma unroll
                        for (int tap = 0; tap < PATCH * PATCH; ++tap) {
                            const uint32_t toff = ((tap / PATCH) * AB_C + tap % PATCH) * 16;
#pragma unroll
                            for (int j = 0; j < KOCT / 2; ++j) {
                                const uint32_t ao = toff + j * 2 * A_OCT_B, bo = toff + j * 2 * B_OCT_B;
                                const uint64_t dah = umma_smem_desc(a_hi + ao, A_OCT_B, ROW_B);
                                const uint64_t dal = umma_smem_desc(a_lo + ao, A_OCT_B, ROW_B);
                                const uint64_t dbh = umma_smem_desc(b_hi + bo, B_OCT_B, ROW_B);
                                const uint64_t dbl = umma_smem_desc(b_lo + bo, B_OCT_B, ROW_B);
                                umma_f16(d, dah, dbh, idesc, (kc | tap | j) != 0);
                                umma_f16(d, dah, dbl, idesc, 1);
                                umma_f16(d, dal, dbh, idesc, 1);
                            }
                        }
                        umma_commit(&empty[stage]);        // smem slot free once these MMAs retire
                        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(&tfull[acc]);               // accumulator complete
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        // ================================ epilogue =========================================
        const int e = threadIdx.x - 128;                    // 0..127 = query row m of the tile
        const int quarter = warp & 3;
        const float sinv = ldexpf(1.f, -(sexp[0] + sexp[1]));
        int acc = 0, acc_phase = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int chunk = item % p.nchunk, qt = (item / p.nchunk) % n_qt, b = item / (p.nchunk * n_qt);
            const int qy = (qt / p.qt_x) * TQ_R + e / TQ_C, qx = (qt % p.qt_x) * TQ_C + e % TQ_C;
            const int rt_b = chunk * p.rt_per_chunk, rt_e = min(n_rt, rt_b + p.rt_per_chunk);
            const float *rinvb = rinv + (size_t)b * p.NR;
            float v1 = -INFINITY, v2 = -INFINITY;
            int i1 = 0x7fffffff, i2 = 0x7fffffff;
            for (int rt = rt_b; rt < rt_e; ++rt) {
                const int ry0 = (rt / p.rt_x) * TR_R, rx0 = (rt % p.rt_x) * TR_C;
                float *rsb = rs + acc * UN;
                int *rib = ri + acc * UN;
#pragma unroll
                for (int k = 0; k < UN / 128; ++k) {
                    const int n = e + k * 128;
                    const int ry = ry0 + n / TR_C, rx = rx0 + n % TR_C;
                    const bool ok = ry < p.rh && rx < p.rw;
                    const int r = ry * p.rw + rx;
                    rsb[n] = ok ? rinvb[r] * sinv : 0.f;
                    rib[n] = ok ? r : -1;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                mbar_wait(&tfull[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * UN;
#pragma unroll 1
                for (int cc = 0; cc < UN / 32; ++cc) {
                    uint32_t reg[32];
                    tmem_ld_32x32(taddr + cc * 32, reg);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int r = rib[cc * 32 + j];
                        if (r >= 0) cand_push(__uint_as_float(reg[j]) * rsb[cc * 32 + j], r, v1, i1, v2, i2);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            if (qy < p.gh && qx < p.gw)
                part[((size_t)b * p.nchunk + chunk) * p.NQ + qy * p.gw + qx] =
                    Candidate{v1, i1 == 0x7fffffff ? -1 : i1, v2, i2 == 0x7fffffff ? -1 : i2};
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// map [B][C8][H][W][8] fp16 as a 4-D tensor (W*8, H, C8, B) with a (80, rows, KOCT, 1) box
static int make_map(CUtensorMap *m, const __half *base, int B, int C8, int H, int W, int rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return C2M_ERR_UNSUPPORTED; }
    cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)C8, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)C8 * H * W * 16};
    cuuint32_t box[4] = {(cuuint32_t)AB_C * 8, (cuuint32_t)rows, (cuuint32_t)KOCT, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return C2M_ERR_CUDA; }
    return C2M_OK;
}

bool corr_umma_supported(const CorrGeom &g) {
    return g.patch == PATCH && g.s_in == 1 && g.s_ref == 1 && g.w >= AB_C && g.h >= A_R && g.wr >= AB_C &&
           g.hr >= B_R;
}

int corr_umma_pick_nchunk(const CorrGeom &g) {
    const int n_qt = ceil_div(g.gh, TQ_R) * ceil_div(g.gw, TQ_C);
    const int n_rt = ceil_div(g.rh, TR_R) * ceil_div(g.rw, TR_C);
    int n = ceil_div(8 * 148, g.B * n_qt);
    if (n > 16) n = 16;
    if (n > n_rt) n = n_rt;
    if (n < 1) n = 1;
    // no empty chunks: chunks hold ceil(n_rt/n) tiles
    const int per = ceil_div(n_rt, n);
    return ceil_div(n_rt, per);
}

int corr_search_umma_launch(const CorrGeom &g, const CorrWorkspace &ws, cudaStream_t st) {
    CUtensorMap tih, til, trh, trl;
    const int C8 = g.Cp / 8;
    int rc;
    if ((rc = make_map(&tih, ws.hi_in, g.B, C8, g.h, g.w, A_R))) return rc;
    if ((rc = make_map(&til, ws.lo_in, g.B, C8, g.h, g.w, A_R))) return rc;
    if ((rc = make_map(&trh, ws.hi_ref, g.B, C8, g.hr, g.wr, B_R))) return rc;
    if ((rc = make_map(&trl, ws.lo_ref, g.B, C8, g.hr, g.wr, B_R))) return rc;

    UmmaParams p;
    p.B = g.B; p.C8 = C8;
    p.gh = g.gh; p.gw = g.gw; p.rh = g.rh; p.rw = g.rw;
    p.qt_x = ceil_div(g.gw, TQ_C); p.qt_y = ceil_div(g.gh, TQ_R);
    p.rt_x = ceil_div(g.rw, TR_C); p.rt_y = ceil_div(g.rh, TR_R);
    p.nchunk = ws.nchunk;
    p.rt_per_chunk = ceil_div(p.rt_x * p.rt_y, ws.nchunk);
    p.NQ = g.NQ; p.NR = g.NR;

    // per launch: the attribute belongs to the (device, context) pair, not to the process
    C2M_CUDA(cudaFuncSetAttribute(corr_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    int dev = 0, sms = 0;
    C2M_CUDA(cudaGetDevice(&dev));
    C2M_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int n_items = g.B * p.qt_x * p.qt_y * p.nchunk;
    const int grid = n_items < sms ? n_items : sms;
    corr_umma_kernel<<<grid, 256, SMEM_BYTES, st>>>(tih, til, trh, trl, ws.rinv, ws.sexp, ws.part, p);
    C2M_LAUNCH_CHECK("corr_umma_kernel");
    return C2M_OK;
}

}  // namespace c2m
