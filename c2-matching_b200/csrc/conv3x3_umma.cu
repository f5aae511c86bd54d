// 3x3 / stride 1 / pad 1 convolution on the sm_100a tensor cores with fp32-grade accuracy
// ("fp16 x 3": split operands hi + lo, hi*hi + hi*lo + lo*hi accumulated in fp32 TMEM).
//
// once the correlation and DCN are fused, the 128 plain 3x3 convolutions of
// RestorationNet (ResidualBlockNoBN chains, mmsr/models/archs/arch_util.py:80-136;
// ref_restoration_arch.py:22-27,92-95,...) are ~80 % of the step when run as exact-fp32 cuDNN
// SIMT kernels (profiles/r01_launches_step_fp32.md).  This kernel keeps fp32-grade results
// (<= 2^-21 relative per product) at tensor-core speed.
//
// Same trick as corr_umma.cu: activations live in the packed-split layout ("PSA")
//      hi, lo : fp16 [B][C/8][H][W][8]
// so ONE TMA box (10 cols x 18 rows x 4 channel octets, origin (x0-1, y0-1): padding = TMA
// out-of-bounds zero fill) per operand half stages a halo'd 16x8-pixel tile in the tcgen05
// K-major no-swizzle layout, and the 9 filter taps are 9 MMAs whose A descriptors differ only
// in start address.  The B operand is the weight matrix [Cout x Cin] of the tap, pre-packed
// (c2m_conv3x3_pack_weights_f32) as split fp16 [Cin/32][tap][octet][Cout][8] and kept RESIDENT in
// shared memory for the whole persistent CTA (<= 147 KB for 64 -> 64), so the only streaming
// traffic is the activation tile: the kernel is HBM-bound on its own input/output, not L2-bound
// on weights.  Epilogue (4 warps, one output pixel per thread): *2^-(sa+sw) + bias, ReLU /
// LeakyReLU, optional residual add (a second PSA tensor), re-split to hi/lo and store — the
// output is directly the next convolution's operand, no fp32 round trip through HBM.
#include "c2m_common.cuh"

namespace c2m {

namespace {
constexpr int T_R = 16, T_C = 8;          // output pixel tile -> UMMA M = 128
constexpr int A_R = T_R + 2, A_C = T_C + 2;
constexpr int KOCT = 4;                   // channel octets per A stage
constexpr int ROW_B = A_C * 16;           // 160 B  (SBO of A)
constexpr int A_OCT_B = A_R * ROW_B;      // 2880 B (LBO of A)
constexpr int A_HALF = KOCT * A_OCT_B;    // 11520 B (hi or lo of one stage)
constexpr int A_STAGE = 2 * A_HALF;       // 23040 B
constexpr int NSTAGE = 3;
constexpr int NACC = 4;
constexpr int W_HDR = 256;                // packed-weight blob header bytes

struct ConvParams {
    int B, C8in, H, W;        // input (C8in = channel octets present in the activation tensor)
    int Cout, N;              // real / padded (multiple of 16) output channels
    int nkc;                  // K chunks of KOCT octets (weights are padded to nkc*KOCT octets)
    int tiles_x, tiles_y;
    int act;                  // 0 none, 1 relu, 2 leaky relu 0.1
    int sa_in, sa_res, sa_out;    // activation scale exponents (values are stored * 2^sa)
    int C8out;
};
}  // namespace

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1)
conv3x3_umma_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
                    const uint8_t *__restrict__ wblob, const float *__restrict__ bias,
                    const __half *__restrict__ res_hi, const __half *__restrict__ res_lo,
                    __half *__restrict__ out_hi, __half *__restrict__ out_lo, const ConvParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t w_half = (uint32_t)p.nkc * 9 * KOCT * p.N * 16;      // bytes of W_hi (== W_lo)
    uint8_t *sW = smem;                                                  // [hi | lo]
    uint8_t *sA = smem + 2 * w_half;
    sA = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(sA) + 127) & ~uintptr_t(127));
    uint64_t *bars = reinterpret_cast<uint64_t *>(sA + NSTAGE * A_STAGE);
    uint64_t *full = bars, *empty = bars + NSTAGE, *tfull = bars + 2 * NSTAGE, *tempty = tfull + NACC,
             *wbar = tempty + NACC;
    uint32_t *tmem_base_p = reinterpret_cast<uint32_t *>(wbar + 1);
    float *sbias = reinterpret_cast<float *>(tmem_base_p + 2);           // [N]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = p.B * p.tiles_x * p.tiles_y;
    const uint32_t tmem_cols = p.N * NACC <= 32 ? 32 : p.N * NACC <= 64 ? 64 : p.N * NACC <= 128 ? 128
                               : p.N * NACC <= 256 ? 256 : 512;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_hi);
        tma_prefetch_desc(&tm_lo);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < NACC; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
        mbar_init(wbar, 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_base_p, tmem_cols);
        tmem_relinquish();
    }
    for (int i = threadIdx.x; i < p.N; i += blockDim.x) sbias[i] = (bias && i < p.Cout) ? bias[i] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_p;
    const int sw = *reinterpret_cast<const int *>(wblob);                // weight scale exponent

    if (warp == 0) {
        // ================================ producer ==========================================
        if (lane == 0) {
            // weights: one bulk copy per half, once per CTA
            mbar_arrive_expect_tx(wbar, 2 * w_half);
            const uint8_t *src = wblob + W_HDR;
            for (uint32_t off = 0; off < 2 * w_half; off += 65536) {
                const uint32_t n = min(65536u, 2 * w_half - off);
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                    ::"r"(smem_u32(sW + off)), "l"(src + off), "r"(n), "r"(smem_u32(wbar))
                    : "memory");
            }
            int stage = 0, phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int b = tile / (p.tiles_x * p.tiles_y), tt = tile % (p.tiles_x * p.tiles_y);
                const int y0 = (tt / p.tiles_x) * T_R, x0 = (tt % p.tiles_x) * T_C;
                for (int kc = 0; kc < p.nkc; ++kc) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t *s = sA + stage * A_STAGE;
                    mbar_arrive_expect_tx(&full[stage], A_STAGE);
                    tma_load_4d(s, &tm_hi, &full[stage], (x0 - 1) * 8, y0 - 1, kc * KOCT, b);
                    tma_load_4d(s + A_HALF, &tm_lo, &full[stage], (x0 - 1) * 8, y0 - 1, kc * KOCT, b);
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ========================================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f16(128, p.N, 0);
            const uint32_t w_hi = smem_u32(sW), w_lo = w_hi + w_half;
            const uint32_t b_lbo = p.N * 16;                 // next channel octet of the weights
            mbar_wait(wbar, 0);
            tc_fence_after();
            int stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + acc * p.N;
                for (int kc = 0; kc < p.nkc; ++kc) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(sA + stage * A_STAGE), a_lo = a_hi + A_HALF;
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const uint32_t toff = ((tap / 3) * A_C + tap % 3) * 16;
                        const uint32_t wtap = ((kc * 9 + tap) * KOCT) * b_lbo;
#pragma unroll
                        for (int j = 0; j < KOCT / 2; ++j) {
                            const uint32_t ao = toff + j * 2 * A_OCT_B, bo = wtap + j * 2 * b_lbo;
                            const uint64_t dah = umma_smem_desc(a_hi + ao, A_OCT_B, ROW_B);
                            const uint64_t dal = umma_smem_desc(a_lo + ao, A_OCT_B, ROW_B);
                            const uint64_t dbh = umma_smem_desc(w_hi + bo, b_lbo, 128);
                            const uint64_t dbl = umma_smem_desc(w_lo + bo, b_lbo, 128);
                            umma_f16(d, dah, dbh, idesc, (kc | tap | j) != 0);
                            umma_f16(d, dah, dbl, idesc, 1);
                            umma_f16(d, dal, dbh, idesc, 1);
                        }
                    }
                    umma_commit(&empty[stage]);
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull[acc]);
                if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ================================ epilogue ==========================================
        const int e = threadIdx.x - 128;
        const int quarter = warp & 3;
        const float out_scale = ldexpf(1.f, -(p.sa_in + sw));
        const float res_scale = ldexpf(1.f, -p.sa_res);
        const float so = ldexpf(1.f, p.sa_out);
        int acc = 0, acc_phase = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int b = tile / (p.tiles_x * p.tiles_y), tt = tile % (p.tiles_x * p.tiles_y);
            const int y = (tt / p.tiles_x) * T_R + e / T_C, x = (tt % p.tiles_x) * T_C + e % T_C;
            const bool ok = y < p.H && x < p.W;
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * p.N;
            for (int c0 = 0; c0 < p.N; c0 += 32) {
                uint32_t reg[32];
                if (p.N - c0 >= 32) {
                    tmem_ld_32x32(taddr + c0, reg);
                } else {   // N = 16 tail: 16 columns
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                        : "=r"(reg[0]), "=r"(reg[1]), "=r"(reg[2]), "=r"(reg[3]), "=r"(reg[4]), "=r"(reg[5]),
                          "=r"(reg[6]), "=r"(reg[7]), "=r"(reg[8]), "=r"(reg[9]), "=r"(reg[10]), "=r"(reg[11]),
                          "=r"(reg[12]), "=r"(reg[13]), "=r"(reg[14]), "=r"(reg[15])
                        : "r"(taddr + c0)
                        : "memory");
                }
                tmem_ld_wait();
                if (ok) {
                    const int ncol = min(32, p.N - c0);
#pragma unroll
                    for (int o8 = 0; o8 < 4; ++o8) {
                        const int oct = c0 / 8 + o8;
                        if (o8 * 8 >= ncol || oct >= p.C8out) break;
                        const size_t off = ((((size_t)b * p.C8out + oct) * p.H + y) * p.W + x) * 8;
                        float r8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                        if (res_hi) {
                            const uint4 rh = *reinterpret_cast<const uint4 *>(res_hi + off);
                            const uint4 rl = *reinterpret_cast<const uint4 *>(res_lo + off);
                            const __half *hh = reinterpret_cast<const __half *>(&rh);
                            const __half *ll = reinterpret_cast<const __half *>(&rl);
#pragma unroll
                            for (int j = 0; j < 8; ++j) r8[j] = (__half2float(hh[j]) + __half2float(ll[j])) * res_scale;
                        }
                        __align__(16) __half h8[8];
                        __align__(16) __half l8[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int o = c0 + o8 * 8 + j;
                            float v = fmaf(__uint_as_float(reg[o8 * 8 + j]), out_scale, sbias[o]);
                            if (p.act == 1) v = fmaxf(v, 0.f);
                            else if (p.act == 2) v = v > 0.f ? v : v * 0.1f;
                            v += r8[j];
                            if (o >= p.Cout) v = 0.f;
                            const float vs = v * so;
                            const __half hh = __float2half_rn(vs);
                            h8[j] = hh;
                            l8[j] = __float2half_rn(vs - __half2float(hh));
                        }
                        *reinterpret_cast<uint4 *>(out_hi + off) = *reinterpret_cast<const uint4 *>(h8);
                        *reinterpret_cast<uint4 *>(out_lo + off) = *reinterpret_cast<const uint4 *>(l8);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------
// PSA <-> strided fp32 converters and the weight pre-pack
// ------------------------------------------------------------------------------------------
__global__ void psa_from_f32_kernel(const float *__restrict__ x, int C, int C8, int H, int W, long long xs_b,
                                    long long xs_c, long long xs_y, long long xs_x, int sa, __half *__restrict__ hi,
                                    __half *__restrict__ lo) {
    // one thread = one (octet, pixel); pixel fastest -> coalesced NCHW reads and 16 B stores
    const long long HW = (long long)H * W;
    const long long n = (long long)C8 * HW;
    const int b = blockIdx.y;
    const float S = ldexpf(1.f, sa);
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int oct = (int)(e / HW);
        const long long pix = e % HW;
        const int y = (int)(pix / W), xx = (int)(pix % W);
        __align__(16) __half h8[8];
        __align__(16) __half l8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = oct * 8 + j;
            const float v = c < C ? x[b * xs_b + c * xs_c + y * xs_y + xx * xs_x] * S : 0.f;
            const __half hh = __float2half_rn(v);
            h8[j] = hh;
            l8[j] = __float2half_rn(v - __half2float(hh));
        }
        const size_t o = (((size_t)b * C8 + oct) * HW + pix) * 8;
        *reinterpret_cast<uint4 *>(hi + o) = *reinterpret_cast<const uint4 *>(h8);
        *reinterpret_cast<uint4 *>(lo + o) = *reinterpret_cast<const uint4 *>(l8);
    }
}

__global__ void psa_to_f32_kernel(const __half *__restrict__ hi, const __half *__restrict__ lo, int C, int C8, int H,
                                  int W, int sa, const float *__restrict__ add, float *__restrict__ out, long long os_b,
                                  long long os_c, long long os_y, long long os_x) {
    const long long HW = (long long)H * W;
    const long long n = (long long)C8 * HW;
    const int b = blockIdx.y;
    const float S = ldexpf(1.f, -sa);
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int oct = (int)(e / HW);
        const long long pix = e % HW;
        const int y = (int)(pix / W), xx = (int)(pix % W);
        const size_t o = (((size_t)b * C8 + oct) * HW + pix) * 8;
        const uint4 rh = *reinterpret_cast<const uint4 *>(hi + o);
        const uint4 rl = *reinterpret_cast<const uint4 *>(lo + o);
        const __half *hh = reinterpret_cast<const __half *>(&rh);
        const __half *ll = reinterpret_cast<const __half *>(&rl);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = oct * 8 + j;
            if (c < C) {
                const long long oi = b * os_b + c * os_c + y * os_y + xx * os_x;
                float v = (__half2float(hh[j]) + __half2float(ll[j])) * S;
                if (add) v += add[oi];
                out[oi] = v;
            }
        }
    }
}

__global__ void wamax_kernel(const float *__restrict__ w, int n, unsigned *__restrict__ bits) {
    float m = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = fmaxf(m, fabsf(w[i]));
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(bits, __float_as_uint(m));
}

// blob: [int sw][unsigned amax_bits]...pad to 256 B | W_hi [nkc][9][KOCT][N][8] | W_lo (same)
__global__ void wpack_kernel(const float *__restrict__ w, int Cin, int Cout, int N, int nkc, uint8_t *__restrict__ blob) {
    __shared__ int s_sw;
    if (threadIdx.x == 0) {
        const float a = __uint_as_float(reinterpret_cast<unsigned *>(blob)[1]);
        int e = 0;
        if (a > 0.f && isfinite(a)) frexpf(a, &e);
        s_sw = 14 - e;                               // |w| * 2^sw <= 2^14
        if (blockIdx.x == 0) reinterpret_cast<int *>(blob)[0] = s_sw;
    }
    __syncthreads();
    const float S = ldexpf(1.f, s_sw);
    const int half_elems = nkc * 9 * KOCT * N * 8;
    __half *hi = reinterpret_cast<__half *>(blob + W_HDR);
    __half *lo = hi + half_elems;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < half_elems; e += gridDim.x * blockDim.x) {
        const int j = e & 7, o = (e >> 3) % N, rest = (e >> 3) / N;
        const int oct = rest % KOCT, tap = (rest / KOCT) % 9, kc = rest / (KOCT * 9);
        const int c = (kc * KOCT + oct) * 8 + j;
        float v = 0.f;
        if (c < Cin && o < Cout) v = w[((size_t)o * Cin + c) * 9 + tap] * S;
        const __half hh = __float2half_rn(v);
        hi[e] = hh;
        lo[e] = __float2half_rn(v - __half2float(hh));
    }
}

// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_act_map(CUtensorMap *m, const void *base, int B, int C8, int H, int W) {
    static EncodeTiledFn enc = nullptr;
    if (!enc) {
        void *pfn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &pfn, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess) {
            set_error("cuTensorMapEncodeTiled entry point not available");
            return C2M_ERR_UNSUPPORTED;
        }
        enc = reinterpret_cast<EncodeTiledFn>(pfn);
    }
    cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)C8, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)C8 * H * W * 16};
    cuuint32_t box[4] = {(cuuint32_t)A_C * 8, (cuuint32_t)A_R, (cuuint32_t)KOCT, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return C2M_ERR_CUDA; }
    return C2M_OK;
}

static inline int pad16(int c) { return (c + 15) / 16 * 16; }
static inline int n_kc(int cin) { return ((cin + 7) / 8 + KOCT - 1) / KOCT; }

}  // namespace c2m

using namespace c2m;

extern "C" size_t c2m_conv3x3_packed_weight_bytes(int Cin, int Cout) {
    if (Cin <= 0 || Cout <= 0) return 0;
    return (size_t)W_HDR + 2 * (size_t)n_kc(Cin) * 9 * KOCT * pad16(Cout) * 16;
}

extern "C" int c2m_conv3x3_supported(int Cin, int Cout) {
    if (Cin <= 0 || Cout <= 0 || pad16(Cout) > 128) return 0;
    const size_t smem = 2 * (size_t)n_kc(Cin) * 9 * KOCT * pad16(Cout) * 16 + NSTAGE * A_STAGE + 4096;
    return smem <= 227 * 1024 ? 1 : 0;
}

extern "C" int c2m_conv3x3_pack_weights_f32(const float *w, int Cin, int Cout, void *packed, c2m_stream_t stream) {
    C2M_CHECK_ARG(w && packed && Cin > 0 && Cout > 0, "conv3x3_pack_weights: bad argument");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    C2M_CUDA(cudaMemsetAsync(packed, 0, W_HDR, st));
    const int n = Cout * Cin * 9;
    wamax_kernel<<<ceil_div(n, 1024) > 64 ? 64 : ceil_div(n, 1024), 256, 0, st>>>(
        w, n, reinterpret_cast<unsigned *>(packed) + 1);
    C2M_LAUNCH_CHECK("wamax_kernel");
    const int N = pad16(Cout), nkc = n_kc(Cin);
    const int total = nkc * 9 * KOCT * N * 8;
    wpack_kernel<<<ceil_div(total, 256) > 296 ? 296 : ceil_div(total, 256), 256, 0, st>>>(
        w, Cin, Cout, N, nkc, reinterpret_cast<uint8_t *>(packed));
    C2M_LAUNCH_CHECK("wpack_kernel");
    return C2M_OK;
}

extern "C" int c2m_psa_from_f32(const float *x, int B, int C, int H, int W, long long xs_b, long long xs_c,
                                long long xs_y, long long xs_x, int sa, void *hi, void *lo, c2m_stream_t stream) {
    C2M_CHECK_ARG(x && hi && lo && B > 0 && C > 0 && H > 0 && W > 0, "psa_from_f32: bad argument");
    const int C8 = (C + 7) / 8;
    const long long n = (long long)C8 * H * W;
    int bx = (int)((n + 255) / 256);
    if (bx > 4096) bx = 4096;
    psa_from_f32_kernel<<<dim3(bx, B), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        x, C, C8, H, W, xs_b, xs_c, xs_y, xs_x, sa, reinterpret_cast<__half *>(hi), reinterpret_cast<__half *>(lo));
    C2M_LAUNCH_CHECK("psa_from_f32_kernel");
    return C2M_OK;
}

extern "C" int c2m_psa_to_f32(const void *hi, const void *lo, int B, int C, int H, int W, int sa, const float *add,
                              float *out, long long os_b, long long os_c, long long os_y, long long os_x,
                              c2m_stream_t stream) {
    C2M_CHECK_ARG(out && hi && lo && B > 0 && C > 0 && H > 0 && W > 0, "psa_to_f32: bad argument");
    const int C8 = (C + 7) / 8;
    const long long n = (long long)C8 * H * W;
    int bx = (int)((n + 255) / 256);
    if (bx > 4096) bx = 4096;
    psa_to_f32_kernel<<<dim3(bx, B), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __half *>(hi), reinterpret_cast<const __half *>(lo), C, C8, H, W, sa, add, out, os_b,
        os_c, os_y, os_x);
    C2M_LAUNCH_CHECK("psa_to_f32_kernel");
    return C2M_OK;
}

extern "C" int c2m_conv3x3_psa(const void *in_hi, const void *in_lo, int B, int Cin, int H, int W, int sa_in,
                               const void *packed_w, const float *bias, int Cout, int act, const void *res_hi,
                               const void *res_lo, int sa_res, void *out_hi, void *out_lo, int sa_out,
                               c2m_stream_t stream) {
    C2M_CHECK_ARG(in_hi && in_lo && packed_w && out_hi && out_lo, "conv3x3_psa: null pointer");
    C2M_CHECK_ARG(B > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0, "conv3x3_psa: bad shape");
    C2M_CHECK_ARG(c2m_conv3x3_supported(Cin, Cout), "conv3x3_psa: %d -> %d channels exceed the resident-weight kernel",
                  Cin, Cout);
    C2M_CHECK_ARG((res_hi == nullptr) == (res_lo == nullptr), "conv3x3_psa: residual needs both halves");
    C2M_CHECK_ARG(act >= 0 && act <= 2, "conv3x3_psa: unknown activation %d", act);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    ConvParams p;
    p.B = B; p.C8in = (Cin + 7) / 8; p.H = H; p.W = W;
    p.Cout = Cout; p.N = pad16(Cout); p.nkc = n_kc(Cin);
    p.tiles_x = ceil_div(W, T_C); p.tiles_y = ceil_div(H, T_R);
    p.act = act; p.sa_in = sa_in; p.sa_res = sa_res; p.sa_out = sa_out;
    p.C8out = (Cout + 7) / 8;
    CUtensorMap mh, ml;
    int rc;
    if ((rc = make_act_map(&mh, in_hi, B, p.C8in, H, W))) return rc;
    if ((rc = make_act_map(&ml, in_lo, B, p.C8in, H, W))) return rc;
    const size_t smem = 2 * (size_t)p.nkc * 9 * KOCT * p.N * 16 + NSTAGE * A_STAGE + 4096 + 1024;
    C2M_CUDA(cudaFuncSetAttribute(conv3x3_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 0;
    C2M_CUDA(cudaGetDevice(&dev));
    C2M_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int n_tiles = B * p.tiles_x * p.tiles_y;
    conv3x3_umma_kernel<<<n_tiles < sms ? n_tiles : sms, 256, smem, st>>>(
        mh, ml, reinterpret_cast<const uint8_t *>(packed_w), bias, reinterpret_cast<const __half *>(res_hi),
        reinterpret_cast<const __half *>(res_lo), reinterpret_cast<__half *>(out_hi), reinterpret_cast<__half *>(out_lo), p);
    C2M_LAUNCH_CHECK("conv3x3_umma_kernel");
    return C2M_OK;
}
