// Shared pieces of the tcgen05 implicit-GEMM kernels (plain 3x3 conv, deformable conv):
// parameter blocks and the epilogue that turns one 128-pixel x N-column TMEM accumulator into
// packed-split (PSA) and / or strided fp32 output with bias, activation, residuals and an optional
// fused PixelShuffle(2).
#pragma once
#include "c2m_common.cuh"

namespace c2m {

constexpr int UC_NMAX = 64;      // output channels per CTA slice

struct ConvParams {
    int B, H, W;
    int nkc_a, nkc;           // K chunks taken from input 1 / in total (input 2 supplies the rest)
    int Cout, N, nslice;      // real couts, couts per CTA slice (multiple of 16), slices; slice s starts at s*N
    int tiles_x, tiles_y, T, n_st;   // tile grid, tiles per item, super-tiles per image
    int act;                  // 0 none, 1 relu, 2 leaky relu 0.1
    int sa_in, sa_res, sa_out;
    int ps;                   // 0 or 2: PixelShuffle(2) applied to the PSA output
    int stacked;              // 1: accumulator has 2N columns, value = col[c] + col[N + c]
    int dbg;                  // tuning experiments only (C2M_CONV_DBG): 1 = no global stores, 2 = no TMA reloads
    unsigned w2_off;          // byte offset of the CTA-pair weight layout inside the packed blob
    int cs;                   // 1: PSA output stored with st.global.cs (streaming: the DCN output must not evict the
                              // gathered input map from L2)
    int C8out, Hout, Wout;    // geometry of the PSA output tensor
    long long os_b, os_c, os_y, os_x;   // fp32 output element strides
    int f32_mode;             // fp32 output: 0 strided scalar stores, 1 strided with os_c == 1 (16 B stores),
                              // 2 octet-planar [B][ceil(Cout/8)][H][W][8] (two 16 B stores per octet)
};

// fp32 output mode of a strided [B,C,H,W] view: 16 B stores need contiguous channels and aligned strides
inline int f32_store_mode(const float *out, const float *add, long long os_b, long long os_c, long long os_y, long long os_x) {
    const bool al = ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(add)) & 15) == 0;
    return (os_c == 1 && al && os_b % 4 == 0 && os_y % 4 == 0 && os_x % 4 == 0) ? 1 : 0;
}
struct ConvPtrs {
    const uint8_t *wblob;
    const float *bias;
    const __half *res_hi, *res_lo, *res2_hi, *res2_lo;
    __half *out_hi, *out_lo;
    float *out_f32;
    const float *add_f32;
};

// One thread = one output pixel (TMEM lane); `taddr` already carries the lane quarter and the
// accumulator's first column.
// Residual registers of one 32-column block (4 channel octets); fetched by `prefetch_residual`
// BEFORE the thread blocks on the accumulator / tcgen05.ld, so the HBM/L2 latency of the residual
// overlaps the MMAs instead of being paid once per octet inside the loop.
struct ResRegs {
    uint4 h[4], l[4];
};

__device__ __forceinline__ void prefetch_residual(const __half *rh, const __half *rl, const ConvParams &p, int b, int y,
                                                  int x, bool ok, int o_base, int c0, ResRegs &r) {
#pragma unroll
    for (int o8 = 0; o8 < 4; ++o8) {
        const int oct = (o_base + c0) / 8 + o8;
        if (ok && c0 + o8 * 8 < p.N && oct < p.C8out) {
            const size_t off = ((((size_t)b * p.C8out + oct) * p.H + y) * p.W + x) * 8;
            r.h[o8] = *reinterpret_cast<const uint4 *>(rh + off);
            r.l[o8] = *reinterpret_cast<const uint4 *>(rl + off);
        } else {
            r.h[o8] = make_uint4(0, 0, 0, 0);
            r.l[o8] = make_uint4(0, 0, 0, 0);
        }
    }
}

// One 32-column block [c0, c0+32) of the accumulator.  r1: prefetched first residual (or null);
// a second residual (q.res2_*, rare) is read in place.
__device__ __forceinline__ void epilogue_store_block(const ConvPtrs &q, const ConvParams &p, uint32_t taddr, int c0, int b,
                                                     int y, int x, bool ok, int o_base, const float *sbias,
                                                     float out_scale, float res_scale, float so,
                                                     const ResRegs *r1 = nullptr) {
                    {
                        uint32_t reg[32];
                        if (p.N - c0 >= 32) {
                            tmem_ld_32x32(taddr + c0, reg);
                        } else {   // 16-column tail
                            asm volatile(
                                "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                                : "=r"(reg[0]), "=r"(reg[1]), "=r"(reg[2]), "=r"(reg[3]), "=r"(reg[4]), "=r"(reg[5]),
                                  "=r"(reg[6]), "=r"(reg[7]), "=r"(reg[8]), "=r"(reg[9]), "=r"(reg[10]), "=r"(reg[11]),
                                  "=r"(reg[12]), "=r"(reg[13]), "=r"(reg[14]), "=r"(reg[15])
                                : "r"(taddr + c0)
                                : "memory");
                        }
                        if (p.stacked) {
                            // B operand was [W_hi | W_lo] stacked along N: columns c and N + c hold the
                            // (x * w_hi) and (x * w_lo) partial sums of output channel c
                            uint32_t reg2[32];
                            if (p.N - c0 >= 32) {
                                tmem_ld_32x32(taddr + p.N + c0, reg2);
                            } else {
                                asm volatile(
                                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                                    : "=r"(reg2[0]), "=r"(reg2[1]), "=r"(reg2[2]), "=r"(reg2[3]), "=r"(reg2[4]), "=r"(reg2[5]),
                                      "=r"(reg2[6]), "=r"(reg2[7]), "=r"(reg2[8]), "=r"(reg2[9]), "=r"(reg2[10]), "=r"(reg2[11]),
                                      "=r"(reg2[12]), "=r"(reg2[13]), "=r"(reg2[14]), "=r"(reg2[15])
                                    : "r"(taddr + p.N + c0)
                                    : "memory");
                            }
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                reg[j] = __float_as_uint(__uint_as_float(reg[j]) + __uint_as_float(reg2[j]));
                        } else {
                            tmem_ld_wait();
                        }
                        if (!ok || (p.dbg & 1)) return;
                        const int ncol = min(32, p.N - c0);
                        float v[32];
    #pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float a = fmaf(__uint_as_float(reg[j]), out_scale, sbias[c0 + j]);
                            if (p.act == 1) a = fmaxf(a, 0.f);
                            else if (p.act == 2) a = a > 0.f ? a : a * 0.1f;
                            v[j] = (j < ncol && o_base + c0 + j < p.Cout) ? a : 0.f;
                        }
                        if (q.out_f32 && p.f32_mode == 2) {
    #pragma unroll
                            for (int o8 = 0; o8 < 4; ++o8) {
                                const int oct = (o_base + c0) / 8 + o8;
                                if (o8 * 8 >= ncol || oct >= p.C8out) break;
                                float4 *dst = reinterpret_cast<float4 *>(
                                    q.out_f32 + ((((size_t)b * p.C8out + oct) * p.H + y) * p.W + x) * 8);
                                dst[0] = make_float4(v[o8 * 8], v[o8 * 8 + 1], v[o8 * 8 + 2], v[o8 * 8 + 3]);
                                dst[1] = make_float4(v[o8 * 8 + 4], v[o8 * 8 + 5], v[o8 * 8 + 6], v[o8 * 8 + 7]);
                            }
                        } else if (q.out_f32) {
                            const long long base = b * p.os_b + y * p.os_y + x * p.os_x + (long long)(o_base + c0) * p.os_c;
    #pragma unroll
                            for (int j4 = 0; j4 < 8; ++j4) {
                                const int j = j4 * 4, o = o_base + c0 + j;
                                if (p.f32_mode == 1 && j + 3 < ncol && o + 3 < p.Cout) {
                                    float4 val = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                                    if (q.add_f32) {
                                        const float4 ad = *reinterpret_cast<const float4 *>(q.add_f32 + base + j);
                                        val.x += ad.x; val.y += ad.y; val.z += ad.z; val.w += ad.w;
                                    }
                                    *reinterpret_cast<float4 *>(q.out_f32 + base + j) = val;
                                } else {
    #pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        if (j + k < ncol && o + k < p.Cout) {
                                            const long long oi = base + (j + k) * p.os_c;
                                            q.out_f32[oi] = q.add_f32 ? v[j + k] + q.add_f32[oi] : v[j + k];
                                        }
                                }
                            }
                        }
                        if (q.out_hi && p.ps == 0) {
    #pragma unroll
                            for (int o8 = 0; o8 < 4; ++o8) {
                                const int oct = (o_base + c0) / 8 + o8;
                                if (o8 * 8 >= ncol || oct >= p.C8out) break;
                                const size_t off = ((((size_t)b * p.C8out + oct) * p.H + y) * p.W + x) * 8;
                                float r8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                                if (r1) {
                                    const __half *hh = reinterpret_cast<const __half *>(&r1->h[o8]);
                                    const __half *ll = reinterpret_cast<const __half *>(&r1->l[o8]);
    #pragma unroll
                                    for (int j = 0; j < 8; ++j) r8[j] = (__half2float(hh[j]) + __half2float(ll[j])) * res_scale;
                                }
                                if (q.res2_hi) {
                                    const uint4 rh = *reinterpret_cast<const uint4 *>(q.res2_hi + off);
                                    const uint4 rl = *reinterpret_cast<const uint4 *>(q.res2_lo + off);
                                    const __half *hh = reinterpret_cast<const __half *>(&rh);
                                    const __half *ll = reinterpret_cast<const __half *>(&rl);
    #pragma unroll
                                    for (int j = 0; j < 8; ++j) r8[j] += (__half2float(hh[j]) + __half2float(ll[j])) * res_scale;
                                }
                                __align__(16) __half h8[8];
                                __align__(16) __half l8[8];
    #pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const int o = o_base + c0 + o8 * 8 + j;
                                    const float vs = (o < p.Cout ? v[o8 * 8 + j] + r8[j] : 0.f) * so;
                                    const __half hh = __float2half_rn(vs);
                                    h8[j] = hh;
                                    l8[j] = __float2half_rn(vs - __half2float(hh));
                                }
                                if (p.cs) {
                                    __stcs(reinterpret_cast<uint4 *>(q.out_hi + off), *reinterpret_cast<const uint4 *>(h8));
                                    __stcs(reinterpret_cast<uint4 *>(q.out_lo + off), *reinterpret_cast<const uint4 *>(l8));
                                } else {
                                    *reinterpret_cast<uint4 *>(q.out_hi + off) = *reinterpret_cast<const uint4 *>(h8);
                                    *reinterpret_cast<uint4 *>(q.out_lo + off) = *reinterpret_cast<const uint4 *>(l8);
                                }
                            }
                        }
                        if (q.out_hi && p.ps == 2) {
                            // PixelShuffle(2): conv channel o = 4c + 2i + j  ->  out[c][2y+i][2x+j]
                            // the 32 columns at c0 hold c = (o_base+c0)/4 .. +7  = exactly one output octet
                            const int oct = (o_base + c0) / 32;
                            if (oct < p.C8out && ncol == 32) {
    #pragma unroll
                                for (int ij = 0; ij < 4; ++ij) {
                                    const size_t off = ((((size_t)b * p.C8out + oct) * p.Hout + 2 * y + (ij >> 1)) * p.Wout +
                                                        2 * x + (ij & 1)) * 8;
                                    __align__(16) __half h8[8];
                                    __align__(16) __half l8[8];
    #pragma unroll
                                    for (int c = 0; c < 8; ++c) {
                                        const float vs = v[c * 4 + ij] * so;
                                        const __half hh = __float2half_rn(vs);
                                        h8[c] = hh;
                                        l8[c] = __float2half_rn(vs - __half2float(hh));
                                    }
                                    *reinterpret_cast<uint4 *>(q.out_hi + off) = *reinterpret_cast<const uint4 *>(h8);
                                    *reinterpret_cast<uint4 *>(q.out_lo + off) = *reinterpret_cast<const uint4 *>(l8);
                                }
                            }
                        }
                    }
}

// ---------------------------------------------------------------------------------------------------------------
// Specialised epilogue of one full 32-column block (4 complete channel octets): the generic function above spends
// ~930 SASS instructions per (warp, tile) on per-element bounds predicates, activation selects and address
// arithmetic — with it the 8 epilogue warps were busy 87 % of the time while the tensor pipe sat at 57 % (ncu,
// profiles/r02_conv_epilogue_bound.md).  Here everything that is uniform per launch is a template parameter and the
// block is straight-line code: ~10 instructions per output element.
//   ACT 0 none / 1 ReLU / 2 LeakyReLU(0.1);  RES: one PSA residual (prefetched);  F32OCT: octet-planar fp32 output
//   instead of PSA;  STACKED: accumulator columns [c, N + c) hold the x*W_hi and x*W_lo partial sums.
// Preconditions (checked by the caller, warp-uniform): c0 + 32 <= N, o_base + c0 + 32 <= Cout, ps == 0, no res2,
// exactly one output kind.  A PSA output scale 2^sa_out is folded in by the caller: `sbias`, `out_scale` and `res_scale`
// arrive pre-multiplied by it (ReLU / LeakyReLU commute with a positive scale), so it costs nothing per element.
// ---------------------------------------------------------------------------------------------------------------
// Branch-free residual prefetch of a FULL 32-column block (all four octets exist): the loads land directly in their
// final registers.  (The generic prefetch_residual guards every octet with a branch; ptxas then copies each loaded
// value inside its basic block, i.e. waits for one load round trip per octet.)
__device__ __forceinline__ void prefetch_residual_full(const __half *rh, const __half *rl, const ConvParams &p, int b, int y,
                                                       int x, bool ok, int o_base, int c0, ResRegs &r) {
    const size_t plane = (size_t)p.H * p.W * 8;
    const size_t off0 = ((size_t)b * p.C8out + (o_base + c0) / 8) * plane + (ok ? ((size_t)y * p.W + x) * 8 : (size_t)0);
#pragma unroll
    for (int o8 = 0; o8 < 4; ++o8) {
        r.h[o8] = __ldg(reinterpret_cast<const uint4 *>(rh + off0 + o8 * plane));
        r.l[o8] = __ldg(reinterpret_cast<const uint4 *>(rl + off0 + o8 * plane));
    }
}

// The residual octets are fetched BEFORE the thread blocks on the accumulator so their HBM / L2 latency overlaps the
// MMAs.  Their fp16 -> fp32 conversion does not depend on the accumulator, so the compiler hoists it above the wait —
// which puts four serialised load round trips in front of every tile (measured: +160 us on a 64->64 @640x640 launch).
// This no-op "modifies" the registers after the wait, so nothing can consume them earlier.
__device__ __forceinline__ void pin_residual(ResRegs &r) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        asm volatile("" : "+r"(r.h[i].x), "+r"(r.h[i].y), "+r"(r.h[i].z), "+r"(r.h[i].w));
        asm volatile("" : "+r"(r.l[i].x), "+r"(r.l[i].y), "+r"(r.l[i].z), "+r"(r.l[i].w));
    }
}
__device__ __forceinline__ float4 lds_f4(const float *p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
    return v;
}

template <int ACT, bool RES, bool F32OCT, bool STACKED>
__device__ __forceinline__ void epilogue_fast_block(const ConvPtrs &q, const ConvParams &p, uint32_t taddr, int c0, int b,
                                                    int y, int x, bool ok, int o_base, const float *sbias,
                                                    float out_scale, float res_scale, ResRegs *r1) {
    uint32_t reg[32];
    tmem_ld_32x32(taddr + c0, reg);
    if (STACKED) {
        uint32_t reg2[32];
        tmem_ld_32x32(taddr + p.N + c0, reg2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) reg[j] = __float_as_uint(__uint_as_float(reg[j]) + __uint_as_float(reg2[j]));
    } else {
        tmem_ld_wait();
    }
    if (RES) pin_residual(*r1);
    if (!ok) return;
    const size_t plane = (size_t)p.H * p.W * 8;                       // elements per channel octet
    const size_t off0 = ((size_t)b * p.C8out + (o_base + c0) / 8) * plane + ((size_t)y * p.W + x) * 8;
#pragma unroll
    for (int o8 = 0; o8 < 4; ++o8) {
        const float4 b0 = lds_f4(sbias + c0 + o8 * 8), b1 = lds_f4(sbias + c0 + o8 * 8 + 4);
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float a = fmaf(__uint_as_float(reg[o8 * 8 + j]), out_scale, bb[j]);
            if (ACT == 1) a = fmaxf(a, 0.f);
            else if (ACT == 2) a = fmaxf(a, 0.1f * a);                  // LeakyReLU(0.1): max(a, 0.1 a)
            v[j] = a;
        }
        if (RES) {
            const __half2 *hh = reinterpret_cast<const __half2 *>(&r1->h[o8]);
            const __half2 *ll = reinterpret_cast<const __half2 *>(&r1->l[o8]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 fh = __half22float2(hh[j]), fl = __half22float2(ll[j]);
                v[2 * j] = fmaf(fh.x + fl.x, res_scale, v[2 * j]);
                v[2 * j + 1] = fmaf(fh.y + fl.y, res_scale, v[2 * j + 1]);
            }
        }
        const size_t off = off0 + o8 * plane;
        if (F32OCT) {
            float4 *dst = reinterpret_cast<float4 *>(q.out_f32 + off);
            dst[0] = make_float4(v[0], v[1], v[2], v[3]);
            dst[1] = make_float4(v[4], v[5], v[6], v[7]);
        } else {
            __align__(16) __half2 h4[4];
            __align__(16) __half2 l4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const __half2 hq = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
                const float2 hf = __half22float2(hq);
                h4[j] = hq;
                l4[j] = __floats2half2_rn(v[2 * j] - hf.x, v[2 * j + 1] - hf.y);
            }
            if (p.cs) {
                __stcs(reinterpret_cast<uint4 *>(q.out_hi + off), *reinterpret_cast<const uint4 *>(h4));
                __stcs(reinterpret_cast<uint4 *>(q.out_lo + off), *reinterpret_cast<const uint4 *>(l4));
            } else {
                *reinterpret_cast<uint4 *>(q.out_hi + off) = *reinterpret_cast<const uint4 *>(h4);
                *reinterpret_cast<uint4 *>(q.out_lo + off) = *reinterpret_cast<const uint4 *>(l4);
            }
        }
    }
}

// warp-uniform: can this launch's full blocks take the specialised epilogue?
__device__ __forceinline__ bool epilogue_fast_ok(const ConvPtrs &q, const ConvParams &p) {
    const bool psa_only = q.out_hi && !q.out_f32;
    const bool oct_only = q.out_f32 && p.f32_mode == 2 && !q.out_hi;
    return p.ps == 0 && !q.res2_hi && !(p.dbg & 1) && (psa_only || (oct_only && !q.res_hi && p.sa_out == 0));
}

template <bool STACKED>
__device__ __forceinline__ void epilogue_fast_dispatch(const ConvPtrs &q, const ConvParams &p, uint32_t taddr, int c0, int b,
                                                       int y, int x, bool ok, int o_base, const float *sbias,
                                                       float out_scale, float res_scale, ResRegs *r1) {
    const bool oct = q.out_f32 != nullptr;
    if (oct) {
        if (p.act == 0) epilogue_fast_block<0, false, true, STACKED>(q, p, taddr, c0, b, y, x, ok, o_base, sbias, out_scale, res_scale, r1);
        else if (p.act == 1) epilogue_fast_block<1, false, true, STACKED>(q, p, taddr, c0, b, y, x, ok, o_base, sbias, out_scale, res_scale, r1);
        else epilogue_fast_block<2, false, true, STACKED>(q, p, taddr, c0, b, y, x, ok, o_base, sbias, out_scale, res_scale, r1);
    } else if (r1) {
        if (p.act == 0) epilogue_fast_block<0, true, false, STACKED>(q, p, taddr, c0, b, y, x, ok, o_base, sbias, out_scale, res_scale, r1);
        else if (p.act == 1) epilogue_fast_block<1, true, false, STACKED>(q, p, taddr, c0, b, y, x, ok, o_base, sbias, out_scale, res_scale, r1);
        else epilogue_fast_block<2, true, false, STACKED>(q, p, taddr, c0, b, y, x, ok, o_base, sbias, out_scale, res_scale, r1);
    } else {
        if (p.act == 0) epilogue_fast_block<0, false, false, STACKED>(q, p, taddr, c0, b, y, x, ok, o_base, sbias, out_scale, res_scale, r1);
        else if (p.act == 1) epilogue_fast_block<1, false, false, STACKED>(q, p, taddr, c0, b, y, x, ok, o_base, sbias, out_scale, res_scale, r1);
        else epilogue_fast_block<2, false, false, STACKED>(q, p, taddr, c0, b, y, x, ok, o_base, sbias, out_scale, res_scale, r1);
    }
}

}  // namespace c2m
