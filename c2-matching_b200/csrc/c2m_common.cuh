// Shared helpers for libc2m_sm100: error plumbing, launch accounting and the sm_100a PTX
// wrappers (mbarrier, TMA, tcgen05/TMEM) used by the kernels.  sm_100a only.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/c2m_sm100.h"

namespace c2m {

// ---------------------------------------------------------------------------- host side
void set_error(const char *fmt, ...);
void count_launch(unsigned n = 1);

// Measurement hook (include/c2m_sm100.h c2m_profile_*): kernel classes timed with CUDA events on the
// launching stream while profiling is enabled.
enum ProfKernel { PROF_CORR_SEARCH = 0, PROF_CONV3X3 = 1, PROF_DCN = 2, PROF_NKERNELS = 3 };
bool prof_enabled();
void *prof_begin(int kernel, double flops, double bytes, cudaStream_t st);   // null when disabled
void prof_end(void *handle, cudaStream_t st);

#define C2M_CHECK_ARG(cond, ...)                 \
    do {                                         \
        if (!(cond)) {                           \
            c2m::set_error(__VA_ARGS__);         \
            return C2M_ERR_INVALID;              \
        }                                        \
    } while (0)

#define C2M_CUDA(call)                                                                        \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            c2m::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                           __LINE__);                                                         \
            return C2M_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

#define C2M_LAUNCH_CHECK(name)                                                         \
    do {                                                                               \
        cudaError_t e__ = cudaGetLastError();                                          \
        if (e__ != cudaSuccess) {                                                      \
            c2m::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__)); \
            return C2M_ERR_CUDA;                                                       \
        }                                                                              \
        c2m::count_launch();                                                           \
    } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
__host__ __device__ static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------- device side
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug (lost arrive, register-pool deadlock, ...) must end in a trap that the
// host sees as a launch failure, not in a kernel that spins until the GPU lease is killed.  One try_wait
// probe suspends for up to ~1 us, so the limit below is tens of seconds; -DC2M_MBAR_SPIN_LIMIT=0 removes
// the counter.
#ifndef C2M_MBAR_SPIN_LIMIT
#define C2M_MBAR_SPIN_LIMIT (1u << 26)
#endif
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
#if C2M_MBAR_SPIN_LIMIT
    uint32_t n = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++n > C2M_MBAR_SPIN_LIMIT) __trap();   // no printf here: it would put a stack frame into every role loop
    }
#else
    while (!mbar_try_wait(bar, parity)) {
    }
#endif
}
// (A hinted variant — try_wait with a suspend-time hint, which compiles to NANOSLEEP.SYNCS — was measured
// for the long epilogue / producer waits: no change in step time or power, so the plain probe loop stays.)

// ---- TMA (cp.async.bulk.tensor), 4-D tiled load into this CTA's shared memory
__device__ __forceinline__ void tma_prefetch_desc(const void *tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst, const void *tmap, uint64_t *bar, int c0, int c1,
                                            int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// ---- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(smem_result)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], fp16/bf16 inputs, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Single-thread issue with split descriptor words (inside `if (elect_one())`).
__device__ __forceinline__ void umma_f16_1(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                           uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Warp-uniform issue: called by ALL 32 lanes of the issuing warp under uniform control flow; one lane, elected inside
// the asm, issues.  With `if (lane == 0) { ... umma_f16(...) }` ptxas wraps every MMA in an ELECT / R2UR.BROADCAST /
// BRA.U.ANY loop (~11 SASS instructions, ~80 cycles per MMA measured) — longer than a 128x64x16 MMA occupies the pipe.
// Descriptors are passed as (low word, high word): the low word is `base + constant` in 32-bit arithmetic, which ptxas
// keeps on the uniform datapath; a 64-bit descriptor value is computed in vector registers and costs R2UR moves.
__device__ __forceinline__ void umma_f16_w(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                           uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_f16_w(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
    umma_f16_w(d_tmem, (uint32_t)a_desc, (uint32_t)(a_desc >> 32), (uint32_t)b_desc, (uint32_t)(b_desc >> 32), idesc, accumulate);
}
__device__ __forceinline__ void umma_commit_w(uint64_t *bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(smem_u32(bar))
        : "memory");
}
// mbarrier arrives once all tcgen05 ops issued so far by this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns -> 32 registers per thread (thread i <-> lane base+i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
          "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
          "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
          "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major, SWIZZLE_NONE ("interleave"): core matrix = 8 rows
// x 16 bytes, rows 16 B apart; `lbo` = byte distance between the two K core matrices of one
// UMMA_K=16 step, `sbo` = byte distance between consecutive 8-row groups along M/N.
// Field layout: cute/arch/mma_sm100_desc.hpp (SmemDescriptor): start[0,14) lbo[16,30) sbo[32,46)
// version[46,48)=1 layout_type[61,64)=0.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
           (1ull << 46);
}
// Same descriptor advanced by `byte_off` (multiple of 16; the 14-bit start-address field cannot
// carry: shared memory is < 2^18 bytes) — one 32-bit add on the low word.
__device__ __forceinline__ uint64_t umma_desc_advance(uint64_t desc, uint32_t byte_off) {
    const uint32_t lo = (uint32_t)desc + (byte_off >> 4);
    return (desc & 0xFFFFFFFF00000000ull) | lo;
}
// Instruction descriptor for kind::f16: D fp32, A/B fp16 (0) or bf16 (1), both K-major.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N, int ab_fmt) {
    return (1u << 4) | ((uint32_t)ab_fmt << 7) | ((uint32_t)ab_fmt << 10) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

// ---- CTA pairs (cta_group::2): cluster of two CTAs on one TPC sharing one UMMA of M = 256.  Only the leader
// (cluster rank 0) issues MMAs; each CTA feeds its own 128 A rows and HALF of the B rows from its own shared memory
// (same offsets in both CTAs) and receives its own 128 accumulator lanes in its own TMEM.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` (a shared::cta address of THIS CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t *smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma2_f16_w(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit2_w(uint64_t *bar) {
    const uint16_t mask = 3;
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
        ::"r"(smem_u32(bar)), "h"(mask)
        : "memory");
}
// arrive (once all tcgen05 ops issued so far by this thread have completed) on the barrier at this shared-memory
// offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit2(uint64_t *bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
#endif  // __CUDACC__

}  // namespace c2m
