// Internal interfaces between the correlation kernels (prep -> search -> rescore).
//
// Data layout in HBM (per feature map of one call, all images b of the batch):
//   p32   [B][HW][Cp]        fp32, pixel-major ("NHWC"), Cp = C rounded up to 8, pad = 0.
//                            Exact values the search is defined on (optionally L2-normalised).
//   hi,lo [B][Cp/8][HW][8]   fp16 split operands of x * 2^sexp:  hi = rn(x*S), lo = rn(x*S - hi).
//                            One 16-byte element per (channel-octet, pixel): a (W*8)-wide row of
//                            one image row is contiguous, so a TMA box of (cols*8, rows, octets)
//                            lands in shared memory as [octet][pixel][8 halfs] = the tcgen05
//                            K-major no-swizzle core-matrix layout with rows 16 B apart.
//   ss    [B][HW]            fp32 per-pixel sum of squares of p32 (for the patch norms).
//   rinv  [B][NR]            fp32 1 / (sqrt(sum_{patch} ss) + 1e-5)   (1 when !is_norm)
//   part  [B][nchunk][NQ]    Candidate: approximate top-4 of each query over one Ref chunk (nchunk <= 32).
//   ovf   [B*NQ]             (query, chunk bit mask) of the (query, chunk) pairs whose top-2 list may be incomplete:
//                            the exhaustive pass re-scans those chunks exactly (see corr_aux.cu).
#pragma once
#include "c2m_common.cuh"

namespace c2m {

constexpr int CORR_TOPK = 4;      // candidates kept per (query, Ref chunk) by the searches
struct __align__(16) Candidate {
    float v[CORR_TOPK];           // approximate scores, best first
    int i[CORR_TOPK];             // Ref patch indices (-1 = empty slot)
    float dropped;                // largest approximate score of the chunk that is NOT in the list
    float pad[3];
};

struct CorrGeom {
    int B, C, Cp;
    int h, w, hr, wr;        // map sizes
    int patch, s_in, s_ref;
    int gh, gw, rh, rw;      // patch grids: input (gh x gw), Ref (rh x rw)
    int NQ, NR;
};

constexpr int CORR_MAX_CHUNKS = 32;

// how a chunk index maps to Ref patches (for the exhaustive fallback)
struct CorrChunkGeom {
    int mode;                // 0: contiguous Ref index ranges of `per` patches (generic search); 1: runs of `per` tiles
    int per;
    int n_rt, rt_x, tile_rows, tile_cols;    // mode 1: tile grid over the Ref patch grid
};

struct CorrOverflow {
    int query;               // b * NQ + q
    unsigned chunks;         // bit c: chunk c must be re-scanned
};

struct CorrWorkspace {
    float *p32_in, *p32_ref;
    __half *hi_in, *lo_in, *hi_ref, *lo_ref;
    float *ss_in, *ss_ref;
    float *rinv;
    Candidate *part;
    CorrOverflow *ovf;       // [B*NQ] overflow list
    unsigned long long *best;   // [B*NQ] packed (exact score, index) of the best candidate so far (best_key)
    float *qnorm;            // [B*NQ] ||P_query|| (exact, deterministic)
    unsigned *amax_bits;     // [2]
    int *sexp;               // [2] scale exponents for (in, ref)
    unsigned *ovf_count;     // [1] entries in ovf
    unsigned *max_pn_bits;   // [1] max over Ref patches of ||P_ref|| (float bits) — window scale when !is_norm
    int nchunk;
    size_t total_bytes;
};

// Sorted top-4 insert with a record of what falls off the list: `dropped` ends up as the largest score the chunk
// produced that is not in v[] (scores that never made it, and entries pushed out).  Strict '>' — ties beyond the
// list count as dropped, and the rescoring pass re-scans a chunk exhaustively whenever `dropped` reaches the
// window, so neither ties nor the list length can lose the true argmax.
__device__ __forceinline__ void cand_push(float s, int r, float (&v)[CORR_TOPK], int (&i)[CORR_TOPK], float &dropped) {
    if (!(s > v[3])) {
        dropped = fmaxf(dropped, s);
        return;
    }
    dropped = fmaxf(dropped, v[3]);
    if (s > v[2]) {
        v[3] = v[2]; i[3] = i[2];
        if (s > v[1]) {
            v[2] = v[1]; i[2] = i[1];
            if (s > v[0]) {
                v[1] = v[0]; i[1] = i[0]; v[0] = s; i[0] = r;
            } else {
                v[1] = s; i[1] = r;
            }
        } else {
            v[2] = s; i[2] = r;
        }
    } else {
        v[3] = s; i[3] = r;
    }
}
__device__ __forceinline__ void cand_init(float (&v)[CORR_TOPK], int (&i)[CORR_TOPK]) {
#pragma unroll
    for (int k = 0; k < CORR_TOPK; ++k) { v[k] = -INFINITY; i[k] = 0x7fffffff; }
}
__device__ __forceinline__ Candidate cand_pack(const float (&v)[CORR_TOPK], const int (&i)[CORR_TOPK], float dropped) {
    Candidate c;
#pragma unroll
    for (int k = 0; k < CORR_TOPK; ++k) { c.v[k] = v[k]; c.i[k] = i[k] == 0x7fffffff ? -1 : i[k]; }
    c.dropped = dropped;
    c.pad[0] = c.pad[1] = c.pad[2] = 0.f;
    return c;
}
__device__ __forceinline__ bool cand_better(float s, int r, float v, int i) {   // exact scores: lower index wins ties
    return s > v || (s == v && r < i);
}

// (score, index) packed so that an unsigned 64-bit max is the lexicographic "better": larger score, then LOWER index
__device__ __forceinline__ unsigned long long best_key(float s, int r) {
    const unsigned b = __float_as_uint(s);
    const unsigned ord = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ((unsigned long long)ord << 32) | (unsigned long long)(0xffffffffu - (unsigned)r);
}
__device__ __forceinline__ void best_unkey(unsigned long long k, float &s, int &r) {
    const unsigned ord = (unsigned)(k >> 32);
    s = __uint_as_float((ord & 0x80000000u) ? (ord & 0x7fffffffu) : ~ord);
    r = (int)(0xffffffffu - (unsigned)(k & 0xffffffffu));
}

int corr_prep_launch(const float *x, int B, int C, int Cp, int HW, int l2norm, int map_slot,
                     const CorrWorkspace &ws, float *p32, __half *hi, __half *lo, float *ss,
                     cudaStream_t st);
int corr_rinv_launch(const CorrGeom &g, const CorrWorkspace &ws, int is_norm, cudaStream_t st);
int corr_search_generic_launch(const CorrGeom &g, const CorrWorkspace &ws, cudaStream_t st);
int corr_search_umma_launch(const CorrGeom &g, const CorrWorkspace &ws, cudaStream_t st);
bool corr_umma_supported(const CorrGeom &g);
int corr_umma_pick_nchunk(const CorrGeom &g, int sms);
void corr_umma_chunk_geom(const CorrGeom &g, int nchunk, CorrChunkGeom &cg);
void corr_generic_chunk_geom(const CorrGeom &g, int nchunk, CorrChunkGeom &cg);
// window_coef: rigorous bound on 2 x |approximate - exact score| per unit of ||P_query|| (see corr_aux.cu)
int corr_rescore_launch(const CorrGeom &g, const CorrWorkspace &ws, const CorrChunkGeom &cg, float window_coef,
                        int is_norm, int norm_input, int64_t *idx, float *val, cudaStream_t st);

}  // namespace c2m
