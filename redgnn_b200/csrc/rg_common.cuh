// Shared device helpers for libredgnn_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "redgnn_b200.h"

#define RG_FULL_MASK 0xffffffffu

#define RG_LAUNCH_CHECK()                                        \
    do {                                                         \
        cudaError_t e__ = cudaGetLastError();                    \
        if (e__ != cudaSuccess) return RG_ERR_CUDA_BASE - (int)e__; \
    } while (0)

#define RG_CUDA_CALL(x)                                          \
    do {                                                         \
        cudaError_t e__ = (x);                                   \
        if (e__ != cudaSuccess) return RG_ERR_CUDA_BASE - (int)e__; \
    } while (0)

static inline size_t rg_align256(size_t x) { return (x + 255) & ~(size_t)255; }
static inline int64_t rg_cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- frontier geometry -------------------------------------------------------------------
__host__ __device__ static inline int rg_words_query(int n_query) { return (n_query + 31) >> 5; }  // Wn
__host__ __device__ static inline int rg_words_ent(int n_ent) { return (n_ent + 31) >> 5; }        // We

// rank(b, e) inside the sorted node list, given the dictionary word of e
__device__ __forceinline__ uint32_t rg_rank(uint2 d, int e) {
    return d.y + __popc(d.x & ((1u << (e & 31)) - 1u));
}

// position of the k-th (0-based) LOWEST set bit of x; requires k < popc(x)
__device__ __forceinline__ int rg_select_low(uint32_t x, int k) {
    int q = 0;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        int c = __popc((x >> q) & ((1u << s) - 1u));
        if (k >= c) {
            k -= c;
            q += s;
        }
    }
    return q;
}

// position of the k-th (0-based) HIGHEST set bit of x; requires k < popc(x)
__device__ __forceinline__ int rg_select_high(uint32_t x, int k) {
    return 31 - rg_select_low(__brev(x), k);
}

// ---- block-wide exclusive scan -------------------------------------------------------------
// smem must hold BLOCK/32 + 1 elements.  Returns the exclusive prefix of v; total = block sum.
template <int BLOCK, typename T>
__device__ __forceinline__ T rg_block_exclusive_scan(T v, T *smem, T &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = BLOCK / 32;
    T incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(RG_FULL_MASK, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) smem[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        T w = (lane < NW) ? smem[lane] : (T)0;
        T wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            T t = __shfl_up_sync(RG_FULL_MASK, wi, o);
            if (lane >= o) wi += t;
        }
        if (lane < NW) smem[lane] = wi - w;
        if (lane == NW - 1) smem[NW] = wi;
    }
    __syncthreads();
    T excl = smem[warp] + (incl - v);
    total = smem[NW];
    __syncthreads();
    return excl;
}

template <int BLOCK, typename T>
__device__ __forceinline__ T rg_block_sum(T v, T *smem) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = BLOCK / 32;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(RG_FULL_MASK, v, o);
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    T r = (T)0;
    if (warp == 0) {
        r = (lane < NW) ? smem[lane] : (T)0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(RG_FULL_MASK, r, o);
    }
    __syncthreads();
    return r;  // valid in warp 0
}

// workspace layout shared by rg_frontier_* and rg_edges_emit
struct RgWorkspace {
    unsigned long long *dict_blockprefix;  // [nbd + 1]
    uint32_t *dict_blocksum;               // [nbd]
    unsigned long long *fact_blockprefix;  // [nbf + 1]
    uint32_t *fact_blocksum;               // [nbf]
    size_t dict_bytes, total_bytes;
};

#define RG_TILE 1024  // items per block in the tiled scans (256 threads x 4)

static inline RgWorkspace rg_carve(void *ws, int32_t n_query, int32_t n_ent, int64_t n_fact) {
    RgWorkspace w;
    int64_t nbd = rg_cdiv((int64_t)n_query * rg_words_ent(n_ent), RG_TILE);
    int64_t nbf = rg_cdiv(n_fact, RG_TILE);
    char *p = (char *)ws;
    size_t off = 0;
    w.dict_blockprefix = (unsigned long long *)(p + off);
    off += rg_align256(8 * (size_t)(nbd + 1));
    w.dict_blocksum = (uint32_t *)(p + off);
    off += rg_align256(4 * (size_t)(nbd + 1));
    w.dict_bytes = off;
    w.fact_blockprefix = (unsigned long long *)(p + off);
    off += rg_align256(8 * (size_t)(nbf + 1));
    w.fact_blocksum = (uint32_t *)(p + off);
    off += rg_align256(4 * (size_t)(nbf + 1));
    w.total_bytes = off;
    return w;
}
