// Fused edge kernels: the B200 replacement of GNNLayer.forward's per-edge work
// (reference Static/transductive/models.py:23-39: gathers, attention, alpha*(hs+hr), scatter-sum)
// and of its autograd.
//
// One warp owns one segment (= one output node in forward, one input node in backward).  32
// candidate slots are probed per step (one lane each: adjacency entry + dictionary word), the
// active ones are compacted with a ballot and then processed 8 at a time by groups of 4 lanes;
// every lane of a group moves D/16 float4 (128-bit) pieces of the D-float rows.  The per-segment
// sum is a fixed function of the slot order (no atomics) => bit-reproducible.
// Segments longer than RG_HEAVY_CHUNK slots are cut into chunks via a device queue; chunk partials
// are added back in chunk order by a fix-up kernel.
#include <algorithm>

#include "rg_common.cuh"

namespace {

#ifndef RG_FWD_MINBLOCKS
#define RG_FWD_MINBLOCKS 5  // CTAs per SM the non-persistent forward is compiled for (48 registers)
#endif
constexpr int kWarpsPerBlock = 8;
constexpr int kBlock = kWarpsPerBlock * 32;

__device__ __forceinline__ float4 ldg4(const float4 *p) { return __ldg(p); }
// ex2.approx / rcp.approx form (5 instructions instead of ~28 for the IEEE division + expf; relative
// error ~2^-21, far inside the 1e-4 parity bound).  Forward and backward use the same function.
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

struct Slot {
    int peer;  // row of the peer node (valid when active)
    int rel;
    bool active;
};

// probe candidate slot `slot` of the segment of query q.  complete_base >= 0: the peer frontier of
// this query holds every entity, so rank = base + entity and no dictionary probe is needed.
template <bool IMPLICIT>
__device__ __forceinline__ Slot probe_slot(const rg_segments &S, const uint2 *drow, int complete_base, int slot,
                                           bool valid) {
    Slot s;
    s.peer = 0;
    s.rel = 0;
    s.active = false;
    if (valid) {
        int2 pr = __ldg(reinterpret_cast<const int2 *>(S.adj) + slot);
        s.rel = pr.y;
        if (IMPLICIT) {
            if (complete_base >= 0) {
                s.active = true;
                s.peer = complete_base + pr.x;
            } else {
                uint2 d = __ldg(drow + (pr.x >> 5));
                uint32_t bit = 1u << (pr.x & 31);
                s.active = (d.x & bit) != 0u;
                s.peer = (int)(d.y + __popc(d.x & (bit - 1u)));
            }
        } else {
            s.peer = pr.x;
            s.active = true;
        }
    }
    return s;
}

// warp-uniform: rank base of query q if its peer frontier is complete, else -1
template <bool IMPLICIT>
__device__ __forceinline__ int complete_base_of(const rg_segments &S, int q) {
    if (!IMPLICIT || S.peer_qinfo == nullptr) return -1;
    const int2 qi = __ldg(reinterpret_cast<const int2 *>(S.peer_qinfo) + q);
    return qi.y == S.n_ent ? qi.x : -1;
}

struct SegRange {
    int q, lo, hi;
};

template <bool IMPLICIT>
__device__ __forceinline__ SegRange seg_range(const rg_segments &S, int64_t seg) {
    SegRange r;
    r.q = __ldg(S.seg_query + seg);
    if (IMPLICIT) {
        int e = __ldg(S.seg_ent + seg);
        r.lo = __ldg(S.ent_ptr + e);
        r.hi = __ldg(S.ent_ptr + e + 1);
    } else {
        r.lo = __ldg(S.seg_ptr + seg);
        r.hi = __ldg(S.seg_ptr + seg + 1);
    }
    return r;
}

// queue the chunks 1.. of a heavy segment (chunk 0 is done by the owner warp); warp-collective
// `own` = slots the owner warp keeps for itself (defaults to one chunk)
template <bool PUBLISH = false>
__device__ __forceinline__ void enqueue_heavy(const rg_heavy &H, int64_t seg, int len, int lane,
                                              int chunk = RG_HEAVY_CHUNK, int own = 0) {
    if (own <= 0) own = chunk;
    const int nch = (len - own + chunk - 1) / chunk;
    int base = 0, slot = 0;
    if (lane == 0) {
        base = atomicAdd(&H.counters[0], nch);
        slot = atomicAdd(&H.counters[1], 1);
    }
    base = __shfl_sync(RG_FULL_MASK, base, 0);
    slot = __shfl_sync(RG_FULL_MASK, slot, 0);
    if (base + nch > H.max_chunks || slot >= H.max_nodes) {
        if (lane == 0) atomicExch(&H.counters[2], 1);
        return;
    }
    if (lane == 0) {
        H.node_seg[slot] = (int)seg;
        H.node_base[slot] = base;
        H.node_n[slot] = nch;
    }
    for (int c = lane; c < nch; c += 32) H.chunk_seg[base + c] = (int)seg;
    if (PUBLISH) {
        // chunk_idx (>= 1) doubles as the "published" flag the in-kernel drain of the persistent forward
        // polls: chunk_seg first, then a release store of chunk_idx
        __threadfence();
        __syncwarp();
        for (int c = lane; c < nch; c += 32)
            asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(H.chunk_idx + base + c), "r"(c + 1) : "memory");
    } else {
        for (int c = lane; c < nch; c += 32) H.chunk_idx[base + c] = c + 1;
    }
}

__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// The persistent forward drains the heavy-chunk queue ITSELF: a warp that has run out of segments takes
// chunks (counters[3]) as soon as their entries are published, until every warp of the grid has finished
// producing (counters[4] == total_warps) and the queue is empty.  The chunk work thereby overlaps the tail
// of the segment work instead of running as a second, latency-bound kernel behind it.
// The drain is OPPORTUNISTIC, never required for correctness: a finished chunk is marked by negating its
// chunk_idx, and k_edge_fwd_chunks runs after the kernel over whatever is still positive (normally nothing:
// a few microseconds).  A warp therefore never waits unboundedly -- if the CTAs of this launch are not all
// co-resident (another stream or process holding SMs), producers it would wait for may not be running;
// after kDrainPolls fruitless polls it simply leaves.  Returns the chunk to process or -1.  Warp-collective.
constexpr int kDrainPolls = 512;   // x up to ~4 us of back-off each: ~2 ms of fruitless waiting at most
__device__ __forceinline__ int heavy_take(const rg_heavy &H, int total_warps, int lane) {
    int c = -1;
    if (lane == 0) {
        // claim the next index once (one atomic per take: a compare-and-swap that never overshoots was measured
        // -- thousands of idle warps retrying it serialise on one L2 line, 8.5 -> 400 ms per step)
        const int t = atomicAdd(&H.counters[3], 1);
        unsigned backoff = 256;   // ns; doubles up to ~4 us: idle warps must not hammer the counters' L2 lines
        for (int polls = 0; polls < kDrainPolls; ++polls) {
            if (ld_acquire(&H.counters[2])) break;               // queue overflow (reported to the caller)
            int reserved = min(ld_acquire(&H.counters[0]), H.max_chunks);
            if (t < reserved) {                                   // exists; its entry is published at once or very soon
                int spins = 0;
                while (ld_acquire(&H.chunk_idx[t]) == 0 && ++spins < 4096) __nanosleep(64);
                if (spins < 4096) c = t;
                break;
            }
            if (ld_acquire(&H.counters[4]) >= total_warps) {      // no producer left: final look at the queue
                reserved = min(ld_acquire(&H.counters[0]), H.max_chunks);
                if (t >= reserved) break;
                continue;
            }
            __nanosleep(backoff);
            if (backoff < 4096) backoff <<= 1;
        }
        // (giving up leaves index t -- should a producer still create it -- to the clean-up kernel)
    }
    return __shfl_sync(RG_FULL_MASK, c, 0);
}

// ------------------------------------------------------------------------------------------
// forward: acc = sum over active slots in [lo, hi) of alpha * (hidden[peer] + rela[rel])
// On return lanes 0..3 hold the reduced row pieces (float4 index v*4 + lane).
// ------------------------------------------------------------------------------------------
template <int D, bool HAS_HIDDEN, bool IMPLICIT, bool SMEM_TAB = false>
__device__ __forceinline__ void fwd_range(const rg_segments &S, int q, int lo, int hi,
                                          const float *__restrict__ hidden, const float *__restrict__ as8,
                                          const float *__restrict__ rela, const float *__restrict__ ar8,
                                          const float *__restrict__ aq8, const float *__restrict__ w8,
                                          float b_alpha, float4 (&acc)[D / 16], const float *s_rela = nullptr,
                                          const float *s_ar8 = nullptr) {
    constexpr int NV = D / 16;
    const int lane = threadIdx.x & 31, grp = lane >> 2, ql = lane & 3;
    const float2 aq2 = __ldg(reinterpret_cast<const float2 *>(aq8 + (size_t)q * 8) + ql);
    const float2 w2 = __ldg(reinterpret_cast<const float2 *>(w8) + ql);
    const uint2 *drow = IMPLICIT ? reinterpret_cast<const uint2 *>(S.peer_dict) + (size_t)q * rg_words_ent(S.n_ent)
                                 : nullptr;
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int cbase = complete_base_of<IMPLICIT>(S, q);

    for (int base = lo; base < hi; base += 32) {
        const int slot = base + lane;
        Slot s = probe_slot<IMPLICIT>(S, drow, cbase, slot, slot < hi);
        const unsigned m = __ballot_sync(RG_FULL_MASK, s.active);
        const int cnt = __popc(m);
        for (int it = 0; it < cnt; it += 8) {
            const int k = it + grp;
            const bool on = k < cnt;
            const int src = on ? ((m == RG_FULL_MASK) ? k : rg_select_low(m, k)) : 0;
            const int p = __shfl_sync(RG_FULL_MASK, s.peer, src);
            const int r = __shfl_sync(RG_FULL_MASK, s.rel, src);
            float4 x[NV];
            float part = 0.f;
            if (on) {
                // relation row / attention row: shared-memory copies when staged (LDS), else global
                const float4 *rp = SMEM_TAB ? reinterpret_cast<const float4 *>(s_rela + r * D)
                                            : reinterpret_cast<const float4 *>(rela + (size_t)r * D);
                float2 z = SMEM_TAB ? reinterpret_cast<const float2 *>(s_ar8 + r * 8)[ql]
                                    : __ldg(reinterpret_cast<const float2 *>(ar8 + (size_t)r * 8) + ql);
                if (HAS_HIDDEN) {
                    const float4 *hp = reinterpret_cast<const float4 *>(hidden + (size_t)p * D);
                    float4 h[NV];
#pragma unroll
                    for (int v = 0; v < NV; ++v) h[v] = ldg4(hp + v * 4 + ql);
                    float2 a = __ldg(reinterpret_cast<const float2 *>(as8 + (size_t)p * 8) + ql);
#pragma unroll
                    for (int v = 0; v < NV; ++v) {
                        float4 t = SMEM_TAB ? rp[v * 4 + ql] : ldg4(rp + v * 4 + ql);
                        x[v] = make_float4(h[v].x + t.x, h[v].y + t.y, h[v].z + t.z, h[v].w + t.w);
                    }
                    z.x += a.x;
                    z.y += a.y;
                } else {
#pragma unroll
                    for (int v = 0; v < NV; ++v) x[v] = SMEM_TAB ? rp[v * 4 + ql] : ldg4(rp + v * 4 + ql);
                }
                z.x += aq2.x;
                z.y += aq2.y;
                part = w2.x * fmaxf(z.x, 0.f) + w2.y * fmaxf(z.y, 0.f);
            } else {
#pragma unroll
                for (int v = 0; v < NV; ++v) x[v] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            part += __shfl_xor_sync(RG_FULL_MASK, part, 1);
            part += __shfl_xor_sync(RG_FULL_MASK, part, 2);
            const float alpha = on ? sigmoidf_(part + b_alpha) : 0.f;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                acc[v].x = fmaf(alpha, x[v].x, acc[v].x);
                acc[v].y = fmaf(alpha, x[v].y, acc[v].y);
                acc[v].z = fmaf(alpha, x[v].z, acc[v].z);
                acc[v].w = fmaf(alpha, x[v].w, acc[v].w);
            }
        }
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            acc[v].x += __shfl_xor_sync(RG_FULL_MASK, acc[v].x, o);
            acc[v].y += __shfl_xor_sync(RG_FULL_MASK, acc[v].y, o);
            acc[v].z += __shfl_xor_sync(RG_FULL_MASK, acc[v].z, o);
            acc[v].w += __shfl_xor_sync(RG_FULL_MASK, acc[v].w, o);
        }
    }
}

template <int D>
__device__ __forceinline__ void store_row(float *dst, const float4 (&acc)[D / 16], int lane) {
    if (lane < 4) {
        float4 *o = reinterpret_cast<float4 *>(dst);
#pragma unroll
        for (int v = 0; v < D / 16; ++v) o[v * 4 + lane] = acc[v];
    }
}

// ------------------------------------------------------------------------------------------
// forward, short segments: the 8 four-lane groups of a warp own 8 consecutive segments and walk
// their slots one by one (no ballot, no cross-group reduction): a power-law graph's median segment
// has a handful of slots, and a whole warp per segment would leave 7 of 8 groups idle.  Segments
// longer than kShortSeg slots are left to the warp-wide fwd_range.  The per-segment sum is again a
// fixed function of the slot order.  Used by the non-persistent kernel (relation tables too large
// for shared memory, small layers, explicit segments); measured on the 1M-entity power-law KG:
// 21.3 -> 15.0 ms per forward.  Inside the 64-register persistent kernel it costs more than it saves
// (FB15k-237 / YAGO shapes: +9 %), so that kernel keeps one segment per warp.
// ------------------------------------------------------------------------------------------
#ifndef RG_SHORT_SEG
#define RG_SHORT_SEG 12
#endif
constexpr int kShortSeg = RG_SHORT_SEG;

template <int D, bool HAS_HIDDEN, bool IMPLICIT, bool SMEM_TAB>
__device__ __forceinline__ void fwd_group(const rg_segments &S, bool mine, int q, int lo, int len,
                                          const float *__restrict__ hidden, const float *__restrict__ as8,
                                          const float *__restrict__ rela, const float *__restrict__ ar8,
                                          const float *__restrict__ aq8, const float *__restrict__ w8, float b_alpha,
                                          float4 (&acc)[D / 16], const float *s_rela, const float *s_ar8) {
    constexpr int NV = D / 16;
    const int ql = threadIdx.x & 3;
    const float2 w2 = __ldg(reinterpret_cast<const float2 *>(w8) + ql);
    float2 aq2 = make_float2(0.f, 0.f);
    const uint2 *drow = nullptr;
    int cbase = -1;
    if (mine) {
        aq2 = __ldg(reinterpret_cast<const float2 *>(aq8 + (size_t)q * 8) + ql);
        if (IMPLICIT) {
            drow = reinterpret_cast<const uint2 *>(S.peer_dict) + (size_t)q * rg_words_ent(S.n_ent);
            cbase = complete_base_of<IMPLICIT>(S, q);
        }
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    int max_len = mine ? len : 0;
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) max_len = max(max_len, __shfl_xor_sync(RG_FULL_MASK, max_len, o));
    for (int i = 0; i < max_len; ++i) {
        const bool valid = mine && i < len;
        Slot sl = probe_slot<IMPLICIT>(S, drow, cbase, lo + i, valid);
        const bool on = sl.active;
        float4 x[NV];
        float part = 0.f;
        if (on) {
            const int p = sl.peer, r = sl.rel;
            const float4 *rp = SMEM_TAB ? reinterpret_cast<const float4 *>(s_rela + r * D)
                                        : reinterpret_cast<const float4 *>(rela + (size_t)r * D);
            float2 z = SMEM_TAB ? reinterpret_cast<const float2 *>(s_ar8 + r * 8)[ql]
                                : __ldg(reinterpret_cast<const float2 *>(ar8 + (size_t)r * 8) + ql);
            if (HAS_HIDDEN) {
                const float4 *hp = reinterpret_cast<const float4 *>(hidden + (size_t)p * D);
                float4 h[NV];
#pragma unroll
                for (int v = 0; v < NV; ++v) h[v] = ldg4(hp + v * 4 + ql);
                float2 a = __ldg(reinterpret_cast<const float2 *>(as8 + (size_t)p * 8) + ql);
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    float4 t = SMEM_TAB ? rp[v * 4 + ql] : ldg4(rp + v * 4 + ql);
                    x[v] = make_float4(h[v].x + t.x, h[v].y + t.y, h[v].z + t.z, h[v].w + t.w);
                }
                z.x += a.x;
                z.y += a.y;
            } else {
#pragma unroll
                for (int v = 0; v < NV; ++v) x[v] = SMEM_TAB ? rp[v * 4 + ql] : ldg4(rp + v * 4 + ql);
            }
            z.x += aq2.x;
            z.y += aq2.y;
            part = w2.x * fmaxf(z.x, 0.f) + w2.y * fmaxf(z.y, 0.f);
        } else {
#pragma unroll
            for (int v = 0; v < NV; ++v) x[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        part += __shfl_xor_sync(RG_FULL_MASK, part, 1);
        part += __shfl_xor_sync(RG_FULL_MASK, part, 2);
        const float alpha = on ? sigmoidf_(part + b_alpha) : 0.f;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            acc[v].x = fmaf(alpha, x[v].x, acc[v].x);
            acc[v].y = fmaf(alpha, x[v].y, acc[v].y);
            acc[v].z = fmaf(alpha, x[v].z, acc[v].z);
            acc[v].w = fmaf(alpha, x[v].w, acc[v].w);
        }
    }
}

// one warp, 8 consecutive segments starting at seg0: short ones group-wise, then the long ones warp-wide
template <int D, bool HAS_HIDDEN, bool IMPLICIT, bool SMEM_TAB>
__device__ __forceinline__ void fwd_block8(const rg_segments &S, int64_t seg0, int64_t n_true,
                                           const float *__restrict__ hidden, const float *__restrict__ as8,
                                           const float *__restrict__ rela, const float *__restrict__ ar8,
                                           const float *__restrict__ aq8, const float *__restrict__ w8, float b_alpha,
                                           float *__restrict__ agg, const rg_heavy &H, int has_heavy,
                                           const float *s_rela, const float *s_ar8) {
    const int lane = threadIdx.x & 31, grp = lane >> 2, ql = lane & 3;
    const int64_t seg = seg0 + grp;
    const bool valid = seg < n_true;
    SegRange r;
    r.q = 0;
    r.lo = 0;
    r.hi = 0;
    if (valid) r = seg_range<IMPLICIT>(S, seg);
    const int len = r.hi - r.lo;
    const bool is_short = valid && len <= kShortSeg;
    float4 acc[D / 16];
    if (__any_sync(RG_FULL_MASK, is_short)) {
        fwd_group<D, HAS_HIDDEN, IMPLICIT, SMEM_TAB>(S, is_short, r.q, r.lo, len, hidden, as8, rela, ar8, aq8, w8,
                                                     b_alpha, acc, s_rela, s_ar8);
        if (is_short) {
            float4 *o = reinterpret_cast<float4 *>(agg + (size_t)seg * D);
#pragma unroll
            for (int v = 0; v < D / 16; ++v) o[v * 4 + ql] = acc[v];
        }
    }
    unsigned long_mask = __ballot_sync(RG_FULL_MASK, valid && !is_short && ql == 0);
    while (long_mask) {  // warp-uniform
        const int src = __ffs(long_mask) - 1;
        long_mask &= long_mask - 1;
        const int q = __shfl_sync(RG_FULL_MASK, r.q, src);
        const int lo = __shfl_sync(RG_FULL_MASK, r.lo, src);
        int hi = __shfl_sync(RG_FULL_MASK, r.hi, src);
        const int64_t sg = seg0 + (src >> 2);
        if (has_heavy && hi - lo > RG_HEAVY_CHUNK) {
            enqueue_heavy(H, sg, hi - lo, lane, RG_HEAVY_SUB, RG_HEAVY_CHUNK);
            hi = lo + RG_HEAVY_CHUNK;
        }
        fwd_range<D, HAS_HIDDEN, IMPLICIT, SMEM_TAB>(S, q, lo, hi, hidden, as8, rela, ar8, aq8, w8, b_alpha, acc,
                                                     s_rela, s_ar8);
        store_row<D>(agg + (size_t)sg * D, acc, lane);
    }
}

template <int D, bool HAS_HIDDEN, bool IMPLICIT, bool BLOCK8>
__global__ void __launch_bounds__(kBlock, RG_FWD_MINBLOCKS) k_edge_fwd(rg_segments S, const float *__restrict__ hidden,
                                                     const float *__restrict__ as8, const float *__restrict__ rela,
                                                     const float *__restrict__ ar8, const float *__restrict__ aq8,
                                                     const float *__restrict__ w8, const float *__restrict__ b_alpha,
                                                     float *__restrict__ agg, rg_heavy H, int has_heavy) {
    const int64_t n_true = S.n_seg_dev ? *S.n_seg_dev : S.n_seg;
    if (BLOCK8) {  // many segments: 8 per warp, the short ones group-wise
        const int64_t seg0 = ((int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5)) * 8;
        if (seg0 >= n_true) return;
        fwd_block8<D, HAS_HIDDEN, IMPLICIT, false>(S, seg0, n_true, hidden, as8, rela, ar8, aq8, w8, __ldg(b_alpha), agg,
                                                   H, has_heavy, nullptr, nullptr);
        return;
    }
    const int lane = threadIdx.x & 31;
    const int64_t seg = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (seg >= n_true) return;
    SegRange r = seg_range<IMPLICIT>(S, seg);
    int hi = r.hi;
    if (has_heavy && r.hi - r.lo > RG_HEAVY_CHUNK) {  // without a queue the owner warp does it all
        enqueue_heavy(H, seg, r.hi - r.lo, lane, RG_HEAVY_SUB, RG_HEAVY_CHUNK);
        hi = r.lo + RG_HEAVY_CHUNK;
    }
    float4 acc[D / 16];
    fwd_range<D, HAS_HIDDEN, IMPLICIT>(S, r.q, r.lo, hi, hidden, as8, rela, ar8, aq8, w8, __ldg(b_alpha), acc);
    store_row<D>(agg + (size_t)seg * D, acc, lane);
}

// Stage the two relation tables in shared memory with the TMA bulk-copy engine (cp.async.bulk, 1-D):
// one elected thread issues two copies that complete on an mbarrier; no thread spends issue slots
// on the ~100 KB transfer.  Block-collective (contains a __syncthreads).
__device__ __forceinline__ void stage_tables(unsigned long long *tab_bar, float *s_a, const float *g_a,
                                             uint32_t bytes_a, float *s_b, const float *g_b, uint32_t bytes_b) {
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(tab_bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes_a + bytes_b)
                     : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(s_a)),
                     "l"(g_a), "r"(bytes_a), "r"(bar)
                     : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(s_b)),
                     "l"(g_b), "r"(bytes_b), "r"(bar)
                     : "memory");
    }
    __syncthreads();  // the barrier is initialised before anyone polls it
    uint32_t ok = 0;
    while (!ok)
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n"
            "selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(bar)
            : "memory");
}

// Persistent variant for the implicit (model) path: 16 warps per CTA, CTAs sized to the SM count,
// the relation tables (rela [rows][D], ar8 [rows][8]) staged once per CTA in shared memory so the
// per-edge relation row comes from LDS (half the L1 wavefronts of the global path), segments
// handed out round-robin (warp w of CTA c takes c*16+w, then += grid*16).
#ifndef RG_PWARPS
#define RG_PWARPS 16
#endif
constexpr int kPWarps = RG_PWARPS;
#ifndef RG_PWARPS_BWD
#define RG_PWARPS_BWD 12
#endif
constexpr int kPWarpsB = RG_PWARPS_BWD;  // persistent backward: 12 warps x 2 CTAs = 85 registers, no spills (-7 % vs 16)

template <int D, bool HAS_HIDDEN>
__global__ void __launch_bounds__(kPWarps * 32, 2) k_edge_fwd_p(rg_segments S, const float *__restrict__ hidden,
                                                               const float *__restrict__ as8,
                                                               const float *__restrict__ rela,
                                                               const float *__restrict__ ar8,
                                                               const float *__restrict__ aq8,
                                                               const float *__restrict__ w8,
                                                               const float *__restrict__ b_alpha,
                                                               float *__restrict__ agg, rg_heavy H, int has_heavy) {
    extern __shared__ __align__(128) float s_tab[];
    __shared__ __align__(8) unsigned long long tab_bar;
    const int rows = S.n_table_rows;
    float *s_rela = s_tab, *s_ar8 = s_tab + (size_t)rows * D;
    stage_tables(&tab_bar, s_rela, rela, (uint32_t)rows * D * 4, s_ar8, ar8, (uint32_t)rows * 32);
    const int lane = threadIdx.x & 31;
    const int64_t n_true = S.n_seg_dev ? *S.n_seg_dev : S.n_seg;
    const float ba = __ldg(b_alpha);
    // work loop: this warp's segments (static round robin), then chunks of heavy segments from the queue
    int64_t seg = (int64_t)blockIdx.x * kPWarps + (threadIdx.x >> 5);
    const int64_t stride = (int64_t)gridDim.x * kPWarps;
    bool draining = false;
    int done_c = 0;
    for (;;) {
        int q, lo, hi;
        float *dst;
        if (!draining) {
            if (seg < n_true) {
                SegRange r = seg_range<true>(S, seg);
                q = r.q, lo = r.lo, hi = r.hi;
                if (has_heavy && r.hi - r.lo > RG_HEAVY_CHUNK) {
                    enqueue_heavy<true>(H, seg, r.hi - r.lo, lane, RG_HEAVY_SUB, RG_HEAVY_CHUNK);
                    hi = r.lo + RG_HEAVY_CHUNK;
                }
                dst = agg + (size_t)seg * D;
                seg += stride;
            } else {
                if (!has_heavy) break;
                draining = true;
                if (lane == 0) {
                    __threadfence();
                    atomicAdd(&H.counters[4], 1);
                }
                continue;
            }
        } else {
            const int c = heavy_take(H, (int)stride, lane);
            if (c < 0) break;
            done_c = c;
            SegRange r = seg_range<true>(S, (int64_t)H.chunk_seg[c]);
            q = r.q;
            lo = r.lo + RG_HEAVY_CHUNK + (H.chunk_idx[c] - 1) * RG_HEAVY_SUB;   // chunk_idx is 1-based
            hi = min(r.hi, lo + RG_HEAVY_SUB);
            dst = H.partial + (size_t)c * D;
        }
        float4 acc[D / 16];
        fwd_range<D, HAS_HIDDEN, true, true>(S, q, lo, hi, hidden, as8, rela, ar8, aq8, w8, ba, acc, s_rela, s_ar8);
        store_row<D>(dst, acc, lane);
        if (draining && lane == 0) H.chunk_idx[done_c] = -H.chunk_idx[done_c];   // finished: the clean-up kernel skips it
    }
}

template <int D, bool HAS_HIDDEN, bool IMPLICIT>
__global__ void __launch_bounds__(kBlock) k_edge_fwd_chunks(rg_segments S, const float *__restrict__ hidden,
                                                            const float *__restrict__ as8,
                                                            const float *__restrict__ rela,
                                                            const float *__restrict__ ar8,
                                                            const float *__restrict__ aq8,
                                                            const float *__restrict__ w8,
                                                            const float *__restrict__ b_alpha, rg_heavy H) {
    const int lane = threadIdx.x & 31;
    if (H.counters[2]) return;  // queue overflow: reported to the caller, nothing to trust
    const int n_chunks = min(H.counters[0], H.max_chunks);
    const int stride = gridDim.x * kWarpsPerBlock;
    for (int c = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); c < n_chunks; c += stride) {
        const int idx = H.chunk_idx[c];
        if (idx <= 0) continue;   // already reduced by the persistent kernel's own drain (warp-uniform)
        const int64_t seg = H.chunk_seg[c];
        SegRange r = seg_range<IMPLICIT>(S, seg);
        const int lo = r.lo + RG_HEAVY_CHUNK + (idx - 1) * RG_HEAVY_SUB;
        const int hi = min(r.hi, lo + RG_HEAVY_SUB);
        float4 acc[D / 16];
        fwd_range<D, HAS_HIDDEN, IMPLICIT>(S, r.q, lo, hi, hidden, as8, rela, ar8, aq8, w8, __ldg(b_alpha), acc);
        store_row<D>(H.partial + (size_t)c * D, acc, lane);
    }
}

// rows[seg] += sum_c partial[base + c], c ascending; ROW floats per row, one warp per heavy node
__global__ void __launch_bounds__(kBlock) k_heavy_fixup(rg_heavy H, int row_floats, float *rows_a, int a_floats,
                                                        float *rows_b, int b_floats) {
    const int lane = threadIdx.x & 31;
    if (H.counters[2]) return;
    const int n_nodes = min(H.counters[1], H.max_nodes);
    const int stride = gridDim.x * kWarpsPerBlock;
    for (int i = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); i < n_nodes; i += stride) {
        const size_t seg = (size_t)H.node_seg[i];
        const int base = H.node_base[i], n = H.node_n[i];
        for (int col = lane; col < row_floats; col += 32) {
            float *dst = nullptr;
            if (col < a_floats) {
                if (rows_a) dst = rows_a + seg * a_floats + col;
            } else if (rows_b) {
                dst = rows_b + seg * b_floats + (col - a_floats);
            }
            if (!dst) continue;
            float v = *dst;
            for (int c = 0; c < n; ++c) v += H.partial[(size_t)(base + c) * row_floats + col];
            *dst = v;
        }
    }
}

// ------------------------------------------------------------------------------------------
// backward, grouped by the input node.  Per-segment results (after the group reduction, valid in
// lanes 0..3): G = sum alpha*g_agg[peer]; small = {Z2 (g_as8 pair), WZ2, GL}.
// ------------------------------------------------------------------------------------------
struct BwdSmall {
    float2 z, wz;
    float gl;
};

template <int D, bool HAS_HIDDEN, bool IMPLICIT, bool SMEM_TAB = false>
__device__ __forceinline__ void bwd_range(const rg_segments &S, int64_t seg, int q, int lo, int hi,
                                          const float *__restrict__ hidden, const float *__restrict__ as8,
                                          const float *__restrict__ rela, const float *__restrict__ ar8,
                                          const float *__restrict__ aq8, const float *__restrict__ w8,
                                          float b_alpha, const float *__restrict__ g_agg, float *g_rela,
                                          float *g_ar8, float4 (&G)[D / 16], BwdSmall &sm,
                                          const float *s_rela = nullptr, const float *s_ar8 = nullptr) {
    constexpr int NV = D / 16;
    const int lane = threadIdx.x & 31, grp = lane >> 2, ql = lane & 3;
    const float2 w2 = __ldg(reinterpret_cast<const float2 *>(w8) + ql);
    float2 zbase = __ldg(reinterpret_cast<const float2 *>(aq8 + (size_t)q * 8) + ql);
    float4 hs[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) hs[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (HAS_HIDDEN) {
        const float4 *hp = reinterpret_cast<const float4 *>(hidden + (size_t)seg * D);
#pragma unroll
        for (int v = 0; v < NV; ++v) hs[v] = ldg4(hp + v * 4 + ql);
        float2 a = __ldg(reinterpret_cast<const float2 *>(as8 + (size_t)seg * 8) + ql);
        zbase.x += a.x;
        zbase.y += a.y;
    }
    const uint2 *drow = IMPLICIT ? reinterpret_cast<const uint2 *>(S.peer_dict) + (size_t)q * rg_words_ent(S.n_ent)
                                 : nullptr;
#pragma unroll
    for (int v = 0; v < NV; ++v) G[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    sm.z = make_float2(0.f, 0.f);
    sm.wz = make_float2(0.f, 0.f);
    sm.gl = 0.f;
    const int cbase = complete_base_of<IMPLICIT>(S, q);

    for (int base = lo; base < hi; base += 32) {
        const int slot = base + lane;
        Slot s = probe_slot<IMPLICIT>(S, drow, cbase, slot, slot < hi);
        const unsigned m = __ballot_sync(RG_FULL_MASK, s.active);
        const int cnt = __popc(m);
        for (int it = 0; it < cnt; it += 8) {
            const int k = it + grp;
            const bool on = k < cnt;
            const int src = on ? ((m == RG_FULL_MASK) ? k : rg_select_low(m, k)) : 0;
            const int p = __shfl_sync(RG_FULL_MASK, s.peer, src);
            const int r = __shfl_sync(RG_FULL_MASK, s.rel, src);
            float4 g[NV];
            float2 z = zbase;
            float dot = 0.f, part = 0.f;
            if (on) {
                const float4 *gp = reinterpret_cast<const float4 *>(g_agg + (size_t)p * D);
                const float4 *rp = SMEM_TAB ? reinterpret_cast<const float4 *>(s_rela + r * D)
                                            : reinterpret_cast<const float4 *>(rela + (size_t)r * D);
#pragma unroll
                for (int v = 0; v < NV; ++v) g[v] = ldg4(gp + v * 4 + ql);
                float2 a = SMEM_TAB ? reinterpret_cast<const float2 *>(s_ar8 + r * 8)[ql]
                                    : __ldg(reinterpret_cast<const float2 *>(ar8 + (size_t)r * 8) + ql);
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    float4 t = SMEM_TAB ? rp[v * 4 + ql] : ldg4(rp + v * 4 + ql);
                    dot = fmaf(g[v].x, hs[v].x + t.x, dot);
                    dot = fmaf(g[v].y, hs[v].y + t.y, dot);
                    dot = fmaf(g[v].z, hs[v].z + t.z, dot);
                    dot = fmaf(g[v].w, hs[v].w + t.w, dot);
                }
                z.x += a.x;
                z.y += a.y;
                part = w2.x * fmaxf(z.x, 0.f) + w2.y * fmaxf(z.y, 0.f);
            } else {
#pragma unroll
                for (int v = 0; v < NV; ++v) g[v] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            part += __shfl_xor_sync(RG_FULL_MASK, part, 1);
            part += __shfl_xor_sync(RG_FULL_MASK, part, 2);
            dot += __shfl_xor_sync(RG_FULL_MASK, dot, 1);
            dot += __shfl_xor_sync(RG_FULL_MASK, dot, 2);
            if (on) {
                const float alpha = sigmoidf_(part + b_alpha);
                const float gl = dot * alpha * (1.f - alpha);
                float2 gz = make_float2(z.x > 0.f ? gl * w2.x : 0.f, z.y > 0.f ? gl * w2.y : 0.f);
                sm.z.x += gz.x;
                sm.z.y += gz.y;
                sm.wz.x = fmaf(gl, fmaxf(z.x, 0.f), sm.wz.x);
                sm.wz.y = fmaf(gl, fmaxf(z.y, 0.f), sm.wz.y);
                sm.gl += gl;
                float4 *gr = reinterpret_cast<float4 *>(g_rela + (size_t)r * D);
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    float4 ag = make_float4(alpha * g[v].x, alpha * g[v].y, alpha * g[v].z, alpha * g[v].w);
                    G[v].x += ag.x;
                    G[v].y += ag.y;
                    G[v].z += ag.z;
                    G[v].w += ag.w;
                    atomicAdd(gr + v * 4 + ql, ag);
                }
                atomicAdd(reinterpret_cast<float2 *>(g_ar8 + (size_t)r * 8) + ql, gz);
            }
        }
    }
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            G[v].x += __shfl_xor_sync(RG_FULL_MASK, G[v].x, o);
            G[v].y += __shfl_xor_sync(RG_FULL_MASK, G[v].y, o);
            G[v].z += __shfl_xor_sync(RG_FULL_MASK, G[v].z, o);
            G[v].w += __shfl_xor_sync(RG_FULL_MASK, G[v].w, o);
        }
        sm.z.x += __shfl_xor_sync(RG_FULL_MASK, sm.z.x, o);
        sm.z.y += __shfl_xor_sync(RG_FULL_MASK, sm.z.y, o);
        sm.wz.x += __shfl_xor_sync(RG_FULL_MASK, sm.wz.x, o);
        sm.wz.y += __shfl_xor_sync(RG_FULL_MASK, sm.wz.y, o);
        sm.gl += __shfl_xor_sync(RG_FULL_MASK, sm.gl, o);
    }
}

// backward, short segments: same group-per-segment scheme as fwd_group (see there)
template <int D, bool HAS_HIDDEN, bool IMPLICIT>
__device__ __forceinline__ void bwd_group(const rg_segments &S, bool mine, int64_t seg, int q, int lo, int len,
                                          const float *__restrict__ hidden, const float *__restrict__ as8,
                                          const float *__restrict__ rela, const float *__restrict__ ar8,
                                          const float *__restrict__ aq8, const float *__restrict__ w8, float b_alpha,
                                          const float *__restrict__ g_agg, float *g_rela, float *g_ar8,
                                          float4 (&G)[D / 16], BwdSmall &sm) {
    constexpr int NV = D / 16;
    const int ql = threadIdx.x & 3;
    const float2 w2 = __ldg(reinterpret_cast<const float2 *>(w8) + ql);
    float2 zbase = make_float2(0.f, 0.f);
    float4 hs[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) hs[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint2 *drow = nullptr;
    int cbase = -1;
    if (mine) {
        zbase = __ldg(reinterpret_cast<const float2 *>(aq8 + (size_t)q * 8) + ql);
        if (HAS_HIDDEN) {
            const float4 *hp = reinterpret_cast<const float4 *>(hidden + (size_t)seg * D);
#pragma unroll
            for (int v = 0; v < NV; ++v) hs[v] = ldg4(hp + v * 4 + ql);
            float2 a = __ldg(reinterpret_cast<const float2 *>(as8 + (size_t)seg * 8) + ql);
            zbase.x += a.x;
            zbase.y += a.y;
        }
        if (IMPLICIT) {
            drow = reinterpret_cast<const uint2 *>(S.peer_dict) + (size_t)q * rg_words_ent(S.n_ent);
            cbase = complete_base_of<IMPLICIT>(S, q);
        }
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) G[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    sm.z = make_float2(0.f, 0.f);
    sm.wz = make_float2(0.f, 0.f);
    sm.gl = 0.f;
    int max_len = mine ? len : 0;
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) max_len = max(max_len, __shfl_xor_sync(RG_FULL_MASK, max_len, o));
    for (int i = 0; i < max_len; ++i) {
        Slot sl = probe_slot<IMPLICIT>(S, drow, cbase, lo + i, mine && i < len);
        const bool on = sl.active;
        const int r = sl.rel;
        float4 g[NV];
        float2 z = zbase;
        float dot = 0.f, part = 0.f;
        if (on) {
            const float4 *gp = reinterpret_cast<const float4 *>(g_agg + (size_t)sl.peer * D);
            const float4 *rp = reinterpret_cast<const float4 *>(rela + (size_t)r * D);
#pragma unroll
            for (int v = 0; v < NV; ++v) g[v] = ldg4(gp + v * 4 + ql);
            float2 a = __ldg(reinterpret_cast<const float2 *>(ar8 + (size_t)r * 8) + ql);
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                float4 t = ldg4(rp + v * 4 + ql);
                dot = fmaf(g[v].x, hs[v].x + t.x, dot);
                dot = fmaf(g[v].y, hs[v].y + t.y, dot);
                dot = fmaf(g[v].z, hs[v].z + t.z, dot);
                dot = fmaf(g[v].w, hs[v].w + t.w, dot);
            }
            z.x += a.x;
            z.y += a.y;
            part = w2.x * fmaxf(z.x, 0.f) + w2.y * fmaxf(z.y, 0.f);
        }
        part += __shfl_xor_sync(RG_FULL_MASK, part, 1);
        part += __shfl_xor_sync(RG_FULL_MASK, part, 2);
        dot += __shfl_xor_sync(RG_FULL_MASK, dot, 1);
        dot += __shfl_xor_sync(RG_FULL_MASK, dot, 2);
        if (on) {
            const float alpha = sigmoidf_(part + b_alpha);
            const float gl = dot * alpha * (1.f - alpha);
            float2 gz = make_float2(z.x > 0.f ? gl * w2.x : 0.f, z.y > 0.f ? gl * w2.y : 0.f);
            sm.z.x += gz.x;
            sm.z.y += gz.y;
            sm.wz.x = fmaf(gl, fmaxf(z.x, 0.f), sm.wz.x);
            sm.wz.y = fmaf(gl, fmaxf(z.y, 0.f), sm.wz.y);
            sm.gl += gl;
            float4 *gr = reinterpret_cast<float4 *>(g_rela + (size_t)r * D);
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                float4 ag = make_float4(alpha * g[v].x, alpha * g[v].y, alpha * g[v].z, alpha * g[v].w);
                G[v].x += ag.x;
                G[v].y += ag.y;
                G[v].z += ag.z;
                G[v].w += ag.w;
                atomicAdd(gr + v * 4 + ql, ag);
            }
            atomicAdd(reinterpret_cast<float2 *>(g_ar8 + (size_t)r * 8) + ql, gz);
        }
    }
}

// The relation-gradient accumulators may be replicated (`copies` > 1, [copies][rows][D] and
// [copies][rows][8]): a CTA adds into copy blockIdx % copies, which spreads the fp32 reductions of
// the few thousand hot sectors over `copies` times as many L2 lines; the caller sums the copies.
template <int D>
__device__ __forceinline__ void select_copy(float *&g_rela, float *&g_ar8, int copies, int rows) {
    if (copies > 1) {
        const size_t c = blockIdx.x % (unsigned)copies;
        g_rela += c * (size_t)rows * D;
        g_ar8 += c * (size_t)rows * 8;
    }
}

// node_small row layout: [0..7] g_as8, [8..15] sum g_l*relu(z), [16] sum g_l, [17..23] zero
__device__ __forceinline__ void store_small(float *dst, const BwdSmall &sm, int lane) {
    if (lane < 4) {
        float2 *o = reinterpret_cast<float2 *>(dst);
        o[lane] = sm.z;
        o[4 + lane] = sm.wz;
        o[8 + lane] = make_float2(lane == 0 ? sm.gl : 0.f, 0.f);
    }
}

template <int D, bool HAS_HIDDEN, bool IMPLICIT, bool BLOCK8>
__global__ void __launch_bounds__(kBlock) k_edge_bwd(rg_segments S, const float *__restrict__ hidden,
                                                     const float *__restrict__ as8, const float *__restrict__ rela,
                                                     const float *__restrict__ ar8, const float *__restrict__ aq8,
                                                     const float *__restrict__ w8, const float *__restrict__ b_alpha,
                                                     const float *__restrict__ g_agg, float *g_hidden,
                                                     float *node_small, float *g_rela, float *g_ar8, int copies,
                                                     rg_heavy H, int has_heavy, int own0) {
    const int lane = threadIdx.x & 31;
    select_copy<D>(g_rela, g_ar8, copies, S.n_table_rows);
    const int64_t n_true = S.n_seg_dev ? *S.n_seg_dev : S.n_seg;
    const float ba = __ldg(b_alpha);
    if (!BLOCK8) {  // one warp per segment
        const int64_t seg = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
        if (seg >= n_true) return;
        SegRange r = seg_range<IMPLICIT>(S, seg);
        int hi = r.hi;
        if (has_heavy && r.hi - r.lo > own0) {  // without a queue the owner warp does it all
            enqueue_heavy(H, seg, r.hi - r.lo, lane, RG_HEAVY_SUB_BWD, own0);
            hi = r.lo + own0;
        }
        float4 G[D / 16];
        BwdSmall sm;
        bwd_range<D, HAS_HIDDEN, IMPLICIT>(S, seg, r.q, r.lo, hi, hidden, as8, rela, ar8, aq8, w8, ba, g_agg, g_rela,
                                           g_ar8, G, sm);
        if (g_hidden) store_row<D>(g_hidden + (size_t)seg * D, G, lane);
        store_small(node_small + (size_t)seg * 24, sm, lane);
        return;
    }
    const int grp = lane >> 2, ql = lane & 3;
    const int64_t seg0 = ((int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5)) * 8;
    if (seg0 >= n_true) return;
    const int64_t seg = seg0 + grp;
    const bool valid = seg < n_true;
    SegRange r;
    r.q = 0;
    r.lo = 0;
    r.hi = 0;
    if (valid) r = seg_range<IMPLICIT>(S, seg);
    const int len = r.hi - r.lo;
    const bool is_short = valid && len <= kShortSeg;
    float4 G[D / 16];
    BwdSmall sm;
    if (__any_sync(RG_FULL_MASK, is_short)) {  // short segments: one per 4-lane group
        bwd_group<D, HAS_HIDDEN, IMPLICIT>(S, is_short, seg, r.q, r.lo, len, hidden, as8, rela, ar8, aq8, w8, ba, g_agg,
                                           g_rela, g_ar8, G, sm);
        if (is_short) {
            if (g_hidden) {
                float4 *o = reinterpret_cast<float4 *>(g_hidden + (size_t)seg * D);
#pragma unroll
                for (int v = 0; v < D / 16; ++v) o[v * 4 + ql] = G[v];
            }
            float2 *o = reinterpret_cast<float2 *>(node_small + (size_t)seg * 24);
            o[ql] = sm.z;
            o[4 + ql] = sm.wz;
            o[8 + ql] = make_float2(ql == 0 ? sm.gl : 0.f, 0.f);
        }
    }
    unsigned long_mask = __ballot_sync(RG_FULL_MASK, valid && !is_short && ql == 0);
    while (long_mask) {  // the long ones, warp-wide, one after the other (warp-uniform loop)
        const int src = __ffs(long_mask) - 1;
        long_mask &= long_mask - 1;
        const int q = __shfl_sync(RG_FULL_MASK, r.q, src);
        const int lo = __shfl_sync(RG_FULL_MASK, r.lo, src);
        int hi = __shfl_sync(RG_FULL_MASK, r.hi, src);
        const int64_t sg = seg0 + (src >> 2);
        if (has_heavy && hi - lo > own0) {  // without a queue the owner warp does it all
            enqueue_heavy(H, sg, hi - lo, lane, RG_HEAVY_SUB_BWD, own0);
            hi = lo + own0;
        }
        bwd_range<D, HAS_HIDDEN, IMPLICIT>(S, sg, q, lo, hi, hidden, as8, rela, ar8, aq8, w8, ba, g_agg, g_rela, g_ar8,
                                           G, sm);
        if (g_hidden) store_row<D>(g_hidden + (size_t)sg * D, G, lane);
        store_small(node_small + (size_t)sg * 24, sm, lane);
    }
}

// Persistent variant (implicit path): warps loop over segments on their own, so a block never idles
// behind its longest segment, and the relation tables are read from shared memory (see k_edge_fwd_p).
template <int D, bool HAS_HIDDEN>
__global__ void __launch_bounds__(kPWarpsB * 32, 2) k_edge_bwd_p(rg_segments S, const float *__restrict__ hidden,
                                                               const float *__restrict__ as8,
                                                               const float *__restrict__ rela,
                                                               const float *__restrict__ ar8,
                                                               const float *__restrict__ aq8,
                                                               const float *__restrict__ w8,
                                                               const float *__restrict__ b_alpha,
                                                               const float *__restrict__ g_agg, float *g_hidden,
                                                               float *node_small, float *g_rela, float *g_ar8,
                                                               int copies, rg_heavy H, int has_heavy, int own0) {
    extern __shared__ __align__(128) float s_tab[];
    __shared__ __align__(8) unsigned long long tab_bar;
    select_copy<D>(g_rela, g_ar8, copies, S.n_table_rows);
    const int rows = S.n_table_rows;
    float *s_rela = s_tab, *s_ar8 = s_tab + (size_t)rows * D;
    stage_tables(&tab_bar, s_rela, rela, (uint32_t)rows * D * 4, s_ar8, ar8, (uint32_t)rows * 32);
    const int lane = threadIdx.x & 31;
    const int64_t n_true = S.n_seg_dev ? *S.n_seg_dev : S.n_seg;
    const float ba = __ldg(b_alpha);
    const int64_t stride = (int64_t)gridDim.x * kPWarpsB;
    for (int64_t seg = (int64_t)blockIdx.x * kPWarpsB + (threadIdx.x >> 5); seg < n_true; seg += stride) {
        SegRange r = seg_range<true>(S, seg);
        int hi = r.hi;
        if (has_heavy && r.hi - r.lo > own0) {
            enqueue_heavy(H, seg, r.hi - r.lo, lane, RG_HEAVY_SUB_BWD, own0);
            hi = r.lo + own0;
        }
        float4 G[D / 16];
        BwdSmall sm;
        bwd_range<D, HAS_HIDDEN, true, true>(S, seg, r.q, r.lo, hi, hidden, as8, rela, ar8, aq8, w8, ba, g_agg, g_rela,
                                             g_ar8, G, sm, s_rela, s_ar8);
        if (g_hidden) store_row<D>(g_hidden + (size_t)seg * D, G, lane);
        store_small(node_small + (size_t)seg * 24, sm, lane);
    }
    // (the heavy chunks run as a second kernel here: draining the queue inside this kernel, as the forward
    // does, was measured slower for the backward -- 1.40 -> 1.53..1.65 ms on the FB15k-237 training step)
}

template <int D, bool HAS_HIDDEN, bool IMPLICIT>
__global__ void __launch_bounds__(kBlock) k_edge_bwd_chunks(rg_segments S, const float *__restrict__ hidden,
                                                            const float *__restrict__ as8,
                                                            const float *__restrict__ rela,
                                                            const float *__restrict__ ar8,
                                                            const float *__restrict__ aq8,
                                                            const float *__restrict__ w8,
                                                            const float *__restrict__ b_alpha,
                                                            const float *__restrict__ g_agg, float *g_rela,
                                                            float *g_ar8, int copies, rg_heavy H, int own0) {
    const int lane = threadIdx.x & 31;
    select_copy<D>(g_rela, g_ar8, copies, S.n_table_rows);
    if (H.counters[2]) return;
    const int n_chunks = min(H.counters[0], H.max_chunks);
    const int stride = gridDim.x * kWarpsPerBlock;
    for (int c = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); c < n_chunks; c += stride) {
        const int64_t seg = H.chunk_seg[c];
        SegRange r = seg_range<IMPLICIT>(S, seg);
        const int lo = r.lo + own0 + (H.chunk_idx[c] - 1) * RG_HEAVY_SUB_BWD;   // chunk_idx is 1-based
        const int hi = min(r.hi, lo + RG_HEAVY_SUB_BWD);
        float4 G[D / 16];
        BwdSmall sm;
        bwd_range<D, HAS_HIDDEN, IMPLICIT>(S, seg, r.q, lo, hi, hidden, as8, rela, ar8, aq8, w8, __ldg(b_alpha), g_agg,
                                           g_rela, g_ar8, G, sm);
        float *row = H.partial + (size_t)c * (D + 24);
        store_row<D>(row, G, lane);
        store_small(row + D, sm, lane);
    }
}

int check_segments(const rg_segments *s) {
    if (!s || s->n_seg < 0 || !s->seg_query || !s->adj) return RG_ERR_BAD_ARG;
    if (s->mode == 0) {
        if (!s->seg_ptr) return RG_ERR_BAD_ARG;
    } else if (s->mode == 1) {
        if (!s->seg_ent || !s->ent_ptr || !s->peer_dict || s->n_ent <= 0) return RG_ERR_BAD_ARG;
    } else {
        return RG_ERR_BAD_ARG;
    }
    if (rg_cdiv(s->n_seg, kWarpsPerBlock) >= (int64_t)INT32_MAX) return RG_ERR_TOO_LARGE;
    return RG_OK;
}

int check_heavy(const rg_heavy *h) {
    if (!h) return RG_OK;
    if (h->max_chunks < 0 || h->max_nodes < 0 || !h->counters) return RG_ERR_BAD_ARG;
    if (h->max_chunks > 0 && (!h->chunk_seg || !h->chunk_idx || !h->partial)) return RG_ERR_BAD_ARG;
    if (h->max_nodes > 0 && (!h->node_seg || !h->node_base || !h->node_n)) return RG_ERR_BAD_ARG;
    return RG_OK;
}

constexpr int kHeavyGrid = 148 * 4;
// 8 segments per warp only pays when that still leaves every SM several blocks of warps
constexpr int64_t kBlock8MinSegs = 8 * 148 * 64;

// 0 = one warp per segment, 1 = eight segments per warp, 2 = persistent + shared-memory tables
inline int pick_variant(const rg_segments *seg, int D) {
    const size_t tab_bytes = (size_t)seg->n_table_rows * (D + 8) * sizeof(float);
    if (seg->mode == 1 && seg->n_table_rows > 0 && tab_bytes <= 110 * 1024 && seg->n_seg >= 4096) return 2;
    return seg->n_seg >= kBlock8MinSegs ? 1 : 0;
}

template <int D, bool HH, bool IM>
int launch_fwd(const rg_segments *seg, const float *hidden, const float *as8, const float *rela, const float *ar8,
               const float *aq8, const float *w8, const float *b_alpha, float *agg, const rg_heavy *heavy,
               cudaStream_t st) {
    rg_heavy H = {};
    const int has_heavy = heavy && heavy->max_chunks > 0 && heavy->max_nodes > 0;
    if (has_heavy) H = *heavy;
    if (seg->n_seg == 0) return RG_OK;
    if (has_heavy) {   // queue state: counters and the published flags (chunk_idx) start from zero
        RG_CUDA_CALL(cudaMemsetAsync(H.counters, 0, 8 * sizeof(int32_t), st));
        RG_CUDA_CALL(cudaMemsetAsync(H.chunk_idx, 0, (size_t)H.max_chunks * sizeof(int32_t), st));
    }
    const size_t tab_bytes = (size_t)seg->n_table_rows * (D + 8) * sizeof(float);
    bool persistent = false;
    if constexpr (IM) {
        // relation tables staged in shared memory when two 16-warp CTAs still fit one SM
        if (pick_variant(seg, D) == 2) {
            auto kern = k_edge_fwd_p<D, HH>;
            RG_CUDA_CALL(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tab_bytes));
            int dev = 0, n_sm = 148, per_sm = 1;
            RG_CUDA_CALL(cudaGetDevice(&dev));
            RG_CUDA_CALL(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
            RG_CUDA_CALL(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kPWarps * 32, tab_bytes));
            if (per_sm >= 1) {
                const int64_t want = rg_cdiv(seg->n_seg, kPWarps);
                const unsigned grid = (unsigned)std::min<int64_t>(want, (int64_t)n_sm * per_sm);
                kern<<<grid, kPWarps * 32, tab_bytes, st>>>(*seg, hidden, as8, rela, ar8, aq8, w8, b_alpha, agg, H,
                                                            has_heavy);
                RG_LAUNCH_CHECK();
                persistent = true;
            }
        }
    }
    if (!persistent) {
        if (seg->n_seg >= kBlock8MinSegs) {
            const unsigned grid = (unsigned)rg_cdiv(seg->n_seg, kWarpsPerBlock * 8);
            k_edge_fwd<D, HH, IM, true><<<grid, kBlock, 0, st>>>(*seg, hidden, as8, rela, ar8, aq8, w8, b_alpha, agg, H,
                                                                 has_heavy);
        } else {
            const unsigned grid = (unsigned)rg_cdiv(seg->n_seg, kWarpsPerBlock);
            k_edge_fwd<D, HH, IM, false><<<grid, kBlock, 0, st>>>(*seg, hidden, as8, rela, ar8, aq8, w8, b_alpha, agg, H,
                                                                  has_heavy);
        }
        RG_LAUNCH_CHECK();
    }
    if (has_heavy) {
        // all chunks (non-persistent main kernel) or whatever the persistent kernel's own drain left positive
        k_edge_fwd_chunks<D, HH, IM><<<kHeavyGrid, kBlock, 0, st>>>(*seg, hidden, as8, rela, ar8, aq8, w8, b_alpha, H);
        RG_LAUNCH_CHECK();
        k_heavy_fixup<<<kHeavyGrid, kBlock, 0, st>>>(H, D, agg, D, nullptr, 0);
        RG_LAUNCH_CHECK();
    }
    return RG_OK;
}

template <int D, bool HH, bool IM>
int launch_bwd(const rg_segments *seg, const float *hidden, const float *as8, const float *rela, const float *ar8,
               const float *aq8, const float *w8, const float *b_alpha, const float *g_agg, float *g_hidden,
               float *node_small, float *g_rela, float *g_ar8, int copies, const rg_heavy *heavy, cudaStream_t st) {
    rg_heavy H = {};
    const int has_heavy = heavy && heavy->max_chunks > 0 && heavy->max_nodes > 0;
    if (has_heavy) H = *heavy;
    if (seg->n_seg == 0) return RG_OK;
    if (has_heavy) {   // queue state: counters and the published flags (chunk_idx) start from zero
        RG_CUDA_CALL(cudaMemsetAsync(H.counters, 0, 8 * sizeof(int32_t), st));
        RG_CUDA_CALL(cudaMemsetAsync(H.chunk_idx, 0, (size_t)H.max_chunks * sizeof(int32_t), st));
    }
    const size_t tab_bytes = (size_t)seg->n_table_rows * (D + 8) * sizeof(float);
    bool persistent = false;
    int own0 = RG_HEAVY_CHUNK_BWD;   // slots of a heavy segment its owner warp keeps; the rest: RG_HEAVY_SUB_BWD pieces
    if constexpr (IM) {
        if (pick_variant(seg, D) == 2) {
            auto kern = k_edge_bwd_p<D, HH>;
            RG_CUDA_CALL(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tab_bytes));
            int dev = 0, n_sm = 148, per_sm = 1;
            RG_CUDA_CALL(cudaGetDevice(&dev));
            RG_CUDA_CALL(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
            RG_CUDA_CALL(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kPWarpsB * 32, tab_bytes));
            if (per_sm >= 1) {
                const int64_t want = rg_cdiv(seg->n_seg, kPWarpsB);
                const unsigned grid = (unsigned)std::min<int64_t>(want, (int64_t)n_sm * per_sm);
                own0 = RG_HEAVY_CHUNK_BWD;
                kern<<<grid, kPWarpsB * 32, tab_bytes, st>>>(*seg, hidden, as8, rela, ar8, aq8, w8, b_alpha, g_agg,
                                                            g_hidden, node_small, g_rela, g_ar8, copies, H, has_heavy,
                                                            own0);
                RG_LAUNCH_CHECK();
                persistent = true;
            }
        }
    }
    if (!persistent) {
        if (seg->n_seg >= kBlock8MinSegs) {
            const unsigned grid = (unsigned)rg_cdiv(seg->n_seg, kWarpsPerBlock * 8);
            own0 = RG_HEAVY_CHUNK_BWD;
            k_edge_bwd<D, HH, IM, true><<<grid, kBlock, 0, st>>>(*seg, hidden, as8, rela, ar8, aq8, w8, b_alpha, g_agg,
                                                                 g_hidden, node_small, g_rela, g_ar8, copies, H,
                                                                 has_heavy, own0);
        } else {
            const unsigned grid = (unsigned)rg_cdiv(seg->n_seg, kWarpsPerBlock);
            // few segments (layer 0 of a training step: one per query): the owner keeps only one small piece, so
            // that a hub subject does not leave one warp walking hundreds of slots while the GPU idles
            own0 = seg->n_seg < 4096 ? RG_HEAVY_SUB_BWD : RG_HEAVY_CHUNK_BWD;
            k_edge_bwd<D, HH, IM, false><<<grid, kBlock, 0, st>>>(*seg, hidden, as8, rela, ar8, aq8, w8, b_alpha, g_agg,
                                                                  g_hidden, node_small, g_rela, g_ar8, copies, H,
                                                                  has_heavy, own0);
        }
        RG_LAUNCH_CHECK();
    }
    if (has_heavy) {
        k_edge_bwd_chunks<D, HH, IM><<<kHeavyGrid, kBlock, 0, st>>>(*seg, hidden, as8, rela, ar8, aq8, w8, b_alpha,
                                                                    g_agg, g_rela, g_ar8, copies, H, own0);
        RG_LAUNCH_CHECK();
        k_heavy_fixup<<<kHeavyGrid, kBlock, 0, st>>>(H, D + 24, g_hidden, D, node_small, 24);
        RG_LAUNCH_CHECK();
    }
    return RG_OK;
}

#define RG_DISPATCH_D(D_, CALL)              \
    switch (D_) {                            \
        case 16: { constexpr int DD = 16; CALL; } break; \
        case 32: { constexpr int DD = 32; CALL; } break; \
        case 48: { constexpr int DD = 48; CALL; } break; \
        case 64: { constexpr int DD = 64; CALL; } break; \
        default: return RG_ERR_UNSUPPORTED;  \
    }

}  // namespace

extern "C" {

int rg_edge_agg_fwd(const rg_segments *seg, int32_t hidden_dim, const float *hidden, const float *as8,
                    const float *rela, const float *ar8, const float *aq8, const float *w8, const float *b_alpha,
                    float *agg, const rg_heavy *heavy, void *stream) {
    int rc = check_segments(seg);
    if (rc) return rc;
    rc = check_heavy(heavy);
    if (rc) return rc;
    if (!rela || !ar8 || !aq8 || !w8 || !b_alpha || !agg) return RG_ERR_BAD_ARG;
    if ((hidden == nullptr) != (as8 == nullptr)) return RG_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const bool hh = hidden != nullptr, im = seg->mode == 1;
#define RG_FWD(HH, IM) \
    RG_DISPATCH_D(hidden_dim, rc = (launch_fwd<DD, HH, IM>(seg, hidden, as8, rela, ar8, aq8, w8, b_alpha, agg, heavy, st)))
    if (hh && im) { RG_FWD(true, true); }
    else if (hh && !im) { RG_FWD(true, false); }
    else if (!hh && im) { RG_FWD(false, true); }
    else { RG_FWD(false, false); }
#undef RG_FWD
    return rc;
}

int rg_edge_agg_bwd(const rg_segments *seg, int32_t hidden_dim, const float *hidden, const float *as8,
                    const float *rela, const float *ar8, const float *aq8, const float *w8, const float *b_alpha,
                    const float *g_agg, float *g_hidden, float *node_small, float *g_rela, float *g_ar8,
                    int32_t grad_copies, const rg_heavy *heavy, void *stream) {
    int rc = check_segments(seg);
    if (rc) return rc;
    rc = check_heavy(heavy);
    if (rc) return rc;
    if (!rela || !ar8 || !aq8 || !w8 || !b_alpha || !g_agg || !node_small || !g_rela || !g_ar8)
        return RG_ERR_BAD_ARG;
    if ((hidden == nullptr) != (as8 == nullptr)) return RG_ERR_BAD_ARG;
    if (!hidden && g_hidden) return RG_ERR_BAD_ARG;
    if (grad_copies < 1 || (grad_copies > 1 && seg->n_table_rows <= 0)) return RG_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const bool hh = hidden != nullptr, im = seg->mode == 1;
#define RG_BWD(HH, IM)                                                                                             \
    RG_DISPATCH_D(hidden_dim, rc = (launch_bwd<DD, HH, IM>(seg, hidden, as8, rela, ar8, aq8, w8, b_alpha, g_agg,   \
                                                           g_hidden, node_small, g_rela, g_ar8, grad_copies,   \
                                                           heavy, st)))
    if (hh && im) { RG_BWD(true, true); }
    else if (hh && !im) { RG_BWD(true, false); }
    else if (!hh && im) { RG_BWD(false, true); }
    else { RG_BWD(false, false); }
#undef RG_BWD
    return rc;
}

int rg_edge_agg_variant(const rg_segments *seg, int32_t hidden_dim) {
    int rc = check_segments(seg);
    if (rc) return rc;
    if (hidden_dim != 16 && hidden_dim != 32 && hidden_dim != 48 && hidden_dim != 64) return RG_ERR_UNSUPPORTED;
    return pick_variant(seg, hidden_dim);
}

}  // extern "C"
