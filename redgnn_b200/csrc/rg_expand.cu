// Frontier expansion kernels: the B200 replacement of DataLoader.get_neighbors
// (reference Static/transductive/load_data.py:106-131, Static/inductive/load_data.py:115-143).
//
// Formulation (sort-free, deterministic, all integer):
//   frontier  = bit matrix over (query, entity), kept entity-major (emask) and query-major with
//               rank prefixes (dict);
//   hop       = one scan over the fact rows: cnt[f] = popc(emask[head[f]]) edges, and
//               emask_next[tail[f]] |= emask[head[f]] (idempotent OR => order independent);
//   dedup     = bit-matrix transpose of emask_next + prefix popcount = sorted unique (b, tail);
//   edge list = exclusive scan of cnt[] gives every fact its output offset; bits are emitted
//               high-to-low => (fact ascending, batch descending) = the reference's order.
#include "rg_common.cuh"

namespace {

constexpr int kBlock = 256;

// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_set_nodes(const int64_t *__restrict__ nodes, int64_t n_nodes,
                                                      int n_query, int n_ent, uint32_t *emask,
                                                      uint32_t *dict, int64_t *counts) {
    const int Wn = rg_words_query(n_query), We = rg_words_ent(n_ent);
    int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n_nodes) return;
    longlong2 be = reinterpret_cast<const longlong2 *>(nodes)[i];
    if (be.x < 0 || be.x >= n_query || be.y < 0 || be.y >= n_ent) {
        atomicOr(reinterpret_cast<unsigned long long *>(counts + RG_CNT_ERR), 1ull);
        return;
    }
    int b = (int)be.x, e = (int)be.y;
    atomicOr(&emask[(size_t)e * Wn + (b >> 5)], 1u << (b & 31));
    atomicOr(&dict[((size_t)b * We + (e >> 5)) * 2], 1u << (e & 31));
}

// popcount of TILE dictionary words per block
__global__ void __launch_bounds__(kBlock) k_dict_reduce(const uint32_t *__restrict__ dict, int64_t n_words,
                                                        uint32_t *blocksum) {
    __shared__ uint32_t sm[kBlock / 32 + 1];
    int64_t base = (int64_t)blockIdx.x * RG_TILE;
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < RG_TILE / kBlock; ++k) {
        int64_t i = base + k * kBlock + threadIdx.x;
        if (i < n_words) c += __popc(dict[2 * i]);
    }
    uint32_t tot = rg_block_sum<kBlock>(c, sm);
    if (threadIdx.x == 0) blocksum[blockIdx.x] = tot;
}

// single block: exclusive scan of nb block sums into 64-bit prefixes; prefix[nb] = total
__global__ void __launch_bounds__(1024) k_scan_blocksums(const uint32_t *__restrict__ blocksum, int64_t nb,
                                                         unsigned long long *prefix, int64_t *count_out) {
    __shared__ unsigned long long sm[1024 / 32 + 1];
    unsigned long long carry = 0;
    for (int64_t base = 0; base < nb; base += 1024) {
        int64_t i = base + threadIdx.x;
        unsigned long long v = (i < nb) ? (unsigned long long)blocksum[i] : 0ull, tot;
        unsigned long long ex = rg_block_exclusive_scan<1024>(v, sm, tot);
        if (i < nb) prefix[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) {
        prefix[nb] = carry;
        if (count_out) *count_out = (int64_t)carry;
    }
}

// write the rank prefix of every dictionary word
__global__ void __launch_bounds__(kBlock) k_dict_apply(uint32_t *dict, int64_t n_words,
                                                       const unsigned long long *__restrict__ blockprefix) {
    __shared__ uint32_t sm[kBlock / 32 + 1];
    constexpr int IPT = RG_TILE / kBlock;
    int64_t first = (int64_t)blockIdx.x * RG_TILE + (int64_t)threadIdx.x * IPT;
    uint32_t c[IPT], s = 0;
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        int64_t i = first + k;
        c[k] = (i < n_words) ? __popc(dict[2 * i]) : 0;
        s += c[k];
    }
    uint32_t tot;
    uint32_t ex = rg_block_exclusive_scan<kBlock>(s, sm, tot);
    uint32_t run = (uint32_t)blockprefix[blockIdx.x] + ex;
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        int64_t i = first + k;
        if (i < n_words) dict[2 * i + 1] = run;
        run += c[k];
    }
}

// per fact: number of edges it produces + OR the head's query set into the tail's row
__global__ void __launch_bounds__(kBlock) k_fact_count(const int32_t *__restrict__ head,
                                                       const int32_t *__restrict__ tail, int64_t n_fact,
                                                       const uint32_t *__restrict__ emask_in,
                                                       uint32_t *emask_out, int Wn, uint32_t *blocksum) {
    __shared__ uint32_t sm[kBlock / 32 + 1];
    int64_t base = (int64_t)blockIdx.x * RG_TILE;
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < RG_TILE / kBlock; ++k) {
        int64_t f = base + k * kBlock + threadIdx.x;
        if (f < n_fact) {
            const uint32_t *row = emask_in + (size_t)head[f] * Wn;
            uint32_t cc = 0;
            for (int w = 0; w < Wn; ++w) cc += __popc(row[w]);
            if (cc) {
                uint32_t *orow = emask_out + (size_t)tail[f] * Wn;
                for (int w = 0; w < Wn; ++w) {
                    uint32_t m = row[w];
                    if (m && (orow[w] & m) != m) atomicOr(&orow[w], m);
                }
            }
            c += cc;
        }
    }
    uint32_t tot = rg_block_sum<kBlock>(c, sm);
    if (threadIdx.x == 0) blocksum[blockIdx.x] = tot;
}

// entity-major emask -> query-major dictionary bits (32x32 bit-block transpose per warp)
__global__ void __launch_bounds__(kBlock) k_transpose(const uint32_t *__restrict__ emask, int n_ent, int Wn,
                                                      int n_query, uint32_t *dict, int We) {
    const int lane = threadIdx.x & 31;
    int64_t task = (int64_t)blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
    if (task >= (int64_t)We * Wn) return;
    int eb = (int)(task / Wn), w = (int)(task % Wn);
    int e = eb * 32 + lane;
    uint32_t x = (e < n_ent) ? emask[(size_t)e * Wn + w] : 0u;
    uint32_t mine = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        uint32_t y = __ballot_sync(RG_FULL_MASK, (x >> j) & 1u);
        if (lane == j) mine = y;
    }
    int b = w * 32 + lane;
    if (b < n_query) dict[((size_t)b * We + eb) * 2] = mine;
}

// enumerate the set bits of the dictionary in (b, e) order
__global__ void __launch_bounds__(kBlock) k_emit_nodes(const uint32_t *__restrict__ dict, int n_query, int We,
                                                       int64_t *nodes64, int32_t *node_b, int32_t *node_e) {
    int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= (int64_t)n_query * We) return;
    uint2 d = reinterpret_cast<const uint2 *>(dict)[i];
    int b = (int)(i / We), wi = (int)(i % We);
    uint32_t bits = d.x;
    size_t pos = d.y;
    while (bits) {
        int bit = __ffs(bits) - 1;
        bits &= bits - 1;
        int e = wi * 32 + bit;
        if (nodes64) reinterpret_cast<longlong2 *>(nodes64)[pos] = make_longlong2(b, e);
        if (node_b) node_b[pos] = b;
        if (node_e) node_e[pos] = e;
        ++pos;
    }
}

__global__ void __launch_bounds__(kBlock) k_remap(const uint32_t *__restrict__ din,
                                                  const uint32_t *__restrict__ dout, int64_t n_words,
                                                  int64_t *remap64, int32_t *remap32, int32_t *inverse32) {
    int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n_words) return;
    uint2 a = reinterpret_cast<const uint2 *>(din)[i];
    if (!a.x) return;
    uint2 o = reinterpret_cast<const uint2 *>(dout)[i];
    uint32_t bits = a.x;
    size_t pos = a.y;
    while (bits) {
        int bit = __ffs(bits) - 1;
        bits &= bits - 1;
        uint32_t r = o.y + __popc(o.x & ((1u << bit) - 1u));
        if (remap64) remap64[pos] = (int64_t)r;
        if (remap32) remap32[pos] = (int32_t)r;
        if (inverse32) inverse32[r] = (int32_t)pos;
        ++pos;
    }
}

// emit sampled_edges[E][6] in reference order (fact ascending, batch index descending).
// blockIdx.x = tile of RG_TILE fact rows; the kEmitParts blocks of blockIdx.y share the tile: each
// rebuilds the tile's edge offsets (cheap next to the edges) and its warps take every
// (8 * kEmitParts)-th fact.  A fact's edges differ only in the query: lane l of a warp owns query
// w*32 + 31 - l of mask word w, so the row position is a popcount (no search), the rows are built
// in shared memory and leave as coalesced 16-byte pieces of one contiguous run.  Queries whose
// frontier is complete (qinfo count == n_ent) skip the dictionary probe: rank = base + entity.
constexpr int kEmitParts = 4;

__device__ __forceinline__ int emit_complete_base(const int32_t *qinfo, int b, int n_query, int n_ent) {
    if (qinfo == nullptr || b >= n_query) return -1;
    const int2 qi = __ldg(reinterpret_cast<const int2 *>(qinfo) + b);
    return qi.y == n_ent ? qi.x : -1;
}

template <bool MASK_SMEM>  // Wn <= 2: the head's query mask and the per-lane query info are cached
__global__ void __launch_bounds__(kBlock) k_emit_edges(const int32_t *__restrict__ head,
                                                       const int32_t *__restrict__ rel,
                                                       const int32_t *__restrict__ tail, int64_t n_fact,
                                                       const uint32_t *__restrict__ emask_in, int Wn,
                                                       const uint32_t *__restrict__ dict_in,
                                                       const uint32_t *__restrict__ dict_out, int We,
                                                       const int32_t *__restrict__ qinfo_in,
                                                       const int32_t *__restrict__ qinfo_out, int n_query, int n_ent,
                                                       const unsigned long long *__restrict__ blockprefix,
                                                       int64_t *edges) {
    __shared__ uint32_t s_off[RG_TILE + 1];
    __shared__ int32_t s_head[RG_TILE];
    __shared__ int2 s_rt[RG_TILE];
    __shared__ uint2 s_mask[MASK_SMEM ? RG_TILE : 1];
    __shared__ uint32_t sm[kBlock / 32 + 1];
    __shared__ __align__(16) longlong2 s_rows[kBlock / 32][96];
    constexpr int IPT = RG_TILE / kBlock;
    const int64_t base = (int64_t)blockIdx.x * RG_TILE;
    const unsigned long long ebase = blockprefix[blockIdx.x];
    if (blockprefix[blockIdx.x + 1] == ebase) return;  // tile without edges (block-uniform)
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        int idx = k * kBlock + threadIdx.x;
        int64_t f = base + idx;
        int h = -1;
        uint32_t c = 0;
        uint2 mk = make_uint2(0u, 0u);
        int2 rt = make_int2(0, 0);
        if (f < n_fact) {
            h = head[f];
            rt = make_int2(rel[f], tail[f]);
            const uint32_t *row = emask_in + (size_t)h * Wn;
            if (MASK_SMEM) {
                mk.x = row[0];
                if (Wn > 1) mk.y = row[1];
                c = __popc(mk.x) + __popc(mk.y);
            } else {
                for (int w = 0; w < Wn; ++w) c += __popc(row[w]);
            }
        }
        s_head[idx] = h;
        s_rt[idx] = rt;
        if (MASK_SMEM) s_mask[idx] = mk;
        s_off[idx] = c;
    }
    __syncthreads();
    {
        uint32_t a[IPT], t = 0;
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
            a[k] = s_off[threadIdx.x * IPT + k];
            t += a[k];
        }
        uint32_t tot;
        uint32_t ex = rg_block_exclusive_scan<kBlock>(t, sm, tot);
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
            s_off[threadIdx.x * IPT + k] = ex;
            ex += a[k];
        }
        if (threadIdx.x == kBlock - 1) s_off[RG_TILE] = tot;
    }
    __syncthreads();
    const uint2 *din = reinterpret_cast<const uint2 *>(dict_in);
    const uint2 *dout = reinterpret_cast<const uint2 *>(dict_out);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bitpos = 31 - lane;
    longlong2 *my_rows = s_rows[warp];
    int cb_in[2] = {-1, -1}, cb_out[2] = {-1, -1};
    if (MASK_SMEM) {
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            cb_in[w] = emit_complete_base(qinfo_in, w * 32 + bitpos, n_query, n_ent);
            cb_out[w] = emit_complete_base(qinfo_out, w * 32 + bitpos, n_query, n_ent);
        }
    }
    for (int idx = blockIdx.y * (kBlock / 32) + warp; idx < RG_TILE; idx += kEmitParts * (kBlock / 32)) {
        const uint32_t off = s_off[idx];
        if (s_off[idx + 1] == off) continue;  // warp-uniform
        const int h = s_head[idx];
        const int2 rt = s_rt[idx];
        longlong2 *dst = reinterpret_cast<longlong2 *>(edges + 6 * (int64_t)(ebase + off));
        for (int w = Wn - 1; w >= 0; --w) {
            uint32_t m;
            if (MASK_SMEM)
                m = w ? s_mask[idx].y : s_mask[idx].x;
            else
                m = emask_in[(size_t)h * Wn + w];
            if (!m) continue;  // warp-uniform
            if ((m >> bitpos) & 1u) {
                const int k = __popc((m >> bitpos) >> 1);  // set bits above mine = rows before mine
                const int b = w * 32 + bitpos;
                const int ci = MASK_SMEM ? (w ? cb_in[1] : cb_in[0]) : emit_complete_base(qinfo_in, b, n_query, n_ent);
                const int co = MASK_SMEM ? (w ? cb_out[1] : cb_out[0])
                                         : emit_complete_base(qinfo_out, b, n_query, n_ent);
                const uint32_t hi_idx = ci >= 0 ? (uint32_t)(ci + h) : rg_rank(din[(size_t)b * We + (h >> 5)], h);
                const uint32_t ti_idx =
                    co >= 0 ? (uint32_t)(co + rt.y) : rg_rank(dout[(size_t)b * We + (rt.y >> 5)], rt.y);
                my_rows[k * 3 + 0] = make_longlong2(b, h);
                my_rows[k * 3 + 1] = make_longlong2(rt.x, rt.y);
                my_rows[k * 3 + 2] = make_longlong2((long long)hi_idx, (long long)ti_idx);
            }
            __syncwarp();
            const int n_piece = 3 * __popc(m);
#pragma unroll
            for (int j = 0; j < 3; ++j)
                if (lane + 32 * j < n_piece) dst[lane + 32 * j] = my_rows[lane + 32 * j];
            dst += n_piece;
            __syncwarp();
        }
    }
}

// per query: {rank of its first node, number of its nodes}; a query whose count equals n_ent has a
// COMPLETE frontier: rank(b, e) = base + e and every candidate edge is active (no dictionary probe)
__global__ void __launch_bounds__(kBlock) k_query_info(const uint32_t *__restrict__ dict, int n_query, int We,
                                                       const int64_t *__restrict__ total, int32_t *qinfo) {
    int q = blockIdx.x * kBlock + threadIdx.x;
    if (q >= n_query) return;
    uint32_t base = dict[((size_t)q * We) * 2 + 1];
    uint32_t next = (q + 1 < n_query) ? dict[((size_t)(q + 1) * We) * 2 + 1] : (uint32_t)*total;
    qinfo[2 * q] = (int32_t)base;
    qinfo[2 * q + 1] = (int32_t)(next - base);
}

int check_frontier(const rg_frontier *fr) {
    if (!fr || !fr->emask || !fr->dict || fr->n_query <= 0 || fr->n_ent <= 0) return RG_ERR_BAD_ARG;
    if ((int64_t)fr->n_query * fr->n_ent >= (int64_t)INT32_MAX) return RG_ERR_TOO_LARGE;
    return RG_OK;
}

// Tiny dictionaries (<= 8 K words: small KGs / small batches): ONE CTA walks all words with a running block
// scan and also writes the per-query {base, count} -- one launch instead of reduce + scan + apply +
// query-info.  (Beyond two passes of the CTA the four-kernel chain is faster again: measured on the
// FB15k-237 shape, 29 K words, +20 us per hop inside the captured forward.)
constexpr int64_t kSmallDictWords = 8 * 1024;
__global__ void __launch_bounds__(1024) k_dict_prefix_small(uint32_t *dict, int64_t n_words, int n_query, int We,
                                                            int64_t *count_out, int32_t *qinfo) {
    __shared__ uint32_t sm[1024 / 32 + 1];
    constexpr int IPT = 4;
    uint32_t carry = 0;
    for (int64_t base = 0; base < n_words; base += 1024 * IPT) {
        const int64_t first = base + (int64_t)threadIdx.x * IPT;
        uint32_t c[IPT], s = 0;
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
            const int64_t i = first + k;
            c[k] = (i < n_words) ? __popc(dict[2 * i]) : 0;
            s += c[k];
        }
        uint32_t tot;
        uint32_t run = carry + rg_block_exclusive_scan<1024>(s, sm, tot);
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
            const int64_t i = first + k;
            if (i < n_words) {
                dict[2 * i + 1] = run;
                if (qinfo && i % We == 0) qinfo[2 * (i / We)] = (int32_t)run;   // first word of a query's row
            }
            run += c[k];
        }
        carry += tot;
    }
    if (threadIdx.x == 0 && count_out) *count_out = (int64_t)carry;
    if (qinfo) {
        __syncthreads();
        for (int q = threadIdx.x; q < n_query; q += 1024) {
            const int32_t b0 = qinfo[2 * q], b1 = (q + 1 < n_query) ? qinfo[2 * (q + 1)] : (int32_t)carry;
            qinfo[2 * q + 1] = b1 - b0;
        }
    }
}

// dictionary bits -> rank prefixes; total written to counts[which]
int dict_prefix(const rg_frontier *fr, const RgWorkspace &w, int64_t *count_out, cudaStream_t st) {
    const int64_t n_words = (int64_t)fr->n_query * rg_words_ent(fr->n_ent);
    if (n_words <= kSmallDictWords) {
        k_dict_prefix_small<<<1, 1024, 0, st>>>(fr->dict, n_words, fr->n_query, rg_words_ent(fr->n_ent), count_out,
                                              fr->qinfo);
        RG_LAUNCH_CHECK();
        return RG_OK;
    }
    const int64_t nb = rg_cdiv(n_words, RG_TILE);
    k_dict_reduce<<<(unsigned)nb, kBlock, 0, st>>>(fr->dict, n_words, w.dict_blocksum);
    RG_LAUNCH_CHECK();
    k_scan_blocksums<<<1, 1024, 0, st>>>(w.dict_blocksum, nb, w.dict_blockprefix, count_out);
    RG_LAUNCH_CHECK();
    k_dict_apply<<<(unsigned)nb, kBlock, 0, st>>>(fr->dict, n_words, w.dict_blockprefix);
    RG_LAUNCH_CHECK();
    if (fr->qinfo) {
        k_query_info<<<(unsigned)rg_cdiv(fr->n_query, kBlock), kBlock, 0, st>>>(fr->dict, fr->n_query,
                                                                               rg_words_ent(fr->n_ent), count_out,
                                                                               fr->qinfo);
        RG_LAUNCH_CHECK();
    }
    return RG_OK;
}

}  // namespace

extern "C" {

size_t rg_frontier_emask_bytes(int32_t n_query, int32_t n_ent) {
    return rg_align256((size_t)n_ent * rg_words_query(n_query) * 4);
}

size_t rg_frontier_dict_bytes(int32_t n_query, int32_t n_ent) {
    return rg_align256((size_t)n_query * rg_words_ent(n_ent) * 8);
}

size_t rg_workspace_bytes(int32_t n_query, int32_t n_ent, int64_t n_fact) {
    return rg_carve(nullptr, n_query, n_ent, n_fact < 0 ? 0 : n_fact).total_bytes;
}

int rg_frontier_from_nodes(const int64_t *nodes, int64_t n_nodes, rg_frontier *fr, int64_t *counts,
                           void *ws, size_t ws_bytes, void *stream) {
    int rc = check_frontier(fr);
    if (rc) return rc;
    if (!nodes || n_nodes < 0 || !counts || !ws) return RG_ERR_BAD_ARG;
    RgWorkspace w = rg_carve(ws, fr->n_query, fr->n_ent, 0);
    if (ws_bytes < w.dict_bytes) return RG_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    RG_CUDA_CALL(cudaMemsetAsync(fr->emask, 0, rg_frontier_emask_bytes(fr->n_query, fr->n_ent), st));
    RG_CUDA_CALL(cudaMemsetAsync(fr->dict, 0, rg_frontier_dict_bytes(fr->n_query, fr->n_ent), st));
    RG_CUDA_CALL(cudaMemsetAsync(counts, 0, RG_COUNTS_WORDS * sizeof(int64_t), st));
    if (n_nodes > 0) {
        k_set_nodes<<<(unsigned)rg_cdiv(n_nodes, kBlock), kBlock, 0, st>>>(
            nodes, n_nodes, fr->n_query, fr->n_ent, fr->emask, fr->dict, counts);
        RG_LAUNCH_CHECK();
    }
    return dict_prefix(fr, w, counts + RG_CNT_N_IN, st);
}

int rg_frontier_step(const rg_graph *g, const rg_frontier *in, rg_frontier *out, int64_t *counts,
                     void *ws, size_t ws_bytes, void *stream) {
    int rc = check_frontier(in);
    if (rc) return rc;
    rc = check_frontier(out);
    if (rc) return rc;
    if (!g || !g->head || !g->tail || g->n_fact <= 0 || !counts || !ws) return RG_ERR_BAD_ARG;
    if (in->n_query != out->n_query || in->n_ent != out->n_ent || in->n_ent != g->n_ent)
        return RG_ERR_BAD_ARG;
    if (in->emask == out->emask || in->dict == out->dict) return RG_ERR_BAD_ARG;
    if (g->n_fact >= (int64_t)INT32_MAX) return RG_ERR_TOO_LARGE;
    RgWorkspace w = rg_carve(ws, in->n_query, in->n_ent, g->n_fact);
    if (ws_bytes < w.total_bytes) return RG_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int Wn = rg_words_query(in->n_query), We = rg_words_ent(in->n_ent);
    const int64_t nbf = rg_cdiv(g->n_fact, RG_TILE);
    RG_CUDA_CALL(cudaMemsetAsync(out->emask, 0, rg_frontier_emask_bytes(out->n_query, out->n_ent), st));
    k_fact_count<<<(unsigned)nbf, kBlock, 0, st>>>(g->head, g->tail, g->n_fact, in->emask, out->emask, Wn,
                                                  w.fact_blocksum);
    RG_LAUNCH_CHECK();
    k_scan_blocksums<<<1, 1024, 0, st>>>(w.fact_blocksum, nbf, w.fact_blockprefix, counts + RG_CNT_E);
    RG_LAUNCH_CHECK();
    const int64_t n_task = (int64_t)We * Wn;
    k_transpose<<<(unsigned)rg_cdiv(n_task, kBlock / 32), kBlock, 0, st>>>(out->emask, out->n_ent, Wn,
                                                                          out->n_query, out->dict, We);
    RG_LAUNCH_CHECK();
    return dict_prefix(out, w, counts + RG_CNT_N_OUT, st);
}

int rg_frontier_nodes(const rg_frontier *fr, int64_t *nodes64, int32_t *node_b, int32_t *node_e,
                      void *stream) {
    int rc = check_frontier(fr);
    if (rc) return rc;
    const int We = rg_words_ent(fr->n_ent);
    const int64_t n_words = (int64_t)fr->n_query * We;
    k_emit_nodes<<<(unsigned)rg_cdiv(n_words, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
        fr->dict, fr->n_query, We, nodes64, node_b, node_e);
    RG_LAUNCH_CHECK();
    return RG_OK;
}

int rg_frontier_remap(const rg_frontier *in, const rg_frontier *out, int64_t *remap64, int32_t *remap32,
                      int32_t *inverse32, void *stream) {
    int rc = check_frontier(in);
    if (rc) return rc;
    rc = check_frontier(out);
    if (rc) return rc;
    if (in->n_query != out->n_query || in->n_ent != out->n_ent) return RG_ERR_BAD_ARG;
    const int64_t n_words = (int64_t)in->n_query * rg_words_ent(in->n_ent);
    k_remap<<<(unsigned)rg_cdiv(n_words, kBlock), kBlock, 0, (cudaStream_t)stream>>>(in->dict, out->dict,
                                                                                    n_words, remap64, remap32,
                                                                                    inverse32);
    RG_LAUNCH_CHECK();
    return RG_OK;
}

int rg_edges_emit(const rg_graph *g, const rg_frontier *in, const rg_frontier *out, const void *ws,
                  size_t ws_bytes, int64_t n_edges, int64_t *edges, void *stream) {
    int rc = check_frontier(in);
    if (rc) return rc;
    rc = check_frontier(out);
    if (rc) return rc;
    if (!g || !g->head || !g->rel || !g->tail || g->n_fact <= 0 || !ws || n_edges < 0) return RG_ERR_BAD_ARG;
    if (n_edges >= (int64_t)INT32_MAX) return RG_ERR_TOO_LARGE;
    if (n_edges == 0) return RG_OK;
    if (!edges) return RG_ERR_BAD_ARG;
    RgWorkspace w = rg_carve(const_cast<void *>(ws), in->n_query, in->n_ent, g->n_fact);
    if (ws_bytes < w.total_bytes) return RG_ERR_WORKSPACE;
    const int64_t nbf = rg_cdiv(g->n_fact, RG_TILE);
    const int Wn = rg_words_query(in->n_query);
    const dim3 grid((unsigned)nbf, kEmitParts);
    auto kern = Wn <= 2 ? k_emit_edges<true> : k_emit_edges<false>;
    kern<<<grid, kBlock, 0, (cudaStream_t)stream>>>(g->head, g->rel, g->tail, g->n_fact, in->emask, Wn, in->dict,
                                                    out->dict, rg_words_ent(in->n_ent), in->qinfo, out->qinfo,
                                                    in->n_query, in->n_ent, w.fact_blockprefix, edges);
    RG_LAUNCH_CHECK();
    return RG_OK;
}

// ---- the public get_neighbors hop as TWO calls (one count read-back in between) -----------------------
// Driven from Python every C call costs ~10 us of host time; a hop was five calls.
int rg_get_neighbors_expand(const rg_graph *g, const int64_t *nodes, int64_t n_nodes, rg_frontier *in, rg_frontier *out,
                            int64_t *counts_in, int64_t *counts_out, void *ws, size_t ws_bytes, void *stream) {
    int rc = rg_frontier_from_nodes(nodes, n_nodes, in, counts_in, ws, ws_bytes, stream);
    if (rc) return rc;
    return rg_frontier_step(g, in, out, counts_out, ws, ws_bytes, stream);
}

int rg_get_neighbors_emit(const rg_graph *g, const rg_frontier *in, const rg_frontier *out, const void *ws,
                          size_t ws_bytes, int64_t n_edges, int64_t *tail_nodes, int64_t *edges,
                          int64_t *old_nodes_new_idx, void *stream) {
    int rc = rg_frontier_nodes(out, tail_nodes, nullptr, nullptr, stream);
    if (rc) return rc;
    rc = rg_frontier_remap(in, out, old_nodes_new_idx, nullptr, nullptr, stream);
    if (rc) return rc;
    return rg_edges_emit(g, in, out, ws, ws_bytes, n_edges, edges, stream);
}

}  // extern "C"
