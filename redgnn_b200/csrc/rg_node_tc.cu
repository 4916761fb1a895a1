// Fused node update on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
// Same contract as k_node_update in rg_node.cu (reference Static/transductive/models.py:41,81-86:
// x = act(W_h agg), h0 re-index, single-step GRU, next layer's Ws_attn, W_final).  The three dense
// contractions per node ([1xD].[DxD], [1xD].[Dx3D] twice = 16.5 kFMA at D=48) are the only real
// GEMMs on the path; on CUDA cores they cost as much as the whole edge kernel.  Here one CTA owns a
// tile of 128 nodes = the 128 TMEM lanes:
//   * operands are staged in shared memory in the canonical K-major no-swizzle UMMA layout
//     (8-row x 16-byte core matrices), each fp32 value split into hi = top 19 bits (what kind::tf32
//     reads) and lo = x - hi, so that  x.w ~= hi.hi + hi.lo + lo.hi  (3xTF32) keeps fp32-level
//     accuracy (~2^-19 relative) -- plain TF32 would break the 1e-4 parity bound;
//   * one elected thread issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=D/2D/3D, K=8 per
//     instruction) with fp32 accumulators in TMEM columns [x | r | z | i_n | h_n];
//   * tcgen05.commit -> mbarrier; 128 x D/16 threads then run the epilogues: warp w owns TMEM lanes
//     32*(w%4).. (= node rows) and the 16-column slice w/4, read with tcgen05.ld.32x32b.x16, apply
//     the gates and write hidden; the as8 / score dot products are combined through spare TMEM
//     columns.  (Splitting the columns over D/16 warps per lane quarter triples the warps that hide
//     the epilogue latency; the tile itself stays 128 rows because of the 227 KB budget.)
//   * the 227 KB budget leaves room for ONE set of operand tiles, so the pipeline is in registers:
//     while GEMM 2 of a tile runs, every thread gathers its slice of the NEXT tile (agg row and the
//     re-indexed h_prev row; the re-index entry itself is fetched two tiles ahead) and writes it
//     into the operand tiles as soon as GEMM 2 has consumed them.
#include "rg_tc.cuh"

namespace {

using namespace rgtc;

template <int D>
struct TcSmem {
    static constexpr int KC = D / 4;                 // 16-byte K chunks per row
    static constexpr int SBO = KC * 128;             // bytes between 8-row groups
    static constexpr int WROWS = 7 * D;              // W_h (D) | W_ih (3D) | W_hh (3D)
    static constexpr int W_BYTES = WROWS * D * 4;    // one precision part
    static constexpr int A_BYTES = kTcRows * D * 4;  // one precision part of one activation tile
    // byte offsets (all multiples of 128)
    static constexpr int W_HI = 0;
    static constexpr int W_LO = W_HI + W_BYTES;
    static constexpr int A_HI = W_LO + W_BYTES;   // agg, then x
    static constexpr int A_LO = A_HI + A_BYTES;
    static constexpr int H_HI = A_LO + A_BYTES;   // h0
    static constexpr int H_LO = H_HI + A_BYTES;
    static constexpr int WS = H_LO + A_BYTES;     // float [9][D]
    static constexpr int BIAS = WS + 9 * D * 4;   // float brz[2D], bin[D], bhn[D]
    static constexpr int BAR = BIAS + 4 * D * 4;  // uint64 mbarrier, uint32 tmem base
    static constexpr int TOTAL = BAR + 16;
    // canonical offset of (row, 16-byte chunk) inside a tile
    __host__ __device__ static constexpr int off(int row, int chunk) {
        return ((row >> 3) * KC + chunk) * 128 + (row & 7) * 16;
    }
};

__device__ __forceinline__ float act_apply_tc(float v, int act) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return tanhf(v);
    return v;
}
// ex2.approx-based forms: relative error ~2^-21, far inside the 1e-4 parity bound, and ~3x fewer
// instructions than expf / tanhf in the latency-critical epilogue
__device__ __forceinline__ float sigmoid_tc(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_tc(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }

// issue the 3xTF32 product  D[tmem_d : N cols] (+)= A(rows 0..127) . B(rows b_row0 .. b_row0+N)^T
// (a fourth lo.lo term was measured and does not help: on the hub rows of the power-law shape the
// remaining error comes from the accumulation inside the tensor core, not from the dropped term)
template <int D>
__device__ __forceinline__ void issue_gemm(uint32_t smem_base, int a_hi, int a_lo, int b_row0, int N, uint32_t tmem_d,
                                           bool accumulate_first) {
    using L = TcSmem<D>;
    const uint32_t idesc = make_idesc(kTcRows, N);
    const uint32_t boff = (uint32_t)(b_row0 >> 3) * L::SBO;
    uint32_t acc = accumulate_first ? 1u : 0u;
#pragma unroll
    for (int ks = 0; ks < D / 8; ++ks) {
        const uint32_t koff = ks * 256;  // two 16-byte chunks = 8 tf32 values
        const uint64_t ah = make_desc(smem_base + a_hi + koff, 128, L::SBO);
        const uint64_t al = make_desc(smem_base + a_lo + koff, 128, L::SBO);
        const uint64_t bh = make_desc(smem_base + L::W_HI + boff + koff, 128, L::SBO);
        const uint64_t bl = make_desc(smem_base + L::W_LO + boff + koff, 128, L::SBO);
        tc_mma_tf32(tmem_d, ah, bh, idesc, acc);
        tc_mma_tf32(tmem_d, ah, bl, idesc, 1u);
        tc_mma_tf32(tmem_d, al, bh, idesc, 1u);
        acc = 1u;
    }
}

template <int D, bool HAS_H0>
__global__ void __launch_bounds__(128 * (D / 16), 1) k_node_update_tc(
    const float *__restrict__ agg, const float *__restrict__ h_prev, const int32_t *__restrict__ src,
    const float *__restrict__ W_h, const float *__restrict__ W_ih, const float *__restrict__ W_hh,
    const float *__restrict__ b_ih, const float *__restrict__ b_hh, const float *__restrict__ Ws_next,
    const float *__restrict__ W_final, int act, int64_t n_nodes_host, const int64_t *__restrict__ n_nodes_dev,
    float *__restrict__ hidden, float *__restrict__ as8, float *__restrict__ score,
    const float *__restrict__ drop_mask, float *__restrict__ saved, int ws_rows) {
    extern __shared__ __align__(1024) uint8_t smem[];
    using L = TcSmem<D>;
    constexpr int KC = L::KC;
    constexpr uint32_t kTmemCols = (6 * D <= 128) ? 128 : (6 * D <= 256 ? 256 : 512);  // 5D accumulators + D scratch
    constexpr int kTcThreads = 128 * (D / 16);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int trow = (warp & 3) * 32 + (tid & 31);  // node row of the tile == TMEM lane
    const int cq = warp >> 2;                       // 16-column slice owned by this thread
    const int64_t n_nodes = n_nodes_dev ? *n_nodes_dev : n_nodes_host;
    const int64_t n_tiles = (n_nodes + kTcRows - 1) / kTcRows;
    if ((int64_t)blockIdx.x >= n_tiles) return;  // block-uniform: nothing allocated yet

    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + L::BAR);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::BAR + 8);
    float *ws = reinterpret_cast<float *>(smem + L::WS);
    float *bias = reinterpret_cast<float *>(smem + L::BIAS);
    const uint32_t smem_base = smem_u32(smem);

    // ---- one-time setup: barrier, TMEM, weights (hi / lo parts, canonical layout) ----
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, kTmemCols);
    for (int i = tid; i < L::WROWS * KC; i += kTcThreads) {
        const int row = i / KC, ch = i % KC;  // row: 0..D-1 W_h, D..4D-1 W_ih, 4D..7D-1 W_hh
        const float *srcw = row < D ? W_h + (size_t)row * D : (row < 4 * D ? W_ih + (size_t)(row - D) * D
                                                                           : W_hh + (size_t)(row - 4 * D) * D);
        float4 x = __ldg(reinterpret_cast<const float4 *>(srcw) + ch), hi, lo;
        split_tf32(x, hi, lo);
        *reinterpret_cast<float4 *>(smem + L::W_HI + L::off(row, ch)) = hi;
        *reinterpret_cast<float4 *>(smem + L::W_LO + L::off(row, ch)) = lo;
    }
    for (int i = tid; i < 9 * D; i += kTcThreads) {
        float v = 0.f;
        if (i < 8 * D) {
            if (Ws_next && i < ws_rows * D) v = Ws_next[i];   // rows >= ws_rows (attn_dim) are zero
        } else if (W_final) {
            v = W_final[i - 8 * D];
        }
        ws[i] = v;
    }
    for (int i = tid; i < 2 * D; i += kTcThreads) bias[i] = b_ih[i] + b_hh[i];
    for (int i = tid; i < D; i += kTcThreads) {
        bias[2 * D + i] = b_ih[2 * D + i];
        bias[3 * D + i] = b_hh[2 * D + i];
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // TMEM columns: [0,D) x | [D,2D) r | [2D,3D) z | [3D,4D) i_n | [4D,5D) h_n
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t phase = 0;

    // this thread's slice (node row trow, 16 columns from c0) of a tile: hi / lo parts into the operand tiles.
    // (Measured and rejected: a separate, row-coalesced thread mapping for this staging -- thread t takes chunk
    // t % KC of rows t / KC + 32 k -- makes the global loads contiguous but puts the 12 chunk writes of a row
    // 128 B apart in the K-major operand layout = on the same shared-memory banks: node update 1.94 -> 2.43 ms.)
    const int c0 = 16 * cq;
    auto stage_slice = [&](const float4 (&a)[4], const float4 (&h)[4]) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float4 hi, lo;
            split_tf32(a[q], hi, lo);
            *reinterpret_cast<float4 *>(smem + L::A_HI + L::off(trow, 4 * cq + q)) = hi;
            *reinterpret_cast<float4 *>(smem + L::A_LO + L::off(trow, 4 * cq + q)) = lo;
            if (HAS_H0) {
                split_tf32(h[q], hi, lo);
                *reinterpret_cast<float4 *>(smem + L::H_HI + L::off(trow, 4 * cq + q)) = hi;
                *reinterpret_cast<float4 *>(smem + L::H_LO + L::off(trow, 4 * cq + q)) = lo;
            }
        }
    };
    auto load_slice = [&](int64_t r, bool ok, int s, float4 (&a)[4], float4 (&h)[4]) {
        const float4 *pa = reinterpret_cast<const float4 *>(agg + (size_t)(ok ? r : 0) * D);
        const float4 *ph = reinterpret_cast<const float4 *>(h_prev + (size_t)(s >= 0 ? s : 0) * D);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            a[q] = ok ? __ldg(pa + 4 * cq + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (HAS_H0) h[q] = (s >= 0) ? __ldg(ph + 4 * cq + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto src_of = [&](int64_t r) -> int { return (HAS_H0 && r < n_nodes) ? __ldg(src + r) : -1; };

    // ---- prologue: stage the first tile; the re-index entry of the following tile is fetched one tile ahead
    //      so that the dependent h_prev[src] gather of the prefetch below starts without a round trip ----
    int s_next;
    {
        const int64_t row = (int64_t)blockIdx.x * kTcRows + trow;
        float4 a[4], h[4];
        load_slice(row, row < n_nodes, src_of(row), a, h);
        s_next = src_of(row + (int64_t)gridDim.x * kTcRows);
        stage_slice(a, h);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row = tile * kTcRows + trow;
        const bool live = row < n_nodes;
        const int64_t nrow = row + (int64_t)gridDim.x * kTcRows;  // this thread's row in the CTA's next tile
        const bool has_next = tile + gridDim.x < n_tiles;          // block-uniform

        // ---- GEMM 1: x = agg . W_h^T  -> TMEM [0, D) ----
        if (tid == 0) {
            tc_fence_after();
            issue_gemm<D>(smem_base, L::A_HI, L::A_LO, 0, D, tmem_base, false);
            tc_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        // epilogue 1: x = act(.) back to shared memory (overwrites the agg tile) as hi / lo
        {
            float v[16];
            tmem_ld16(t_lane + c0, v);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float4 x = make_float4(act_apply_tc(v[4 * q], act), act_apply_tc(v[4 * q + 1], act),
                                       act_apply_tc(v[4 * q + 2], act), act_apply_tc(v[4 * q + 3], act)),
                       hi, lo;
                if (live) {  // training: keep x_act for the backward pass, apply the (scaled) dropout mask
                    const size_t o = (size_t)row * D + c0 + 4 * q;
                    if (saved) *reinterpret_cast<float4 *>(saved + il_off(row, c0 / 4 + q, KC)) = x;
                    if (drop_mask) {
                        const float4 mk = __ldg(reinterpret_cast<const float4 *>(drop_mask + o));
                        x = make_float4(x.x * mk.x, x.y * mk.y, x.z * mk.z, x.w * mk.w);
                    }
                }
                split_tf32(x, hi, lo);
                *reinterpret_cast<float4 *>(smem + L::A_HI + L::off(trow, c0 / 4 + q)) = hi;
                *reinterpret_cast<float4 *>(smem + L::A_LO + L::off(trow, c0 / 4 + q)) = lo;
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();

        // ---- GEMM 2: [r z i_n] = x . W_ih^T ; [r z] += h0 . W_hh[r,z]^T ; h_n = h0 . W_hh[n]^T ----
        if (tid == 0) {
            tc_fence_after();
            issue_gemm<D>(smem_base, L::A_HI, L::A_LO, D, 3 * D, tmem_base + D, false);
            if (HAS_H0) {
                issue_gemm<D>(smem_base, L::H_HI, L::H_LO, 4 * D, 2 * D, tmem_base + D, true);
                issue_gemm<D>(smem_base, L::H_HI, L::H_LO, 6 * D, D, tmem_base + 4 * D, false);
            }
            tc_commit(bar);
        }
        // prefetch the next tile's rows into registers: the gather latency hides behind GEMM 2
        float4 na[4], nh[4];
        load_slice(nrow, has_next && nrow < n_nodes, has_next ? s_next : -1, na, nh);
        s_next = src_of(nrow + (int64_t)gridDim.x * kTcRows);
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        // GEMM 2 has consumed both operand tiles: take this thread's h0 slice out of the h0 tile, then
        // overwrite the same positions of both tiles with the next tile's rows
        float h0v[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float4 hh = make_float4(0.f, 0.f, 0.f, 0.f), hl = hh;
            if (HAS_H0) {
                hh = *reinterpret_cast<const float4 *>(smem + L::H_HI + L::off(trow, c0 / 4 + q));
                hl = *reinterpret_cast<const float4 *>(smem + L::H_LO + L::off(trow, c0 / 4 + q));
            }
            h0v[4 * q] = hh.x + hl.x; h0v[4 * q + 1] = hh.y + hl.y;
            h0v[4 * q + 2] = hh.z + hl.z; h0v[4 * q + 3] = hh.w + hl.w;
        }
        if (has_next) stage_slice(na, nh);
        // epilogue 2: gates and the hidden slice; attention projection / score partials
        {
            float vr[16], vz[16], vi[16], vh[16], hn[16];
            tmem_ld16(t_lane + D + c0, vr);
            tmem_ld16(t_lane + 2 * D + c0, vz);
            tmem_ld16(t_lane + 3 * D + c0, vi);
            if (HAS_H0) tmem_ld16(t_lane + 4 * D + c0, vh);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float sv[4][4];  // r, z, n, W_hn h0 + b_hn of these 4 columns (training only)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = 4 * q + e, c = c0 + j;
                    const float rg = sigmoid_tc(vr[j] + bias[c]);
                    const float zg = sigmoid_tc(vz[j] + bias[D + c]);
                    const float hlin = (HAS_H0 ? vh[j] : 0.f) + bias[3 * D + c];
                    const float ng = tanh_tc(vi[j] + bias[2 * D + c] + rg * hlin);
                    hn[j] = (1.0f - zg) * ng + zg * h0v[j];
                    sv[0][e] = rg; sv[1][e] = zg; sv[2][e] = ng; sv[3][e] = hlin;
                }
                if (saved && live) {  // planes 1..5 of saved[6][il_plane_floats(n, D)] (lane-interleaved, rg_tc.cuh)
                    const size_t plane = il_plane_floats(n_nodes_host, D), o = il_off(row, c0 / 4 + q, KC);
#pragma unroll
                    for (int pl = 0; pl < 4; ++pl)
                        *reinterpret_cast<float4 *>(saved + (pl + 1) * plane + o) =
                            make_float4(sv[pl][0], sv[pl][1], sv[pl][2], sv[pl][3]);
                    *reinterpret_cast<float4 *>(saved + 5 * plane + o) =
                        make_float4(h0v[4 * q], h0v[4 * q + 1], h0v[4 * q + 2], h0v[4 * q + 3]);
                }
            }
            if (live) {
                float4 *po = reinterpret_cast<float4 *>(hidden + (size_t)row * D + c0);
#pragma unroll
                for (int q = 0; q < 4; ++q) po[q] = make_float4(hn[4 * q], hn[4 * q + 1], hn[4 * q + 2], hn[4 * q + 3]);
            }
            if (as8 || score) {
                // partial dot products over this thread's 16 columns, exchanged through spare TMEM columns
                // [5D + 16 cq, +16) of the row's lane (the operand tiles already hold the next tile) and
                // summed in slice order by the cq == 0 thread of the row
                float part[16];
#pragma unroll
                for (int a = 0; a < 16; ++a) part[a] = 0.f;
#pragma unroll
                for (int a = 0; a < 9; ++a) {
                    float sacc = 0.f;
#pragma unroll
                    for (int j = 0; j < 16; ++j) sacc = fmaf(hn[j], ws[a * D + c0 + j], sacc);
                    part[a] = sacc;
                }
                if (cq != 0) tmem_st16(t_lane + 5 * D + c0, part);
                tc_fence_before();
                __syncthreads();
                tc_fence_after();
                if (cq == 0) {
#pragma unroll
                    for (int k = 1; k < D / 16; ++k) {
                        float other[16];
                        tmem_ld16(t_lane + 5 * D + 16 * k, other);
#pragma unroll
                        for (int a = 0; a < 9; ++a) part[a] += other[a];
                    }
                    if (live) {
                        if (as8) {
                            float4 *po = reinterpret_cast<float4 *>(as8 + (size_t)row * 8);
                            po[0] = make_float4(part[0], part[1], part[2], part[3]);
                            po[1] = make_float4(part[4], part[5], part[6], part[7]);
                        }
                        if (score) score[row] = part[8];
                    }
                }
            }
        }
        // TMEM reads of this tile are done and the staged rows of the next tile are visible to the
        // async proxy before the next GEMM 1 is issued
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
    }
    if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

template <int D, bool HH>
int launch_node_tc(const float *agg, const float *h_prev, const int32_t *src, const float *W_h, const float *W_ih,
                   const float *W_hh, const float *b_ih, const float *b_hh, const float *Ws_next,
                   const float *W_final, int act, int64_t n_nodes, const int64_t *n_nodes_dev, float *hidden,
                   float *as8, float *score, const float *drop_mask, float *saved, int ws_rows, cudaStream_t st) {
    constexpr size_t smem = TcSmem<D>::TOTAL;
    static_assert(smem <= 232448, "tile does not fit the 227 KB shared memory of one CTA");
    auto kern = k_node_update_tc<D, HH>;
    RG_CUDA_CALL(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, n_sm = 148;
    RG_CUDA_CALL(cudaGetDevice(&dev));
    RG_CUDA_CALL(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    const int64_t n_tiles = (n_nodes + kTcRows - 1) / kTcRows;
    const int grid = (int)(n_tiles < n_sm ? n_tiles : n_sm);  // persistent: one CTA per SM
    kern<<<grid, 128 * (D / 16), smem, st>>>(agg, h_prev, src, W_h, W_ih, W_hh, b_ih, b_hh, Ws_next, W_final, act,
                                         n_nodes, n_nodes_dev, hidden, as8, score, drop_mask, saved, ws_rows);
    RG_LAUNCH_CHECK();
    return RG_OK;
}

}  // namespace

// internal entry used by rg_node_update (rg_node.cu); returns RG_ERR_UNSUPPORTED when D has no
// tensor-core instantiation (the tile must fit 227 KB of shared memory: D <= 48)
int rg_node_update_tc(int32_t hidden_dim, int64_t n_nodes, const int64_t *n_nodes_dev, const float *agg,
                      const float *h_prev, const int32_t *src, const float *W_h, const float *W_ih,
                      const float *W_hh, const float *b_ih, const float *b_hh, const float *Ws_next,
                      const float *W_final, int32_t act, float *hidden, float *as8, float *score,
                      const float *drop_mask, float *saved, int32_t ws_rows, cudaStream_t st) {
#define RG_NODE_TC(DD)                                                                                             \
    return h_prev ? launch_node_tc<DD, true>(agg, h_prev, src, W_h, W_ih, W_hh, b_ih, b_hh, Ws_next, W_final, act,  \
                                             n_nodes, n_nodes_dev, hidden, as8, score, drop_mask, saved, ws_rows, st)       \
                  : launch_node_tc<DD, false>(agg, h_prev, src, W_h, W_ih, W_hh, b_ih, b_hh, Ws_next, W_final, act, \
                                              n_nodes, n_nodes_dev, hidden, as8, score, drop_mask, saved, ws_rows, st)
    switch (hidden_dim) {
        case 16: RG_NODE_TC(16);
        case 32: RG_NODE_TC(32);
        case 48: RG_NODE_TC(48);
        default: return RG_ERR_UNSUPPORTED;
    }
#undef RG_NODE_TC
}
