// Dense part of the TRAINING BACKWARD of the node update, native (no library GEMMs):
// autograd of  hidden = GRU(dropout(act(W_h agg)), h0)  (reference Static/transductive/models.py:41,
// 81-84; triggered at base_model.py:61).
//
//   rg_node_bwd    (k_node_bwd_tc, tcgen05 + TMEM, 128-node tiles like the forward kernel):
//       g   = g_hidden [+ g_small . w_small] [+ g_h0_next[remap]]              upstream gradient
//       g_n' = g (1-z)(1-n^2);  g_z' = g (h0-n) z (1-z);  g_r' = g_n' hl r (1-r)   GRU gate pre-activations
//       G4  = [g_r' | g_z' | g_n' | g_n' r]                                     (kept for the weight gradients)
//       g_x = [g_r' g_z' g_n'] . W_ih       g_h0 = g z + [g_r' g_z' g_n' r] . W_hh
//       g_pre = g_x * mask * act'(x)        g_agg = g_pre . W_h
//     The gate gradients never touch shared memory: every thread writes its (row, 16-column) slice as
//     hi / lo TF32 parts straight into TMEM with tcgen05.st, and the MMAs take their A operand FROM
//     TMEM (tcgen05.mma ... [d], [a], b-desc: lane = node row, 8 consecutive columns = one K slice);
//     only the transposed weights (B operands, hi / lo, K-major canonical layout) live in shared
//     memory.  3xTF32 (hi.hi + hi.lo + lo.hi) as in the forward kernel keeps fp32-level accuracy.
//   rg_node_wgrad  (k_node_wgrad, CUDA cores, exact fp32, deterministic): the reductions over NODES
//       dW_ih = G[r,z,n]^T x_in   dW_hh = G[r,z,nr]^T h0   dW_h = g_pre^T agg   dW_small = g_small^T hidden
//       bias sums = column sums of G4
//     as 8x8 register tiles over 32-node slabs staged in shared memory by TMA bulk copies (mbarrier
//     ring); per-CTA partials are summed in a fixed order by k_wgrad_reduce (no atomics).
// saved / G4 / g_pre use the lane-interleaved plane layout of rg_tc.cuh (coalesced for lane == row).
// Both stop at the device-side node count, so upper-bound (shape-static) buffers cost nothing.
#include "rg_tc.cuh"

namespace {

using namespace rgtc;

template <int D>
struct BwdSmem {
    static constexpr int KC = D / 4;               // 16-byte K chunks per row
    static constexpr int SBO = KC * 128;           // bytes between 8-row groups
    static constexpr int WROWS = 7 * D;            // T_r (2D) | T_z (2D) | T_n (D) | T_nr (D) | T_h (D)
    static constexpr int W_BYTES = WROWS * D * 4;  // one precision part
    static constexpr int W_HI = 0;
    static constexpr int W_LO = W_BYTES;
    static constexpr int WSM = 2 * W_BYTES;        // float w_small[8][D]
    static constexpr int BAR = WSM + 8 * D * 4;    // uint64 mbarrier, uint32 tmem base
    static constexpr int TOTAL = BAR + 16;
    __host__ __device__ static constexpr int off(int row, int chunk) {
        return ((row >> 3) * KC + chunk) * 128 + (row & 7) * 16;
    }
};

__device__ __forceinline__ void tmem_st16_nowait(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        :
        : "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
          "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
          "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
          "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
          "r"(__float_as_uint(v[15]))
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// hi / lo TF32 parts of 16 values into two TMEM column ranges of this thread's lane
__device__ __forceinline__ void tmem_put_split(uint32_t t_hi, uint32_t t_lo, const float (&v)[16]) {
    float hi[16], lo[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        hi[i] = tf32_hi(v[i]);
        lo[i] = v[i] - hi[i];
    }
    tmem_st16_nowait(t_hi, hi);
    tmem_st16_nowait(t_lo, lo);
}

__device__ __forceinline__ void ld16(const float *p, float (&v)[16]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 x = __ldg(reinterpret_cast<const float4 *>(p) + q);
        v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
    }
}
// this thread's 16-column slice (chunks 4cq .. 4cq+3) of a row of a lane-interleaved plane
__device__ __forceinline__ void ld16_il(const float *plane, int64_t row, int cq, int KC, float (&v)[16]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 x = __ldg(reinterpret_cast<const float4 *>(plane + il_off(row, 4 * cq + q, KC)));
        v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
    }
}
__device__ __forceinline__ void st16_il(float *plane, int64_t row, int cq, int KC, const float (&v)[16]) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
        *reinterpret_cast<float4 *>(plane + il_off(row, 4 * cq + q, KC)) =
            make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}
__device__ __forceinline__ void st16(float *p, const float (&v)[16]) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
        reinterpret_cast<float4 *>(p)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

// 3xTF32 product with the A operand in TMEM:  D[tmem_d : N cols] (+)= A(cols a_hi.., a_lo..) . B(rows b_row0..)^T
template <int D>
__device__ __forceinline__ void issue_gemm_ts(uint32_t smem_base, uint32_t tmem_a_hi, uint32_t tmem_a_lo, int b_row0,
                                              int N, uint32_t tmem_d, uint32_t &acc) {
    using L = BwdSmem<D>;
    const uint32_t idesc = make_idesc(kTcRows, N);
    const uint32_t boff = (uint32_t)(b_row0 >> 3) * L::SBO;
#pragma unroll
    for (int ks = 0; ks < D / 8; ++ks) {
        const uint32_t koff = ks * 256;  // two 16-byte chunks = 8 tf32 values
        const uint64_t bh = make_desc(smem_base + L::W_HI + boff + koff, 128, L::SBO);
        const uint64_t bl = make_desc(smem_base + L::W_LO + boff + koff, 128, L::SBO);
        tc_mma_tf32_ts(tmem_d, tmem_a_hi + 8 * ks, bh, idesc, acc);
        tc_mma_tf32_ts(tmem_d, tmem_a_hi + 8 * ks, bl, idesc, 1u);
        tc_mma_tf32_ts(tmem_d, tmem_a_lo + 8 * ks, bh, idesc, 1u);
        acc = 1u;
    }
}

template <int D, bool HAS_H0>
__global__ void __launch_bounds__(128 * (D / 16), 1) k_node_bwd_tc(
    const float *__restrict__ g_hidden, const float *__restrict__ g_small, int g_small_stride,
    const float *__restrict__ w_small, int w_small_rows, const float *__restrict__ g_h0_next, const int32_t *__restrict__ remap,
    const float *__restrict__ saved,
    int64_t plane_rows, const float *__restrict__ drop_mask, const float *__restrict__ W_h,
    const float *__restrict__ W_ih, const float *__restrict__ W_hh, int act, int64_t n_nodes_host,
    const int64_t *__restrict__ n_nodes_dev, float *__restrict__ G4, float *__restrict__ g_pre_out,
    float *__restrict__ g_agg, float *__restrict__ g_h0) {
    extern __shared__ __align__(1024) uint8_t smem[];
    using L = BwdSmem<D>;
    constexpr int KC = L::KC;
    constexpr uint32_t kTmemCols = (10 * D <= 256) ? 256 : 512;
    constexpr int kThreads = 128 * (D / 16);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int trow = (warp & 3) * 32 + (tid & 31);  // node row of the tile == TMEM lane
    const int cq = warp >> 2, c0 = 16 * cq;         // 16-column slice owned by this thread
    const int64_t n_nodes = n_nodes_dev ? *n_nodes_dev : n_nodes_host;
    const int64_t n_tiles = (n_nodes + kTcRows - 1) / kTcRows;
    if ((int64_t)blockIdx.x >= n_tiles) return;  // block-uniform: nothing allocated yet

    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + L::BAR);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::BAR + 8);
    float *wsm = reinterpret_cast<float *>(smem + L::WSM);
    const uint32_t smem_base = smem_u32(smem);

    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, kTmemCols);
    // transposed weights as B operands (row = output column n, K = gate / feature index k), hi / lo parts:
    // 128-bit loads along n (contiguous in the source), scalar stores into the K-major canonical layout
    for (int i = tid; i < L::WROWS * D / 4; i += kThreads) {
        const int k = i % D, r4 = (i / D) * 4;   // rows r4 .. r4+3 of the B tile, K index k
        const float *srcw;
        if (r4 < 4 * D) {  // T_r, T_z: n < D -> W_ih[gate*D + k][n], else W_hh[gate*D + k][n - D]
            const int gate = r4 / (2 * D), n = r4 % (2 * D);
            srcw = n < D ? W_ih + (size_t)(gate * D + k) * D + n : W_hh + (size_t)(gate * D + k) * D + (n - D);
        } else if (r4 < 5 * D) {
            srcw = W_ih + (size_t)(2 * D + k) * D + (r4 - 4 * D);
        } else if (r4 < 6 * D) {
            srcw = W_hh + (size_t)(2 * D + k) * D + (r4 - 5 * D);
        } else {
            srcw = W_h + (size_t)k * D + (r4 - 6 * D);
        }
        const float4 x = __ldg(reinterpret_cast<const float4 *>(srcw));
        const float v[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float hi = tf32_hi(v[e]);
            const int o = L::off(r4 + e, k >> 2) + (k & 3) * 4;
            *reinterpret_cast<float *>(smem + L::W_HI + o) = hi;
            *reinterpret_cast<float *>(smem + L::W_LO + o) = v[e] - hi;
        }
    }
    for (int i = tid; i < 8 * D; i += kThreads) wsm[i] = (w_small && i < w_small_rows * D) ? w_small[i] : 0.f;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    // TMEM columns: gate operands [2g D, 2g D + D) hi, [2g D + D, 2g D + 2D) lo for g = r, z, n, nr;
    // accumulators g_x [8D, 9D), g_h0 [9D, 10D); second GEMM: g_pre hi/lo in [0, 2D), g_agg in [2D, 3D)
    uint32_t phase = 0;
    const size_t plane = il_plane_floats(plane_rows, D);      // saved planes, G4 planes: lane-interleaved
    const size_t gplane = il_plane_floats(n_nodes_host, D);

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row = tile * kTcRows + trow;
        const bool live = row < n_nodes;
        const size_t o = (size_t)(live ? row : 0) * D + c0;
        float gh0d[16];
        {
            float g[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) g[j] = 0.f;
            if (live) {
                if (g_hidden) ld16(g_hidden + o, g);
                if (g_small) {
                    const float4 s0 = __ldg(reinterpret_cast<const float4 *>(g_small + (size_t)row * g_small_stride));
                    const float4 s1 = __ldg(reinterpret_cast<const float4 *>(g_small + (size_t)row * g_small_stride) + 1);
                    const float s[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
                    for (int a = 0; a < 8; ++a)
#pragma unroll
                        for (int j = 0; j < 16; ++j) g[j] = fmaf(s[a], wsm[a * D + c0 + j], g[j]);
                }
                if (g_h0_next) {
                    float t[16];
                    ld16(g_h0_next + (size_t)__ldg(remap + row) * D + c0, t);
#pragma unroll
                    for (int j = 0; j < 16; ++j) g[j] += t[j];
                }
            }
            float r[16], z[16], nn[16], hl[16], h0[16];
            if (live) {
                ld16_il(saved + plane, row, cq, KC, r);
                ld16_il(saved + 2 * plane, row, cq, KC, z);
                ld16_il(saved + 3 * plane, row, cq, KC, nn);
                ld16_il(saved + 4 * plane, row, cq, KC, hl);
                ld16_il(saved + 5 * plane, row, cq, KC, h0);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) r[j] = z[j] = nn[j] = hl[j] = h0[j] = 0.f;
            }
            float grp[16], gzp[16], gnp[16], gnr[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                gnp[j] = g[j] * (1.f - z[j]) * (1.f - nn[j] * nn[j]);
                gzp[j] = g[j] * (h0[j] - nn[j]) * z[j] * (1.f - z[j]);
                grp[j] = gnp[j] * hl[j] * r[j] * (1.f - r[j]);
                gnr[j] = gnp[j] * r[j];
                gh0d[j] = g[j] * z[j];
            }
            if (live) {   // G4: four lane-interleaved planes [4][il_rows(n)][D]
                st16_il(G4, row, cq, KC, grp);
                st16_il(G4 + gplane, row, cq, KC, gzp);
                st16_il(G4 + 2 * gplane, row, cq, KC, gnp);
                st16_il(G4 + 3 * gplane, row, cq, KC, gnr);
            }
            tmem_put_split(t_lane + 0 * D + c0, t_lane + 1 * D + c0, grp);
            tmem_put_split(t_lane + 2 * D + c0, t_lane + 3 * D + c0, gzp);
            tmem_put_split(t_lane + 4 * D + c0, t_lane + 5 * D + c0, gnp);
            if (HAS_H0) tmem_put_split(t_lane + 6 * D + c0, t_lane + 7 * D + c0, gnr);
            tmem_wait_st();
        }
        tc_fence_before();
        __syncthreads();

        // ---- GEMM batch 1: g_x = G[r,z,n] . W_ih ; g_h0 = G[r,z,nr] . W_hh ----
        if (tid == 0) {
            tc_fence_after();
            uint32_t acc = 0u;
            constexpr int N2 = HAS_H0 ? 2 * D : D;
            issue_gemm_ts<D>(smem_base, tmem_base + 0 * D, tmem_base + 1 * D, 0, N2, tmem_base + 8 * D, acc);
            issue_gemm_ts<D>(smem_base, tmem_base + 2 * D, tmem_base + 3 * D, 2 * D, N2, tmem_base + 8 * D, acc);
            issue_gemm_ts<D>(smem_base, tmem_base + 4 * D, tmem_base + 5 * D, 4 * D, D, tmem_base + 8 * D, acc);
            if (HAS_H0) issue_gemm_ts<D>(smem_base, tmem_base + 6 * D, tmem_base + 7 * D, 5 * D, D, tmem_base + 9 * D, acc);
            tc_commit(bar);
        }
        // while the MMAs run: this thread's slice of x_act / dropout mask for the epilogue
        float xa[16], mk[16];
        if (live) {
            ld16_il(saved, row, cq, KC, xa);
            if (drop_mask) ld16(drop_mask + o, mk);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        {
            float v[16];
            tmem_ld16(t_lane + 8 * D + c0, v);
            if (HAS_H0) {
                float vh[16];
                tmem_ld16(t_lane + 9 * D + c0, vh);
                if (live) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) vh[j] += gh0d[j];
                    st16(g_h0 + o, vh);
                }
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float gp = 0.f;
                if (live) {
                    gp = drop_mask ? v[j] * mk[j] : v[j];
                    if (act == 1) gp = xa[j] > 0.f ? gp : 0.f;
                    else if (act == 2) gp *= 1.f - xa[j] * xa[j];
                }
                v[j] = gp;
            }
            if (live) st16_il(g_pre_out, row, cq, KC, v);
            tmem_put_split(t_lane + 0 * D + c0, t_lane + 1 * D + c0, v);
            tmem_wait_st();
        }
        tc_fence_before();
        __syncthreads();

        // ---- GEMM batch 2: g_agg = g_pre . W_h ----
        if (tid == 0) {
            tc_fence_after();
            uint32_t acc = 0u;
            issue_gemm_ts<D>(smem_base, tmem_base + 0 * D, tmem_base + 1 * D, 6 * D, D, tmem_base + 2 * D, acc);
            tc_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        {
            float v[16];
            tmem_ld16(t_lane + 2 * D + c0, v);
            if (live) st16(g_agg + o, v);
        }
        // all TMEM reads of this tile are done before the next tile's tcgen05.st / MMAs overwrite the columns
        tc_fence_before();
        __syncthreads();
    }
    if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

template <int D, bool HH>
int launch_node_bwd(const float *g_hidden, const float *g_small, int g_small_stride, const float *w_small,
                    int w_small_rows, const float *g_h0_next,
                    const int32_t *remap, const float *saved, int64_t plane_rows, const float *drop_mask,
                    const float *W_h, const float *W_ih, const float *W_hh, int act, int64_t n_nodes,
                    const int64_t *n_nodes_dev, float *G4, float *g_pre, float *g_agg, float *g_h0, cudaStream_t st) {
    constexpr size_t smem = BwdSmem<D>::TOTAL;
    static_assert(smem <= 232448, "weights do not fit the 227 KB shared memory of one CTA");
    auto kern = k_node_bwd_tc<D, HH>;
    RG_CUDA_CALL(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, n_sm = 148;
    RG_CUDA_CALL(cudaGetDevice(&dev));
    RG_CUDA_CALL(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    const int64_t n_tiles = (n_nodes + kTcRows - 1) / kTcRows;
    const int grid = (int)(n_tiles < n_sm ? n_tiles : n_sm);  // persistent: one CTA per SM (it owns all of TMEM)
    kern<<<grid, 128 * (D / 16), smem, st>>>(g_hidden, g_small, g_small_stride, w_small, w_small_rows, g_h0_next, remap, saved,
                                             plane_rows,
                                             drop_mask,
                                             W_h, W_ih, W_hh, act, n_nodes, n_nodes_dev, G4, g_pre, g_agg, g_h0);
    RG_LAUNCH_CHECK();
    return RG_OK;
}

// ------------------------------------------------------------------------------------------------
// weight gradients: reductions over nodes on CUDA cores
//
// A slab = one 32-node tile.  Every source is ONE contiguous run of memory per slab: the
// lane-interleaved planes (x_act, h0, the four G4 planes, g_pre; rg_tc.cuh) because a 32-row tile of
// them is contiguous by construction, the row-major ones (agg, hidden, dropout mask, g_small) because
// consecutive rows are.  A slab is therefore staged with up to 11 TMA bulk copies (cp.async.bulk,
// completion on an mbarrier) issued by one thread into a 2-deep ring: no thread spends registers or
// issue slots on the loads and the copy of slab s+1 overlaps the FMAs of slab s.  Each thread owns an
// 8 x 8 register tile of one product space; in shared memory element (row k, col i) sits at
//   interleaved: ((i / 4) * 32 + k) * 4 + i % 4      row-major: k * width + i.
// ------------------------------------------------------------------------------------------------
constexpr int kWgNodes = 32;   // nodes per ring stage = one lane-interleaved tile
constexpr int kWgStages = 3;   // 3 x <= 60 KB: one CTA per SM (two thread halves splitting a slab's rows were measured
constexpr int kWgCtas = 148;   // slower: 0.61 -> 0.69 ms on the FB15k-237 training step)
constexpr int kWgHalves = 1;

struct WgLayout {   // float offsets of the sources inside one ring stage (-1 = source absent)
    int ox, om, oh, oa, ohid, og4, ogp, ogs, stage;
};

template <int D>
struct Wg {
    static constexpr int KS = kWgNodes;
    static constexpr int T1 = (D / 8) * (3 * D / 8), T3 = (D / 8) * (D / 8), T4 = D / 8, T5 = 4 * D / 8;
    static constexpr int TILES = 2 * T1 + T3 + T4 + T5;
    static constexpr int HALF = ((TILES + 31) / 32) * 32;   // threads of one half (one thread per register tile)
    static constexpr int THREADS = kWgHalves * HALF;
    // output layout
    static constexpr int O_WIH = 0, O_WHH = 3 * D * D, O_WH = 6 * D * D, O_WS = 7 * D * D, O_B = 7 * D * D + 8 * D;
    static constexpr int OUT = O_B + 4 * D;
    static WgLayout layout(bool has_mask, bool has_h0, bool has_small, int small_stride) {
        WgLayout l;
        constexpr int IL = il_tile_floats(D);       // one lane-interleaved tile (padded chunks)
        int o = 0;
        l.ox = o, o += IL;
        l.om = has_mask ? o : -1, o += has_mask ? KS * D : 0;
        l.oh = has_h0 ? o : -1, o += has_h0 ? IL : 0;
        l.oa = o, o += KS * D;
        l.ohid = has_small ? o : -1, o += has_small ? KS * D : 0;
        l.og4 = o, o += 4 * IL;
        l.ogp = o, o += IL;
        l.ogs = has_small ? o : -1, o += has_small ? KS * small_stride : 0;
        l.stage = o;
        return l;
    }
};

__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

template <int D, bool HAS_H0>
__global__ void __launch_bounds__(Wg<D>::THREADS, 1) k_node_wgrad(
    const float *__restrict__ saved, int64_t plane_rows, const float *__restrict__ drop_mask,
    const float *__restrict__ agg, const float *__restrict__ hidden, const float *__restrict__ G4,
    const float *__restrict__ g_pre, const float *__restrict__ g_small, int g_small_stride, int64_t n_nodes_host,
    const int64_t *__restrict__ n_nodes_dev, float *__restrict__ partial, WgLayout L) {
    using W = Wg<D>;
    extern __shared__ __align__(128) float wg_smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(wg_smem + kWgStages * L.stage);
    const int tid = threadIdx.x, half = tid / W::HALF;
    const int64_t n_nodes = n_nodes_dev ? *n_nodes_dev : n_nodes_host;
    const int64_t n_slabs = (n_nodes + kWgNodes - 1) / kWgNodes;
    const int64_t my_slabs = n_slabs > (int64_t)blockIdx.x ? (n_slabs - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const bool has_mask = drop_mask != nullptr, has_small = g_small != nullptr;

    // this thread's 8 x 8 tile.  x / g operand of node k, half h (4 floats): stage[xo + k * xk + h * xh], likewise g;
    // out = first output element; kind 0 = outer product (out[j * D + i]), 1 = column sums, 2 = idle
    int xo = 0, xk = 4, xh = kIlChunk, go = 0, gk = 4, gh = kIlChunk, mo = -1, out = 0, kind = 2;
    {
        constexpr int IB = D / 8;
        auto il = [](int base, int col) { return base + (col / 4) * kIlChunk; };   // interleaved: chunk offset, k * 4 added
        auto g4 = [&](int col) { return il(L.og4 + (col / D) * il_tile_floats(D), col % D); };
        int t = tid % W::HALF;
        if (t < W::T1) {
            xo = il(L.ox, 8 * (t % IB)), go = g4(8 * (t / IB));
            out = W::O_WIH + 8 * (t / IB) * D + 8 * (t % IB), kind = 0;
            if (has_mask) mo = L.om + 8 * (t % IB);
        } else if ((t -= W::T1) < W::T1) {
            const int jc = 8 * (t / IB);
            xo = il(L.oh, 8 * (t % IB)), go = g4(jc < 2 * D ? jc : jc + D);   // gates r, z, then g_n * r (fourth plane)
            out = W::O_WHH + jc * D + 8 * (t % IB), kind = HAS_H0 ? 0 : 2;
        } else if ((t -= W::T1) < W::T3) {
            xo = L.oa + 8 * (t % IB), xk = D, xh = 4;                          // agg: row-major
            go = il(L.ogp, 8 * (t / IB));
            out = W::O_WH + 8 * (t / IB) * D + 8 * (t % IB), kind = 0;
        } else if ((t -= W::T3) < W::T4) {
            xo = L.ohid + 8 * t, xk = D, xh = 4;                               // hidden, g_small: row-major
            go = L.ogs, gk = g_small_stride, gh = 4, out = W::O_WS + 8 * t, kind = has_small ? 0 : 2;
        } else if ((t -= W::T4) < W::T5) {
            go = g4(8 * t), out = W::O_B + 8 * t, kind = 1;
        }
    }
    if (tid == 0) {
        for (int s = 0; s < kWgStages; ++s) mbar_init(full + s, 1);
        fence_mbar_init();
    }
    __syncthreads();

    const size_t plane = il_plane_floats(plane_rows, D);      // saved planes
    const size_t gplane = il_plane_floats(n_nodes_host, D);   // G4 planes
    auto issue = [&](int64_t it) {   // thread 0: TMA bulk copies of this CTA's it-th slab into ring stage it % kWgStages
        const int s = (int)(it % kWgStages);
        float *st = wg_smem + s * L.stage;
        const int64_t base = ((int64_t)blockIdx.x + it * gridDim.x) * kWgNodes;
        const int64_t left = n_nodes_host - base;    // rows that exist in the ROW-MAJOR buffers (in-bounds copy)
        const uint32_t rows = (uint32_t)(left < kWgNodes ? left : kWgNodes);
        const uint32_t tb = il_tile_floats(D) * 4;   // a full interleaved tile (planes are padded to 32 rows)
        const uint32_t rb = rows * D * 4;
        uint32_t total = tb * 6 + rb;                                     // x_act, 4 x G4, g_pre; agg
        if (has_mask) total += rb;
        if (HAS_H0) total += tb;
        if (has_small) total += rb + rows * g_small_stride * 4;
        mbar_expect_tx(full + s, total);
        const size_t toff = (size_t)(base >> 5) * il_tile_floats(D);   // this tile inside an interleaved plane
        bulk_g2s(st + L.ox, saved + toff, tb, full + s);
        if (has_mask) bulk_g2s(st + L.om, drop_mask + (size_t)base * D, rb, full + s);
        if (HAS_H0) bulk_g2s(st + L.oh, saved + 5 * plane + toff, tb, full + s);
        bulk_g2s(st + L.oa, agg + (size_t)base * D, rb, full + s);
#pragma unroll
        for (int g = 0; g < 4; ++g)
            bulk_g2s(st + L.og4 + g * il_tile_floats(D), G4 + g * gplane + toff, tb, full + s);
        bulk_g2s(st + L.ogp, g_pre + toff, tb, full + s);
        if (has_small) {
            bulk_g2s(st + L.ohid, hidden + (size_t)base * D, rb, full + s);
            bulk_g2s(st + L.ogs, g_small + (size_t)base * g_small_stride, rows * g_small_stride * 4, full + s);
        }
    };
    if (tid == 0)
        for (int64_t it = 0; it < kWgStages && it < my_slabs; ++it) issue(it);

    float acc[8][8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;

    for (int64_t it = 0; it < my_slabs; ++it) {
        const int s = (int)(it % kWgStages);
        const float *st = wg_smem + s * L.stage;
        const int64_t base = ((int64_t)blockIdx.x + it * gridDim.x) * kWgNodes;
        const int rows = (int)(n_nodes - base < kWgNodes ? n_nodes - base : kWgNodes);   // true rows only
        mbar_wait(full + s, (uint32_t)((it / kWgStages) & 1));
        if (kind == 0) {
            const float *xp = st + xo, *gp = st + go;
            if (mo >= 0) {
                const float *mp = st + mo;
                for (int k = half; k < rows; k += kWgHalves) {
                    const float4 x0 = *reinterpret_cast<const float4 *>(xp + k * xk);
                    const float4 x1 = *reinterpret_cast<const float4 *>(xp + k * xk + xh);
                    const float4 m0 = *reinterpret_cast<const float4 *>(mp + k * D);
                    const float4 m1 = *reinterpret_cast<const float4 *>(mp + k * D + 4);
                    const float4 g0 = *reinterpret_cast<const float4 *>(gp + k * gk);
                    const float4 g1 = *reinterpret_cast<const float4 *>(gp + k * gk + gh);
                    const float xv[8] = {x0.x * m0.x, x0.y * m0.y, x0.z * m0.z, x0.w * m0.w,
                                         x1.x * m1.x, x1.y * m1.y, x1.z * m1.z, x1.w * m1.w};
                    const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j)
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[j][i] = fmaf(gv[j], xv[i], acc[j][i]);
                }
            } else {
#pragma unroll 4
                for (int k = half; k < rows; k += kWgHalves) {
                    const float4 x0 = *reinterpret_cast<const float4 *>(xp + k * xk);
                    const float4 x1 = *reinterpret_cast<const float4 *>(xp + k * xk + xh);
                    const float4 g0 = *reinterpret_cast<const float4 *>(gp + k * gk);
                    const float4 g1 = *reinterpret_cast<const float4 *>(gp + k * gk + gh);
                    const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                    const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j)
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[j][i] = fmaf(gv[j], xv[i], acc[j][i]);
                }
            }
        } else if (kind == 1) {
            const float *gp = st + go;
            for (int k = half; k < rows; k += kWgHalves) {
                const float4 g0 = *reinterpret_cast<const float4 *>(gp + k * gk);
                const float4 g1 = *reinterpret_cast<const float4 *>(gp + k * gk + gh);
                acc[0][0] += g0.x; acc[0][1] += g0.y; acc[0][2] += g0.z; acc[0][3] += g0.w;
                acc[0][4] += g1.x; acc[0][5] += g1.y; acc[0][6] += g1.z; acc[0][7] += g1.w;
            }
        }
        __syncthreads();   // every thread is done reading stage s
        if (tid == 0 && it + kWgStages < my_slabs) {
            fence_proxy_async();   // generic-proxy reads of the stage are ordered before the async-proxy refill
            issue(it + kWgStages);
        }
    }
    float *po = partial + ((size_t)blockIdx.x * kWgHalves + half) * W::OUT + out;
    if (kind == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            reinterpret_cast<float4 *>(po + j * D)[0] = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
            reinterpret_cast<float4 *>(po + j * D)[1] = make_float4(acc[j][4], acc[j][5], acc[j][6], acc[j][7]);
        }
    } else if (kind == 1) {
        reinterpret_cast<float4 *>(po)[0] = make_float4(acc[0][0], acc[0][1], acc[0][2], acc[0][3]);
        reinterpret_cast<float4 *>(po)[1] = make_float4(acc[0][4], acc[0][5], acc[0][6], acc[0][7]);
    }
}

// sum over CTAs of partial[cta][o] in a FIXED order (8 interleaved groups of CTAs, each summed ascending,
// then a fixed tree over the groups): deterministic, and 8x shorter dependent chains than one thread per
// output.  The sums go straight to their destinations: either the packed `out` vector (rg_node_wgrad), or
// the parameter gradients themselves (rg_node_wgrad_into: the GRU weights / biases ACCUMULATE over the
// layers, W_h / the small projection are per layer).  Outputs of idle tiles (no h0 / no g_small) are zero.
struct WgDst {
    float *out;                                   // packed [OUT] or NULL
    float *wih, *whh, *bih, *bhh;                 // += (GRU parameters, shared by all layers)
    float *wh, *ws;                               // =  (W_h [D][D]; small projection, first ws_rows rows of [8][D])
    int ws_rows;
};

template <int D, bool HAS_H0>
__global__ void __launch_bounds__(256) k_wgrad_reduce(const float *__restrict__ partial, int n_ctas, int has_small,
                                                      WgDst dst) {
    using W = Wg<D>;
    __shared__ float sm[8][32];
    const int ol = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int o = blockIdx.x * 32 + ol;
    float s = 0.f;
    if (o < W::OUT) {
        const bool idle = (!HAS_H0 && o >= W::O_WHH && o < W::O_WH) || (!has_small && o >= W::O_WS && o < W::O_B);
        if (!idle) {
            float s2[2] = {0.f, 0.f};
            int c = grp;
            for (; c + 8 < n_ctas; c += 16) {
                s2[0] += __ldg(partial + (size_t)c * W::OUT + o);
                s2[1] += __ldg(partial + (size_t)(c + 8) * W::OUT + o);
            }
            if (c < n_ctas) s2[0] += __ldg(partial + (size_t)c * W::OUT + o);
            s = s2[0] + s2[1];
        }
    }
    sm[grp][ol] = s;
    __syncthreads();
    if (grp != 0 || o >= W::OUT) return;
    const float v = ((sm[0][ol] + sm[1][ol]) + (sm[2][ol] + sm[3][ol])) + ((sm[4][ol] + sm[5][ol]) + (sm[6][ol] + sm[7][ol]));
    if (dst.out) {
        dst.out[o] = v;
        return;
    }
    if (o < W::O_WHH) {
        dst.wih[o] += v;
    } else if (o < W::O_WH) {
        dst.whh[o - W::O_WHH] += v;
    } else if (o < W::O_WS) {
        dst.wh[o - W::O_WH] = v;
    } else if (o < W::O_B) {
        const int j = o - W::O_WS;
        if (dst.ws && j < dst.ws_rows * D) dst.ws[j] = v;
    } else {                                      // column sums of g_r', g_z', g_n', g_n' r
        const int j = o - W::O_B, g = j / D, c = j % D;
        if (g < 3) dst.bih[j] += v;               // b_ih: r, z, n
        if (g < 2) dst.bhh[j] += v;               // b_hh: r, z, (n from g_n' r)
        if (g == 3) dst.bhh[2 * D + c] += v;
    }
}

template <int D, bool HH>
int launch_wgrad(const float *saved, int64_t plane_rows, const float *drop_mask, const float *agg, const float *hidden,
                 const float *G4, const float *g_pre, const float *g_small, int g_small_stride, int64_t n_nodes,
                 const int64_t *n_nodes_dev, float *partial, WgDst dst, cudaStream_t st) {
    using W = Wg<D>;
    const WgLayout L = W::layout(drop_mask != nullptr, HH, g_small != nullptr, g_small_stride);
    const size_t smem = (size_t)kWgStages * L.stage * 4 + 64;   // + mbarriers
    auto kern = k_node_wgrad<D, HH>;
    RG_CUDA_CALL(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t n_slabs = rg_cdiv(n_nodes, kWgNodes);
    const int grid = (int)(n_slabs < kWgCtas ? n_slabs : kWgCtas);
    kern<<<grid, W::THREADS, smem, st>>>(saved, plane_rows, drop_mask, agg, hidden, G4, g_pre, g_small, g_small_stride,
                                        n_nodes, n_nodes_dev, partial, L);
    RG_LAUNCH_CHECK();
    k_wgrad_reduce<D, HH><<<(W::OUT + 31) / 32, 256, 0, st>>>(partial, grid * kWgHalves, g_small != nullptr, dst);
    RG_LAUNCH_CHECK();
    return RG_OK;
}

}  // namespace

extern "C" int rg_node_bwd(int32_t hidden_dim, int64_t n_nodes, const int64_t *n_nodes_dev, const float *g_hidden,
                           const float *g_small, int32_t g_small_stride, const float *w_small, int32_t w_small_rows,
                           const float *g_h0_next,
                           const int32_t *remap, const float *saved, int64_t saved_plane_rows, const float *drop_mask,
                           const float *W_h, const float *W_ih, const float *W_hh, int32_t act, int32_t has_h0, float *G4,
                           float *g_pre, float *g_agg, float *g_h0, void *stream) {
    if (n_nodes < 0 || !saved || !W_h || !W_ih || !W_hh || !G4 || !g_pre || !g_agg || act < 0 || act > 2)
        return RG_ERR_BAD_ARG;
    if (!g_hidden && !g_small && !g_h0_next) return RG_ERR_BAD_ARG;
    if ((g_small == nullptr) != (w_small == nullptr) || (g_h0_next == nullptr) != (remap == nullptr)) return RG_ERR_BAD_ARG;
    if (g_small && (g_small_stride < 8 || g_small_stride % 4 || w_small_rows < 1 || w_small_rows > 8)) return RG_ERR_BAD_ARG;
    if (has_h0 && !g_h0) return RG_ERR_BAD_ARG;
    if (n_nodes == 0) return RG_OK;
    const int64_t plane = saved_plane_rows > 0 ? saved_plane_rows : n_nodes;
    if (plane < n_nodes) return RG_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
#define RG_NB(DD)                                                                                                      \
    return has_h0 ? launch_node_bwd<DD, true>(g_hidden, g_small, g_small_stride, w_small, w_small_rows, g_h0_next, remap, saved, plane, \
                                              drop_mask, W_h, W_ih, W_hh, act, n_nodes, n_nodes_dev, G4, g_pre, g_agg,  \
                                              g_h0, st)                                                                \
                  : launch_node_bwd<DD, false>(g_hidden, g_small, g_small_stride, w_small, w_small_rows, g_h0_next, remap, saved,     \
                                               plane, drop_mask, W_h, W_ih, W_hh, act, n_nodes, n_nodes_dev, G4, g_pre, \
                                               g_agg, g_h0, st)
    switch (hidden_dim) {
        case 16: RG_NB(16);
        case 32: RG_NB(32);
        case 48: RG_NB(48);
        default: return RG_ERR_UNSUPPORTED;
    }
#undef RG_NB
}

extern "C" int32_t rg_node_wgrad_ctas(void) { return kWgCtas * kWgHalves; }   // partial rows the caller provides

extern "C" int64_t rg_node_wgrad_out_floats(int32_t hidden_dim) {
    switch (hidden_dim) {
        case 16: return Wg<16>::OUT;
        case 32: return Wg<32>::OUT;
        case 48: return Wg<48>::OUT;
        default: return 0;
    }
}

extern "C" int rg_node_wgrad(int32_t hidden_dim, int64_t n_nodes, const int64_t *n_nodes_dev, const float *saved,
                             int64_t saved_plane_rows, const float *drop_mask, const float *agg, const float *hidden,
                             const float *G4, const float *g_pre, const float *g_small, int32_t g_small_stride,
                             int32_t has_h0, float *partial, float *out, float *g_wih, float *g_whh, float *g_bih,
                             float *g_bhh, float *g_wh, float *g_ws, int32_t ws_rows, void *stream) {
    if (n_nodes <= 0 || !saved || !agg || !G4 || !g_pre || !partial) return RG_ERR_BAD_ARG;
    if (!out && (!g_wih || !g_whh || !g_bih || !g_bhh || !g_wh)) return RG_ERR_BAD_ARG;
    if (g_small && (!hidden || g_small_stride < 8 || g_small_stride % 4)) return RG_ERR_BAD_ARG;
    if (ws_rows < 0 || ws_rows > 8) return RG_ERR_BAD_ARG;
    const int64_t plane = saved_plane_rows > 0 ? saved_plane_rows : n_nodes;
    if (plane < n_nodes) return RG_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    WgDst dst = {out, g_wih, g_whh, g_bih, g_bhh, g_wh, g_ws, ws_rows};
#define RG_WG(DD)                                                                                                    \
    return has_h0 ? launch_wgrad<DD, true>(saved, plane, drop_mask, agg, hidden, G4, g_pre, g_small, g_small_stride,  \
                                           n_nodes, n_nodes_dev, partial, dst, st)                                   \
                  : launch_wgrad<DD, false>(saved, plane, drop_mask, agg, hidden, G4, g_pre, g_small, g_small_stride, \
                                            n_nodes, n_nodes_dev, partial, dst, st)
    switch (hidden_dim) {
        case 16: RG_WG(16);
        case 32: RG_WG(32);
        case 48: RG_WG(48);
        default: return RG_ERR_UNSUPPORTED;
    }
#undef RG_WG
}
