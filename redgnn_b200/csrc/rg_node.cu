// Fused node update: everything RED_GNN_*.forward does per NODE after the edge aggregation
// (reference Static/transductive/models.py:41 and :81-86):
//     x      = act(W_h . agg[j])                       GNNLayer.forward :41
//     h0[j]  = hidden_prev[src[j]]  (0 for new nodes)  zeros().index_copy_(1, old_nodes_new_idx, h0) :81
//     hidden = GRU_cell(x, h0[j])                      self.gate :83   (dropout :82 is identity in eval)
//     as8    = Ws_next . hidden                        the next layer's Ws_attn(hs), hoisted per node
//     score  = W_final . hidden                        :86 (last layer)
// One persistent CTA keeps all weights in shared memory (transposed, conflict-free) and walks tiles
// of 64 nodes; each thread owns an 8-row x (D/16)-column register tile, so the three small GEMMs run
// at FMA rate out of shared memory and the D-float rows touch HBM exactly once (read agg, read
// h_prev row, write hidden).  Inference only: training keeps the torch-composed path so autograd
// sees it.
#include <algorithm>
#include <cstdlib>

#include "rg_common.cuh"

namespace {

constexpr int kThreads = 128;  // 16 column groups x 8 row groups
constexpr int kTM = 64;        // nodes per tile
constexpr int kRT = 8;         // rows per thread

template <int D>
struct NodeSmem {
    static constexpr int S = D + 1;  // padded row stride of the activation tiles
    // float offsets
    static constexpr int WhT = 0;                  // [D][D]
    static constexpr int WihT = WhT + D * D;       // [D][3D]
    static constexpr int WhhT = WihT + D * 3 * D;  // [D][3D]
    static constexpr int Brz = WhhT + D * 3 * D;   // [2D] b_ih + b_hh of the r and z gates
    static constexpr int Bin = Brz + 2 * D;        // [D]
    static constexpr int Bhn = Bin + D;            // [D]
    static constexpr int Ws = Bhn + D;             // [9][D] rows 0..7 = Ws_next, row 8 = W_final
    static constexpr int A = Ws + 9 * D;           // [TM][S] agg tile, later the hidden tile
    static constexpr int X = A + kTM * S;          // [TM][S]
    static constexpr int H0 = X + kTM * S;         // [TM][S]
    static constexpr int Total = H0 + kTM * S;
};

__device__ __forceinline__ float act_apply(float v, int act) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return tanhf(v);
    return v;
}

__device__ __forceinline__ float sigmoid_precise(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int D, bool HAS_H0>
__global__ void __launch_bounds__(kThreads) k_node_update(
    const float *__restrict__ agg, const float *__restrict__ h_prev, const int32_t *__restrict__ src,
    const float *__restrict__ W_h, const float *__restrict__ W_ih, const float *__restrict__ W_hh,
    const float *__restrict__ b_ih, const float *__restrict__ b_hh, const float *__restrict__ Ws_next,
    const float *__restrict__ W_final, int act, int64_t n_nodes_host, const int64_t *__restrict__ n_nodes_dev,
    float *__restrict__ hidden, float *__restrict__ as8, float *__restrict__ score) {
    extern __shared__ float sm[];
    const int64_t n_nodes = n_nodes_dev ? *n_nodes_dev : n_nodes_host;
    using L = NodeSmem<D>;
    constexpr int S = L::S, CT = D / 16;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;

    // ---- weights -> shared memory, transposed to [k][col] ----
    for (int i = tid; i < D * D; i += kThreads) {
        int c = i / D, k = i % D;
        sm[L::WhT + k * D + c] = W_h[i];
    }
    for (int i = tid; i < 3 * D * D; i += kThreads) {
        int c = i / D, k = i % D;  // c in [0, 3D): gate-major output column
        sm[L::WihT + k * 3 * D + c] = W_ih[i];
        sm[L::WhhT + k * 3 * D + c] = HAS_H0 ? W_hh[i] : 0.f;
    }
    for (int i = tid; i < 2 * D; i += kThreads) sm[L::Brz + i] = b_ih[i] + b_hh[i];
    for (int i = tid; i < D; i += kThreads) {
        sm[L::Bin + i] = b_ih[2 * D + i];
        sm[L::Bhn + i] = b_hh[2 * D + i];
    }
    for (int i = tid; i < 9 * D; i += kThreads) {
        float v = 0.f;
        if (i < 8 * D) {
            if (Ws_next) v = Ws_next[i];
        } else if (W_final) {
            v = W_final[i - 8 * D];
        }
        sm[L::Ws + i] = v;
    }
    __syncthreads();

    const int64_t n_tiles = (n_nodes + kTM - 1) / kTM;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * kTM;
        const int rows = (int)min((int64_t)kTM, n_nodes - row0);
        // ---- stage agg rows and the re-indexed previous state ----
        constexpr int V = D / 4;  // float4 per row
        for (int i = tid; i < kTM * V; i += kThreads) {
            const int r = i / V, v = i % V;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), h = a;
            if (r < rows) {
                a = __ldg(reinterpret_cast<const float4 *>(agg + (size_t)(row0 + r) * D) + v);
                if (HAS_H0) {
                    const int s = __ldg(src + row0 + r);
                    if (s >= 0) h = __ldg(reinterpret_cast<const float4 *>(h_prev + (size_t)s * D) + v);
                }
            }
            float *pa = sm + L::A + r * S + v * 4;
            pa[0] = a.x; pa[1] = a.y; pa[2] = a.z; pa[3] = a.w;
            float *ph = sm + L::H0 + r * S + v * 4;
            ph[0] = h.x; ph[1] = h.y; ph[2] = h.z; ph[3] = h.w;
        }
        __syncthreads();

        // ---- phase A: X = act(agg . W_h^T) ----
        {
            float acc[kRT][CT];
#pragma unroll
            for (int i = 0; i < kRT; ++i)
#pragma unroll
                for (int j = 0; j < CT; ++j) acc[i][j] = 0.f;
            const float *pa = sm + L::A + (ty * kRT) * S;
#pragma unroll 4
            for (int k = 0; k < D; ++k) {
                float w[CT];
#pragma unroll
                for (int j = 0; j < CT; ++j) w[j] = sm[L::WhT + k * D + tx + 16 * j];
#pragma unroll
                for (int i = 0; i < kRT; ++i) {
                    const float a = pa[i * S + k];
#pragma unroll
                    for (int j = 0; j < CT; ++j) acc[i][j] = fmaf(a, w[j], acc[i][j]);
                }
            }
#pragma unroll
            for (int i = 0; i < kRT; ++i)
#pragma unroll
                for (int j = 0; j < CT; ++j)
                    sm[L::X + (ty * kRT + i) * S + tx + 16 * j] = act_apply(acc[i][j], act);
        }
        __syncthreads();

        // ---- phase B: GRU cell ----
        {
            float aR[kRT][CT], aZ[kRT][CT], aI[kRT][CT], aH[kRT][CT];
#pragma unroll
            for (int i = 0; i < kRT; ++i)
#pragma unroll
                for (int j = 0; j < CT; ++j) aR[i][j] = aZ[i][j] = aI[i][j] = aH[i][j] = 0.f;
            const float *px = sm + L::X + (ty * kRT) * S;
            const float *ph = sm + L::H0 + (ty * kRT) * S;
#pragma unroll 2
            for (int k = 0; k < D; ++k) {
                float wir[CT], wiz[CT], win[CT], whr[CT], whz[CT], whn[CT];
                const float *wi = sm + L::WihT + k * 3 * D + tx;
                const float *wh = sm + L::WhhT + k * 3 * D + tx;
#pragma unroll
                for (int j = 0; j < CT; ++j) {
                    wir[j] = wi[16 * j];
                    wiz[j] = wi[D + 16 * j];
                    win[j] = wi[2 * D + 16 * j];
                    if (HAS_H0) {
                        whr[j] = wh[16 * j];
                        whz[j] = wh[D + 16 * j];
                        whn[j] = wh[2 * D + 16 * j];
                    }
                }
#pragma unroll
                for (int i = 0; i < kRT; ++i) {
                    const float x = px[i * S + k];
                    const float h = HAS_H0 ? ph[i * S + k] : 0.f;
#pragma unroll
                    for (int j = 0; j < CT; ++j) {
                        aR[i][j] = fmaf(x, wir[j], aR[i][j]);
                        aZ[i][j] = fmaf(x, wiz[j], aZ[i][j]);
                        aI[i][j] = fmaf(x, win[j], aI[i][j]);
                        if (HAS_H0) {
                            aR[i][j] = fmaf(h, whr[j], aR[i][j]);
                            aZ[i][j] = fmaf(h, whz[j], aZ[i][j]);
                            aH[i][j] = fmaf(h, whn[j], aH[i][j]);
                        }
                    }
                }
            }
            // gates; the hidden tile overwrites the agg tile (free since the barrier after phase A)
#pragma unroll
            for (int i = 0; i < kRT; ++i) {
                const int r = ty * kRT + i;
#pragma unroll
                for (int j = 0; j < CT; ++j) {
                    const int c = tx + 16 * j;
                    const float rg = sigmoid_precise(aR[i][j] + sm[L::Brz + c]);
                    const float zg = sigmoid_precise(aZ[i][j] + sm[L::Brz + D + c]);
                    const float ng = tanhf(aI[i][j] + sm[L::Bin + c] + rg * (aH[i][j] + sm[L::Bhn + c]));
                    const float h0 = HAS_H0 ? sm[L::H0 + r * S + c] : 0.f;
                    const float hn = (1.0f - zg) * ng + zg * h0;
                    sm[L::A + r * S + c] = hn;
                    if (r < rows) hidden[(size_t)(row0 + r) * D + c] = hn;
                }
            }
        }
        __syncthreads();

        // ---- phase C: next layer's attention projection and / or the final score ----
        if (as8 || score) {
            for (int o = tid; o < kTM * 9; o += kThreads) {
                const int r = o / 9, a = o % 9;
                if (r >= rows) continue;
                if (a < 8 ? (as8 == nullptr) : (score == nullptr)) continue;
                const float *ph = sm + L::A + r * S;
                const float *pw = sm + L::Ws + a * D;
                float s = 0.f;
#pragma unroll 8
                for (int k = 0; k < D; ++k) s = fmaf(ph[k], pw[k], s);
                if (a < 8)
                    as8[(size_t)(row0 + r) * 8 + a] = s;
                else
                    score[row0 + r] = s;
            }
        }
        __syncthreads();
    }
}

template <int D, bool HH>
int launch_node(const float *agg, const float *h_prev, const int32_t *src, const float *W_h, const float *W_ih,
                const float *W_hh, const float *b_ih, const float *b_hh, const float *Ws_next,
                const float *W_final, int act, int64_t n_nodes, const int64_t *n_nodes_dev, float *hidden,
                float *as8, float *score, cudaStream_t st) {
    constexpr size_t smem = sizeof(float) * NodeSmem<D>::Total;
    auto kern = k_node_update<D, HH>;
    RG_CUDA_CALL(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, n_sm = 148, per_sm = 1;
    RG_CUDA_CALL(cudaGetDevice(&dev));
    RG_CUDA_CALL(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    RG_CUDA_CALL(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
    if (per_sm < 1) per_sm = 1;
    const int64_t n_tiles = (n_nodes + kTM - 1) / kTM;
    const int grid = (int)std::min<int64_t>(n_tiles, (int64_t)n_sm * per_sm);
    kern<<<grid, kThreads, smem, st>>>(agg, h_prev, src, W_h, W_ih, W_hh, b_ih, b_hh, Ws_next, W_final, act,
                                       n_nodes, n_nodes_dev, hidden, as8, score);
    RG_LAUNCH_CHECK();
    return RG_OK;
}

}  // namespace

__global__ void __launch_bounds__(256) k_scatter_scores(int64_t n_host, const int64_t *__restrict__ n_dev,
                                                        const int32_t *__restrict__ node_b,
                                                        const int32_t *__restrict__ node_e,
                                                        const float *__restrict__ score, int n_ent_out,
                                                        float *__restrict__ out) {
    const int64_t n = n_dev ? *n_dev : n_host;
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) out[(size_t)node_b[i] * n_ent_out + node_e[i]] = score[i];
}

// backward of the score scatter: g_node[j] = g_scores_all[node_b[j]][node_e[j]] (0 past the count)
__global__ void __launch_bounds__(256) k_gather_scores(int64_t n_host, const int64_t *__restrict__ n_dev,
                                                       const int32_t *__restrict__ node_b,
                                                       const int32_t *__restrict__ node_e,
                                                       const float *__restrict__ g_all, int n_ent_out,
                                                       float *__restrict__ g_node, int stride) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n_host) return;
    const int64_t n = n_dev ? *n_dev : n_host;
    const float g = i < n ? g_all[(size_t)node_b[i] * n_ent_out + node_e[i]] : 0.f;
    if (stride == 8) {   // row of a g_small [n][8] operand of rg_node_bwd: {g, 0 x 7}
        float4 *o = reinterpret_cast<float4 *>(g_node + i * 8);
        o[0] = make_float4(g, 0.f, 0.f, 0.f);
        o[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        g_node[i * stride] = g;
    }
}

// partial[q][slice][0..23] = sum over the slice's share of the query's node rows [base, base+count)
// of rows24[.][0..23]; the caller adds the kQuerySlices partials (fixed order => deterministic)
constexpr int kQuerySlices = 32;
__global__ void __launch_bounds__(240) k_query_sum24(const float *__restrict__ rows24,
                                                     const int32_t *__restrict__ qinfo,
                                                     float *__restrict__ partial) {
    __shared__ float sm[10][24];
    const int q = blockIdx.x, sl = blockIdx.y, k = threadIdx.x % 24, t = threadIdx.x / 24;  // 10 row-threads x 24 cols
    const int base = qinfo[2 * q], cnt = qinfo[2 * q + 1];
    const int per = (cnt + kQuerySlices - 1) / kQuerySlices;
    const int lo = sl * per, hi = min(cnt, lo + per);
    float acc = 0.f;
    for (int r = lo + t; r < hi; r += 10) acc += rows24[(size_t)(base + r) * 24 + k];
    sm[t][k] = acc;
    __syncthreads();
    if (t == 0) {
        float s = 0.f;
        for (int i = 0; i < 10; ++i) s += sm[i][k];
        partial[((size_t)q * kQuerySlices + sl) * 24 + k] = s;
    }
}

extern "C" int rg_gather_scores(int64_t n_nodes, const int64_t *n_nodes_dev, const int32_t *node_b,
                                const int32_t *node_e, const float *g_scores_all, int32_t n_ent_out, float *g_node,
                                int32_t out_stride, void *stream) {
    if (n_nodes < 0 || !node_b || !node_e || !g_scores_all || !g_node || n_ent_out <= 0 || out_stride < 1)
        return RG_ERR_BAD_ARG;
    if (n_nodes == 0) return RG_OK;
    k_gather_scores<<<(unsigned)rg_cdiv(n_nodes, 256), 256, 0, (cudaStream_t)stream>>>(
        n_nodes, n_nodes_dev, node_b, node_e, g_scores_all, n_ent_out, g_node, out_stride);
    RG_LAUNCH_CHECK();
    return RG_OK;
}

extern "C" int rg_query_sum8(int32_t n_query, const float *rows24, const int32_t *qinfo, float *partial,
                             void *stream) {
    if (n_query <= 0 || !rows24 || !qinfo || !partial) return RG_ERR_BAD_ARG;
    k_query_sum24<<<dim3(n_query, kQuerySlices), 240, 0, (cudaStream_t)stream>>>(rows24, qinfo, partial);
    RG_LAUNCH_CHECK();
    return RG_OK;
}

extern "C" int rg_scatter_scores(int64_t n_nodes, const int64_t *n_nodes_dev, const int32_t *node_b,
                                 const int32_t *node_e, const float *score, int32_t n_ent_out,
                                 float *scores_all, void *stream) {
    if (n_nodes < 0 || !node_b || !node_e || !score || !scores_all || n_ent_out <= 0) return RG_ERR_BAD_ARG;
    if (n_nodes == 0) return RG_OK;
    k_scatter_scores<<<(unsigned)rg_cdiv(n_nodes, 256), 256, 0, (cudaStream_t)stream>>>(
        n_nodes, n_nodes_dev, node_b, node_e, score, n_ent_out, scores_all);
    RG_LAUNCH_CHECK();
    return RG_OK;
}

// tensor-core (tcgen05 / TMEM, 3xTF32) variant, rg_node_tc.cu
int rg_node_update_tc(int32_t hidden_dim, int64_t n_nodes, const int64_t *n_nodes_dev, const float *agg,
                      const float *h_prev, const int32_t *src, const float *W_h, const float *W_ih,
                      const float *W_hh, const float *b_ih, const float *b_hh, const float *Ws_next,
                      const float *W_final, int32_t act, float *hidden, float *as8, float *score,
                      const float *drop_mask, float *saved, int32_t ws_rows, cudaStream_t st);

// training forward: tensor-core kernel only (hidden_dim <= 48); also writes saved[6][n][D] =
// {act(W_h agg) before dropout, r, z, n, W_hn h0 + b_hn, h0} for the backward pass
extern "C" int rg_node_update_train(int32_t hidden_dim, int64_t n_nodes, const int64_t *n_nodes_dev,
                                    const float *agg, const float *h_prev,
                                    const int32_t *src, const float *W_h, const float *W_ih, const float *W_hh,
                                    const float *b_ih, const float *b_hh, int32_t act, const float *drop_mask,
                                    float *hidden, float *saved, const float *Ws_next, int32_t ws_rows,
                                    const float *W_final, float *as8, float *score, void *stream) {
    if (n_nodes < 0 || !agg || !W_h || !W_ih || !W_hh || !b_ih || !b_hh || !hidden || !saved) return RG_ERR_BAD_ARG;
    if ((h_prev == nullptr) != (src == nullptr) || act < 0 || act > 2) return RG_ERR_BAD_ARG;
    if ((as8 != nullptr) != (Ws_next != nullptr) || (score != nullptr) != (W_final != nullptr)) return RG_ERR_BAD_ARG;
    if (Ws_next && (ws_rows < 1 || ws_rows > 8)) return RG_ERR_BAD_ARG;
    if (hidden_dim > 48) return RG_ERR_UNSUPPORTED;
    if (n_nodes == 0) return RG_OK;
    return rg_node_update_tc(hidden_dim, n_nodes, n_nodes_dev, agg, h_prev, src, W_h, W_ih, W_hh, b_ih, b_hh, Ws_next,
                             W_final, act, hidden, as8, score, drop_mask, saved, ws_rows, (cudaStream_t)stream);
}

extern "C" int rg_node_update(int32_t hidden_dim, int64_t n_nodes, const int64_t *n_nodes_dev, const float *agg,
                              const float *h_prev,
                              const int32_t *src, const float *W_h, const float *W_ih, const float *W_hh,
                              const float *b_ih, const float *b_hh, const float *Ws_next, const float *W_final,
                              int32_t act, float *hidden, float *as8, float *score, void *stream) {
    if (n_nodes < 0 || !agg || !W_h || !W_ih || !W_hh || !b_ih || !b_hh || !hidden) return RG_ERR_BAD_ARG;
    if ((h_prev == nullptr) != (src == nullptr)) return RG_ERR_BAD_ARG;
    if ((as8 != nullptr) != (Ws_next != nullptr) || (score != nullptr) != (W_final != nullptr)) return RG_ERR_BAD_ARG;
    if (act < 0 || act > 2) return RG_ERR_BAD_ARG;
    if (n_nodes == 0) return RG_OK;
    cudaStream_t st = (cudaStream_t)stream;
    // default: tensor cores (hidden_dim <= 48 fits one CTA's shared memory); REDGNN_NODE_SIMT=1 forces
    // the CUDA-core kernel (also the path for hidden_dim 64)
    const char *force_simt = std::getenv("REDGNN_NODE_SIMT");
    if (hidden_dim <= 48 && !(force_simt && force_simt[0] == '1'))
        return rg_node_update_tc(hidden_dim, n_nodes, n_nodes_dev, agg, h_prev, src, W_h, W_ih, W_hh, b_ih, b_hh,
                                 Ws_next, W_final, act, hidden, as8, score, nullptr, nullptr, 8, st);
#define RG_NODE(DD)                                                                                              \
    return h_prev ? launch_node<DD, true>(agg, h_prev, src, W_h, W_ih, W_hh, b_ih, b_hh, Ws_next, W_final, act,   \
                                          n_nodes, n_nodes_dev, hidden, as8, score, st)                                       \
                  : launch_node<DD, false>(agg, h_prev, src, W_h, W_ih, W_hh, b_ih, b_hh, Ws_next, W_final, act,  \
                                           n_nodes, n_nodes_dev, hidden, as8, score, st)
    switch (hidden_dim) {
        case 16: RG_NODE(16);
        case 32: RG_NODE(32);
        case 48: RG_NODE(48);
        case 64: RG_NODE(64);
        default: return RG_ERR_UNSUPPORTED;
    }
#undef RG_NODE
}

// ------------------------------------------------------------------------------------------------
// Filtered ranking on the device: utils.cal_ranks (reference Static/transductive/utils.py:7-14)
// without the (n, n_ent) D2H copy and scipy.stats.rankdata.  For query row s and answer t:
//   s'        = (s - min(s)) + 1e-8f                                   (float32, like numpy)
//   full      = #{e : s'[e] > s'[t]} + (#{e : s'[e] == s'[t]} + 1) / 2  rankdata(-s', 'average')
//   filtered  = 1 + #{e in filt : fs[e] > fs[t]},  fs = s' on the filter set, 0 elsewhere  ('min')
//   rank      = full - filtered + 1
// One CTA per query; answers of a query are processed in order (ascending entity id).
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int kRankThreads = 256;

__device__ __forceinline__ float block_min(float v, float *sm) {
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = sm[0];
    for (int i = 1; i < kRankThreads / 32; ++i) r = fminf(r, sm[i]);
    __syncthreads();
    return r;
}
__device__ __forceinline__ int block_sum_int(int v, int *sm) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    int r = 0;
    for (int i = 0; i < kRankThreads / 32; ++i) r += sm[i];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kRankThreads) k_filtered_ranks(const float *__restrict__ scores, int n_ent,
                                                                 const int32_t *__restrict__ ans_ptr,
                                                                 const int32_t *__restrict__ ans_idx,
                                                                 const int32_t *__restrict__ flt_ptr,
                                                                 const int32_t *__restrict__ flt_idx,
                                                                 double *__restrict__ ranks) {
    __shared__ float smf[kRankThreads / 32];
    __shared__ int smi[kRankThreads / 32];
    const int q = blockIdx.x;
    const float *s = scores + (size_t)q * n_ent;
    float mn = INFINITY;
    for (int e = threadIdx.x; e < n_ent; e += kRankThreads) mn = fminf(mn, s[e]);
    mn = block_min(mn, smf);
    const int f0 = flt_ptr[q], f1 = flt_ptr[q + 1];
    for (int a = ans_ptr[q]; a < ans_ptr[q + 1]; ++a) {
        const int t = ans_idx[a];
        const float st = (s[t] - mn) + 1e-8f;
        int greater = 0, equal = 0, fgreater = 0, t_in_filter = 0;
        for (int e = threadIdx.x; e < n_ent; e += kRankThreads) {
            const float v = (s[e] - mn) + 1e-8f;
            greater += v > st;
            equal += v == st;
        }
        for (int i = f0 + threadIdx.x; i < f1; i += kRankThreads) {
            const int e = flt_idx[i];
            t_in_filter |= (e == t);
            fgreater += ((s[e] - mn) + 1e-8f) > st;
        }
        greater = block_sum_int(greater, smi);
        equal = block_sum_int(equal, smi);
        fgreater = block_sum_int(fgreater, smi);
        t_in_filter = block_sum_int(t_in_filter, smi);
        if (threadIdx.x == 0) {
            // answer outside its own filter set: fs[t] = 0 and every filter entry (> 0) outranks it
            const int fcount = t_in_filter ? fgreater : (f1 - f0);
            const double full = (double)greater + ((double)equal + 1.0) * 0.5;
            ranks[a] = full - (1.0 + (double)fcount) + 1.0;
        }
    }
}
}  // namespace

extern "C" int rg_filtered_ranks(int32_t n_query, int32_t n_ent, const float *scores, const int32_t *ans_ptr,
                                 const int32_t *ans_idx, const int32_t *flt_ptr, const int32_t *flt_idx,
                                 double *ranks, void *stream) {
    if (n_query < 0 || n_ent <= 0 || !scores || !ans_ptr || !ans_idx || !flt_ptr || !flt_idx || !ranks)
        return RG_ERR_BAD_ARG;
    if (n_query == 0) return RG_OK;
    k_filtered_ranks<<<n_query, kRankThreads, 0, (cudaStream_t)stream>>>(scores, n_ent, ans_ptr, ans_idx, flt_ptr,
                                                                       flt_idx, ranks);
    RG_LAUNCH_CHECK();
    return RG_OK;
}
