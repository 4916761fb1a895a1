// Per-relation / per-query side of a GNN layer (reference Static/transductive/models.py:29-36):
// the attention is factorised, so besides the per-edge kernels a layer needs two small tables
//   ar8[r] = Wr_attn . rela[r]                     (one row per relation, 2R+1 rows)
//   aq8[b] = Wqr_attn . rela[q_rel[b]] + b_qr      (one row per query)
// and, in training, the parameter gradients that flow back through them.  Both are a few hundred rows
// of D = 48: ONE kernel each instead of ~10 / ~25 framework launches per layer (which, at a few
// microseconds apiece inside a captured step, had grown to a fifth of the FB15k-237 training step).
#include "rg_common.cuh"

namespace {

// thread per output: ar8 [rows][8], aq8 [n][8], w8 [8]
__global__ void __launch_bounds__(256) k_attn_tables(int D, int A, int rows, int n, const float *__restrict__ rela,
                                                     const float *__restrict__ Wr, const float *__restrict__ Wqr,
                                                     const float *__restrict__ bqr, const float *__restrict__ w_alpha,
                                                     const int64_t *__restrict__ q_rel, float *__restrict__ ar8,
                                                     float *__restrict__ aq8, float *__restrict__ w8) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int n_ar = rows * 8, n_aq = n * 8;
    if (i < n_ar) {
        const int r = i >> 3, k = i & 7;
        float s = 0.f;
        if (k < A)
            for (int c = 0; c < D; ++c) s = fmaf(__ldg(rela + (size_t)r * D + c), __ldg(Wr + k * D + c), s);
        ar8[i] = s;
    } else if (i < n_ar + n_aq) {
        const int j = i - n_ar, b = j >> 3, k = j & 7;
        float s = 0.f;
        if (k < A) {
            const float *row = rela + (size_t)q_rel[b] * D;
            s = __ldg(bqr + k);
            for (int c = 0; c < D; ++c) s = fmaf(__ldg(row + c), __ldg(Wqr + k * D + c), s);
        }
        aq8[j] = s;
    } else if (i < n_ar + n_aq + 8) {
        const int k = i - n_ar - n_aq;
        w8[k] = k < A ? __ldg(w_alpha + k) : 0.f;
    }
}

// Every parameter gradient of a layer's attention / relation side, written in place.  CTAs 0 .. R-1 own
// kPgRows relation rows each (rela_embed.weight rows: sum of the accumulator copies + g_ar8 . Wr + the query
// part, queries in ascending order => deterministic); the LAST CTA computes the small tensors (Wr_attn,
// Wqr_attn weight / bias, w_alpha weight / bias), which need reductions over ALL rows / queries.
// (One CTA for everything took 81 us per layer: 8 copies x 475 rows x 48 floats read by 1024 threads.)
constexpr int kPgThreads = 512;
constexpr int kPgRows = 16;

__global__ void __launch_bounds__(kPgThreads) k_attn_param_grads(
    int D, int A, int rows, int n, int copies, const float *__restrict__ rela, const float *__restrict__ Wr,
    const float *__restrict__ Wqr, const int64_t *__restrict__ q_rel, const float *__restrict__ g_rela_c,
    const float *__restrict__ g_ar8_c, const float *__restrict__ q_part, int q_slices, float *__restrict__ g_rela,
    float *__restrict__ g_Wr, float *__restrict__ g_Wqr, float *__restrict__ g_bqr, float *__restrict__ g_w_alpha,
    float *__restrict__ g_b_alpha) {
    extern __shared__ float sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float *s_gaq = sm;                        // [n][8]     per-query sums of g_as8 (= g_aq8)
    float *s_w = s_gaq + (size_t)n * 8;       // [2][8][D]  Wr, Wqr (zero rows above A)
    float *s_gar = s_w + 16 * D;              // [rows or kPgRows][8]  sum over copies of g_ar8
    for (int i = tid; i < n * 8; i += kPgThreads) {
        const int b = i >> 3, k = i & 7;
        float s = 0.f;
        for (int sl = 0; sl < q_slices; ++sl) s += __ldg(q_part + ((size_t)b * q_slices + sl) * 24 + k);
        s_gaq[i] = s;
    }
    for (int i = tid; i < 2 * 8 * D; i += kPgThreads) {
        const int which = i / (8 * D), k = (i / D) & 7, c = i % D;
        s_w[i] = k < A ? __ldg((which ? Wqr : Wr) + k * D + c) : 0.f;
    }
    const int n_row_ctas = gridDim.x - 1;
    if ((int)blockIdx.x < n_row_ctas) {
        // ---- rela_embed.weight rows [r0, r1) ----
        const int r0 = blockIdx.x * kPgRows, r1 = min(rows, r0 + kPgRows), nr = r1 - r0;
        for (int i = tid; i < nr * 8; i += kPgThreads) {
            float s = 0.f;
            for (int c = 0; c < copies; ++c) s += __ldg(g_ar8_c + (size_t)c * rows * 8 + (size_t)r0 * 8 + i);
            s_gar[i] = s;
        }
        __syncthreads();
        for (int i = tid; i < nr * D; i += kPgThreads) {
            const int rl = i / D, c = i % D, r = r0 + rl;
            float s = 0.f;
            for (int cp = 0; cp < copies; ++cp) s += __ldg(g_rela_c + ((size_t)cp * rows + r) * D + c);
#pragma unroll
            for (int k = 0; k < 8; ++k) s = fmaf(s_gar[rl * 8 + k], s_w[k * D + c], s);
            for (int b = 0; b < n; ++b) {           // + g_aq8[b] . Wqr[:, c] for the queries whose relation is r
                if ((int)q_rel[b] != r) continue;
                float t = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) t = fmaf(s_gaq[b * 8 + k], s_w[8 * D + k * D + c], t);
                s += t;
            }
            g_rela[(size_t)r * D + c] = s;
        }
        return;
    }
    // ---- last CTA: the small tensors ----
    for (int i = tid; i < rows * 8; i += kPgThreads) {
        float s = 0.f;
        for (int c = 0; c < copies; ++c) s += __ldg(g_ar8_c + (size_t)c * rows * 8 + i);
        s_gar[i] = s;
    }
    // w_alpha.weight / bias: columns 8..8+A and 16 of the per-query partial sums, one warp per column
    if (warp < 9) {
        const int col = warp < 8 ? 8 + warp : 16;
        float s = 0.f;
        for (int i = lane; i < n * q_slices; i += 32) s += __ldg(q_part + (size_t)i * 24 + col);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(RG_FULL_MASK, s, o);
        if (lane == 0) {
            if (warp < 8) {
                if (warp < A) g_w_alpha[warp] = s;
            } else {
                g_b_alpha[0] = s;
            }
        }
    }
    __syncthreads();
    // Wr_attn.weight[k][c] = sum_r g_ar8[r][k] rela[r][c];  Wqr_attn.weight[k][c] = sum_b g_aq8[b][k] rela[q_rel[b]][c]
    for (int o = tid; o < 2 * A * D; o += kPgThreads) {
        const bool second = o >= A * D;
        const int j = second ? o - A * D : o, k = j / D, c = j % D;
        float s = 0.f;
        if (!second) {
            float s4[4] = {0.f, 0.f, 0.f, 0.f};
            int r = 0;
            for (; r + 4 <= rows; r += 4) {
#pragma unroll
                for (int u = 0; u < 4; ++u) s4[u] = fmaf(s_gar[(r + u) * 8 + k], __ldg(rela + (size_t)(r + u) * D + c), s4[u]);
            }
            for (; r < rows; ++r) s4[0] = fmaf(s_gar[r * 8 + k], __ldg(rela + (size_t)r * D + c), s4[0]);
            s = (s4[0] + s4[1]) + (s4[2] + s4[3]);
            g_Wr[j] = s;
        } else {
            for (int b = 0; b < n; ++b) s = fmaf(s_gaq[b * 8 + k], __ldg(rela + (size_t)q_rel[b] * D + c), s);
            g_Wqr[j] = s;
        }
    }
    if (tid < A) {   // Wqr_attn.bias[k] = sum_b g_aq8[b][k]
        float s = 0.f;
        for (int b = 0; b < n; ++b) s += s_gaq[b * 8 + tid];
        g_bqr[tid] = s;
    }
}

}  // namespace

extern "C" int rg_attn_tables(int32_t hidden_dim, int32_t attn_dim, int32_t n_rows, int32_t n_query, const float *rela,
                              const float *Wr, const float *Wqr, const float *bqr, const float *w_alpha,
                              const int64_t *q_rel, float *ar8, float *aq8, float *w8, void *stream) {
    if (hidden_dim <= 0 || attn_dim < 1 || attn_dim > 8 || n_rows <= 0 || n_query <= 0) return RG_ERR_BAD_ARG;
    if (!rela || !Wr || !Wqr || !bqr || !w_alpha || !q_rel || !ar8 || !aq8 || !w8) return RG_ERR_BAD_ARG;
    const int total = (n_rows + n_query + 1) * 8;
    k_attn_tables<<<(unsigned)rg_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(hidden_dim, attn_dim, n_rows, n_query,
                                                                                 rela, Wr, Wqr, bqr, w_alpha, q_rel, ar8,
                                                                                 aq8, w8);
    RG_LAUNCH_CHECK();
    return RG_OK;
}

extern "C" int rg_attn_param_grads(int32_t hidden_dim, int32_t attn_dim, int32_t n_rows, int32_t n_query,
                                   int32_t grad_copies, const float *rela, const float *Wr, const float *Wqr,
                                   const int64_t *q_rel, const float *g_rela_copies, const float *g_ar8_copies,
                                   const float *q_part, int32_t q_slices, float *g_rela, float *g_Wr, float *g_Wqr,
                                   float *g_bqr, float *g_w_alpha, float *g_b_alpha, void *stream) {
    if (hidden_dim <= 0 || hidden_dim > 64 || attn_dim < 1 || attn_dim > 8 || n_rows <= 0 || n_query <= 0 ||
        grad_copies < 1 || q_slices < 1)
        return RG_ERR_BAD_ARG;
    if (!rela || !Wr || !Wqr || !q_rel || !g_rela_copies || !g_ar8_copies || !q_part || !g_rela || !g_Wr || !g_Wqr ||
        !g_bqr || !g_w_alpha || !g_b_alpha)
        return RG_ERR_BAD_ARG;
    const size_t smem = ((size_t)n_rows * 8 + (size_t)n_query * 8 + 16 * hidden_dim) * sizeof(float);
    if (smem > 200 * 1024) return RG_ERR_TOO_LARGE;
    RG_CUDA_CALL(cudaFuncSetAttribute(k_attn_param_grads, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)rg_cdiv(n_rows, kPgRows) + 1;
    k_attn_param_grads<<<grid, kPgThreads, smem, (cudaStream_t)stream>>>(
        hidden_dim, attn_dim, n_rows, n_query, grad_copies, rela, Wr, Wqr, q_rel, g_rela_copies, g_ar8_copies, q_part,
        q_slices, g_rela, g_Wr, g_Wqr, g_bqr, g_w_alpha, g_b_alpha);
    RG_LAUNCH_CHECK();
    return RG_OK;
}

// ---- fused training loss on the per-node scores (reference Static/transductive/base_model.py:58-60) ------
//   loss_q = -scores_all[q][obj_q] + logsumexp_e scores_all[q][e]
// evaluated WITHOUT materialising the dense (n, n_ent) score matrix: the visited nodes of query q are the
// rows [base, base + count) of `score`, every unvisited entity scores exactly 0 (models.py:87-88), so
//   logsumexp = m + log( sum_visited exp(s - m) + (n_ent - count) exp(-m) ),  m = max(max_visited s, 0 if any unvisited)
// and the positive is the visited row whose entity is obj_q (0 if unvisited).  Also writes the gradient
//   g_score[row] = softmax probability - [entity == obj_q]
// as rows {g, 0 x 7} of the g_small operand of rg_node_bwd (stride 8).  One CTA per query; fixed order.
namespace {
constexpr int kLossThreads = 256;
__device__ __forceinline__ float block_reduce(float v, float *sm, bool is_max) {
    for (int o = 16; o > 0; o >>= 1) {
        const float t = __shfl_xor_sync(RG_FULL_MASK, v, o);
        v = is_max ? fmaxf(v, t) : v + t;
    }
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = sm[0];
    for (int i = 1; i < kLossThreads / 32; ++i) r = is_max ? fmaxf(r, sm[i]) : r + sm[i];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kLossThreads) k_node_loss(const float *__restrict__ score,
                                                            const int32_t *__restrict__ node_e,
                                                            const int32_t *__restrict__ qinfo,
                                                            const int64_t *__restrict__ obj, int n_ent,
                                                            float *__restrict__ loss_q, float *__restrict__ g_small) {
    __shared__ float sm[kLossThreads / 32];
    __shared__ float s_pos;
    const int q = blockIdx.x, base = qinfo[2 * q], cnt = qinfo[2 * q + 1];
    const int target = (int)obj[q];
    if (threadIdx.x == 0) s_pos = 0.f;
    float mx = cnt < n_ent ? 0.f : -INFINITY;
    for (int i = threadIdx.x; i < cnt; i += kLossThreads) mx = fmaxf(mx, __ldg(score + base + i));
    mx = block_reduce(mx, sm, true);
    float s = 0.f;
    for (int i = threadIdx.x; i < cnt; i += kLossThreads) {
        const float v = __ldg(score + base + i);
        s += __expf(v - mx);
        if (__ldg(node_e + base + i) == target) s_pos = v;   // at most one row per query holds the target
    }
    s = block_reduce(s, sm, false) + (float)(n_ent - cnt) * __expf(-mx);
    if (threadIdx.x == 0) loss_q[q] = -s_pos + mx + __logf(s);
    const float inv = 1.f / s;
    for (int i = threadIdx.x; i < cnt; i += kLossThreads) {
        const float p = __expf(__ldg(score + base + i) - mx) * inv;
        const float g = p - (__ldg(node_e + base + i) == target ? 1.f : 0.f);
        float4 *o = reinterpret_cast<float4 *>(g_small + (size_t)(base + i) * 8);
        o[0] = make_float4(g, 0.f, 0.f, 0.f);
        o[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
}  // namespace

extern "C" int rg_node_loss(int32_t n_query, int32_t n_ent, const float *score, const int32_t *node_e,
                            const int32_t *qinfo, const int64_t *obj, float *loss_q, float *g_small, void *stream) {
    if (n_query <= 0 || n_ent <= 0 || !score || !node_e || !qinfo || !obj || !loss_q || !g_small) return RG_ERR_BAD_ARG;
    k_node_loss<<<n_query, kLossThreads, 0, (cudaStream_t)stream>>>(score, node_e, qinfo, obj, n_ent, loss_q, g_small);
    RG_LAUNCH_CHECK();
    return RG_OK;
}
