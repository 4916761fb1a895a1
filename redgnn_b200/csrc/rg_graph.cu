// Device graph build: the CSR-by-tail / CSR-by-head views of a KG from its fact arrays.
// Replaces the scipy CSR construction of DataLoader.load_graph (reference
// Static/transductive/load_data.py:76-81, M_sub = csr_matrix(...)) -- re-run by shuffle_train every
// epoch -- with a stable device radix sort (cub, not on the per-query hot path).
#include <cub/cub.cuh>

#include "rg_common.cuh"

namespace {

struct BuildWs {
    int32_t *keys_out, *iota, *order, *deg;
    void *cub_tmp;
    size_t cub_bytes, total;
};

size_t cub_bytes_for(int64_t F, int32_t n_ent) {
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const int32_t *)nullptr, (int32_t *)nullptr, (const int32_t *)nullptr,
                                    (int32_t *)nullptr, (int)F);
    cub::DeviceScan::ExclusiveSum(nullptr, b, (const int32_t *)nullptr, (int32_t *)nullptr, n_ent + 1);
    return a > b ? a : b;
}

BuildWs carve_build(void *ws, int64_t F, int32_t n_ent) {
    BuildWs w;
    char *p = (char *)ws;
    size_t off = 0;
    w.keys_out = (int32_t *)(p + off); off += rg_align256(4 * (size_t)F);
    w.iota = (int32_t *)(p + off);     off += rg_align256(4 * (size_t)F);
    w.order = (int32_t *)(p + off);    off += rg_align256(4 * (size_t)F);
    w.deg = (int32_t *)(p + off);      off += rg_align256(4 * ((size_t)n_ent + 1));
    w.cub_bytes = cub_bytes_for(F, n_ent);
    w.cub_tmp = p + off;               off += rg_align256(w.cub_bytes);
    w.total = off;
    return w;
}

__global__ void k_iota_hist(int64_t F, const int32_t *__restrict__ key, int32_t *iota, int32_t *deg) {
    int64_t f = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (f >= F) return;
    iota[f] = (int32_t)f;
    atomicAdd(&deg[key[f]], 1);   // integer counts: order independent
}

__global__ void k_gather_adj(int64_t F, const int32_t *__restrict__ order, const int32_t *__restrict__ other,
                             const int32_t *__restrict__ rel, int2 *adj) {
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= F) return;
    const int f = order[i];
    adj[i] = make_int2(other[f], rel[f]);
}

// rows keyed by `key`, entries (other, rel) in fact order inside a row
int build_csr(int64_t F, int32_t n_ent, const int32_t *key, const int32_t *other, const int32_t *rel, int32_t *ptr,
              int32_t *adj, BuildWs &w, cudaStream_t st) {
    RG_CUDA_CALL(cudaMemsetAsync(w.deg, 0, 4 * ((size_t)n_ent + 1), st));
    k_iota_hist<<<(unsigned)rg_cdiv(F, 256), 256, 0, st>>>(F, key, w.iota, w.deg);
    RG_LAUNCH_CHECK();
    int end_bit = 1;
    while ((1ll << end_bit) < n_ent) ++end_bit;
    size_t bytes = w.cub_bytes;
    RG_CUDA_CALL(cub::DeviceRadixSort::SortPairs(w.cub_tmp, bytes, key, w.keys_out, (const int32_t *)w.iota, w.order,
                                                 (int)F, 0, end_bit, st));
    bytes = w.cub_bytes;
    RG_CUDA_CALL(cub::DeviceScan::ExclusiveSum(w.cub_tmp, bytes, (const int32_t *)w.deg, ptr, n_ent + 1, st));
    k_gather_adj<<<(unsigned)rg_cdiv(F, 256), 256, 0, st>>>(F, w.order, other, rel, reinterpret_cast<int2 *>(adj));
    RG_LAUNCH_CHECK();
    return RG_OK;
}

// shuffle_train on the device: permuted pool rows -> fact rows, their inverses, self-loops
__global__ void k_resplit(const int32_t *__restrict__ pool, const int32_t *__restrict__ perm, int64_t n_keep,
                          int32_t n_ent, int32_t n_rel, int32_t *head, int32_t *rel, int32_t *tail) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n_keep) {
        const int64_t src = perm[i];
        const int h = pool[3 * src], r = pool[3 * src + 1], t = pool[3 * src + 2];
        head[i] = h, rel[i] = r, tail[i] = t;
        head[n_keep + i] = t, rel[n_keep + i] = r + n_rel, tail[n_keep + i] = h;
    } else if (i < n_keep + n_ent) {
        const int e = (int)(i - n_keep);
        const int64_t row = 2 * n_keep + e;
        head[row] = e, rel[row] = 2 * n_rel, tail[row] = e;
    }
}

}  // namespace

extern "C" int rg_graph_resplit(const int32_t *pool, const int32_t *perm, int64_t n_keep, int32_t n_ent,
                                int32_t n_rel, int32_t *head, int32_t *rel, int32_t *tail, void *stream) {
    if (!pool || !perm || n_keep < 0 || n_ent <= 0 || n_rel <= 0 || !head || !rel || !tail) return RG_ERR_BAD_ARG;
    if (2 * n_keep + n_ent >= (int64_t)INT32_MAX) return RG_ERR_TOO_LARGE;
    k_resplit<<<(unsigned)rg_cdiv(n_keep + n_ent, 256), 256, 0, (cudaStream_t)stream>>>(pool, perm, n_keep, n_ent,
                                                                                        n_rel, head, rel, tail);
    RG_LAUNCH_CHECK();
    return RG_OK;
}

extern "C" size_t rg_graph_build_workspace_bytes(int32_t n_ent, int64_t n_fact) {
    if (n_ent <= 0 || n_fact <= 0) return 0;
    return carve_build(nullptr, n_fact, n_ent).total;
}

extern "C" int rg_graph_build(const int32_t *head, const int32_t *rel, const int32_t *tail, int32_t n_ent,
                              int64_t n_fact, int32_t *in_ptr, int32_t *in_adj, int32_t *out_ptr, int32_t *out_adj,
                              void *ws, size_t ws_bytes, void *stream) {
    if (!head || !rel || !tail || n_ent <= 0 || n_fact <= 0 || !in_ptr || !in_adj || !out_ptr || !out_adj || !ws)
        return RG_ERR_BAD_ARG;
    if (n_fact >= (int64_t)INT32_MAX) return RG_ERR_TOO_LARGE;
    BuildWs w = carve_build(ws, n_fact, n_ent);
    if (ws_bytes < w.total) return RG_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = build_csr(n_fact, n_ent, tail, head, rel, in_ptr, in_adj, w, st);   // pull side: rows by tail
    if (rc) return rc;
    return build_csr(n_fact, n_ent, head, tail, rel, out_ptr, out_adj, w, st);  // push side: rows by head
}
