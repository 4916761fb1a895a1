/*
 * redgnn_b200.h -- C ABI of libredgnn_b200.so: the B200 (sm_100a) implementation of RED-GNN's
 * query-conditioned relational-digraph propagation path.
 *
 * The reference (LARS-research/RED-GNN, Static/{transductive,inductive}) is pure Python; the
 * "FFI" a maintainer would bind is therefore a ctypes stub.  Every entry
 * point below names the reference code it replaces (paths relative to /root/reference/Static/).
 *
 * Contract:
 *   - plain C types only: device pointers, sizes, a cudaStream_t passed as void*;
 *   - the library allocates nothing, keeps no handles and no global state; all buffers
 *     (graph arrays, frontier state, workspaces, outputs) are owned by the caller;
 *   - every call is asynchronous on the given stream and never synchronises the host; the only
 *     host read-back on the path is the caller's own copy of the 8-word `counts` block;
 *   - return value: 0 on success, negative rg_status otherwise; nothing is thrown.
 *
 * Data layout in HBM
 *   graph      : head[F], rel[F], tail[F] int32 in REFERENCE ROW ORDER (triples, then the
 *                n_ent self-loops (e, 2*n_rel, e));  CSR-by-tail  in_ptr[n_ent+1], in_adj[F]
 *                = (head, rel) pairs and CSR-by-head out_ptr[n_ent+1], out_adj[F] = (tail, rel)
 *                pairs, both stable in fact order.
 *   frontier   : the per-query node set of one layer, stored twice:
 *                emask[n_ent][Wn] uint32, Wn = ceil(n_query/32): bit b of row e  <=> (b,e) in set
 *                dict [n_query][We] {uint32 bits, uint32 prefix}, We = ceil(n_ent/32):
 *                  bit (e%32) of word e/32 of row b <=> (b,e) in set;  prefix = number of set
 *                  members that precede this word in (b, e) lexicographic order, so that
 *                  rank(b,e) = prefix + popc(bits & ((1<<(e%32))-1)) is the row of (b,e) in the
 *                  reference's sorted `torch.unique(dim=0)` node list.
 */
#ifndef REDGNN_B200_H
#define REDGNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RG_ABI_VERSION 1

typedef enum rg_status {
    RG_OK = 0,
    RG_ERR_BAD_ARG = -1,      /* null pointer, negative size, inconsistent shapes            */
    RG_ERR_UNSUPPORTED = -2,  /* hidden_dim not in {16,32,48,64} or attn_dim > 8              */
    RG_ERR_WORKSPACE = -3,    /* workspace smaller than rg_workspace_bytes()                  */
    RG_ERR_TOO_LARGE = -4,    /* an index space of this call would exceed 2^31-1              */
    RG_ERR_IO = -5,           /* rg_text_*: the file cannot be opened / mapped                */
    RG_ERR_PARSE = -6,        /* rg_text_*: a line does not hold exactly three names          */
    RG_ERR_UNKNOWN_NAME = -7, /* rg_text_*: a name is in neither entity2id nor relation2id    */
    RG_ERR_HOST = -8,         /* rg_text_*: host allocation failed                            */
    RG_ERR_CUDA_BASE = -1000  /* -(1000 + cudaError_t) for a CUDA launch / API failure        */
} rg_status;

/* word indices of the int64 `counts[RG_COUNTS_WORDS]` device block written by the frontier calls */
#define RG_COUNTS_WORDS 8
#define RG_CNT_N_IN 0     /* unique input nodes  (rows of head_nodes)                         */
#define RG_CNT_E 1        /* edges of this hop                                                */
#define RG_CNT_N_OUT 2    /* rows of tail_nodes                                               */
#define RG_CNT_ERR 3      /* bit 0: node out of range in rg_frontier_from_nodes               */
#define RG_CNT_HEAVY 4    /* overflow flag of a heavy-segment queue (edge kernels)            */

/* Static graph of one KG.  Replaces DataLoader.KG / M_sub, tKG / tM_sub (transductive/
 * load_data.py:76-89) and tra_KG/tra_sub, ind_KG/ind_sub (inductive/load_data.py:88-98). */
typedef struct rg_graph {
    int32_t n_ent;
    int32_t n_rel;           /* R; relation ids run 0..2R (2R = self-loop)                    */
    int64_t n_fact;          /* F = rows incl. inverses and self-loops                        */
    const int32_t *head, *rel, *tail;   /* [F] reference row order                            */
    const int32_t *in_ptr;   /* [n_ent+1] CSR by tail                                         */
    const int32_t *in_adj;   /* [F][2]  (head, rel), fact order inside a row                  */
    const int32_t *out_ptr;  /* [n_ent+1] CSR by head                                         */
    const int32_t *out_adj;  /* [F][2]  (tail, rel), fact order inside a row                  */
} rg_graph;

/* Node set of one layer (see "Data layout").  Replaces the (N,2) `nodes` LongTensor of
 * RED_GNN_*.forward (transductive/models.py:73,78) as the carried state between hops. */
typedef struct rg_frontier {
    int32_t n_query;
    int32_t n_ent;
    uint32_t *emask;         /* [n_ent][Wn]                                                   */
    uint32_t *dict;          /* [n_query][We][2]                                              */
    int32_t *qinfo;          /* optional [n_query][2] = {rank of the query's first node, node count};
                                count == n_ent marks a COMPLETE frontier (rank = base + entity)    */
} rg_frontier;

/* Segment description for the fused edge kernels.
 * mode 0 (explicit): segment s owns adj[seg_ptr[s] .. seg_ptr[s+1]) = (peer_row, rel) pairs and
 *                    belongs to query seg_query[s].  Used for arbitrary caller-provided edge lists
 *                    (the public GNNLayer.forward(edges) surface, models.py:23-43).
 * mode 1 (implicit): segment s is node (seg_query[s], seg_ent[s]); its candidate edges are the
 *                    CSR row ent_ptr/ent_adj of that entity = (peer_entity, rel); a candidate is an
 *                    edge iff (query, peer_entity) is in `peer_dict`, whose rank gives peer_row.
 *                    No per-layer edge list exists in HBM on this path. */
typedef struct rg_segments {
    int32_t mode;
    int32_t n_ent;                 /* implicit: dictionary row length We = ceil(n_ent/32)      */
    int64_t n_seg;                 /* number of segments (an UPPER BOUND when n_seg_dev is set) */
    const int64_t *n_seg_dev;      /* optional device-resident true count (no host read-back)  */
    const int32_t *seg_query;      /* [n_seg]                                                  */
    const int32_t *seg_ptr;        /* explicit: [n_seg+1]                                      */
    const int32_t *adj;            /* explicit: [E][2]; implicit: ent_adj [F][2]               */
    const int32_t *seg_ent;        /* implicit: [n_seg]                                        */
    const int32_t *ent_ptr;        /* implicit: [n_ent+1]                                      */
    const uint32_t *peer_dict;     /* implicit: [n_query][We][2]                               */
    const int32_t *peer_qinfo;     /* implicit, optional: rg_frontier.qinfo of the peer frontier     */
    int32_t n_table_rows;          /* rows of the relation tables (2R+1); > 0 lets the edge kernels
                                      stage rela / ar8 in shared memory when they fit            */
} rg_segments;

/* Queue for segments longer than RG_HEAVY_CHUNK candidate slots: they are cut into chunks that
 * other warps reduce into `partial`, then summed in chunk order (deterministic).  The persistent
 * kernels drain the queue themselves (idle warps take chunks while others still work on segments);
 * the other variants run a second kernel over it. */
#ifndef RG_HEAVY_CHUNK
#define RG_HEAVY_CHUNK 256        /* rg_edge_agg_fwd: slots a heavy segment's owner warp keeps ...            */
#endif
#ifndef RG_HEAVY_SUB
#define RG_HEAVY_SUB 256          /* ... and the size of the pieces the rest is cut into (queue sizing unit) */
#endif
#ifndef RG_HEAVY_CHUNK_BWD
#define RG_HEAVY_CHUNK_BWD 256    /* rg_edge_agg_bwd: slots a heavy segment's owner warp keeps ...            */
#endif
#ifndef RG_HEAVY_SUB_BWD
#define RG_HEAVY_SUB_BWD 128      /* ... and the size of the pieces the rest is cut into (queue sizing unit) */
#endif
typedef struct rg_heavy {
    int32_t max_chunks;
    int32_t max_nodes;
    int32_t *counters;       /* [8] zeroed by the library: n_chunks, n_nodes, overflow, taken, warps done */
    int32_t *chunk_seg;      /* [max_chunks]                                                   */
    int32_t *chunk_idx;      /* [max_chunks]                                                   */
    int32_t *node_seg;       /* [max_nodes]                                                    */
    int32_t *node_base;      /* [max_nodes]                                                    */
    int32_t *node_n;         /* [max_nodes]                                                    */
    float *partial;          /* [max_chunks][row_floats]                                       */
} rg_heavy;

int rg_abi_version(void);
const char *rg_strerror(int status);

/* sizes of the caller-allocated state */
size_t rg_frontier_emask_bytes(int32_t n_query, int32_t n_ent);
size_t rg_frontier_dict_bytes(int32_t n_query, int32_t n_ent);
size_t rg_workspace_bytes(int32_t n_query, int32_t n_ent, int64_t n_fact);

/* ---- text side of the graph store (host-only: no GPU, no stream) -----------------
 * DataLoader.read_triples (transductive/load_data.py:58-67, inductive/load_data.py:76-86): every line
 * of facts/train/valid/test.txt is `h, r, t = line.strip().split()` mapped through entity2id /
 * relation2id (transductive/load_data.py:11-25 by line number, inductive/load_data.py:14-40 from
 * "name id" files).  The caller hands both dictionaries over as flat arrays; the file is mapped
 * read-only and parsed by `n_threads` host threads (<= 0: all cores; at most 64 and one per MiB of file)
 * split at line boundaries.
 *   lines      : as Python's text-mode iteration yields them ("\n", "\r\n" or a lone "\r" end a line;
 *                a last line without terminator counts when it is not empty);
 *   separators : what str.split() accepts in UTF-8 text (ASCII white space, 0x1c-0x1f, U+0085, U+00A0,
 *                U+1680, U+2000-200A, U+2028/9, U+202F, U+205F, U+3000); names compare as bytes.
 * Errors follow the reference's first failure in file order: a line without exactly three names is
 * RG_ERR_PARSE (its ValueError), a name missing from its dictionary RG_ERR_UNKNOWN_NAME (its
 * KeyError), an unreadable file RG_ERR_IO (its FileNotFoundError); *err_line = 0-based line.
 * The file must not shrink while the call runs (it is mapped, not copied). */
typedef struct rg_name_table {
    const char *bytes;       /* the names back to back (UTF-8 as in the file, no terminators)  */
    const int64_t *off;      /* [n+1]: name k is bytes[off[k], off[k+1])                        */
    const int32_t *id;       /* [n]: the dictionary value; a name listed twice: the later wins  */
    int64_t n;
} rg_name_table;

int rg_text_count_lines(const char *path, int64_t *n_lines);
/* out[cap_rows][3] int32 (h, r, t) in file order; *n_rows = lines of the file (set whenever the file
 * can be read; RG_ERR_BAD_ARG if it exceeds cap_rows, nothing is written then). */
int rg_text_parse_triples(const char *path, const rg_name_table *ent, const rg_name_table *rel, int32_t *out,
                          int64_t cap_rows, int64_t *n_rows, int64_t *err_line, int32_t n_threads);

/* Graph build: the CSR views of rg_graph from the fact arrays (replaces the scipy csr_matrix of
 * load_graph, transductive/load_data.py:76-81, re-run every epoch by shuffle_train :152-164).
 * head/rel/tail: device int32 [n_fact] in reference row order INCLUDING the self-loop block.
 * Outputs (caller-allocated): in_ptr/out_ptr [n_ent+1], in_adj/out_adj [n_fact][2]; rows are stable
 * in fact order.  Every entity id must lie in [0, n_ent). */
size_t rg_graph_build_workspace_bytes(int32_t n_ent, int64_t n_fact);
int rg_graph_build(const int32_t *head, const int32_t *rel, const int32_t *tail, int32_t n_ent,
                   int64_t n_fact, int32_t *in_ptr, int32_t *in_adj, int32_t *out_ptr, int32_t *out_adj,
                   void *ws, size_t ws_bytes, void *stream);

/* shuffle_train (transductive/load_data.py:152-164) without the host round trip of the KG: `pool`
 * [n_all][3] int32 is the fixed union of the fact and train triples (file order, device resident),
 * `perm` the host-drawn np.random.permutation(n_all) (device int32; only its first n_keep = n_all*3/4
 * entries are read).  Writes the new KG in reference row order: rows [0, n_keep) = pool[perm[i]],
 * rows [n_keep, 2 n_keep) their inverses (t, r + n_rel, h) (double_triple :69-74), then the n_ent
 * self-loops (e, 2 n_rel, e) (:77-79).  Follow with rg_graph_build on the same arrays. */
int rg_graph_resplit(const int32_t *pool, const int32_t *perm, int64_t n_keep, int32_t n_ent,
                     int32_t n_rel, int32_t *head, int32_t *rel, int32_t *tail, void *stream);

/* ---- expansion: DataLoader.get_neighbors (transductive/load_data.py:106-131,
 *      inductive/load_data.py:115-143) --------------------------------------------------------- */

/* nodes[N][2] int64 (batch_idx, entity), any order, duplicates allowed -> frontier state.
 * Replaces the node_1hot construction (load_data.py:115).  counts[RG_CNT_N_IN] = unique nodes. */
int rg_frontier_from_nodes(const int64_t *nodes, int64_t n_nodes, rg_frontier *fr,
                           int64_t *counts, void *ws, size_t ws_bytes, void *stream);

/* One hop: which facts have their head in which query's frontier (M_sub.dot(node_1hot),
 * load_data.py:116), the per-query dedup of the tails (torch.unique, :123) and the sizes
 * counts[RG_CNT_E], counts[RG_CNT_N_OUT].  `ws` keeps the per-block edge offsets that
 * rg_edges_emit() reads; it must be the same buffer, untouched in between. */
int rg_frontier_step(const rg_graph *g, const rg_frontier *in, rg_frontier *out,
                     int64_t *counts, void *ws, size_t ws_bytes, void *stream);

/* Sorted unique node list of a frontier (tail_nodes, load_data.py:123): any of the three
 * outputs may be NULL.  nodes64 is [N][2] int64, node_b / node_e are int32 [N]. */
int rg_frontier_nodes(const rg_frontier *fr, int64_t *nodes64, int32_t *node_b, int32_t *node_e,
                      void *stream);

/* old_nodes_new_idx (load_data.py:127-129): row of every `in` node inside `out`'s node list.
 * inverse32[N'] (optional, caller pre-fills it with -1) receives the opposite map: the `in` row of
 * an `out` node, which is what the h0 re-index zeros().index_copy_(1, old_nodes_new_idx, h0)
 * (transductive/models.py:81) needs when it is fused into the node update as a gather. */
int rg_frontier_remap(const rg_frontier *in, const rg_frontier *out, int64_t *remap64,
                      int32_t *remap32, int32_t *inverse32, void *stream);

/* sampled_edges[E][6] int64 = (batch, head, rel, tail, head_index, tail_index) in the
 * reference's order (fact row ascending, batch index descending; load_data.py:117-125). */
int rg_edges_emit(const rg_graph *g, const rg_frontier *in, const rg_frontier *out,
                  const void *ws, size_t ws_bytes, int64_t n_edges, int64_t *edges, void *stream);

/* The public get_neighbors hop (load_data.py:106-131) as two calls around the one count read-back:
 * rg_get_neighbors_expand = rg_frontier_from_nodes + rg_frontier_step (counts_in / counts_out as above),
 * rg_get_neighbors_emit   = rg_frontier_nodes (tail_nodes [N'][2]) + rg_frontier_remap (old_nodes_new_idx
 *                           [N]) + rg_edges_emit (sampled_edges [E][6]), all int64. */
int rg_get_neighbors_expand(const rg_graph *g, const int64_t *nodes, int64_t n_nodes, rg_frontier *in,
                            rg_frontier *out, int64_t *counts_in, int64_t *counts_out, void *ws,
                            size_t ws_bytes, void *stream);
int rg_get_neighbors_emit(const rg_graph *g, const rg_frontier *in, const rg_frontier *out, const void *ws,
                          size_t ws_bytes, int64_t n_edges, int64_t *tail_nodes, int64_t *edges,
                          int64_t *old_nodes_new_idx, void *stream);

/* ---- propagation: GNNLayer.forward (transductive/models.py:23-43) and its autograd ------------
 * Attention is factorised: as8[N][8] = hidden @ Ws^T, ar8[2R+1][8] = rela @ Wr^T,
 * aq8[n][8] = rela[q_rel] @ Wqr^T + b_qr (columns >= attn_dim are zero), w8[8] = w_alpha (zero
 * padded).  Per edge (peer p -> segment s, relation r, query q):
 *     alpha = sigmoid(b_alpha + sum_k w8[k] * relu(as8[p][k] + ar8[r][k] + aq8[q][k]))
 *     agg[s] += alpha * (hidden[p] + rela[r])               (models.py:35-39)
 * `hidden` and `as8` may both be NULL (layer 0: hidden == 0).  Segments must be grouped by the
 * OUTPUT node; the sum runs in slot order, no atomics: bit-reproducible run to run. */
int rg_edge_agg_fwd(const rg_segments *seg, int32_t hidden_dim, const float *hidden,
                    const float *as8, const float *rela, const float *ar8, const float *aq8,
                    const float *w8, const float *b_alpha, float *agg, const rg_heavy *heavy,
                    void *stream);

/* Backward of the above, segments grouped by the INPUT node p (peers are output rows):
 *   g_hidden[p]   = sum_e alpha_e * g_agg[s_e]                      (may be NULL with hidden)
 *   node_small[p] = { g_as8[p][0..7], sum_e g_l*relu(z)[0..7], sum_e g_l, 7 x 0 }   ([N][24])
 *   g_rela[r]    += sum_e alpha_e * g_agg[s_e]     (fp32 atomics into a caller-zeroed buffer)
 *   g_ar8[r]     += sum_e g_z,e                    (fp32 atomics into a caller-zeroed buffer)
 * with g_l = <g_agg[s], hidden[p]+rela[r]> * alpha(1-alpha), g_z = g_l * w8 * [z > 0].
 * grad_copies >= 1: g_rela is [grad_copies][n_table_rows][D] and g_ar8 [grad_copies][n_table_rows][8]
 * (needs seg->n_table_rows > 0 when > 1); CTAs spread their atomics over the copies and the caller
 * sums them -- the few thousand accumulator sectors otherwise serialise in L2. */
int rg_edge_agg_bwd(const rg_segments *seg, int32_t hidden_dim, const float *hidden,
                    const float *as8, const float *rela, const float *ar8, const float *aq8,
                    const float *w8, const float *b_alpha, const float *g_agg, float *g_hidden,
                    float *node_small, float *g_rela, float *g_ar8, int32_t grad_copies,
                    const rg_heavy *heavy, void *stream);

/* Which kernel variant rg_edge_agg_fwd / rg_edge_agg_bwd pick for these segments (pure host
 * function, no launch; lets the parity tests assert that a shape exercises the path it is meant
 * to): 0 = one warp per segment, 1 = eight segments per warp (short ones per 4-lane group),
 * 2 = persistent CTAs with the relation tables staged in shared memory; negative rg_status. */
int rg_edge_agg_variant(const rg_segments *seg, int32_t hidden_dim);

/* ---- node update: models.py:41 (act(W_h agg)), :81 (h0 re-index), :83 (single-step nn.GRU, gate
 *      order r,z,n; dropout :82 is the identity in eval mode), next layer's Ws_attn(hidden) and
 *      :86 W_final(hidden).  Inference only (no saved state for autograd).
 *   hidden[j] = GRU(act(W_h agg[j]), src[j] >= 0 ? h_prev[src[j]] : 0)
 *   as8[j]    = Ws_next[8][D] . hidden[j]     (optional; rows >= attn_dim of Ws_next are zero)
 *   score[j]  = W_final[D] . hidden[j]        (optional)
 * h_prev and src are both NULL at layer 0 (h0 == 0).  act: 0 identity, 1 relu, 2 tanh.
 * n_nodes is an upper bound when n_nodes_dev (device-resident true count) is given. */
int rg_node_update(int32_t hidden_dim, int64_t n_nodes, const int64_t *n_nodes_dev, const float *agg,
                   const float *h_prev,
                   const int32_t *src, const float *W_h, const float *W_ih, const float *W_hh,
                   const float *b_ih, const float *b_hh, const float *Ws_next, const float *W_final,
                   int32_t act, float *hidden, float *as8, float *score, void *stream);

/* Training forward of the same node update (tensor-core kernel, hidden_dim <= 48): applies the
 * caller's dropout mask (models.py:82; values 0 or 1/(1-p), NULL = no dropout) between act(W_h agg)
 * and the GRU and writes six planes saved[6] = {act(W_h agg), r, z, n, W_hn h0 + b_hn, h0} for the
 * backward pass (rg_node_bwd / rg_node_wgrad).  Each plane is LANE-INTERLEAVED: 32-row tiles, inside a
 * tile [chunk = col / 4][row % 32][4 floats] + 4 pad floats per chunk, i.e. element (row, col) sits at
 *   ((row / 32) * (D / 4) + col / 4) * 132 + (row % 32) * 4 + col % 4      (floats)
 * and a plane of n rows holds ceil(n / 32) * (D / 4) * 132 floats (csrc/rg_tc.cuh; coalesced for the
 * tensor-core kernels, whose lanes are node rows).  G4 (four planes) and g_pre of rg_node_bwd use the same
 * layout.  Optionally also emits the NEXT layer's attention projection as8 = Ws_next[ws_rows][D] . hidden
 * and / or the scores W_final . hidden. */
int rg_node_update_train(int32_t hidden_dim, int64_t n_nodes, const int64_t *n_nodes_dev,
                         const float *agg, const float *h_prev,
                         const int32_t *src, const float *W_h, const float *W_ih, const float *W_hh,
                         const float *b_ih, const float *b_hh, int32_t act, const float *drop_mask,
                         float *hidden, float *saved, const float *Ws_next, int32_t ws_rows,
                         const float *W_final, float *as8, float *score, void *stream);

/* Training BACKWARD of the node update (autograd of models.py:41,81-84), dense part, native:
 * rg_node_bwd (tensor cores, hidden_dim <= 48), per node row j < n_nodes:
 *   g      = g_hidden[j] (+ g_small[j][0..w_small_rows) . w_small[w_small_rows][D]) (+ g_h0_next[remap[j]])
 *            (g_small rows are g_small_stride floats apart: 8, or 24 for rg_edge_agg_bwd's node_small)
 *   G4[j]  = [g_r' | g_z' | g_n' | g_n' r]     gradients of the GRU gate pre-activations ([n][4D])
 *   g_pre[j] = ([g_r' g_z' g_n'] . W_ih) * drop_mask * act'(x)
 *   g_agg[j] = g_pre[j] . W_h                                       (input of rg_edge_agg_bwd)
 *   g_h0[j]  = g z + [g_r' g_z' g_n' r] . W_hh                      (has_h0 only)
 * g_small / w_small fold the next layer's attention projection (g_as8 . Ws_attn) or the score
 * head (g_score . W_final) into the upstream gradient; g_h0_next / remap (old_nodes_new_idx as
 * int32) fold the GRU-state path of the next layer in as a gather instead of a scatter.  Any of the
 * three upstream parts may be NULL, not all.  `saved` is what rg_node_update_train wrote.
 * rg_node_wgrad (CUDA cores, exact fp32, deterministic): the reductions over nodes
 *   out = [ dW_ih[3D][D] | dW_hh[3D][D] | dW_h[D][D] | dW_small[8][D] | column sums of G4 [4D] ]
 * with dW_ih = G4[r,z,n]^T (x * mask), dW_hh = G4[r,z,nr]^T h0, dW_h = g_pre^T agg,
 * dW_small = g_small^T hidden.  `partial` is scratch of rg_node_wgrad_ctas() * rg_node_wgrad_out_floats()
 * floats (per-CTA partial sums, added in CTA order). */
int rg_node_bwd(int32_t hidden_dim, int64_t n_nodes, const int64_t *n_nodes_dev, const float *g_hidden,
                const float *g_small, int32_t g_small_stride, const float *w_small, int32_t w_small_rows,
                const float *g_h0_next,
                const int32_t *remap, const float *saved, int64_t saved_plane_rows, const float *drop_mask,
                const float *W_h, const float *W_ih, const float *W_hh, int32_t act, int32_t has_h0, float *G4,
                float *g_pre, float *g_agg, float *g_h0, void *stream);
int32_t rg_node_wgrad_ctas(void);
int64_t rg_node_wgrad_out_floats(int32_t hidden_dim);
/* `out` != NULL: the packed vector above.  `out` == NULL: the sums go straight into the parameter
 * gradients -- g_wih / g_whh / g_bih / g_bhh (GRU, shared by all layers) are ACCUMULATED, g_wh [D][D] and
 * the first ws_rows rows of the small projection g_ws (may be NULL) are written. */
int rg_node_wgrad(int32_t hidden_dim, int64_t n_nodes, const int64_t *n_nodes_dev, const float *saved,
                  int64_t saved_plane_rows, const float *drop_mask, const float *agg, const float *hidden,
                  const float *G4, const float *g_pre, const float *g_small, int32_t g_small_stride, int32_t has_h0,
                  float *partial, float *out, float *g_wih, float *g_whh, float *g_bih, float *g_bhh, float *g_wh,
                  float *g_ws, int32_t ws_rows, void *stream);

/* ---- per-relation / per-query side of a layer (models.py:29-36) ----------------------------------
 * rg_attn_tables: ar8[r] = Wr_attn . rela[r] (n_rows x 8), aq8[b] = Wqr_attn . rela[q_rel[b]] + b_qr
 * (n_query x 8), w8 = w_alpha padded to 8; columns >= attn_dim are zero.  Weights in the reference's own
 * shapes (Wr, Wqr: [attn_dim][D]).
 * rg_attn_param_grads: every parameter gradient of that side from what rg_edge_agg_bwd (g_rela / g_ar8
 * accumulator copies) and rg_query_sum8 (q_part [n][q_slices][24]) produced, written in place:
 *   g_rela [n_rows][D] = sum_copies g_rela + g_ar8 . Wr + scatter_b(g_aq8[b] . Wqr -> row q_rel[b])
 *   g_Wr, g_Wqr [attn_dim][D], g_bqr, g_w_alpha [attn_dim], g_b_alpha [1].  Fixed summation order. */
int rg_attn_tables(int32_t hidden_dim, int32_t attn_dim, int32_t n_rows, int32_t n_query, const float *rela,
                   const float *Wr, const float *Wqr, const float *bqr, const float *w_alpha,
                   const int64_t *q_rel, float *ar8, float *aq8, float *w8, void *stream);
/* Fused training loss on the per-node scores (base_model.py:58-60) without the dense (n, n_ent) matrix:
 *   loss_q[q] = -scores_all[q][obj[q]] + logsumexp_e scores_all[q][e]
 * with the visited nodes of query q = rows [qinfo[q].base, +count) of `score` / `node_e` and every
 * unvisited entity scoring exactly 0 (models.py:87-88).  Also writes d loss_q / d score as rows
 * {g, 0 x 7} of g_small [N][8] (the upstream operand of rg_node_bwd for the last layer). */
int rg_node_loss(int32_t n_query, int32_t n_ent, const float *score, const int32_t *node_e,
                 const int32_t *qinfo, const int64_t *obj, float *loss_q, float *g_small, void *stream);
int rg_attn_param_grads(int32_t hidden_dim, int32_t attn_dim, int32_t n_rows, int32_t n_query,
                        int32_t grad_copies, const float *rela, const float *Wr, const float *Wqr,
                        const int64_t *q_rel, const float *g_rela_copies, const float *g_ar8_copies,
                        const float *q_part, int32_t q_slices, float *g_rela, float *g_Wr, float *g_Wqr,
                        float *g_bqr, float *g_w_alpha, float *g_b_alpha, void *stream);
/* Glue of the graph-captured training step (all shape-static, true counts read on the device):
 *   rg_gather_scores: backward of rg_scatter_scores, g_node[j * out_stride] = g_scores_all[b_j][e_j] (0 past n);
 *                     out_stride 8 writes whole rows {g, 0 x 7} = the g_small operand of rg_node_bwd;
 *   rg_query_sum8   : partial[q][32][0..23] = slice sums of rows24[.][0..23] over the node rows of query q
 *                     (rg_frontier.qinfo ranges); adding the 32 slices gives, in a fixed order, the
 *                     per-query attention-bias gradient (cols 0..7) and the w_alpha / b_alpha sums. */
int rg_gather_scores(int64_t n_nodes, const int64_t *n_nodes_dev, const int32_t *node_b,
                     const int32_t *node_e, const float *g_scores_all, int32_t n_ent_out, float *g_node,
                     int32_t out_stride, void *stream);
int rg_query_sum8(int32_t n_query, const float *rows24, const int32_t *qinfo, float *partial, void *stream);

/* Filtered ranking on the device: utils.cal_ranks (transductive/utils.py:7-14: rankdata 'average'
 * full rank minus 'min' filtered rank plus one, scores shifted by their row minimum + 1e-8 in
 * float32) for the answers of every query, without the (n, n_ent) D2H copy.  ans_ptr / flt_ptr are
 * CSR offsets [n+1] into ans_idx / flt_idx (entity ids; answers ascending per query so that the
 * output order equals ranks[np.nonzero(ranks)] of the reference).  ranks[ans_ptr[n]] float64. */
int rg_filtered_ranks(int32_t n_query, int32_t n_ent, const float *scores, const int32_t *ans_ptr,
                      const int32_t *ans_idx, const int32_t *flt_ptr, const int32_t *flt_idx,
                      double *ranks, void *stream);

/* scores_all[node_b[j]][node_e[j]] = score[j] for j < n (models.py:87-88; scores_all is zeroed by
 * the caller, so unvisited entities keep an exact 0). */
int rg_scatter_scores(int64_t n_nodes, const int64_t *n_nodes_dev, const int32_t *node_b,
                      const int32_t *node_e, const float *score, int32_t n_ent_out, float *scores_all,
                      void *stream);

#ifdef __cplusplus
}
#endif
#endif /* REDGNN_B200_H */
