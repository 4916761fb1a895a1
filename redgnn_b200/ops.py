"""autograd wrapper of the fused edge kernels (rg_edge_agg_fwd / rg_edge_agg_bwd).

Replaces the per-edge part of GNNLayer.forward (reference Static/transductive/models.py:29-39:
hidden[sub], rela_embed(rel), rela_embed(q_rel)[r_idx], attention, alpha*(hs+hr), scatter-sum)
and what autograd would do for it (index_select / embedding backward / scatter backward).
"""
import ctypes as C

import torch

from . import _lib
from ._lib import lib, check, ptr, stream_ptr


class Segments(object):
    """Python owner of an rg_segments description (see include/redgnn_b200.h)."""

    def __init__(self, mode, n_seg, seg_query, adj, n_ent=0, seg_ptr=None, seg_ent=None, ent_ptr=None,
                 peer_dict=None, heavy_bound=(0, 0)):
        self.mode, self.n_seg, self.n_ent = int(mode), int(n_seg), int(n_ent)
        self.seg_query, self.adj = seg_query, adj
        self.seg_ptr, self.seg_ent, self.ent_ptr, self.peer_dict = seg_ptr, seg_ent, ent_ptr, peer_dict
        self.heavy_bound = (int(heavy_bound[0]), int(heavy_bound[1]))   # (max_chunks, max_nodes)
        self.n_edges = int(adj.shape[0]) if mode == 0 else None       # implicit: set by the caller if known
        self.n_seg_dev = None      # optional int64 device scalar: true segment count (n_seg = upper bound)
        self.peer_qinfo = None     # implicit: per-query {base, count} of the peer frontier
        self.n_table_rows = 0      # 2R+1 when known: lets the edge kernels stage rela / ar8 in smem
        self._c = None

    @staticmethod
    def implicit(node_b, node_e, ent_ptr, ent_adj, peer_frontier, heavy_per_query):
        """Segments = nodes (node_b[i], node_e[i]); edges pulled from the CSR row of node_e[i] and
        filtered / ranked through `peer_frontier`'s dictionary."""
        n_query = peer_frontier.n_query
        seg = Segments(1, node_b.shape[0], node_b, ent_adj, n_ent=peer_frontier.n_ent, seg_ent=node_e,
                       ent_ptr=ent_ptr, peer_dict=peer_frontier.dict,
                       heavy_bound=(heavy_per_query[0] * n_query, heavy_per_query[1] * n_query))
        seg.peer_qinfo = peer_frontier.qinfo
        return seg

    @staticmethod
    def explicit(seg_index, peer_index, rel, query, n_seg, backward=False):
        """Group an arbitrary edge list by `seg_index` (stable, so the in-segment order is the
        caller's edge order).  All arguments are int64 cuda tensors of length E.  `backward`: the
        segments feed rg_edge_agg_bwd (its heavy-segment chunk size differs)."""
        order = torch.sort(seg_index, stable=True)[1]
        deg = torch.bincount(seg_index, minlength=n_seg)
        seg_ptr = torch.zeros(n_seg + 1, dtype=torch.int64, device=seg_index.device)
        seg_ptr[1:] = torch.cumsum(deg, 0)
        adj = torch.stack([peer_index[order], rel[order]], dim=1).to(torch.int32).contiguous()
        seg_query = torch.zeros(n_seg, dtype=torch.int32, device=seg_index.device)
        seg_query[seg_index] = query.to(torch.int32)
        ck = _lib.RG_HEAVY_CHUNK_BWD if backward else _lib.RG_HEAVY_CHUNK
        if deg.numel():
            bound = (int(((deg - 1).clamp_(min=0) // ck).sum()), int((deg > ck).sum()))
        else:
            bound = (0, 0)
        return Segments(0, n_seg, seg_query, adj, seg_ptr=seg_ptr.to(torch.int32).contiguous(), heavy_bound=bound)

    def c_struct(self):
        if self._c is None:
            self._c = _lib.RgSegments(self.mode, self.n_ent, self.n_seg,
                                      self.n_seg_dev.data_ptr() if self.n_seg_dev is not None else None,
                                      self.seg_query.data_ptr(),
                                      self.seg_ptr.data_ptr() if self.seg_ptr is not None else None,
                                      self.adj.data_ptr(),
                                      self.seg_ent.data_ptr() if self.seg_ent is not None else None,
                                      self.ent_ptr.data_ptr() if self.ent_ptr is not None else None,
                                      self.peer_dict.data_ptr() if self.peer_dict is not None else None,
                                      self.peer_qinfo.data_ptr() if self.peer_qinfo is not None else None,
                                      self.n_table_rows)
        return self._c


class _Heavy(object):
    """Queue + partial rows for segments longer than RG_HEAVY_CHUNK slots (allocated per call)."""

    def __init__(self, bound, row_floats, device):
        self.max_chunks, self.max_nodes = bound
        self.struct = None
        if self.max_chunks > 0 and self.max_nodes > 0:
            i32 = lambda n: torch.empty(n, dtype=torch.int32, device=device)
            self.counters = torch.zeros(8, dtype=torch.int32, device=device)
            self.bufs = [i32(self.max_chunks), i32(self.max_chunks), i32(self.max_nodes), i32(self.max_nodes),
                         i32(self.max_nodes)]
            self.partial = torch.empty((self.max_chunks, row_floats), dtype=torch.float32, device=device)
            self.struct = _lib.RgHeavy(self.max_chunks, self.max_nodes, self.counters.data_ptr(),
                                       *[b.data_ptr() for b in self.bufs], self.partial.data_ptr())

    def ref(self):
        return C.byref(self.struct) if self.struct is not None else None


# debugging aid (tests): synchronise after every edge kernel and fail if a heavy-segment queue
# overflowed its statically computed capacity (which would silently drop chunks)
DEBUG_CHECK_HEAVY = False


def _check_heavy(heavy, what):
    if DEBUG_CHECK_HEAVY and heavy.struct is not None:
        c = heavy.counters.cpu()
        if int(c[2]) != 0 or int(c[0]) > heavy.max_chunks or int(c[1]) > heavy.max_nodes:
            raise _lib.RgError("%s: heavy-segment queue overflow (chunks %d/%d, nodes %d/%d)" % (
                what, int(c[0]), heavy.max_chunks, int(c[1]), heavy.max_nodes))


def _f32c(t):
    return None if t is None else t.detach().to(torch.float32).contiguous()


def edge_agg_forward(fwd_seg, hidden, as8, rela, ar8, aq8, w8, b_alpha):
    """Raw launcher (no autograd): returns agg [n_seg, D]."""
    _lib.require_cuda(rela, ar8, aq8, w8, b_alpha, hidden, as8)
    d = rela.shape[1]
    if fwd_seg.n_table_rows != rela.shape[0]:
        fwd_seg.n_table_rows, fwd_seg._c = rela.shape[0], None
    agg = torch.empty((fwd_seg.n_seg, d), dtype=torch.float32, device=rela.device)
    heavy = _Heavy(fwd_seg.heavy_bound, d, rela.device)
    with _lib.Stats.timed("edge_fwd", (fwd_seg, d, hidden is not None)):
        check(lib.rg_edge_agg_fwd(C.byref(fwd_seg.c_struct()), d, ptr(hidden), ptr(as8), ptr(rela), ptr(ar8),
                                  ptr(aq8), ptr(w8), ptr(b_alpha), ptr(agg), heavy.ref(), stream_ptr()))
    _lib.Stats.launches += 3 if heavy.struct is not None else 1
    _check_heavy(heavy, "rg_edge_agg_fwd")
    return agg


def edge_agg_backward(bwd_seg, hidden, as8, rela, ar8, aq8, w8, b_alpha, g_agg, n_query):
    """Raw launcher: returns (g_hidden | None, g_as8 | None, g_rela, g_ar8, g_aq8, g_w8, g_b)."""
    d = rela.shape[1]
    n_in = bwd_seg.n_seg
    if bwd_seg.n_table_rows != rela.shape[0]:
        bwd_seg.n_table_rows, bwd_seg._c = rela.shape[0], None
    dev = rela.device
    g_hidden = torch.empty((n_in, d), dtype=torch.float32, device=dev) if hidden is not None else None
    node_small = torch.empty((n_in, 24), dtype=torch.float32, device=dev)
    copies = _lib.GRAD_COPIES
    g_rela = torch.zeros((copies,) + tuple(rela.shape), dtype=torch.float32, device=dev)
    g_ar8 = torch.zeros((copies,) + tuple(ar8.shape), dtype=torch.float32, device=dev)
    heavy = _Heavy(bwd_seg.heavy_bound, d + 24, dev)
    with _lib.Stats.timed("edge_bwd", (bwd_seg, d, hidden is not None)):
        check(lib.rg_edge_agg_bwd(C.byref(bwd_seg.c_struct()), d, ptr(hidden), ptr(as8), ptr(rela), ptr(ar8),
                                  ptr(aq8), ptr(w8), ptr(b_alpha), ptr(g_agg), ptr(g_hidden), ptr(node_small),
                                  ptr(g_rela), ptr(g_ar8), copies, heavy.ref(), stream_ptr()))
    g_rela, g_ar8 = g_rela.sum(0), g_ar8.sum(0)
    _lib.Stats.launches += 3 if heavy.struct is not None else 1
    _check_heavy(heavy, "rg_edge_agg_bwd")
    g_as8 = node_small[:, :8]
    g_aq8 = torch.zeros((n_query, 8), dtype=torch.float32, device=dev)
    g_aq8.index_add_(0, bwd_seg.seg_query.long(), g_as8)
    g_w8 = node_small[:, 8:16].sum(0)
    g_b = node_small[:, 16].sum().reshape(1)
    return g_hidden, (g_as8.contiguous() if hidden is not None else None), g_rela, g_ar8, g_aq8, g_w8, g_b


ACT_CODES = {"idd": 0, "relu": 1, "tanh": 2}


def attn_tables(layer, q_rel):
    """Inference: (ar8 [2R+1, 8], aq8 [n, 8], w8 [8]) of a GNNLayer in one kernel (rg_attn_tables):
    Wr_attn(rela_embed), Wqr_attn(rela_embed[q_rel]) + bias, w_alpha -- models.py:29-36, factorised."""
    rela = layer.rela_embed.weight
    rows, d = rela.shape
    n, a = q_rel.shape[0], layer.attn_dim
    ar8 = torch.empty((rows, 8), dtype=torch.float32, device=rela.device)
    aq8 = torch.empty((n, 8), dtype=torch.float32, device=rela.device)
    w8 = torch.empty(8, dtype=torch.float32, device=rela.device)
    check(lib.rg_attn_tables(d, a, rows, n, ptr(rela), ptr(layer.Wr_attn.weight), ptr(layer.Wqr_attn.weight),
                             ptr(layer.Wqr_attn.bias), ptr(layer.w_alpha.weight), ptr(q_rel.contiguous()), ptr(ar8),
                             ptr(aq8), ptr(w8), stream_ptr()))
    _lib.Stats.launches += 1
    return ar8, aq8, w8


def scatter_scores(node_b, node_e, score, n_query, n_ent_out, n_dev=None):
    """scores_all (n, n_ent) with exact zeros for unvisited entities (rg_scatter_scores)."""
    out = torch.zeros((n_query, n_ent_out), dtype=torch.float32, device=score.device)
    check(lib.rg_scatter_scores(score.shape[0], ptr(n_dev), ptr(node_b), ptr(node_e), ptr(score), n_ent_out,
                                ptr(out), stream_ptr()))
    _lib.Stats.launches += 1
    return out


def node_update(agg, h_prev, src, W_h, gate, act_code, Ws_next8=None, W_final=None, n_dev=None):
    """Fused inference node update (rg_node_update): returns (hidden, as8 | None, score | None).
    gate: the nn.GRU module (weight_ih_l0 [3D,D], weight_hh_l0, bias_ih_l0, bias_hh_l0)."""
    _lib.require_cuda(agg, h_prev, src, W_h)
    n, d = agg.shape
    hidden = torch.empty((n, d), dtype=torch.float32, device=agg.device)
    as8 = torch.empty((n, 8), dtype=torch.float32, device=agg.device) if Ws_next8 is not None else None
    score = torch.empty((n,), dtype=torch.float32, device=agg.device) if W_final is not None else None
    with _lib.Stats.timed("node_update", (n, d)):
        check(lib.rg_node_update(d, n, ptr(n_dev), ptr(agg), ptr(h_prev), ptr(src), ptr(_f32c(W_h)), ptr(_f32c(gate.weight_ih_l0)),
                                 ptr(_f32c(gate.weight_hh_l0)), ptr(_f32c(gate.bias_ih_l0)),
                                 ptr(_f32c(gate.bias_hh_l0)), ptr(Ws_next8), ptr(_f32c(W_final)), act_code,
                                 ptr(hidden), ptr(as8), ptr(score), stream_ptr()))
    _lib.Stats.launches += 1
    return hidden, as8, score


class NodeUpdateTrain(torch.autograd.Function):
    """Training node update: hidden = GRU(dropout(act(W_h agg)), h0) with h0[j] = h_prev[src[j]].
    Forward = the tcgen05 kernel (rg_node_update_train, also saves the gates); backward = rg_node_bwd
    (tcgen05 data gradients) + rg_node_wgrad (weight-gradient reductions), no library GEMMs."""

    @staticmethod
    def forward(ctx, agg, h_prev, W_h, w_ih, w_hh, b_ih, b_hh, mask, src, remap, act_code):
        agg, h_prev, W_h, w_ih, w_hh, b_ih, b_hh, mask = (_f32c(t) for t in (agg, h_prev, W_h, w_ih, w_hh, b_ih,
                                                                             b_hh, mask))
        n, d = agg.shape
        hidden = torch.empty((n, d), dtype=torch.float32, device=agg.device)
        saved = torch.empty((6, _lib.il_plane_floats(n, d)), dtype=torch.float32, device=agg.device)   # lane-interleaved
        with _lib.Stats.timed("node_update_train", (n, d)):
            check(lib.rg_node_update_train(d, n, None, ptr(agg), ptr(h_prev), ptr(src), ptr(W_h), ptr(w_ih), ptr(w_hh),
                                           ptr(b_ih), ptr(b_hh), act_code, ptr(mask), ptr(hidden), ptr(saved),
                                           None, 0, None, None, None, stream_ptr()))
        _lib.Stats.launches += 1
        ctx.save_for_backward(agg, saved, W_h, w_ih, w_hh, mask if mask is not None else agg.new_empty(0),
                              remap if remap is not None else agg.new_empty(0, dtype=torch.int64))
        ctx.act_code, ctx.has_h0, ctx.has_mask = act_code, h_prev is not None, mask is not None
        return hidden

    @staticmethod
    def backward(ctx, g_h):
        agg, saved, W_h, w_ih, w_hh, mask, remap = ctx.saved_tensors
        n, d = agg.shape
        dev = agg.device
        g_h = g_h.to(torch.float32).contiguous()
        e = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        pl = _lib.il_plane_floats(n, d)
        G4, g_pre, g_agg = e(4, pl), e(pl), e(n, d)
        g_h0 = e(n, d) if ctx.has_h0 else None
        mk = mask if ctx.has_mask else None
        with _lib.Stats.timed("node_bwd", (n, d)):
            check(lib.rg_node_bwd(d, n, None, ptr(g_h), None, 8, None, 0, None, None, ptr(saved), n, ptr(mk), ptr(W_h),
                                  ptr(w_ih), ptr(w_hh), ctx.act_code, int(ctx.has_h0), ptr(G4), ptr(g_pre), ptr(g_agg),
                                  ptr(g_h0), stream_ptr()))
        out_floats = int(lib.rg_node_wgrad_out_floats(d))
        partial, wg = e(int(lib.rg_node_wgrad_ctas()) * out_floats), e(out_floats)
        with _lib.Stats.timed("node_wgrad", (n, d)):
            check(lib.rg_node_wgrad(d, n, None, ptr(saved), n, ptr(mk), ptr(agg), None, ptr(G4), ptr(g_pre), None, 8,
                                    int(ctx.has_h0), ptr(partial), ptr(wg), None, None, None, None, None, None, 0,
                                    stream_ptr()))
        _lib.Stats.launches += 3
        o_whh, o_wh, o_ws, o_b = 3 * d * d, 6 * d * d, 7 * d * d, 7 * d * d + 8 * d
        d_wih, d_whh, d_wh = wg[:o_whh].view(3 * d, d), wg[o_whh:o_wh].view(3 * d, d), wg[o_wh:o_ws].view(d, d)
        b4 = wg[o_b:].view(4, d)                               # column sums of g_r', g_z', g_n', g_n' r
        d_bih, d_bhh = b4[:3].reshape(-1), torch.cat([b4[0], b4[1], b4[3]])
        g_prev = g_h0.index_select(0, remap) if ctx.has_h0 else None
        return g_agg, g_prev, d_wh, d_wih, d_whh, d_bih, d_bhh, None, None, None, None


class EdgeAggregate(torch.autograd.Function):
    """agg[s] = sum_{edges e into s} alpha_e * (hidden[p_e] + rela[r_e]),
    alpha_e = sigmoid(b_alpha + sum_k w8[k] relu(as8[p_e][k] + ar8[r_e][k] + aq8[q_e][k]))."""

    @staticmethod
    def forward(ctx, hidden, as8, rela, ar8, aq8, w8, b_alpha, fwd_seg, bwd_seg):
        hidden, as8, rela, ar8, aq8, w8, b_alpha = (_f32c(t) for t in (hidden, as8, rela, ar8, aq8, w8, b_alpha))
        agg = edge_agg_forward(fwd_seg, hidden, as8, rela, ar8, aq8, w8, b_alpha)
        ctx.has_hidden = hidden is not None
        saved = [rela, ar8, aq8, w8, b_alpha] + ([hidden, as8] if ctx.has_hidden else [])
        ctx.save_for_backward(*saved)
        ctx.bwd_seg = bwd_seg
        return agg

    @staticmethod
    def backward(ctx, g_agg):
        saved = ctx.saved_tensors
        rela, ar8, aq8, w8, b_alpha = saved[:5]
        hidden, as8 = (saved[5], saved[6]) if ctx.has_hidden else (None, None)
        g_agg = g_agg.to(torch.float32).contiguous()
        g_hidden, g_as8, g_rela, g_ar8, g_aq8, g_w8, g_b = edge_agg_backward(
            ctx.bwd_seg, hidden, as8, rela, ar8, aq8, w8, b_alpha, g_agg, aq8.shape[0])
        return g_hidden, g_as8, g_rela, g_ar8, g_aq8, g_w8, g_b, None, None


def edge_aggregate(hidden, as8, rela, ar8, aq8, w8, b_alpha, fwd_seg, bwd_seg):
    return EdgeAggregate.apply(hidden, as8, rela, ar8, aq8, w8, b_alpha, fwd_seg, bwd_seg)
