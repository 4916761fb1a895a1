"""Host-side mirror of the reference's model surface for the propagation path.

`GNNLayer` and `RedGNN` (base of RED_GNN_trans / RED_GNN_induc) keep the reference's constructor
and forward signatures and parameter names (reference Static/transductive/models.py:5-89,
Static/inductive/models.py:5-89), so a reference `state_dict` loads unchanged and
Static/*/base_model.py / train.py run unmodified on top of them.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .ops import (Segments, edge_aggregate, edge_agg_forward, node_update, scatter_scores, ACT_CODES,
                  NodeUpdateTrain, attn_tables)

SUPPORTED_DIMS = (16, 32, 48, 64)


def _pad8(x):
    """Pad the last (attention) dimension to the kernels' fixed width of 8 with zeros."""
    a = x.shape[-1]
    return x if a == 8 else F.pad(x, (0, 8 - a))


class GNNLayer(torch.nn.Module):
    """models.py:5-43.  Same parameters; the per-edge work runs in the fused CUDA kernels."""

    def __init__(self, in_dim, out_dim, attn_dim, n_rel, act=lambda x: x):
        super(GNNLayer, self).__init__()
        if in_dim not in SUPPORTED_DIMS:
            raise ValueError("redgnn_b200: hidden_dim must be one of %s, got %d" % (SUPPORTED_DIMS, in_dim))
        if not 1 <= attn_dim <= 8:
            raise ValueError("redgnn_b200: attn_dim must be in 1..8, got %d" % attn_dim)
        self.n_rel = n_rel
        self.in_dim = in_dim
        self.out_dim = out_dim
        self.attn_dim = attn_dim
        self.act = act

        self.rela_embed = nn.Embedding(2 * n_rel + 1, in_dim)
        self.Ws_attn = nn.Linear(in_dim, attn_dim, bias=False)
        self.Wr_attn = nn.Linear(in_dim, attn_dim, bias=False)
        self.Wqr_attn = nn.Linear(in_dim, attn_dim)
        self.w_alpha = nn.Linear(attn_dim, 1)
        self.W_h = nn.Linear(in_dim, out_dim, bias=False)

    def propagate(self, q_rel, hidden, fwd_seg, bwd_seg):
        """models.py:23-43: edge aggregation followed by act(W_h .)."""
        return self.act(self.W_h(self.aggregate(q_rel, hidden, fwd_seg, bwd_seg)))

    def aggregate(self, q_rel, hidden, fwd_seg, bwd_seg):
        """Shared core.  hidden may be None (== all zeros, layer 0).  The attention projections are
        per node / per relation / per query (tiny dense maps, left to torch + autograd); everything
        per edge is one fused kernel forward and one backward."""
        rela = self.rela_embed.weight
        as8 = _pad8(self.Ws_attn(hidden)) if hidden is not None else None
        ar8 = _pad8(self.Wr_attn(rela))
        aq8 = _pad8(self.Wqr_attn(rela[q_rel]))
        if hidden is None and torch.is_grad_enabled():
            # the reference multiplies Ws_attn by an all-zero hidden at layer 0: its gradient is an
            # explicit zero (not None), which Adam's weight decay still acts on -- keep that
            aq8 = aq8 + 0.0 * self.Ws_attn.weight.sum()
        w8 = _pad8(self.w_alpha.weight).reshape(8)
        return edge_aggregate(hidden, as8, rela, ar8, aq8, w8, self.w_alpha.bias, fwd_seg, bwd_seg)

    def forward(self, q_sub, q_rel, hidden, edges, n_node, old_nodes_new_idx):
        """models.py:23-43 for a caller-provided edge list
        edges[E,6] = (batch_idx, head, rela, tail, old_idx, new_idx)."""
        _lib.require_cuda(hidden, edges)
        if edges.shape[0] == 0:      # no messages at all: scatter() of nothing is zeros (models.py:39-41)
            return self.act(self.W_h(torch.zeros((int(n_node), self.in_dim), device=hidden.device)))
        sub, rel, obj, r_idx = edges[:, 4], edges[:, 2], edges[:, 5], edges[:, 0]
        fwd_seg = Segments.explicit(obj, sub, rel, r_idx, int(n_node))
        bwd_seg = Segments.explicit(sub, obj, rel, r_idx, int(hidden.shape[0]), backward=True) \
            if torch.is_grad_enabled() else None
        return self.propagate(q_rel, hidden, fwd_seg, bwd_seg)


class RedGNN(torch.nn.Module):
    """Common body of RED_GNN_trans / RED_GNN_induc (models.py:45-89)."""

    def __init__(self, params, loader):
        super(RedGNN, self).__init__()
        self.n_layer = params.n_layer
        self.hidden_dim = params.hidden_dim
        self.attn_dim = params.attn_dim
        self.n_rel = params.n_rel
        self.loader = loader
        acts = {'relu': nn.ReLU(), 'tanh': torch.tanh, 'idd': lambda x: x}
        act = acts[params.act]
        self.act_name = params.act

        self.gnn_layers = nn.ModuleList([GNNLayer(self.hidden_dim, self.hidden_dim, self.attn_dim, self.n_rel, act=act)
                                         for _ in range(self.n_layer)])
        self.dropout = nn.Dropout(params.dropout)
        self.W_final = nn.Linear(self.hidden_dim, 1, bias=False)
        self.gate = nn.GRU(self.hidden_dim, self.hidden_dim)
        self._last_stats = None

    # single-step GRU cell with the nn.GRU parameters, in exact fp32 (cuDNN's RNN path may pick
    # TF32 tensor-core math, which breaks the 1e-4 parity bound)
    def _gate(self, x, h):
        g = self.gate
        gi = F.linear(x, g.weight_ih_l0, g.bias_ih_l0)
        gh = F.linear(h, g.weight_hh_l0, g.bias_hh_l0)
        i_r, i_z, i_n = gi.chunk(3, dim=1)
        h_r, h_z, h_n = gh.chunk(3, dim=1)
        r = torch.sigmoid(i_r + h_r)
        z = torch.sigmoid(i_z + h_z)
        n = torch.tanh(i_n + r * h_n)
        return (1.0 - z) * n + z * h

    @staticmethod
    def _to_device(x, dev):
        """numpy / list (the reference's calling convention) or a tensor already on the device."""
        if isinstance(x, torch.Tensor):
            return x.to(device=dev, dtype=torch.int64, non_blocking=True)
        return torch.as_tensor(np.asarray(x), dtype=torch.int64).to(dev, non_blocking=True)

    # upper-bound (n_query * n_ent rows) buffers allowed for the sync-free inference path, in bytes
    ASYNC_BUDGET_BYTES = 24 << 30

    @property
    def last_stats(self):
        """{'edges': [E per layer], 'nodes': N of the last layer} of the latest forward (reads the
        device-side hop counts, i.e. synchronises, on first access)."""
        st = self._last_stats
        if isinstance(st, list):
            fr0 = getattr(st[0], "source_frontier", None)       # frontier of the query subjects: holds RG_CNT_ERR
            counts = torch.stack([fr.counts for fr in st] + ([fr0.counts] if fr0 is not None else [])).cpu()
            if fr0 is not None:
                if int(counts[-1][_lib.RG_CNT_ERR]):
                    raise _lib.RgError("query subject out of range for this graph (n_ent=%d)" % fr0.n_ent)
                counts = counts[:-1]
            for fr, c in zip(st, counts):
                fr.n_nodes = int(c[_lib.RG_CNT_N_OUT])
                fr.n_edges = int(c[_lib.RG_CNT_E])
            st = {"edges": [int(c[_lib.RG_CNT_E]) for c in counts], "nodes": int(counts[-1][_lib.RG_CNT_N_OUT])}
            self._last_stats = st
        return st

    # The sync-free forward has fixed launch geometry for a given (batch size, KG), so it is captured
    # once into a CUDA graph and replayed: the ~100 launches / allocations of a forward cost one
    # cudaGraphLaunch instead of milliseconds of host time.
    use_cuda_graph = True
    check_tensor_inputs = True         # False: sync-free serving; range errors then surface in `last_stats`
    inference_in_eval = True
    fused_train_node_update = True     # autograd path: node update in the tcgen05 kernel (hidden_dim <= 48)
    MAX_CACHED_GRAPHS = 4

    def _run_graph(self, q_sub, q_rel, graph, n_ent_out):
        n = q_sub.shape[0]
        key = (n, id(graph), graph.epoch, n_ent_out, tuple(p.data_ptr() for p in self.parameters()))
        cache = self.__dict__.setdefault("_graph_cache", {})
        entry = cache.get(key)
        if entry is None:
            sub_buf, rel_buf = q_sub.clone(), q_rel.clone()
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):               # warm-up outside capture (lazy inits, workspaces)
                self._run_async(sub_buf, rel_buf, graph, n_ent_out)
            cur.wait_stream(side)
            before = _lib.Stats.launches
            cg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cg):
                out = self._run_async(sub_buf, rel_buf, graph, n_ent_out)
            # "kg" / "ws": the captured launches have the KG arrays' and the expansion scratch's addresses baked in
            entry = {"graph": cg, "sub": sub_buf, "rel": rel_buf, "out": out, "frontiers": self._last_stats,
                     "launches": _lib.Stats.launches - before, "kg": graph, "ws": graph.workspace(n)}
            _lib.Stats.launches = before
            while len(cache) >= self.MAX_CACHED_GRAPHS:
                cache.pop(next(iter(cache)))
            cache[key] = entry
        entry["sub"].copy_(q_sub)
        entry["rel"].copy_(q_rel)
        entry["graph"].replay()
        _lib.Stats.launches += entry["launches"]
        self._last_stats = entry["frontiers"]
        return entry["out"].clone()

    # Training: forward and hand-written backward of the whole path captured as two CUDA graphs
    # (train_graph.py); the eager autograd path below remains for hidden_dim 64 / profiling.
    graph_train = True
    grads_in_place = False             # see train_graph.TrainStepFunction
    MAX_CACHED_TRAIN_GRAPHS = 2

    def _run_train_graph(self, q_sub, q_rel, graph, n_ent_out, objs=None):
        from .train_graph import TrainStepRunner, TrainStepFunction, TrainLossFunction
        p_drop = float(self.dropout.p) if self.training else 0.0
        key = (q_sub.shape[0], id(graph), graph.epoch, n_ent_out, p_drop,
               tuple(p.data_ptr() for p in self.parameters()))
        cache = self.__dict__.setdefault("_train_graph_cache", {})
        runner = cache.get(key)
        if runner is None:
            for k in [k for k, r in cache.items() if r.graph is graph and r.kg_epoch != graph.epoch]:
                cache.pop(k)                     # runners whose heavy-queue bounds the rebuilt KG outgrew
            while len(cache) >= self.MAX_CACHED_TRAIN_GRAPHS:
                cache.pop(next(iter(cache)))
            saved_p, self.dropout.p = self.dropout.p, p_drop
            try:
                runner = TrainStepRunner(self, graph, q_sub.shape[0], n_ent_out)
            finally:
                self.dropout.p = saved_p
            cache[key] = runner
        params = [runner.params[k] for k in runner.names]
        if objs is not None:
            out = TrainLossFunction.apply(runner, q_sub, q_rel, objs, *params)
        else:
            out = TrainStepFunction.apply(runner, q_sub, q_rel, *params)
        self._last_stats = runner.frontiers
        self._last_runner = runner
        return out

    def loss(self, subs, rels, objs, mode=None):
        """The training loss of base_model.py:58-60,
            sum_q ( -scores[q, objs[q]] + logsumexp_e scores[q, e] ),   scores = self.forward(subs, rels, mode),
        WITHOUT the dense (n, n_ent) score matrix when the graph-captured training step is available
        (train() mode, hidden_dim <= 48): rg_node_loss evaluates it on the per-node scores (unvisited entities
        score exactly 0) and its gradient feeds the node backward directly.  Otherwise the dense formula."""
        dev = self.W_final.weight.device
        n, d = len(subs), self.hidden_dim
        graph, n_ent_out = self._graph_and_width(mode)
        ok = (torch.is_grad_enabled() and self.training and self.graph_train and d <= 48 and n > 0 and dev.type == 'cuda'
              and _lib.Stats.timing is None and n * graph.n_ent * d * 4 * 9 * self.n_layer <= self.ASYNC_BUDGET_BYTES)
        objs_t = self._to_device(objs, dev)
        if ok:
            with torch.cuda.device(dev):
                self._last_runner = None
                q_sub, q_rel = self._to_device(subs, dev), self._to_device(rels, dev)
                if not isinstance(subs, torch.Tensor):
                    s_np = np.asarray(subs)
                    if s_np.min() < 0 or s_np.max() >= graph.n_ent:
                        raise _lib.RgError("query subject out of range for this graph (n_ent=%d)" % graph.n_ent)
                return self._run_train_graph(q_sub, q_rel, graph, n_ent_out, objs=objs_t)
        scores = self.forward(subs, rels) if mode is None else self.forward(subs, rels, mode)
        pos = scores[torch.arange(n, device=dev), objs_t]
        mx = scores.max(1, keepdim=True)[0]
        return torch.sum(-pos + mx.squeeze(1) + torch.log(torch.sum(torch.exp(scores - mx), 1)))

    def flat_grad(self):
        """With `grads_in_place`: the flat gradient buffer of the latest graph-captured training step
        (every parameter's .grad is a view of it, in named_parameters() order), else None."""
        r = self.__dict__.get("_last_runner")
        return r.flat_grad if r is not None and self.grads_in_place else None

    def _run_async(self, q_sub, q_rel, graph, n_ent_out):
        """Inference without ANY host synchronisation: every per-layer buffer is sized by the upper
        bound n_query * n_ent and the kernels read the true node counts from device memory
        (the reference syncs twice per layer: models.py:78 D2H, load_data.py:119 H2D)."""
        dev = q_sub.device
        n, d = q_sub.shape[0], self.hidden_dim
        cap = n * graph.n_ent
        batch = torch.arange(n, device=dev)
        fr = graph.frontier_from_nodes(torch.stack([batch, q_sub], dim=1), n)
        hidden, as8, scores, src = None, None, None, None
        frontiers = []
        for i in range(self.n_layer):
            fr_next = graph.step(fr)
            n_dev = fr_next.counts[_lib.RG_CNT_N_OUT:_lib.RG_CNT_N_OUT + 1]
            nb, ne = fr_next.nodes32(cap)
            fwd_seg = Segments.implicit(nb, ne, graph.in_ptr, graph.in_adj, fr, graph.heavy_in)
            fwd_seg.n_seg_dev = n_dev
            fwd_seg.frontier = fr_next                    # lets bench.py resolve E / N' after the fact
            layer = self.gnn_layers[i]
            rela = layer.rela_embed.weight
            ar8, aq8, w8 = attn_tables(layer, q_rel)         # per-relation / per-query attention tables, one kernel
            agg = edge_agg_forward(fwd_seg, hidden, as8, rela, ar8, aq8, w8, layer.w_alpha.bias)
            last = i == self.n_layer - 1
            ws_next = None if last else F.pad(self.gnn_layers[i + 1].Ws_attn.weight,
                                              (0, 0, 0, 8 - self.attn_dim)).contiguous()
            src = fr.inverse_remap_to(fr_next, cap) if hidden is not None else None
            hidden, as8, scores = node_update(agg, hidden, src, layer.W_h.weight, self.gate,
                                              ACT_CODES[self.act_name], ws_next,
                                              self.W_final.weight if last else None, n_dev=n_dev)
            if i == 0:
                fr_next.source_frontier = fr
            frontiers.append(fr_next)
            fr = fr_next
        self._last_stats = frontiers                      # resolved lazily by `last_stats`
        return scatter_scores(nb, ne, scores, n, n_ent_out, n_dev=n_dev)

    def _run(self, subs, rels, graph, n_ent_out):
        dev = self.W_final.weight.device
        if dev.type != 'cuda':
            raise _lib.RgError("redgnn_b200: the model must live on a CUDA device (call .cuda()); no CPU path exists")
        if len(subs) == 0:
            return torch.zeros((0, n_ent_out), device=dev)
        # the library launches on the CURRENT device / stream: make the model's device current
        with torch.cuda.device(dev):
            return self._run_on_device(subs, rels, graph, n_ent_out, dev)

    def _run_on_device(self, subs, rels, graph, n_ent_out, dev):
        self._last_runner = None           # set again by the graph-captured training path only
        n = len(subs)
        d = self.hidden_dim
        if not isinstance(subs, torch.Tensor):
            s = np.asarray(subs)
            if len(s) and (s.min() < 0 or s.max() >= graph.n_ent):
                raise _lib.RgError("query subject out of range for this graph (n_ent=%d)" % graph.n_ent)
        q_sub, q_rel = self._to_device(subs, dev), self._to_device(rels, dev)
        if isinstance(subs, torch.Tensor) and self.check_tensor_inputs and n > 0:
            # the kernels drop an out-of-range (query, entity) silently (all-zero scores): one 2-scalar
            # read-back keeps the eager path's error behaviour for tensor inputs too
            lo, hi = torch.aminmax(q_sub)
            lo, hi = torch.stack([lo, hi]).tolist()
            if lo < 0 or hi >= graph.n_ent:
                raise _lib.RgError("query subject out of range for this graph (n_ent=%d)" % graph.n_ent)
        # eval(): the reference's evaluate() never differentiates (base_model.py:106 takes .data), so an
        # eval-mode forward runs the inference kernels and returns a tensor WITHOUT autograd history
        # unless `inference_in_eval` is switched off; train() mode always keeps autograd.
        need_grad = torch.is_grad_enabled() and (self.training or not self.inference_in_eval)
        if not need_grad and not (self.training and self.dropout.p > 0) and n > 0 \
                and n * graph.n_ent * (d + 10) * 4 * 4 <= self.ASYNC_BUDGET_BYTES:
            with torch.no_grad():    # eval() with autograd on still lands here (inference_in_eval)
                if self.use_cuda_graph and _lib.Stats.timing is None:
                    return self._run_graph(q_sub, q_rel, graph, n_ent_out)
                return self._run_async(q_sub, q_rel, graph, n_ent_out)
        if need_grad and self.graph_train and d <= 48 and n > 0 and _lib.Stats.timing is None \
                and n * graph.n_ent * d * 4 * 9 * self.n_layer <= self.ASYNC_BUDGET_BYTES:
            return self._run_train_graph(q_sub, q_rel, graph, n_ent_out)

        batch = torch.arange(n, device=dev)
        fr = graph.frontier_from_nodes(torch.stack([batch, q_sub], dim=1), n)
        node_b, node_e = batch.to(torch.int32), q_sub.to(torch.int32)
        n_nodes = n
        # inference: per-node work (W_h, act, h0 re-index, GRU, next Ws_attn, W_final) runs in the fused
        # node kernel; with autograd on (or active dropout) it stays torch-composed so autograd sees it
        fused = not need_grad and not (self.training and self.dropout.p > 0)
        h0 = None if fused else torch.zeros((n, d), device=dev)
        hidden, as8, scores = None, None, None
        edges_per_layer = []
        for i in range(self.n_layer):
            fr_next = graph.step(fr)
            _, n_edges, n_next, err = fr_next.read_counts(also=fr if i == 0 else None)
            if err:
                raise _lib.RgError("query subject out of range for this graph (n_ent=%d)" % graph.n_ent)
            nb, ne = fr_next.nodes32(n_next)
            fwd_seg = Segments.implicit(nb, ne, graph.in_ptr, graph.in_adj, fr, graph.heavy_in)
            fwd_seg.n_edges = n_edges
            layer = self.gnn_layers[i]
            if fused:
                rela = layer.rela_embed.weight
                ar8 = _pad8(layer.Wr_attn(rela))
                aq8 = _pad8(layer.Wqr_attn(rela[q_rel]))
                w8 = _pad8(layer.w_alpha.weight).reshape(8)
                agg = edge_agg_forward(fwd_seg, hidden, as8, rela.contiguous(), ar8.contiguous(), aq8.contiguous(),
                                       w8.contiguous(), layer.w_alpha.bias)
                last = i == self.n_layer - 1
                ws_next = None if last else F.pad(self.gnn_layers[i + 1].Ws_attn.weight,
                                                  (0, 0, 0, 8 - self.attn_dim)).contiguous()
                src = fr.inverse_remap_to(fr_next, n_next) if hidden is not None else None
                hidden, as8, scores = node_update(agg, hidden, src, layer.W_h.weight, self.gate,
                                                  ACT_CODES[self.act_name], ws_next,
                                                  self.W_final.weight if last else None)
            else:
                bwd_seg = None
                if need_grad:
                    bwd_seg = Segments.implicit(node_b, node_e, graph.out_ptr, graph.out_adj, fr_next, graph.heavy_out)
                    bwd_seg.n_edges = n_edges
                if need_grad and d <= 48 and self.fused_train_node_update:
                    # per-node work in the tcgen05 kernel (gates saved for an explicit backward)
                    remap, src = fr.remap_both(fr_next, n_nodes, n_next)
                    agg = layer.aggregate(q_rel, hidden, fwd_seg, bwd_seg)
                    mask = None
                    if self.training and self.dropout.p > 0:
                        keep = 1.0 - self.dropout.p
                        mask = (torch.rand((n_next, d), device=dev) < keep).to(torch.float32).div_(keep)
                    g = self.gate
                    hidden = NodeUpdateTrain.apply(agg, hidden, layer.W_h.weight, g.weight_ih_l0, g.weight_hh_l0,
                                                   g.bias_ih_l0, g.bias_hh_l0, mask,
                                                   src if hidden is not None else None,
                                                   remap if hidden is not None else None, ACT_CODES[self.act_name])
                else:
                    remap = fr.remap_to(fr_next, n_nodes)
                    hidden = layer.propagate(q_rel, hidden, fwd_seg, bwd_seg)
                    h0 = torch.zeros((n_next, d), device=dev).index_copy_(0, remap, h0)
                    hidden = self.dropout(hidden)
                    hidden = self._gate(hidden, h0)
                    h0 = hidden
            fr, node_b, node_e, n_nodes = fr_next, nb, ne, n_next
            edges_per_layer.append(n_edges)

        if not fused:
            scores = self.W_final(hidden).squeeze(-1)
        scores_all = torch.zeros((n, n_ent_out), device=dev)
        scores_all[node_b.long(), node_e.long()] = scores
        self._last_stats = {"edges": edges_per_layer, "nodes": n_nodes}
        return scores_all
