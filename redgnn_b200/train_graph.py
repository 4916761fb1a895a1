"""Graph-captured training step of RED_GNN_* (forward + backward of the whole path).

The eager autograd path (layers.RedGNN._run) costs ~90 framework ops and one host read-back per
layer; at the reference's batch sizes (n_batch 3..100, Static/transductive/train.py:46-111) the
step is host-bound by a wide margin.  Here the forward of a (batch size, KG) pair is made
shape-static exactly like the inference path (upper-bound buffers of n_query * n_ent rows, true
counts read on the device), the backward is written out by hand on the same buffers (fused edge
backward, tensor-core node backward, weight-gradient reduction kernel), and both are captured
once as CUDA graphs.  A single torch.autograd.Function replays them, so `loss.backward()` and the
optimiser in the caller's loop (Static/*/base_model.py:49-70) work unchanged.

Invariant that makes upper-bound buffers safe: EVERY consumer of a per-node buffer (edge kernels,
node update, node backward, weight-gradient reduction, score scatter / gather, per-query sums) reads
the true node count from device memory and never touches a row past it, so stale rows -- including
NaN/Inf rows a diverged step left behind, which the reference loop survives by re-randomising NaN
parameters, base_model.py:65-69 -- are simply never read (tests/test_robustness_gpu.py injects one).

Forward and backward both run on the full upper-bound buffers and every kernel stops at the
device-side node count, so there is ONE captured forward and ONE captured backward per runner and no
host synchronisation inside a training step.  The dense part of the backward is native
(csrc/rg_node_bwd.cu: tcgen05 data gradients with the gate gradients as TMEM-resident A operands,
CUDA-core weight gradients); what is left to torch are the per-relation / per-query projections
(a few hundred rows) and the loss / optimiser of the caller.
"""
import ctypes as C

import torch
import torch.nn.functional as F

from . import _lib
from ._lib import lib, check, ptr, stream_ptr
from .ops import Segments, _Heavy, ACT_CODES


class TrainStepRunner(object):
    """Static buffers + the two CUDA graphs for one (model, KG, batch size)."""

    def __init__(self, model, graph, n, n_ent_out):
        self.model, self.graph, self.n, self.n_ent_out = model, graph, int(n), int(n_ent_out)
        dev = model.W_final.weight.device
        self.dev, self.d, self.a, self.n_layer = dev, model.hidden_dim, model.attn_dim, model.n_layer
        # rows of every per-node buffer: the upper bound n_query * n_ent
        self.cap = self.n * graph.n_ent
        self.act_code = ACT_CODES[model.act_name]
        self.p_drop = float(model.dropout.p)
        self.names = [k for k, _ in model.named_parameters()]
        self.params = dict(model.named_parameters())
        sizes = [self.params[k].numel() for k in self.names]
        self.flat_grad = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        self.grad_views, off = {}, 0
        for k, sz in zip(self.names, sizes):
            self.grad_views[k] = self.flat_grad[off:off + sz].view_as(self.params[k])
            off += sz
        self.sub = torch.zeros(self.n, dtype=torch.int64, device=dev)
        self.rel = torch.zeros(self.n, dtype=torch.int64, device=dev)
        self.g_out = torch.zeros((self.n, self.n_ent_out), dtype=torch.float32, device=dev)
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
        # persistent, finite-by-construction buffers (see module docstring)
        self.agg = [z(self.cap, self.d) for _ in range(self.n_layer)]
        self.hidden = [z(self.cap, self.d) for _ in range(self.n_layer)]
        # saved gate planes / G4 / g_pre: lane-interleaved planes (csrc/rg_tc.cuh)
        self.plane = _lib.il_plane_floats(self.cap, self.d)
        self.saved = [z(6, self.plane) for _ in range(self.n_layer)]
        self.as8 = [z(self.cap, 8) for _ in range(self.n_layer - 1)]       # next layer's Ws_attn(hidden), per node
        rows = 2 * model.n_rel + 1                                         # attention tables (rg_attn_tables)
        self.ar8 = [z(rows, 8) for _ in range(self.n_layer)]
        self.aq8 = [z(self.n, 8) for _ in range(self.n_layer)]
        self.w8 = [z(8) for _ in range(self.n_layer)]
        self.score_node = z(self.cap)
        self.wg_out_floats = int(lib.rg_node_wgrad_out_floats(self.d))
        self.wg_partial = torch.empty(int(lib.rg_node_wgrad_ctas()) * self.wg_out_floats, dtype=torch.float32, device=dev)
        self.ws = graph.workspace(self.n)      # keeps the expansion scratch the captured graphs point at alive
        self.kg_epoch = graph.epoch
        self.L = None
        self.scores = None
        self.version = 0
        self.bwd_graphs = {}           # one entry: (CUDAGraph, launches)
        self.bwd_loss_graph, self.g_small_loss, self.loss_q = None, None, None   # fused-loss variant (lazy)
        self.bwd_replays = 0
        self._build()

    # ------------------------------------------------------------------------------------------
    def _forward(self):
        m, g, n, d, a, cap, dev = self.model, self.graph, self.n, self.d, self.a, self.cap, self.dev
        q_sub, q_rel = self.sub, self.rel
        batch = torch.arange(n, device=dev)
        fr = g.frontier_from_nodes(torch.stack([batch, q_sub], dim=1), n)
        self.fr0 = fr
        # (an out-of-range subject is dropped by the frontier kernels and flagged in RG_CNT_ERR; clamped here so
        # that the layer-0 backward segments, which index the CSR by it, stay in bounds)
        node_b, node_e = batch.to(torch.int32), q_sub.clamp(0, g.n_ent - 1).to(torch.int32)
        hidden, n_in_dev, L = None, None, []
        gate = m.gate
        for i in range(self.n_layer):
            layer = m.gnn_layers[i]
            fr_next = g.step(fr)
            n_dev = fr_next.counts[_lib.RG_CNT_N_OUT:_lib.RG_CNT_N_OUT + 1]
            nb, ne = fr_next.nodes32(cap)
            # src: row of an output node inside the previous layer (-1 = new node) for the h0 gather of the
            # forward; remap: old_nodes_new_idx as int32 for the g_h0 gather of the backward (layers >= 1)
            remap, src = fr.remaps32(fr_next, cap, cap) if hidden is not None else (None, None)
            rela = layer.rela_embed.weight
            ar8, aq8, w8 = self.ar8[i], self.aq8[i], self.w8[i]
            check(lib.rg_attn_tables(d, a, rela.shape[0], n, ptr(rela), ptr(layer.Wr_attn.weight),
                                     ptr(layer.Wqr_attn.weight), ptr(layer.Wqr_attn.bias), ptr(layer.w_alpha.weight),
                                     ptr(q_rel), ptr(ar8), ptr(aq8), ptr(w8), stream_ptr()))
            as8 = self.as8[i - 1] if hidden is not None else None   # written by the previous node update
            fwd_seg = Segments.implicit(nb, ne, g.in_ptr, g.in_adj, fr, g.heavy_in)
            fwd_seg.n_seg_dev, fwd_seg.n_table_rows = n_dev, rela.shape[0]
            bwd_seg = Segments.implicit(node_b, node_e, g.out_ptr, g.out_adj, fr_next, g.heavy_out)
            bwd_seg.n_seg_dev = n_in_dev                         # None at layer 0: exactly n query nodes
            bwd_seg.n_table_rows = rela.shape[0]
            heavy = _Heavy(fwd_seg.heavy_bound, d, dev)
            check(lib.rg_edge_agg_fwd(C.byref(fwd_seg.c_struct()), d, ptr(hidden), ptr(as8), ptr(rela), ptr(ar8),
                                      ptr(aq8), ptr(w8), ptr(layer.w_alpha.bias), ptr(self.agg[i]), heavy.ref(),
                                      stream_ptr()))
            _lib.Stats.launches += (2 if heavy.struct is not None else 1) + 2      # + tables, node update
            mask = None
            if self.p_drop > 0:
                keep = 1.0 - self.p_drop
                mask = (torch.rand((cap, d), device=dev) < keep).to(torch.float32).div_(keep)
            last = i == self.n_layer - 1
            # the node kernel also emits the NEXT layer's attention projection Ws_attn(hidden) / the scores
            ws_next = None if last else m.gnn_layers[i + 1].Ws_attn.weight
            check(lib.rg_node_update_train(d, cap, ptr(n_dev), ptr(self.agg[i]), ptr(hidden), ptr(src),
                                           ptr(layer.W_h.weight), ptr(gate.weight_ih_l0), ptr(gate.weight_hh_l0),
                                           ptr(gate.bias_ih_l0), ptr(gate.bias_hh_l0), self.act_code, ptr(mask),
                                           ptr(self.hidden[i]), ptr(self.saved[i]), ptr(ws_next), a,
                                           ptr(m.W_final.weight) if last else None,
                                           ptr(self.as8[i]) if not last else None,
                                           ptr(self.score_node) if last else None, stream_ptr()))
            L.append(dict(fr_in=fr, fr_out=fr_next, n_dev=n_dev, nb=nb, ne=ne, src=src, remap=remap, rela=rela,
                          w8=w8, ar8=ar8, aq8=aq8, as8=as8, hidden_prev=hidden, mask=mask, bwd_seg=bwd_seg,
                          heavy=heavy, fwd_seg=fwd_seg))
            hidden, n_in_dev, fr, node_b, node_e = self.hidden[i], n_dev, fr_next, nb, ne
        scores = torch.zeros((n, self.n_ent_out), dtype=torch.float32, device=dev)
        check(lib.rg_scatter_scores(cap, ptr(n_in_dev), ptr(node_b), ptr(node_e), ptr(self.score_node), self.n_ent_out,
                                    ptr(scores), stream_ptr()))
        _lib.Stats.launches += 1
        self.L = L
        return scores

    # ------------------------------------------------------------------------------------------
    def _backward(self, caps=None, from_loss=False):
        """Hand-written backward of the whole path on the buffers of the last forward replay.  Every
        kernel stops at the device-side node counts, so nothing here depends on the true sizes
        (`caps` is accepted for compatibility and ignored): ONE captured variant, no host read-back,
        and every parameter gradient is written (or accumulated) IN PLACE in the flat gradient buffer.
        Per layer, last to first:  rg_node_bwd (tensor cores: GRU / W_h data gradients, with the next
        layer's attention-projection and GRU-state paths folded in as gathers)  ->  rg_node_wgrad
        (weight gradients)  ->  rg_edge_agg_bwd (fused edge backward)  ->  rg_query_sum8 +
        rg_attn_param_grads (per-relation / per-query parameter gradients)."""
        m, n, d, a, dev, cap = self.model, self.n, self.d, self.a, self.dev, self.cap
        st = stream_ptr
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
        e = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        gv = self.grad_views
        self.flat_grad.zero_()                                   # the GRU gradients accumulate over the layers
        last = self.L[-1]
        # upstream of the last layer: g_hidden = g_score * W_final, expressed as g_small . w_small (1 row);
        # g_score comes from the dense score gradient (gather) or straight from the fused loss kernel
        if from_loss:
            g_small = self.g_small_loss
        else:
            g_small = e(cap, 8)
            check(lib.rg_gather_scores(cap, ptr(last["n_dev"]), ptr(last["nb"]), ptr(last["ne"]), ptr(self.g_out),
                                       self.n_ent_out, ptr(g_small), 8, st()))
            _lib.Stats.launches += 1
        g_small_stride, w_small, w_rows, ws_dst = 8, m.W_final.weight, 1, gv["W_final.weight"]
        g_hid_e, g_h0_next, remap = None, None, None
        gate = m.gate
        w_ih, w_hh = gate.weight_ih_l0, gate.weight_hh_l0
        copies = _lib.GRAD_COPIES
        for i in reversed(range(self.n_layer)):
            lay, layer = self.L[i], m.gnn_layers[i]
            pre = "gnn_layers.%d." % i
            has_h0 = lay["hidden_prev"] is not None
            G4, g_pre, g_agg = e(4, self.plane), e(self.plane), e(cap, d)
            g_h0 = e(cap, d) if has_h0 else None
            check(lib.rg_node_bwd(d, cap, ptr(lay["n_dev"]), ptr(g_hid_e), ptr(g_small), g_small_stride, ptr(w_small),
                                  w_rows, ptr(g_h0_next), ptr(remap), ptr(self.saved[i]), cap, ptr(lay["mask"]),
                                  ptr(layer.W_h.weight), ptr(w_ih), ptr(w_hh), self.act_code, int(has_h0), ptr(G4),
                                  ptr(g_pre), ptr(g_agg), ptr(g_h0), st()))
            check(lib.rg_node_wgrad(d, cap, ptr(lay["n_dev"]), ptr(self.saved[i]), cap, ptr(lay["mask"]),
                                    ptr(self.agg[i]), ptr(self.hidden[i]), ptr(G4), ptr(g_pre), ptr(g_small),
                                    g_small_stride, int(has_h0), ptr(self.wg_partial), None,
                                    ptr(gv["gate.weight_ih_l0"]), ptr(gv["gate.weight_hh_l0"]),
                                    ptr(gv["gate.bias_ih_l0"]), ptr(gv["gate.bias_hh_l0"]), ptr(gv[pre + "W_h.weight"]),
                                    ptr(ws_dst), w_rows, st()))
            _lib.Stats.launches += 3
            # fused edge backward on the same implicit segments (grouped by the layer's INPUT nodes)
            hidden_prev = lay["hidden_prev"]
            bwd_seg, rela = lay["bwd_seg"], lay["rela"]
            rows = rela.shape[0]
            cap_in = cap if i > 0 else n
            node_small = e(cap_in, 24)                           # every consumer stops at the true input-node count
            g_hid_e = e(cap_in, d) if hidden_prev is not None else None
            acc = z(copies * rows * (d + 8))                     # relation-gradient accumulator copies (atomics)
            g_rela_c, g_ar8_c = acc[:copies * rows * d], acc[copies * rows * d:]
            heavy = _Heavy(bwd_seg.heavy_bound, d + 24, dev)
            check(lib.rg_edge_agg_bwd(C.byref(bwd_seg.c_struct()), d, ptr(hidden_prev), ptr(lay["as8"]), ptr(rela),
                                      ptr(lay["ar8"]), ptr(lay["aq8"]), ptr(lay["w8"]), ptr(layer.w_alpha.bias),
                                      ptr(g_agg), ptr(g_hid_e), ptr(node_small), ptr(g_rela_c), ptr(g_ar8_c), copies,
                                      heavy.ref(), st()))
            lay["heavy_bwd"] = heavy
            q_part = e(n, 32, 24)
            check(lib.rg_query_sum8(n, ptr(node_small), ptr(lay["fr_in"].qinfo), ptr(q_part), st()))
            check(lib.rg_attn_param_grads(d, a, rows, n, copies, ptr(rela), ptr(layer.Wr_attn.weight),
                                          ptr(layer.Wqr_attn.weight), ptr(self.rel), ptr(g_rela_c), ptr(g_ar8_c),
                                          ptr(q_part), 32, ptr(gv[pre + "rela_embed.weight"]),
                                          ptr(gv[pre + "Wr_attn.weight"]), ptr(gv[pre + "Wqr_attn.weight"]),
                                          ptr(gv[pre + "Wqr_attn.bias"]), ptr(gv[pre + "w_alpha.weight"]),
                                          ptr(gv[pre + "w_alpha.bias"]), st()))
            _lib.Stats.launches += 2 + (3 if heavy.struct is not None else 1)
            if hidden_prev is not None:
                # upstream of layer i-1 = edge part (g_hid_e) + attention-projection part (g_as8 . Ws_attn, g_as8 =
                # columns 0..7 of node_small) + GRU-state part (g_h0 gathered through old_nodes_new_idx): all three
                # are summed inside rg_node_bwd; rg_node_wgrad of layer i-1 turns g_as8 into this layer's Ws_attn gradient
                g_small, g_small_stride, w_small, w_rows = node_small, 24, layer.Ws_attn.weight, a
                ws_dst = gv[pre + "Ws_attn.weight"]
                g_h0_next, remap = g_h0, lay["remap"]
        # layer 0 multiplies Ws_attn by an all-zero hidden: its gradient is an explicit zero (flat_grad.zero_ above)

    # ------------------------------------------------------------------------------------------
    def _build(self):
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        full = (self.cap,) * self.n_layer
        with torch.cuda.stream(side), torch.no_grad():       # eager warm-up (lazy inits, workspaces)
            self._forward()
            self._backward(full)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.fwd_graph = torch.cuda.CUDAGraph()
        before = _lib.Stats.launches
        with torch.no_grad():
            with torch.cuda.graph(self.fwd_graph):
                self.scores = self._forward()
        # kernels of THIS library inside each graph (added to the bookkeeping at every replay)
        self.fwd_launches = _lib.Stats.launches - before
        _lib.Stats.launches = before
        self.frontiers = [lay["fr_out"] for lay in self.L]
        self.frontiers[0].source_frontier = self.fr0       # lets RedGNN.last_stats surface RG_CNT_ERR lazily
        self._capture_backward(full, warm=False)

    def node_counts(self):
        """True node count of every layer of the forward just replayed (synchronises); also surfaces the
        device-side range check of the query subjects (tensor inputs are not checked on the host)."""
        c = torch.stack([fr.counts[_lib.RG_CNT_N_OUT] for fr in self.frontiers]
                        + [self.fr0.counts[_lib.RG_CNT_ERR]]).cpu()
        if int(c[-1]):
            raise _lib.RgError("query subject out of range for this graph (n_ent=%d)" % self.graph.n_ent)
        return [int(x) for x in c[:-1]]

    def _capture_backward(self, caps, warm=True, from_loss=False):
        g = torch.cuda.CUDAGraph()
        before = _lib.Stats.launches
        with torch.no_grad(), torch.cuda.graph(g, pool=self.fwd_graph.pool()):
            self._backward(from_loss=from_loss)
        launches = _lib.Stats.launches - before
        _lib.Stats.launches = before
        if from_loss:
            self.bwd_loss_graph = (g, launches)
        else:
            self.bwd_graphs[caps] = (g, launches)

    # ---- fused loss (base_model.py:58-60) on the per-node scores: no dense (n, n_ent) round trip ----------
    def fused_loss(self, objs):
        """After a forward replay: per-query losses [n] and, as a side effect, d loss / d score in
        `g_small_loss` (the upstream operand of the last layer's rg_node_bwd)."""
        if self.g_small_loss is None:
            self.g_small_loss = torch.zeros((self.cap, 8), dtype=torch.float32, device=self.dev)
            self.loss_q = torch.zeros(self.n, dtype=torch.float32, device=self.dev)
        last = self.L[-1]
        check(lib.rg_node_loss(self.n, self.n_ent_out, ptr(self.score_node), ptr(last["ne"]), ptr(last["fr_out"].qinfo),
                               ptr(objs), ptr(self.loss_q), ptr(self.g_small_loss), stream_ptr()))
        _lib.Stats.launches += 1
        return self.loss_q

    def replay_backward_from_loss(self):
        if self.bwd_loss_graph is None:
            self._capture_backward(None, from_loss=True)
        g, launches = self.bwd_loss_graph
        g.replay()
        _lib.Stats.launches += launches
        self.bwd_replays += 1

    def replay_backward(self):
        """One captured backward (all kernels read the true node counts on the device): no host
        synchronisation anywhere in the training step.  The range check of the query subjects is the
        caller's (RedGNN checks numpy inputs on the host, tensor inputs by `check_tensor_inputs`)."""
        self.bwd_replays += 1
        (g, launches), = self.bwd_graphs.values()
        g.replay()
        _lib.Stats.launches += launches


class TrainStepFunction(torch.autograd.Function):
    """scores = forward graph; parameter gradients = backward graph.  The parameters are inputs only
    so that autograd routes the gradients to them; the kernels read them in place.

    `model.grads_in_place` (off by default): every parameter's `.grad` IS a view of the runner's flat
    gradient buffer, which the backward graph fills directly -- no per-parameter copies, and a
    multi-GPU step all-reduces that one buffer with zero packing kernels (dist.allreduce_flat).
    Valid when this forward is the parameters' only path to the loss (base_model.py:56-61 is)."""

    @staticmethod
    def forward(ctx, runner, q_sub, q_rel, *params):
        runner.sub.copy_(q_sub)
        runner.rel.copy_(q_rel)
        runner.fwd_graph.replay()
        _lib.Stats.launches += runner.fwd_launches
        runner.version += 1
        ctx.runner, ctx.version = runner, runner.version
        ctx.in_place = bool(getattr(runner.model, "grads_in_place", False))
        if ctx.in_place:
            for k in runner.names:
                runner.params[k].grad = runner.grad_views[k]
        return runner.scores.clone()

    @staticmethod
    def backward(ctx, g_scores):
        r = ctx.runner
        if ctx.version != r.version:
            raise _lib.RgError("redgnn_b200: a graph-captured training forward was followed by another forward of "
                               "the same batch size before its backward; set model.graph_train = False for "
                               "interleaved forward passes")
        r.g_out.copy_(g_scores)
        r.replay_backward()
        if ctx.in_place:                       # .grad views already hold the result
            return (None, None, None) + (None,) * len(r.names)
        flat = r.flat_grad.clone()
        out, off = [], 0
        for k in r.names:
            p = r.params[k]
            out.append(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        return (None, None, None) + tuple(out)


class TrainLossFunction(torch.autograd.Function):
    """loss = sum_q ( -score[q, obj_q] + logsumexp_e score[q, e] )  (base_model.py:58-60) of a graph-captured
    training forward, computed on the per-node scores by rg_node_loss: the dense (n, n_ent) score matrix
    is neither read nor differentiated through."""

    @staticmethod
    def forward(ctx, runner, q_sub, q_rel, objs, *params):
        runner.sub.copy_(q_sub)
        runner.rel.copy_(q_rel)
        runner.fwd_graph.replay()
        _lib.Stats.launches += runner.fwd_launches
        runner.version += 1
        ctx.runner, ctx.version = runner, runner.version
        ctx.in_place = bool(getattr(runner.model, "grads_in_place", False))
        if ctx.in_place:
            for k in runner.names:
                runner.params[k].grad = runner.grad_views[k]
        return runner.fused_loss(objs).sum()

    @staticmethod
    def backward(ctx, g_loss):
        r = ctx.runner
        if ctx.version != r.version:
            raise _lib.RgError("redgnn_b200: a graph-captured training forward was followed by another forward of "
                               "the same batch size before its backward")
        r.g_small_loss.mul_(g_loss)                    # d loss / d score, scaled by the upstream gradient (usually 1)
        r.replay_backward_from_loss()
        if ctx.in_place:
            return (None, None, None, None) + (None,) * len(r.names)
        flat = r.flat_grad.clone()
        out, off = [], 0
        for k in r.names:
            p = r.params[k]
            out.append(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        return (None, None, None, None) + tuple(out)
