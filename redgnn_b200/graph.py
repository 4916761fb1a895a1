"""Device-resident knowledge graph + frontier expansion (the get_neighbors side of the path).

Replaces, for one KG:  DataLoader.KG / M_sub (reference Static/transductive/load_data.py:76-89,
Static/inductive/load_data.py:88-98) and DataLoader.get_neighbors (transductive :106-131,
inductive :115-143).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, ptr, stream_ptr


def _csr(key, other, rel, n_ent):
    """CSR rows keyed by `key`, stable in fact order: ptr int32 [n_ent+1], adj int32 [F,2]=(other,rel)."""
    order = torch.sort(key, stable=True)[1]
    adj = torch.stack([other[order], rel[order]], dim=1).to(torch.int32).contiguous()
    deg = torch.bincount(key, minlength=n_ent)
    p = torch.zeros(n_ent + 1, dtype=torch.int64, device=key.device)
    p[1:] = torch.cumsum(deg, 0)
    return p.to(torch.int32).contiguous(), adj, deg


class DeviceGraph(object):
    """head/rel/tail int32 [F] in REFERENCE ROW ORDER ([triples ; self-loops (e, 2R, e)]) plus the
    CSR-by-tail and CSR-by-head views the fused edge kernels pull from."""

    def __init__(self, triples, n_ent, n_rel, device):
        """triples: int array [T,3] (h, r, t) WITHOUT the self-loop block (it is appended here,
        exactly like load_graph, load_data.py:76-81)."""
        tri = torch.as_tensor(np.asarray(triples, dtype=np.int64).reshape(-1, 3))
        ids = torch.arange(n_ent, dtype=torch.int64)
        loops = torch.stack([ids, torch.full_like(ids, 2 * n_rel), ids], dim=1)
        kg = torch.cat([tri, loops], dim=0).to(device)
        self.device = torch.device(device)
        self.n_ent, self.n_rel, self.n_fact = int(n_ent), int(n_rel), int(kg.shape[0])
        h, r, t = kg[:, 0].contiguous(), kg[:, 1].contiguous(), kg[:, 2].contiguous()
        self.head, self.rel, self.tail = (x.to(torch.int32).contiguous() for x in (h, r, t))
        if self.device.type == 'cuda':
            # device build through the C ABI (stable radix sort of the fact ids by tail / by head)
            i32 = lambda *shape: torch.empty(shape, dtype=torch.int32, device=self.device)
            self.in_ptr, self.in_adj = i32(n_ent + 1), i32(self.n_fact, 2)
            self.out_ptr, self.out_adj = i32(n_ent + 1), i32(self.n_fact, 2)
            with torch.cuda.device(self.device):
                ws = torch.empty(lib.rg_graph_build_workspace_bytes(n_ent, self.n_fact), dtype=torch.uint8,
                                 device=self.device)
                check(lib.rg_graph_build(ptr(self.head), ptr(self.rel), ptr(self.tail), n_ent, self.n_fact,
                                         ptr(self.in_ptr), ptr(self.in_adj), ptr(self.out_ptr), ptr(self.out_adj),
                                         ptr(ws), ws.numel(), stream_ptr()))
            in_deg = (self.in_ptr[1:] - self.in_ptr[:-1]).long()
            out_deg = (self.out_ptr[1:] - self.out_ptr[:-1]).long()
        else:
            # host-side construction of the same arrays (data preparation only; used by the CPU tests of
            # the loaders -- every kernel entry point still requires CUDA tensors)
            self.in_ptr, self.in_adj, in_deg = _csr(t, h, r, n_ent)
            self.out_ptr, self.out_adj, out_deg = _csr(h, t, r, n_ent)
        ck = _lib.RG_HEAVY_CHUNK
        # per-query upper bounds for the heavy-segment queues of the edge kernels
        self.heavy_in = (int(((in_deg - 1) // ck).sum()), int((in_deg > ck).sum()))
        self.heavy_out = (int(((out_deg - 1) // ck).sum()), int((out_deg > ck).sum()))
        self._ws = {}
        self._c = None

    # ---- ctypes views -------------------------------------------------------------------
    def c_struct(self):
        if self._c is None:
            self._c = _lib.RgGraph(self.n_ent, self.n_rel, self.n_fact, self.head.data_ptr(), self.rel.data_ptr(),
                                   self.tail.data_ptr(), self.in_ptr.data_ptr(), self.in_adj.data_ptr(),
                                   self.out_ptr.data_ptr(), self.out_adj.data_ptr())
        return self._c

    def kg_numpy(self):
        return torch.stack([self.head, self.rel, self.tail], 1).cpu().numpy().astype(np.int64)

    def workspace(self, n_query):
        ws = self._ws.get(n_query)
        if ws is None:
            nbytes = lib.rg_workspace_bytes(n_query, self.n_ent, self.n_fact)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            if len(self._ws) > 8:
                self._ws.clear()
            self._ws[n_query] = ws
        return ws

    # ---- frontier API ---------------------------------------------------------------------
    def new_frontier(self, n_query):
        return Frontier(n_query, self.n_ent, self.device)

    def frontier_from_nodes(self, nodes64, n_query):
        """nodes64: cuda int64 [N,2] (batch_idx, entity); any order, duplicates allowed."""
        _lib.require_cuda(nodes64)
        nodes64 = nodes64.contiguous()
        fr = self.new_frontier(n_query)
        ws = self.workspace(n_query)
        check(lib.rg_frontier_from_nodes(ptr(nodes64), nodes64.shape[0], C.byref(fr.c_struct()), ptr(fr.counts),
                                         ptr(ws), ws.numel(), stream_ptr()))
        _lib.Stats.launches += 5      # k_set_nodes + dict reduce / scan / apply + query info
        return fr

    def step(self, fr_in):
        """One hop.  Returns the next frontier; sizes are in fr_out.counts (device) until
        `fr_out.read_counts()` synchronises."""
        fr_out = self.new_frontier(fr_in.n_query)
        ws = self.workspace(fr_in.n_query)
        check(lib.rg_frontier_step(C.byref(self.c_struct()), C.byref(fr_in.c_struct()), C.byref(fr_out.c_struct()),
                                   ptr(fr_out.counts), ptr(ws), ws.numel(), stream_ptr()))
        _lib.Stats.launches += 7      # fact_count, scan, transpose, dict reduce / scan / apply, query info
        return fr_out

    def emit_edges(self, fr_in, fr_out, n_edges):
        edges = torch.empty((n_edges, 6), dtype=torch.int64, device=self.device)
        ws = self.workspace(fr_in.n_query)
        check(lib.rg_edges_emit(C.byref(self.c_struct()), C.byref(fr_in.c_struct()), C.byref(fr_out.c_struct()),
                                ptr(ws), ws.numel(), n_edges, ptr(edges), stream_ptr()))
        _lib.Stats.launches += 1
        return edges

    def get_neighbors(self, nodes, n_query=None):
        """Drop-in body of DataLoader.get_neighbors for this KG: returns cuda int64 tensors
        (tail_nodes[N',2], sampled_edges[E,6], old_nodes_new_idx[N]) bit-identical to the
        reference's, in the reference's order."""
        if isinstance(nodes, np.ndarray):
            if n_query is None:
                n_query = int(nodes[:, 0].max()) + 1 if len(nodes) else 1
            nodes = torch.as_tensor(np.ascontiguousarray(nodes, dtype=np.int64)).to(self.device, non_blocking=True)
        else:
            nodes = nodes.to(device=self.device, dtype=torch.int64)
            if n_query is None:
                n_query = int(nodes[:, 0].max().item()) + 1 if nodes.shape[0] else 1
        fr_in = self.frontier_from_nodes(nodes, n_query)
        fr_out = self.step(fr_in)
        n_in, n_edges, n_out, err = fr_out.read_counts(also=fr_in)
        if err:
            raise _lib.RgError("get_neighbors: node out of range (batch_idx >= %d or entity >= %d)"
                               % (n_query, self.n_ent))
        tail_nodes = fr_out.nodes64(n_out)
        remap = fr_in.remap_to(fr_out, n_in)
        edges = self.emit_edges(fr_in, fr_out, n_edges)
        return tail_nodes, edges, remap


class Frontier(object):
    """Per-layer node set (see include/redgnn_b200.h "Data layout")."""

    def __init__(self, n_query, n_ent, device):
        self.n_query, self.n_ent = int(n_query), int(n_ent)
        self.emask = torch.empty(lib.rg_frontier_emask_bytes(n_query, n_ent) // 4, dtype=torch.int32, device=device)
        self.dict = torch.empty(lib.rg_frontier_dict_bytes(n_query, n_ent) // 4, dtype=torch.int32, device=device)
        self.counts = torch.zeros(_lib.RG_COUNTS_WORDS, dtype=torch.int64, device=device)
        self.qinfo = torch.empty(2 * n_query, dtype=torch.int32, device=device)
        self.n_nodes = None      # host copies of the hop counts, filled by read_counts() / RedGNN.last_stats
        self.n_edges = None
        self._c = None

    def c_struct(self):
        if self._c is None:
            self._c = _lib.RgFrontier(self.n_query, self.n_ent, self.emask.data_ptr(), self.dict.data_ptr(),
                                      self.qinfo.data_ptr())
        return self._c

    def read_counts(self, also=None):
        """The one host synchronisation of a hop: (n_in, n_edges, n_out, err)."""
        if also is not None:
            both = torch.stack([self.counts, also.counts]).cpu()
            c, a = both[0], both[1]
            n_in, err = int(a[_lib.RG_CNT_N_IN]), int(a[_lib.RG_CNT_ERR])
            also.n_nodes = n_in
        else:
            c = self.counts.cpu()
            n_in, err = -1, 0
        self.n_nodes = int(c[_lib.RG_CNT_N_OUT])
        return n_in, int(c[_lib.RG_CNT_E]), self.n_nodes, err

    def nodes64(self, n_nodes):
        out = torch.empty((n_nodes, 2), dtype=torch.int64, device=self.emask.device)
        check(lib.rg_frontier_nodes(C.byref(self.c_struct()), ptr(out), None, None, stream_ptr()))
        _lib.Stats.launches += 1
        return out

    def nodes32(self, n_nodes):
        b = torch.empty(n_nodes, dtype=torch.int32, device=self.emask.device)
        e = torch.empty(n_nodes, dtype=torch.int32, device=self.emask.device)
        check(lib.rg_frontier_nodes(C.byref(self.c_struct()), None, ptr(b), ptr(e), stream_ptr()))
        _lib.Stats.launches += 1
        return b, e

    def remap_to(self, fr_out, n_nodes):
        out = torch.empty(n_nodes, dtype=torch.int64, device=self.emask.device)
        check(lib.rg_frontier_remap(C.byref(self.c_struct()), C.byref(fr_out.c_struct()), ptr(out), None, None,
                                    stream_ptr()))
        _lib.Stats.launches += 1
        return out

    def remap_both(self, fr_out, n_nodes, n_out):
        """(old_nodes_new_idx int64 [n_nodes], inverse int32 [n_out]) in one launch."""
        fwd = torch.empty(n_nodes, dtype=torch.int64, device=self.emask.device)
        inv = torch.full((n_out,), -1, dtype=torch.int32, device=self.emask.device)
        check(lib.rg_frontier_remap(C.byref(self.c_struct()), C.byref(fr_out.c_struct()), ptr(fwd), None, ptr(inv),
                                    stream_ptr()))
        _lib.Stats.launches += 1
        return fwd, inv

    def inverse_remap_to(self, fr_out, n_out):
        """int32 [n_out]: row of each `fr_out` node inside this (previous) frontier, -1 for new nodes."""
        inv = torch.full((n_out,), -1, dtype=torch.int32, device=self.emask.device)
        check(lib.rg_frontier_remap(C.byref(self.c_struct()), C.byref(fr_out.c_struct()), None, None, ptr(inv),
                                    stream_ptr()))
        _lib.Stats.launches += 1
        return inv
