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
        tri_np = np.asarray(triples, dtype=np.int64).reshape(-1, 3)
        self._validate(tri_np, n_ent, n_rel)
        tri = torch.as_tensor(tri_np)
        ids = torch.arange(n_ent, dtype=torch.int64)
        loops = torch.stack([ids, torch.full_like(ids, 2 * n_rel), ids], dim=1)
        kg = torch.cat([tri, loops], dim=0).to(device)
        self.device = torch.device(device)
        self.n_ent, self.n_rel, self.n_fact = int(n_ent), int(n_rel), int(kg.shape[0])
        h, r, t = kg[:, 0].contiguous(), kg[:, 1].contiguous(), kg[:, 2].contiguous()
        self.head, self.rel, self.tail = (x.to(torch.int32).contiguous() for x in (h, r, t))
        self.heavy_in, self.heavy_out = (0, 0), (0, 0)
        self.epoch = 0            # bumped when an in-place rebuild outgrows the heavy-queue bounds (see rebuild)
        self._ws_cur = None
        self._build_ws = None
        self._c = None
        if self.device.type == 'cuda':
            # device build through the C ABI (stable radix sort of the fact ids by tail / by head)
            i32 = lambda *shape: torch.empty(shape, dtype=torch.int32, device=self.device)
            self.in_ptr, self.in_adj = i32(n_ent + 1), i32(self.n_fact, 2)
            self.out_ptr, self.out_adj = i32(n_ent + 1), i32(self.n_fact, 2)
            self._build_csr(grow_only=False)
        else:
            # host-side construction of the same arrays (data preparation only; used by the CPU tests of
            # the loaders -- every kernel entry point still requires CUDA tensors)
            self.in_ptr, self.in_adj, in_deg = _csr(t, h, r, n_ent)
            self.out_ptr, self.out_adj, out_deg = _csr(h, t, r, n_ent)
            self._set_heavy(in_deg, out_deg, False)

    @staticmethod
    def _validate(tri, n_ent, n_rel):
        """The kernels index by these ids without bounds checks (the reference fails in scipy's
        csr_matrix on such input, load_data.py:80): refuse them here."""
        if len(tri) == 0:
            return
        if tri[:, [0, 2]].min() < 0 or tri[:, [0, 2]].max() >= n_ent:
            raise _lib.RgError("DeviceGraph: entity id outside [0, %d)" % n_ent)
        if tri[:, 1].min() < 0 or tri[:, 1].max() > 2 * n_rel:
            raise _lib.RgError("DeviceGraph: relation id outside [0, %d]" % (2 * n_rel))

    def _build_csr(self, grow_only):
        with torch.cuda.device(self.device):
            if self._build_ws is None:
                self._build_ws = torch.empty(lib.rg_graph_build_workspace_bytes(self.n_ent, self.n_fact),
                                             dtype=torch.uint8, device=self.device)
            ws = self._build_ws
            check(lib.rg_graph_build(ptr(self.head), ptr(self.rel), ptr(self.tail), self.n_ent, self.n_fact,
                                     ptr(self.in_ptr), ptr(self.in_adj), ptr(self.out_ptr), ptr(self.out_adj),
                                     ptr(ws), ws.numel(), stream_ptr()))
        if not grow_only:
            self._build_ws = None        # first build: the sort scratch (~20 B/fact) is kept only once rebuilds start
        in_deg = (self.in_ptr[1:] - self.in_ptr[:-1]).long()
        out_deg = (self.out_ptr[1:] - self.out_ptr[:-1]).long()
        self._set_heavy(in_deg, out_deg, grow_only)

    def _set_heavy(self, in_deg, out_deg, grow_only):
        """Per-query upper bounds (chunks, nodes) for the heavy-segment queues of the edge kernels.
        Captured CUDA graphs have queue buffers of these sizes baked in, so an in-place rebuild only
        ever GROWS them (with 10 % headroom) and bumps `epoch`, which is part of the capture keys."""
        ck, ckb = _lib.RG_HEAVY_CHUNK, _lib.RG_HEAVY_CHUNK_BWD     # forward pulls in-edges, backward pushes out-edges
        stats = torch.stack([((in_deg - 1) // ck).sum(), (in_deg > ck).sum(),
                             ((out_deg - 1) // ckb).sum(), (out_deg > ckb).sum()]).cpu().tolist()
        exact_in, exact_out = (int(stats[0]), int(stats[1])), (int(stats[2]), int(stats[3]))
        if not grow_only:
            self.heavy_in, self.heavy_out = exact_in, exact_out
            return
        fits = all(a <= b for a, b in zip(exact_in + exact_out, self.heavy_in + self.heavy_out))
        if not fits:
            grow = lambda new, old: tuple(max(o, n + n // 10 + 1) if n > o else o for n, o in zip(new, old))
            self.heavy_in, self.heavy_out = grow(exact_in, self.heavy_in), grow(exact_out, self.heavy_out)
            self.epoch += 1

    def rebuild(self, triples):
        """Same-size KG in place (shuffle_train re-splits keep the row count): all device addresses
        stay valid, so CUDA graphs captured on this KG keep working."""
        tri = np.asarray(triples, dtype=np.int64).reshape(-1, 3)
        if len(tri) + self.n_ent != self.n_fact or self.device.type != 'cuda':
            raise _lib.RgError("DeviceGraph.rebuild needs a CUDA graph and the same number of triples")
        self._validate(tri, self.n_ent, self.n_rel)
        dev = torch.as_tensor(tri.astype(np.int32)).to(self.device, non_blocking=True)
        n = len(tri)
        self.head[:n].copy_(dev[:, 0])
        self.rel[:n].copy_(dev[:, 1])
        self.tail[:n].copy_(dev[:, 2])
        self._build_csr(grow_only=True)

    def resplit(self, pool_dev, perm_dev, n_keep):
        """shuffle_train on the device (rg_graph_resplit + rg_graph_build), in place."""
        if 2 * n_keep + self.n_ent != self.n_fact:
            raise _lib.RgError("DeviceGraph.resplit: row count changes (%d -> %d)" % (self.n_fact, 2 * n_keep + self.n_ent))
        with torch.cuda.device(self.device):
            check(lib.rg_graph_resplit(ptr(pool_dev), ptr(perm_dev), n_keep, self.n_ent, self.n_rel, ptr(self.head),
                                       ptr(self.rel), ptr(self.tail), stream_ptr()))
        self._build_csr(grow_only=True)

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
        """Scratch for the scans of one expansion call chain.  One buffer, grown on demand and never
        freed while something may still point at it: captured CUDA graphs have its address baked in,
        so whoever captures keeps a reference to the tensor returned here (RedGNN._run_graph,
        TrainStepRunner); replacing `_ws_cur` with a larger buffer then leaves theirs alive."""
        nbytes = lib.rg_workspace_bytes(n_query, self.n_ent, self.n_fact)
        ws = self._ws_cur
        if ws is None or ws.numel() < nbytes:
            ws = self._ws_cur = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
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
        _lib.Stats.launches += 1 + self._dict_launches(n_query)      # k_set_nodes + dictionary prefix
        return fr

    def step(self, fr_in):
        """One hop.  Returns the next frontier; sizes are in fr_out.counts (device) until
        `fr_out.read_counts()` synchronises."""
        fr_out = self.new_frontier(fr_in.n_query)
        ws = self.workspace(fr_in.n_query)
        check(lib.rg_frontier_step(C.byref(self.c_struct()), C.byref(fr_in.c_struct()), C.byref(fr_out.c_struct()),
                                   ptr(fr_out.counts), ptr(ws), ws.numel(), stream_ptr()))
        _lib.Stats.launches += 3 + self._dict_launches(fr_in.n_query)   # fact_count, scan, transpose + dictionary prefix
        return fr_out

    def _dict_launches(self, n_query):
        """Kernels of the dictionary-prefix step: one for small dictionaries (k_dict_prefix_small), else four."""
        return 1 if n_query * ((self.n_ent + 31) // 32) <= 8 * 1024 else 4

    def emit_edges(self, fr_in, fr_out, n_edges):
        edges = torch.empty((n_edges, 6), dtype=torch.int64, device=self.device)
        ws = self.workspace(fr_in.n_query)
        check(lib.rg_edges_emit(C.byref(self.c_struct()), C.byref(fr_in.c_struct()), C.byref(fr_out.c_struct()),
                                ptr(ws), ws.numel(), n_edges, ptr(edges), stream_ptr()))
        _lib.Stats.launches += 1
        return edges

    def get_neighbors(self, nodes, n_query=None, spans=None):
        """Drop-in body of DataLoader.get_neighbors for this KG: returns cuda int64 tensors
        (tail_nodes[N',2], sampled_edges[E,6], old_nodes_new_idx[N]) bit-identical to the
        reference's, in the reference's order.  `spans` (bench.py): a list that receives CUDA events
        around the two device phases of the hop and its sizes."""
        if isinstance(nodes, np.ndarray):
            if n_query is None:
                n_query = int(nodes[:, 0].max()) + 1 if len(nodes) else 1
            nodes = torch.as_tensor(np.ascontiguousarray(nodes, dtype=np.int64)).to(self.device, non_blocking=True)
        else:
            nodes = nodes.to(device=self.device, dtype=torch.int64)
            if n_query is None:
                n_query = int(nodes[:, 0].max().item()) + 1 if nodes.shape[0] else 1
        nodes = nodes.contiguous()
        fr_in, fr_out = self.new_frontier(n_query), self.new_frontier(n_query)
        ws = self.workspace(n_query)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if spans is not None else None
        with torch.cuda.device(self.device):
            if ev:
                ev[0].record()
            check(lib.rg_get_neighbors_expand(C.byref(self.c_struct()), ptr(nodes), nodes.shape[0],
                                              C.byref(fr_in.c_struct()), C.byref(fr_out.c_struct()), ptr(fr_in.counts),
                                              ptr(fr_out.counts), ptr(ws), ws.numel(), stream_ptr()))
            if ev:
                ev[1].record()
            n_in, n_edges, n_out, err = fr_out.read_counts(also=fr_in)      # the one host synchronisation of a hop
            if err:
                raise _lib.RgError("get_neighbors: node out of range (batch_idx >= %d or entity >= %d)"
                                   % (n_query, self.n_ent))
            tail_nodes = torch.empty((n_out, 2), dtype=torch.int64, device=self.device)
            edges = torch.empty((n_edges, 6), dtype=torch.int64, device=self.device)
            remap = torch.empty(n_in, dtype=torch.int64, device=self.device)
            if ev:
                ev[2].record()
            check(lib.rg_get_neighbors_emit(C.byref(self.c_struct()), C.byref(fr_in.c_struct()), C.byref(fr_out.c_struct()),
                                            ptr(ws), ws.numel(), n_edges, ptr(tail_nodes), ptr(edges), ptr(remap),
                                            stream_ptr()))
            if ev:
                ev[3].record()
                spans.append((ev, n_in, n_edges, n_out))
        _lib.Stats.launches += 4 + 2 * self._dict_launches(n_query) + 3
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

    def remaps32(self, fr_out, n_nodes, n_out):
        """(old_nodes_new_idx int32 [n_nodes], inverse int32 [n_out]) in one launch; both buffers may be
        upper bounds (rows past the true counts stay unwritten / -1)."""
        fwd = torch.empty(n_nodes, dtype=torch.int32, device=self.emask.device)
        inv = torch.full((n_out,), -1, dtype=torch.int32, device=self.emask.device)
        check(lib.rg_frontier_remap(C.byref(self.c_struct()), C.byref(fr_out.c_struct()), None, ptr(fwd), ptr(inv),
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
