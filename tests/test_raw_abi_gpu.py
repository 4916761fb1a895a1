"""GPU: the ctypes stub of INTEGRATION.md section 2, verbatim in spirit -- raw ctypes against
libredgnn_b200.so with nothing from the package but the library path -- reproduces the oracle's
get_neighbors bit for bit.  Shows the C ABI is usable by a reference maintainer on its own."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import redgnn_oracle as O
from helpers import assert_expansion_equal

pytestmark = pytest.mark.gpu
LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "redgnn_b200", "libredgnn_b200.so")


class RgGraph(C.Structure):
    _fields_ = [("n_ent", C.c_int32), ("n_rel", C.c_int32), ("n_fact", C.c_int64)] + \
               [(k, C.c_void_p) for k in ("head", "rel", "tail", "in_ptr", "in_adj", "out_ptr", "out_adj")]


class RgFrontier(C.Structure):
    _fields_ = [("n_query", C.c_int32), ("n_ent", C.c_int32), ("emask", C.c_void_p), ("dict", C.c_void_p),
                ("qinfo", C.c_void_p)]


def test_integration_stub_matches_oracle(tiny_dir):
    lib = C.CDLL(LIB)
    lib.rg_frontier_emask_bytes.restype = lib.rg_frontier_dict_bytes.restype = C.c_size_t
    lib.rg_workspace_bytes.restype = C.c_size_t
    lib.rg_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int64]
    vp = C.c_void_p
    lib.rg_frontier_from_nodes.argtypes = [vp, C.c_int64, C.POINTER(RgFrontier), vp, vp, C.c_size_t, vp]
    lib.rg_frontier_step.argtypes = [C.POINTER(RgGraph), C.POINTER(RgFrontier), C.POINTER(RgFrontier), vp, vp,
                                     C.c_size_t, vp]
    lib.rg_frontier_nodes.argtypes = [C.POINTER(RgFrontier), vp, vp, vp, vp]
    lib.rg_frontier_remap.argtypes = [C.POINTER(RgFrontier), C.POINTER(RgFrontier), vp, vp, vp, vp]
    lib.rg_edges_emit.argtypes = [C.POINTER(RgGraph), C.POINTER(RgFrontier), C.POINTER(RgFrontier), vp, C.c_size_t,
                                  C.c_int64, vp, vp]
    D = O.TransductiveData(tiny_dir)
    og = D.test_graph
    kg = torch.as_tensor(og.KG.astype(np.int32)).cuda()
    head, rel, tail = (kg[:, i].contiguous() for i in range(3))
    g = RgGraph(og.n_ent, og.n_rel, og.n_fact, head.data_ptr(), rel.data_ptr(), tail.data_ptr(), None, None, None,
                None)                                            # the CSR views are only needed by the edge kernels
    nodes = np.stack([np.arange(9), np.arange(9) * 5 % og.n_ent], 1)
    want = O.get_neighbors(og, nodes)

    n = int(nodes[:, 0].max()) + 1
    st = vp(torch.cuda.current_stream().cuda_stream)
    new = lambda: (torch.empty(lib.rg_frontier_emask_bytes(n, g.n_ent) // 4, dtype=torch.int32, device='cuda'),
                   torch.empty(lib.rg_frontier_dict_bytes(n, g.n_ent) // 4, dtype=torch.int32, device='cuda'))
    (em0, d0), (em1, d1) = new(), new()
    f0 = RgFrontier(n, g.n_ent, em0.data_ptr(), d0.data_ptr(), None)
    f1 = RgFrontier(n, g.n_ent, em1.data_ptr(), d1.data_ptr(), None)
    ws = torch.empty(lib.rg_workspace_bytes(n, g.n_ent, g.n_fact), dtype=torch.uint8, device='cuda')
    c0, c1 = torch.zeros(8, dtype=torch.int64, device='cuda'), torch.zeros(8, dtype=torch.int64, device='cuda')
    nd = torch.as_tensor(nodes, dtype=torch.int64).cuda()
    assert lib.rg_frontier_from_nodes(vp(nd.data_ptr()), len(nd), C.byref(f0), vp(c0.data_ptr()), vp(ws.data_ptr()),
                                      ws.numel(), st) == 0
    assert lib.rg_frontier_step(C.byref(g), C.byref(f0), C.byref(f1), vp(c1.data_ptr()), vp(ws.data_ptr()),
                                ws.numel(), st) == 0
    n_in, n_edges, n_out = int(c0[0]), int(c1[1]), int(c1[2])          # the one host sync of the hop
    tail_nodes = torch.empty((n_out, 2), dtype=torch.int64, device='cuda')
    edges = torch.empty((n_edges, 6), dtype=torch.int64, device='cuda')
    remap = torch.empty(n_in, dtype=torch.int64, device='cuda')
    assert lib.rg_frontier_nodes(C.byref(f1), vp(tail_nodes.data_ptr()), None, None, st) == 0
    assert lib.rg_frontier_remap(C.byref(f0), C.byref(f1), vp(remap.data_ptr()), None, None, st) == 0
    assert lib.rg_edges_emit(C.byref(g), C.byref(f0), C.byref(f1), vp(ws.data_ptr()), ws.numel(), n_edges,
                             vp(edges.data_ptr()), st) == 0
    torch.cuda.synchronize()
    assert_expansion_equal((tail_nodes, edges, remap), want, "raw ctypes stub")
