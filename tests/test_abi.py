"""CPU: the C-ABI library loads and exports every symbol include/redgnn_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "redgnn_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rg_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    names = declared_functions()
    for must in ("rg_abi_version", "rg_strerror", "rg_frontier_from_nodes", "rg_frontier_step", "rg_frontier_nodes",
                 "rg_frontier_remap", "rg_edges_emit", "rg_edge_agg_fwd", "rg_edge_agg_bwd", "rg_workspace_bytes"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from redgnn_b200 import _lib
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_functions():
        assert hasattr(raw, name), "symbol %s declared in the header but not exported" % name
        assert name in _lib.SIGNATURES, "symbol %s has no ctypes signature" % name
    assert set(_lib.SIGNATURES) == set(declared_functions())


def test_version_sizes_and_errors_without_gpu():
    from redgnn_b200 import _lib
    lib = _lib.lib
    assert lib.rg_abi_version() == _lib.RG_ABI_VERSION
    assert lib.rg_strerror(0) == b"ok"
    assert b"2^31" in lib.rg_strerror(-4)
    assert lib.rg_frontier_emask_bytes(33, 100) >= 100 * 2 * 4
    assert lib.rg_frontier_dict_bytes(33, 100) >= 33 * 4 * 8
    assert lib.rg_workspace_bytes(8, 1000, 50000) > 0
    # argument validation happens before any CUDA call
    assert lib.rg_frontier_nodes(None, None, None, None, None) == -1
    assert lib.rg_edge_agg_fwd(None, 48, *([None] * 8), None, None) == -1
    assert lib.rg_scatter_scores(5, None, None, None, None, 3, None, None) == -1
    fr = _lib.RgFrontier(70000, 70000, 1, 1, None)
    assert lib.rg_frontier_nodes(ctypes.byref(fr), None, None, None, None) == -4
    # the round-2 entry points: null pointers / bad sizes are refused the same way
    assert lib.rg_node_bwd(48, 10, None, None, None, 8, None, 0, None, None, None, 10, None, None, None, None, 1, 1,
                           None, None, None, None, None) == -1
    assert lib.rg_node_wgrad(48, 10, None, None, 10, None, None, None, None, None, None, 8, 1, None, None, None, None,
                             None, None, None, None, 0, None) == -1
    assert lib.rg_attn_tables(48, 9, 10, 2, *([None] * 10)) == -1                      # attn_dim > 8
    assert lib.rg_attn_param_grads(48, 5, 10, 2, 0, *([None] * 7), 32, *([None] * 7)) == -1   # grad_copies < 1
    assert lib.rg_node_loss(0, 10, *([None] * 7)) == -1
    assert lib.rg_graph_resplit(None, None, 5, 10, 2, None, None, None, None) == -1
    assert lib.rg_get_neighbors_expand(None, None, 0, None, None, None, None, None, 0, None) == -1
    assert lib.rg_node_wgrad_ctas() >= 148 and lib.rg_node_wgrad_out_floats(48) == 7 * 48 * 48 + 8 * 48 + 4 * 48
    assert lib.rg_node_wgrad_out_floats(64) == 0                                       # tensor-core kernels: hidden_dim <= 48
    assert lib.rg_edge_agg_variant(None, 48) == -1


def test_struct_layout_matches_header():
    from redgnn_b200 import _lib
    assert ctypes.sizeof(_lib.RgGraph) == 16 + 7 * 8
    assert ctypes.sizeof(_lib.RgFrontier) == 8 + 3 * 8
    assert ctypes.sizeof(_lib.RgSegments) == 16 + 8 * 8 + 8
    assert ctypes.sizeof(_lib.RgHeavy) == 8 + 7 * 8
    assert ctypes.sizeof(_lib.RgNameTable) == 4 * 8


def test_no_cpu_fallback():
    import torch
    from redgnn_b200 import _lib
    with pytest.raises(_lib.RgError):
        _lib.require_cuda(torch.zeros(3))
