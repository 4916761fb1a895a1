"""ctypes binding of libredgnn_b200.so (the C ABI declared in include/redgnn_b200.h).

There is no fallback: if the shared library is missing the import fails loudly, and every
compute entry point requires CUDA tensors.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("REDGNN_B200_LIB") or os.path.join(_HERE, "libredgnn_b200.so")   # env: developer builds

RG_ABI_VERSION = 1
RG_COUNTS_WORDS = 8
RG_CNT_N_IN, RG_CNT_E, RG_CNT_N_OUT, RG_CNT_ERR = 0, 1, 2, 3
GRAD_COPIES = int(os.environ.get("REDGNN_GRAD_COPIES", "8"))   # relation-gradient accumulator replicas
RG_HEAVY_CHUNK = int(os.environ.get("REDGNN_HEAVY_SUB", "256"))   # = RG_HEAVY_SUB: queue-sizing unit of the forward (env: A/B builds)
RG_HEAVY_CHUNK_BWD = int(os.environ.get("REDGNN_HEAVY_SUB_BWD", "128"))   # = RG_HEAVY_SUB_BWD: queue-sizing unit of the backward


class RgGraph(C.Structure):
    _fields_ = [("n_ent", C.c_int32), ("n_rel", C.c_int32), ("n_fact", C.c_int64),
                ("head", C.c_void_p), ("rel", C.c_void_p), ("tail", C.c_void_p),
                ("in_ptr", C.c_void_p), ("in_adj", C.c_void_p),
                ("out_ptr", C.c_void_p), ("out_adj", C.c_void_p)]


class RgFrontier(C.Structure):
    _fields_ = [("n_query", C.c_int32), ("n_ent", C.c_int32),
                ("emask", C.c_void_p), ("dict", C.c_void_p), ("qinfo", C.c_void_p)]


class RgSegments(C.Structure):
    _fields_ = [("mode", C.c_int32), ("n_ent", C.c_int32), ("n_seg", C.c_int64), ("n_seg_dev", C.c_void_p),
                ("seg_query", C.c_void_p), ("seg_ptr", C.c_void_p), ("adj", C.c_void_p),
                ("seg_ent", C.c_void_p), ("ent_ptr", C.c_void_p), ("peer_dict", C.c_void_p),
                ("peer_qinfo", C.c_void_p), ("n_table_rows", C.c_int32)]


class RgHeavy(C.Structure):
    _fields_ = [("max_chunks", C.c_int32), ("max_nodes", C.c_int32), ("counters", C.c_void_p),
                ("chunk_seg", C.c_void_p), ("chunk_idx", C.c_void_p), ("node_seg", C.c_void_p),
                ("node_base", C.c_void_p), ("node_n", C.c_void_p), ("partial", C.c_void_p)]


class RgNameTable(C.Structure):
    _fields_ = [("bytes", C.c_void_p), ("off", C.c_void_p), ("id", C.c_void_p), ("n", C.c_int64)]


RG_ERR_BAD_ARG = -1
RG_ERR_IO, RG_ERR_PARSE, RG_ERR_UNKNOWN_NAME = -5, -6, -7

# name -> (restype, argtypes); kept in one table so tests can check it against the header
SIGNATURES = {
    "rg_abi_version": (C.c_int, []),
    "rg_strerror": (C.c_char_p, [C.c_int]),
    "rg_frontier_emask_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "rg_frontier_dict_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "rg_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int64]),
    "rg_text_count_lines": (C.c_int, [C.c_char_p, C.POINTER(C.c_int64)]),
    "rg_text_parse_triples": (C.c_int, [C.c_char_p, C.POINTER(RgNameTable), C.POINTER(RgNameTable), C.c_void_p,
                                        C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int32]),
    "rg_graph_build_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int64]),
    "rg_graph_build": (C.c_int, [C.c_void_p] * 3 + [C.c_int32, C.c_int64] + [C.c_void_p] * 5 + [C.c_size_t, C.c_void_p]),
    "rg_graph_resplit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32] + [C.c_void_p] * 4),
    "rg_frontier_from_nodes": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(RgFrontier), C.c_void_p,
                                         C.c_void_p, C.c_size_t, C.c_void_p]),
    "rg_frontier_step": (C.c_int, [C.POINTER(RgGraph), C.POINTER(RgFrontier), C.POINTER(RgFrontier),
                                   C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "rg_frontier_nodes": (C.c_int, [C.POINTER(RgFrontier), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rg_frontier_remap": (C.c_int, [C.POINTER(RgFrontier), C.POINTER(RgFrontier), C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "rg_get_neighbors_expand": (C.c_int, [C.POINTER(RgGraph), C.c_void_p, C.c_int64, C.POINTER(RgFrontier),
                                          C.POINTER(RgFrontier), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                          C.c_void_p]),
    "rg_get_neighbors_emit": (C.c_int, [C.POINTER(RgGraph), C.POINTER(RgFrontier), C.POINTER(RgFrontier), C.c_void_p,
                                        C.c_size_t, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rg_node_update": (C.c_int, [C.c_int32, C.c_int64] + [C.c_void_p] * 11 + [C.c_int32] + [C.c_void_p] * 4),
    "rg_node_update_train": (C.c_int, [C.c_int32, C.c_int64] + [C.c_void_p] * 9 + [C.c_int32] + [C.c_void_p] * 4
                             + [C.c_int32] + [C.c_void_p] * 4),
    "rg_node_bwd": (C.c_int, [C.c_int32, C.c_int64] + [C.c_void_p] * 3 + [C.c_int32, C.c_void_p, C.c_int32]
                    + [C.c_void_p] * 3 + [C.c_int64]
                    + [C.c_void_p] * 4 + [C.c_int32, C.c_int32] + [C.c_void_p] * 5),
    "rg_node_wgrad_ctas": (C.c_int32, []),
    "rg_node_wgrad_out_floats": (C.c_int64, [C.c_int32]),
    "rg_node_wgrad": (C.c_int, [C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64] + [C.c_void_p] * 6
                      + [C.c_int32, C.c_int32] + [C.c_void_p] * 8 + [C.c_int32, C.c_void_p]),
    "rg_node_loss": (C.c_int, [C.c_int32, C.c_int32] + [C.c_void_p] * 7),
    "rg_attn_tables": (C.c_int, [C.c_int32] * 4 + [C.c_void_p] * 10),
    "rg_attn_param_grads": (C.c_int, [C.c_int32] * 5 + [C.c_void_p] * 7 + [C.c_int32] + [C.c_void_p] * 7),
    "rg_gather_scores": (C.c_int, [C.c_int64] + [C.c_void_p] * 4 + [C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "rg_query_sum8": (C.c_int, [C.c_int32] + [C.c_void_p] * 4),
    "rg_filtered_ranks": (C.c_int, [C.c_int32, C.c_int32] + [C.c_void_p] * 7),
    "rg_scatter_scores": (C.c_int, [C.c_int64] + [C.c_void_p] * 4 + [C.c_int32, C.c_void_p, C.c_void_p]),
    "rg_edges_emit": (C.c_int, [C.POINTER(RgGraph), C.POINTER(RgFrontier), C.POINTER(RgFrontier), C.c_void_p,
                                C.c_size_t, C.c_int64, C.c_void_p, C.c_void_p]),
    "rg_edge_agg_fwd": (C.c_int, [C.POINTER(RgSegments), C.c_int32] + [C.c_void_p] * 8
                        + [C.POINTER(RgHeavy), C.c_void_p]),
    "rg_edge_agg_variant": (C.c_int, [C.POINTER(RgSegments), C.c_int32]),
    "rg_edge_agg_bwd": (C.c_int, [C.POINTER(RgSegments), C.c_int32] + [C.c_void_p] * 12
                        + [C.c_int32, C.POINTER(RgHeavy), C.c_void_p]),
}


def _load():
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            "redgnn_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C redgnn_b200/csrc`. There is no CPU or PyTorch fallback for this path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    got = lib.rg_abi_version()
    if got != RG_ABI_VERSION:
        raise ImportError("redgnn_b200: ABI version mismatch (library %d, binding %d)" % (got, RG_ABI_VERSION))
    return lib


lib = _load()


class RgError(RuntimeError):
    pass


class Stats(object):
    """Launch bookkeeping for bench.py: how many of this library's kernels were launched, and
    (when `timing` is a list) CUDA-event pairs around the named kernels on the launching stream."""
    launches = 0
    timing = None

    @classmethod
    def timed(cls, name, meta=None):
        return _Timed(name, meta) if cls.timing is not None else _NULL


class _Null(object):
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NULL = _Null()


class _Timed(object):
    def __init__(self, name, meta):
        import torch
        self.name, self.meta = name, meta
        self.t0, self.t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def __enter__(self):
        self.t0.record()
        return self

    def __exit__(self, *a):
        self.t1.record()
        Stats.timing.append((self.name, self.meta, self.t0, self.t1))
        return False


def il_plane_floats(rows, d):
    """Floats of one lane-interleaved plane of `rows` x d (csrc/rg_tc.cuh: 32-row tiles, 132-float chunks)."""
    return -(-int(rows) // 32) * (d // 4) * 132


def check(rc):
    if rc != 0:
        raise RgError("libredgnn_b200: %s (status %d)" % (lib.rg_strerror(rc).decode(), rc))


def ptr(t):
    """Device (or host) address of a tensor, or NULL."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RgError("redgnn_b200 kernels need CUDA tensors; got a %s tensor (no CPU fallback exists)"
                          % t.device)
