"""ctypes binding of the C ABI in include/ggnn_b200.h (libggnn_b200.so, sm_100a).

There is deliberately no fallback: if the CUDA library is missing or fails to load, importing
this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GGNN_B200_LIB") or os.path.join(_HERE, "libggnn_b200.so")

EUCLIDEAN, COSINE = 0, 1
L = 4
ERR_INVALID, ERR_UNSUPPORTED = 10001, 10002


class GraphConfig(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in
                ("N", "D", "KBuild", "KF", "G", "S", "S0", "S0_off", "SG", "SG_off", "N_all", "ST_all")] + \
               [(n, C.c_uint32 * L) for n in ("Bs", "Ns", "Ns_offsets", "STs_offsets")]


class GraphOffsets(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in ("graph", "translation", "selection", "nn1_stats", "total")]


class QueryShape(C.Structure):
    _fields_ = [("cache_size", C.c_uint32), ("sorted_size", C.c_uint32), ("block_dim_x", C.c_uint32)]


class QueryParams(C.Structure):
    _fields_ = [
        ("D", C.c_uint32), ("measure", C.c_int32), ("KQuery", C.c_uint32), ("sorted_size", C.c_uint32),
        ("cache_size", C.c_uint32), ("block_dim_x", C.c_uint32), ("tau_query", C.c_float),
        ("max_iterations", C.c_uint32), ("N_base", C.c_int32), ("KBuild", C.c_uint32),
        ("num_starting_points", C.c_uint32),
        ("d_base", C.c_void_p), ("d_query", C.c_void_p), ("d_graph", C.c_void_p),
        ("d_starting_points", C.c_void_p), ("d_nn1_stats", C.c_void_p), ("d_query_results", C.c_void_p),
        ("d_query_results_dists", C.c_void_p), ("d_stats", C.c_void_p),
        ("shards_per_gpu", C.c_uint32), ("on_gpu_shard_id", C.c_uint32), ("d_work_counter", C.c_void_p),
        # fused shard-merge exchange (all zero = off)
        ("n_scatter", C.c_uint32), ("scatter_slot", C.c_uint32), ("scatter_rows", C.c_uint32),
        ("scatter_dists_offset", C.c_size_t), ("d_scatter_dst", C.c_void_p), ("d_scatter_flags", C.c_void_p),
        ("d_scatter_done", C.c_void_p),
        ("base_type", C.c_uint32),   # 0 fp32, 1 native uint8 rows (d_base / d_query are uint8 then)
        ("d_base_interleaved", C.c_void_p),   # optional interleaved copy of an fp32 base (ggnn_b200_interleave_rows)
    ]


class BfQueryParams(C.Structure):
    _fields_ = [
        ("D", C.c_uint32), ("measure", C.c_int32), ("KQuery", C.c_uint32), ("N_base", C.c_int32),
        ("d_base", C.c_void_p), ("d_query", C.c_void_p), ("d_query_results", C.c_void_p),
        ("d_query_results_dists", C.c_void_p), ("d_workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class BuildPassStats(C.Structure):
    _fields_ = [("kernel", C.c_uint32), ("layer_top", C.c_uint32), ("layer_btm", C.c_uint32), ("points", C.c_uint32),
                ("pops", C.c_uint64), ("dists", C.c_uint64), ("ms", C.c_float)]


EXPORTS = [
    "ggnn_b200_last_error", "ggnn_b200_version", "ggnn_b200_graph_config_init", "ggnn_b200_graph_blob_bytes",
    "ggnn_b200_graph_blob_offsets", "ggnn_b200_build_scratch_bytes", "ggnn_b200_query_shape_init",
    "ggnn_b200_query", "ggnn_b200_bf_query", "ggnn_b200_bf_query_workspace_bytes", "ggnn_b200_top", "ggnn_b200_nn1_stats", "ggnn_b200_select",
    "ggnn_b200_merge", "ggnn_b200_sym", "ggnn_b200_sym_buffer_merge", "ggnn_b200_build_graph",
    "ggnn_b200_merge_topk", "ggnn_b200_widen_u8",
    "ggnn_b200_ipc_alloc", "ggnn_b200_ipc_open", "ggnn_b200_ipc_close", "ggnn_b200_ipc_free", "ggnn_b200_peer_enable",
    "ggnn_b200_wait_flag", "ggnn_b200_refine_graph", "ggnn_b200_rng_create", "ggnn_b200_rng_fill_build", "ggnn_b200_rng_destroy",
    "ggnn_b200_build_stats_begin", "ggnn_b200_build_stats_end", "ggnn_b200_interleave_rows",
    "ggnn_b200_bf_query_u8", "ggnn_b200_bf_query_u8_workspace_bytes", "ggnn_b200_debug_i8_pack", "ggnn_b200_debug_i8_mma",
]

_lib = None


class GGNNError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        l = C.CDLL(LIB_PATH)
        l.ggnn_b200_last_error.restype = C.c_char_p
        l.ggnn_b200_version.restype = C.c_char_p
        l.ggnn_b200_graph_blob_bytes.restype = C.c_size_t
        l.ggnn_b200_build_scratch_bytes.restype = C.c_size_t
        l.ggnn_b200_graph_blob_offsets.restype = None
        vp, u32, i32, f32, sz = C.c_void_p, C.c_uint32, C.c_int32, C.c_float, C.c_size_t
        cfgp = C.POINTER(GraphConfig)
        l.ggnn_b200_graph_config_init.argtypes = [cfgp, u32, u32, u32]
        l.ggnn_b200_graph_blob_bytes.argtypes = [cfgp]
        l.ggnn_b200_graph_blob_offsets.argtypes = [cfgp, C.POINTER(GraphOffsets)]
        l.ggnn_b200_build_scratch_bytes.argtypes = [cfgp]
        l.ggnn_b200_query_shape_init.argtypes = [C.POINTER(QueryShape), u32, u32, u32]
        l.ggnn_b200_query.argtypes = [C.POINTER(QueryParams), u32, vp]
        l.ggnn_b200_bf_query.argtypes = [C.POINTER(BfQueryParams), u32, vp]
        l.ggnn_b200_bf_query_workspace_bytes.restype = C.c_size_t
        l.ggnn_b200_bf_query_workspace_bytes.argtypes = [u32, i32, u32, u32, u32]
        l.ggnn_b200_top.argtypes = [cfgp, vp, i32, u32, vp, vp, vp]
        l.ggnn_b200_nn1_stats.argtypes = [vp, u32, vp, vp, vp]
        l.ggnn_b200_select.argtypes = [cfgp, u32, vp, vp, vp, vp]
        l.ggnn_b200_merge.argtypes = [cfgp, vp, i32, f32, u32, u32, vp, vp, vp, vp]
        l.ggnn_b200_sym.argtypes = [cfgp, vp, i32, f32, u32, vp, vp, vp, vp]
        l.ggnn_b200_sym_buffer_merge.argtypes = [cfgp, u32, vp, vp, vp, vp]
        l.ggnn_b200_build_graph.argtypes = [cfgp, vp, i32, f32, u32, vp, vp, vp, sz, vp]
        l.ggnn_b200_merge_topk.argtypes = [vp, vp, u32, sz, sz, u32, u32, u32, C.c_int64, vp, vp, vp]
        l.ggnn_b200_widen_u8.argtypes = [vp, vp, sz, vp]
        l.ggnn_b200_ipc_alloc.argtypes = [sz, C.POINTER(vp), C.c_char_p]
        l.ggnn_b200_ipc_open.argtypes = [C.c_char_p, C.POINTER(vp)]
        l.ggnn_b200_ipc_close.argtypes = [vp]
        l.ggnn_b200_ipc_free.argtypes = [vp]
        l.ggnn_b200_peer_enable.argtypes = [C.c_int]
        l.ggnn_b200_wait_flag.argtypes = [vp, u32, u32, vp, vp]
        l.ggnn_b200_refine_graph.argtypes = [cfgp, vp, i32, f32, vp, vp, sz, vp]
        l.ggnn_b200_rng_create.argtypes = [C.POINTER(vp), C.c_uint64]
        l.ggnn_b200_rng_fill_build.argtypes = [vp, cfgp, vp, vp]
        l.ggnn_b200_rng_destroy.argtypes = [vp]
        l.ggnn_b200_rng_destroy.restype = None
        l.ggnn_b200_interleave_rows.argtypes = [vp, vp, u32, u32, vp]
        l.ggnn_b200_bf_query_u8.argtypes = [C.POINTER(BfQueryParams), u32, vp]
        l.ggnn_b200_bf_query_u8_workspace_bytes.restype = C.c_size_t
        l.ggnn_b200_bf_query_u8_workspace_bytes.argtypes = [u32, i32, u32, u32, u32]
        l.ggnn_b200_debug_i8_pack.argtypes = [vp, u32, u32, u32, vp, vp, vp]
        l.ggnn_b200_debug_i8_mma.argtypes = [vp, vp, u32, vp, vp]
        l.ggnn_b200_build_stats_begin.argtypes = []
        l.ggnn_b200_build_stats_end.argtypes = [C.POINTER(BuildPassStats), u32, C.POINTER(u32)]
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().ggnn_b200_last_error().decode()
        if rc == ERR_INVALID:
            raise ValueError(msg)
        if rc == ERR_UNSUPPORTED:
            raise NotImplementedError(msg)
        raise GGNNError(f"CUDA error {rc}: {msg}")


def graph_config(N, D, KBuild):
    c = GraphConfig()
    check(lib().ggnn_b200_graph_config_init(C.byref(c), N, D, KBuild))
    return c


def graph_offsets(cfg):
    o = GraphOffsets()
    lib().ggnn_b200_graph_blob_offsets(C.byref(cfg), C.byref(o))
    return o


def query_shape(D, KQuery, max_iterations):
    s = QueryShape()
    check(lib().ggnn_b200_query_shape_init(C.byref(s), D, KQuery, max_iterations))
    return s
