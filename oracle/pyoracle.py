"""ctypes wrapper around oracle/libggnn_oracle.so (the CPU restatement of the reference).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (ggnn_b200) never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libggnn_oracle.so")

EUCLIDEAN, COSINE = 0, 1
L = 4


class GraphConfig(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in
                ("N", "D", "KBuild", "KF", "G", "S", "S0", "S0_off", "SG", "SG_off", "N_all", "ST_all")] + \
               [(n, C.c_uint32 * L) for n in ("Bs", "Ns", "Ns_offsets", "STs_offsets")]

    def as_dict(self):
        d = {}
        for n, t in self._fields_:
            v = getattr(self, n)
            d[n] = list(v) if not isinstance(v, int) else v
        return d


class GraphView(C.Structure):
    _fields_ = [("graph", C.c_void_p), ("translation", C.c_void_p), ("selection", C.c_void_p),
                ("nn1_stats", C.c_void_p)]


class QueryLaunch(C.Structure):
    _fields_ = [("cache_size", C.c_uint32), ("sorted_size", C.c_uint32), ("block_dim_x", C.c_uint32)]


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("ggnn_oracle.c", "ggnn_oracle.h")]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "libggnn_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_graph_blob_bytes.restype = C.c_size_t
        _lib.orc_distance.restype = C.c_float
        _lib.orc_bf_block_dim.restype = C.c_uint32
        _lib.orc_cache_create.restype = C.c_void_p
        _lib.orc_cache_pop.restype = C.c_int32
    return _lib


def set_threads(n):
    lib().orc_set_threads(C.c_int(int(n)))


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def graph_config(N, D, KBuild):
    c = GraphConfig()
    lib().orc_graph_config_init(C.byref(c), C.c_uint32(N), C.c_uint32(D), C.c_uint32(KBuild))
    return c


def blob_bytes(cfg):
    return lib().orc_graph_blob_bytes(C.byref(cfg))


class Graph:
    """numpy views into a reference-layout graph blob (graph.cpp:48-84)."""

    def __init__(self, cfg, blob=None):
        self.cfg = cfg
        nbytes = blob_bytes(cfg)
        if blob is None:
            blob = np.zeros(nbytes, dtype=np.uint8)
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        assert blob.nbytes == nbytes, (blob.nbytes, nbytes)
        self.blob = blob
        K = cfg.KBuild
        o = cfg.N_all * K * 4
        self.graph = blob[:o].view(np.int32).reshape(cfg.N_all, K)
        self.translation = blob[o:o + cfg.ST_all * 4].view(np.int32)
        self.selection = blob[o + cfg.ST_all * 4:o + cfg.ST_all * 8].view(np.int32)
        self.nn1_stats = blob[o + cfg.ST_all * 8:o + cfg.ST_all * 8 + 8].view(np.float32)
        self.view = GraphView()
        lib().orc_graph_view_init(C.byref(self.view), C.byref(cfg), _p(blob))

    def layer_graph(self, l):
        c = self.cfg
        return self.graph[c.Ns_offsets[l]:c.Ns_offsets[l] + c.Ns[l]]

    def layer_translation(self, l):
        c = self.cfg
        return self.translation[c.STs_offsets[l]:c.STs_offsets[l] + c.Ns[l]] if l else None

    def layer_selection(self, l):
        c = self.cfg
        return self.selection[c.STs_offsets[l]:c.STs_offsets[l] + c.Ns[l]] if l else None

    def start_points(self):
        return self.layer_translation(L - 1)


def query_launch_params(D, KQuery, max_iters):
    o = QueryLaunch()
    rc = lib().orc_query_launch_params(C.c_uint32(D), C.c_uint32(KQuery), C.c_uint32(max_iters), C.byref(o))
    if rc:
        raise ValueError("unsupported query parameters")
    return o.cache_size, o.sorted_size, o.block_dim_x


def bf_block_dim(D):
    return lib().orc_bf_block_dim(C.c_uint32(D))


def construction_config(D, min_block):
    b, i = C.c_uint32(), C.c_uint32()
    lib().orc_construction_config(C.c_uint32(D), C.c_uint32(min_block), C.byref(b), C.byref(i))
    return b.value, i.value


def distance(q, b, measure=EUCLIDEAN, vblock=32, items=4):
    q, b = _f32(q), _f32(b)
    return float(lib().orc_distance(_p(q), _p(b), C.c_uint32(q.size), C.c_int(measure), C.c_uint32(vblock),
                                    C.c_uint32(items)))


def bf_query(base, query, K, measure=EUCLIDEAN):
    base, query = _f32(base), _f32(query)
    ids = np.empty((query.shape[0], K), np.int32)
    dists = np.empty((query.shape[0], K), np.float32)
    lib().orc_bf_query(_p(base), C.c_uint32(base.shape[0]), _p(query), C.c_uint32(query.shape[0]),
                       C.c_uint32(base.shape[1]), C.c_uint32(K), C.c_int(measure), _p(ids), _p(dists))
    return ids, dists


def query(base, query, graph0, start_points, nn1_stats, KQuery, tau_query, max_iterations=400,
          measure=EUCLIDEAN, shards_per_gpu=1, on_gpu_shard_id=0, out=None, with_stats=False):
    base, query = _f32(base), _f32(query)
    graph0, start_points, nn1_stats = _i32(graph0), _i32(start_points), _f32(nn1_stats)
    Nq = query.shape[0]
    if out is None:
        ids = np.empty((Nq, KQuery * shards_per_gpu), np.int32)
        dists = np.empty((Nq, KQuery * shards_per_gpu), np.float32)
    else:
        ids, dists = out
    stats = np.zeros((Nq, 2), np.uint32) if with_stats else None
    lib().orc_query(_p(base), C.c_uint32(base.shape[0]), _p(query), C.c_uint32(Nq), C.c_uint32(base.shape[1]),
                    C.c_int(measure), _p(graph0), C.c_uint32(graph0.shape[1]), _p(start_points),
                    C.c_uint32(start_points.size), _p(nn1_stats), C.c_uint32(KQuery), C.c_float(tau_query),
                    C.c_uint32(max_iterations), C.c_uint32(shards_per_gpu), C.c_uint32(on_gpu_shard_id),
                    _p(ids), _p(dists), _p(stats))
    return (ids, dists, stats) if with_stats else (ids, dists)


def top(g, base, layer, measure=EUCLIDEAN):
    base = _f32(base)
    nn1 = np.zeros(g.cfg.Ns[layer], np.float32)
    lib().orc_top(C.byref(g.cfg), _p(base), C.c_int(measure), C.c_uint32(layer), C.byref(g.view), _p(nn1))
    return nn1


def nn1_stats(nn1):
    nn1 = _f32(nn1)
    out = np.zeros(2, np.float32)
    lib().orc_nn1_stats(_p(nn1), C.c_uint32(nn1.size), _p(out))
    return out


def select(g, layer, nn1, rng):
    nn1, rng = _f32(nn1), _f32(rng)
    lib().orc_select(C.byref(g.cfg), C.c_uint32(layer), _p(nn1), _p(rng), C.byref(g.view))


def merge(g, base, layer_top, layer_btm, tau_build, measure=EUCLIDEAN):
    base = _f32(base)
    nn1 = np.zeros(g.cfg.N, np.float32)
    lib().orc_merge(C.byref(g.cfg), _p(base), C.c_int(measure), C.c_float(tau_build), C.c_uint32(layer_top),
                    C.c_uint32(layer_btm), C.byref(g.view), _p(nn1))
    return nn1


def sym(g, base, layer, tau_build, measure=EUCLIDEAN):
    base = _f32(base)
    KF = g.cfg.KBuild // 2
    sb = np.full((g.cfg.Ns[layer], KF), -1, np.int32)
    sa = np.zeros(g.cfg.Ns[layer], np.uint32)
    lib().orc_sym(C.byref(g.cfg), _p(base), C.c_int(measure), C.c_float(tau_build), C.c_uint32(layer),
                  C.byref(g.view), _p(sb), _p(sa))
    return sb, sa


def sym_buffer_merge(g, layer, sym_buffer, sym_atomic):
    sb, sa = _i32(sym_buffer), np.ascontiguousarray(sym_atomic, np.uint32)
    lib().orc_sym_buffer_merge(C.byref(g.cfg), C.c_uint32(layer), _p(sb), _p(sa), C.byref(g.view))


def build_graph(cfg, base, tau_build, rng, refinement_iterations=2, measure=EUCLIDEAN):
    base, rng = _f32(base), _f32(rng)
    assert rng.size >= cfg.Ns[0] + cfg.Ns[1] + cfg.Ns[2]
    g = Graph(cfg)
    lib().orc_build(C.byref(cfg), _p(base), C.c_int(measure), C.c_float(tau_build),
                    C.c_uint32(refinement_iterations), _p(rng), _p(g.blob))
    return g


def merge_results(ids, dists, K, spg_N_shard):
    """ids/dists: [n_parts, Nq, K_in] per-partition sorted lists -> [Nq, K]."""
    ids, dists = _i32(ids), _f32(dists)
    P, Nq, K_in = ids.shape
    oi = np.empty((Nq, K), np.int32)
    od = np.empty((Nq, K), np.float32)
    lib().orc_merge_results(_p(ids), _p(dists), C.c_uint32(P), C.c_uint32(Nq), C.c_uint32(K_in), C.c_uint32(K),
                            C.c_uint32(spg_N_shard), _p(oi), _p(od))
    return oi, od


def evaluate(gt, results, KQuery, base=None, query=None, measure=EUCLIDEAN):
    """-> dict(c1, c1_dup, cK, cK_dup, rK, rK_dup) (eval.cpp:176-242)."""
    gt, results = _i32(gt), _i32(results)
    out = np.zeros(6, np.float32)
    b = _f32(base) if base is not None else None
    q = _f32(query) if query is not None else None
    D = b.shape[1] if b is not None else 0
    lib().orc_eval(_p(b), C.c_uint32(b.shape[0] if b is not None else 0), _p(q), C.c_uint32(results.shape[0]),
                   C.c_uint32(D), C.c_int(measure), _p(gt), C.c_uint32(gt.shape[1]), _p(results),
                   C.c_uint32(KQuery), _p(out))
    return dict(zip(("c1", "c1_dup", "cK", "cK_dup", "rK", "rK_dup"), map(float, out)))


class Cache:
    """SimpleKNNCache push/pop emulation (simple_knn_cache.cuh:126-239) for known-answer tests."""

    def __init__(self, best, sorted_size, cache_size, vblock=32, xi=0.0):
        self.best, self.sorted, self.cache = best, sorted_size, cache_size
        self.h = C.c_void_p(lib().orc_cache_create(C.c_uint32(best), C.c_uint32(sorted_size),
                                                   C.c_uint32(cache_size), C.c_uint32(vblock)))
        lib().orc_cache_set_xi(self.h, C.c_float(xi))

    def push(self, key, dist):
        lib().orc_cache_push(self.h, C.c_int32(key), C.c_float(dist))

    def pop(self):
        return lib().orc_cache_pop(self.h)

    def state(self):
        keys = np.empty(self.cache, np.int32)
        dists = np.empty(self.sorted, np.float32)
        ph, vh = C.c_uint32(), C.c_uint32()
        lib().orc_cache_state(self.h, _p(keys), _p(dists), C.byref(ph), C.byref(vh))
        return keys, dists, ph.value, vh.value

    def __del__(self):
        try:
            lib().orc_cache_destroy(self.h)
        except Exception:
            pass
