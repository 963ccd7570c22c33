"""Host-side mirror of the reference's Python module `ggnn` (src/ggnn/python/nanobind.cu:131-301) on top
of the C ABI (include/ggnn_b200.h).  Same class / method names, arguments, defaults and return types:

    my_ggnn = GGNN(); my_ggnn.set_base(base); my_ggnn.build(24, 0.5)
    indices, dists = my_ggnn.query(query, 10, 0.64, 400)
    gt, _ = my_ggnn.bf_query(query, 10)
    Evaluator(base, query, gt, 10).evaluate_results(indices)

torch is used for device memory, streams and peer copies only; every computation on the path is a
kernel of libggnn_b200.so.  No CPU fallback: without a CUDA device the calls raise.
"""
import ctypes as C
import math
import os
from enum import IntEnum

import numpy as np
import torch

from . import _lib, swap


class DistanceMeasure(IntEnum):  # include/ggnn/base/def.h:27-30
    Euclidean = 0
    Cosine = 1


_log_level = 0


def set_log_level(level):  # nanobind.cu:151
    global _log_level
    _log_level = int(level)


def _log(level, *a):
    if _log_level >= level:
        print("[ggnn_b200]", *a, flush=True)


def _as_tensor(x, what):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    if hasattr(x, "tensor"):  # FloatDataset & co
        x = x.tensor
    if not isinstance(x, torch.Tensor):
        raise TypeError(f"{what} must be a torch tensor or numpy array")
    if x.dim() != 2:
        raise ValueError(f"{what} must be 2-dimensional [N, D]")
    return x


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Graph:
    """Read-only view of one shard's graph (nanobind.cu:295-300; layout src/ggnn/base/graph.cpp:33-92)."""

    def __init__(self, cfg, blob):
        self.config = cfg
        self.blob = blob  # uint8 tensor on the device (or CPU)
        o = _lib.graph_offsets(cfg)
        K = cfg.KBuild
        self.graph = blob[o.graph:o.graph + cfg.N_all * K * 4].view(torch.int32).view(cfg.N_all, K)
        self.translation = blob[o.translation:o.translation + cfg.ST_all * 4].view(torch.int32)
        self.selection = blob[o.selection:o.selection + cfg.ST_all * 4].view(torch.int32)
        self.nn1_stats = blob[o.nn1_stats:o.nn1_stats + 8].view(torch.float32)

    def layer_graph(self, l):
        c = self.config
        return self.graph[c.Ns_offsets[l]:c.Ns_offsets[l] + c.Ns[l]]

    def layer_translation(self, l):
        c = self.config
        return self.translation[c.STs_offsets[l]:c.STs_offsets[l] + c.Ns[l]] if l else None

    def layer_selection(self, l):
        c = self.config
        return self.selection[c.STs_offsets[l]:c.STs_offsets[l] + c.Ns[l]] if l else None


class _Shard:
    def __init__(self, device, base, global_id, pool=None, host_rows=None):
        self.device = device
        self.base = base          # [N_shard, D] fp32 on `device` (None in swap mode: see ggnn_b200/swap.py; None for a
                                  # uint8 base until a kernel needs widened rows: GGNN._f32)
        self.base_u8 = None       # [N_shard, D] uint8 on `device`: native rows of a uint8 base (resident mode)
        self.base_il = None       # optional interleaved copy of fp32 rows (GGNN_B200_INTERLEAVED_BASE=1, see _interleaved)
        self.global_id = global_id
        self.graph = None         # Graph (resident mode)
        self.pool = pool          # swap.ShardPool of this shard's GPU, or None = everything stays resident
        self.host_rows = host_rows  # swap mode: view of the host base tensor


class QueryFuture:
    """Handle of a query enqueued by GGNN.query_async()."""

    def __init__(self, events, ids, dists):
        self._events, self._ids, self._dists = events, ids, dists

    def done(self):
        return all(e.query() for e in self._events)

    def result(self):
        for e in self._events:
            e.synchronize()
        return self._ids, self._dists


class GGNN:
    """Drop-in for ggnn.GGNN (include/ggnn/base/ggnn.cuh:41-182, nanobind.cu:184-267)."""

    MIN_D, MAX_D = 1, 4096
    MIN_KBUILD, MAX_KBUILD = 2, 512

    def __init__(self):
        self._base = None
        self._gpus = [0]
        self._shard_size = 0
        self._workdir = "."
        self._results_on_gpu = False
        self._shards = []
        self._kbuild = None
        self._measure = None
        self._work_counters = {}
        self._base_dtype = torch.float32
        self._host_streams = {}
        self._host_rr = 0
        self._pools = []          # per GPU: swap.ShardPool or None (all shards resident)

    # ---- configuration (ggnn.cu:53-60, 420-454) ----
    def set_working_directory(self, path):
        self._workdir = str(path)

    def set_cpu_memory_limit(self, limit):
        # pinned host memory the swapped-out graphs may use; beyond it they go to part_<id>.ggnn files in the working
        # directory (gpu_instance.cu:177-200).  Irrelevant while all shards of a GPU fit into its memory.
        self._cpu_limit = int(limit)

    def set_reserved_gpu_memory(self, reserved):
        # GPU memory to leave free for queries / results when counting how many shards fit (gpu_instance.cu:150-175)
        self._reserved = int(reserved)

    def set_gpus(self, gpu_ids):
        if self._shards:
            raise RuntimeError("GPUs cannot be changed after the graph has been set up")  # ggnn.cu:146-152
        self._gpus = [int(g) for g in gpu_ids]
        if not self._gpus:
            raise ValueError("at least one GPU is required")

    def set_shard_size(self, n_shard):
        if self._shards:
            raise RuntimeError("shard size cannot be changed after the graph has been set up")
        self._shard_size = int(n_shard)

    def set_return_results_on_gpu(self, flag=True):
        self._results_on_gpu = bool(flag)

    def set_base(self, base):
        base = _as_tensor(base, "base")
        if base.dtype not in (torch.float32, torch.uint8):
            raise TypeError("base must be float32 or uint8 (lib.h:26-28)")
        # uint8 base vectors (the reference's BaseT = uint8_t instantiation, lib.h:26-28) stay uint8: the traversal
        # kernel reads the 1-byte rows natively (a quarter of the gather traffic and of the memory; integer dp4a
        # arithmetic, exact for D <= 256).  Every distance of the reference is computed on static_cast<float>(value)
        # (distance.cuh:131-148), so the kernels without a native variant (construction, brute force, other query shapes)
        # run on rows widened to fp32 on the device for the duration of the call: bit-identical results.
        self._base_dtype = base.dtype
        if self._shards:
            raise RuntimeError("base cannot be changed after the graph has been set up")
        if not (self.MIN_D <= base.shape[1] <= self.MAX_D):
            raise ValueError("unsupported dimension")
        self._base = base.contiguous()  # (used in place; the reference copies)

    # ---- sharding (ggnn.cu:154-203) ----
    def _prepare(self, k_build):
        if self._base is None:
            raise RuntimeError("The base needs to be set before building a graph.")
        if not torch.cuda.is_available():
            raise RuntimeError("ggnn_b200 needs a CUDA device (there is no CPU fallback)")
        if not (self.MIN_KBUILD <= k_build <= self.MAX_KBUILD):
            raise ValueError("KBuild out of range")
        if self._shards:
            if self._kbuild != k_build:
                raise RuntimeError("graph already set up with a different KBuild")
            return
        N = self._base.shape[0]
        n_shard = self._shard_size or N
        if N % n_shard:
            raise ValueError("base size must be divisible by the shard size")
        num_shards = N // n_shard
        if num_shards % len(self._gpus):
            raise ValueError("number of shards must be divisible by the number of GPUs")
        self._spg = num_shards // len(self._gpus)
        self._n_shard = n_shard
        self._kbuild = k_build
        l = _lib.lib()  # fail early if the CUDA library is missing
        cfg = self._cfg()
        blob_bytes = l.ggnn_b200_graph_blob_bytes(C.byref(cfg))
        shard_bytes = n_shard * self._base.shape[1] * 4 + blob_bytes
        scratch_bytes = l.ggnn_b200_build_scratch_bytes(C.byref(cfg))
        for gi, gpu in enumerate(self._gpus):
            dev = torch.device("cuda", gpu)
            # how many shards fit on this GPU (gpu_instance.cu:135-227); a base that already lives on the GPU stays there
            n_buf = self._spg
            if not self._base.is_cuda:
                free_bytes, _ = torch.cuda.mem_get_info(dev)
                n_buf = swap.plan_gpu_buffers(free_bytes, getattr(self, "_reserved", 0), scratch_bytes, shard_bytes, self._spg,
                                              int(os.environ.get("GGNN_B200_GPU_SHARD_BUFFERS", "0")))
            pool = None
            if n_buf < self._spg:
                n_cpu = swap.plan_cpu_buffers(getattr(self, "_cpu_limit", None), blob_bytes, self._spg)
                pool = swap.ShardPool(dev, n_buf, n_shard, self._base.shape[1], blob_bytes, n_cpu, self._workdir)
                _log(1, f"GPU {gpu}: {n_buf} of {self._spg} shards resident, {n_cpu} graphs in pinned host memory, rest on disk")
            self._pools.append(pool)
            for s in range(self._spg):
                gid = gi * self._spg + s
                rows = self._base[gid * n_shard:(gid + 1) * n_shard]
                if pool is None and rows.dtype == torch.uint8:
                    sh = _Shard(dev, None, gid)
                    sh.base_u8 = rows.to(dev, non_blocking=True).contiguous()
                    self._shards.append(sh)
                elif pool is None:
                    self._shards.append(_Shard(dev, rows.to(dev, non_blocking=True).contiguous(), gid))
                else:
                    self._shards.append(_Shard(dev, None, gid, pool, rows))

    def _f32(self, sh):
        """fp32 rows of a resident shard; the rows of a uint8 base are widened on the device on demand (exact)"""
        if sh.base is None and sh.base_u8 is not None:
            with torch.cuda.device(sh.device):
                out = torch.empty(sh.base_u8.shape, dtype=torch.float32, device=sh.device)
                _lib.check(_lib.lib().ggnn_b200_widen_u8(_ptr(sh.base_u8), _ptr(out), sh.base_u8.numel(), _stream_ptr(sh.device)))
            sh.base = out
        return sh.base

    def _interleaved(self, sh, sh_base):
        """GGNN_B200_INTERLEAVED_BASE=1: a second copy of a resident fp32 shard whose rows hold element 32c + t at 4t + c
        (ggnn_b200_interleave_rows), so that the traversal kernel reads a lane's four dims with one 16-byte load.  Same
        results; doubles the memory of the base, hence opt-in.  D = 128 only."""
        if sh.base_il is None:
            with torch.cuda.device(sh.device):
                out = torch.empty_like(sh_base)
                _lib.check(_lib.lib().ggnn_b200_interleave_rows(_ptr(sh_base), _ptr(out), sh_base.shape[0], sh_base.shape[1],
                                                                _stream_ptr(sh.device)))
            sh.base_il = out
        return sh.base_il

    @staticmethod
    def _drop_f32(sh):
        """release the widened copy of a uint8 shard again (GGNN_B200_KEEP_WIDENED=1 keeps it)"""
        if sh.base_u8 is not None and not os.environ.get("GGNN_B200_KEEP_WIDENED"):
            sh.base = None

    def _cfg(self):
        return _lib.graph_config(self._n_shard, self._base.shape[1], self._kbuild)

    # ---- build / store / load ----
    def build(self, k_build, tau_build, refinement_iterations=2, measure=DistanceMeasure.Euclidean):
        self._prepare(int(k_build))
        cfg = self._cfg()
        l = _lib.lib()
        blob_bytes = l.ggnn_b200_graph_blob_bytes(C.byref(cfg))
        scratch_bytes = l.ggnn_b200_build_scratch_bytes(C.byref(cfg))
        for i, sh in enumerate(self._shards):
            with torch.cuda.device(sh.device):
                if sh.pool is None:
                    base, blob = self._f32(sh), torch.zeros(blob_bytes, dtype=torch.uint8, device=sh.device)
                else:  # swap mode: build in a pool slot, start loading the next shard of this GPU meanwhile
                    base, blob = sh.pool.acquire(sh.global_id, sh.host_rows)
                    blob.zero_()
                    nxt = self._shards[i + 1] if i + 1 < len(self._shards) else None
                    if nxt is not None and nxt.pool is sh.pool:
                        sh.pool.prefetch(nxt.global_id, nxt.host_rows, keep=(sh.global_id,))
                scratch = torch.empty(scratch_bytes, dtype=torch.uint8, device=sh.device)
                # one cuRAND generator (XORWOW, seed 1234) per GPU whose sequence continues over the shards built on it,
                # like the reference's (graph_construction.cu:96-102,127): every shard gets the reference's selection
                rng = self._rng(sh.device)
                d_rng = torch.empty(cfg.Ns[0] + cfg.Ns[1] + cfg.Ns[2], dtype=torch.float32, device=sh.device)
                _lib.check(l.ggnn_b200_rng_fill_build(rng, C.byref(cfg), _ptr(d_rng), _stream_ptr(sh.device)))
                _lib.check(l.ggnn_b200_build_graph(C.byref(cfg), _ptr(base), int(measure), float(tau_build),
                                                   int(refinement_iterations), _ptr(d_rng), _ptr(blob), _ptr(scratch),
                                                   scratch_bytes, _stream_ptr(sh.device)))
                if sh.pool is None:
                    sh.graph = Graph(cfg, blob)
                else:
                    sh.pool.mark_built(sh.global_id)
                torch.cuda.current_stream(sh.device).synchronize()
                del scratch, base
                self._drop_f32(sh)
        self._measure = int(measure)

    def _rng(self, device):
        rngs = self.__dict__.setdefault("_rngs", {})
        if device not in rngs:
            h = C.c_void_p()
            _lib.check(_lib.lib().ggnn_b200_rng_create(C.byref(h), 1234))
            rngs[device] = h
        return rngs[device]

    def __del__(self):
        try:
            for h in self.__dict__.get("_rngs", {}).values():
                _lib.lib().ggnn_b200_rng_destroy(h)
        except Exception:  # interpreter shutdown
            pass

    def _has_graph(self):
        return bool(self._shards) and all((sh.graph is not None) if sh.pool is None else (sh.global_id in sh.pool.has_graph)
                                          for sh in self._shards)

    def store(self):
        if not self._has_graph():
            raise RuntimeError("There is no graph to store.")
        os.makedirs(self._workdir, exist_ok=True)
        for sh in self._shards:  # gpu_instance.cu:86-115: part_<global_shard_id>.ggnn = raw blob
            path = os.path.join(self._workdir, f"part_{sh.global_id}.ggnn")
            if sh.pool is None:
                sh.graph.blob.cpu().numpy().tofile(path)
            else:
                sh.pool.blob_to_file(sh.global_id, path)

    def load(self, k_build):
        self._prepare(int(k_build))
        cfg = self._cfg()
        nbytes = _lib.lib().ggnn_b200_graph_blob_bytes(C.byref(cfg))
        for sh in self._shards:
            path = os.path.join(self._workdir, f"part_{sh.global_id}.ggnn")
            if os.path.getsize(path) != nbytes:  # gpu_instance.cu:454-455 validates by size only
                raise RuntimeError(f"{path}: unexpected file size")
            if sh.pool is not None:
                sh.pool.adopt_file(sh.global_id)   # read when the shard is swapped in
                continue
            blob = torch.from_numpy(np.fromfile(path, dtype=np.uint8)).to(sh.device)
            sh.graph = Graph(cfg, blob)

    def _resident(self, sh, keep=()):
        """-> (base rows, Graph) of a shard on its device; swap mode: valid until the shard is evicted"""
        if sh.pool is None:
            return sh.base, sh.graph
        base, blob = sh.pool.acquire(sh.global_id, sh.host_rows, keep)
        return base, Graph(self._cfg(), blob)

    def get_graph(self, global_shard_id=0):
        sh = self._shards[global_shard_id]
        if sh.pool is None:
            return sh.graph
        with torch.cuda.device(sh.device):
            return self._resident(sh)[1]

    # ---- query ----
    def _counter(self, device):
        # one scheduling counter per (device, stream): batches enqueued on different streams may overlap
        key = (device, torch.cuda.current_stream(device).cuda_stream)
        if key not in self._work_counters:
            self._work_counters[key] = torch.zeros(1, dtype=torch.int32, device=device)
        return self._work_counters[key]

    def _query_device(self, gpu_index, q_dev, k_query, tau_query, max_iterations, measure, scatter=None):
        """all shards of one GPU -> sorted [Nq, K] ids (GPU-local numbering) + dists on that GPU.
        scatter (exchange.ScatterTarget): the kernels store their lists (shard-local ids) straight into the gathered
        buffers of the destination GPUs instead (fused shard-merge exchange) and nothing is returned"""
        l = _lib.lib()
        shards = self._shards[gpu_index * self._spg:(gpu_index + 1) * self._spg]
        dev = shards[0].device
        Nq = q_dev.shape[0]
        with torch.cuda.device(dev):
            ids = dists = None
            if scatter is None:
                ids = torch.empty((Nq, k_query * self._spg), dtype=torch.int32, device=dev)
                dists = torch.empty((Nq, k_query * self._spg), dtype=torch.float32, device=dev)
            q_f32 = None
            order = list(range(len(shards)))
            pool = shards[0].pool
            if pool is not None:
                # swap mode: alternate the direction from call to call, so that the shards left on the GPU by the
                # previous call are searched first (gpu_instance.cu:669-670, 740); one counter per GPU (pool)
                if pool.query_calls % 2:
                    order.reverse()
                pool.query_calls += 1
            for oi, s in enumerate(order):
                sh = shards[s]
                sh_base, sh_graph = self._resident(sh)
                if sh.pool is not None and oi + 1 < len(order):
                    nxt = shards[order[oi + 1]]
                    sh.pool.prefetch(nxt.global_id, nxt.host_rows, keep=(sh.global_id,))
                cfg = sh_graph.config
                p = _lib.QueryParams()
                p.D, p.measure, p.KQuery = cfg.D, int(measure), int(k_query)
                p.tau_query, p.max_iterations = float(tau_query), int(max_iterations)
                p.N_base, p.KBuild, p.num_starting_points = cfg.N, cfg.KBuild, cfg.S
                # uint8 base: native 1-byte rows where the kernel has a variant, else rows widened on the device
                native = (sh.pool is None and sh.base_u8 is not None and q_dev.dtype == torch.uint8 and
                          cfg.D in (32, 64, 96, 128, 256) and int(k_query) <= 47 and
                          not os.environ.get("GGNN_B200_NO_NATIVE_U8"))
                if native:
                    p.base_type = 1
                    p.d_base, p.d_query = sh.base_u8.data_ptr(), q_dev.data_ptr()
                else:
                    if sh_base is None:
                        sh_base = self._f32(sh)
                    if q_dev.dtype == torch.uint8:
                        if q_f32 is None:
                            q_f32 = q_dev.float()
                    p.d_base, p.d_query = sh_base.data_ptr(), (q_f32 if q_f32 is not None else q_dev).data_ptr()
                    if (sh.pool is None and sh.base_u8 is None and cfg.D == 128 and int(measure) == 0 and
                            os.environ.get("GGNN_B200_INTERLEAVED_BASE")):
                        p.d_base_interleaved = self._interleaved(sh, sh_base).data_ptr()
                p.d_graph = sh_graph.graph.data_ptr()
                p.d_starting_points = sh_graph.layer_translation(_lib.L - 1).data_ptr()
                p.d_nn1_stats = sh_graph.nn1_stats.data_ptr()
                if scatter is None:
                    p.d_query_results, p.d_query_results_dists = ids.data_ptr(), dists.data_ptr()
                    p.shards_per_gpu, p.on_gpu_shard_id = self._spg, s
                else:
                    scatter.apply(p, s)
                p.d_work_counter = self._counter(dev).data_ptr()
                _lib.check(l.ggnn_b200_query(C.byref(p), Nq, _stream_ptr(dev)))
            if scatter is not None:
                return None
            if self._spg > 1:  # replaces the segmented sort gpu_instance.cu:745-790
                out_i = torch.empty((Nq, k_query), dtype=torch.int32, device=dev)
                out_d = torch.empty((Nq, k_query), dtype=torch.float32, device=dev)
                _lib.check(l.ggnn_b200_merge_topk(_ptr(ids), _ptr(dists), self._spg, k_query, k_query * self._spg,
                                                  k_query, Nq, k_query, 0, _ptr(out_i), _ptr(out_d), _stream_ptr(dev)))
                ids, dists = out_i, out_d
        return ids, dists

    def query(self, query, k_query, tau_query, max_iterations=400, measure=DistanceMeasure.Euclidean):
        if not self._has_graph():
            raise RuntimeError("There is no graph to query.")
        query = _as_tensor(query, "query")
        if query.dtype != getattr(self, "_base_dtype", torch.float32):
            raise ValueError("query data type has to match base data type")  # ggnn.cu:524-540
        if query.shape[1] != self._base.shape[1]:
            raise ValueError("query dimension does not match the base")
        l = _lib.lib()
        k_query = int(k_query)
        n_gpus = len(self._gpus)
        # (the reference refuses results on the GPU for several GPUs, ggnn.cu:299-306: its merge runs on the CPU.  Here the
        # merged lists live on the first GPU, so the flag is honoured for any number of GPUs.)
        if n_gpus == 1 and not self._results_on_gpu and not query.is_cuda:
            chunks = int(os.environ.get("GGNN_B200_QUERY_CHUNKS", "0")) or (2 if query.shape[0] >= 4096 else 1)
            if self._pools and self._pools[0] is not None:
                chunks = 1  # swap mode: one stream at a time owns the shard slots (see swap.ShardPool)
            return self._enqueue_host_query(query, k_query, tau_query, max_iterations, measure, chunks).result()
        dev0 = self._shards[0].device
        gather = self._local_gather(query.shape[0], k_query) if n_gpus > 1 else None
        per_gpu = []
        for gi in range(n_gpus):
            dev = self._shards[gi * self._spg].device
            q_dev = query.to(dev, non_blocking=True).contiguous()
            per_gpu.append(self._query_device(gi, q_dev, k_query, tau_query, max_iterations, measure,
                                              scatter=gather.target(gi) if gather else None))
        if n_gpus == 1:
            ids, dists = per_gpu[0]
        elif gather is not None:
            # fused shard-merge exchange: every GPU's kernels stored their lists into the first GPU's buffer (peer stores)
            ids, dists = gather.merge(query.shape[0], self._n_shard)
        else:  # no peer access: peer copies + one merge kernel (replaces ResultMerger::merge, result_merger.cpp:51-149)
            with torch.cuda.device(dev0):
                Nq = query.shape[0]
                all_i = torch.empty((n_gpus, Nq, k_query), dtype=torch.int32, device=dev0)
                all_d = torch.empty((n_gpus, Nq, k_query), dtype=torch.float32, device=dev0)
                for gi, (i_, d_) in enumerate(per_gpu):
                    torch.cuda.current_stream(i_.device).synchronize()
                    all_i[gi].copy_(i_, non_blocking=True)
                    all_d[gi].copy_(d_, non_blocking=True)
                ids = torch.empty((Nq, k_query), dtype=torch.int32, device=dev0)
                dists = torch.empty((Nq, k_query), dtype=torch.float32, device=dev0)
                _lib.check(l.ggnn_b200_merge_topk(_ptr(all_i), _ptr(all_d), n_gpus, Nq * k_query, k_query, k_query, Nq,
                                                  k_query, self._spg * self._n_shard, _ptr(ids), _ptr(dists),
                                                  _stream_ptr(dev0)))
        if self._results_on_gpu:
            return ids, dists
        return ids.cpu(), dists.cpu()

    def _local_gather(self, n_query, k_query):
        """exchange.LocalGather for this (Nq, K), or None when the GPUs cannot access each other's memory"""
        from . import exchange
        key = (int(n_query), int(k_query))
        cache = self.__dict__.setdefault("_gathers", {})
        if key not in cache:
            if os.environ.get("GGNN_B200_NO_PEER_GATHER"):
                cache[key] = None
            else:
                try:
                    cache.clear()  # one buffer at a time
                    cache[key] = exchange.LocalGather([self._shards[g * self._spg].device for g in range(len(self._gpus))],
                                                      n_query, k_query, self._spg)
                except (NotImplementedError, _lib.GGNNError) as e:
                    _log(1, f"peer gather unavailable ({e}); using peer copies")
                    cache[key] = None
        return cache[key]

    def query_async(self, query, k_query, tau_query, max_iterations=400, measure=DistanceMeasure.Euclidean):
        """query() for a HOST query tensor on a single GPU without waiting: the host->device copy, the traversal and
        the device->host copy of the results are enqueued on one of this instance's own streams; .result() of the
        returned QueryFuture waits for them and hands out (ids, dists) in pinned host memory.  Several batches may be
        in flight at once: their copies overlap the other batches' kernels, and the SMs a batch's last long queries
        leave idle are filled by the next batch.  (The reference's query is synchronous only: ggnn.cu:506-551.)"""
        if not self._has_graph():
            raise RuntimeError("There is no graph to query.")
        query = _as_tensor(query, "query")
        if query.dtype != getattr(self, "_base_dtype", torch.float32):
            raise ValueError("query data type has to match base data type")
        if query.shape[1] != self._base.shape[1]:
            raise ValueError("query dimension does not match the base")
        if len(self._gpus) != 1 or query.is_cuda:
            raise RuntimeError("query_async takes a host tensor and a single GPU")
        return self._enqueue_host_query(query, int(k_query), tau_query, max_iterations, measure, 1)

    def _enqueue_host_query(self, query, k_query, tau_query, max_iterations, measure, n_chunks):
        dev = self._shards[0].device
        Nq = query.shape[0]
        if dev not in self._host_streams:
            self._host_streams[dev] = [torch.cuda.Stream(dev) for _ in range(4)]
        streams = self._host_streams[dev]
        out_i = torch.empty((Nq, k_query), dtype=torch.int32, pin_memory=True)
        out_d = torch.empty((Nq, k_query), dtype=torch.float32, pin_memory=True)
        n_chunks = max(1, min(int(n_chunks), len(streams), max(1, Nq)))
        bounds = [Nq * c // n_chunks for c in range(n_chunks + 1)]
        events = []
        with torch.cuda.device(dev):
            start = torch.cuda.current_stream(dev).record_event()  # after whatever the caller enqueued before
            for c in range(n_chunks):
                lo, hi = bounds[c], bounds[c + 1]
                if hi == lo:
                    continue
                st = streams[self._host_rr % len(streams)]
                self._host_rr += 1
                st.wait_event(start)
                with torch.cuda.stream(st):
                    q_dev = query[lo:hi].to(dev, non_blocking=True)
                    ids, dists = self._query_device(0, q_dev.contiguous(), k_query, tau_query, max_iterations, measure)
                    out_i[lo:hi].copy_(ids, non_blocking=True)
                    out_d[lo:hi].copy_(dists, non_blocking=True)
                    events.append(st.record_event())
        return QueryFuture(events, out_i, out_d)

    def bf_query(self, query, k_gt=100, measure=DistanceMeasure.Euclidean):
        if self._base is None:
            raise RuntimeError("The base needs to be set before running a brute-force query.")
        if len(self._gpus) > 1:
            raise RuntimeError("bf_query supports only a single GPU")  # ggnn.cu:338-339
        if not torch.cuda.is_available():
            raise RuntimeError("ggnn_b200 needs a CUDA device (there is no CPU fallback)")
        query = _as_tensor(query, "query")
        if query.dtype != getattr(self, "_base_dtype", torch.float32):
            raise ValueError("query data type has to match base data type")
        dev = torch.device("cuda", self._gpus[0])
        with torch.cuda.device(dev):
            single = self._shards[0] if (self._shards and len(self._shards) == 1 and self._shards[0].pool is None) else None
            # uint8 base: exact integer contraction on the int8 tensor cores, rows never widened (GGNN_B200_NO_I8_BF=1:
            # widen and take the fp32 path instead -- identical results)
            ws_u8 = 0
            if query.dtype == torch.uint8 and not os.environ.get("GGNN_B200_NO_I8_BF"):
                ws_u8 = _lib.lib().ggnn_b200_bf_query_u8_workspace_bytes(self._base.shape[1], int(measure), int(k_gt),
                                                                         self._base.shape[0], query.shape[0])
            if ws_u8:
                rows = single.base_u8 if (single is not None and single.base_u8 is not None) else self._base.to(dev).contiguous()
                ids, dists = self._bf_query_rows(rows, query.to(dev).contiguous(), int(k_gt), int(measure))
                return (ids, dists) if self._results_on_gpu else (ids.cpu(), dists.cpu())
            if single is not None:
                base = self._f32(single)
            else:
                base = self._base.to(dev).contiguous().float()
            ids, dists = self._bf_query_rows(base, query.to(dev).contiguous().float(), int(k_gt), int(measure))
            del base
            if single is not None:
                torch.cuda.current_stream(dev).synchronize()
                self._drop_f32(single)
        if self._results_on_gpu:
            return ids, dists
        return ids.cpu(), dists.cpu()

    @staticmethod
    def _bf_query_rows(base, q, k_gt, measure=0):
        """exact kNN of device-resident queries against device-resident base rows (any row range of a base)"""
        l = _lib.lib()
        dev = base.device
        with torch.cuda.device(dev):
            Nq = q.shape[0]
            ids = torch.empty((Nq, k_gt), dtype=torch.int32, device=dev)
            dists = torch.empty((Nq, k_gt), dtype=torch.float32, device=dev)
            p = _lib.BfQueryParams()
            p.D, p.measure, p.KQuery, p.N_base = base.shape[1], int(measure), int(k_gt), base.shape[0]
            p.d_base, p.d_query = base.data_ptr(), q.data_ptr()
            p.d_query_results, p.d_query_results_dists = ids.data_ptr(), dists.data_ptr()
            # scratch for the tensor-core contraction path (0 = shape not covered -> exact SIMT scan)
            u8 = base.dtype == torch.uint8  # native uint8 rows: ggnn_b200_bf_query_u8 (the caller checked the shape)
            if u8 and q.dtype != torch.uint8:
                raise ValueError("query data type has to match base data type")
            size_fn = l.ggnn_b200_bf_query_u8_workspace_bytes if u8 else l.ggnn_b200_bf_query_workspace_bytes
            ws_bytes = size_fn(base.shape[1], int(measure), int(k_gt), base.shape[0], Nq)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes else None
            p.d_workspace, p.workspace_bytes = (ws.data_ptr() if ws is not None else None), ws_bytes
            _lib.check((l.ggnn_b200_bf_query_u8 if u8 else l.ggnn_b200_bf_query)(C.byref(p), Nq, _stream_ptr(dev)))
            if ws is not None:
                ws.record_stream(torch.cuda.current_stream(dev))
        return ids, dists


class Evaluation:  # include/ggnn/base/eval.h:39-48, nanobind.cu:280-293
    def __init__(self, k_query, c1, c1_dup, c_k_query, c_k_query_dup, r_k_query, r_k_query_dup):
        self.k_query, self.c1, self.c1_dup = k_query, c1, c1_dup
        self.c_k_query, self.c_k_query_dup = c_k_query, c_k_query_dup
        self.r_k_query, self.r_k_query_dup = r_k_query, r_k_query_dup

    def __repr__(self):
        return (f"c@1 (=r@1): {self.c1:.6g} +duplicates: {self.c1_dup:.6g}\n"
                f"c@{self.k_query}: {self.c_k_query:.6g} +duplicates: {self.c_k_query_dup:.6g}\n"
                f"r@{self.k_query}: {self.r_k_query:.6g} +duplicates: {self.r_k_query_dup:.6g}")


class Evaluator:
    """Recall metrics of the reference (src/ggnn/base/eval.cpp:88-242): c@1, c@K (= "recall@K"), r@K and
    their duplicate-aware variants (ground-truth prefixes extended over distance ties <= 1e-6)."""

    def __init__(self, base, query, gt, k_query, measure=DistanceMeasure.Euclidean):
        self.k_query = int(k_query)
        gt = _as_tensor(gt, "gt").cpu().to(torch.int64)
        self.gt = gt
        Nq, Kgt = gt.shape
        self.end1 = self.endk = None
        if base is None or query is None:
            return
        base = _as_tensor(base, "base").cpu().float()
        query = _as_tensor(query, "query").cpu().float()
        eps = 1e-6
        # distances of every ground-truth entry to its query (eval.cpp:37-65)
        g = base[gt.reshape(-1)].view(Nq, Kgt, -1)
        q = query[:Nq].unsqueeze(1)
        if int(measure) == 0:
            d = ((g - q) ** 2).sum(-1).sqrt()
        else:
            a_norm = (g * g).sum(-1)
            prod = a_norm * a_norm  # reference quirk: b_norm is computed from a (eval.cpp:52)
            d = torch.where(prod > 0, (1.0 - (g * q).sum(-1) / prod.sqrt()).abs(), torch.ones_like(a_norm))
        # eval.cpp:135-167: extend while dist_k - dist_ref <= eps (stops at the first larger one)
        def run_len(ref_col, start):
            ok = (d[:, start:] - d[:, ref_col:ref_col + 1]) <= eps
            return torch.cumprod(ok.to(torch.int64), dim=1).sum(1)
        self.end1 = 1 + run_len(0, 1)
        if self.k_query <= Kgt:
            self.endk = self.k_query + run_len(self.k_query - 1, self.k_query)
        else:
            self.endk = torch.full((Nq,), Kgt, dtype=torch.int64)

    def evaluate_results(self, results):
        res = _as_tensor(results, "results").cpu().to(torch.int64)
        K = self.k_query
        Nq = res.shape[0]
        gt = self.gt[:Nq]
        Kgt = gt.shape[1]
        has_dup = self.end1 is not None
        end1 = self.end1[:Nq] if has_dup else torch.ones(Nq, dtype=torch.int64)
        endk = self.endk[:Nq] if has_dup else torch.full((Nq,), K, dtype=torch.int64)
        eq = res[:, :K].unsqueeze(2) == gt.unsqueeze(1)            # [Nq, K(result), Kgt]
        kg = torch.arange(Kgt).view(1, 1, Kgt)
        within = kg < endk.view(-1, 1, 1)
        eq = eq & within
        c1 = (eq[:, 0, 0]).sum().item()
        rK = (eq[:, :, 0]).sum().item() if K > 0 else 0
        rK_dup = rK
        c1_dup = (eq[:, 0, :] & (kg[0] < end1.view(-1, 1))).sum().item()
        cK = (eq & (kg < K)).sum().item()
        cK_dup = eq.sum().item()
        inv_q, inv_r = 1.0 / Nq, 1.0 / (Nq * K)
        nan = float("nan")
        return Evaluation(K, c1 * inv_q, c1_dup * inv_q if has_dup else nan, cK * inv_r,
                          cK_dup * inv_r if has_dup else nan, rK * inv_q, rK_dup * inv_q if has_dup else nan)


class _Dataset:
    """FloatDataset / UCharDataset / IntDataset of the reference (nanobind.cu:157-182): thin wrapper
    around a 2-D tensor with fvecs/bvecs/ivecs IO (src/ggnn/base/dataset.cu:118-233)."""
    dtype = torch.float32
    np_dtype = np.float32

    def __init__(self, tensor):
        self.tensor = tensor

    @classmethod
    def load(cls, file, from_=0, num=2 ** 32 - 1, pin_memory=False):
        # memory-mapped: only the requested rows are read (a subset of a billion-vector file stays cheap)
        raw = np.memmap(file, dtype=np.uint8, mode="r")
        d = int(np.frombuffer(raw[:4].tobytes(), dtype=np.int32)[0])
        esz = np.dtype(cls.np_dtype).itemsize
        rec = 4 + d * esz
        n_total = raw.size // rec
        lo, hi = min(from_, n_total), min(n_total, from_ + num)
        if num != 2 ** 32 - 1 and hi - lo != num:
            raise ValueError("Dataset contains fewer vectors than requested.")  # dataset.cu:158-161 (CHECK_EQ there)
        body = raw[lo * rec:hi * rec].reshape(hi - lo, rec)[:, 4:]
        t = torch.from_numpy(np.ascontiguousarray(body).view(cls.np_dtype).reshape(hi - lo, d).copy())
        del raw
        if pin_memory and torch.cuda.is_available():
            t = t.pin_memory()
        return cls(t)

    def store(self, file):
        t = self.tensor.cpu().numpy()
        n, d = t.shape
        out = np.empty((n, 4 + d * t.dtype.itemsize), dtype=np.uint8)
        out[:, :4] = np.frombuffer(np.int32(d).tobytes(), dtype=np.uint8)
        out[:, 4:] = t.view(np.uint8).reshape(n, -1)
        out.tofile(file)

    @property
    def N(self):
        return self.tensor.shape[0]

    @property
    def D(self):
        return self.tensor.shape[1]

    def numel(self):
        return self.tensor.numel()

    def clone(self):
        return type(self)(self.tensor.clone())

    @property
    def view(self):
        return self.tensor

    @property
    def device(self):
        return self.tensor.device


class FloatDataset(_Dataset):
    dtype, np_dtype = torch.float32, np.float32


class UCharDataset(_Dataset):
    dtype, np_dtype = torch.uint8, np.uint8


class IntDataset(_Dataset):
    dtype, np_dtype = torch.int32, np.int32
