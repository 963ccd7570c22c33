"""Host-side logic of the Python mirror of the reference API (no GPU): argument checks, sharding rules,
Evaluator, dataset IO, result-merge semantics."""
import os

import numpy as np
import pytest
import torch

import ggnn_b200 as ggnn
from ggnn_b200 import distributed as D
from oracle import pyoracle as O


def test_api_surface_matches_reference_module():
    # nanobind.cu:131-301
    for name in ("GGNN", "DistanceMeasure", "Evaluator", "Evaluation", "FloatDataset", "UCharDataset", "IntDataset",
                 "set_log_level"):
        assert hasattr(ggnn, name)
    g = ggnn.GGNN()
    for m in ("set_base", "set_working_directory", "set_cpu_memory_limit", "set_reserved_gpu_memory", "set_gpus",
              "set_shard_size", "set_return_results_on_gpu", "build", "load", "store", "query", "bf_query", "get_graph"):
        assert callable(getattr(g, m))
    assert int(ggnn.DistanceMeasure.Euclidean) == 0 and int(ggnn.DistanceMeasure.Cosine) == 1


def test_misuse_raises_like_the_reference():
    g = ggnn.GGNN()
    with pytest.raises(RuntimeError):
        g.build(24, 0.5)                      # base not set (ggnn.cu:209)
    with pytest.raises(RuntimeError):
        g.query(torch.zeros(4, 8), 10, 0.5)   # no graph (ggnn.cu:282)
    with pytest.raises(RuntimeError):
        g.store()
    with pytest.raises(ValueError):
        g.set_base(torch.zeros(10, 5000))     # D > 4096
    with pytest.raises(TypeError):
        g.set_base(torch.zeros(10, 8, dtype=torch.int32))
    g.set_base(torch.zeros(10, 8, dtype=torch.uint8))   # uint8 base vectors are accepted (widened to fp32)
    g.set_base(torch.zeros(100, 8))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            g.build(24, 0.5)


def test_shard_layout_rules():
    assert D.shard_layout(100, 25, 2) == (4, 2)       # ggnn_main_multi_gpu.cpp:62-65 shape
    assert D.shard_layout(100_000_000, 12_500_000, 8) == (8, 1)
    assert D.local_rows(3, 100_000_000, 12_500_000, 8) == (37_500_000, 50_000_000)
    assert D.local_rows(1, 100, 25, 2) == (50, 100)
    with pytest.raises(ValueError):
        D.shard_layout(100, 30, 2)   # N % N_shard != 0
    with pytest.raises(ValueError):
        D.shard_layout(100, 20, 2)   # 5 shards on 2 GPUs


def test_evaluator_matches_oracle_restatement():
    rng = np.random.default_rng(3)
    base = rng.random((500, 16), dtype=np.float32)
    base[10] = base[11]                       # exact duplicates -> duplicate-aware metrics differ
    query = rng.random((40, 16), dtype=np.float32)
    query[0] = base[10]
    gt, _ = O.bf_query(base, query, 20)
    res = gt[:, :10].copy()
    res[::3, 2] = 499
    res[0, 0] = gt[0, 1]
    for measure in (0, 1):
        e = ggnn.Evaluator(base, query, gt, 10, measure).evaluate_results(res)
        o = O.evaluate(gt, res, 10, base, query, measure)
        got = dict(c1=e.c1, c1_dup=e.c1_dup, cK=e.c_k_query, cK_dup=e.c_k_query_dup, rK=e.r_k_query, rK_dup=e.r_k_query_dup)
        for k in o:
            assert got[k] == pytest.approx(o[k], abs=1e-6), (measure, k)
    e = ggnn.Evaluator(None, None, gt, 10).evaluate_results(res)
    assert np.isnan(e.c1_dup) and e.c_k_query == pytest.approx(O.evaluate(gt, res, 10)["cK"], abs=1e-6)
    assert "c@10" in repr(e)


def test_fvecs_roundtrip(tmp_path):
    x = torch.rand(7, 5)
    p = os.path.join(tmp_path, "x.fvecs")
    ggnn.FloatDataset(x).store(p)
    raw = np.fromfile(p, dtype=np.int32)
    assert raw[0] == 5 and raw.size == 7 * 6    # dataset.cu:118-233: [int32 D][D values] per row
    y = ggnn.FloatDataset.load(p)
    assert torch.equal(x, y.tensor) and (y.N, y.D) == (7, 5)
    z = ggnn.FloatDataset.load(p, 2, 3)
    assert torch.equal(x[2:5], z.tensor)
    i = torch.randint(0, 1000, (4, 3), dtype=torch.int32)
    p2 = os.path.join(tmp_path, "g.ivecs")
    ggnn.IntDataset(i).store(p2)
    assert torch.equal(ggnn.IntDataset.load(p2).tensor, i)


def _write_vecs(path, arr):
    """[int32 D][D values] per row (src/ggnn/base/dataset.cu:204-215)"""
    arr = np.ascontiguousarray(arr)
    with open(path, "wb") as f:
        for row in arr:
            f.write(np.int32(arr.shape[1]).tobytes())
            f.write(row.tobytes())


def test_cpp_host_api_datasets_io_and_evaluator(tmp_path):
    """include/ggnn/ggnn.hpp on the CPU: dataset factories / type erasure / [fbi]vecs IO / argument checks inside the
    C++ test program; its Evaluator output must equal the oracle's restatement of eval.cpp for float and uint8 data"""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(tmp_path, "host_api_test")
    subprocess.check_call(["g++", "-std=c++20", "-O1", f"-I{root}/include", "-I/usr/local/cuda/include",
                           f"{root}/tests/cpp/host_api_test.cpp", f"-L{root}/ggnn_b200", "-lggnn_b200",
                           "-L/usr/local/cuda/lib64", "-lcudart", f"-Wl,-rpath,{root}/ggnn_b200", "-o", exe])
    rng = np.random.default_rng(7)
    base_u8 = rng.integers(0, 256, (300, 16), dtype=np.uint8)
    base_u8[10] = base_u8[11]
    base_u8[12] = base_u8[11]                  # a run of exact duplicates
    query_u8 = rng.integers(0, 256, (25, 16), dtype=np.uint8)
    query_u8[0] = base_u8[10]
    base = (base_u8.astype(np.float32) / 255).astype(np.float32)
    query = (query_u8.astype(np.float32) / 255).astype(np.float32)
    K = 10
    gt, _ = O.bf_query(base, query, 20)
    res = gt[:, :K].copy()
    res[::3, 2] = 299
    res[0, 0] = gt[0, 1]
    res[1, 5] = res[1, 4]                      # the same id twice in one result row: every match counts (eval.cpp:206-227)
    for name, arr in (("base.fvecs", base), ("query.fvecs", query), ("base.bvecs", base_u8), ("query.bvecs", query_u8),
                      ("gt.ivecs", gt.astype(np.int32)), ("res.ivecs", res.astype(np.int32))):
        _write_vecs(os.path.join(tmp_path, name), arr)
    out = json.loads(subprocess.check_output([exe, str(tmp_path), str(K)]))
    assert out["base"] == [300, 16] and out["gt"] == [25, 20]
    for tag, b, q in (("float", base, query), ("uint8", base_u8.astype(np.float32), query_u8.astype(np.float32))):
        for measure in (0, 1):
            o = O.evaluate(gt, res, K, b, q, measure)
            got = out[f"{tag}_{measure}"]
            for k in o:
                assert got[k] == pytest.approx(o[k], abs=1e-6), (tag, measure, k)
    o = O.evaluate(gt, res, K)
    assert out["nodup"]["cK"] == pytest.approx(o["cK"], abs=1e-6) and out["nodup"]["c1_dup"] == -1
    # files written by the C++ side are readable by the Python mirror
    assert torch.equal(ggnn.FloatDataset.load(os.path.join(tmp_path, "rt.fvecs")).tensor, torch.arange(12.).view(3, 4))


def test_oracle_merge_semantics():
    ids = np.array([[[0, 1, 2]], [[0, 1, 2]]], np.int32)               # 2 partitions, 1 query, K_in 3
    d = np.array([[[0.1, 0.4, 0.9]], [[0.2, 0.4, 0.5]]], np.float32)
    oi, od = O.merge_results(ids, d, 3, 1000)
    assert od.tolist() == [[np.float32(0.1), np.float32(0.2), np.float32(0.4)]]
    assert oi.tolist() == [[0, 1000, 1]]  # tie 0.4: lower partition first; ids re-based (result_merger.cpp:115-116)


# ------------------------------------------------------------------------------------------------
# The device keeps the best list + prioQ ring in registers (WarpLists) or shared memory (SmemLists) and updates ALL
# slots of a push at once (ggnn_b200/csrc/common.cuh); the reference walks the list block by block from the top
# (simple_knn_cache.cuh:160-211).  The model below is the device's update rule written in numpy; it must leave
# exactly the state of the oracle's lock-step emulation of the reference -- for every block width, ring position
# (wrapped or not: the ring-wrap quirk of SURVEY appendix A included) and list size, also the ones beyond one warp.
# ------------------------------------------------------------------------------------------------
class _SlotParallelLists:
    EMPTY = -1

    def __init__(self, best, sorted_size):
        self.best, self.sorted = best, sorted_size
        self.key = np.full(sorted_size, self.EMPTY, np.int32)
        self.dist = np.full(sorted_size, np.inf, np.float32)
        self.head = best

    def push(self, k, d):
        d = np.float32(d)
        if (self.key == k).any():
            return
        p = np.arange(self.sorted)
        pk = np.roll(self.key, 1)
        pd = np.roll(self.dist, 1)
        recv = (p >= 1) & (p != self.best) & (p != self.head) & (pd >= d) & (pk != self.EMPTY)
        asd = np.where(recv, pd, self.dist)
        nk = np.where(recv, pk, self.key)
        pa = np.roll(asd, 1)
        pa[self.best] = asd[self.sorted - 1]
        active = self.dist >= d
        has_prev = (p != 0) & (p != self.head)
        ins = active & (~has_prev | (pa < d))
        self.key = np.where(ins, np.int32(k), nk).astype(np.int32)
        self.dist = np.where(ins, d, asd).astype(np.float32)

    def pop(self, criteria):
        k, dd = int(self.key[self.head]), self.dist[self.head]
        if k == self.EMPTY or dd >= criteria:
            return self.EMPTY
        self.key[self.head], self.dist[self.head] = self.EMPTY, np.inf
        self.head = self.best if self.head + 1 >= self.sorted else self.head + 1
        return k


@pytest.mark.parametrize("best,sorted_size,cache,vblock", [
    (10, 32, 512, 32), (10, 64, 256, 32), (25, 64, 256, 32), (100, 128, 512, 32), (10, 32, 1024, 64),
    (150, 192, 512, 32), (239, 256, 512, 32), (300, 320, 512, 32), (1000, 1024, 2048, 128), (40, 64, 512, 256)])
def test_slot_parallel_list_update_equals_reference_block_loop(best, sorted_size, cache, vblock):
    rng = np.random.default_rng(best * 7 + sorted_size)
    ref = O.Cache(best, sorted_size, cache, vblock, xi=0.25)
    dev = _SlotParallelLists(best, sorted_size)
    n_ops = 6 * sorted_size
    next_key = 0
    for step in range(n_ops):
        if rng.random() < 0.3 and step > sorted_size // 2:
            # pop: criteria() = dist[BEST-1] + xi (simple_knn_cache.cuh:121-124, 223)
            _, dists, _, _ = ref.state()
            want = ref.pop()
            got = dev.pop(np.float32(dists[best - 1] + np.float32(0.25)))
            assert got == want, step
        else:
            if rng.random() < 0.1 and next_key > 0:
                key = int(rng.integers(0, next_key))       # a key that may already be cached (dedupe, :132-146)
            else:
                key, next_key = next_key, next_key + 1
            # distances from a small grid -> many exact ties (the shift / insert rules differ on >= vs <)
            d = np.float32(rng.integers(0, 4 * sorted_size) / 8.0)
            ref.push(key, d)
            dev.push(key, d)
        keys, dists, ph, _ = ref.state()
        assert ph == dev.head, step
        assert np.array_equal(keys[:sorted_size], dev.key), step
        assert np.array_equal(dists, dev.dist), step


# ------------------------------------------------------------------------------------------------
# shard swapping GPU <-> pinned RAM <-> disk (ggnn_b200/swap.py; reference: gpu_instance.cu:135-227, 370-467)
# ------------------------------------------------------------------------------------------------
def test_swap_planning_rules():
    from ggnn_b200 import swap
    GB = 1 << 30
    assert swap.plan_gpu_buffers(180 * GB, 1 * GB, 2 * GB, 8 * GB, 8) == 8          # everything fits: resident
    assert swap.plan_gpu_buffers(40 * GB, 1 * GB, 2 * GB, 8 * GB, 8) == 4           # (40-1-2)/8
    assert swap.plan_gpu_buffers(40 * GB, 30 * GB, 2 * GB, 8 * GB, 8) == 1
    with pytest.raises(RuntimeError):
        swap.plan_gpu_buffers(8 * GB, 1 * GB, 2 * GB, 8 * GB, 8)                     # not even one shard
    assert swap.plan_gpu_buffers(1, 0, 0, 8 * GB, 8, override=3) == 3
    assert swap.plan_cpu_buffers(None, GB, 5) == 5
    assert swap.plan_cpu_buffers(int(2.5 * GB), GB, 5) == 2
    assert swap.plan_cpu_buffers(0, GB, 5) == 0


def test_shard_pool_lru_writeback_and_disk(tmp_path):
    """the pool on the CPU device (same code path minus streams): least-recently-used replacement, dirty graphs are
    written back to host memory first and to part files beyond the CPU budget, files can be adopted and stored"""
    from ggnn_b200 import swap
    rows, dim, blob_bytes = 4, 3, 64
    host = torch.arange(5 * rows * dim, dtype=torch.float32).view(5 * rows, dim)
    pool = swap.ShardPool("cpu", 2, rows, dim, blob_bytes, n_cpu=1, workdir=str(tmp_path))

    def rows_of(g):
        return host[g * rows:(g + 1) * rows]

    # "build" shards 0..4 through two slots
    for g in range(5):
        base, blob = pool.acquire(g, rows_of(g))
        assert torch.equal(base, rows_of(g)) and int(blob.sum()) == 0      # base rows loaded, empty graph
        blob.fill_(g + 1)
        pool.mark_built(g)
    assert list(pool.resident) == [3, 4] and pool.stats["evictions"] == 3
    assert set(pool.host_blob) == {0} and pool.on_disk == {1, 2}           # one graph fits the CPU budget, the rest on disk
    assert os.path.getsize(os.path.join(tmp_path, "part_1.ggnn")) == blob_bytes
    # reading them back in reverse order (the query loop alternates its direction): 4 and 3 are still resident
    for g in (4, 3, 2, 1, 0):
        base, blob = pool.acquire(g, rows_of(g), keep=())
        assert torch.equal(base, rows_of(g)) and bool((blob == g + 1).all()), g
    assert pool.stats["disk_reads"] == 2 and pool.stats["loads"] == 5 + 3
    assert pool.on_disk == {1, 2, 3, 4}                                    # 3 and 4 were dirty when evicted
    # prefetch never evicts the shard being worked on
    pool.acquire(0, rows_of(0))
    pool.prefetch(2, rows_of(2), keep=(0,))
    assert 0 in pool.resident and 2 in pool.resident
    # store: newest copy wins; adopt: an existing file replaces whatever is cached
    out = os.path.join(tmp_path, "out.ggnn")
    pool.blob_to_file(0, out)
    assert np.fromfile(out, dtype=np.uint8).tolist() == [1] * blob_bytes
    np.full(blob_bytes, 9, np.uint8).tofile(os.path.join(tmp_path, "part_0.ggnn"))
    pool.adopt_file(0)
    _, blob = pool.acquire(0, rows_of(0))
    assert bool((blob == 9).all())
    with pytest.raises(RuntimeError):
        swap.ShardPool("cpu", 1, rows, dim, blob_bytes, 0, str(tmp_path)).blob_to_file(7, out)


def test_benchmark_cli_mirrors_the_reference_flags():
    """tools/ggnn_benchmark.py: flag names / defaults of examples/cpp-and-cuda/ggnn_benchmark.cpp:37-50 and its tau lists"""
    from tools import ggnn_benchmark as B
    a = B.parse_args(["--base", "b.fvecs", "--query", "q.fvecs"])
    assert (a.k_build, a.tau_build, a.refinement_iterations, a.k_query, a.max_iterations) == (24, 0.5, 2, 10, 200)
    assert (a.measure, a.shard_size, a.gpus, a.grid_search, a.subset, a.gt, a.graph_dir) == ("euclidean", 0, [0], False, 0, "", "")
    a = B.parse_args(["--base", "b.bvecs", "--query", "q.bvecs", "--gpu_ids", "0 1 3", "--grid_search", "--measure", "cosine"])
    assert a.gpus == [0, 1, 3] and a.grid_search and a.measure == "cosine"
    assert B.tau_values(False) == [0.34, 0.41, 0.51, 0.64]
    g = B.tau_values(True)
    assert len(g) == 84 and g[0] == 0.0 and abs(g[69] - 0.69) < 1e-9 and abs(g[70] - 0.7) < 1e-9 and abs(g[-1] - 2.0) < 1e-9
    assert B.dataset_class(ggnn, "x.bvecs") is ggnn.UCharDataset and B.dataset_class(ggnn, "x.ivecs") is ggnn.IntDataset
    with pytest.raises(SystemExit):
        B.dataset_class(ggnn, "x.txt")


def test_import_ggnn_compat_package_and_reference_python_examples_start():
    """compat/python/ggnn re-exports the API under the reference's module name; the reference's own example scripts
    (only present in the build container) run unchanged up to the first call that needs a GPU"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=os.path.join(root, "compat", "python") + os.pathsep + root)
    code = ("import ggnn; assert ggnn.GGNN and ggnn.DistanceMeasure.Cosine == 1 and ggnn.Evaluator and ggnn.FloatDataset "
            "and ggnn.UCharDataset and ggnn.IntDataset and callable(ggnn.set_log_level); print(ggnn.__version__)")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert out.returncode == 0 and "b200" in out.stdout, out.stderr
    ex = "/root/reference/examples/python"
    if not os.path.isdir(ex) or torch.cuda.is_available():
        return
    for script in ("ggnn_pytorch.py", "ggnn_pytorch_multi_gpu.py"):
        p = subprocess.run([sys.executable, os.path.join(ex, script)], env=env, capture_output=True, text=True, timeout=300)
        # no GPU here: the scripts get as far as build(), which refuses to run without a CUDA device
        assert p.returncode != 0 and "no CPU fallback" in p.stderr, p.stderr[-500:]


def test_dataset_load_ranges_and_errors(tmp_path):
    """subset loading reads only the requested rows; asking for more rows than the file holds is an error like in the
    reference (src/ggnn/base/dataset.cu:158-161)"""
    x = torch.arange(40, dtype=torch.float32).view(10, 4)
    p = os.path.join(tmp_path, "x.fvecs")
    ggnn.FloatDataset(x).store(p)
    assert torch.equal(ggnn.FloatDataset.load(p, 3, 4).tensor, x[3:7])
    assert torch.equal(ggnn.FloatDataset.load(p, 9, 1).tensor, x[9:])
    assert torch.equal(ggnn.FloatDataset.load(p, 6).tensor, x[6:])          # num = all remaining rows
    with pytest.raises(ValueError):
        ggnn.FloatDataset.load(p, 8, 5)
    u = torch.randint(0, 256, (6, 16), dtype=torch.uint8)
    pb = os.path.join(tmp_path, "u.bvecs")
    ggnn.UCharDataset(u).store(pb)
    assert os.path.getsize(pb) == 6 * (4 + 16)
    assert torch.equal(ggnn.UCharDataset.load(pb, 2, 3).tensor, u[2:5])


def test_bench_query_blocks_are_prefix_stable():
    """bench.gen_gpu draws queries in blocks of 10 000: the first blocks of a larger request equal a smaller request (the
    batches of a bench step are different queries, batch 0 is the batch recall is quoted on)"""
    import bench
    dev = torch.device("cpu")
    for kind in ("manifold8", "manifoldcos8", "uniform"):
        _, q1 = bench.gen_gpu(64, 10_000, 16, kind, 1234, dev)
        _, q3 = bench.gen_gpu(64, 25_000, 16, kind, 1234, dev)
        assert q3.shape == (25_000, 16) and torch.equal(q3[:10_000], q1)
        assert not torch.equal(q3[10_000:20_000], q1)
    b1, _ = bench.gen_gpu(3000, 10, 16, "manifold8", 1234, dev, shard_index=2)
    out = torch.empty((3000, 16))
    b2, _ = bench.gen_gpu(3000, 10, 16, "manifold8", 1234, dev, shard_index=2, out=out)
    assert b2.data_ptr() == out.data_ptr() and torch.equal(b1, b2)     # generated in place into a slice of a larger base
    c1, c2 = bench.config1_data(), bench.config1_data()
    assert torch.equal(c1[0], c2[0]) and torch.equal(c1[1], c2[1]) and c1[0].shape == (10_000, 128)


def test_scatter_target_fills_the_exchange_fields():
    """exchange.ScatterTarget -> ggnn_b200_query_params: no local result buffer, shard-local ids, slot = first slot of the
    rank + shard index on the GPU"""
    from ggnn_b200 import _lib
    from ggnn_b200.exchange import ScatterTarget, _align
    t = ScatterTarget(8, 6, 10_000, 3_200_000, 0x1000, 0x2000, 0x3000)
    p = _lib.QueryParams()
    p.d_query_results, p.shards_per_gpu, p.on_gpu_shard_id = 0x99, 4, 3
    t.apply(p, 1)
    assert p.d_query_results is None and p.d_query_results_dists is None
    assert (p.shards_per_gpu, p.on_gpu_shard_id, p.n_scatter, p.scatter_slot, p.scatter_rows) == (1, 0, 8, 7, 10_000)
    assert (p.scatter_dists_offset, p.d_scatter_dst, p.d_scatter_flags, p.d_scatter_done) == (3_200_000, 0x1000, 0x2000, 0x3000)
    assert _align(1, 256) == 256 and _align(512, 256) == 512
