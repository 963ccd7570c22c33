"""The CPU oracle is pinned against (a) the reference's own GraphConfig known answers (SURVEY.md 8c,
produced by compiling src/ggnn/base/graph_config.cpp) and (b) outputs of the UNMODIFIED reference library
run on a B200 (tests/golden/*.npz, see tests/golden/README.md)."""
import numpy as np
import pytest

from oracle import pyoracle as O

# N, K, KF, S, G, S0, S0_off, SG, SG_off, Bs, Ns, N_all, ST_all, blob bytes  
KAT = [
    (10_000, 24, 12, 32, 7, 29, 53, 4, 4, [343, 49, 7, 1], [10000, 1568, 224, 32], 11_824, 1_824, 1_149_704),
    (25_000, 24, 12, 32, 9, 34, 214, 3, 5, [729, 81, 9, 1], [25000, 2592, 288, 32], 27_912, 2_912, 2_702_856),
    (100_000, 24, 12, 32, 15, 29, 2125, 2, 2, [3375, 225, 15, 1], [100000, 7200, 480, 32], 107_712, 7_712, 10_402_056),
    (1_000_000, 24, 12, 32, 32, 30, 16960, 1, 0, [32768, 1024, 32, 1], [1000000, 32768, 1024, 32], 1_033_824, 33_824, 99_517_704),
    (10_000_000, 24, 12, 32, 68, 31, 252608, 0, 32, [314432, 4624, 68, 1], [10000000, 147968, 2176, 32], 10_150_176, 150_176, 975_618_312),
    (12_500_000, 24, 12, 32, 73, 32, 51456, 0, 32, [389017, 5329, 73, 1], [12500000, 170528, 2336, 32], 12_672_896, 172_896, 1_217_981_192),
    (100_000_000, 24, 12, 32, 146, 32, 411648, 0, 32, [3112136, 21316, 146, 1], [100000000, 682112, 4672, 32], 100_686_816, 686_816, 9_671_428_872),
    (1_000_000, 40, 20, 32, 31, 33, 16897, 1, 1, [29791, 961, 31, 1], [1000000, 30752, 992, 32], 1_031_776, 31_776, 165_338_376),
    (1_000_000, 96, 48, 64, 25, 64, 0, 2, 14, [15625, 625, 25, 1], [1000000, 40000, 1600, 64], 1_041_664, 41_664, 400_332_296),
]


@pytest.mark.parametrize("row", KAT, ids=[f"N{r[0]}_K{r[1]}" for r in KAT])
def test_graph_config_known_answers(row):
    N, K, KF, S, G, S0, S0_off, SG, SG_off, Bs, Ns, N_all, ST_all, blob = row
    c = O.graph_config(N, 128, K)
    assert (c.KF, c.S, c.G, c.S0, c.S0_off, c.SG, c.SG_off) == (KF, S, G, S0, S0_off, SG, SG_off)
    assert list(c.Bs) == Bs and list(c.Ns) == Ns
    assert (c.N_all, c.ST_all) == (N_all, ST_all)
    assert O.blob_bytes(c) == blob


def test_launch_parameter_table():
    # SURVEY.md appendix B (query_kernels.cu:77-110 evaluated by hand)
    assert O.query_launch_params(128, 10, 400) == (512, 32, 32)
    assert O.query_launch_params(96, 10, 400) == (512, 32, 32)
    assert O.query_launch_params(128, 10, 200) == (256, 64, 32)
    assert O.query_launch_params(128, 100, 400) == (512, 128, 32)
    assert O.query_launch_params(128, 10, 1000) == (1024, 32, 64)
    assert O.query_launch_params(256, 10, 400) == (512, 32, 64)
    assert O.bf_block_dim(128) == 32 and O.bf_block_dim(960) == 256
    assert O.construction_config(128, 128) == (128, 4) and O.construction_config(2048, 64) == (256, 8)


def test_push_ring_wrap_quirk_trace():
    """with a wrapped ring, push loses the entry at physical slot SORTED-1 and
    duplicates the one at physical BEST (simple_knn_cache.cuh:160-211)."""
    c = O.Cache(10, 32, 512, xi=1e9)
    for i in range(10):
        c.push(100 + i, float(i))
    for _ in range(10):
        assert c.pop() >= 0
    for i in range(22):
        c.push(200 + i, 20.0 + i)
    keys, d, head, _ = c.state()
    assert head == 20
    c.push(999, 25.5)
    keys, d, head, _ = c.state()
    order = [(int(keys[10 + (head - 10 + j) % 22]), float(d[10 + (head - 10 + j) % 22])) for j in range(22)]
    expect = [(200, 20.0), (201, 21.0), (202, 22.0), (203, 23.0), (204, 24.0), (205, 25.0), (999, 25.5), (206, 26.0),
              (207, 27.0), (208, 28.0), (209, 29.0), (210, 30.0), (212, 32.0), (212, 32.0), (213, 33.0), (214, 34.0),
              (215, 35.0), (216, 36.0), (217, 37.0), (218, 38.0), (219, 39.0), (220, 40.0)]
    assert order == expect


def test_push_clean_when_ring_not_wrapped():
    c = O.Cache(4, 32, 256)
    for k, dd in [(1, 5.0), (2, 3.0), (3, 4.0), (2, 1.0), (4, 3.0)]:
        c.push(k, dd)
    keys, d, head, _ = c.state()
    assert list(keys[:4]) == [4, 2, 3, 1] and list(d[:4]) == [3.0, 3.0, 4.0, 5.0]  # new goes before equal, dup key dropped
    assert list(keys[4:8]) == [4, 2, 3, 1]


@pytest.mark.parametrize("name", ["l2_10k", "cos_10k"])
def test_oracle_query_matches_reference_dump(golden, name):
    g = golden[name]
    cfg = O.graph_config(g["N"], g["D"], g["kbuild"])
    gr = O.Graph(cfg, g["blob"])
    n = 400
    ids, dists = O.query(g["base"], g["query"][:n], gr.layer_graph(0), gr.start_points(), gr.nn1_stats, g["kquery"],
                         g["tau_query"], g["max_it"], g["measure"])
    assert np.array_equal(ids, g["query_ids"][:n])
    assert np.array_equal(dists, g["query_dists"][:n])


@pytest.mark.parametrize("name", ["l2_10k", "cos_10k"])
def test_oracle_bf_matches_reference_dump(golden, name):
    g = golden[name]
    n = 48
    ids, dists = O.bf_query(g["base"], g["query"][:n], g["kquery"], g["measure"])
    assert np.array_equal(ids, g["bf_ids"][:n])
    assert np.array_equal(dists, g["bf_dists"][:n])


@pytest.mark.parametrize("name", ["l2_10k", "cos_10k"])
def test_oracle_top_and_select_reproduce_reference_graph(golden, name):
    """Deterministic construction stages: the top-layer segment kNN (`top`) of layer 3 depends only on
    translation[3]; the reference's sym pass then only rewrites the foreign half.  The local half
    (first KL columns) of graph[3] must match the oracle's `top` bit for bit."""
    g = golden[name]
    cfg = O.graph_config(g["N"], g["D"], g["kbuild"])
    ref = O.Graph(cfg, g["blob"])
    mine = O.Graph(cfg, g["blob"].copy())
    mine.layer_graph(3)[:] = -7
    O.top(mine, g["base"], 3, g["measure"])
    KL = cfg.KBuild - cfg.KBuild // 2
    assert np.array_equal(mine.layer_graph(3)[:, :KL], ref.layer_graph(3)[:, :KL])


def test_oracle_eval_matches_definition():
    rng = np.random.default_rng(0)
    gt = np.stack([rng.permutation(1000)[:20] for _ in range(50)]).astype(np.int32)
    res = gt[:, :10].copy()
    res[:, 5:] = 5000 + np.arange(5)  # 5 of 10 correct
    res[::2, 0] = 7777                # top-1 wrong for every second query
    e = O.evaluate(gt, res, 10)
    assert e["cK"] == pytest.approx((50 * 5 - 25) / 500)
    assert e["c1"] == pytest.approx(0.5) and e["rK"] == pytest.approx(0.5)


@pytest.mark.parametrize("D,hi", [(128, 256), (96, 256), (64, 4), (32, 2)])
def test_uint8_distances_are_exact_integers_in_the_reference_arithmetic(D, hi):
    """The premise of the uint8 kernels (csrc/query.cu fetch_u8, csrc/bf_i8.cu): the reference computes on
    static_cast<float>(value) (include/ggnn/cuda_utils/distance.cuh:104-139) and for D <= 128 every partial sum is an
    integer below 2^24, so its fp32 result IS the integer |b|^2 - 2 q.b + |q|^2 -- and the K best by (distance, index)
    (k_best_list.cuh:92,100) are what an integer brute force returns.  Checked on the oracle, ties by the hundred included."""
    rng = np.random.default_rng(D + hi)
    base = rng.integers(0, hi, (3000, D), dtype=np.uint8)
    query = rng.integers(0, hi, (40, D), dtype=np.uint8)
    base[5] = base[3]
    base[100] = 255 if hi == 256 else hi - 1   # the largest norm
    query[0] = base[3]
    K = 32
    o_ids, o_d = O.bf_query(base.astype(np.float32), query.astype(np.float32), K, 0)
    b, q = base.astype(np.int64), query.astype(np.int64)
    d = (b * b).sum(1)[None, :] - 2 * (q @ b.T) + (q * q).sum(1)[:, None]
    assert d.max() < 2 ** 24
    key = d * (1 << 32) + np.arange(base.shape[0])[None, :]
    top = np.sort(key, axis=1)[:, :K]
    assert np.array_equal(o_ids, (top & 0xffffffff).astype(np.int32))
    assert np.array_equal(o_d, (top >> 32).astype(np.float32))


def test_bf_i8_check_layout_model_matches_a_byte_by_byte_loop():
    """tools/bf_i8_check.py checks the operand packing of csrc/bf_i8.cu against pack_model(); pack_model() itself is
    checked here against the definition: byte b of 16-byte chunk c of row r of tile t at t * 16 KB + r * 128 + ((c ^ (r & 7)) << 4) + b
    (the SWIZZLE_128B image of a K-major UMMA operand), rows past the end and bytes past D zero."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        "bf_i8_check", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "bf_i8_check.py"))
    T = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(T)
    rng = np.random.default_rng(0)
    for n, n_pad, D in ((300, 384, 96), (128, 128, 128), (5, 128, 32)):
        rows = rng.integers(0, 256, (n, D), dtype=np.uint8)
        want = np.zeros(n_pad * 128, dtype=np.uint8)
        for row in range(n):
            t, r = row >> 7, row & 127
            for c in range(D // 16):
                off = t * 16384 + r * 128 + ((c ^ (r & 7)) << 4)
                want[off:off + 16] = rows[row, c * 16:(c + 1) * 16]
        assert np.array_equal(T.pack_model(rows, n_pad), want)


# ------------------------------------------------------------------------------------------------
# host-side result handling: oracle restatements and the product's Evaluator against outputs of the UNMODIFIED reference
# (tests/golden/host_merge_eval.npz, produced by tools/gen_host_golden.py + oracle/ref_host_check.cpp; no GPU involved)
# ------------------------------------------------------------------------------------------------
def _host_golden():
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_host_golden", os.path.join(root, "tools", "gen_host_golden.py"))
    G = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(G)
    return G, np.load(os.path.join(root, "tests", "golden", "host_merge_eval.npz"))


def test_oracle_result_merge_matches_reference_result_merger():
    """orc_merge_results == ggnn::ResultMerger::merge (src/ggnn/base/result_merger.cpp:51-149) on per-GPU sorted lists with
    distinct distances: ids (with the reference's partition * spg * N_shard offset) and distances, for several GPUs x shards
    per GPU, and the single-GPU copy-through of the first KQuery entries"""
    G, z = _host_golden()
    assert [tuple(c) for c in z["merge_cases"]] == G.MERGE_CASES
    for i, case in enumerate(G.MERGE_CASES):
        num_gpus, spg, Nq, K, N_shard = case
        ids, d = G.merge_inputs(case, 100 + i)
        o_ids, o_d = O.merge_results(ids, d, K, spg * N_shard)
        assert np.array_equal(o_ids, z[f"merge{i}_ids"]), case
        assert np.array_equal(o_d, z[f"merge{i}_dists"]), case


def test_oracle_and_product_evaluator_match_reference_evaluator():
    """orc_eval and ggnn_b200.Evaluator == ggnn::Evaluator (src/ggnn/base/eval.cpp:88-242): c@1, c@K, r@K and their
    duplicate-aware variants on float and uint8 data, Euclidean and cosine (incl. the b_norm quirk of eval.cpp:52), bases
    with duplicate rows, K_gt < KQuery and KQuery = 1"""
    import torch
    import ggnn_b200 as ggnn
    G, z = _host_golden()
    assert [tuple(c) for c in z["eval_cases"]] == G.EVAL_CASES
    for i, case in enumerate(G.EVAL_CASES):
        N, Nq, D, Kgt, K, measure, is_u8 = case
        base, query, gt, res = G.eval_inputs(case, 200 + i)
        want = z[f"eval{i}_values"]
        o = O.evaluate(gt, res, K, base, query, measure)
        got = np.array([o[k] for k in ("c1", "c1_dup", "cK", "cK_dup", "rK", "rK_dup")], np.float32)
        assert np.array_equal(got, want), (case, got, want)
        e = ggnn.Evaluator(torch.from_numpy(base), torch.from_numpy(query), torch.from_numpy(gt), K,
                           ggnn.DistanceMeasure(measure)).evaluate_results(torch.from_numpy(res))
        mine = np.array([e.c1, e.c1_dup, e.c_k_query, e.c_k_query_dup, e.r_k_query, e.r_k_query_dup], np.float64)
        assert np.allclose(mine, want.astype(np.float64), rtol=0, atol=1e-6), (case, mine, want)


def test_vecs_io_matches_reference_dataset_store_and_load(tmp_path):
    """fvecs / bvecs / ivecs: tests/golden/ref_store.* were written by the reference's Dataset<T>::store
    (src/ggnn/base/dataset.cu:215-226); our store() must produce the same bytes, our load() the same rows -- whole file
    and the (from, num) ranges the reference's own load returned (dataset.cu:118-213) -- and asking for more rows than
    the file holds must fail like the reference's CHECK_EQ"""
    import os
    import torch
    import ggnn_b200 as ggnn
    G, z = _host_golden()
    cls = {"fvecs": ggnn.FloatDataset, "bvecs": ggnn.UCharDataset, "ivecs": ggnn.IntDataset}
    for i, case in enumerate(G.IO_CASES):
        _, dt, ext, N, D, (lo, num) = case
        data = G.io_inputs(case, 300 + i)
        ref_file = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"ref_store.{ext}")
        mine = str(tmp_path / f"mine.{ext}")
        cls[ext](torch.from_numpy(data)).store(mine)
        assert open(mine, "rb").read() == open(ref_file, "rb").read()
        assert np.array_equal(cls[ext].load(ref_file).tensor.numpy(), data)
        part = cls[ext].load(ref_file, lo, num).tensor.numpy()
        assert np.array_equal(part, z[f"io{i}_loaded"]) and np.array_equal(part, data[lo:lo + num])
        with pytest.raises(ValueError):
            cls[ext].load(ref_file, 1, N)  # fewer vectors than requested
