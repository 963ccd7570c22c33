"""The reference's benchmark procedure (examples/cpp-and-cuda/ggnn_benchmark.cpp, docs/source/benchmarking.rst) on the
Python API of this repo -- same flags, same flow: load fvecs/bvecs, build or load the graph, ground truth from an
ivecs file or by bf_query (exported if a path is given), Evaluator, then either the four documented tau_query values
(0.34 / 0.41 / 0.51 / 0.64) or the grid search (0.00..0.69 step 0.01, 0.7..2.0 step 0.1), with the time of every query
call.  (The reference's own, unmodified C++ program also runs on this library: examples/ref_ggnn_benchmark_on_b200.)

  python tools/ggnn_benchmark.py --base sift_base.fvecs --query sift_query.fvecs --gt sift_groundtruth.ivecs \\
      --graph_dir /tmp/graphs --k_build 24 --tau_build 0.5 --k_query 10 --max_iterations 400
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def parse_args(argv=None):
    ap = argparse.ArgumentParser(description="GGNN benchmark procedure on ggnn_b200")
    ap.add_argument("--base", required=True, help="file with base vectors (fvecs/bvecs)")
    ap.add_argument("--subset", type=int, default=0, help="number of base vectors to use")
    ap.add_argument("--query", required=True, help="file with query vectors (fvecs/bvecs)")
    ap.add_argument("--gt", default="", help="file with ground-truth indices (ivecs); computed and exported if missing")
    ap.add_argument("--graph_dir", default="", help="directory to store and load ggnn graph files")
    ap.add_argument("--k_build", type=int, default=24)
    ap.add_argument("--tau_build", type=float, default=0.5)
    ap.add_argument("--refinement_iterations", type=int, default=2)
    ap.add_argument("--k_query", type=int, default=10)
    ap.add_argument("--max_iterations", type=int, default=200)
    ap.add_argument("--measure", default="euclidean", choices=["euclidean", "cosine"])
    ap.add_argument("--shard_size", type=int, default=0)
    ap.add_argument("--gpu_ids", default="0", help="GPU ids, separated by spaces")
    ap.add_argument("--grid_search", action="store_true")
    a = ap.parse_args(argv)
    if a.tau_build < 0 or a.refinement_iterations < 0:
        ap.error("tau_build and refinement_iterations have to be non-negative")
    a.gpus = [int(g) for g in a.gpu_ids.split()]
    return a


def tau_values(grid_search):
    if grid_search:  # ggnn_benchmark.cpp:186-192
        return [i * 0.01 for i in range(70)] + [i * 0.1 for i in range(7, 21)]
    return [0.34, 0.41, 0.51, 0.64]  # :193-200


def dataset_class(ggnn, path):
    if path.endswith(".fvecs"):
        return ggnn.FloatDataset
    if path.endswith(".bvecs"):
        return ggnn.UCharDataset
    if path.endswith(".ivecs"):
        return ggnn.IntDataset
    raise SystemExit(f"Could not guess file type from {path}. fvecs, bvecs, or ivecs file required.")


def main(argv=None):
    a = parse_args(argv)
    import ggnn_b200 as ggnn
    for f in (a.base, a.query):
        if not os.path.exists(f):
            raise SystemExit(f"file has to exist: {f}")
    measure = ggnn.DistanceMeasure.Euclidean if a.measure == "euclidean" else ggnn.DistanceMeasure.Cosine
    base = dataset_class(ggnn, a.base).load(a.base, 0, a.subset or 2 ** 32 - 1, True)
    query = dataset_class(ggnn, a.query).load(a.query, 0, 2 ** 32 - 1, True)
    print(f"base {base.N} x {base.D}, query {query.N} x {query.D}", flush=True)

    idx = ggnn.GGNN()
    idx.set_working_directory(a.graph_dir or ".")
    idx.set_base(base.tensor)
    idx.set_gpus(a.gpus)
    idx.set_shard_size(a.shard_size)
    if a.graph_dir and os.path.isfile(os.path.join(a.graph_dir, "part_0.ggnn")):
        idx.load(a.k_build)
        print("graph loaded", flush=True)
    else:
        t0 = time.time()
        idx.build(a.k_build, a.tau_build, a.refinement_iterations, measure)
        print(f"graph built in {time.time() - t0:.2f} s", flush=True)
        if a.graph_dir:
            idx.store()

    if a.gt and os.path.isfile(a.gt):
        gt = ggnn.IntDataset.load(a.gt).tensor
    else:
        if len(a.gpus) > 1:
            raise SystemExit("bf_query supports a single GPU: provide a ground-truth file")
        gt, _ = idx.bf_query(query.tensor, 100, measure)
        if a.gt:
            ggnn.IntDataset(gt.cpu()).store(a.gt)
    ev = ggnn.Evaluator(base.tensor, query.tensor, gt, a.k_query, measure)

    for tau in tau_values(a.grid_search):
        t0 = time.time()
        ids, _ = idx.query(query.tensor, a.k_query, tau, a.max_iterations, measure)
        dt = time.time() - t0
        print(f"--\nQuery with tau_query {tau:.2f} max iterations {a.max_iterations}: {dt * 1e3:.2f} ms "
              f"({query.N / dt:.0f} queries/s)\n{ev.evaluate_results(ids)}", flush=True)


if __name__ == "__main__":
    main()
