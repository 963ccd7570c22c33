"""One shard of any size on one GPU: build, then the traversal kernel alone (several launches, CUDA events) with its
pop / distance counters and algorithmic bytes -- the target of the ncu captures on a base much larger than L2
(BASELINE config 4's 12.5M x 128 shard, config 3's 10M x 96 cosine).
  python tools/shard_probe.py N D kind measure [reps] [tau] [max_it]
  ncu --set full -k regex:query_kernel -s 3 -c 1 ... python tools/shard_probe.py 12500000 128 manifold8 0"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import ggnn_b200 as ggnn  # noqa: E402


def main():
    N, D, kind, measure = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
    reps = int(sys.argv[5]) if len(sys.argv) > 5 else 8
    tau = float(sys.argv[6]) if len(sys.argv) > 6 else 0.64
    max_it = int(sys.argv[7]) if len(sys.argv) > 7 else 400
    Nq, K = 10_000, 10
    dev = torch.device("cuda", 0)
    base, query = bench.gen_gpu(N, Nq, D, kind, 1234, dev)
    idx = ggnn.GGNN()
    idx.set_return_results_on_gpu(True)
    idx.set_base(base)
    torch.cuda.synchronize()
    t0 = time.time()
    idx.build(24, 0.5, 2, measure)
    torch.cuda.synchronize()
    rep = {"N": N, "D": D, "kind": kind, "measure": measure, "build_s": time.time() - t0, "tau_query": tau, "max_iterations": max_it}
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * reps)]
    for r in range(reps):
        ev[2 * r].record()
        ids, _ = idx.query(query, K, tau, max_it, measure)
        ev[2 * r + 1].record()
    torch.cuda.synchronize()
    ms = [ev[2 * r].elapsed_time(ev[2 * r + 1]) for r in range(reps)]
    n_iter, n_dist, alg = bench._query_stats(idx, query, K, tau, max_it, measure)
    peak, _ = bench.measured_peaks()
    rep.update({"query_ms": ms, "query_ms_median": float(np.median(ms[1:] or ms)), "pops_per_query": n_iter / Nq,
                "dists_per_query": n_dist / Nq, "algorithmic_bytes_per_launch": alg,
                "algorithmic_gbs": alg / (float(np.median(ms[1:] or ms)) * 1e-3) / 1e9, "hbm_peak_gbs": peak})
    if os.environ.get("PROBE_RECALL"):
        gt, _ = idx.bf_query(query[:2000].contiguous(), K, measure)
        rep["recall_at_10_first_2000"] = bench.recall_at_k(gt, ids[:2000], K)
    if os.environ.get("PROBE_BF"):  # exact ground truth for the whole batch: time, recall, checksum
        idx.bf_query(query, K, measure)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gt, _ = idx.bf_query(query, K, measure)
        e1.record()
        torch.cuda.synchronize()
        rep["bf_query_ms"] = e0.elapsed_time(e1)
        rep["recall_at_10"] = bench.recall_at_k(gt, ids, K)
        rep["bf_ids_crc32"] = bench._crc(gt)
    print(json.dumps(rep), flush=True)


if __name__ == "__main__":
    main()
