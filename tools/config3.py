"""BASELINE config 3 (DEEP10M-shape): 10 000 000 x 96 fp32, cosine, k_query 10, build + query on one B200,
next to the unmodified reference (oracle/_ref) on the same data.  usage: python tools/config3.py [N] [Nq]"""
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
from tools.gpu_check import run_ref  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    Nq = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
    D, K, tau, it = 96, 10, 0.64, 400
    dev = torch.device("cuda", 0)
    base, query = bench.gen_gpu(N, Nq, D, "manifoldcos8", 1234, dev)
    idx = ggnn.GGNN()
    idx.set_return_results_on_gpu(True)
    idx.set_base(base)
    torch.cuda.synchronize()
    t0 = time.time()
    idx.build(24, 0.5, 2, ggnn.DistanceMeasure.Cosine)
    torch.cuda.synchronize()
    rep = {"N": N, "Nq": Nq, "D": D, "measure": "cosine", "ours_build_s": time.time() - t0}
    t0 = time.time()
    gt, _ = idx.bf_query(query, K, ggnn.DistanceMeasure.Cosine)
    torch.cuda.synchronize()
    rep["ours_bf_s"] = time.time() - t0
    idx.query(query, K, tau, it, ggnn.DistanceMeasure.Cosine)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ids, _ = idx.query(query, K, tau, it, ggnn.DistanceMeasure.Cosine)
    e1.record()
    torch.cuda.synchronize()
    rep["ours_query_ms"] = e0.elapsed_time(e1) / 5
    rep["ours_recall"] = bench.recall_at_k(gt, ids, K)
    print(json.dumps(rep), flush=True)
    wd = "/tmp/config3"
    os.makedirs(wd, exist_ok=True)
    base.cpu().numpy().tofile(os.path.join(wd, "base.bin"))
    query.cpu().numpy().tofile(os.path.join(wd, "query.bin"))
    gt_h = gt.cpu()
    del idx, base
    torch.cuda.empty_cache()
    try:
        r = run_ref(wd, n=N, nq=Nq, d=D, measure=1, kbuild=24, tau_build=0.5, refine=2, build=1, kquery=K, tau_query=tau,
                    max_iter=it, query_reps=4, gpu_reps=4, bf=K, dump=1)
        rid = np.fromfile(os.path.join(wd, "query_ids.bin"), np.int32).reshape(Nq, K)
        rgt = np.fromfile(os.path.join(wd, "bf_ids.bin"), np.int32).reshape(Nq, K)
        rep["reference"] = {"build_s": r["build_s"], "bf_s": r["bf_s"], "query_e2e_ms": r["query_e2e_ms"],
                            "query_gpu_kernel_ms": r["query_gpu_kernel_ms"],
                            "recall": bench.recall_at_k(torch.from_numpy(rgt), torch.from_numpy(rid), K),
                            "bf_ids_equal_ours": bool(np.array_equal(rgt, gt_h.numpy()))}
    except Exception as e:  # pragma: no cover
        rep["reference"] = {"error": str(e)[-400:]}
    for f in os.listdir(wd):
        os.remove(os.path.join(wd, f))
    print(json.dumps(rep), flush=True)
    json.dump(rep, open(os.path.join(ROOT, "gpurun_out", f"config3_{N}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
