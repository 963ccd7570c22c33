"""Operating-point sweep on the GPU box (procedure of the reference's ggnn_benchmark.cpp:186-200): build once,
ground truth by bf_query, then tau_query x max_iterations -> recall@10 and kernel ms.  Because ggnn_b200's
query kernel returns results identical to the reference's on the same graph, the recall column is the
reference's too; the reference's time at selected points comes from oracle/_ref/ref_driver on OUR graph file."""
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
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    kinds = sys.argv[2].split(",") if len(sys.argv) > 2 else ["clustered", "uniform"]
    Nq, D, K = 10_000, 128, 10
    dev = torch.device("cuda", 0)
    out = {}
    for kind in kinds:
        base, query = bench.gen_gpu(N, Nq, D, kind, 1234, dev)
        idx = ggnn.GGNN()
        idx.set_return_results_on_gpu(True)
        idx.set_base(base)
        torch.cuda.synchronize()
        t0 = time.time()
        idx.build(24, 0.5, 2)
        torch.cuda.synchronize()
        build_s = time.time() - t0
        t0 = time.time()
        gt, _ = idx.bf_query(query, K)
        torch.cuda.synchronize()
        bf_s = time.time() - t0
        rows = []
        quick = len(sys.argv) > 3 and sys.argv[3] == "quick"
        for max_it in ((200, 400) if quick else (200, 400, 1000)):
            for tau in ((0.34, 0.41, 0.51, 0.64, 0.8, 1.0) if quick else (0.34, 0.41, 0.51, 0.64, 0.8, 1.0, 1.5, 2.0)):
                idx.query(query, K, tau, max_it)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ids, _ = idx.query(query, K, tau, max_it)
                e1.record()
                torch.cuda.synchronize()
                rows.append({"tau": tau, "max_it": max_it, "recall": bench.recall_at_k(gt, ids, K), "ms": e0.elapsed_time(e1)})
                print(kind, rows[-1], flush=True)
        rep = {"N": N, "build_s": build_s, "bf_s": bf_s, "rows": rows, "nn1_stats": idx.get_graph(0).nn1_stats.tolist()}
        # reference timing on OUR graph at a few operating points
        wd = os.path.join("/tmp", f"sweep_{kind}")
        os.makedirs(wd, exist_ok=True)
        base.cpu().numpy().tofile(os.path.join(wd, "base.bin"))
        query.cpu().numpy().tofile(os.path.join(wd, "query.bin"))
        idx.set_working_directory(wd)
        idx.store()
        refs = []
        for tau, max_it in (((0.64, 400),) if quick else ((0.51, 200), (0.64, 400), (1.0, 400), (1.5, 1000))):
            try:
                r = run_ref(wd, n=N, nq=Nq, d=D, measure=0, kbuild=24, build=0, kquery=K, tau_query=tau, max_iter=max_it,
                            query_reps=3, gpu_reps=3, bf=0, dump=0)
                refs.append({"tau": tau, "max_it": max_it, "e2e_ms": r["query_e2e_ms"], "kernel_ms": r["query_gpu_kernel_ms"]})
                print(kind, "reference", refs[-1], flush=True)
            except Exception as e:
                refs.append({"tau": tau, "max_it": max_it, "error": str(e)[-300:]})
        # one full reference build for the build-time comparison
        try:
            if quick:
                raise RuntimeError("skipped (quick)")
            r = run_ref(wd, n=N, nq=Nq, d=D, measure=0, kbuild=24, tau_build=0.5, refine=2, build=1, kquery=K, tau_query=0.64,
                        max_iter=400, query_reps=2, gpu_reps=0, bf=K, dump=1)
            rid = np.fromfile(os.path.join(wd, "query_ids.bin"), np.int32).reshape(Nq, K)
            rgt = np.fromfile(os.path.join(wd, "bf_ids.bin"), np.int32).reshape(Nq, K)
            rep["reference_build"] = {"build_s": r["build_s"], "bf_s": r["bf_s"], "e2e_ms": r["query_e2e_ms"],
                                      "recall_tau0.64_it400": bench.recall_at_k(torch.from_numpy(rgt), torch.from_numpy(rid), K),
                                      "bf_ids_equal_ours": bool(np.array_equal(rgt, gt.cpu().numpy()))}
            print(kind, "reference build", rep["reference_build"], flush=True)
        except Exception as e:
            rep["reference_build"] = {"error": str(e)[-300:]}
        rep["reference_on_our_graph"] = refs
        for f in os.listdir(wd):
            os.remove(os.path.join(wd, f))
        out[kind] = rep
        del idx, base, query
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"sweep_{N}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
