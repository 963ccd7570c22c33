"""Launch-parameter exploration for the traversal kernel on the GPU box (build once, time many settings).
usage: [GGNN_B200_LIB=...] python tools/tune_query.py kinds rows_list modes_list tag"""
import itertools
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import ggnn_b200 as ggnn  # noqa: E402


def main():
    kinds = sys.argv[1].split(",") if len(sys.argv) > 1 else ["manifold8", "manifold16"]
    rows_l = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "8,16,24").split(",")]
    modes = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "0,1").split(",")]
    tag = sys.argv[4] if len(sys.argv) > 4 else "default"
    dev = torch.device("cuda", 0)
    res = []
    for kind in kinds:
        base, query = bench.gen_gpu(1_000_000, 10_000, 128, kind, 1234, dev)
        idx = ggnn.GGNN()
        idx.set_return_results_on_gpu(True)
        idx.set_base(base)
        idx.build(24, 0.5, 2)
        ref_ids = None
        for pf, mode, rows in itertools.product((1, 2), modes, rows_l):
            os.environ["GGNN_B200_QUERY_PREFETCH"] = str(pf)
            os.environ["GGNN_B200_STAGE_MODE"] = str(mode)
            os.environ["GGNN_B200_QUERY_STAGE_ROWS"] = str(rows)
            try:
                idx.query(query, 10, 0.64, 400)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    ids, _ = idx.query(query, 10, 0.64, 400)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 5
                if ref_ids is None:
                    ref_ids = ids.clone()
                same = bool(torch.equal(ids, ref_ids))
            except Exception as e:
                ms, same = None, str(e)[:80]
            res.append({"tag": tag, "kind": kind, "prefetch": pf, "stage_mode": mode, "stage_rows": rows, "ms": ms, "same": same})
            print(res[-1], flush=True)
        del idx, base, query
        torch.cuda.empty_cache()
    with open(os.path.join(ROOT, "gpurun_out", "tune_query.jsonl"), "a") as f:
        for r in res:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
