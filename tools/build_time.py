"""Build time of the 1M x 128 bench graph for different staging modes of the merge kernel + recall check."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import ggnn_b200 as ggnn  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "manifold8"
base, query = bench.gen_gpu(1_000_000, 10_000, 128, kind, 1234, torch.device("cuda", 0))
for mode, pers in (("0", "1"), ("3", "1"), ("0", "1"), ("3", "1")):
    os.environ["GGNN_B200_BUILD_STAGE_MODE"] = mode
    os.environ["GGNN_B200_BUILD_PERSISTENT"] = pers
    idx = ggnn.GGNN()
    idx.set_return_results_on_gpu(True)
    idx.set_base(base)
    torch.cuda.synchronize()
    t = time.time()
    idx.build(24, 0.5, 2)
    torch.cuda.synchronize()
    dt = time.time() - t
    ids, _ = idx.query(query, 10, 0.64, 400)
    gt, _ = idx.bf_query(query, 10)
    rec = ggnn.Evaluator(None, None, gt.cpu(), 10).evaluate_results(ids.cpu()).c_k_query
    print(f"build stage mode {mode} persistent {pers}: {dt:.3f} s, recall@10 {rec:.4f}", flush=True)
    del idx
