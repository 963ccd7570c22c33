"""Warm build time (third build of the process) of the 1M x 128 bench graph for the loaded library + recall check."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import ggnn_b200 as ggnn  # noqa: E402

base, query = bench.gen_gpu(1_000_000, 10_000, 128, "manifold8", 1234, torch.device("cuda", 0))
ts = []
for _ in range(3):
    idx = ggnn.GGNN()
    idx.set_return_results_on_gpu(True)
    idx.set_base(base)
    torch.cuda.synchronize()
    t = time.time()
    idx.build(24, 0.5, 2)
    torch.cuda.synchronize()
    ts.append(time.time() - t)
ids, _ = idx.query(query, 10, 0.64, 400)
gt, _ = idx.bf_query(query, 10)
print(json.dumps({"lib": os.environ.get("GGNN_B200_LIB", "default"), "sym_warps": os.environ.get("GGNN_B200_SYM_WARPS_PER_SM"),
                  "build_s": ts, "recall": bench.recall_at_k(gt, ids, 10)}), flush=True)
