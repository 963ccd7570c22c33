"""bf_query timing for one K on the bench workload under different tensor-path settings (env switches of bf_tc.cu)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import ggnn_b200 as ggnn  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dev = torch.device("cuda", 0)
base, query = bench.gen_gpu(1_000_000, 10_000, 128, "manifold8", 1234, dev)
g = ggnn.GGNN()
g.set_return_results_on_gpu(True)
g.set_base(base)
g._prepare(24)
ref = None
for merger in ("1", "0"):
    for splits in ("0", "2", "4", "6"):
        os.environ["GGNN_B200_BF_MERGER"] = merger
        os.environ["GGNN_B200_BF_SPLITS"] = splits
        g.bf_query(query, K)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        for r in range(3):
            ev[2 * r].record()
            ids, d = g.bf_query(query, K)
            ev[2 * r + 1].record()
        torch.cuda.synchronize()
        if ref is None:
            ref = ids.clone()
        print(json.dumps({"K": K, "merger": merger, "splits": splits,
                          "ms": float(np.median([ev[2 * r].elapsed_time(ev[2 * r + 1]) for r in range(3)])),
                          "same_ids": bool(torch.equal(ids, ref))}), flush=True)
