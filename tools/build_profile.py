"""Build one 1M x 128 graph (bench workload); used under ncu to capture the construction kernels."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import ggnn_b200 as ggnn  # noqa: E402

base, _ = bench.gen_gpu(1_000_000, 16, 128, bench.DEF["kind"], 1234, torch.device("cuda", 0))
g = ggnn.GGNN()
g.set_base(base)
torch.cuda.synchronize()
t0 = time.time()
g.build(24, 0.5, 2)
torch.cuda.synchronize()
print("build_s", time.time() - t0)
