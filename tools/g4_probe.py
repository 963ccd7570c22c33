"""Correctness probe of the gather4 staging mode on a small problem (run before timing it)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ggnn_b200 as ggnn  # noqa: E402

rng = np.random.default_rng(1)
base = torch.from_numpy(rng.random((20000, 128), dtype=np.float32))
query = torch.from_numpy(rng.random((512, 128), dtype=np.float32))
g = ggnn.GGNN()
g.set_base(base)
g.build(24, 0.5)
os.environ["GGNN_B200_STAGE_MODE"] = "0"
i0, d0 = g.query(query, 10, 0.64, 400)
for pad in ("1", "0"):
    os.environ["GGNN_B200_STAGE_MODE"] = "3"
    os.environ["GGNN_B200_GATHER4_PAD_VALID"] = pad
    i3, d3 = g.query(query, 10, 0.64, 400)
    print("pad_valid", pad, "ids equal", bool(torch.equal(i0, i3)), "dists equal", bool(torch.equal(d0, d3)), flush=True)
