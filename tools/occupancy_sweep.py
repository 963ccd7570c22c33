"""Resident warps per SM vs. per-warp shared memory for the traversal kernel: times one 10k batch and a 4x batch
(throughput regime) for combinations of stage rows and visited-hash size.  Results -> gpurun_out/occupancy_sweep.jsonl
usage: python tools/occupancy_sweep.py [kind]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import ggnn_b200 as ggnn  # noqa: E402
from tools.tail_probe import run  # noqa: E402


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "manifold8"
    base, query = bench.gen_gpu(1_000_000, 10_000, 128, kind, 1234, torch.device("cuda", 0))
    idx = ggnn.GGNN()
    idx.set_return_results_on_gpu(True)
    idx.set_base(base)
    idx.build(24, 0.5, 2)
    q4 = query.repeat(4, 1)
    ref = None
    out = []
    for rows, hs, mode, pad in ((16, 0, 0, 0), (16, 0, 3, 0)):
        os.environ["GGNN_B200_QUERY_STAGE_ROWS"] = str(rows)
        os.environ["GGNN_B200_QUERY_HASH_SLOTS"] = str(hs)
        os.environ["GGNN_B200_STAGE_MODE"] = str(mode)
        os.environ["GGNN_B200_GATHER4_PAD_VALID"] = str(pad)
        t1 = run(idx, query, 10, 0.64, 400)
        t4 = run(idx, q4, 10, 0.64, 400)
        ids, _ = idx.query(query, 10, 0.64, 400)
        if ref is None:
            ref = ids.clone()
        r = {"kind": kind, "stage_rows": rows, "hash_slots": hs or 512, "stage_mode": mode, "pad_valid": pad, "ms_10k": t1, "ms_per_10k_at_40k": t4 / 4,
             "same_ids": bool(torch.equal(ids, ref))}
        print(r, flush=True)
        out.append(r)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "occupancy_sweep.jsonl"), "a") as f:
        for r in out:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
