"""Tensor-core bf_query vs the exact SIMT scan (both through GGNN.bf_query): ids and distances must be identical."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import ggnn_b200 as ggnn  # noqa: E402


def run(base, query, K, tc):
    os.environ["GGNN_B200_BF_TC"] = "1" if tc else "0"
    g = ggnn.GGNN()
    g.set_return_results_on_gpu(True)
    g.set_base(base)
    ids, d = g.bf_query(query, K)
    torch.cuda.synchronize()
    t0 = time.time()
    ids, d = g.bf_query(query, K)
    torch.cuda.synchronize()
    return ids, d, (time.time() - t0) * 1e3


def main():
    dev = torch.device("cuda", 0)
    cases = [(20_000, 500, 128, 10, "manifold8"), (5_000, 130, 64, 32, "uniform"), (100_000, 1000, 96, 10, "manifold16"),
             (1_000_000, 10_000, 128, 10, "manifold8"), (1_000_000, 10_000, 128, 10, "uniform")]
    if len(sys.argv) > 1:
        cases = cases[:int(sys.argv[1])]
    if os.environ.get("BF_CASE"):
        cases = [cases[int(os.environ["BF_CASE"])]]
    ok = True
    for N, Nq, D, K, kind in cases:
        base, query = bench.gen_gpu(N, Nq, D, kind, 7, dev)
        base[17] = base[3]  # exact duplicate rows -> tie on (dist, idx)
        i1, d1, t1 = run(base, query, K, True)
        i0, d0, t0 = run(base, query, K, False)
        same = bool(torch.equal(i0, i1) and torch.equal(d0, d1))
        frac = float((i0 == i1).all(1).float().mean())
        print(f"N={N} Nq={Nq} D={D} K={K} {kind}: identical={same} (rows equal {frac:.5f}) tensor {t1:.2f} ms | exact SIMT {t0:.2f} ms", flush=True)
        ok = ok and same
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
