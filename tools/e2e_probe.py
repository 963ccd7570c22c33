"""End-to-end (pinned host -> device -> pinned host) queries/s of GGNN.query_async() against the number of batches
kept in flight, and of the synchronous GGNN.query()."""
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
idx = ggnn.GGNN()
idx.set_base(base)
idx.build(24, 0.5, 2)
q_host = query.cpu().pin_memory()
steps = 40


def run_async(n, depth):
    pending, last = [], None
    for _ in range(n):
        pending.append(idx.query_async(q_host, 10, 0.64, 400))
        if len(pending) >= depth:
            last = pending.pop(0).result()
    while pending:
        last = pending.pop(0).result()
    return last


for depth in (1, 2, 3, 4, 6, 8):
    run_async(8, depth)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run_async(steps, depth)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    print(f"query_async depth {depth}: {ms:.3f} ms/step = {10_000 / ms / 1e3:.2f} M q/s", flush=True)
for chunks in (1, 2, 4):
    os.environ["GGNN_B200_QUERY_CHUNKS"] = str(chunks)
    idx.query(q_host, 10, 0.64, 400)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        idx.query(q_host, 10, 0.64, 400)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    print(f"query (sync) {chunks} chunk(s): {ms:.3f} ms/step = {10_000 / ms / 1e3:.2f} M q/s", flush=True)
# host-side cost of one enqueue (no waiting)
t0 = time.perf_counter()
fs = [idx.query_async(q_host, 10, 0.64, 400) for _ in range(4)]
t1 = time.perf_counter()
for f in fs:
    f.result()
print(f"host time per enqueue: {(t1 - t0) * 1e6 / 4:.0f} us", flush=True)
