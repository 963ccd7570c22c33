"""A/B timing of the traversal kernel for one build of the library (GGNN_B200_LIB=... selects it): the bench workload
(1M x 128 manifold8, 10 000 queries, k 10, tau 0.64, 400 iterations), fp32 and native uint8 rows, one launch alone and
two streams alternating; results are checksummed so that variants can be checked against each other."""
import json
import os
import sys
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import ggnn_b200 as ggnn  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    D = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    kind = sys.argv[3] if len(sys.argv) > 3 else "manifold8"
    measure = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    dev = torch.device("cuda", 0)
    base, qs = bench.gen_gpu(N, 40_000, D, kind, 1234, dev)
    batches = [qs[i * 10000:(i + 1) * 10000] for i in range(4)]
    idx = ggnn.GGNN()
    idx.set_return_results_on_gpu(True)
    idx.set_base(base)
    idx.build(24, 0.5, 2, measure)
    if measure or kind != "manifold8":   # no uint8 twin: fp32 only
        out = {"lib": os.environ.get("GGNN_B200_LIB", "default"), "N": N, "D": D, "kind": kind, "measure": measure}
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(80)]
        idx.query(batches[0], 10, 0.64, 400, measure)
        for i in range(40):
            ev[2 * i].record()
            idx.query(batches[i % 4], 10, 0.64, 400, measure)
            ev[2 * i + 1].record()
        torch.cuda.synchronize()
        out["single_ms"] = float(np.median([ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(40)]))
        print(json.dumps(out), flush=True)
        return
    g8 = ggnn.GGNN()
    g8.set_return_results_on_gpu(True)
    g8.set_base(base.to(torch.uint8))
    g8._prepare(24)
    g8._shards[0].graph = idx.get_graph(0)
    g8._measure = 0
    out = {"lib": os.environ.get("GGNN_B200_LIB", "default")}
    streams = [torch.cuda.Stream(dev) for _ in range(2)]
    for tag, g, conv in (("f32", idx, lambda q: q), ("u8", g8, lambda q: q.to(torch.uint8))):
        qb = [conv(q) for q in batches]
        r = g.query(qb[0], 10, 0.64, 400)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(80)]
        for i in range(40):
            ev[2 * i].record()
            g.query(qb[i % 4], 10, 0.64, 400)
            ev[2 * i + 1].record()
        torch.cuda.synchronize()
        single = float(np.median([ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(40)]))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream(dev)
        e0.record(cur)
        for s in streams:
            s.wait_event(e0)
        for i in range(80):
            with torch.cuda.stream(streams[i % 2]):
                g.query(qb[i % 4], 10, 0.64, 400)
        for s in streams:
            cur.wait_stream(s)
        e1.record(cur)
        torch.cuda.synchronize()
        out[tag] = {"single_ms": single, "piped_ms": e0.elapsed_time(e1) / 80,
                    "crc": zlib.crc32(r[0].cpu().numpy().tobytes() + r[1].cpu().numpy().tobytes()) & 0xffffffff}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
