"""Is the traversal kernel's time set by throughput or by its longest queries?  Prints the distribution of
pops per query and times the same batch in natural order, longest-first (LPT) and shortest-first order."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import ggnn_b200 as ggnn  # noqa: E402
from ggnn_b200 import _lib  # noqa: E402


def run(idx, query, K, tau, it, stats=None):
    gr = idx.get_graph(0)
    cfg = gr.config
    Nq = query.shape[0]
    ids = torch.empty((Nq, K), dtype=torch.int32, device="cuda")
    d = torch.empty((Nq, K), dtype=torch.float32, device="cuda")
    wc = torch.zeros(1, dtype=torch.int32, device="cuda")
    p = _lib.QueryParams()
    p.D, p.measure, p.KQuery, p.tau_query, p.max_iterations = cfg.D, 0, K, tau, it
    p.N_base, p.KBuild, p.num_starting_points = cfg.N, cfg.KBuild, cfg.S
    p.d_base, p.d_query, p.d_graph = idx._shards[0].base.data_ptr(), query.data_ptr(), gr.graph.data_ptr()
    p.d_starting_points, p.d_nn1_stats = gr.layer_translation(3).data_ptr(), gr.nn1_stats.data_ptr()
    p.d_query_results, p.d_query_results_dists = ids.data_ptr(), d.data_ptr()
    p.d_stats = stats.data_ptr() if stats is not None else None
    p.shards_per_gpu, p.on_gpu_shard_id, p.d_work_counter = 1, 0, wc.data_ptr()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(_lib.lib().ggnn_b200_query(C.byref(p), Nq, st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        _lib.check(_lib.lib().ggnn_b200_query(C.byref(p), Nq, st))
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "manifold8"
    base, query = bench.gen_gpu(1_000_000, 10_000, 128, kind, 1234, torch.device("cuda", 0))
    idx = ggnn.GGNN()
    idx.set_return_results_on_gpu(True)
    idx.set_base(base)
    idx.build(24, 0.5, 2)
    stats = torch.zeros((10_000, 2), dtype=torch.int32, device="cuda")
    t_nat = run(idx, query, 10, 0.64, 400, stats)
    pops = stats[:, 0].float()
    dists = stats[:, 1].float()
    qs = torch.tensor([0.5, 0.9, 0.99, 0.999, 1.0], device="cuda")
    print(kind, "pops quantiles", torch.quantile(pops, qs).tolist(), "mean", pops.mean().item())
    print(kind, "dists quantiles", torch.quantile(dists, qs).tolist(), "mean", dists.mean().item())
    order = torch.argsort(dists, descending=True)
    t_lpt = run(idx, query[order].contiguous(), 10, 0.64, 400)
    t_spt = run(idx, query[order.flip(0)].contiguous(), 10, 0.64, 400)
    t_rep = run(idx, query.repeat(4, 1), 10, 0.64, 400)
    print(kind, f"natural {t_nat:.3f} ms | longest-first {t_lpt:.3f} ms | shortest-first {t_spt:.3f} ms | 4x batch {t_rep:.3f} ms ({t_rep/4:.3f} per 10k)")


if __name__ == "__main__":
    main()
