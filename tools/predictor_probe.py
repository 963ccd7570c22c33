"""Can a cheap pre-pass predict which queries run long (to launch them first)?  Correlates candidate predictors with
the measured pops / distance evaluations per query and times the kernel with the batch ordered by each predictor."""
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
    dev = torch.device("cuda", 0)
    base, query = bench.gen_gpu(1_000_000, 10_000, 128, kind, 1234, dev)
    idx = ggnn.GGNN()
    idx.set_return_results_on_gpu(True)
    idx.set_base(base)
    idx.build(24, 0.5, 2)
    stats = torch.zeros((10_000, 2), dtype=torch.int32, device=dev)
    t_nat = run(idx, query, 10, 0.64, 400, stats)
    work = stats[:, 1].float()
    gr = idx.get_graph(0)
    sp = gr.layer_translation(3).long()
    d_sp = torch.cdist(query, base[sp]) ** 2                    # [Nq, 32] squared distances to the start points
    nn1_max = gr.nn1_stats[1]
    preds = {
        "min_start_dist": d_sp.min(1).values,
        "mean_start_dist": d_sp.mean(1),
        "kth10_start_dist": d_sp.kthvalue(10, dim=1).values,
        "spread_start": d_sp.kthvalue(10, dim=1).values - d_sp.min(1).values,
        "query_norm": (query * query).sum(1),
    }
    # two-hop: distance to the best start point's neighbours
    best_sp = sp[d_sp.argmin(1)]
    nb = gr.layer_graph(0)[best_sp].long().clamp_min(0)          # [Nq, 24]
    d_nb = ((base[nb] - query[:, None, :]) ** 2).sum(-1)
    preds["min_2hop_dist"] = d_nb.min(1).values
    preds["gain_2hop"] = d_sp.min(1).values - d_nb.min(1).values
    print(f"{kind}: natural order {t_nat:.3f} ms; oracle LPT {run(idx, query[torch.argsort(work, descending=True)].contiguous(), 10, 0.64, 400):.3f} ms")
    for name, p in preds.items():
        c = torch.corrcoef(torch.stack([p.float(), work]))[0, 1].item()
        order = torch.argsort(p, descending=(c > 0))
        t = run(idx, query[order].contiguous(), 10, 0.64, 400)
        print(f"  {name:18s} corr {c:+.3f}  ordered-by-predictor {t:.3f} ms", flush=True)
    # random order (is the natural order unlucky?)
    t = run(idx, query[torch.randperm(10_000, device=dev)].contiguous(), 10, 0.64, 400)
    print(f"  random order {t:.3f} ms")


if __name__ == "__main__":
    main()
