"""GPU-box check: run the UNMODIFIED reference (oracle/_ref) and ggnn_b200 side by side on the same
inputs, compare results and timings, and (re)generate the golden fixtures under tests/golden/.

  python tools/gpu_check.py golden          # config-1 sized fixtures (N=10000, D=128) -> gpurun_out/golden/
  python tools/gpu_check.py compare N Nq D  # larger side-by-side comparison (no fixtures)

TEST / MEASUREMENT TOOLING: it executes oracle/_ref and the CPU oracle as checkers only.
"""
import json
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ggnn_b200 as ggnn  # noqa: E402
from ggnn_b200 import _lib  # noqa: E402
from oracle import pyoracle as O  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "ref_driver")


def gen(N, Nq, D, seed=1234, kind="uniform"):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.random((N, D), dtype=np.float32), rng.random((Nq, D), dtype=np.float32)
    if kind == "clustered":  # SIFT-like: mixture of Gaussians, clipped to [0,255], integers stored as fp32
        nc = 1000
        centers = rng.random((nc, D), dtype=np.float32) * 160 + 20
        def draw(n):
            c = rng.integers(0, nc, n)
            x = centers[c] + rng.standard_normal((n, D), dtype=np.float32) * 25
            return np.clip(np.rint(x), 0, 255).astype(np.float32)
        return draw(N), draw(Nq)
    if kind == "normal":  # DEEP-like: unit-normalised Gaussian
        def draw(n):
            x = rng.standard_normal((n, D), dtype=np.float32)
            return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)
        return draw(N), draw(Nq)
    raise ValueError(kind)


def run_ref(workdir, **kw):
    args = [REF, f"dir={workdir}"] + [f"{k}={v}" for k, v in kw.items()]
    t0 = time.time()
    p = subprocess.run(args, capture_output=True, text=True)
    dt = time.time() - t0
    if p.returncode != 0:
        raise RuntimeError(f"ref_driver failed rc={p.returncode}\n{p.stdout[-2000:]}\n{p.stderr[-4000:]}")
    line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    out["wall_s"] = dt
    return out


def load_blob(path, cfg):
    return O.Graph(cfg, np.fromfile(path, dtype=np.uint8))


def recall(gt, ids, K):
    hits = 0
    for a, b in zip(gt[:, :K], ids[:, :K]):
        hits += len(set(a.tolist()) & set(b.tolist()))
    return hits / (gt.shape[0] * K)


def timed(fn, reps=3):
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.time()
        r = fn()
        torch.cuda.synchronize()
        ts.append((time.time() - t0) * 1e3)
    return r, ts


def side_by_side(workdir, N, Nq, D, measure=0, kind="uniform", kbuild=24, tau_build=0.5, refine=2, K=10, tau_q=0.64,
                 max_it=400, bf_k=10, oracle_check=True, report=None):
    os.makedirs(workdir, exist_ok=True)
    report = report if report is not None else {}
    base, query = gen(N, Nq, D, kind=kind)
    base.tofile(os.path.join(workdir, "base.bin"))
    query.tofile(os.path.join(workdir, "query.bin"))

    # ---- reference: build + store + bf + query (own graph) ----
    ref = run_ref(workdir, n=N, nq=Nq, d=D, measure=measure, kbuild=kbuild, tau_build=tau_build, refine=refine,
                  build=1, kquery=K, tau_query=tau_q, max_iter=max_it, query_reps=4, gpu_reps=4, bf=bf_k, dump=1)
    report["reference"] = ref
    cfg_o = O.graph_config(N, D, kbuild)
    ref_graph = load_blob(os.path.join(workdir, "part_0.ggnn"), cfg_o)
    ref_ids = np.fromfile(os.path.join(workdir, "query_ids.bin"), np.int32).reshape(Nq, K)
    ref_d = np.fromfile(os.path.join(workdir, "query_dists.bin"), np.float32).reshape(Nq, K)
    ref_bf = np.fromfile(os.path.join(workdir, "bf_ids.bin"), np.int32).reshape(Nq, bf_k)
    ref_bfd = np.fromfile(os.path.join(workdir, "bf_dists.bin"), np.float32).reshape(Nq, bf_k)
    report["ref_recall"] = recall(ref_bf, ref_ids, K)

    # ---- ours on the REFERENCE graph ----
    g = ggnn.GGNN()
    g.set_working_directory(workdir)
    g.set_base(torch.from_numpy(base))
    g.load(kbuild)
    tq = torch.from_numpy(query).pin_memory()
    (ids, dists), t_e2e = timed(lambda: g.query(tq, K, tau_q, max_it, measure))
    g.set_return_results_on_gpu(True)
    tq_gpu = tq.cuda()
    _, t_gpu = timed(lambda: g.query(tq_gpu, K, tau_q, max_it, measure), reps=5)
    g.set_return_results_on_gpu(False)
    ids, dists = ids.numpy(), dists.numpy()
    report["ours_on_ref_graph"] = {
        "e2e_ms": t_e2e, "gpu_ms": t_gpu,
        "ids_equal_frac": float((ids == ref_ids).all(axis=1).mean()),
        "ids_exact": bool(np.array_equal(ids, ref_ids)),
        "dists_exact": bool(np.array_equal(dists, ref_d)),
        "dists_max_rel": float(np.max(np.abs(dists - ref_d) / np.maximum(np.abs(ref_d), 1e-30))) if np.isfinite(ref_d).all() else None,
        "recall": recall(ref_bf, ids, K),
    }
    (bf_ids, bf_d), t_bf = timed(lambda: g.bf_query(tq, bf_k, measure), reps=2)
    bf_ids, bf_d = bf_ids.numpy(), bf_d.numpy()
    report["bf"] = {"ms": t_bf, "ids_exact": bool(np.array_equal(bf_ids, ref_bf)),
                    "dists_exact": bool(np.array_equal(bf_d, ref_bfd)),
                    "ids_equal_frac": float((bf_ids == ref_bf).all(axis=1).mean())}

    if oracle_check:  # CPU oracle vs the reference itself: pins the oracle
        nq_o = min(Nq, 500)
        t0 = time.time()
        o_ids, o_d = O.query(base, query[:nq_o], ref_graph.layer_graph(0), ref_graph.start_points(), ref_graph.nn1_stats,
                             K, tau_q, max_it, measure)
        report["oracle_vs_reference_query"] = {"n": nq_o, "ids_exact": bool(np.array_equal(o_ids, ref_ids[:nq_o])),
                                               "dists_exact": bool(np.array_equal(o_d, ref_d[:nq_o])),
                                               "s": time.time() - t0}
        if N <= 20000:
            nb = min(Nq, 200)
            o_bf, o_bfd = O.bf_query(base, query[:nb], bf_k, measure)
            report["oracle_vs_reference_bf"] = {"n": nb, "ids_exact": bool(np.array_equal(o_bf, ref_bf[:nb])),
                                                "dists_exact": bool(np.array_equal(o_bfd, ref_bfd[:nb]))}

    # ---- ours: own build ----
    g2 = ggnn.GGNN()
    g2.set_base(torch.from_numpy(base))
    torch.cuda.synchronize()
    t0 = time.time()
    g2.build(kbuild, tau_build, refine, measure)
    torch.cuda.synchronize()
    report["ours_build_s"] = time.time() - t0
    own = g2.get_graph(0)
    (ids2, d2), t2 = timed(lambda: g2.query(tq, K, tau_q, max_it, measure))
    ids2 = ids2.numpy()
    sel_eq = bool(np.array_equal(own.selection.cpu().numpy(), ref_graph.selection))
    tr_eq = bool(np.array_equal(own.translation.cpu().numpy(), ref_graph.translation))
    own_g = own.graph.cpu().numpy()
    report["ours_own_graph"] = {
        "e2e_ms": t2, "recall": recall(ref_bf, ids2, K), "selection_exact": sel_eq, "translation_exact": tr_eq,
        "nn1_stats": own.nn1_stats.cpu().numpy().tolist(), "ref_nn1_stats": ref_graph.nn1_stats.tolist(),
        "top_layer_graph_equal_frac": float((own_g[cfg_o.Ns_offsets[3]:] == ref_graph.graph[cfg_o.Ns_offsets[3]:]).mean()),
        "layer0_row_overlap": float(np.mean([len(set(a) & set(b)) / len(a) for a, b in
                                             zip(own_g[:2000].tolist(), ref_graph.graph[:2000].tolist())])),
    }
    # reference query kernel on OUR graph (exchangeable blob)
    own.blob.cpu().numpy().tofile(os.path.join(workdir, "part_0.ggnn"))
    ref2 = run_ref(workdir, n=N, nq=Nq, d=D, measure=measure, kbuild=kbuild, build=0, kquery=K, tau_query=tau_q,
                   max_iter=max_it, query_reps=1, gpu_reps=0, bf=0, dump=1)
    r_ids = np.fromfile(os.path.join(workdir, "query_ids.bin"), np.int32).reshape(Nq, K)
    report["reference_on_our_graph"] = {"ids_exact_vs_ours": bool(np.array_equal(r_ids, ids2)),
                                        "recall": recall(ref_bf, r_ids, K), "kernel_ms": ref2.get("query_kernel_ms")}
    return report, dict(base=base, query=query, ref_graph=ref_graph, ref_ids=ref_ids, ref_d=ref_d, ref_bf=ref_bf,
                        ref_bfd=ref_bfd)


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "golden"
    out_root = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_root, exist_ok=True)
    print("device:", torch.cuda.get_device_name(0), "| nproc:", os.cpu_count(), flush=True)
    if mode == "golden":
        rep = {}
        for name, measure, kind, D in (("l2_10k", 0, "uniform", 128), ("cos_10k", 1, "normal", 96)):
            wd = os.path.join(out_root, "golden", name)
            r, data = side_by_side(wd, 10000, 2000, D, measure=measure, kind=kind)
            rep[name] = r
            gd = os.path.join(out_root, "golden_fixtures")
            os.makedirs(gd, exist_ok=True)
            np.savez_compressed(os.path.join(gd, f"{name}.npz"), graph_blob=data["ref_graph"].blob,
                                query_ids=data["ref_ids"], query_dists=data["ref_d"], bf_ids=data["ref_bf"],
                                bf_dists=data["ref_bfd"],
                                meta=np.array([10000, 2000, D, measure, 24, 10, 400], dtype=np.int64),
                                tau=np.array([0.5, 0.64], dtype=np.float32))
            print(json.dumps({name: r}, indent=1), flush=True)
        json.dump(rep, open(os.path.join(out_root, "golden_report.json"), "w"), indent=1)
    else:
        N, Nq, D = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
        kind = sys.argv[5] if len(sys.argv) > 5 else "uniform"
        measure = int(sys.argv[6]) if len(sys.argv) > 6 else 0
        wd = os.path.join(out_root, f"cmp_{N}_{D}_{kind}")
        r, _ = side_by_side(wd, N, Nq, D, measure=measure, kind=kind, oracle_check=True)
        print(json.dumps(r, indent=1), flush=True)
        json.dump(r, open(os.path.join(out_root, f"compare_{N}_{D}_{kind}.json"), "w"), indent=1)
        for f in ("base.bin", "query.bin", "part_0.ggnn"):  # keep gpurun_out small
            try:
                os.remove(os.path.join(wd, f))
            except OSError:
                pass


if __name__ == "__main__":
    main()
