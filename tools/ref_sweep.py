"""Operating-point search on the REFERENCE first (procedure of examples/cpp-and-cuda/ggnn_benchmark.cpp:186-200):
the unmodified reference (oracle/_ref/ref_driver) builds its own graph on bench.py's synthetic data and answers one
10 000-query batch for every (tau_query, max_iterations) pair; recall@10 against its own brute force and its own
kernel time are recorded.  usage: python tools/ref_sweep.py kind [N] [D] [measure]  ->  gpurun_out/ref_sweep_<kind>_<N>.json"""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    kind = sys.argv[1]
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    D = int(sys.argv[3]) if len(sys.argv) > 3 else 128
    measure = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    Nq, K = 10_000, 10
    wd = f"/tmp/ref_sweep_{kind}"
    os.makedirs(wd, exist_ok=True)
    base, query = bench.gen_gpu(N, Nq, D, kind, 1234, torch.device("cuda", 0))
    base.cpu().numpy().tofile(os.path.join(wd, "base.bin"))
    query.cpu().numpy().tofile(os.path.join(wd, "query.bin"))
    del base, query
    torch.cuda.empty_cache()
    args = [os.path.join(ROOT, "oracle", "_ref", "ref_driver"), f"dir={wd}", f"n={N}", f"nq={Nq}", f"d={D}", f"measure={measure}",
            "kbuild=24", "tau_build=0.5", "refine=2", "build=1", f"kquery={K}", "tau_query=0.64", "max_iter=400", "query_reps=2",
            "gpu_reps=0", f"bf={K}", "dump=0", "sweep=0.34,0.41,0.51,0.64,0.8,1.0,1.25,1.5,2.0", "sweep_iters=200,400,1000,2000"]
    p = subprocess.run(args, capture_output=True, text=True)
    if p.returncode != 0:
        raise SystemExit(p.stderr[-2000:])
    r = json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1])
    out = {"kind": kind, "N": N, "D": D, "measure": measure, "reference_build_s": r.get("build_s"), "bf_s": r.get("bf_s"),
           "sweep": r.get("sweep")}
    ok = [s for s in out["sweep"] if s["recall"] >= 0.99]
    out["fastest_point_with_recall_ge_0.99"] = min(ok, key=lambda s: s["kernel_ms"]) if ok else None
    for f in os.listdir(wd):
        os.remove(os.path.join(wd, f))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"ref_sweep_{kind}_{N}.json"), "w"), indent=1)
    print(json.dumps(out["fastest_point_with_recall_ge_0.99"]))
    for s in out["sweep"]:
        print(s)


if __name__ == "__main__":
    main()
