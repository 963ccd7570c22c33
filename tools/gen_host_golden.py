"""Generates tests/golden/host_merge_eval.npz: outputs of the UNMODIFIED reference's host-side result handling --
ggnn::ResultMerger::merge (src/ggnn/base/result_merger.cpp:51-149) and ggnn::Evaluator (src/ggnn/base/eval.cpp:88-242) --
on seeded inputs, produced by oracle/_ref/ref_host_check (oracle/ref_host_check.cpp linked against the reference library
that oracle/build_ref.sh compiles from /root/reference).  Needs no GPU.  The inputs are regenerated from the seeds by
tests/test_oracle_golden.py through the functions below, so only the reference's OUTPUTS are stored.

    python tools/gen_host_golden.py        # in the container that has /root/reference
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_host_check")
OUT = os.path.join(ROOT, "tests", "golden", "host_merge_eval.npz")

# num_gpus, shards_per_gpu, N_query, KQuery, N_shard
MERGE_CASES = [(2, 1, 50, 10, 1000), (4, 2, 33, 7, 500), (8, 1, 20, 100, 12345), (3, 4, 17, 1, 77), (1, 3, 10, 5, 100), (1, 1, 9, 4, 50)]
# N, N_query, D, K_gt, KQuery, measure, is_uint8
EVAL_CASES = [(500, 40, 16, 20, 10, 0, 0), (500, 40, 16, 20, 10, 1, 0), (300, 25, 32, 100, 10, 0, 1), (300, 25, 32, 12, 10, 1, 1),
              (200, 30, 8, 5, 10, 0, 0), (400, 20, 24, 30, 1, 0, 0)]


def merge_inputs(case, seed):
    """per GPU: ids / dists [N_query, KQuery * spg], every row sorted by distance (what the per-GPU segmented sort of
    src/ggnn/base/gpu_instance.cu:745-790 leaves), all distances of a query distinct (the reference's tie order is arbitrary)"""
    num_gpus, spg, Nq, K, N_shard = case
    rng = np.random.default_rng(seed)
    per = K * spg
    d = np.empty((num_gpus, Nq, per), np.float32)
    for n in range(Nq):
        vals = (rng.permutation(num_gpus * per * 4)[:num_gpus * per].astype(np.float32) + 1) * 0.25
        d[:, n, :] = np.sort(vals.reshape(num_gpus, per), axis=1)
    ids = rng.integers(0, spg * N_shard, (num_gpus, Nq, per), dtype=np.int32)
    return ids, d


def eval_inputs(case, seed):
    """base with duplicate rows (distance ties in the ground truth), exact ground truth by the reference's own distance
    order, results = ground truth with some entries replaced / swapped"""
    N, Nq, D, Kgt, K, measure, is_u8 = case
    rng = np.random.default_rng(seed)
    if is_u8:
        base = rng.integers(0, 256, (N, D), dtype=np.uint8)
        query = rng.integers(0, 256, (Nq, D), dtype=np.uint8)
    else:
        base = rng.random((N, D), dtype=np.float32)
        query = rng.random((Nq, D), dtype=np.float32)
    dup = rng.integers(0, N, N // 4)
    base[rng.integers(0, N, N // 4)] = base[dup]   # duplicates
    query[0] = base[dup[0]]                        # distance 0, several times
    b, q = base.astype(np.float64), query.astype(np.float64)
    if measure == 0:
        dist = ((q[:, None, :] - b[None, :, :]) ** 2).sum(2)
    else:
        dist = 1.0 - (q @ b.T) / np.maximum(np.linalg.norm(q, axis=1)[:, None] * np.linalg.norm(b, axis=1)[None, :], 1e-30)
    gt = np.argsort(dist, axis=1, kind="stable")[:, :Kgt].astype(np.int32)
    res = np.empty((Nq, K), np.int32)
    for n in range(Nq):
        row = gt[n, :K].copy() if Kgt >= K else np.concatenate([gt[n], rng.integers(0, N, K - Kgt).astype(np.int32)])
        for k in range(K):
            u = rng.random()
            if u < 0.15:
                row[k] = rng.integers(0, N)                       # a wrong neighbour
            elif u < 0.3 and Kgt > K:
                row[k] = gt[n, rng.integers(K, Kgt)]              # one from beyond the K-th (a duplicate of it, maybe)
        if rng.random() < 0.5 and K > 1:
            row[[0, 1]] = row[[1, 0]]
        res[n] = row
    return base, query, gt, res


# vecs IO (src/ggnn/base/dataset.cu:118-226): type code of ref_host_check, numpy dtype, file extension, N, D, (from, num) of the range read
IO_CASES = [(0, np.float32, "fvecs", 7, 5, (2, 3)), (1, np.uint8, "bvecs", 9, 16, (0, 9)), (2, np.int32, "ivecs", 6, 3, (5, 1))]


def io_inputs(case, seed):
    _, dt, _, N, D, _ = case
    rng = np.random.default_rng(seed)
    return (rng.random((N, D), dtype=np.float32) if dt == np.float32 else rng.integers(0, 200, (N, D)).astype(dt))


def run_driver(mode, payload):
    with tempfile.TemporaryDirectory() as tmp:
        fin, fout = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
        with open(fin, "wb") as f:
            for part in payload:
                f.write(np.ascontiguousarray(part).tobytes())
        subprocess.run([DRIVER, mode, fin, fout], check=True)
        return np.fromfile(fout, dtype=np.uint8)


def main():
    if not os.path.exists(DRIVER):
        sys.exit(f"{DRIVER} is missing: bash oracle/build_ref.sh (needs /root/reference)")
    out = {"merge_cases": np.array(MERGE_CASES, np.int64), "eval_cases": np.array(EVAL_CASES, np.int64)}
    for i, case in enumerate(MERGE_CASES):
        ids, d = merge_inputs(case, 100 + i)
        num_gpus, spg, Nq, K, N_shard = case
        payload = [np.array(case, np.uint32)]
        for g in range(num_gpus):
            payload += [ids[g], d[g]]
        raw = run_driver("merge", payload)
        out[f"merge{i}_ids"] = raw[:Nq * K * 4].view(np.int32).reshape(Nq, K).copy()
        out[f"merge{i}_dists"] = raw[Nq * K * 4:].view(np.float32).reshape(Nq, K).copy()
    for i, case in enumerate(EVAL_CASES):
        base, query, gt, res = eval_inputs(case, 200 + i)
        raw = run_driver("eval", [np.array(case, np.uint32), base, query, gt, res])
        Nq = case[1]
        out[f"eval{i}_values"] = raw[:24].view(np.float32).copy()
        out[f"eval{i}_top1_end"] = raw[24:24 + 4 * Nq].view(np.uint32).copy()
        out[f"eval{i}_topk_end"] = raw[24 + 4 * Nq:].view(np.uint32).copy()
        print(case, out[f"eval{i}_values"], "dups:", int((out[f"eval{i}_top1_end"] > 1).sum()), int((out[f"eval{i}_topk_end"] > case[4]).sum()))
    # vecs files: written by the reference's Dataset<T>::store (committed next to the npz), ranges read back by its load
    for i, case in enumerate(IO_CASES):
        code, dt, ext, N, D, (lo, num) = case
        data = io_inputs(case, 300 + i)
        path = os.path.join(ROOT, "tests", "golden", f"ref_store.{ext}")
        with tempfile.TemporaryDirectory() as tmp:
            fin, fout = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
            with open(fin, "wb") as f:
                f.write(np.array([code, N, D], np.uint32).tobytes() + data.tobytes())
            subprocess.run([DRIVER, "store", fin, path], check=True)
            subprocess.run([DRIVER, "load", path, str(lo), str(num), fout], check=True)
            raw = np.fromfile(fout, dtype=np.uint8)
        n, d, esz = raw[:12].view(np.uint32)
        assert (n, d, esz) == (num, D, np.dtype(dt).itemsize), (n, d, esz)
        out[f"io{i}_loaded"] = raw[12:].view(dt).reshape(n, d).copy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
