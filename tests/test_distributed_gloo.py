"""world_size-2 gloo test of the sharded search plumbing (rank->rows mapping, query broadcast or replication, all-gather of
the per-rank top-K lists, id re-basing).  The per-shard search and the merge are stood in for by the CPU
oracle here (tests only); on GPUs they are the CUDA kernels (tests/test_gpu_parity.py, bench.py --gpus N)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ggnn_b200 import distributed as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import pyoracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)
    N, Nq, Dm, K = 512, 24, 16, 5
    base = rng.random((N, Dm), dtype=np.float32)
    query_full = rng.random((Nq, Dm), dtype=np.float32)
    lo, hi = D.local_rows(rank, N, N // world, world)
    q = torch.from_numpy(query_full.copy()) if rank == 0 else torch.zeros(Nq, Dm)

    def local_query(qt):  # exact per-shard search (stand-in for the per-GPU kernel)
        i, d = O.bf_query(base[lo:hi], qt.numpy(), K)
        return torch.from_numpy(i), torch.from_numpy(d)

    def merge(all_i, all_d, off):
        i, d = O.merge_results(all_i.numpy(), all_d.numpy(), K, off)
        return torch.from_numpy(i), torch.from_numpy(d)

    ids, dists = D.distributed_query(local_query, merge, q, K, hi - lo)
    gi, gd = O.bf_query(base, query_full, K)
    ok = bool(np.array_equal(ids.numpy(), gi) and np.array_equal(dists.numpy(), gd) and torch.equal(q, torch.from_numpy(query_full)))
    # replicated queries (bench.py default at N > 1: every rank already holds the batch, no broadcast): same result,
    # and a rank whose copy differs is NOT overwritten (there is no hidden collective on the input side)
    q_rep = torch.from_numpy(query_full.copy())
    ids2, dists2 = D.distributed_query(local_query, merge, q_rep, K, hi - lo, broadcast=False)
    ok = ok and bool(np.array_equal(ids2.numpy(), gi) and np.array_equal(dists2.numpy(), gd))
    marker = torch.full((Nq, Dm), float(rank))
    D.distributed_query(lambda qt: local_query(torch.from_numpy(query_full.copy())), merge, marker, K, hi - lo, broadcast=False)
    ok = ok and bool((marker == float(rank)).all())
    open(os.path.join(out_dir, f"rank{rank}.txt"), "w").write("ok" if ok else "mismatch")
    dist.destroy_process_group()


def test_two_rank_sharded_search_equals_single_shard(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert open(os.path.join(tmp_path, f"rank{r}.txt")).read() == "ok"
