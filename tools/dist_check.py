"""Multi-GPU correctness check, run under torchrun (one rank per GPU):
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
(1) sharded brute force (NCCL all_gather + merge kernel) == single-GPU brute force over the whole base,
(2) sharded ANN search == k-way merge of the per-shard results (oracle merge semantics), recall vs (1),
(3) rank 0 additionally runs the in-process multi-GPU path GGNN.set_gpus([0..W-1]) and compares with (2)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import ggnn_b200 as ggnn  # noqa: E402
from ggnn_b200 import distributed as gd  # noqa: E402
from oracle import pyoracle as O  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    N, Nq, D, K = 200_000, 2_000, 128, 10
    base, query = bench.gen_gpu(N, Nq, D, "manifold8", 1234, dev, shard_index=rank)
    idx = ggnn.GGNN()
    idx.set_gpus([local])
    idx.set_return_results_on_gpu(True)
    idx.set_base(base)
    idx.build(24, 0.5, 2)
    q = query.clone() if rank == 0 else torch.zeros_like(query)
    ids, dists = gd.distributed_query(lambda t: idx.query(t, K, 0.64, 400), gd.gpu_merge, q, K, N)
    assert torch.equal(q, query), "broadcast"
    gt, gtd = gd.distributed_query(lambda t: idx.bf_query(t, K), gd.gpu_merge, q, K, N, broadcast=False)
    # gather everything on rank 0 for the checks
    bases = [torch.empty_like(base) for _ in range(world)]
    dist.all_gather(bases, base)
    loc_i, loc_d = idx.query(q, K, 0.64, 400)
    li = [torch.empty_like(loc_i) for _ in range(world)]
    ld = [torch.empty_like(loc_d) for _ in range(world)]
    dist.all_gather(li, loc_i)
    dist.all_gather(ld, loc_d)
    ok = True
    if rank == 0:
        whole = torch.cat(bases)
        one = ggnn.GGNN()
        one.set_gpus([local])
        one.set_return_results_on_gpu(True)
        one.set_base(whole)
        g1, g1d = one.bf_query(query, K)
        bf_ok = bool(torch.equal(g1d, gtd)) and float((g1 == gt).float().mean()) > 0.9999
        e_i, e_d = O.merge_results(torch.stack(li).cpu().numpy(), torch.stack(ld).cpu().numpy(), K, N)
        ann_ok = bool(np.array_equal(ids.cpu().numpy(), e_i) and np.array_equal(dists.cpu().numpy(), e_d))
        rec = bench.recall_at_k(gt, ids, K)
        print(f"[dist_check] world={world} sharded bf == single-GPU bf: {bf_ok}; sharded ANN == merged shards: {ann_ok}; "
              f"recall@10 {rec:.4f}", flush=True)
        ok = bf_ok and ann_ok and rec > 0.97
        # in-process multi-GPU path (peer copies + merge kernel), same shards
        try:
            multi = ggnn.GGNN()
            multi.set_gpus(list(range(world)))
            multi.set_shard_size(N)
            multi.set_base(whole.cpu())
            multi.build(24, 0.5, 2)
            mi, md = multi.query(query.cpu(), K, 0.64, 400)
            same_d = bool(torch.equal(md, dists.cpu()))
            print(f"[dist_check] in-process set_gpus({world}) distances == torchrun result: {same_d}; "
                  f"ids equal frac {float((mi == ids.cpu()).float().mean()):.5f}", flush=True)
            ok = ok and same_d
        except Exception as e:  # pragma: no cover
            print("[dist_check] in-process multi-GPU path failed:", repr(e), flush=True)
            ok = False
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
