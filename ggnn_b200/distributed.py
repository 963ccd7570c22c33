"""One-process-per-GPU sharded search (torch.distributed; NCCL over NVLink on GPUs, gloo in the CPU tests).

The reference shards the base by rows into equal contiguous shards, builds one independent graph per
shard, searches every shard with the same query batch and merges the per-shard top-K lists
(src/ggnn/base/ggnn.cu:154-203, 278-330; src/ggnn/base/result_merger.cpp:51-149).  It does all of that
in one process with one host thread per GPU, D2H copies and a CPU heap merge.  Here rank r owns GPU r
and global shards [r*spg, (r+1)*spg); the only exchange is the natural one: the query batch is
broadcast from rank 0 and the per-rank sorted [Nq, K] (id, dist) lists are all-gathered (one collective, Nq*K*8
bytes per rank) and merged by one kernel (ggnn_b200_merge_topk) -- no collective on the data path of the search.
"""
import torch
import torch.distributed as dist

from .exchange import make_exchange, PeerExchange, NcclExchange  # noqa: F401  (the device-side exchange)


def shard_layout(N, n_shard, world_size):
    """-> (num_shards, shards_per_gpu); same preconditions as the reference (ggnn.cu:162-183)."""
    n_shard = n_shard or N // world_size
    if n_shard <= 0 or N % n_shard:
        raise ValueError("base size must be divisible by the shard size")
    num_shards = N // n_shard
    if num_shards % world_size:
        raise ValueError("number of shards must be divisible by the number of GPUs")
    return num_shards, num_shards // world_size


def local_rows(rank, N, n_shard, world_size):
    """contiguous base row range [lo, hi) owned by `rank` (gpu_instance.cu:485-492)."""
    _, spg = shard_layout(N, n_shard, world_size)
    n_shard = n_shard or N // world_size
    return rank * spg * n_shard, (rank + 1) * spg * n_shard


def distributed_query(local_query_fn, merge_fn, query, k, rows_per_rank, group=None, src=0, broadcast=True):
    """
    local_query_fn(query) -> (ids [Nq,K] int32 in rank-local numbering, dists [Nq,K] fp32), sorted ascending
    merge_fn(all_ids [W,Nq,K], all_dists [W,Nq,K], id_offset_per_list) -> (ids [Nq,K], dists [Nq,K])
    Returns the merged global result on every rank.
    """
    world = dist.get_world_size(group)
    if broadcast and world > 1:
        dist.broadcast(query, src=src, group=group)
    ids, dists = local_query_fn(query)
    if world == 1:
        return ids, dists
    # ONE collective per batch: (ids, dists) travel packed as [2, Nq, K] 32-bit words per rank
    packed = torch.stack((ids.contiguous(), dists.contiguous().view(torch.int32)))
    gathered = torch.empty((world,) + tuple(packed.shape), dtype=torch.int32, device=ids.device)
    dist.all_gather(list(gathered.unbind(0)), packed, group=group)
    return merge_fn(gathered[:, 0], gathered[:, 1].view(torch.float32), rows_per_rank)


def exchange_query(idx, exchange, q_dev, k, tau_query, max_iterations=400, measure=0, pipe=0):
    """One batch of the sharded search through the fused exchange (ggnn_b200/exchange.py): the traversal kernels of this
    rank's shards store their lists into every rank's gathered buffer, then this rank waits for everybody's lists and
    merges them.  -> (ids [Nq, k] global numbering, dists) on this rank's GPU; every rank gets the full result.
    All ranks must call this in the same order with the same `pipe`."""
    tgt, b = exchange.target(pipe)
    idx._query_device(0, q_dev, int(k), tau_query, max_iterations, measure, scatter=tgt)
    return exchange.finish(pipe, b, q_dev.shape[0], idx._n_shard)


def gpu_merge(all_ids, all_dists, id_offset_per_list, k=None):
    """merge_fn backed by the CUDA kernel (no CPU path)."""
    import ctypes as C
    from . import _lib
    W, Nq, K_in = all_ids.shape
    k = k or K_in
    dev = all_ids.device
    if dev.type != "cuda":
        raise RuntimeError("gpu_merge needs CUDA tensors (there is no CPU fallback)")
    out_i = torch.empty((Nq, k), dtype=torch.int32, device=dev)
    out_d = torch.empty((Nq, k), dtype=torch.float32, device=dev)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    # the W lists may be strided views of one gathered buffer: rows of a list must be dense, lists any distance apart
    if all_ids.stride() != all_dists.stride() or tuple(all_ids.stride()[1:]) != (K_in, 1):
        all_ids, all_dists = all_ids.contiguous(), all_dists.contiguous()
    _lib.check(_lib.lib().ggnn_b200_merge_topk(C.c_void_p(all_ids.data_ptr()), C.c_void_p(all_dists.data_ptr()), W,
                                               all_ids.stride(0), K_in, K_in, Nq, k, int(id_offset_per_list),
                                               C.c_void_p(out_i.data_ptr()), C.c_void_p(out_d.data_ptr()), stream))
    return out_i, out_d
