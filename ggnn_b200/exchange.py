"""Fused shard-merge exchange over peer memory (NVLink / NVSwitch).

The reference merges the per-shard top-K lists on the CPU: every GPU sorts its [Nq, K*spg] buffer, copies it to the
host and a heap merge runs over the per-GPU lists (src/ggnn/base/gpu_instance.cu:714-790,
src/ggnn/base/result_merger.cpp:51-149).  Here the traversal kernel itself delivers the lists: its epilogue stores each
query's K (id, dist) pairs straight into the gathered buffer of every destination GPU (peer-mapped memory) and the last
warp of the launch bumps a flag word there (ggnn_b200_query_params.n_scatter ...).  The receiver waits for the flag on
its own stream (ggnn_b200_wait_flag, one polling thread) and merges the lists in place (ggnn_b200_merge_topk).  No
collective call, no packing kernel, no allocation per batch.

  PeerExchange   one process per GPU (torchrun): buffers are cudaMalloc'ed and shared through CUDA IPC handles
                 (exchanged once, through torch.distributed)
  NcclExchange   same interface on pre-allocated buffers with ncclAllGather (fallback when IPC is not permitted)
  LocalGather    one process driving several GPUs (GGNN.set_gpus([...])): peer access, destination = first GPU
"""
import ctypes as C

import torch

from . import _lib


class ScatterTarget:
    """the scatter fields of ggnn_b200_query_params for one batch (see include/ggnn_b200.h)"""
    __slots__ = ("n_dst", "slot0", "rows", "dists_offset", "dst_table", "flag_table", "done")

    def __init__(self, n_dst, slot0, rows, dists_offset, dst_table, flag_table, done):
        self.n_dst, self.slot0, self.rows, self.dists_offset = n_dst, slot0, rows, dists_offset
        self.dst_table, self.flag_table, self.done = dst_table, flag_table, done

    def apply(self, p, on_gpu_shard):
        p.d_query_results, p.d_query_results_dists = None, None
        p.shards_per_gpu, p.on_gpu_shard_id = 1, 0
        p.n_scatter, p.scatter_slot, p.scatter_rows = self.n_dst, self.slot0 + on_gpu_shard, self.rows
        p.scatter_dists_offset = self.dists_offset
        p.d_scatter_dst, p.d_scatter_flags, p.d_scatter_done = self.dst_table, self.flag_table, self.done


def _align(v, a):
    return (v + a - 1) // a * a


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _ExchangeBase:
    """rows x k lists from world*slots_per_rank shards; n_pipes independent pipelines (one per CUDA stream in flight),
    each double-buffered: a rank may start scattering batch b+1 of a pipeline while a slower rank still merges batch b"""

    def __init__(self, device, rows, k, slots_per_rank, n_pipes, group):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.device = device
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.rows, self.k, self.spr, self.n_pipes = int(rows), int(k), int(slots_per_rank), int(n_pipes)
        self.n_slots = self.world * self.spr
        self.list_words = self.rows * self.k
        self.seq = [0] * self.n_pipes
        self.out = [(torch.empty((self.rows, self.k), dtype=torch.int32, device=device),
                     torch.empty((self.rows, self.k), dtype=torch.float32, device=device)) for _ in range(2 * self.n_pipes)]

    def _buffer(self, pipe):
        return 2 * pipe + (self.seq[pipe] & 1)

    def _merge(self, b, ids_ptr, dists_ptr, n_query, id_offset_per_list, k_out=None):
        out_i, out_d = self.out[b]
        k_out = k_out or self.k
        _lib.check(_lib.lib().ggnn_b200_merge_topk(C.c_void_p(ids_ptr), C.c_void_p(dists_ptr), self.n_slots, self.list_words,
                                                   self.k, self.k, n_query, k_out, int(id_offset_per_list),
                                                   C.c_void_p(out_i.data_ptr()), C.c_void_p(out_d.data_ptr()),
                                                   _stream_ptr(self.device)))
        # (the kernel writes rows of k_out values densely)
        return (out_i.view(-1)[:n_query * k_out].view(n_query, k_out), out_d.view(-1)[:n_query * k_out].view(n_query, k_out))


class PeerExchange(_ExchangeBase):
    mode = "peer stores from the traversal kernel's epilogue (CUDA IPC over NVLink) + flag wait + in-place merge kernel"

    def __init__(self, device, rows, k, slots_per_rank=1, n_pipes=2, group=None, timeout_ms=20000):
        super().__init__(device, rows, k, slots_per_rank, n_pipes, group)
        l = _lib.lib()
        dist = self.dist
        self.timeout_ms = int(timeout_ms)
        n_buf = 2 * self.n_pipes
        self.dists_offset = self.n_slots * self.list_words * 4
        self.buf_bytes = _align(2 * self.dists_offset, 256)
        self.flags_off = n_buf * self.buf_bytes
        total = self.flags_off + n_buf * 128   # one flag word per 128-byte line
        self._own, self._peers, self._opened = None, [0] * self.world, []
        ok, err = 1, ""
        handle = C.create_string_buffer(64)
        try:
            with torch.cuda.device(device):
                p = C.c_void_p()
                _lib.check(l.ggnn_b200_ipc_alloc(total, C.byref(p), handle))
                self._own = p.value
        except Exception as e:  # noqa: BLE001
            ok, err = 0, repr(e)
        handles = [None] * self.world
        dist.all_gather_object(handles, (ok, bytes(handle.raw)), group=group)
        if all(h[0] for h in handles):
            try:
                with torch.cuda.device(device):
                    for r in range(self.world):
                        if r == self.rank:
                            self._peers[r] = self._own
                        else:
                            p = C.c_void_p()
                            _lib.check(l.ggnn_b200_ipc_open(handles[r][1], C.byref(p)))
                            self._peers[r] = p.value
                            self._opened.append(p.value)
            except Exception as e:  # noqa: BLE001
                ok, err = 0, repr(e)
        else:
            ok = 0
        flags = [None] * self.world
        dist.all_gather_object(flags, ok, group=group)
        if not all(flags):
            self._release()
            raise RuntimeError(f"peer-memory exchange unavailable ({err or 'a peer rank failed'})")
        tbl = [[[self._peers[r] + b * self.buf_bytes for r in range(self.world)],
                [self._peers[r] + self.flags_off + b * 128 for r in range(self.world)]] for b in range(n_buf)]
        self._tbl = torch.tensor(tbl, dtype=torch.int64, device=device)
        self._done = torch.zeros(32 * self.n_pipes, dtype=torch.int32, device=device)
        self._timed_out = torch.zeros(1, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group=group)

    def target(self, pipe=0):
        """-> (ScatterTarget for the query launches of the next batch of `pipe`, buffer index)"""
        b = self._buffer(pipe)
        t = ScatterTarget(self.world, self.rank * self.spr, self.rows, self.dists_offset, self._tbl[b, 0].data_ptr(),
                          self._tbl[b, 1].data_ptr(), self._done[32 * pipe:].data_ptr())
        return t, b

    def finish(self, pipe, b, n_query, id_offset_per_list, k_out=None):
        """wait until all world*spr lists of this batch have arrived, merge them -> (ids, dists) [n_query, k]
        (views of per-buffer outputs: valid until the same pipeline has been used twice more)"""
        uses = self.seq[pipe] // 2 + 1
        expected = (self.n_slots * uses) & 0xffffffff
        own = self._own + b * self.buf_bytes
        _lib.check(_lib.lib().ggnn_b200_wait_flag(C.c_void_p(self._own + self.flags_off + b * 128), expected, self.timeout_ms,
                                                  C.c_void_p(self._timed_out.data_ptr()), _stream_ptr(self.device)))
        self.seq[pipe] += 1
        return self._merge(b, own, own + self.dists_offset, n_query, id_offset_per_list, k_out)

    def check(self):
        """raises if a flag wait ever timed out (a peer did not deliver); synchronises"""
        if int(self._timed_out.item()):
            raise RuntimeError("shard-merge exchange: timed out waiting for a peer's lists")

    def _release(self):
        l = _lib.lib()
        for p in self._opened:
            l.ggnn_b200_ipc_close(C.c_void_p(p))
        self._opened = []
        if self._own:
            l.ggnn_b200_ipc_free(C.c_void_p(self._own))
            self._own = None

    def close(self):
        torch.cuda.synchronize(self.device)
        self.dist.barrier(group=self.group)
        with torch.cuda.device(self.device):
            for p in self._opened:
                _lib.lib().ggnn_b200_ipc_close(C.c_void_p(p))
            self._opened = []
        self.dist.barrier(group=self.group)
        with torch.cuda.device(self.device):
            self._release()


class NcclExchange(_ExchangeBase):
    mode = "traversal kernel writes the packed send buffer, ncclAllGather into pre-allocated buffers, in-place merge kernel"

    def __init__(self, device, rows, k, slots_per_rank=1, n_pipes=2, group=None, groups=None):
        super().__init__(device, rows, k, slots_per_rank, n_pipes, group)
        n_buf = 2 * self.n_pipes
        # collectives of one communicator serialise: one process group per pipeline
        self.groups = groups or [group] * self.n_pipes
        w = self.spr * self.list_words
        self._send = [torch.empty(2 * w, dtype=torch.int32, device=device) for _ in range(n_buf)]
        self._recv_i = [torch.empty(self.world * w, dtype=torch.int32, device=device) for _ in range(n_buf)]
        self._recv_d = [torch.empty(self.world * w, dtype=torch.int32, device=device) for _ in range(n_buf)]
        self._tbl = torch.tensor([[s.data_ptr()] for s in self._send], dtype=torch.int64, device=device)
        self.dists_offset = w * 4

    def target(self, pipe=0):
        b = self._buffer(pipe)
        return ScatterTarget(1, 0, self.rows, self.dists_offset, self._tbl[b].data_ptr(), None, None), b

    def finish(self, pipe, b, n_query, id_offset_per_list, k_out=None):
        w = self.spr * self.list_words
        g = self.groups[pipe]
        self.dist.all_gather_into_tensor(self._recv_i[b], self._send[b][:w], group=g)
        self.dist.all_gather_into_tensor(self._recv_d[b], self._send[b][w:], group=g)
        self.seq[pipe] += 1
        return self._merge(b, self._recv_i[b].data_ptr(), self._recv_d[b].data_ptr(), n_query, id_offset_per_list, k_out)

    def check(self):
        pass

    def close(self):
        torch.cuda.synchronize(self.device)


def make_exchange(device, rows, k, slots_per_rank=1, n_pipes=2, group=None, groups=None, prefer="peer"):
    """PeerExchange if CUDA IPC works on this box for every rank, else NcclExchange (all ranks take the same branch)"""
    if prefer == "peer":
        try:
            return PeerExchange(device, rows, k, slots_per_rank, n_pipes, group)
        except RuntimeError as e:
            import sys
            print(f"[ggnn_b200] {e}; using ncclAllGather", file=sys.stderr, flush=True)
    return NcclExchange(device, rows, k, slots_per_rank, n_pipes, group, groups)


class LocalGather:
    """one process, several GPUs (GGNN.set_gpus): every GPU's traversal kernel stores its lists into one buffer on the
    first GPU through peer access; that GPU merges them once the other devices' streams have got there (events, no host
    synchronisation).  Raises NotImplementedError when the devices cannot access each other (caller falls back to copies)."""

    def __init__(self, devices, rows, k, slots_per_gpu):
        l = _lib.lib()
        self.devices, self.rows, self.k, self.spg = list(devices), int(rows), int(k), int(slots_per_gpu)
        self.n_slots = len(self.devices) * self.spg
        self.list_words = self.rows * self.k
        d0 = self.devices[0]
        for d in self.devices[1:]:
            with torch.cuda.device(d):
                _lib.check(l.ggnn_b200_peer_enable(d0.index))
        self.dists_offset = self.n_slots * self.list_words * 4
        self.buf = torch.empty(2 * self.n_slots * self.list_words, dtype=torch.int32, device=d0)
        self.tbl = [torch.tensor([self.buf.data_ptr()], dtype=torch.int64, device=d) for d in self.devices]

    def target(self, gpu_index):
        return ScatterTarget(1, gpu_index * self.spg, self.rows, self.dists_offset, self.tbl[gpu_index].data_ptr(), None, None)

    def merge(self, n_query, id_offset_per_list):
        d0 = self.devices[0]
        with torch.cuda.device(d0):
            st0 = torch.cuda.current_stream(d0)
            for d in self.devices[1:]:
                st0.wait_event(torch.cuda.current_stream(d).record_event())
            out_i = torch.empty((n_query, self.k), dtype=torch.int32, device=d0)
            out_d = torch.empty((n_query, self.k), dtype=torch.float32, device=d0)
            _lib.check(_lib.lib().ggnn_b200_merge_topk(C.c_void_p(self.buf.data_ptr()),
                                                       C.c_void_p(self.buf.data_ptr() + self.dists_offset), self.n_slots,
                                                       self.list_words, self.k, self.k, n_query, self.k, int(id_offset_per_list),
                                                       C.c_void_p(out_i.data_ptr()), C.c_void_p(out_d.data_ptr()), _stream_ptr(d0)))
            # the buffer is reused by the next call: the other devices must not overwrite it before this merge has read it
            ev = st0.record_event()
            for d in self.devices[1:]:
                torch.cuda.current_stream(d).wait_event(ev)
        return out_i, out_d
