"""Shard swapping GPU <-> pinned host RAM <-> disk, for bases whose shards do not all fit on their GPU.

Mirrors what the reference's GPUInstance does (src/ggnn/base/gpu_instance.cu:135-227 allocateGraph: as many GPU
buffers as the free memory minus the reserved amount allows, CPU buffers up to the CPU memory limit, the rest as
`part_<global_shard_id>.ggnn` files in the working directory; :370-467 swapOutPart / swapInPart; :246-255 prefetch
of the next shard) with one difference in mechanism: loads run on a side CUDA stream and are ordered against the
compute stream with events instead of IO threads + joins.  Only host-side data movement lives here; every kernel still
goes through the C ABI.  With enough memory (every BASELINE config on a 180 GB B200) nothing of this is used: the
shards stay resident (`plan_gpu_buffers` returns shards_per_gpu).
"""
import os
from collections import OrderedDict

import numpy as np
import torch


def plan_gpu_buffers(free_bytes, reserved_bytes, scratch_bytes, shard_bytes, shards_per_gpu, override=0):
    """How many (base, graph) shard buffers one GPU gets (gpu_instance.cu:150-175).  `override` > 0 forces the count
    (testing: env GGNN_B200_GPU_SHARD_BUFFERS)."""
    if override > 0:
        return min(int(override), shards_per_gpu)
    avail = int(free_bytes) - int(reserved_bytes) - int(scratch_bytes)
    n = avail // int(shard_bytes) if avail > 0 else 0
    if n < 1:
        raise RuntimeError("not enough GPU memory for a single shard (base + graph + build scratch); "
                           "use a smaller shard size")
    return int(min(n, shards_per_gpu))


def plan_cpu_buffers(cpu_limit_bytes, blob_bytes, n_shards):
    """How many graph blobs may stay in pinned host memory; the others live on disk (gpu_instance.cu:177-200)."""
    if cpu_limit_bytes is None:
        return n_shards
    return int(max(0, min(n_shards, int(cpu_limit_bytes) // int(blob_bytes))))


class ShardPool:
    """`n_buffers` device-resident (base rows, graph blob) slots for the `shards` of one GPU, least recently used
    replacement.  Shard state outside the GPU: base rows = a view of the user's host base tensor (never copied),
    graph blob = a pinned host tensor (the first `n_cpu` shards) or `part_<id>.ggnn` in `workdir`."""

    def __init__(self, device, n_buffers, rows, dim, blob_bytes, n_cpu, workdir, pin=True):
        self.device = torch.device(device)
        self.rows, self.dim, self.blob_bytes = rows, dim, blob_bytes
        self.workdir = workdir
        self.n_cpu = n_cpu
        self.pin = pin and self.device.type == "cuda"
        self.slots = [(torch.empty((rows, dim), dtype=torch.float32, device=self.device),
                       torch.zeros(blob_bytes, dtype=torch.uint8, device=self.device)) for _ in range(n_buffers)]
        self.resident = OrderedDict()     # global shard id -> slot index, least recently used first
        self.free = list(range(n_buffers))
        self.dirty = set()                # shards whose device blob is newer than the host / disk copy
        self.host_blob = {}               # shard id -> pinned host tensor
        self.on_disk = set()
        self.has_graph = set()            # shards that have a graph anywhere
        self.ready = {}                   # shard id -> event of its load; kept until the shard is evicted, so that EVERY
                                          # stream that acquires the shard waits for the load (not only the first one)
        self.query_calls = 0              # direction toggle of this GPU's shard loop (gpu_instance.cu:669-670)
        self.copy_stream = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None
        self.stats = {"loads": 0, "evictions": 0, "writebacks": 0, "disk_writes": 0, "disk_reads": 0}

    # ---- where a shard's graph lives outside the GPU ----
    def _path(self, gid):
        return os.path.join(self.workdir, f"part_{gid}.ggnn")

    def _writeback(self, gid, slot):
        """device blob -> pinned host tensor or disk (gpu_instance.cu:370-420)"""
        blob = self.slots[slot][1]
        self.stats["writebacks"] += 1
        if gid in self.host_blob or len(self.host_blob) < self.n_cpu:
            if gid not in self.host_blob:
                self.host_blob[gid] = torch.empty(self.blob_bytes, dtype=torch.uint8, pin_memory=self.pin)
            self.host_blob[gid].copy_(blob)          # synchronous: the slot is reused right after
        else:
            os.makedirs(self.workdir, exist_ok=True)
            blob.cpu().numpy().tofile(self._path(gid))
            self.on_disk.add(gid)
            self.stats["disk_writes"] += 1
        self.dirty.discard(gid)

    def _load(self, gid, slot, host_base_rows):
        base_buf, blob_buf = self.slots[slot]
        self.stats["loads"] += 1
        base_buf.copy_(host_base_rows, non_blocking=True)
        if gid in self.host_blob:
            blob_buf.copy_(self.host_blob[gid], non_blocking=True)
        elif gid in self.on_disk:
            blob_buf.copy_(torch.from_numpy(np.fromfile(self._path(gid), dtype=np.uint8)))
            self.stats["disk_reads"] += 1
        else:
            blob_buf.zero_()                          # no graph yet (before build)

    def _take_slot(self, keep=()):
        if self.free:
            return self.free.pop()
        for victim in self.resident:                  # least recently used first
            if victim not in keep:
                break
        else:
            raise RuntimeError("all GPU shard buffers are in use")
        slot = self.resident.pop(victim)
        self.stats["evictions"] += 1
        if self.device.type == "cuda":
            # Kernels of ANY stream may still read the victim (GGNN.query() pipelines chunks over several internal
            # streams, query_async keeps batches in flight) and its own load may still be running: wait for the whole
            # device before the slot is overwritten.  (Swapping is bound by the host->device copies, not by this.)
            torch.cuda.synchronize(self.device)
            self.ready.pop(victim, None)
        if victim in self.dirty:
            self._writeback(victim, slot)
        return slot

    # ---- API ----
    def prefetch(self, gid, host_base_rows, keep=()):
        """start loading a shard on the copy stream (no-op if resident or no slot can be freed without `keep`)"""
        if gid in self.resident:
            return
        if not self.free and all(r in keep for r in self.resident):
            return
        slot = self._take_slot(keep)
        if self.copy_stream is not None:
            with torch.cuda.stream(self.copy_stream):
                self._load(gid, slot, host_base_rows)
                self.ready[gid] = self.copy_stream.record_event()
        else:
            self._load(gid, slot, host_base_rows)
        self.resident[gid] = slot

    def acquire(self, gid, host_base_rows, keep=()):
        """-> (base [rows, dim] fp32, blob uint8) on the device, valid until the shard is evicted; the current stream
        waits for the load"""
        self.prefetch(gid, host_base_rows, keep)
        if gid not in self.resident:
            raise RuntimeError("no GPU shard buffer available")
        self.resident.move_to_end(gid)
        ev = self.ready.get(gid)
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)
        return self.slots[self.resident[gid]]

    def mark_built(self, gid):
        self.dirty.add(gid)
        self.has_graph.add(gid)

    def adopt_file(self, gid):
        """load(): the graph of this shard is the part file already in the working directory"""
        self.on_disk.add(gid)
        self.has_graph.add(gid)
        self.dirty.discard(gid)
        self.host_blob.pop(gid, None)
        if gid in self.resident:                      # drop a stale device copy
            if self.device.type == "cuda":
                torch.cuda.synchronize(self.device)
            self.ready.pop(gid, None)
            self.free.append(self.resident.pop(gid))

    def blob_to_file(self, gid, path):
        """store(): newest copy of the shard's graph -> file"""
        if gid in self.resident and (gid in self.dirty or (gid not in self.host_blob and gid not in self.on_disk)):
            if self.device.type == "cuda":
                torch.cuda.current_stream(self.device).synchronize()
            self.slots[self.resident[gid]][1].cpu().numpy().tofile(path)
        elif gid in self.host_blob:
            self.host_blob[gid].numpy().tofile(path)
        elif gid in self.on_disk:
            if os.path.abspath(path) != os.path.abspath(self._path(gid)):
                np.fromfile(self._path(gid), dtype=np.uint8).tofile(path)
        else:
            raise RuntimeError("There is no graph to store.")
