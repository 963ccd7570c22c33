#!/usr/bin/env python
"""bench.py -- queries/s @ recall@10 of the GGNN batched query hot path on SIFT1M-shape synthetic fp32 data.

  python bench.py --gpus N --steps K --warmup W             our arm (ggnn_b200, sm_100a kernels)
  python bench.py --impl reference --gpus N --steps K ...   the UNMODIFIED reference CUDA library
                                                            (oracle/_ref, built from /root/reference) through
                                                            its own public API ggnn::GGNN, same data and parameters
N > 1: launched by torchrun, one rank per GPU; rank r owns one 1M-vector shard (the reference's row sharding), every
rank holds the query batch, the traversal kernel's epilogue stores the per-shard top-K lists straight into every rank's
gathered buffer over NVLink (CUDA IPC peer memory) and one kernel merges them (ggnn_b200/exchange.py).

A "step" = --batches-per-step (25) independent query batches of 10 000 queries each (BASELINE's batch), every batch one
search of the whole base; the batches of a step are DIFFERENT queries.  Rank 0 prints ONE JSON line.
  value    queries/s with the query batches resident in HBM and results left in HBM (CUDA events, max over ranks)
           -- for N > 1 multiplied by the number of shards searched per query (weak scaling: per-GPU work fixed)
  e2e      the same through the public API (GGNN.query_async / GGNN.query) with pinned HOST query tensors and the
           results copied back to the host inside the timed region
  roofline algorithmic bytes of the traversal / kernel time, against the measured HBM copy peak
  cpu_baseline  the CPU oracle port of the same traversal, all host cores, bounded query sample
  config4  BASELINE config 4 (100M x 128 = 8 shards of 12.5M over the N GPUs, plain queries/s: strong scaling)
  extra    N = 1 only: bf_query (config 5), config 3 (10M x 96 cosine build + query), warm build time, a harder
           operating point, the C++ host API (include/ggnn/ggnn.hpp) on the same graph
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DEF = dict(n_base=1_000_000, n_query=10_000, dim=128, k_build=24, tau_build=0.5, refine=2, k_query=10,
           tau_query=0.64, max_iterations=400, kind="manifold8", seed=1234)


def gen_gpu(N, Nq, D, kind, seed, device, shard_index=0, out=None):
    """synthetic SIFT1M-shape data generated on the device (fp32, no dataset files offline).
    'manifold<d>' (default manifold8: at the reference's documented SIFT1M operating point tau_query=0.64,
                  max_iterations=400 it reaches the documented recall@10 of 0.99, measured for both implementations): SIFT-like -- points on a d-dimensional linear manifold embedded in D dims
                  (low intrinsic dimension like real descriptors), per-dim std 40 around 128, unit Gaussian noise,
                  clipped to [0,255] and rounded: integer-valued vectors stored as fp32, like SIFT.
    'uniform'   : U[0,1) in every dim (the reference's README example; intrinsic dimension = D, a very hard ANN
                  instance: neither the reference nor this implementation gets useful recall at 1M points).
    'clustered' : mixture of 1000 isotropic Gaussians (intrinsic dimension = D, equally hard).
    Queries are drawn from the same distribution with a separate generator (the first Nq' < Nq queries of a larger
    request are the queries of the smaller one); shards differ by `shard_index`.  `out`: [N, D] tensor to fill with
    the base rows (a slice of a larger base) instead of allocating."""
    g = torch.Generator(device=device).manual_seed(seed)
    gb = torch.Generator(device=device).manual_seed(seed + 17 * (shard_index + 1))

    def alloc(n, is_base):
        return out if (is_base and out is not None) else torch.empty((n, D), device=device)
    QB = 10_000  # queries are drawn in blocks of 10 000, so the first blocks of a larger request equal a smaller request
    if kind == "uniform":
        q = torch.cat([torch.rand((min(QB, Nq - lo), D), generator=g, device=device) for lo in range(0, Nq, QB)])
        b = torch.rand((N, D), generator=gb, device=device)
        if out is not None:
            out.copy_(b)
            b = out
        return b, q
    if kind.startswith("manifoldcos"):  # DEEP-like: unit-normalised vectors with low intrinsic dimension (cosine)
        d = int(kind[len("manifoldcos"):] or 8)
        A = torch.randn((d, D), generator=g, device=device) / (d ** 0.5)

        def draw_c(n, gen, is_base=False, chunk=1 << 20):
            o = alloc(n, is_base)
            for lo in range(0, n, chunk):
                m = min(chunk, n - lo)
                x = torch.randn((m, d), generator=gen, device=device) @ A + 0.02 * torch.randn((m, D), generator=gen, device=device)
                o[lo:lo + m] = x / x.norm(dim=1, keepdim=True)
            return o
        query = draw_c(Nq, g, chunk=QB)
        return draw_c(N, gb, True), query
    if kind.startswith("manifold"):
        d = int(kind[len("manifold"):] or 16)
        A = torch.randn((d, D), generator=g, device=device) / (d ** 0.5)

        def draw(n, gen, is_base=False, chunk=1 << 20):
            o = alloc(n, is_base)
            for lo in range(0, n, chunk):  # chunked to bound temporaries
                m = min(chunk, n - lo)
                z = torch.randn((m, d), generator=gen, device=device)
                x = (z @ A) * 40.0 + 128.0 + torch.randn((m, D), generator=gen, device=device)
                o[lo:lo + m] = x.round_().clamp_(0, 255)
            return o
        query = draw(Nq, g, chunk=QB)
        return draw(N, gb, True), query
    nc = 1000
    centers = torch.rand((nc, D), generator=g, device=device) * 160 + 20

    def draw(n, gen, is_base=False, chunk=1 << 20):
        o = alloc(n, is_base)
        for lo in range(0, n, chunk):
            m = min(chunk, n - lo)
            c = torch.randint(0, nc, (m,), generator=gen, device=device)
            x = centers[c] + torch.randn((m, D), generator=gen, device=device) * 25
            o[lo:lo + m] = x.round_().clamp_(0, 255)
        return o
    query = draw(Nq, g, chunk=QB)
    return draw(N, gb, True), query


class ClockSampler:
    """samples SM clock and clock-event (throttle) reasons of one GPU DURING the timed region: NVML polled every
    ~2 ms from a thread (nvidia-smi -lms is too coarse for a millisecond-scale region)."""

    def __init__(self, gpu):
        self.gpu, self.sm, self.reasons, self.stop_flag, self.t, self.err = gpu, [], set(), False, None, None
        self.max_mhz = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def _poll(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(get_reasons(self.h))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception as e:  # pragma: no cover
                self.err = repr(e)
                break
            time.sleep(0.002)

    def stop(self):
        self.stop_flag = True
        if self.t:
            self.t.join(timeout=1.0)
        out = {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.sm)}
        if self.err:
            out["error"] = self.err
        return out


def measured_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def recall_at_k(gt, ids, K):
    gt, ids = gt[:, :K].long(), ids[:, :K].long()
    return float((ids.unsqueeze(2) == gt.unsqueeze(1)).any(2).float().mean())


# ------------------------------------------------------------------------------------------------
def _events(n):
    return [torch.cuda.Event(enable_timing=True) for _ in range(n)]


def _query_stats(idx, query, K, tau, max_it, measure=0):
    """pops / distance evaluations of one batch on shard 0 (counters emitted by the traversal kernel) and its
    algorithmic bytes: 4*D*(1 + n_dist) + 4*KBuild*n_pop + 4*S + 8*K per query"""
    from ggnn_b200 import _lib
    dev = query.device
    gr = idx.get_graph(0)
    cfg = gr.config
    nq = query.shape[0]
    stats = torch.zeros((nq, 2), dtype=torch.int32, device=dev)
    tmp_i = torch.empty((nq, K), dtype=torch.int32, device=dev)
    tmp_d = torch.empty((nq, K), dtype=torch.float32, device=dev)
    p = _lib.QueryParams()
    p.D, p.measure, p.KQuery, p.tau_query, p.max_iterations = cfg.D, measure, K, tau, max_it
    p.N_base, p.KBuild, p.num_starting_points = cfg.N, cfg.KBuild, cfg.S
    p.d_base, p.d_query, p.d_graph = idx._shards[0].base.data_ptr(), query.data_ptr(), gr.graph.data_ptr()
    p.d_starting_points, p.d_nn1_stats = gr.layer_translation(3).data_ptr(), gr.nn1_stats.data_ptr()
    p.d_query_results, p.d_query_results_dists, p.d_stats = tmp_i.data_ptr(), tmp_d.data_ptr(), stats.data_ptr()
    p.shards_per_gpu, p.on_gpu_shard_id = 1, 0
    _lib.check(_lib.lib().ggnn_b200_query(C.byref(p), nq, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    n_iter, n_dist = stats.to(torch.float64).sum(0).tolist()
    # (n_dist already includes the S start-point evaluations)
    alg = 4.0 * cfg.D * (nq + n_dist) + 4.0 * cfg.KBuild * n_iter + (4.0 * cfg.S + 8.0 * K) * nq
    return n_iter, n_dist, alg


def _crc(t):
    return zlib.crc32(t.cpu().numpy().tobytes()) & 0xffffffff


def run_ours(a):
    import torch.distributed as dist
    import ggnn_b200 as ggnn
    from ggnn_b200 import distributed as gd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torchrun (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B, Nq, K = max(1, a.batches_per_step), a.n_query, a.k_query
    base, q_all = gen_gpu(a.n_base, Nq * B, a.dim, a.kind, a.seed, dev, shard_index=rank)
    batches = [q_all[b * Nq:(b + 1) * Nq] for b in range(B)]
    query = batches[0]
    idx = ggnn.GGNN()
    idx.set_gpus([local])
    idx.set_return_results_on_gpu(True)
    idx.set_base(base)
    torch.cuda.synchronize()
    t0 = time.time()
    idx.build(a.k_build, a.tau_build, a.refine)
    torch.cuda.synchronize()
    build_s = time.time() - t0
    build_warm_s, build_passes = None, None
    if world == 1:  # the first build of a process pays module load / context set-up: time a second one (fresh object)
        from ggnn_b200 import _lib
        tmp = ggnn.GGNN()
        tmp.set_gpus([local])
        tmp.set_base(base)
        torch.cuda.synchronize()
        _lib.check(_lib.lib().ggnn_b200_build_stats_begin())   # per-launch counters + event times of merge / sym
        t0 = time.time()
        tmp.build(a.k_build, a.tau_build, a.refine)
        torch.cuda.synchronize()
        build_warm_s = time.time() - t0
        ps = (_lib.BuildPassStats * 64)()
        n_ps = C.c_uint32(0)
        _lib.check(_lib.lib().ggnn_b200_build_stats_end(ps, 64, C.byref(n_ps)))
        build_passes = []
        peak_b, _ = measured_peaks()
        for i in range(n_ps.value):
            s_ = ps[i]
            if s_.layer_btm != 0:
                continue   # (the upper layers are 3 % of the build)
            # algorithmic bytes: 4*D per point read + 4*D per distance evaluation (sym: one row read serves
            # both of its distances) + 4*KBuild per pop
            alg = 4.0 * a.dim * (s_.points + s_.dists) + 4.0 * a.k_build * s_.pops
            build_passes.append({"kernel": "sym" if s_.kernel else "merge", "layer_top": s_.layer_top, "ms": s_.ms,
                                 "pops_per_point": s_.pops / s_.points, "dists_per_point": s_.dists / s_.points,
                                 "algorithmic_gbs": alg / (s_.ms * 1e-3) / 1e9, "frac_of_hbm_peak": alg / (s_.ms * 1e-3) / 1e9 / peak_b})
        del tmp

    def local_query(q):
        return idx.query(q, K, a.tau_query, a.max_iterations)

    # N > 1: one exchange pipeline per CUDA stream in flight (each double-buffered); the NCCL fallback additionally
    # gets one communicator per pipeline (collectives of one communicator serialise)
    n_pipes = max(1, a.streams)
    exchange = None
    if world > 1:
        chk = torch.stack((q_all.double().sum(), q_all[::97].double().sum()))  # the seeded generator must agree everywhere
        ref_chk = chk.clone()
        dist.broadcast(ref_chk, src=0)
        assert torch.equal(chk, ref_chk), "query batches differ between ranks"
        groups = [dist.new_group(backend="nccl") for _ in range(n_pipes)] if a.exchange == "nccl" else None
        exchange = gd.make_exchange(dev, Nq, K, 1, n_pipes, None, groups, prefer=a.exchange)

    def step_device(q, pipe=0):
        if world == 1:
            return local_query(q)
        return gd.exchange_query(idx, exchange, q, K, a.tau_query, a.max_iterations, 0, pipe)

    # ground truth + recall (untimed): exact brute force on every shard, merged the same way
    def bf_local(q):
        return idx.bf_query(q, K)
    if world == 1:
        gt, _ = bf_local(query)
    else:
        gt, _ = gd.distributed_query(bf_local, gd.gpu_merge, query, K, a.n_base, broadcast=False)
    ids, dists = step_device(query)
    ids, dists = ids.clone(), dists.clone()
    rec = recall_at_k(gt, ids, K)
    if world > 1:  # the fused exchange must give exactly what the plain all_gather + merge path gives
        ids_ag, d_ag = gd.distributed_query(local_query, gd.gpu_merge, query, K, a.n_base, broadcast=False)
        assert torch.equal(ids, ids_ag) and torch.equal(dists, d_ag), "fused exchange and all_gather results differ"

    # algorithmic bytes per launch on this rank: mean over the batches of a step (untimed stats launches)
    st = [_query_stats(idx, q, K, a.tau_query, a.max_iterations) for q in batches[:min(B, 5)]]
    n_iter, n_dist, alg_bytes = [float(np.mean([s[i] for s in st])) for i in range(3)]

    # ---- timed region 1: device-resident ----
    # A step = B independent batches.  With --streams 2 (default) consecutive batches are enqueued on two alternating
    # CUDA streams, so the next batch's CTAs fill the SMs that the previous batch's last, long queries leave idle
    # (a single launch's tail is one query latency).  --streams 1 serialises the batches.
    sampler = ClockSampler(local)
    n_streams = n_pipes
    streams = [torch.cuda.Stream(dev) for _ in range(n_streams)] if n_streams > 1 else [torch.cuda.current_stream(dev)]

    def run_steps(n):
        i = 0
        for _ in range(n):
            for q in batches:
                with torch.cuda.stream(streams[i % n_streams]):
                    step_device(q, i % n_streams)
                i += 1

    run_steps(a.warmup)
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e_start, e_end = _events(2)
    cur = torch.cuda.current_stream(dev)
    e_start.record(cur)
    for s_ in streams:
        s_.wait_event(e_start)
    run_steps(a.steps)
    for s_ in streams:
        cur.wait_stream(s_)
    e_end.record(cur)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = e_start.elapsed_time(e_end)
    clocks = sampler.stop() if rank == 0 else None
    if exchange is not None:
        exchange.check()

    # dominant kernel alone (the traversal kernel of this rank, one launch at a time), CUDA events on the launch stream
    n_k = min(a.steps * B, 100)
    kev = _events(2 * n_k)
    for s in range(n_k):
        kev[2 * s].record()
        local_query(batches[s % B])
        kev[2 * s + 1].record()
    torch.cuda.synchronize()
    kernel_ms = float(np.mean([kev[2 * s].elapsed_time(kev[2 * s + 1]) for s in range(n_k)]))

    # ---- timed region 2: end to end through the public API, host buffers ----
    q_host = [q.cpu().pin_memory() for q in batches]
    depth = max(1, a.e2e_depth)
    if world == 1:
        idx.set_return_results_on_gpu(False)

        def e2e_sync(n):   # one synchronous GGNN.query() per batch: H2D + kernel + D2H inside, the reference's only mode
            r = None
            for _ in range(n):
                for qh in q_host:
                    r = idx.query(qh, K, a.tau_query, a.max_iterations)
            return r

        def e2e_async(n):  # up to `depth` batches in flight through GGNN.query_async(); every batch still copies its own
            pending, last = [], None   # queries host->device and its own results device->host inside the timed region
            for _ in range(n):
                for qh in q_host:
                    pending.append(idx.query_async(qh, K, a.tau_query, a.max_iterations))
                    if len(pending) >= depth:
                        last = pending.pop(0).result()
            while pending:
                last = pending.pop(0).result()
            return last
    else:
        qd = [torch.empty_like(query) for _ in range(n_streams)]
        pin_i = [torch.empty((Nq, K), dtype=torch.int32, pin_memory=True) for _ in range(n_streams)]
        pin_d = [torch.empty((Nq, K), dtype=torch.float32, pin_memory=True) for _ in range(n_streams)]
        evs = [None] * n_streams

        def e2e_pipelined(n, n_pipe):
            # every rank copies the batch host->device over its own PCIe link (what the reference does: one H2D per GPU,
            # gpu_instance.cu:638-641), searches its shard, the lists are exchanged and merged, rank 0 copies the result back
            i = 0
            for _ in range(n):
                for qh in q_host:
                    pipe = i % n_pipe
                    i += 1
                    if evs[pipe] is not None:
                        evs[pipe].synchronize()  # the batch that used this pipeline's buffers has delivered its result
                    with torch.cuda.stream(streams[pipe]):
                        qd[pipe].copy_(qh, non_blocking=True)
                        r = step_device(qd[pipe], pipe)
                        if rank == 0:
                            pin_i[pipe].copy_(r[0], non_blocking=True)
                            pin_d[pipe].copy_(r[1], non_blocking=True)
                        evs[pipe] = streams[pipe].record_event()
            for e in evs:
                if e is not None:
                    e.synchronize()
            return pin_i[(i - 1) % n_pipe], pin_d[(i - 1) % n_pipe]

        def e2e_sync(n):
            return e2e_pipelined(n, 1)

        def e2e_async(n):
            return e2e_pipelined(n, n_streams)

    def timed_host(fn, n):
        fn(max(1, min(a.warmup, 3)))   # (the first calls allocate the pinned result buffers)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn(n)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3, r
    e2e_sync_ms, r_sync = timed_host(e2e_sync, a.steps)
    r_sync = tuple(t.clone() for t in r_sync)   # (N > 1: the pinned result buffers are reused by the next run)
    e2e_ms, e2e_depth = e2e_sync_ms, 1
    if depth > 1 and (world == 1 or n_streams > 1):
        e2e_ms, r_async = timed_host(e2e_async, a.steps)
        e2e_depth = depth if world == 1 else n_streams
        if rank == 0:
            assert torch.equal(r_async[0], r_sync[0]) and torch.equal(r_async[1], r_sync[1]), "async and sync results differ"
    if exchange is not None:
        exchange.check()
    if world == 1:
        idx.set_return_results_on_gpu(True)

    times = torch.tensor([total_ms, e2e_ms, kernel_ms, e2e_sync_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, kernel_ms, e2e_sync_ms = times.tolist()

    config4 = None
    if a.config4:
        del q_host
        try:
            config4 = run_config4(a, dev, world, rank, local)
        except Exception as e:  # noqa: BLE001  (the headline line must still be printed)
            config4 = {"error": repr(e)[-400:]}
            if world > 1:
                raise

    if rank != 0:
        if exchange is not None:
            exchange.close()
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = total_ms / a.steps
    shards = world
    q_per_step = Nq * B
    qps = q_per_step / (ms_per_step * 1e-3)
    e2e_qps = q_per_step / (e2e_ms / a.steps * 1e-3)
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    gr = idx.get_graph(0)
    cpu = cpu_baseline(a, idx, base, query, gr)
    extra = None
    if world == 1 and a.extras:
        extra = run_extras(a, idx, base, query, gt, dev, build_s, build_warm_s, build_passes)
    out = {
        "metric": "queries/sec @ recall@10", "value": qps * shards, "unit": "queries/s" if shards == 1 else "queries/s x shards searched",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"SIFT1M-shape {a.n_base}x{a.dim} fp32 ({a.kind}) per GPU shard, {Nq} queries per batch, "
                               f"Euclidean, k_build={a.k_build} tau_build={a.tau_build} refine={a.refine}, "
                               f"k_query={K} tau_query={a.tau_query} max_iterations={a.max_iterations}",
                   "batches_per_step": B, "queries_per_step": q_per_step,
                   "shards": shards, "vectors_total": a.n_base * shards, "queries_per_s": qps, "recall_at_10": rec,
                   "l2_policy": "inputs_larger_than_l2 (512 MB base per shard, gather-random; every batch of a step holds different queries)",
                   "parallelism": (f"base row-sharded x{shards}, queries replicated, " + exchange.mode if shards > 1 else "single shard"),
                   "build_s": build_s, "build_warm_s": build_warm_s,
                   "pipelining": f"{n_streams} CUDA stream(s): independent query batches overlap their tails" if n_streams > 1
                   else "none (batches serialised on one stream)",
                   "single_batch_ms": kernel_ms, "ms_per_batch": ms_per_step / B},
        "recall_at_10": rec,
        "e2e": {"value": e2e_qps * shards, "unit": "queries/s" if shards == 1 else "queries/s x shards searched",
                "h2d_bytes_per_step": q_per_step * a.dim * 4 * shards,
                "d2h_bytes_per_step": q_per_step * K * 8,
                "mode": ((f"GGNN.query_async(), {e2e_depth} batches in flight" if shards == 1 else
                          f"{e2e_depth} batches in flight: pinned H2D on every rank, per-shard query, exchange, merge, D2H on rank 0")
                         if e2e_depth > 1 else "one synchronous call per batch"),
                "sync_value": q_per_step / (e2e_sync_ms / a.steps * 1e-3) * shards,
                "sync_mode": "one synchronous GGNN.query() per batch (the reference's only mode)"},
        "gpu_launches": a.steps * B * (1 + (2 if shards > 1 else 0)),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": load_traffic(), "kernel": "query_kernel", "kernel_ms": kernel_ms,
                     "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                     "pops_per_query": n_iter / Nq, "dists_per_query": n_dist / Nq,
                     "note": "kernel_ms = one launch alone (its tail included); achieved at the pipelined rate = "
                             f"{alg_bytes / (ms_per_step / B * 1e-3) / 1e9:.0f} GB/s (a third of the gathers hit in L2)"},
        "cpu_baseline": cpu,
    }
    if config4 is not None:
        out["config4"] = config4
    if extra is not None:
        out["extra"] = extra
    emit(out)
    if exchange is not None:
        exchange.close()
    if world > 1:
        dist.destroy_process_group()


def run_config4(a, dev, world, rank, local):
    """BASELINE config 4: 100M x 128 fp32 = 8 shards of 12.5M vectors over the N GPUs (8/N shards per GPU, the
    reference's partitioning: src/ggnn/base/ggnn.cu:154-203, gpu_instance.cu:485-492; usage
    examples/cpp-and-cuda/ggnn_main_multi_gpu.cpp:62-65), generated on the device per shard, k_query 10.  Plain
    queries/s: the total work is fixed, so the ratio between the lines of different N is the strong-scaling speed-up."""
    import torch.distributed as dist
    import ggnn_b200 as ggnn
    from ggnn_b200 import distributed as gd
    n_shards, n_shard = a.c4_shards, a.c4_shard_size
    if n_shards % world:
        return {"skipped": f"{n_shards} shards do not divide over {world} GPUs"}
    spg = n_shards // world
    Nq, K, D = a.n_query, a.k_query, a.dim
    Bc = max(1, a.c4_batches)
    torch.cuda.empty_cache()
    base = torch.empty((spg * n_shard, D), dtype=torch.float32, device=dev)
    q_all = None
    for s in range(spg):
        _, q_all = gen_gpu(n_shard, Nq * Bc, D, a.kind, a.seed, dev, shard_index=rank * spg + s, out=base[s * n_shard:(s + 1) * n_shard])
    batches = [q_all[b * Nq:(b + 1) * Nq] for b in range(Bc)]
    idx = ggnn.GGNN()
    idx.set_gpus([local])
    idx.set_shard_size(n_shard)
    idx.set_return_results_on_gpu(True)
    idx.set_base(base)
    torch.cuda.synchronize()
    t0 = time.time()
    idx.build(a.k_build, a.tau_build, a.refine)
    torch.cuda.synchronize()
    build_s = time.time() - t0
    n_pipes = max(1, a.streams)
    exchange = None
    if world > 1:
        groups = [dist.new_group(backend="nccl") for _ in range(n_pipes)] if a.exchange == "nccl" else None
        exchange = gd.make_exchange(dev, Nq, K, spg, n_pipes, None, groups, prefer=a.exchange)

    def step(q, pipe=0):
        if world == 1:
            return idx.query(q, K, a.tau_query, a.max_iterations)
        return gd.exchange_query(idx, exchange, q, K, a.tau_query, a.max_iterations, 0, pipe)

    # ground truth: exact brute force per shard (bounded scratch), merged like the search results
    def bf_local(q):
        parts = [ggnn.GGNN._bf_query_rows(base[s * n_shard:(s + 1) * n_shard], q, K) for s in range(spg)]
        if spg == 1:
            return parts[0]
        return gd.gpu_merge(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]), n_shard)
    nq_gt = min(Nq, a.c4_gt_queries)
    q_gt = batches[0][:nq_gt].contiguous()
    if world == 1:
        gt, _ = bf_local(q_gt)
    else:
        gt, _ = gd.distributed_query(bf_local, gd.gpu_merge, q_gt, K, spg * n_shard, broadcast=False)
    ids, _ = step(batches[0])
    rec = recall_at_k(gt, ids[:nq_gt], K)
    torch.cuda.synchronize()

    streams = [torch.cuda.Stream(dev) for _ in range(n_pipes)] if n_pipes > 1 else [torch.cuda.current_stream(dev)]

    def run(n):
        i = 0
        for _ in range(n):
            for q in batches:
                with torch.cuda.stream(streams[i % n_pipes]):
                    step(q, i % n_pipes)
                i += 1
    run(2)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = _events(2)
    cur = torch.cuda.current_stream(dev)
    e0.record(cur)
    for s_ in streams:
        s_.wait_event(e0)
    run(a.c4_steps)
    for s_ in streams:
        cur.wait_stream(s_)
    e1.record(cur)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    # one shard's traversal kernel alone + its algorithmic bytes (base >> L2: no L2 hits to speak of)
    n_iter, n_dist, alg = _query_stats(idx, batches[0], K, a.tau_query, a.max_iterations)
    gr = idx.get_graph(0)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        exchange.check()
        exchange.close()
    ms = float(t.item())
    ms_batch = ms / (a.c4_steps * Bc)
    peak, _ = measured_peaks()
    out = {"workload": f"{n_shards} shards x {n_shard} x {D} fp32 ({a.kind}) = {n_shards * n_shard} vectors, {spg} shard(s) per GPU, "
                       f"{Nq} queries per batch, k_query={K} tau_query={a.tau_query} max_iterations={a.max_iterations}",
           "vectors_total": n_shards * n_shard, "shards": n_shards, "shards_per_gpu": spg, "n_gpus": world,
           "queries_per_s": Nq / (ms_batch * 1e-3), "ms_per_batch": ms_batch, "scaling": "strong",
           "batches_timed": a.c4_steps * Bc, "recall_at_10": rec, "recall_queries": nq_gt, "build_s": build_s,
           "build_s_per_shard": build_s / spg,
           "shard0_pops_per_query": n_iter / Nq, "shard0_dists_per_query": n_dist / Nq,
           "shard0_algorithmic_bytes_per_launch": alg,
           "algorithmic_gbs_per_gpu": alg * spg / (ms_batch * 1e-3) / 1e9, "hbm_peak_gbs": peak,
           "graph_blob_bytes_per_shard": int(gr.blob.numel())}
    del idx, base, exchange
    torch.cuda.empty_cache()
    return out


def run_extras(a, idx, base, query, gt, dev, build_s, build_warm_s, build_passes=None):
    """N = 1 only, untimed by the headline: the other BASELINE configs on the same box"""
    import ggnn_b200 as ggnn
    K, Nq = a.k_query, a.n_query
    ex = {"build": {"first_in_process_s": build_s, "warm_s": build_warm_s,
                    "what": f"{a.n_base}x{a.dim} k_build={a.k_build} tau_build={a.tau_build} refine={a.refine}",
                    "layer0_passes": build_passes}}

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        ev = _events(2 * reps)
        for r in range(reps):
            ev[2 * r].record()
            out = fn()
            ev[2 * r + 1].record()
        torch.cuda.synchronize()
        return out, float(np.median([ev[2 * r].elapsed_time(ev[2 * r + 1]) for r in range(reps)]))
    # config 1: the README example -- 10 000 x 128 uniform base and 10 000 queries as CPU-resident tensors, everything
    # (host -> device copies, results back to the host) inside the timed call
    try:
        bc, qc = config1_data()
        c1 = ggnn.GGNN()
        c1.set_base(bc)
        t0 = time.perf_counter()
        c1.build(a.k_build, a.tau_build, a.refine)
        torch.cuda.synchronize()
        c1_build = time.perf_counter() - t0
        c1.query(qc, K, a.tau_query, a.max_iterations)
        ts = []
        for _ in range(10):
            t0 = time.perf_counter()
            ci, _ = c1.query(qc, K, a.tau_query, a.max_iterations)
            ts.append((time.perf_counter() - t0) * 1e3)
        cg, _ = c1.bf_query(qc, K)
        ex["config1"] = {"workload": "README example: 10000x128 U[0,1) fp32 base + 10000 queries as CPU tensors (torch.manual_seed(1234)), "
                                     f"k_build={a.k_build} tau_build={a.tau_build}, k_query={K} tau_query={a.tau_query} max_iterations={a.max_iterations}",
                         "build_s": c1_build, "query_ms_host_to_host": float(np.median(ts)), "queries_per_s": 10_000 / (float(np.median(ts)) * 1e-3),
                         "recall_at_10": recall_at_k(cg, ci, K), "bf_ids_crc32": _crc(cg)}
        del c1
    except Exception as e:  # noqa: BLE001
        ex["config1"] = {"error": repr(e)[-300:]}
    # config 5: bf_query 1M x 10k as a tensor-core contraction (+ exact re-rank), k = 10 and the API default k = 100
    try:
        (bi, bd), ms10 = timed(lambda: idx.bf_query(query, K))
        (bi100, _), ms100 = timed(lambda: idx.bf_query(query, 100), reps=3)
        flops = 2.0 * Nq * a.n_base * a.dim
        ex["bf_query"] = {"workload": f"{a.n_base}x{a.dim} base, {Nq} queries (BASELINE config 5)",
                          "k10_ms": ms10, "k100_ms": ms100, "k10_useful_tflops": flops / (ms10 * 1e-3) / 1e12,
                          "ids_crc32_k10": _crc(bi), "ids_crc32_k100": _crc(bi100),
                          "same_as_ground_truth_used_for_recall": bool(torch.equal(bi, gt)),
                          "tensor_pipe": load_profile_note("bf_tc")}
    except Exception as e:  # noqa: BLE001
        ex["bf_query"] = {"error": repr(e)[-300:]}
    # native uint8 rows: bench.py's vectors are integer-valued in [0, 255], so the SAME base, queries and
    # graph can be searched from 1-byte rows -- results must be identical, the gather traffic is a quarter
    try:
        if bool((base == base.round()).all()) and float(base.min()) >= 0 and float(base.max()) <= 255:
            g8 = ggnn.GGNN()
            g8.set_return_results_on_gpu(True)
            g8.set_base(base.to(torch.uint8))
            g8._prepare(a.k_build)
            g8._shards[0].graph = idx.get_graph(0)
            g8._measure = 0
            q8 = query.to(torch.uint8)
            (i8, d8), ms8 = timed(lambda: g8.query(q8, K, a.tau_query, a.max_iterations), reps=10)
            (i32, d32), ms32 = timed(lambda: idx.query(query, K, a.tau_query, a.max_iterations), reps=10)
            st2 = [torch.cuda.Stream(dev) for _ in range(2)]

            def piped(g, q, n=20):
                for i in range(n):
                    with torch.cuda.stream(st2[i % 2]):
                        g.query(q, K, a.tau_query, a.max_iterations)
                for s_ in st2:
                    torch.cuda.current_stream(dev).wait_stream(s_)
            _, p8 = timed(lambda: piped(g8, q8), reps=2)
            _, p32 = timed(lambda: piped(idx, query), reps=2)
            n_iter, n_dist, alg32 = _query_stats(idx, query, K, a.tau_query, a.max_iterations)
            cfg = idx.get_graph(0).config
            alg8 = 1.0 * cfg.D * (Nq + n_dist) + 4.0 * cfg.KBuild * n_iter + (4.0 * cfg.S + 8.0 * K) * Nq
            ex["uint8_native"] = {"what": "same base / queries / graph as the headline, vectors stored as uint8 (1-byte rows, integer dp4a distances)",
                                  "single_batch_ms": ms8, "fp32_single_batch_ms": ms32, "pipelined_ms_per_batch": p8 / 20,
                                  "fp32_pipelined_ms_per_batch": p32 / 20, "queries_per_s_pipelined": Nq / (p8 / 20 * 1e-3),
                                  "results_identical_to_fp32": bool(torch.equal(i8, i32) and torch.equal(d8, d32)),
                                  "algorithmic_bytes_per_launch": alg8, "fp32_algorithmic_bytes_per_launch": alg32,
                                  "base_bytes": int(base.numel()), "fp32_base_bytes": int(base.numel() * 4)}
            # ... and bf_query on the uint8 rows: exact integer contraction on the int8 tensor cores (csrc/bf_i8.cu)
            try:
                (b8i, b8d), ms_bf8 = timed(lambda: g8.bf_query(q8, K), reps=3)
                (b8i100, _), ms_bf8_100 = timed(lambda: g8.bf_query(q8, 100), reps=3)
                ex["bf_query_uint8"] = {"what": f"{a.n_base}x{a.dim} uint8 base, {Nq} queries: tcgen05.mma.kind::i8 (u8 x u8 -> s32, exact) "
                                                "+ integer re-rank, rows never widened",
                                        "k10_ms": ms_bf8, "k100_ms": ms_bf8_100,
                                        "k10_useful_tops": 2.0 * Nq * a.n_base * a.dim / (ms_bf8 * 1e-3) / 1e12,
                                        "ids_identical_to_fp32_bf_query": bool(torch.equal(b8i, gt)),
                                        "ids_crc32_k10": _crc(b8i), "ids_crc32_k100": _crc(b8i100),
                                        "tensor_pipe": load_profile_note("bf_i8")}
            except Exception as e:  # noqa: BLE001
                ex["bf_query_uint8"] = {"error": repr(e)[-300:]}
            del g8
    except Exception as e:  # noqa: BLE001
        ex["uint8_native"] = {"error": repr(e)[-300:]}
    # opt-in interleaved copy of the fp32 base (GGNN_B200_INTERLEAVED_BASE=1): one 16-byte shared-memory load per lane and
    # row instead of four 4-byte loads, same results, twice the memory of the base
    try:
        st2 = [torch.cuda.Stream(dev) for _ in range(2)]

        def piped32(n=20):
            for i in range(n):
                with torch.cuda.stream(st2[i % 2]):
                    idx.query(query, K, a.tau_query, a.max_iterations)
            for s_ in st2:
                torch.cuda.current_stream(dev).wait_stream(s_)
        (r_nat, d_nat), ms_nat = timed(lambda: idx.query(query, K, a.tau_query, a.max_iterations), reps=15)
        _, p_nat = timed(piped32, reps=3)
        os.environ["GGNN_B200_INTERLEAVED_BASE"] = "1"
        try:
            (r_il, d_il), ms_il = timed(lambda: idx.query(query, K, a.tau_query, a.max_iterations), reps=15)
            _, p_il = timed(piped32, reps=3)
        finally:
            del os.environ["GGNN_B200_INTERLEAVED_BASE"]
            idx._shards[0].base_il = None
        ex["interleaved_base"] = {"what": "GGNN_B200_INTERLEAVED_BASE=1: second copy of the fp32 base with a lane's four dims adjacent (LDS.128)",
                                  "single_batch_ms": ms_il, "natural_single_batch_ms": ms_nat, "pipelined_ms_per_batch": p_il / 20,
                                  "natural_pipelined_ms_per_batch": p_nat / 20, "results_identical": bool(torch.equal(r_il, r_nat) and torch.equal(d_il, d_nat)),
                                  "extra_bytes": int(base.numel() * 4)}
    except Exception as e:  # noqa: BLE001
        ex["interleaved_base"] = {"error": repr(e)[-300:]}
    # a harder data set (intrinsic dimension 16): operating point found by sweeping the REFERENCE first
    # (profiles/r02_reference_sweep_manifold16.json, procedure of ggnn_benchmark.cpp:186-200)
    try:
        hb, hq = gen_gpu(a.n_base, Nq, a.dim, a.hard_kind, a.seed, dev)
        h = ggnn.GGNN()
        h.set_return_results_on_gpu(True)
        h.set_base(hb)
        h.build(a.k_build, a.tau_build, a.refine)
        hgt, _ = h.bf_query(hq, K)
        (hi, _), hms = timed(lambda: h.query(hq, K, a.hard_tau, a.hard_iterations), reps=7)
        ex["hard_operating_point"] = {"workload": f"{a.n_base}x{a.dim} ({a.hard_kind}), {Nq} queries, tau_query={a.hard_tau} "
                                                  f"max_iterations={a.hard_iterations}", "single_batch_ms": hms,
                                      "queries_per_s": Nq / (hms * 1e-3), "recall_at_10": recall_at_k(hgt, hi, K),
                                      "bf_ids_crc32": _crc(hgt)}
        del h, hb, hq
    except Exception as e:  # noqa: BLE001
        ex["hard_operating_point"] = {"error": repr(e)[-300:]}
    # C++ host API on the same graph: include/ggnn/ggnn.hpp through examples/host_bench (loads part_0.ggnn)
    try:
        ex["cpp_host"] = cpp_host_bench(a, idx, base, query)
    except Exception as e:  # noqa: BLE001
        ex["cpp_host"] = {"error": repr(e)[-300:]}
    # config 3: 10M x 96 cosine, build + query on one GPU -- in a process of its own (tools/shard_probe.py), so that the
    # figure does not depend on what this process has allocated and freed before
    try:
        torch.cuda.empty_cache()
        env = dict(os.environ, PROBE_BF="1")
        p3 = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "shard_probe.py"), str(a.c3_n), "96", "manifoldcos8", "1", "12",
                             str(a.tau_query), str(a.max_iterations)], capture_output=True, text=True, timeout=900, env=env)
        if p3.returncode != 0:
            raise RuntimeError(p3.stderr[-300:])
        r3 = json.loads([ln for ln in p3.stdout.splitlines() if ln.startswith("{")][-1])
        peak, _ = measured_peaks()
        ex["config3"] = {"workload": f"{a.c3_n}x96 fp32 (manifoldcos8), cosine, {Nq} queries, k_query={K} tau_query={a.tau_query} "
                                     f"max_iterations={a.max_iterations} (own process: tools/shard_probe.py)",
                         "build_s": r3["build_s"], "bf_query_ms": r3.get("bf_query_ms"), "single_batch_ms": r3["query_ms_median"],
                         "queries_per_s": Nq / (r3["query_ms_median"] * 1e-3), "recall_at_10": r3.get("recall_at_10"),
                         "pops_per_query": r3["pops_per_query"], "dists_per_query": r3["dists_per_query"],
                         "roofline": {"bound": "hbm", "achieved": r3["algorithmic_gbs"], "peak": peak, "unit": "GB/s",
                                      "frac": r3["algorithmic_gbs"] / peak,
                                      "algorithmic_bytes_per_launch": r3["algorithmic_bytes_per_launch"]},
                         "bf_ids_crc32": r3.get("bf_ids_crc32")}
    except Exception as e:  # noqa: BLE001
        ex["config3"] = {"error": repr(e)[-300:]}
    return ex


def config1_data():
    """BASELINE config 1 (README.md:94-95 of the reference): torch.rand on the CPU, base first, then the queries"""
    g = torch.Generator().manual_seed(1234)
    return torch.rand((10_000, 128), generator=g), torch.rand((10_000, 128), generator=g)


def load_profile_note(name):
    """figures taken from the committed ncu captures (profiles/), keyed by kernel"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_figures.json")))[name]
    except Exception:
        return None


def cpp_host_bench(a, idx, base, query):
    """the same query batch through the C++20 host API (include/ggnn/ggnn.hpp -> C ABI) in a separate process: graph
    loaded from the part file stored here, queries in pinned host memory, synchronous ggnn::GGNN::query per batch"""
    exe = os.path.join(ROOT, "examples", "host_bench")
    if not os.path.exists(exe):
        return {"unavailable": "examples/host_bench not built"}
    wd = os.path.join("/tmp", f"ggnn_cpp_host_{os.getpid()}")
    os.makedirs(wd, exist_ok=True)
    try:
        base.cpu().numpy().tofile(os.path.join(wd, "base.bin"))
        query.cpu().numpy().tofile(os.path.join(wd, "query.bin"))
        idx.set_working_directory(wd)
        idx.store()
        p = subprocess.run([exe, wd, str(a.n_base), str(a.n_query), str(a.dim), str(a.k_build), str(a.k_query), str(a.tau_query),
                            str(a.max_iterations), "20"], capture_output=True, text=True, timeout=300)
        if p.returncode != 0:
            return {"error": f"rc={p.returncode}: {p.stderr[-300:]}"}
        r = json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1])
        ids_py, _ = idx.query(query, a.k_query, a.tau_query, a.max_iterations)
        r["ids_crc32_python_api"] = _crc(ids_py)
        r["same_ids_as_python_api"] = r.get("ids_crc32") == r["ids_crc32_python_api"]
        return r
    finally:
        for f in os.listdir(wd):
            os.remove(os.path.join(wd, f))
        os.rmdir(wd)


def load_traffic():
    """dram bytes per launch of the traversal kernel from the committed ncu capture (profiles/), or null"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "query_kernel_traffic.json")))["dram_bytes_per_launch"]
    except Exception:
        return None


def cpu_baseline(a, idx, base, query, gr):
    """CPU oracle port of the same traversal (same graph), all host cores, bounded query sample (~10-30 s)."""
    from oracle import pyoracle as O
    cores = os.cpu_count() or 1
    O.set_threads(cores)  # torchrun exports OMP_NUM_THREADS=1
    torch.set_num_threads(cores)
    n = a.n_query
    b = base.cpu().numpy()
    q = query[:n].cpu().numpy()
    g0 = gr.layer_graph(0).cpu().numpy()
    sp = gr.layer_translation(3).cpu().numpy()
    ns = gr.nn1_stats.cpu().numpy()
    O.query(b, q[:cores], g0, sp, ns, a.k_query, a.tau_query, a.max_iterations)  # warm
    t0 = time.perf_counter()
    reps = 0
    while reps < 200 and time.perf_counter() - t0 < 10.0:  # bounded sample: ~10 s of CPU work
        O.query(b, q, g0, sp, ns, a.k_query, a.tau_query, a.max_iterations)
        reps += 1
    dt = time.perf_counter() - t0
    out = {"value": n * reps / dt, "unit": "queries/s", "cores": cores, "kind": "port",
           "sample": f"all {n} queries of one batch x {reps} passes, same graph and parameters, OpenMP over queries ({dt:.2f} s)"}
    # exact brute force on the host cores (torch CPU SGEMM formulation), for context
    try:
        nb = min(n, 256)
        bt, qt = base.cpu(), query[:nb].cpu()
        t0 = time.perf_counter()
        d = (qt * qt).sum(1, keepdim=True) + (bt * bt).sum(1).unsqueeze(0) - 2.0 * qt @ bt.t()
        d.topk(a.k_query, dim=1, largest=False)
        out["cpu_bruteforce_qps"] = nb / (time.perf_counter() - t0)
        out["cpu_bruteforce_sample"] = f"{nb} queries, torch CPU sgemm + topk"
    except Exception as e:  # pragma: no cover
        out["cpu_bruteforce_qps"] = None
        out["cpu_bruteforce_sample"] = str(e)
    return out


# ------------------------------------------------------------------------------------------------
def run_reference(a):
    """the UNMODIFIED reference (oracle/_ref) through ggnn::GGNN on the same config; rank 0 only.
    value = its own device-timed traversal kernel (the cudaEvent figure it logs per call, gpu_instance.cu:687-712), summed
    over the batches of a step; e2e = wall clock around its synchronous ggnn::GGNN::query() on pinned host buffers."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    drv = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    base_line = {"impl": "reference", "metric": "queries/sec @ recall@10", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup}
    if not os.path.exists(drv):
        emit({"impl": "reference", "unavailable": "oracle/_ref/ref_driver not built (bash oracle/build_ref.sh needs /root/reference)"})
        return
    shards = a.gpus
    B, Nq = max(1, a.batches_per_step), a.n_query
    wd = os.path.join("/tmp", f"ggnn_ref_bench_{os.getpid()}")
    os.makedirs(wd, exist_ok=True)
    dev = torch.device("cuda", 0)
    parts, q_all = [], None
    for s in range(shards):
        b, q_all = gen_gpu(a.n_base, Nq * B, a.dim, a.kind, a.seed, dev, shard_index=s)
        parts.append(b.cpu().numpy())
    np.concatenate(parts).tofile(os.path.join(wd, "base.bin"))
    q_all.cpu().numpy().tofile(os.path.join(wd, "query.bin"))
    del parts
    torch.cuda.empty_cache()
    reps = (a.warmup + a.steps) * B
    args = [drv, f"dir={wd}", f"n={a.n_base * shards}", f"nq={Nq * B}", f"batch={Nq}", f"d={a.dim}", "measure=0",
            f"kbuild={a.k_build}", f"tau_build={a.tau_build}", f"refine={a.refine}", "build=1",
            f"build_reps={2 if shards == 1 else 1}", f"kquery={a.k_query}",
            f"tau_query={a.tau_query}", f"max_iter={a.max_iterations}", f"query_reps={reps}", f"gpu_reps={reps if shards == 1 else 0}",
            f"bf={a.k_query if shards == 1 else 0}", f"bf_nq={Nq}", "dump=1", f"gpus={shards}", f"shard={a.n_base}"]
    p = subprocess.run(args, capture_output=True, text=True)
    if p.returncode != 0:
        emit({"impl": "reference", "unavailable": f"ref_driver rc={p.returncode}: {p.stderr[-300:]}"})
        return
    r = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
    skip = a.warmup * B
    e2e = r["query_e2e_ms"][skip:]
    # device-timed kernel of every call (N = 1: results left on the GPU; N > 1: the reference cannot keep results on the
    # GPUs, its per-shard kernel times of one call overlap across GPUs -> take the slowest GPU's kernel per call)
    if r.get("query_gpu_kernel_ms"):
        kern = r["query_gpu_kernel_ms"][skip:]
    else:
        km = r.get("query_kernel_ms") or []
        kern = [max(km[i * shards:(i + 1) * shards]) for i in range(skip, len(km) // shards)] if km else e2e
    rec = bf_crc = None
    try:
        ids = np.fromfile(os.path.join(wd, "query_ids.bin"), np.int32).reshape(Nq * B, a.k_query)[:Nq]
        gt = np.fromfile(os.path.join(wd, "bf_ids.bin"), np.int32).reshape(-1, a.k_query)[:Nq]
        rec = recall_at_k(torch.from_numpy(gt), torch.from_numpy(ids), a.k_query)
        bf_crc = zlib.crc32(gt.tobytes()) & 0xffffffff
    except Exception:
        pass
    for f in os.listdir(wd):
        os.remove(os.path.join(wd, f))
    config1 = None
    if shards == 1:  # BASELINE config 1 through the reference's API (same tensors as our arm's extra.config1)
        try:
            bc, qc = config1_data()
            bc.numpy().tofile(os.path.join(wd, "base.bin"))
            qc.numpy().tofile(os.path.join(wd, "query.bin"))
            p1 = subprocess.run([drv, f"dir={wd}", "n=10000", "nq=10000", "d=128", "measure=0", f"kbuild={a.k_build}",
                                 f"tau_build={a.tau_build}", f"refine={a.refine}", "build=1", f"kquery={a.k_query}",
                                 f"tau_query={a.tau_query}", f"max_iter={a.max_iterations}", "query_reps=11", "gpu_reps=0",
                                 f"bf={a.k_query}", "dump=1"], capture_output=True, text=True)
            r1 = json.loads([l for l in p1.stdout.splitlines() if l.startswith("{")][-1])
            i1 = np.fromfile(os.path.join(wd, "query_ids.bin"), np.int32).reshape(10000, a.k_query)
            g1 = np.fromfile(os.path.join(wd, "bf_ids.bin"), np.int32).reshape(10000, a.k_query)
            config1 = {"build_s": r1.get("build_s"), "query_ms_host_to_host": float(np.median(r1["query_e2e_ms"][1:])),
                       "kernel_ms": float(np.median(r1["query_kernel_ms"][1:])) if r1.get("query_kernel_ms") else None,
                       "recall_at_10": recall_at_k(torch.from_numpy(g1), torch.from_numpy(i1), a.k_query),
                       "bf_ids_crc32": zlib.crc32(g1.tobytes()) & 0xffffffff}
            for f in os.listdir(wd):
                os.remove(os.path.join(wd, f))
        except Exception as e:  # noqa: BLE001
            config1 = {"error": repr(e)[-300:]}
    ms_step = float(np.sum(kern)) / a.steps
    e2e_step = float(np.sum(e2e)) / a.steps
    q_per_step = Nq * B
    val = q_per_step / (ms_step * 1e-3) * shards
    e2e_val = q_per_step / (e2e_step * 1e-3) * shards
    unit = "queries/s" if shards == 1 else "queries/s x shards searched"
    wall_gpu = r.get("query_gpu_ms")
    out = dict(base_line)
    out.update({
        "value": val, "unit": unit, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "recall_at_10": rec,
        "config": {"workload": f"SIFT1M-shape {a.n_base}x{a.dim} fp32 ({a.kind}) per GPU shard, {Nq} queries per batch, Euclidean, "
                               f"k_build={a.k_build} tau_build={a.tau_build} refine={a.refine}, k_query={a.k_query} "
                               f"tau_query={a.tau_query} max_iterations={a.max_iterations}",
                   "batches_per_step": B, "queries_per_step": q_per_step, "shards": shards,
                   "value_is": "the reference's own cudaEvent time of its traversal kernel per call, summed over the step",
                   "reference_build_s": r.get("build_s"), "reference_build_s_earlier_in_process": r.get("build_s_all"),
                   "reference_kernel_ms_per_batch": float(np.mean(kern)),
                   "reference_wall_ms_per_batch_results_on_gpu": float(np.mean(wall_gpu[skip:])) if wall_gpu else None,
                   "bf_ids_crc32_k10": bf_crc, "bf_s": r.get("bf_s"), "config1": config1},
        "e2e": {"value": e2e_val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "mode": "wall clock around one synchronous ggnn::GGNN::query() per batch, pinned host query, results to the host"},
        "cpu_baseline": {"value": e2e_val, "unit": unit, "kind": "reference", "cores": shards,
                         "sample": "the reference has no CPU implementation of this path: this arm runs its own CUDA "
                                   "kernels (unmodified sources, compiled for sm_100a) through ggnn::GGNN::query on "
                                   "pinned host buffers; host threads = 1 per GPU (+ CPU merge threads for N>1)"},
    })
    emit(out)


_REAL_STDOUT = None


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)   # one step = --batches-per-step batches of 10 000 queries (~11 ms)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batches-per-step", dest="batches_per_step", type=int, default=25)
    ap.add_argument("--streams", type=int, default=2)
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: 'peer' = the traversal kernel stores its lists into every rank's buffer over NVLink (CUDA IPC); "
                         "'nccl' = ncclAllGather on pre-allocated buffers (also the automatic fallback)")
    ap.add_argument("--e2e-depth", dest="e2e_depth", type=int, default=2,
                    help="batches in flight in the end-to-end measurement (1 = synchronous GGNN.query() per batch only)")
    ap.add_argument("--no-config4", dest="config4", action="store_false", help="skip BASELINE config 4 (100M vectors in 8 shards)")
    ap.add_argument("--c4-shards", dest="c4_shards", type=int, default=8)
    ap.add_argument("--c4-shard-size", dest="c4_shard_size", type=int, default=12_500_000)
    ap.add_argument("--c4-steps", dest="c4_steps", type=int, default=5)
    ap.add_argument("--c4-batches", dest="c4_batches", type=int, default=4)
    ap.add_argument("--c4-gt-queries", dest="c4_gt_queries", type=int, default=2000)
    ap.add_argument("--no-extras", dest="extras", action="store_false", help="N = 1: skip bf_query / config 3 / hard point / C++ host")
    ap.add_argument("--c3-n", dest="c3_n", type=int, default=10_000_000)
    ap.add_argument("--hard-kind", dest="hard_kind", default="manifold16")
    ap.add_argument("--hard-tau", dest="hard_tau", type=float, default=1.0)
    ap.add_argument("--hard-iterations", dest="hard_iterations", type=int, default=200)
    for k, v in DEF.items():
        ap.add_argument("--" + k.replace("_", "-"), type=type(v), default=v)
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    # the contract is ONE JSON line on stdout: libraries that write to file descriptor 1 themselves (NCCL prints its
    # version there when a communicator is created) are sent to stderr; only emit() writes to the real stdout
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    main()
