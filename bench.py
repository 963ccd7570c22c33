#!/usr/bin/env python
"""bench.py -- queries/s @ recall@10 of the GGNN batched query hot path on SIFT1M-shape synthetic fp32 data.

  python bench.py --gpus N --steps K --warmup W             our arm (ggnn_b200, sm_100a kernels)
  python bench.py --impl reference --gpus N --steps K ...   the UNMODIFIED reference CUDA library
                                                            (oracle/_ref, built from /root/reference) through
                                                            its own public API ggnn::GGNN, same data and parameters
N > 1: launched by torchrun, one rank per GPU; rank r owns one 1M-vector shard (the reference's row sharding),
the query batch is broadcast, per-rank top-K lists are all-gathered over NCCL and merged by one kernel.

A "step" = one search of the whole query batch (10 000 queries).  Rank 0 prints ONE JSON line.
  value    queries/s with the query batch resident in HBM and results left in HBM (CUDA events, max over ranks)
           -- for N > 1 multiplied by the number of shards searched per query (weak scaling: per-GPU work fixed)
  e2e      the same through the public API GGNN.query() with a pinned HOST query tensor and results copied
           back to the host inside the timed region
  roofline algorithmic bytes of the traversal / kernel time, against the measured HBM copy peak
  cpu_baseline  the CPU oracle port of the same traversal, all host cores, bounded query sample
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DEF = dict(n_base=1_000_000, n_query=10_000, dim=128, k_build=24, tau_build=0.5, refine=2, k_query=10,
           tau_query=0.64, max_iterations=400, kind="manifold8", seed=1234)


def gen_gpu(N, Nq, D, kind, seed, device, shard_index=0):
    """synthetic SIFT1M-shape data generated on the device (fp32, no dataset files offline).
    'manifold<d>' (default manifold8: at the reference's documented SIFT1M operating point tau_query=0.64,
                  max_iterations=400 it reaches the documented recall@10 of 0.99, measured for both implementations): SIFT-like -- points on a d-dimensional linear manifold embedded in D dims
                  (low intrinsic dimension like real descriptors), per-dim std 40 around 128, unit Gaussian noise,
                  clipped to [0,255] and rounded: integer-valued vectors stored as fp32, like SIFT.
    'uniform'   : U[0,1) in every dim (the reference's README example; intrinsic dimension = D, a very hard ANN
                  instance: neither the reference nor this implementation gets useful recall at 1M points).
    'clustered' : mixture of 1000 isotropic Gaussians (intrinsic dimension = D, equally hard).
    Queries are drawn from the same distribution with a separate generator; shards differ by `shard_index`."""
    g = torch.Generator(device=device).manual_seed(seed)
    gb = torch.Generator(device=device).manual_seed(seed + 17 * (shard_index + 1))
    if kind == "uniform":
        return torch.rand((N, D), generator=gb, device=device), torch.rand((Nq, D), generator=g, device=device)
    if kind.startswith("manifoldcos"):  # DEEP-like: unit-normalised vectors with low intrinsic dimension (cosine)
        d = int(kind[len("manifoldcos"):] or 8)
        A = torch.randn((d, D), generator=g, device=device) / (d ** 0.5)

        def draw_c(n, gen):
            out = torch.empty((n, D), device=device)
            for lo in range(0, n, 1 << 20):
                m = min(1 << 20, n - lo)
                x = torch.randn((m, d), generator=gen, device=device) @ A + 0.02 * torch.randn((m, D), generator=gen, device=device)
                out[lo:lo + m] = x / x.norm(dim=1, keepdim=True)
            return out
        query = draw_c(Nq, g)
        return draw_c(N, gb), query
    if kind.startswith("manifold"):
        d = int(kind[len("manifold"):] or 16)
        A = torch.randn((d, D), generator=g, device=device) / (d ** 0.5)

        def draw(n, gen):
            out = torch.empty((n, D), device=device)
            for lo in range(0, n, 1 << 20):  # chunked to bound temporaries
                m = min(1 << 20, n - lo)
                z = torch.randn((m, d), generator=gen, device=device)
                x = (z @ A) * 40.0 + 128.0 + torch.randn((m, D), generator=gen, device=device)
                out[lo:lo + m] = x.round_().clamp_(0, 255)
            return out
        query = draw(Nq, g)
        return draw(N, gb), query
    nc = 1000
    centers = torch.rand((nc, D), generator=g, device=device) * 160 + 20

    def draw(n, gen):
        c = torch.randint(0, nc, (n,), generator=gen, device=device)
        x = centers[c] + torch.randn((n, D), generator=gen, device=device) * 25
        return x.round_().clamp_(0, 255)
    query = draw(Nq, g)
    return draw(N, gb), query


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
def run_ours(a):
    import torch.distributed as dist
    import ggnn_b200 as ggnn
    from ggnn_b200 import _lib
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

    base, query = gen_gpu(a.n_base, a.n_query, a.dim, a.kind, a.seed, dev, shard_index=rank)
    idx = ggnn.GGNN()
    idx.set_gpus([local])
    idx.set_return_results_on_gpu(True)
    idx.set_base(base)
    torch.cuda.synchronize()
    t0 = time.time()
    idx.build(a.k_build, a.tau_build, a.refine)
    torch.cuda.synchronize()
    build_s = time.time() - t0
    K = a.k_query

    def local_query(q):
        return idx.query(q, K, a.tau_query, a.max_iterations)

    # N > 1: every pipeline (CUDA stream) gets its own NCCL communicator and its own receive buffer for the broadcast
    # query batch, so the collectives of batches in flight neither serialise on one communicator nor alias
    n_pipes = max(1, a.streams)
    groups = [dist.new_group(backend="nccl") for _ in range(n_pipes)] if world > 1 else [None] * n_pipes
    # Query distribution at N > 1.  "replicated" (default): every rank holds the batch -- device-resident for `value`,
    # its own pinned host copy for `e2e` (each rank copies host->device over its own PCIe link, exactly what the
    # reference does: one H2D per GPU, gpu_instance.cu:638-641) -- so the only collective is the top-K exchange.
    # "broadcast": the batch lives on rank 0 only and is broadcast over NVLink inside the timed region.
    bcast = world > 1 and a.query_distribution == "broadcast"
    if world > 1:  # the seeded generator must have produced the same batch everywhere
        chk = torch.stack((query.double().sum(), query[::97].double().sum()))
        ref_chk = chk.clone()
        dist.broadcast(ref_chk, src=0)
        assert torch.equal(chk, ref_chk), "query batches differ between ranks"
    q_bufs = [query if (rank == 0 or world == 1 or not bcast) else torch.empty_like(query) for _ in range(n_pipes)]

    def step_device(pipe=0):
        if world == 1:
            return local_query(query)
        return gd.distributed_query(local_query, gd.gpu_merge, q_bufs[pipe], K, a.n_base, group=groups[pipe], broadcast=bcast)

    # ground truth + recall (untimed): exact brute force on every shard, merged the same way
    def bf_local(q):
        return idx.bf_query(q, K)
    if world == 1:
        gt, _ = bf_local(query)
    else:
        gt, _ = gd.distributed_query(bf_local, gd.gpu_merge, query, K, a.n_base, broadcast=False)
    ids, dists = step_device()
    rec = recall_at_k(gt, ids, K)

    # algorithmic bytes of one step on this rank: counters from an untimed stats launch
    gr = idx.get_graph(0)
    cfg = gr.config
    stats = torch.zeros((a.n_query, 2), dtype=torch.int32, device=dev)
    tmp_i = torch.empty((a.n_query, K), dtype=torch.int32, device=dev)
    tmp_d = torch.empty((a.n_query, K), dtype=torch.float32, device=dev)
    p = _lib.QueryParams()
    p.D, p.measure, p.KQuery, p.tau_query, p.max_iterations = cfg.D, 0, K, a.tau_query, a.max_iterations
    p.N_base, p.KBuild, p.num_starting_points = cfg.N, cfg.KBuild, cfg.S
    p.d_base, p.d_query, p.d_graph = idx._shards[0].base.data_ptr(), query.data_ptr(), gr.graph.data_ptr()
    p.d_starting_points, p.d_nn1_stats = gr.layer_translation(3).data_ptr(), gr.nn1_stats.data_ptr()
    p.d_query_results, p.d_query_results_dists, p.d_stats = tmp_i.data_ptr(), tmp_d.data_ptr(), stats.data_ptr()
    p.shards_per_gpu, p.on_gpu_shard_id = 1, 0
    _lib.check(_lib.lib().ggnn_b200_query(C.byref(p), a.n_query, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    st = stats.to(torch.float64).sum(0).tolist()
    n_iter, n_dist = st[0], st[1]
    S = cfg.S
    alg_bytes = 4.0 * cfg.D * (a.n_query * (1 + 0) + n_dist) + 4.0 * cfg.KBuild * n_iter + (4.0 * S + 8.0 * K) * a.n_query
    # (n_dist already includes the S start-point evaluations)

    # ---- timed region 1: device-resident ----
    # Steps are independent query batches.  With --streams 2 (default) consecutive batches are enqueued on two
    # alternating CUDA streams, so the next batch's CTAs fill the SMs that the previous batch's last, long
    # queries leave idle (the kernel's tail is one query latency).  --streams 1 serialises the batches.
    sampler = ClockSampler(local)
    n_streams = n_pipes
    streams = [torch.cuda.Stream(dev) for _ in range(n_streams)] if n_streams > 1 else [torch.cuda.current_stream(dev)]
    outs = [None] * n_streams

    def run_steps(n):
        for s in range(n):
            st = streams[s % n_streams]
            with torch.cuda.stream(st):
                outs[s % n_streams] = step_device(s % n_streams)

    run_steps(a.warmup)
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream(dev)
    e_start.record(cur)
    for st in streams:
        st.wait_event(e_start)
    run_steps(a.steps)
    for st in streams:
        cur.wait_stream(st)
    e_end.record(cur)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = e_start.elapsed_time(e_end)
    clocks = sampler.stop() if rank == 0 else None

    # dominant kernel alone (the traversal kernel of this rank), CUDA events on the launch stream
    kev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * a.steps)]
    for s in range(a.steps):
        kev[2 * s].record()
        local_query(query)
        kev[2 * s + 1].record()
    torch.cuda.synchronize()
    kernel_ms = float(np.mean([kev[2 * s].elapsed_time(kev[2 * s + 1]) for s in range(a.steps)]))

    # ---- timed region 2: end to end through the public API, host buffers ----
    idx.set_return_results_on_gpu(False)
    q_host = query.cpu().pin_memory()

    def step_e2e():
        if world == 1:
            return idx.query(q_host, K, a.tau_query, a.max_iterations)   # H2D + kernels + D2H inside
        qd = q_host.to(dev, non_blocking=True) if (rank == 0 or not bcast) else torch.empty_like(query)
        idx.set_return_results_on_gpu(True)
        r = gd.distributed_query(local_query, gd.gpu_merge, qd, K, a.n_base, broadcast=bcast)
        return r[0].cpu(), r[1].cpu()
    for _ in range(max(3, a.warmup)):   # (the first calls allocate the pinned result buffers)
        step_e2e()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_sync_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms, e2e_depth = e2e_sync_ms, 1
    if world == 1 and a.e2e_depth > 1:
        # the same K steps with up to e2e_depth batches in flight through GGNN.query_async(): every step still copies
        # its own queries host->device and its own results device->host inside the timed region
        def run_async(n):
            pending, last = [], None
            for _ in range(n):
                pending.append(idx.query_async(q_host, K, a.tau_query, a.max_iterations))
                if len(pending) >= a.e2e_depth:
                    last = pending.pop(0).result()
            while pending:
                last = pending.pop(0).result()
            return last
        run_async(max(4, a.warmup))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r_async = run_async(a.steps)
        torch.cuda.synchronize()
        e2e_ms, e2e_depth = (time.perf_counter() - t0) * 1e3, a.e2e_depth
        r_sync = step_e2e()
        assert torch.equal(r_async[0], r_sync[0]) and torch.equal(r_async[1], r_sync[1]), "async and sync results differ"
    elif world > 1 and a.e2e_depth > 1 and n_streams > 1:
        # N > 1: one batch in flight per pipeline (stream + communicator).  Rank 0 copies the batch host->device, it is
        # broadcast, every rank searches its shard, the lists are gathered and merged, rank 0 copies the result back.
        idx.set_return_results_on_gpu(True)
        qd = [torch.empty_like(query) for _ in range(n_streams)]
        pin_i = [torch.empty((a.n_query, K), dtype=torch.int32, pin_memory=True) for _ in range(n_streams)]
        pin_d = [torch.empty((a.n_query, K), dtype=torch.float32, pin_memory=True) for _ in range(n_streams)]
        evs = [None] * n_streams

        def run_async(n):
            for s_ in range(n):
                pipe = s_ % n_streams
                if evs[pipe] is not None:
                    evs[pipe].synchronize()  # the step that used this pipeline's buffers has delivered its result
                with torch.cuda.stream(streams[pipe]):
                    if rank == 0 or not bcast:
                        qd[pipe].copy_(q_host, non_blocking=True)
                    r = gd.distributed_query(local_query, gd.gpu_merge, qd[pipe], K, a.n_base, group=groups[pipe], broadcast=bcast)
                    if rank == 0:
                        pin_i[pipe].copy_(r[0], non_blocking=True)
                        pin_d[pipe].copy_(r[1], non_blocking=True)
                    evs[pipe] = streams[pipe].record_event()
            for e in evs:
                if e is not None:
                    e.synchronize()
        run_async(max(4, a.warmup))
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run_async(a.steps)
        torch.cuda.synchronize()
        e2e_ms, e2e_depth = (time.perf_counter() - t0) * 1e3, n_streams
        if rank == 0:
            r_sync = step_e2e()
            last = (a.steps - 1) % n_streams
            assert torch.equal(pin_i[last], r_sync[0]) and torch.equal(pin_d[last], r_sync[1]), "async and sync results differ"
        else:
            step_e2e()
    times = torch.tensor([total_ms, e2e_ms, kernel_ms, e2e_sync_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, kernel_ms, e2e_sync_ms = times.tolist()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = total_ms / a.steps
    shards = world
    qps = a.n_query / (ms_per_step * 1e-3)
    e2e_qps = a.n_query / (e2e_ms / a.steps * 1e-3)
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    cpu = cpu_baseline(a, idx, base, query, gr)
    out = {
        "metric": "queries/sec @ recall@10", "value": qps * shards, "unit": "queries/s" if shards == 1 else "queries/s x shards searched",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"SIFT1M-shape {a.n_base}x{a.dim} fp32 ({a.kind}) per GPU shard, {a.n_query} queries, "
                               f"Euclidean, k_build={a.k_build} tau_build={a.tau_build} refine={a.refine}, "
                               f"k_query={K} tau_query={a.tau_query} max_iterations={a.max_iterations}",
                   "shards": shards, "vectors_total": a.n_base * shards, "queries_per_s": qps, "recall_at_10": rec,
                   "l2_policy": "inputs_larger_than_l2 (512 MB base per shard, gather-random)",
                   "parallelism": (f"base row-sharded x{shards}, queries {a.query_distribution}, NCCL all_gather of [Nq,K] + merge kernel"
                                   if shards > 1 else "single shard"),
                   "build_s": build_s,
                   "pipelining": f"{n_streams} CUDA stream(s): independent query batches overlap their tails" if n_streams > 1
                   else "none (batches serialised on one stream)",
                   "single_batch_ms": kernel_ms},
        "recall_at_10": rec,
        "e2e": {"value": e2e_qps * shards, "unit": "queries/s" if shards == 1 else "queries/s x shards searched",
                "h2d_bytes_per_step": a.n_query * a.dim * 4 * (1 if (bcast or shards == 1) else shards),
                "d2h_bytes_per_step": a.n_query * K * 8,
                "mode": ((f"GGNN.query_async(), {e2e_depth} batches in flight" if shards == 1 else
                          f"{e2e_depth} batches in flight: " + ("pinned H2D on rank 0, broadcast" if bcast else "pinned H2D on every rank") +
                          ", per-shard query, all_gather, merge, D2H on rank 0")
                         if e2e_depth > 1 else "one synchronous call per step"),
                "sync_value": a.n_query / (e2e_sync_ms / a.steps * 1e-3) * shards},
        "gpu_launches": a.steps * (1 + (1 if shards > 1 else 0)),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": load_traffic(), "kernel": "query_kernel", "kernel_ms": kernel_ms,
                     "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                     "pops_per_query": n_iter / a.n_query, "dists_per_query": n_dist / a.n_query},
        "cpu_baseline": cpu,
    }
    emit(out)
    if world > 1:
        dist.destroy_process_group()


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
           "sample": f"all {n} queries x {reps} passes, same graph and parameters, OpenMP over queries ({dt:.2f} s)"}
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
    """the UNMODIFIED reference (oracle/_ref) through ggnn::GGNN on the same config; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    drv = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    base_line = {"impl": "reference", "metric": "queries/sec @ recall@10", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup}
    if not os.path.exists(drv):
        emit({"impl": "reference", "unavailable": "oracle/_ref/ref_driver not built (bash oracle/build_ref.sh needs /root/reference)"})
        return
    shards = a.gpus
    wd = os.path.join("/tmp", f"ggnn_ref_bench_{os.getpid()}")
    os.makedirs(wd, exist_ok=True)
    dev = torch.device("cuda", 0)
    parts, query = [], None
    for s in range(shards):
        b, q = gen_gpu(a.n_base, a.n_query, a.dim, a.kind, a.seed, dev, shard_index=s)
        parts.append(b.cpu().numpy())
        query = q
    np.concatenate(parts).tofile(os.path.join(wd, "base.bin"))
    query.cpu().numpy().tofile(os.path.join(wd, "query.bin"))
    del parts
    torch.cuda.empty_cache()
    reps = a.warmup + a.steps
    args = [drv, f"dir={wd}", f"n={a.n_base * shards}", f"nq={a.n_query}", f"d={a.dim}", "measure=0",
            f"kbuild={a.k_build}", f"tau_build={a.tau_build}", f"refine={a.refine}", "build=1", f"kquery={a.k_query}",
            f"tau_query={a.tau_query}", f"max_iter={a.max_iterations}", f"query_reps={reps}", f"gpu_reps={reps if shards == 1 else 0}",
            f"bf={a.k_query if shards == 1 else 0}", "dump=1", f"gpus={shards}", f"shard={a.n_base}"]
    p = subprocess.run(args, capture_output=True, text=True)
    if p.returncode != 0:
        emit({"impl": "reference", "unavailable": f"ref_driver rc={p.returncode}: {p.stderr[-300:]}"})
        return
    r = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
    e2e = r["query_e2e_ms"][a.warmup:]
    gpu = r["query_gpu_ms"][a.warmup:] if r.get("query_gpu_ms") else e2e  # N>1: the reference cannot keep results on the GPUs
    rec = None
    try:
        ids = np.fromfile(os.path.join(wd, "query_ids.bin"), np.int32).reshape(a.n_query, a.k_query)
        gt = np.fromfile(os.path.join(wd, "bf_ids.bin"), np.int32).reshape(a.n_query, a.k_query)
        rec = recall_at_k(torch.from_numpy(gt), torch.from_numpy(ids), a.k_query)
    except Exception:
        pass
    for f in os.listdir(wd):
        os.remove(os.path.join(wd, f))
    ms = float(np.mean(gpu))
    e2e_ms = float(np.mean(e2e))
    val = a.n_query / (ms * 1e-3) * shards
    e2e_val = a.n_query / (e2e_ms * 1e-3) * shards
    unit = "queries/s" if shards == 1 else "queries/s x shards searched"
    out = dict(base_line)
    out.update({
        "value": val, "unit": unit, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "recall_at_10": rec,
        "config": {"workload": f"SIFT1M-shape {a.n_base}x{a.dim} fp32 ({a.kind}) per GPU shard, {a.n_query} queries, Euclidean, "
                               f"k_build={a.k_build} tau_build={a.tau_build} refine={a.refine}, k_query={a.k_query} "
                               f"tau_query={a.tau_query} max_iterations={a.max_iterations}", "shards": shards,
                   "reference_build_s": r.get("build_s"), "reference_kernel_ms": r.get("query_gpu_kernel_ms")},
        "e2e": {"value": e2e_val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
    ap.add_argument("--steps", type=int, default=50)   # one step = one 10 000-query batch (~0.5 ms)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=2)
    ap.add_argument("--query-distribution", dest="query_distribution", default="replicated", choices=["replicated", "broadcast"],
                    help="N > 1: every rank holds / copies the query batch itself (like the reference's one H2D per GPU), "
                         "or rank 0 broadcasts it over NVLink inside the timed region")
    ap.add_argument("--e2e-depth", dest="e2e_depth", type=int, default=2,
                    help="batches in flight in the end-to-end measurement (1 = synchronous GGNN.query() per step)")
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
