"""uint8 brute force on the int8 tensor cores (csrc/bf_i8.cu) -- staged check, every stage prints its own verdict:
  1. pack : ggnn_b200_debug_i8_pack against a numpy model of the tile-major 128-byte-swizzled layout, integer norms
  2. mma  : ggnn_b200_debug_i8_mma (one 128 x 128 x D product) against an int64 matmul, with a diagnosis on mismatch
  3. e2e  : GGNN.bf_query on uint8 tensors -- int8 tensor path vs rows widened to fp32 (GGNN_B200_NO_I8_BF=1) vs an exact
            int64 brute force in torch; shapes with ragged tiles, heavy distance ties, K up to 128
  4. time : 1M x 128 base, 10 000 queries (BASELINE config 5 shape on uint8 values)
Usage: python tools/bf_i8_check.py [stages, e.g. 123]"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ggnn_b200 as ggnn  # noqa: E402
from ggnn_b200 import _lib  # noqa: E402

DEV = torch.device("cuda", 0)


def pack_model(rows, n_pad):
    """numpy model of i8_pack_kernel: 16 KB per 128-row tile, 16-byte chunk c of row r at chunk c ^ (r & 7)"""
    n, D = rows.shape
    out = np.zeros((n_pad // 128, 128, 8, 16), dtype=np.uint8)
    padded = np.zeros((n_pad, 128), dtype=np.uint8)
    padded[:n, :D] = rows
    ch = padded.reshape(n_pad // 128, 128, 8, 16)
    r = np.arange(128)
    for c in range(8):
        out[:, r, c ^ (r & 7), :] = ch[:, r, c, :]
    return out.reshape(-1)


def stage_pack():
    l = _lib.lib()
    ok = True
    rng = np.random.default_rng(5)
    for D in (32, 64, 96, 128):
        n, n_pad = 300, 384
        rows = rng.integers(0, 256, (n, D), dtype=np.uint8)
        d_rows = torch.from_numpy(rows).to(DEV)
        d_tiled = torch.full((n_pad * 128,), 0xAB, dtype=torch.uint8, device=DEV)
        d_norms = torch.zeros(n, dtype=torch.int32, device=DEV)
        _lib.check(l.ggnn_b200_debug_i8_pack(d_rows.data_ptr(), n, n_pad, D, d_tiled.data_ptr(), d_norms.data_ptr(), None))
        torch.cuda.synchronize()
        same = np.array_equal(d_tiled.cpu().numpy(), pack_model(rows, n_pad))
        norms = (rows.astype(np.int64) ** 2).sum(1)
        same_n = np.array_equal(d_norms.cpu().numpy().astype(np.int64), norms)
        print(f"[pack] D={D}: layout {'ok' if same else 'DIFFERS'}, norms {'ok' if same_n else 'DIFFER'}", flush=True)
        ok = ok and same and same_n
    return ok


def stage_mma():
    l = _lib.lib()
    ok = True
    rng = np.random.default_rng(6)
    for D, hi in ((128, 256), (32, 256), (64, 256), (96, 256), (128, 4)):
        rows = rng.integers(0, hi, (256, D), dtype=np.uint8)
        if hi == 256:
            rows[3] = 255  # the largest possible products
            rows[130] = 255
        tiled = torch.from_numpy(pack_model(rows, 256)).to(DEV)
        out = torch.full((128, 128), -7, dtype=torch.int32, device=DEV)
        _lib.check(l.ggnn_b200_debug_i8_mma(tiled.data_ptr(), tiled.data_ptr() + 16384, D // 32, out.data_ptr(), None))
        torch.cuda.synchronize()
        got = out.cpu().numpy().astype(np.int64)
        a, b = rows[:128].astype(np.int64), rows[128:].astype(np.int64)
        want = a @ b.T
        same = np.array_equal(got, want)
        print(f"[mma] D={D} values<{hi}: {'ok' if same else 'DIFFERS'}", flush=True)
        if not same:
            ok = False
            sa, sb = rows[:128].astype(np.int8).astype(np.int64), rows[128:].astype(np.int8).astype(np.int64)
            alts = {"transposed": want.T, "A signed": sa @ b.T, "B signed": a @ sb.T, "both signed": sa @ sb.T,
                    "first 32 k only": a[:, :32] @ b[:, :32].T, "untouched (-7)": np.full_like(want, -7), "zero": np.zeros_like(want)}
            for name, alt in alts.items():
                print(f"      equals '{name}': {np.array_equal(got, alt)} ({(got == alt).mean():.4f} of the entries)")
            print(f"      entries equal to the expected product: {(got == want).mean():.4f}; rows fully right: "
                  f"{(got == want).all(1).sum()} / 128, columns fully right: {(got == want).all(0).sum()} / 128")
            print("      got[0:4, 0:6]\n", got[:4, :6], "\n      want[0:4, 0:6]\n", want[:4, :6], flush=True)
    return ok


def exact_reference(base, query, K):
    """(dist, id)-sorted exact top K in int64, distances as the reference's fp32 (exact integers)"""
    b, q = base.to(torch.float64), query.to(torch.float64)  # (no integer matmul on the GPU; exact below 2^53)
    d = ((q * q).sum(1, keepdim=True) - 2 * (q @ b.T) + (b * b).sum(1).unsqueeze(0)).to(torch.int64)
    key = d * (1 << 32) + torch.arange(b.shape[0], device=b.device, dtype=torch.int64).unsqueeze(0)
    top = torch.topk(key, K, dim=1, largest=False, sorted=True).values
    return (top & 0xffffffff).to(torch.int32), (top >> 32).to(torch.float32)


def run_bf(base, query, K, native):
    if native:
        os.environ.pop("GGNN_B200_NO_I8_BF", None)
    else:
        os.environ["GGNN_B200_NO_I8_BF"] = "1"
    g = ggnn.GGNN()
    g.set_return_results_on_gpu(True)
    g.set_base(base)
    ids, d = g.bf_query(query, K)
    torch.cuda.synchronize()
    t = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ids, d = g.bf_query(query, K)
        e1.record()
        torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    os.environ.pop("GGNN_B200_NO_I8_BF", None)
    return ids, d, min(t)


def gen(N, Nq, D, kind, seed):
    gcpu = torch.Generator().manual_seed(seed)
    if kind == "uniform":
        base = torch.randint(0, 256, (N, D), generator=gcpu, dtype=torch.uint8)
        query = torch.randint(0, 256, (Nq, D), generator=gcpu, dtype=torch.uint8)
    elif kind == "ties":  # values 0..3: distances collide all the time
        base = torch.randint(0, 4, (N, D), generator=gcpu, dtype=torch.uint8)
        query = torch.randint(0, 4, (Nq, D), generator=gcpu, dtype=torch.uint8)
    elif kind == "same":  # identical rows: every row ties, the candidate lists overflow -> exact scan
        base = torch.randint(0, 256, (1, D), generator=gcpu, dtype=torch.uint8).repeat(N, 1)
        query = torch.randint(0, 256, (Nq, D), generator=gcpu, dtype=torch.uint8)
    elif kind == "max":  # rows of 0 and 255: the largest distances / products
        base = (torch.randint(0, 2, (N, D), generator=gcpu, dtype=torch.uint8) * 255)
        query = (torch.randint(0, 2, (Nq, D), generator=gcpu, dtype=torch.uint8) * 255)
    else:  # SIFT-like: low-dimensional manifold + noise, clipped to bytes
        g = torch.Generator(device=DEV).manual_seed(seed)
        A = torch.randn(8, D, generator=g, device=DEV)
        zb = torch.rand(N, 8, generator=g, device=DEV)
        zq = torch.rand(Nq, 8, generator=g, device=DEV)
        f = lambda z: torch.clamp(60 + 25 * (torch.sin(3 * z) @ A) + 6 * torch.randn(z.shape[0], D, generator=g, device=DEV), 0, 255).to(torch.uint8)
        return f(zb), f(zq)
    base[min(17, N - 1)] = base[3]  # exact duplicate rows -> tie on (dist, idx)
    return base.to(DEV), query.to(DEV)


def stage_e2e():
    cases = [(5_000, 130, 128, 10, "uniform"), (1_000, 77, 32, 1, "uniform"), (20_000, 500, 64, 32, "ties"), (12_345, 257, 96, 100, "uniform"),
             (30_000, 300, 128, 128, "ties"), (4_096, 128, 128, 10, "max"), (200, 50, 128, 10, "uniform"), (20_000, 40, 64, 10, "same"), (100_000, 1000, 128, 10, "manifold"),
             (300_000, 2000, 128, 100, "manifold")]
    ok = True
    for N, Nq, D, K, kind in cases:
        base, query = gen(N, Nq, D, kind, 11)
        i1, d1, t1 = run_bf(base, query, K, True)
        i0, d0, t0 = run_bf(base, query, K, False)
        same = bool(torch.equal(i0, i1) and torch.equal(d0, d1))
        line = f"[e2e] N={N} Nq={Nq} D={D} K={K} {kind}: int8 path == widened path: {same}"
        if N * Nq <= 3e8:
            ri, rd = exact_reference(base, query, K)
            ex = bool(torch.equal(ri, i1) and torch.equal(rd, d1))
            line += f", == exact int64 reference: {ex}"
            same = same and ex
        if not same:
            line += f" (rows with equal ids {float((i0 == i1).all(1).float().mean()):.4f}, equal dists {float((d0 == d1).all(1).float().mean()):.4f})"
        print(line + f" | int8 {t1:.2f} ms, widened {t0:.2f} ms", flush=True)
        ok = ok and same
    return ok


def stage_time():
    ok = True
    for K in (10, 100):
        base, query = gen(1_000_000, 10_000, 128, "manifold", 3)
        i1, d1, t1 = run_bf(base, query, K, True)
        i0, d0, t0 = run_bf(base, query, K, False)
        os.environ["GGNN_B200_BF_I8_ROTATE"] = "0"  # every CTA of a split starts at the split's first tile
        i1n, d1n, t1n = run_bf(base, query, K, True)
        os.environ.pop("GGNN_B200_BF_I8_ROTATE")
        same = bool(torch.equal(i0, i1) and torch.equal(d0, d1) and torch.equal(i1n, i1) and torch.equal(d1n, d1))
        print(f"[time] 1M x 128 uint8, 10 000 queries, K={K}: identical={same} | int8 tensor path {t1:.2f} ms "
              f"({2 * 1e6 * 1e4 * 128 / t1 / 1e9:.0f} useful Tops/s; {t1n:.2f} ms with all CTAs of a split starting at the same tile) | "
              f"widened 3xTF32 path {t0:.2f} ms", flush=True)
        ok = ok and same
    return ok


def stage_profile():
    """one call per path and K at the 1M shape, for a kernel launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file <csv> python tools/bf_i8_check.py 5"""
    base, query = gen(1_000_000, 10_000, 128, "manifold", 3)
    os.environ["GGNN_B200_BF_DEBUG"] = "1"  # candidate statistics on stderr
    for native in (True, False):
        if not native:
            os.environ["GGNN_B200_NO_I8_BF"] = "1"
        g = ggnn.GGNN()
        g.set_return_results_on_gpu(True)
        g.set_base(base)
        for K in (10, 100):
            g.bf_query(query, K)
            torch.cuda.synchronize()
    os.environ.pop("GGNN_B200_NO_I8_BF", None)
    os.environ.pop("GGNN_B200_BF_DEBUG", None)
    return True


def main():
    stages = sys.argv[1] if len(sys.argv) > 1 else "1234"
    t0 = time.time()
    res = {}
    for key, fn in (("1", stage_pack), ("2", stage_mma), ("3", stage_e2e), ("4", stage_time), ("5", stage_profile)):
        if key in stages:
            try:
                res[fn.__name__] = fn()
            except Exception as e:  # keep going: the later stages still tell something
                print(f"[{fn.__name__}] raised {type(e).__name__}: {e}", flush=True)
                res[fn.__name__] = False
    print(f"bf_i8_check: {res} in {time.time() - t0:.1f} s", flush=True)
    sys.exit(0 if all(res.values()) else 1)


if __name__ == "__main__":
    main()
