#!/usr/bin/env python
"""GPU-box probe: raw pinned H2D/D2H bandwidth and the end-to-end walt_engine_map_se time for
several host chunk sizes (full configs[1] workload).  Prints one JSON line per setting."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    args = bench.argparse.Namespace(gpus=1, steps=3, warmup=3, genome_mb=float(os.environ.get("GENOME_MB", 3100)),
                                    reads=10_000_000, read_len=150, group_width=8, table_depth=0)
    torch.cuda.set_device(0)
    # raw copies
    h = torch.empty(1_500_000_000, dtype=torch.uint8).pin_memory()
    d = torch.empty_like(h, device="cuda")
    for name, src, dst in (("h2d", h, d), ("d2h", d, h)):
        best = 1e9
        for _ in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        print(json.dumps({"probe": name, "gb_per_s": 1.5 / best, "ms": best * 1e3}), flush=True)
    del h, d
    wl = bench.Workload(args, 0, 0)
    from walt_b200.engine import BEST_DT, PinnedArray
    n, rl = wl.n, wl.rl
    h_reads = PinnedArray((n * rl,), np.uint8)
    h_offs = PinnedArray((n + 1,), np.uint64)
    h_out = PinnedArray((n,), BEST_DT)
    h_reads.array[:] = wl.d_reads.cpu().numpy()
    h_offs.array[:] = np.arange(n + 1, dtype=np.uint64) * np.uint64(rl)
    for chunk in [int(x) for x in os.environ.get("CHUNKS", "131072,262144,524288,1048576,2097152").split(",")]:
        wl.e.set_chunk_reads(chunk)
        for _ in range(2):
            wl.e.map_se(h_reads.array, h_offs.array, ag=False, m=6, b=5000, out=h_out.array)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3):
            wl.e.map_se(h_reads.array, h_offs.array, ag=False, m=6, b=5000, out=h_out.array)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
        print(json.dumps({"probe": "e2e", "chunk_reads": chunk, "ms": dt * 1e3, "reads_per_s": n / dt,
                          "env": {k: v for k, v in os.environ.items() if k.startswith("WALT_")}}), flush=True)
    packed_sweep(wl, h_reads, h_offs, h_out, n)


def packed_sweep(wl, h_reads, h_offs, h_out, n):
    """the same sweep for walt_engine_map_se_packed (the call the walt program makes)"""
    from walt_b200 import host as wh
    from walt_b200.engine import PinnedArray
    h_pk = PinnedArray(((n * wl.rl >> 2) + n + 16,), np.uint8)
    wh.pack_reads_2bit(h_reads.array, h_offs.array, out=h_pk.array)
    for chunk in [int(x) for x in os.environ.get("CHUNKS", "131072,262144,524288,1048576,2097152").split(",")]:
        wl.e.set_chunk_reads(chunk)
        for _ in range(2):
            wl.e.map_se_packed(h_pk.array, h_offs.array, ag=False, m=6, b=5000, out=h_out.array)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3):
            wl.e.map_se_packed(h_pk.array, h_offs.array, ag=False, m=6, b=5000, out=h_out.array)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
        print(json.dumps({"probe": "e2e_packed", "chunk_reads": chunk, "ms": dt * 1e3, "reads_per_s": n / dt}), flush=True)


if __name__ == "__main__":
    main()
