#!/usr/bin/env python
"""GPU-box probe for the 1 -> N end-to-end curve (one node).

  torchrun --nproc-per-node N tools/multi_gpu_probe.py pcie
      every rank copies pinned host memory to / from its own GPU at the same time: the node's aggregate
      H2D / D2H / bidirectional ceiling at N concurrent GPUs (what bounds the end-to-end numbers).
  python tools/multi_gpu_probe.py group N
      ONE process drives N engines through walt_group (include/walt_b200.h): index built on device 0,
      replicated device to device (walt_engine_clone_index), then walt_group_map_se_packed over N x 10 M
      reads in pinned host memory: aggregate reads/s through the C ABI, results compared with N = 1.
Prints JSON lines."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pcie():
    import torch.distributed as dist
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    nb = 1 << 30
    h_in = torch.empty(nb, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nb, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(nb, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(nb, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps=4):
        best = 1e9
        for _ in range(reps):
            sync()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            t = torch.tensor([dt], device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = min(best, float(t.item()))
        return best

    def h2d():
        d_a.copy_(h_in, non_blocking=True)

    def d2h():
        h_out.copy_(d_b, non_blocking=True)

    def both():
        with torch.cuda.stream(s1):
            d_a.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_b, non_blocking=True)

    res = {}
    for name, fn, nbytes in (("h2d", h2d, nb), ("d2h", d2h, nb), ("bidir", both, 2 * nb)):
        t = timed(fn)
        res[name + "_gb_per_s_aggregate"] = world * nbytes / t / 1e9
        res[name + "_gb_per_s_per_gpu"] = nbytes / t / 1e9
    if rank == 0:
        print(json.dumps({"probe": "pcie", "n_gpus": world, "bytes_per_copy": nb, **res,
                          "cpus": os.cpu_count(), "how": "1 GiB pinned <-> device copies on all ranks at once, max over ranks, best of 4"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def group(n_dev):
    import bench
    import walt_b200
    from walt_b200 import engine as eng
    from walt_b200 import host as wh
    from walt_b200.engine import BEST_DT, PinnedArray
    n, rl, total = 10_000_000, 150, int(float(os.environ.get("GENOME_MB", 3100)) * 1e6)
    lengths = bench.chrom_lengths(total)
    names = [f"chr{i + 1}" for i in range(22)] + ["chrX", "chrY"]
    t0 = time.perf_counter()
    g = walt_b200.Group(list(range(n_dev)))
    e0 = g.engines[0]
    e0.set_chromosomes(lengths, names)
    torch.cuda.set_device(0)
    d_fwd = torch.empty(eng.packed_genome_bytes(total), dtype=torch.uint8, device="cuda:0")
    eng.synth_genome_device(0, total, 3, d_fwd.data_ptr())
    e0.build_from_device_genome(d_fwd.data_ptr(), which=(0, 1))
    t_build = time.perf_counter() - t0
    t0 = time.perf_counter()
    for e in g.engines[1:]:
        e.clone_index_from(e0)
    t_clone = time.perf_counter() - t0
    # N x 10 M reads in pinned host memory (every GPU's share comes from its own seed)
    h_reads = PinnedArray((n_dev * n * rl,), np.uint8)
    d_reads = torch.empty(n * rl, dtype=torch.uint8, device="cuda:0")
    for i in range(n_dev):
        e0.synth_reads_device(d_fwd.data_ptr(), n, rl, 4 + 1000 * i, False, d_reads.data_ptr())
        h_reads.array[i * n * rl:(i + 1) * n * rl] = d_reads.cpu().numpy()
    del d_fwd, d_reads
    nn = n_dev * n
    h_offs = PinnedArray((nn + 1,), np.uint64)
    h_offs.array[:] = np.arange(nn + 1, dtype=np.uint64) * np.uint64(rl)
    h_pk = PinnedArray(((nn * rl >> 2) + nn + 16,), np.uint8)
    wh.pack_reads_2bit(h_reads.array, h_offs.array, out=h_pk.array)
    h_out = PinnedArray((nn,), BEST_DT)
    out = {"probe": "group", "n_gpus": n_dev, "index_build_s": round(t_build, 2), "index_clone_s": round(t_clone, 2),
           "index_bytes_per_gpu": e0.hbm_bytes(), "clone_gb_per_s_per_gpu": e0.hbm_bytes() * max(1, n_dev - 1) / max(t_clone, 1e-9) / 1e9 / max(1, n_dev - 1)}
    for name, fn in (("packed", lambda: g.map_se_packed(h_pk.array, h_offs.array, m=6, b=5000, out=h_out.array)),
                     ("ascii", lambda: g.map_se(h_reads.array, h_offs.array, m=6, b=5000, out=h_out.array))):
        for _ in range(2):
            fn()
        t0 = time.perf_counter()
        for _ in range(3):
            fn()
        dt = (time.perf_counter() - t0) / 3
        out[f"e2e_{name}_reads_per_s"] = nn / dt
        out[f"e2e_{name}_ms"] = dt * 1e3
    # the first GPU's share must equal what one engine alone returns for it
    ref, _ = e0.map_se_packed(h_pk.array, h_offs.array[: n + 1], m=6, b=5000)
    out["first_share_identical_to_single_engine"] = bool(np.array_equal(ref, h_out.array[:n]))
    out["unique_frac"] = float((h_out.array["times"] == 1).mean())
    print(json.dumps(out), flush=True)
    for h in (h_reads, h_offs, h_pk, h_out):
        h.free()
    g.close()


if __name__ == "__main__":
    if sys.argv[1] == "pcie":
        pcie()
    else:
        group(int(sys.argv[2]))
