#!/usr/bin/env python
"""bench.py -- reads mapped/sec of the WALT single-end hot path (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--genome-mb 3100] [--reads 10000000] [--read-len 150]

Workload (synthetic, seeded, generated on the device; there is no network):
  * genome  : 24 chromosomes with the hg19 length profile, 3.1 Gb in total, i.i.d. uniform ACGT
  * index   : the _CT00/_CT01 sub-indexes built from it by the engine's device makedb
              (bit-identical to the reference builder's output, tests/test_gpu_parity.py)
  * reads   : 10 M single-end 150 bp directional bisulfite reads per GPU (SURVEY.md 8(d) model),
              -m 6 -b 5000
A "step" is one pass of the hot path (both strand passes of mapping.cpp:486-500) over the
batch.  `value` times the kernel with every buffer resident in HBM (CUDA events on the
launching stream); `e2e` times the C-ABI call a host program makes (walt_engine_map_se) with
pinned host buffers, host<->device copies inside the timed region.  Multi-GPU: one process per
GPU (torchrun), full index replica per GPU, reads sharded, no collective on the data path
("weak" scaling: 10 M reads per GPU); barrier + max over ranks.

`--impl reference` times the UNMODIFIED reference (oracle/_ref/libwaltref.so, the reference's
own SingleEndMapping object code under OpenMP with every host core) on a bounded sample of the
same workload; the index it maps against is exported from the device builder (set-up, not
timed) because the reference's single-threaded makedb needs ~2 h at this genome size.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HG19_MB = [249.25, 243.20, 198.02, 191.15, 180.92, 171.12, 159.14, 146.36, 141.21, 135.53, 135.01, 133.85,
           115.17, 107.35, 102.53, 90.35, 81.20, 78.08, 59.13, 63.03, 48.13, 51.30, 155.27, 59.37]
METRIC = "reads mapped/sec (SE 150bp, hg19-size synthetic)"
UNIT = "reads/s"
M, B = 6, 5000


def chrom_lengths(total_bases):
    w = np.array(HG19_MB) / np.sum(HG19_MB)
    lens = np.floor(w * total_bases).astype(np.int64)
    lens[0] += total_bases - lens.sum()
    return lens.astype(np.uint32)


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 8:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for nm, v in zip(names, r[4:8]):
                if v.strip().lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def measured_peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def algorithmic_bytes(ctr, n_reads, rl):
    """bytes a step must touch, from the oracle's deterministic work counters."""
    s = min(50, (rl - 2) // 3)
    per_probe = 4 + -(-3 * (s - 12) // 4)
    seed = 8 * ctr["n_lookups"] + 2 * ctr["sum_log2_bucket"] * per_probe
    verify = ctr["n_cand"] * (4 + -(-rl // 4))
    io = n_reads * (-(-rl // 4) + 16)
    return {"seed": seed / n_reads, "verify": verify / n_reads, "io": io / n_reads,
            "total": (seed + verify + io) / n_reads}


class Workload:
    """Per-rank synthetic genome, resident index and read batch."""

    def __init__(self, args, device, rank):
        import torch
        import walt_b200
        from walt_b200 import engine as eng
        self.torch, self.eng = torch, eng
        self.device = device
        self.n, self.rl = args.reads, args.read_len
        total = int(args.genome_mb * 1e6)
        self.lengths = chrom_lengths(total)
        self.names = [f"chr{i + 1}" for i in range(22)] + ["chrX", "chrY"]
        t0 = time.time()
        self.e = walt_b200.Engine(device)
        self.e.set_group_width(args.group_width)
        self.e.set_table_depth(args.table_depth)
        self.e.set_chromosomes(self.lengths, self.names)
        d_fwd = torch.empty(eng.packed_genome_bytes(total), dtype=torch.uint8, device=f"cuda:{device}")
        eng.synth_genome_device(device, total, 3, d_fwd.data_ptr())
        self.e.build_from_device_genome(d_fwd.data_ptr(), which=(0, 1))
        torch.cuda.synchronize()
        self.t_index = time.time() - t0
        self.d_reads = torch.empty(self.n * self.rl, dtype=torch.uint8, device=f"cuda:{device}")
        # every rank maps its own shard: different read seed per rank
        self.e.synth_reads_device(d_fwd.data_ptr(), self.n, self.rl, 4 + 1000 * rank, False, self.d_reads.data_ptr())
        del d_fwd
        torch.cuda.empty_cache()
        self.d_offs = torch.arange(self.n + 1, dtype=torch.int64, device=f"cuda:{device}") * self.rl
        self.d_out = torch.zeros(self.n * 16, dtype=torch.uint8, device=f"cuda:{device}")
        torch.cuda.synchronize()

    def host_index(self):
        """Export both sub-indexes into reference-owned Genome/HashTable objects."""
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import refio
        L = refio.ref_lib()
        out = []
        for which, strand in ((0, "+"), (1, "-")):
            info = self.e.subindex_info(which)
            h = C.c_void_p(L.waltref_index_alloc(C.c_uint32(len(self.lengths)),
                                                 self.lengths.ctypes.data_as(C.c_void_p),
                                                 C.c_uint32(info["index_size"])))
            L.waltref_index_set_strand(h, C.c_char(strand.encode()))
            got = C.c_uint32()
            self.e._check(self.e.L.walt_engine_export_subindex(
                self.e.h, C.c_int(which), C.c_void_p(L.waltref_index_sequence(h)),
                C.c_void_p(L.waltref_index_counter(h)), C.c_void_p(L.waltref_index_index(h)), C.byref(got)))
            out.append(h)
        return out

    def sample_reads(self, n):
        n = min(n, self.n)
        buf = self.d_reads[: n * self.rl].cpu().numpy()
        offs = np.arange(n + 1, dtype=np.uint64) * np.uint64(self.rl)
        return buf, offs


def reference_pass(hidx, buf, offs, threads):
    """Both strand passes of the reference's OpenMP loop (mapping.cpp:486-500); -> (seconds, BestMatch[])."""
    import refio
    L = refio.ref_lib()
    n = len(offs) - 1
    best = refio.init_best(n, M)
    t = 0.0
    for h, strand in zip(hidx, "+-"):
        t += L.waltref_time_se(h, buf.ctypes.data_as(C.c_void_p), offs.ctypes.data_as(C.c_void_p), C.c_uint32(n),
                               C.c_char(strand.encode()), C.c_int(0), C.c_uint32(B),
                               best.ctypes.data_as(C.c_void_p), C.c_int(threads))
    return t, best


def oracle_counters(wl, hidx, buf, offs):
    """Work counters of the C oracle on a sample (feeds the algorithmic-bytes model)."""
    import refio
    L = refio.ref_lib()
    Lo = refio.oracle_lib()
    starts = np.concatenate([[0], np.cumsum(wl.lengths.astype(np.uint64))]).astype(np.uint32)
    n = len(offs) - 1
    best = refio.init_best(n, M)
    ctr = refio.WoCounters()
    for h, strand in zip(hidx, "+-"):
        ix = refio.WoIndex(L.waltref_index_sequence(h), int(L.waltref_index_genome_len(h)), len(wl.lengths),
                           starts.ctypes.data, L.waltref_index_counter(h), L.waltref_index_index(h),
                           int(L.waltref_index_index_size(h)))
        Lo.wo_se_map_batch(C.byref(ix), buf.ctypes.data_as(C.c_void_p), offs.ctypes.data_as(C.c_void_p),
                           C.c_uint32(n), C.c_char(strand.encode()), C.c_int(0), C.c_uint32(B),
                           best.ctypes.data_as(C.c_void_p), C.byref(ctr))
    return ctr.asdict(), best


def config_dict(args, extra=None):
    d = {"workload": f"configs[1]: {args.genome_mb:g} Mb synthetic genome (24 chr, hg19 profile), "
                     f"{args.reads} SE {args.read_len} bp bisulfite reads per GPU, -m {M} -b {B}",
         "genome_mb": args.genome_mb, "reads_per_gpu": args.reads, "read_len": args.read_len,
         "max_mismatches": M, "bucket_limit": B, "parallelism": f"reads sharded x{args.gpus}, index replicated",
         "l2": "inputs larger than L2 (index >= 13 GB randomly gathered, 1.5 GB of reads streamed per step)"}
    if extra:
        d.update(extra)
    return d


def run_reference(args):
    rank, local, world = dist_env()
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refio
    if not refio.have_reference():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref was not built (no /root/reference at build time)"}))
        return 0
    wl = Workload(args, local, 0)
    hidx = wl.host_index()
    threads = os.cpu_count() or 1
    cal_n = min(50000, wl.n)
    buf, offs = wl.sample_reads(cal_n)
    t_cal, _ = reference_pass(hidx, buf, offs, threads)
    per_step_s = 8.0
    n = int(max(cal_n, min(wl.n, cal_n * per_step_s / max(t_cal, 1e-6))))
    buf, offs = wl.sample_reads(n)
    for _ in range(args.warmup):
        reference_pass(hidx, buf, offs, threads)
    t = 0.0
    for _ in range(args.steps):
        dt, _ = reference_pass(hidx, buf, offs, threads)
        t += dt
    value = n * args.steps / t
    sample = f"first {n} reads of rank 0's batch per step, both strand passes, OpenMP loop only"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config_dict(args, {"reference_sample_reads": n}),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def run_ours(args):
    import torch
    rank, local, world = dist_env()
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    else:
        dist = None
        torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    wl = Workload(args, local, rank)
    e, n, rl = wl.e, wl.n, wl.rl
    stream = torch.cuda.current_stream()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        e.map_se_device(wl.d_reads.data_ptr(), wl.d_offs.data_ptr(), n, rl, wl.d_out.data_ptr(), ag=False, m=M, b=B,
                        stream=stream.cuda_stream)

    # ---- side legs on rank 0 (not timed): oracle counters, parity spot check, CPU baseline ----
    cpu_baseline, alg, parity = None, None, None
    if rank == 0 and not args.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import refio
        device_step()
        torch.cuda.synchronize()
        hidx = wl.host_index() if refio.have_reference() else None
        if hidx is not None:
            ns = min(20000, n)
            buf, offs = wl.sample_reads(ns)
            ctr, obest = oracle_counters(wl, hidx, buf, offs)
            alg = algorithmic_bytes(ctr, ns, rl)
            got = wl.d_out[: ns * 16].cpu().numpy().view(refio.BEST_DT)
            bad = sum(int((got[f] != obest[f]).sum()) for f in ("genome_pos", "times", "mismatch", "strand"))
            parity = {"sample_reads": ns, "fields_differing_vs_oracle": bad,
                      "unique_frac": float((got["times"] == 1).mean())}
            if world == 1:
                threads = os.cpu_count() or 1
                cal_n = min(50000, n)
                buf, offs = wl.sample_reads(cal_n)
                t_cal, _ = reference_pass(hidx, buf, offs, threads)
                cn = int(max(cal_n, min(n, cal_n * 12.0 / max(t_cal, 1e-6))))
                buf, offs = wl.sample_reads(cn)
                t_ref, rbest = reference_pass(hidx, buf, offs, threads)
                gotc = wl.d_out[: cn * 16].cpu().numpy().view(refio.BEST_DT)
                badc = sum(int((gotc[f] != rbest[f]).sum()) for f in ("genome_pos", "times", "mismatch", "strand"))
                parity["reference_sample_reads"] = cn
                parity["fields_differing_vs_reference"] = badc
                cpu_baseline = {"value": cn / t_ref, "unit": UNIT, "cores": threads, "kind": "reference",
                                "sample": f"first {cn} reads of the batch, both strand passes, OpenMP loop of the "
                                          f"unmodified reference (oracle/_ref/libwaltref.so), {t_ref:.1f} s"}
            L = refio.ref_lib()
            for h in hidx:
                L.waltref_index_free(h)

    # ---- device-resident timing (`value`) ----
    for _ in range(args.warmup):
        device_step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record(stream)
    for i in range(args.steps):
        device_step()
        evs[i + 1].record(stream)
    barrier()
    clocks = sampler.stop()
    step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    t_dev = evs[0].elapsed_time(evs[-1]) / 1e3

    if args.no_e2e:   # kernel experiments only: not a bench line
        if rank == 0:
            print(json.dumps({"experiment": True, "kernel_ms": float(np.mean(step_ms)), "reads_per_s": n * 1e3 / float(np.mean(step_ms)),
                              "env": {k: v for k, v in os.environ.items() if k.startswith("WALT_")}}))
        return 0
    # ---- end-to-end timing through the C ABI with pinned host buffers (`e2e`) ----
    from walt_b200.engine import BEST_DT, PinnedArray
    h_reads = PinnedArray((n * rl,), np.uint8)
    h_offs = PinnedArray((n + 1,), np.uint64)
    h_out = PinnedArray((n,), BEST_DT)
    h_reads.array[:] = wl.d_reads.cpu().numpy()
    h_offs.array[:] = np.arange(n + 1, dtype=np.uint64) * np.uint64(rl)
    for _ in range(max(1, args.warmup // 2)):
        e.map_se(h_reads.array, h_offs.array, ag=False, m=M, b=B, out=h_out.array)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e.map_se(h_reads.array, h_offs.array, ag=False, m=M, b=B, out=h_out.array)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    launches_e2e = e.stats()["n_kernel_launches"]
    same = bool(np.array_equal(h_out.array[:1000].view(np.uint8), wl.d_out[:16000].cpu().numpy()))
    barrier()

    from walt_b200.sharding import max_over_ranks
    t_dev, t_e2e = max_over_ranks([t_dev, t_e2e], dist, dev)
    if rank == 0:
        total_reads = n * world * args.steps
        peak, peak_kind = measured_peak_gbs()
        kernel_s = float(np.mean(step_ms)) / 1e3
        roof = None
        if alg is not None:
            achieved = alg["total"] * n / kernel_s / 1e9
            roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": None, "peak_kind": peak_kind, "kernel": "se_map_kernel",
                    "algorithmic_bytes_per_read": alg, "kernel_ms": kernel_s * 1e3}
        line = {"metric": METRIC, "value": total_reads / t_dev, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": config_dict(args, {"index_build_s": round(wl.t_index, 1), "hbm_index_bytes": e.hbm_bytes(),
                                             "table_depth": e.subindex_info(0)["depth"],
                                             "group_width": args.group_width}),
                "clocks": clocks,
                "e2e": {"value": total_reads / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(n * rl + 8 * (n + 1)),
                        "d2h_bytes_per_step": int(16 * n), "ms_per_step": 1e3 * t_e2e / args.steps,
                        "kernel_launches_per_step": launches_e2e, "matches_device_path": same},
                "gpu_launches": args.steps,
                "roofline": roof, "cpu_baseline": cpu_baseline, "parity_check": parity}
        print(json.dumps(line))
    h_reads.free(); h_offs.free(); h_out.free()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genome-mb", type=float, default=3100.0)
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--no-cpu", action="store_true", help="skip the oracle/reference side legs")
    ap.add_argument("--no-e2e", action="store_true", help="kernel experiment: print the device timing only")
    ap.add_argument("--group-width", type=int, default=8, help="lanes that own one read (8, 16, 32)")
    ap.add_argument("--table-depth", type=int, default=0, help="prefix-table depth (0 = auto)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    sys.exit(run_reference(args) if args.impl == "reference" else run_ours(args))


if __name__ == "__main__":
    main()
