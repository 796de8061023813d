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

`--workload se_small | se_ag | pe | pe_stress` (configs[0], [2], [3], [4]) measure the other
configurations with the same contract; the default line is configs[1] and carries a `configs` block
with the device-resident timing, roofline and reference parity of the other four at full size.

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
VERIFY = {"n_families": 278, "rep_pct": 64, "div_per_mille": 2, "genome_seed": 21}   # ~4500 copies per family at 1000 Mb
METRICS = {"verify": ("candidates verified/sec by verify_kernel (SE 150bp, ~4500 candidates per seed lookup, -b 5000)", "candidates/s"),
           "se_small": ("reads mapped/sec (SE 100bp, 10 Mb synthetic genome)", "reads/s"),
           "se": ("reads mapped/sec (SE 150bp, hg19-size synthetic)", "reads/s"),
           "se_ag": ("reads mapped/sec (SE 150bp -A, hg19-size synthetic)", "reads/s"),
           "pe": ("read pairs mapped/sec (PE 2x150bp -k 50 -L 1000, hg19-size synthetic)", "pairs/s"),
           "pe_stress": ("read pairs mapped/sec (PBAT PE 2x150bp -P -m 8 -b 5000, repeat-heavy synthetic genome, "
                         "30% adaptor read-through)", "pairs/s")}
MISMATCHES = {"verify": 6, "se_small": 6, "se": 6, "se_ag": 6, "pe": 6, "pe_stress": 8}
# genome (Mb), reads (pairs) per GPU, read length of BASELINE.json's configs[0..4]
FULL_SIZE = {"verify": (1000.0, 8192, 150), "se_small": (10.0, 100_000, 100), "se": (3100.0, 10_000_000, 150), "se_ag": (3100.0, 10_000_000, 150),
             "pe": (3100.0, 5_000_000, 150), "pe_stress": (3100.0, 5_000_000, 150)}
CONFIG_NO = {"se_small": 0, "se": 1, "se_ag": 2, "pe": 3, "pe_stress": 4}
M, B, TOP_K, FRAG = 6, 5000, 50, 1000


def chrom_lengths(total_bases, kind="se"):
    if kind == "se_small":     # four chromosomes
        w = np.array([0.4, 0.3, 0.2, 0.1])
        lens = np.floor(w * total_bases).astype(np.int64)
        lens[0] += total_bases - lens.sum()
        return lens.astype(np.uint32)
    w = np.array(HG19_MB) / np.sum(HG19_MB)
    lens = np.floor(w * total_bases).astype(np.int64)
    lens[0] += total_bases - lens.sum()
    return lens.astype(np.uint32)


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def bind_to_gpu_numa_node(gpu_index):
    """Pin this rank (and so the first-touch placement of its pinned host buffers) to the CPUs NVML
    reports as local to its GPU: with several ranks per box the host<->device copies otherwise cross
    sockets.  Returns the number of CPUs bound to, or None if nothing was changed."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[gpu_index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else gpu_index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        local = {c for c in range(ncpu) if (int(words[c // 64]) >> (c % 64)) & 1}
        allowed = os.sched_getaffinity(0) & local
        if allowed and allowed != os.sched_getaffinity(0):
            os.sched_setaffinity(0, allowed)
            return len(allowed)
    except Exception:
        pass
    return None


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML every few milliseconds on a background
    thread while the timed regions run (the nvidia-smi loop of B200_PROFILING.md cannot sample a
    region that lasts tens of milliseconds)."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, gpu_index, period_s=0.004):
        self.gpu, self.period = gpu_index, period_s
        self.rows, self.stop_flag, self.thread, self.h = [], False, None, None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.rows.append((time.perf_counter(), float(sm), int(rs)))
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self.h is None:
            return
        import threading
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def mark(self):
        return time.perf_counter()

    def stop(self, windows=None):
        """windows: list of (t0, t1) perf_counter intervals that count as "timed region"."""
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.h is None or self.thread is None:
            return out
        self.stop_flag = True
        self.thread.join(timeout=2)
        rows = self.rows
        if windows:
            inside = [r for r in rows if any(a <= r[0] <= b for a, b in windows)]
            rows = inside or rows
        if not rows:
            return out
        reasons = sorted(nm for nm, bit in self.REASONS.items() if any(r[2] & bit for r in rows))
        return {"sm_mhz": float(np.median([r[1] for r in rows])), "sm_max_mhz": self.max_sm, "reasons": reasons,
                "samples": len(rows), "how": "NVML, sampled every 4 ms inside the timed regions"}


def kernel_source_hash():
    """sha256 (16 hex digits) of the CUDA sources the mapping kernels are compiled from"""
    import hashlib
    h = hashlib.sha256()
    for f in ("walt_core.cuh", "walt_engine.cu"):
        h.update(open(os.path.join(ROOT, "walt_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def committed_traffic(kind, n_reads, genome_mb):
    """DRAM bytes per step (all mapping kernels of one device-resident step) from the committed
    `ncu --set full` capture of this same command (profiles/traffic.json) -- only if that capture was
    taken on the kernel sources that are being measured now (source hash) and on the same workload."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[kind]
        if t["reads_per_step"] == n_reads and abs(t["genome_mb"] - genome_mb) < 1e-6 and t["source_sha16"] == kernel_source_hash():
            return t
    except Exception:
        pass
    return None


def measured_peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def algorithmic_bytes(ctr, n_reads, rl, pe=False):
    """bytes a step must touch, from the oracle's deterministic work counters.
    n_reads counts reads (SE) or pairs (PE: two packed reads in, one 72-byte summary out)."""
    s = min(50, (rl - 2) // 3)
    per_probe = 4 + -(-3 * (s - 12) // 4)
    seed = 8 * ctr["n_lookups"] + 2 * ctr["sum_log2_bucket"] * per_probe
    verify = ctr["n_cand"] * (4 + -(-rl // 4))
    io = n_reads * ((2 * -(-rl // 4) + 72) if pe else (-(-rl // 4) + 16))
    return {"seed": seed / n_reads, "verify": verify / n_reads, "io": io / n_reads,
            "total": (seed + verify + io) / n_reads}


HOST_INDEX_CACHE = {}


def free_host_indexes(hidx=None):
    """free exported reference indexes: the given ones unless they are cached, or (None) the whole cache"""
    import refio
    L = refio.ref_lib()
    if hidx is None:
        for h in HOST_INDEX_CACHE.values():
            L.waltref_index_free(h)
        HOST_INDEX_CACHE.clear()
        return
    cached = {int(h.value) for h in HOST_INDEX_CACHE.values()}
    for h in hidx.values():
        if int(h.value) not in cached:
            L.waltref_index_free(h)


class Workload:
    """Per-rank synthetic genome, resident index and read batch."""

    def __init__(self, args, device, rank, kind=None, genome_mb=None, reads=None, read_len=None, verify=None):
        import torch
        import walt_b200
        from walt_b200 import engine as eng
        self.torch, self.eng = torch, eng
        self.device = device
        self.kind = kind or getattr(args, "workload", "se")
        self.genome_mb = float(genome_mb if genome_mb is not None else args.genome_mb)
        self.verify = dict(VERIFY, **(verify or {}))
        self.ag = self.kind == "se_ag"
        self.is_pe = self.kind in ("pe", "pe_stress")
        self.pbat = self.kind == "pe_stress"
        self.m = MISMATCHES[self.kind]
        self.which = {"verify": (0, 1), "se_small": (0, 1), "se": (0, 1), "se_ag": (2, 3), "pe": (0, 1, 2, 3), "pe_stress": (0, 1, 2, 3)}[self.kind]
        self.n, self.rl = int(reads if reads is not None else args.reads), int(read_len if read_len is not None else args.read_len)
        total = int(self.genome_mb * 1e6)
        self.lengths = chrom_lengths(total, self.kind)
        self.names = [f"chr{i + 1}" for i in range(len(self.lengths))] if self.kind == "se_small" else \
            [f"chr{i + 1}" for i in range(22)] + ["chrX", "chrY"]
        t0 = time.time()
        self.e = walt_b200.Engine(device)
        self.e.set_group_width(args.group_width)
        self.e.set_table_depth(args.table_depth)
        self.e.set_chromosomes(self.lengths, self.names)
        dev = f"cuda:{device}"
        d_fwd = torch.empty(eng.packed_genome_bytes(total), dtype=torch.uint8, device=dev)
        if self.kind == "verify":
            eng.synth_verify_genome_device(device, total, self.verify["genome_seed"], self.verify["n_families"], self.verify["rep_pct"],
                                           self.verify["div_per_mille"], d_fwd.data_ptr())
        else:
            eng.synth_genome_device(device, total, 3, d_fwd.data_ptr(), repeats=self.kind == "pe_stress")
        self.e.build_from_device_genome(d_fwd.data_ptr(), which=self.which)
        torch.cuda.synchronize()
        self.t_index = time.time() - t0
        self.build_info = self.e.last_build_info()   # of the last sub-index built
        self.d_reads = torch.empty(self.n * self.rl, dtype=torch.uint8, device=dev)
        self.d_reads2 = None
        # every rank maps its own shard: different read seed per rank
        if self.is_pe:
            self.d_reads2 = torch.empty(self.n * self.rl, dtype=torch.uint8, device=dev)
            # d_reads / d_reads2 are the first / second read FILE; under PBAT the first file holds the
            # A-rich mate, i.e. the directional library with its mates exchanged
            t_rich, a_rich = (self.d_reads2, self.d_reads) if self.pbat else (self.d_reads, self.d_reads2)
            self.e.synth_pairs_device(d_fwd.data_ptr(), self.n, self.rl, 5 + 1000 * rank, t_rich.data_ptr(),
                                      a_rich.data_ptr(), readthrough_pct=30 if self.kind == "pe_stress" else 0)
        elif self.kind == "verify":
            self.e.synth_verify_reads_device(d_fwd.data_ptr(), self.n, self.rl, 4 + 1000 * rank, self.verify["genome_seed"],
                                             self.verify["n_families"], self.verify["rep_pct"], self.d_reads.data_ptr())
        else:
            self.e.synth_reads_device(d_fwd.data_ptr(), self.n, self.rl, 4 + 1000 * rank, self.ag, self.d_reads.data_ptr())
        del d_fwd
        torch.cuda.empty_cache()
        self.d_offs = torch.arange(self.n + 1, dtype=torch.int64, device=dev) * self.rl
        out_bytes = eng.PE_RESULT_DT.itemsize if self.is_pe else 16
        self.out_dt = eng.PE_RESULT_DT if self.is_pe else eng.BEST_DT
        self.d_out = torch.zeros(self.n * out_bytes, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()

    def device_step(self, stream):
        if self.is_pe:
            self.e.map_pe_device(self.d_reads.data_ptr(), self.d_offs.data_ptr(), self.d_reads2.data_ptr(),
                                 self.d_offs.data_ptr(), self.n, self.rl, self.d_out.data_ptr(), m=self.m, b=B, top_k=TOP_K,
                                 frag_range=FRAG, pbat=self.pbat, stream=stream)
        else:
            self.e.map_se_device(self.d_reads.data_ptr(), self.d_offs.data_ptr(), self.n, self.rl, self.d_out.data_ptr(),
                                 ag=self.ag, m=self.m, b=B, stream=stream)

    def launches_per_step(self):
        return self.e.stats()["n_kernel_launches"]   # launches of the last device call

    def close(self):
        self.e.close()
        self.d_reads = self.d_reads2 = self.d_offs = self.d_out = None
        self.torch.cuda.empty_cache()

    def host_index(self):
        """Export the resident sub-indexes into reference-owned Genome/HashTable objects.  The uniform
        genome of configs[1..3] is the same (same seed, same size, deterministic builder), so its exports
        are kept for the next configuration (HOST_INDEX_CACHE; freed by free_host_indexes)."""
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import refio
        L = refio.ref_lib()
        out = {}
        shared = self.kind in ("se", "se_ag", "pe")
        for which in self.which:
            key = (self.genome_mb, which)
            if shared and key in HOST_INDEX_CACHE:
                out[which] = HOST_INDEX_CACHE[key]
                continue
            info = self.e.subindex_info(which)
            h = C.c_void_p(L.waltref_index_alloc(C.c_uint32(len(self.lengths)),
                                                 self.lengths.ctypes.data_as(C.c_void_p),
                                                 C.c_uint32(info["index_size"])))
            L.waltref_index_set_strand(h, C.c_char(b"-" if which & 1 else b"+"))
            got = C.c_uint32()
            self.e._check(self.e.L.walt_engine_export_subindex(
                self.e.h, C.c_int(which), C.c_void_p(L.waltref_index_sequence(h)),
                C.c_void_p(L.waltref_index_counter(h)), C.c_void_p(L.waltref_index_index(h)), C.byref(got)))
            out[which] = h
            if shared:
                HOST_INDEX_CACHE[(self.genome_mb, which)] = h
        return out

    def sample_reads(self, n, mate=1):
        """first n reads; for pairs `mate` is the bisulfite ROLE (1 = C->T mate, 2 = G->A mate)"""
        n = min(n, self.n)
        if self.is_pe and self.pbat:
            mate = 3 - mate
        src = self.d_reads if mate == 1 else self.d_reads2
        buf = src[: n * self.rl].cpu().numpy()
        offs = np.arange(n + 1, dtype=np.uint64) * np.uint64(self.rl)
        return buf, offs


def reference_pass(wl, hidx, n, threads):
    """The reference's OpenMP loops over the first n reads (pairs) of the batch: both strand passes
    of mapping.cpp:486-500, or the 2 mates x 2 strands of paired.cpp:642-672.
    -> (seconds inside the loops, results: BestMatch[] | {mate: (ranked, sizes)})."""
    import refio
    L = refio.ref_lib()
    if not wl.is_pe:
        buf, offs = wl.sample_reads(n)
        best = refio.init_best(n, wl.m)
        t = 0.0
        for which, strand in zip(wl.which, "+-"):
            t += L.waltref_time_se(hidx[which], buf.ctypes.data_as(C.c_void_p), offs.ctypes.data_as(C.c_void_p),
                                   C.c_uint32(n), C.c_char(strand.encode()), C.c_int(int(wl.ag)), C.c_uint32(B),
                                   best.ctypes.data_as(C.c_void_p), C.c_int(threads))
        return t, best
    t, res = 0.0, {}
    for mate, ag, pair in ((1, 0, (0, 1)), (2, 1, (2, 3))):
        buf, offs = wl.sample_reads(n, mate)
        hp = C.c_void_p(L.waltref_heaps_alloc(C.c_uint32(n), C.c_uint32(TOP_K)))
        for which, strand in zip(pair, "+-"):
            t += L.waltref_time_pe(hidx[which], hp, buf.ctypes.data_as(C.c_void_p), offs.ctypes.data_as(C.c_void_p),
                                   C.c_uint32(n), C.c_char(strand.encode()), C.c_int(ag), C.c_uint32(wl.m), C.c_uint32(B),
                                   C.c_int(threads))
        ranked = np.zeros((n, TOP_K), dtype=refio.CAND_DT)
        sizes = np.zeros(n, dtype=np.uint32)
        L.waltref_heaps_drain(hp, C.c_uint32(n), C.c_uint32(TOP_K), ranked.ctypes.data_as(C.c_void_p),
                              sizes.ctypes.data_as(C.c_void_p))
        L.waltref_heaps_free(hp)
        res[mate] = (ranked, sizes)
    return t, res


def oracle_counters(wl, hidx, n):
    """Work counters of the C oracle on the first n reads (pairs): feeds the algorithmic-bytes model."""
    import refio
    L = refio.ref_lib()
    Lo = refio.oracle_lib()
    starts = np.concatenate([[0], np.cumsum(wl.lengths.astype(np.uint64))]).astype(np.uint32)
    ctr = refio.WoCounters()

    def view(which):
        h = hidx[which]
        return refio.WoIndex(L.waltref_index_sequence(h), int(L.waltref_index_genome_len(h)), len(wl.lengths),
                             starts.ctypes.data, L.waltref_index_counter(h), L.waltref_index_index(h),
                             int(L.waltref_index_index_size(h)))
    if not wl.is_pe:
        buf, offs = wl.sample_reads(n)
        best = refio.init_best(n, wl.m)
        for which, strand in zip(wl.which, "+-"):
            ix = view(which)
            Lo.wo_se_map_batch(C.byref(ix), buf.ctypes.data_as(C.c_void_p), offs.ctypes.data_as(C.c_void_p),
                               C.c_uint32(n), C.c_char(strand.encode()), C.c_int(int(wl.ag)), C.c_uint32(B),
                               best.ctypes.data_as(C.c_void_p), C.byref(ctr))
        return ctr.asdict(), best
    for mate, ag, pair in ((1, 0, (0, 1)), (2, 1, (2, 3))):
        buf, offs = wl.sample_reads(n, mate)
        cands = np.zeros((n, TOP_K), dtype=refio.CAND_DT)
        sizes = np.zeros(n, dtype=np.uint32)
        for which, strand in zip(pair, "+-"):
            ix = view(which)
            Lo.wo_pe_map_batch(C.byref(ix), buf.ctypes.data_as(C.c_void_p), offs.ctypes.data_as(C.c_void_p),
                               C.c_uint32(n), C.c_char(strand.encode()), C.c_int(ag), C.c_uint32(wl.m), C.c_uint32(B),
                               C.c_uint32(TOP_K), cands.ctypes.data_as(C.c_void_p), sizes.ctypes.data_as(C.c_void_p),
                               C.byref(ctr))
    return ctr.asdict(), None


def parity_vs_reference(wl, n, ref_result):
    """Fields of the results the TIMED device-resident call left in wl.d_out that differ from the
    unmodified reference's on the first n reads (pairs).  SE: BestMatch.  PE: the walt_pe_result of
    walt_engine_map_pe_device against what MergePairedEndResults (paired.cpp:438-570; oracle
    restatement wo_pe_result_batch) derives from the drained heaps of the reference's
    PairEndMapping (libwaltref.so): pairing result, winning candidates, GetBestMatch4Single."""
    import refio
    if not wl.is_pe:
        got = wl.d_out[: n * 16].cpu().numpy().view(refio.BEST_DT)
        return sum(int((got[f] != ref_result[f]).sum()) for f in ("genome_pos", "times", "mismatch", "strand"))
    Lo = refio.oracle_lib()
    starts = np.concatenate([[0], np.cumsum(wl.lengths.astype(np.uint64))]).astype(np.uint32)
    lengths = np.ascontiguousarray(wl.lengths, np.uint32)
    chroms = refio.WoChroms(len(lengths), starts.ctypes.data, lengths.ctypes.data)
    offs = np.arange(n + 1, dtype=np.uint64) * np.uint64(wl.rl)
    (r1, s1), (r2, s2) = ref_result[1], ref_result[2]          # bisulfite roles: 1 = C->T mate, 2 = G->A mate
    want = np.zeros(n, dtype=wl.out_dt)
    Lo.wo_pe_result_batch(C.byref(chroms), r1.ctypes.data_as(C.c_void_p), s1.ctypes.data_as(C.c_void_p),
                          offs.ctypes.data_as(C.c_void_p), r2.ctypes.data_as(C.c_void_p), s2.ctypes.data_as(C.c_void_p),
                          offs.ctypes.data_as(C.c_void_p), C.c_uint32(n), C.c_uint32(TOP_K), C.c_uint32(wl.m), C.c_int(FRAG),
                          want.ctypes.data_as(C.c_void_p))
    if wl.pbat:   # derived oracle: the reference on exchanged mates, per-mate fields handed back
        sw = want.copy()
        sw["pair"]["best_i"], sw["pair"]["best_j"] = want["pair"]["best_j"], want["pair"]["best_i"]
        sw["c1"], sw["c2"], sw["single1"], sw["single2"] = want["c2"], want["c1"], want["single2"], want["single1"]
        want = sw
    got = wl.d_out[: n * wl.out_dt.itemsize].cpu().numpy().view(wl.out_dt)
    bad = 0
    for f in ("best_times", "best_i", "best_j", "frag_len"):
        bad += int((got["pair"][f] != want["pair"][f]).sum())
    for rec, fields in (("c1", ("genome_pos", "mismatch", "strand")), ("c2", ("genome_pos", "mismatch", "strand")),
                        ("single1", ("genome_pos", "times", "mismatch", "strand")),
                        ("single2", ("genome_pos", "times", "mismatch", "strand"))):
        for f in fields:
            bad += int((got[rec][f] != want[rec][f]).sum())
    return bad


def workload_text(kind, genome_mb, reads, read_len):
    return {"verify": f"verification micro-benchmark: {genome_mb:g} Mb synthetic genome, {VERIFY['rep_pct']} % of its 512-base tiles "
                      f"copies of {VERIFY['n_families']} families of 400 bases ({VERIFY['div_per_mille']} substitutions per 1000 bases), "
                      f"{reads} SE {read_len} bp bisulfite reads from inside copies per step, -m {M} -b {B}",
            "se_small": f"configs[0]: {genome_mb:g} Mb synthetic genome (4 chr), {reads} SE {read_len} bp bisulfite reads per GPU, "
                        f"-m {M} -b {B}",
            "se": f"configs[1]: {genome_mb:g} Mb synthetic genome (24 chr, hg19 profile), {reads} SE "
                  f"{read_len} bp bisulfite reads per GPU, -m {M} -b {B}",
            "se_ag": f"configs[2]: {genome_mb:g} Mb synthetic genome (24 chr, hg19 profile), {reads} A-rich SE "
                     f"{read_len} bp reads per GPU, -A -m {M} -b {B}",
            "pe": f"configs[3]: {genome_mb:g} Mb synthetic genome (24 chr, hg19 profile), {reads} pairs "
                  f"2x{read_len} bp per GPU, -m {M} -b {B} -k {TOP_K} -L {FRAG}",
            "pe_stress": f"configs[4]: {genome_mb:g} Mb repeat-heavy synthetic genome (2000 repeat families, ~40 % of "
                         f"the bases), {reads} PBAT pairs 2x{read_len} bp per GPU, 30 % adaptor read-through "
                         f"(as clipped by -C), -P -m 8 -b {B} -k {TOP_K} -L {FRAG}"}[kind]


def config_dict(args, extra=None):
    d = {"workload": workload_text(args.workload, args.genome_mb, args.reads, args.read_len), "genome_mb": args.genome_mb,
         "reads_per_gpu": args.reads, "read_len": args.read_len,
         "max_mismatches": MISMATCHES[args.workload], "bucket_limit": B, "parallelism": f"reads sharded x{args.gpus}, index replicated",
         "l2": "inputs larger than L2 (index >= 13 GB per strand randomly gathered, >= 1.5 GB of reads streamed per step)"
               if args.genome_mb >= 500 else "L2 flushed between timed steps (a 256 MB buffer is rewritten)"}
    if extra:
        d.update(extra)
    return d


# ---------------------------------------------------------------------------------------------
# files in -> files out: the `walt` program against the reference `walt` on the same inputs
# ---------------------------------------------------------------------------------------------
def write_fastq_fixed(path, seqs, n, rl):
    """FASTQ with fixed-width records (vectorised): @r%09d / seq / + / qualities."""
    rec = 11 + 1 + rl + 3 + rl + 1
    step = 250_000
    digits = 10 ** np.arange(8, -1, -1, dtype=np.int64)
    with open(path, "wb") as f:
        for i0 in range(0, n, step):
            m = min(step, n - i0)
            a = np.empty((m, rec), np.uint8)
            a[:, 0] = ord("@"); a[:, 1] = ord("r")
            idx = np.arange(i0, i0 + m, dtype=np.int64)
            a[:, 2:11] = (idx[:, None] // digits[None, :] % 10 + 48).astype(np.uint8)
            a[:, 11] = 10
            a[:, 12:12 + rl] = seqs[i0 * rl:(i0 + m) * rl].reshape(m, rl)
            a[:, 12 + rl] = 10; a[:, 13 + rl] = ord("+"); a[:, 14 + rl] = 10
            a[:, 15 + rl:15 + 2 * rl] = (35 + (idx[:, None] * 7 + np.arange(rl)[None, :] * 13) % 38).astype(np.uint8)
            a[:, 15 + 2 * rl] = 10
            f.write(a.tobytes())


def cli_leg(device, genome_mb, n_reads, rl, workdir=None, keep=False, small=False):
    """makedb-equivalent on the device -> .dbindex files; synthetic FASTQ; then `walt_b200/bin/walt`
    and the unmodified reference `oracle/_ref/walt -t <cores>` on the same files, wall clock of each
    whole process (index load, FASTQ parse, mapping, SAM formatting), outputs compared byte for byte."""
    import hashlib
    import shutil
    import subprocess
    import tempfile
    import torch
    import walt_b200
    from walt_b200 import engine as eng
    from walt_b200 import host as wh
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refio
    total = int(genome_mb * 1e6)
    lengths = chrom_lengths(total, "se_small" if small else "se")
    names = [f"chr{i + 1}" for i in range(len(lengths))] if small else [f"chr{i + 1}" for i in range(22)] + ["chrX", "chrY"]
    work = tempfile.mkdtemp(prefix="walt_cli_", dir=workdir or os.environ.get("WALT_BENCH_TMP") or None)
    out = {"genome_mb": genome_mb, "reads": n_reads, "read_len": rl}
    try:
        t0 = time.time()
        e = walt_b200.Engine(device)
        e.set_chromosomes(lengths, names)
        dev = f"cuda:{device}"
        d_fwd = torch.empty(eng.packed_genome_bytes(total), dtype=torch.uint8, device=dev)
        eng.synth_genome_device(device, total, 7, d_fwd.data_ptr())
        idx = os.path.join(work, "g.dbindex")
        chroms = wh.Chroms(names=names, lengths=lengths)
        size_of_index = 0
        for which, sfx in enumerate(("_CT00", "_CT01", "_GA10", "_GA11")):
            e.build_from_device_genome(d_fwd.data_ptr(), which=(which,))
            seq, counter, index = e.export_subindex(which, total)
            wh.write_subindex(idx + sfx, "-" if which & 1 else "+", seq, counter, index)
            size_of_index = max(size_of_index, int(index.size))
        wh.write_dbindex_header(idx, chroms, size_of_index)
        d_reads = torch.empty(n_reads * rl, dtype=torch.uint8, device=dev)
        e.synth_reads_device(d_fwd.data_ptr(), n_reads, rl, 11, False, d_reads.data_ptr())
        fq = os.path.join(work, "reads.fastq")
        write_fastq_fixed(fq, d_reads.cpu().numpy(), n_reads, rl)
        del d_fwd, d_reads
        e.close()
        torch.cuda.empty_cache()
        if os.environ.get("WALT_CLI_SYNC", "0") != "0":
            os.sync()     # experiment: set-up wrote ~10 GB; wait for the write-back before any program is timed on those
                          # files (measured: no help -- profiles/r02_bench_cli_startup_split.json)
        out["setup_s"] = round(time.time() - t0, 1)
        out["fastq_bytes"] = os.path.getsize(fq)
        out["index_bytes"] = sum(os.path.getsize(idx + s) for s in ("", "_CT00", "_CT01", "_GA10", "_GA11"))
        cores = os.cpu_count() or 1
        opts = ["-i", idx, "-r", fq, "-sam", "-u", "-a", "-m", str(M), "-b", str(B)]
        n_gpus = int(os.environ.get("WALT_CLI_GPUS", "1"))      # walt -gpus N: index read once, cloned device to device
        out["gpus"] = n_gpus

        def timed(cmd, env=None):
            t = time.perf_counter()
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
            dt = time.perf_counter() - t
            if r.returncode != 0:
                raise RuntimeError(f"{cmd[0]} failed: {r.stderr.decode()[-500:]}")
            return dt, r.stderr.decode()

        def digest(path):
            h = hashlib.sha256()
            with open(path, "rb") as f:
                for blk in iter(lambda: f.read(1 << 24), b""):
                    h.update(blk)
            return h.hexdigest(), os.path.getsize(path)

        ours_bin = os.path.join(ROOT, "walt_b200", "bin", "walt")
        env = dict(os.environ, WALT_TIMING="1")
        if n_gpus == 1:
            env["CUDA_VISIBLE_DEVICES"] = str(device)
        o_out = os.path.join(work, "ours.sam")
        runs = []
        for _ in range(2):   # the second run has every input in the page cache (as has the reference's)
            dt, err = timed([ours_bin] + opts + ["-o", o_out] + (["-gpus", str(n_gpus)] if n_gpus > 1 else []), env)
            runs.append(dt)
            stages = [l for l in err.splitlines() if l.startswith("[walt timing]")]
        out["ours_s"] = min(runs)
        out["ours_runs_s"] = [round(x, 3) for x in runs]
        out["ours_stages"] = stages if stages else None
        out["ours_reads_per_s"] = n_reads / min(runs)
        d_ours = digest(o_out)
        ms_ours = open(o_out + ".mapstats", "rb").read()
        # experiments on the same files: WALT_CLI_VARIANTS="NAME=VALUE,NAME=VALUE;..." -> one more run per setting
        for spec in filter(None, os.environ.get("WALT_CLI_VARIANTS", "").split(";")):
            v_env = dict(env, **dict(kv.split("=", 1) for kv in spec.split(",")))
            v_out = os.path.join(work, "variant.sam")
            dt, err = timed([ours_bin] + opts + ["-o", v_out] + (["-gpus", str(n_gpus)] if n_gpus > 1 else []), v_env)
            out.setdefault("variants", []).append({"env": spec, "s": round(dt, 3), "identical": digest(v_out) == d_ours,
                                                   "stages": [l for l in err.splitlines() if l.startswith("[walt timing]")]})
            os.remove(v_out)
        out["sam_bytes"] = d_ours[1]
        if refio.have_reference():
            r_out = os.path.join(work, "ref.sam")
            ref_bin = os.path.join(ROOT, "oracle", "_ref", "walt")
            dt, _ = timed([ref_bin] + opts + ["-o", r_out, "-t", str(cores)])
            out["reference_s"] = dt
            out["reference_reads_per_s"] = n_reads / dt
            out["reference_threads"] = cores
            out["outputs_identical"] = bool(digest(r_out) == d_ours and open(r_out + ".mapstats", "rb").read() == ms_ours)
            out["speedup"] = dt / min(runs)
        return out
    finally:
        if not keep:
            shutil.rmtree(work, ignore_errors=True)


def makedb_leg(device, genome_mb, repeats=False):
    """`walt_b200/bin/makedb` against the unmodified `oracle/_ref/makedb` on the same FASTA file:
    whole-process wall clock of each, the five index files compared byte for byte."""
    import shutil
    import subprocess
    import tempfile
    import torch
    from walt_b200 import engine as eng
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refio
    import synth
    total = int(genome_mb * 1e6)
    lengths = chrom_lengths(total)
    work = tempfile.mkdtemp(prefix="walt_makedb_", dir=os.environ.get("WALT_BENCH_TMP") or None)
    out = {"genome_mb": genome_mb, "repeat_heavy": bool(repeats)}
    try:
        # the genome comes from the device generator so that it matches the other workloads' model
        d = torch.empty(eng.packed_genome_bytes(total), dtype=torch.uint8, device=f"cuda:{device}")
        eng.synth_genome_device(device, total, 9, d.data_ptr(), repeats=repeats)
        words = d.cpu().numpy().view(np.uint64)[1:]      # word 0 is the pad; first base in the top bits
        codes = np.zeros(words.size * 32, np.uint8)
        w = words.astype(np.uint64)
        for i in range(32):
            codes[i::32] = ((w >> np.uint64(62 - 2 * i)) & np.uint64(3)).astype(np.uint8)
        seq = np.frombuffer(b"ACGT", np.uint8)[codes[:total]]
        del d, w, codes
        torch.cuda.empty_cache()
        chroms, at = [], 0
        for i, ln in enumerate(lengths):
            chroms.append((f"chr{i + 1}", seq[at:at + int(ln)]))
            at += int(ln)
        fa = os.path.join(work, "genome.fa")
        synth.write_fasta(fa, chroms)
        out["fasta_bytes"] = os.path.getsize(fa)
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=str(device))
        ours = os.path.join(work, "ours.dbindex")
        t = time.perf_counter()
        r = subprocess.run([os.path.join(ROOT, "walt_b200", "bin", "makedb"), "-c", fa, "-o", ours],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
        out["ours_s"] = time.perf_counter() - t
        if r.returncode != 0:
            raise RuntimeError("makedb failed: " + r.stderr.decode()[-400:])
        out["index_bytes"] = sum(os.path.getsize(ours + sfx) for sfx in ("", "_CT00", "_CT01", "_GA10", "_GA11"))
        if refio.have_reference():
            ref = os.path.join(work, "ref.dbindex")
            t = time.perf_counter()
            refio.ref_makedb(fa, ref)
            out["reference_s"] = time.perf_counter() - t
            out["speedup"] = out["reference_s"] / out["ours_s"]
            same = True
            for sfx in ("", "_CT00", "_CT01", "_GA10", "_GA11"):
                a, b = ours + sfx, ref + sfx
                same = same and os.path.getsize(a) == os.path.getsize(b) and \
                    subprocess.run(["cmp", "-s", a, b]).returncode == 0
            out["files_identical"] = bool(same)
        return out
    finally:
        shutil.rmtree(work, ignore_errors=True)


def run_cli(args):
    rank, local, world = dist_env()
    if rank != 0:
        return 0
    r = cli_leg(local, args.cli_genome_mb, args.cli_reads, args.read_len)
    mk = None
    if args.makedb_genome_mb > 0:
        try:
            mk = makedb_leg(local, args.makedb_genome_mb, repeats=args.makedb_repeats)
        except Exception as ex:
            mk = {"error": str(ex)[-300:]}
    line = {"metric": "reads mapped/sec, walt program: .dbindex + FASTQ files in, SAM file out (process wall clock)",
            "value": r["ours_reads_per_s"], "unit": "reads/s", "n_gpus": 1, "higher_is_better": True, "data": "synthetic",
            "config": {"workload": f"walt -i <{r['genome_mb']:g} Mb index> -r <{r['reads']} SE {r['read_len']} bp reads> -sam -u -a "
                                   f"-m {M} -b {B}; reference: the unmodified walt -t <cores> on the same files"},
            "cli": r, "makedb": mk}
    emit(json.dumps(line))
    return 0


def run_verify(args):
    import torch
    rank, local, world = dist_env()
    if rank != 0:
        return 0
    torch.cuda.set_device(local)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    stream = torch.cuda.current_stream()
    emit(json.dumps(verify_leg(args, local, stream, torch.cuda.synchronize)))
    return 0


def run_reference(args):
    rank, local, world = dist_env()
    if rank != 0:
        return 0
    metric, unit = METRICS[args.workload]
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refio
    if not refio.have_reference():
        emit(json.dumps({"impl": "reference", "unavailable": "oracle/_ref was not built (no /root/reference at build time)"}))
        return 0
    wl = Workload(args, local, 0)
    hidx = wl.host_index()
    threads = os.cpu_count() or 1
    cal_n = min(50000, wl.n)
    t_cal, _ = reference_pass(wl, hidx, cal_n, threads)
    per_step_s = 8.0
    n = int(max(cal_n, min(wl.n, cal_n * per_step_s / max(t_cal, 1e-6))))
    for _ in range(args.warmup):
        reference_pass(wl, hidx, n, threads)
    t = 0.0
    for _ in range(args.steps):
        dt, _ = reference_pass(wl, hidx, n, threads)
        t += dt
    value = n * args.steps / t
    sample = (f"first {n} {'pairs' if wl.is_pe else 'reads'} of rank 0's batch per step, every strand pass, "
              f"OpenMP loops only")
    line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config_dict(args, {"reference_sample_reads": n}),
            "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(json.dumps(line))
    return 0


def time_device_steps(wl, steps, warmup, stream, barrier, flush_l2=False):
    """W untimed steps, then K timed ones (CUDA events on the launching stream).  flush_l2: the index of
    the workload fits the L2, so a 256 MB buffer is rewritten between the steps and every step is timed
    by its own pair of events; otherwise one pair of events brackets the K steps back to back."""
    import torch
    for _ in range(warmup):
        wl.device_step(stream.cuda_stream)
    barrier()
    t0 = time.perf_counter()
    if flush_l2:
        junk = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{wl.device}")
        step_ms = []
        for _ in range(steps):
            junk.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            wl.device_step(stream.cuda_stream)
            b.record(stream)
            torch.cuda.synchronize()
            step_ms.append(a.elapsed_time(b))
        barrier()
        return step_ms, sum(step_ms) / 1e3, (t0, time.perf_counter())
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    evs[0].record(stream)
    for i in range(steps):
        wl.device_step(stream.cuda_stream)
        evs[i + 1].record(stream)
    barrier()
    step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
    return step_ms, evs[0].elapsed_time(evs[-1]) / 1e3, (t0, time.perf_counter())


def own_floor_bytes(stats, n, rl, pe):
    """Sector-level minimum of the table/fingerprint design for the work the engine did (its own lookup
    and candidate counters): per lookup one 32-byte sector of the prefix table and one of the entry
    array, per candidate the two sectors its 150-base window spans plus its 8-byte entry, per read the
    read itself and its result."""
    io = (2 * rl + 72) if pe else (rl + 16)
    return (64.0 * stats["n_lookups"] + 72.0 * stats["n_candidates"]) / n + io


def side_legs(wl, unit, ref_seconds, want_baseline):
    """Rank 0, not timed: the C oracle's work counters (algorithmic bytes), the parity of the results
    the timed device-resident step left on the device against the oracle and the unmodified reference,
    and the reference's OpenMP loops on a bounded sample (the CPU baseline)."""
    import refio
    out = {"alg": None, "parity": None, "cpu_baseline": None}
    if not refio.have_reference():
        return out
    n, pe = wl.n, wl.is_pe
    hidx = wl.host_index()
    try:
        ns = min(20000, n)
        ctr, obest = oracle_counters(wl, hidx, ns)
        out["alg"] = algorithmic_bytes(ctr, ns, wl.rl, pe)
        out["alg"]["candidates_per_read"] = ctr["n_cand"] / ns
        parity = {"sample_reads": ns}
        if not pe:
            got = wl.d_out[: ns * 16].cpu().numpy().view(refio.BEST_DT)
            parity["fields_differing_vs_oracle"] = sum(int((got[f] != obest[f]).sum())
                                                       for f in ("genome_pos", "times", "mismatch", "strand"))
            parity["unique_frac"] = float((got["times"] == 1).mean())
        else:
            got = wl.d_out[: ns * wl.out_dt.itemsize].cpu().numpy().view(wl.out_dt)
            parity["unique_pair_frac"] = float((got["pair"]["best_times"] == 1).mean())
        threads = os.cpu_count() or 1
        cal_n = min(50000, n)
        t_cal, _ = reference_pass(wl, hidx, cal_n, threads)
        cn = int(max(cal_n, min(n, cal_n * ref_seconds / max(t_cal, 1e-6))))
        if pe:
            cn = max(cn, min(n, 1_000_000))     # the timed paired-end path is checked on at least a million pairs
        t_ref, rres = reference_pass(wl, hidx, cn, threads)
        parity["reference_sample_reads"] = cn
        parity["fields_differing_vs_reference"] = parity_vs_reference(wl, cn, rres)
        parity["what"] = ("walt_pe_result of the timed walt_engine_map_pe_device call vs MergePairedEndResults on the "
                          "reference's drained heaps" if pe else "BestMatch of the timed walt_engine_map_se_device call vs the "
                          "reference's SingleEndMapping")
        out["parity"] = parity
        base = {"value": cn / t_ref, "unit": unit, "cores": threads, "kind": "reference",
                "sample": f"first {cn} {'pairs' if pe else 'reads'} of the batch, every strand pass, "
                          f"OpenMP loops of the unmodified reference (oracle/_ref/libwaltref.so), {t_ref:.1f} s"}
        out["cpu_baseline"] = base if want_baseline else {"value": base["value"], "cores": threads}
    finally:
        free_host_indexes(hidx)
    return out


def roofline_dict(wl, alg, kernel_s, dstats, genome_mb):
    """HBM roofline of one device-resident step.  `frac` follows SURVEY.md 8(d): the ALGORITHMIC bytes of
    the reference's method (its ~110 binary-search probes per lookup) over the step's device time -- a
    reference-equivalent throughput, not a statement about the memory system.  `dram_frac` is the
    honest utilisation: DRAM bytes of the committed same-source ncu capture over the same time;
    `own_floor_bytes` is the sector-level minimum of what this design has to touch."""
    peak, peak_kind = measured_peak_gbs()
    n, pe = wl.n, wl.is_pe
    achieved = alg["total"] * n / kernel_s / 1e9
    t = committed_traffic(wl.kind, n, genome_mb)
    floor = own_floor_bytes(dstats, n, wl.rl, pe) if dstats else None
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": t["dram_bytes_per_step"] if t else None, "peak_kind": peak_kind,
            "kernel": "pe_log_kernel (park / take-over) + pe_heap_kernel + pair_kernel" if pe else "se_map_kernel (park / take-over)",
            "kernel_ms": kernel_s * 1e3, "algorithmic_bytes_per_read": alg,
            "seeding": {"algorithmic_bytes_per_read": alg["seed"], "share": alg["seed"] / alg["total"]},
            "verification": {"algorithmic_bytes_per_read": alg["verify"], "share": alg["verify"] / alg["total"],
                             "candidates_per_read": alg.get("candidates_per_read"),
                             "note": "seeding and verification run fused in the mapping kernels, so they share the "
                                     "step's time; `bench.py --workload verify` times verification alone"},
            "dram_frac": (t["dram_bytes_per_step"] / kernel_s / 1e9 / peak) if t else None,
            "dram_bytes_per_read": (t["dram_bytes_per_step"] / n) if t else None,
            "traffic_source": (t.get("source") if t else
                               "no ncu capture of these kernel sources (hash %s) on this workload in profiles/traffic.json"
                               % kernel_source_hash()),
            "own_floor_bytes": floor,
            "own_floor_frac": (floor * n / kernel_s / 1e9 / peak) if floor else None,
            "engine_counters_per_step": dstats,
            "note": "frac = SURVEY.md 8(d) algorithmic bytes (the reference's binary-search probes) / device time / "
                    "measured copy bandwidth; dram_frac = measured DRAM bytes (ncu, same kernel sources) / device time / "
                    "the same peak; own_floor = 64 B per lookup + 72 B per candidate + the read and its result"}
    return roof


def verify_leg(args, local, stream, barrier, steps=None, warmup=None):
    """the micro-benchmark proper, plus two variants that show what bounds it: longer reads (more algorithmic
    bytes per 32-byte sector moved) and a repeat set that fits the L2 (windows come from the L2, only the
    entry array streams from HBM)"""
    out = verify_one(args, local, stream, barrier, steps, warmup)
    out["variants"] = []
    for name, kw in (("reads of 192 bases (the longest the verification kernel takes)", {"read_len": 192}),
                     ("half the repeat copies (139 families, 32 % of the tiles: their windows fit the 126 MB L2)",
                      {"verify": {"n_families": 139, "rep_pct": 32}})):
        try:
            v = verify_one(args, local, stream, barrier, steps, warmup, parity=False, **kw)
            out["variants"].append({"what": name, "value": v["value"], "verify_kernel_ms_per_step": v["verify_kernel_ms_per_step"],
                                    "slots_per_read": v["slots_per_read"], "frac": v["roofline"]["frac"],
                                    "achieved": v["roofline"]["achieved"],
                                    "algorithmic_bytes_per_candidate": v["roofline"]["algorithmic_bytes_per_candidate"]})
        except Exception as ex:
            out["variants"].append({"what": name, "error": str(ex)[-200:]})
    return out


def verify_one(args, local, stream, barrier, steps=None, warmup=None, parity=True, read_len=None, verify=None):
    """Verification alone (north_star's ">= 50 % of HBM roofline on verification"): reads
    whose six seed lookups each meet thousands of candidates.  The engine parks them; verify_kernel then
    checks every candidate of every run, a warp per 32 slots -- that kernel's launches are timed by CUDA
    events on the stream it runs on (walt_engine_set_kernel_timing) and the slots it verified counted on
    the device.  bytes_verify = slots * (4 + ceil(rl / 4)) as in SURVEY.md 8(d)."""
    import refio
    gmb, nreads, rl = FULL_SIZE["verify"]
    rl = read_len or rl
    steps, warmup = steps or args.steps, warmup or args.warmup
    wl = Workload(args, local, 0, kind="verify", genome_mb=gmb, reads=nreads, read_len=rl, verify=verify)
    e = wl.e
    e.set_kernel_timing(True)
    for _ in range(warmup):
        wl.device_step(stream.cuda_stream)
    e.device_stats()
    step_ms, t_dev, _ = time_device_steps(wl, steps, 0, stream, barrier)
    st = e.device_stats()
    peak, peak_kind = measured_peak_gbs()
    t_verify = st["verify_ns"] / 1e9
    slots = st["n_verify_slots"]
    per = 4 + -(-rl // 4)
    out = {"workload": workload_text("verify", gmb, nreads, rl), "steps": steps, "warmup": warmup,
           "metric": METRICS["verify"][0], "value": slots / t_verify if t_verify else None, "unit": METRICS["verify"][1],
           "verify_kernel_ms_per_step": 1e3 * t_verify / steps, "step_ms": float(np.mean(step_ms)),
           "slots_verified_per_step": slots // steps, "slots_per_read": slots / steps / nreads, "reads_parked_per_step": st["n_parked"] // steps,
           "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peak, "peak_kind": peak_kind,
                        "achieved": slots * per / t_verify / 1e9 if t_verify else None,
                        "frac": slots * per / t_verify / 1e9 / peak if t_verify else None,
                        "algorithmic_bytes_per_candidate": per,
                        "traffic": None, "dram_frac": None,
                        "how": "device time of the verify_kernel launches (CUDA events on their stream) over the timed steps; "
                               "slots counted by the kernel"},
           "index_build_s": round(wl.t_index, 1)}
    t = committed_traffic("verify", nreads, gmb)
    if t and t_verify:
        out["roofline"]["traffic"] = t["dram_bytes_per_step"]
        out["roofline"]["dram_frac"] = t["dram_bytes_per_step"] * steps / t_verify / 1e9 / peak
    if parity and not args.no_cpu and refio.have_reference():   # parity of the whole step (park, verify, fold) on a sample
        try:
            hidx = wl.host_index()
            ns = min(512, nreads)
            ctr, obest = oracle_counters(wl, hidx, ns)
            got = wl.d_out[: ns * 16].cpu().numpy().view(refio.BEST_DT)
            out["parity_check"] = {"sample_reads": ns, "candidates_per_read_oracle": ctr["n_cand"] / ns,
                                   "fields_differing_vs_oracle": sum(int((got[f] != obest[f]).sum())
                                                                     for f in ("genome_pos", "times", "mismatch", "strand"))}
            free_host_indexes(hidx)
        except Exception as ex:
            out["parity_check"] = {"error": str(ex)[-300:]}
    wl.close()
    return out


def other_configs(args, local, stream, barrier):
    """configs[0], [2], [3], [4] at full size on rank 0's GPU: device-resident timing, roofline,
    parity of the timed path against the reference.  Each builds its own genome and index."""
    out = []
    for kind in ("se_ag", "pe", "se_small", "pe_stress"):
        if kind == "se_small" and HOST_INDEX_CACHE:
            free_host_indexes()              # the uniform genome's exports are not needed any more
        gmb, nreads, rl = FULL_SIZE[kind]
        scale = float(os.environ.get("WALT_BENCH_SCALE", "1"))   # flow tests only: shrink the other configurations
        if scale != 1.0 and kind != "se_small":
            gmb, nreads = gmb * scale, int(nreads * scale)
        entry = {"config": CONFIG_NO[kind], "workload": workload_text(kind, gmb, nreads, rl)}
        try:
            t0 = time.time()
            wl = Workload(args, local, 0, kind=kind, genome_mb=gmb, reads=nreads, read_len=rl)
            metric, unit = METRICS[kind]
            small = gmb < 500
            step_ms, t_dev, _ = time_device_steps(wl, args.steps, args.warmup, stream, barrier, flush_l2=small)
            launches = args.steps * wl.launches_per_step()
            wl.e.device_stats()
            wl.device_step(stream.cuda_stream)
            dstats = wl.e.device_stats()
            kernel_s = float(np.mean(step_ms)) / 1e3
            entry.update({"metric": metric, "value": nreads * args.steps / t_dev, "unit": unit, "ms_per_step": 1e3 * t_dev / args.steps,
                          "steps": args.steps, "warmup": args.warmup, "gpu_launches": launches,
                          "index_build_s": round(wl.t_index, 1), "hbm_index_bytes": wl.e.hbm_bytes(),
                          "l2": "flushed between timed steps" if small else "inputs larger than L2"})
            legs = side_legs(wl, unit, 5.0, False)
            if legs["alg"] is not None:
                r = roofline_dict(wl, legs["alg"], kernel_s, dstats, gmb)
                entry["roofline"] = {k: r[k] for k in ("frac", "achieved", "dram_frac", "traffic", "own_floor_bytes", "own_floor_frac",
                                                       "kernel_ms", "algorithmic_bytes_per_read")}
            entry["parity_check"] = legs["parity"]
            entry["fields_differing_vs_reference"] = (legs["parity"] or {}).get("fields_differing_vs_reference")
            entry["reference"] = legs["cpu_baseline"]
            wl.close()
            if kind == "se_small":   # files in -> files out at configs[0] size, against the reference program
                try:
                    entry["cli"] = cli_leg(local, gmb, nreads, rl, small=True)
                except Exception as ex:
                    entry["cli"] = {"error": str(ex)[-300:]}
            entry["wall_s"] = round(time.time() - t0, 1)
        except Exception as ex:
            entry["error"] = str(ex)[-400:]
        out.append(entry)
    free_host_indexes()
    out.sort(key=lambda c: c["config"])
    return out


def run_ours(args):
    import torch
    rank, local, world = dist_env()
    metric, unit = METRICS[args.workload]
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"     # keep stdout to the one JSON line
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    else:
        dist = None
        torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    numa_cpus = bind_to_gpu_numa_node(local) if world > 1 else None
    wl = Workload(args, local, rank)
    e, n, rl = wl.e, wl.n, wl.rl
    pe = wl.is_pe
    stream = torch.cuda.current_stream()
    small = args.genome_mb < 500

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (`value`) ----
    sampler = ClockSampler(local)
    sampler.start()
    step_ms, t_dev, win = time_device_steps(wl, args.steps, args.warmup, stream, barrier, flush_l2=small)
    windows = [win]
    launches_dev = args.steps * wl.launches_per_step()   # kernels of the timed device-resident steps
    e.device_stats()
    wl.device_step(stream.cuda_stream)
    dstats = e.device_stats()    # work counters of one device-resident step (its results stay in wl.d_out)

    # ---- side legs on rank 0 (not timed): oracle counters, parity of the timed path, CPU baseline ----
    legs = {"alg": None, "parity": None, "cpu_baseline": None}
    if rank == 0 and not args.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        legs = side_legs(wl, unit, 12.0 if world == 1 else 2.0, world == 1)
        if world > 1:
            legs["cpu_baseline"] = None
    alg, parity, cpu_baseline = legs["alg"], legs["parity"], legs["cpu_baseline"]

    if args.no_e2e:   # kernel experiments only: not a bench line
        sampler.stop(windows)
        if rank == 0:
            emit(json.dumps({"experiment": True, "workload": wl.kind, "kernel_ms": float(np.mean(step_ms)),
                              "per_s": n * 1e3 / float(np.mean(step_ms)), "stats": dstats,
                              "env": {k: v for k, v in os.environ.items() if k.startswith("WALT_")}}))
        return 0
    # ---- end-to-end timing through the C ABI with pinned host buffers (`e2e`) ----
    from walt_b200.engine import PinnedArray
    from walt_b200.sharding import max_over_ranks
    h_offs = PinnedArray((n + 1,), np.uint64)
    h_offs.array[:] = np.arange(n + 1, dtype=np.uint64) * np.uint64(rl)
    h_out = PinnedArray((n,), wl.out_dt)
    h_reads = PinnedArray((n * rl,), np.uint8)
    h_reads.array[:] = wl.d_reads.cpu().numpy()
    h_reads2 = None
    if pe:
        h_reads2 = PinnedArray((n * rl,), np.uint8)
        h_reads2.array[:] = wl.d_reads2.cpu().numpy()

    def e2e_step():
        if pe:
            e.map_pe_compact(h_reads.array, h_offs.array, h_reads2.array, h_offs.array, m=wl.m, b=B, top_k=TOP_K,
                             frag_range=FRAG, pbat=wl.pbat, out=h_out.array)
        else:
            e.map_se(h_reads.array, h_offs.array, ag=wl.ag, m=wl.m, b=B, out=h_out.array)

    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    windows.append((t0, t0 + t_e2e))
    launches_e2e = e.stats()["n_kernel_launches"]
    same = bool(np.array_equal(h_out.array.view(np.uint8), wl.d_out.cpu().numpy()))
    ascii_out = h_out.array.copy()

    # ---- the same call on 2-bit packed host batches (walt_engine_map_se_packed; `e2e_packed`) ----
    from walt_b200 import host as wh
    pk_bytes = (n * rl >> 2) + n + 16
    h_pk = PinnedArray((pk_bytes,), np.uint8)
    wh.pack_reads_2bit(h_reads.array, h_offs.array, out=h_pk.array)
    h_pk2 = None
    if pe:
        h_pk2 = PinnedArray((pk_bytes,), np.uint8)
        wh.pack_reads_2bit(h_reads2.array, h_offs.array, out=h_pk2.array)
    h_out.array.view(np.uint8)[:] = 0

    def e2e_packed_step():
        if pe:
            e.map_pe_compact_packed(h_pk.array, h_offs.array, h_pk2.array, h_offs.array, m=wl.m, b=B, top_k=TOP_K,
                                    frag_range=FRAG, pbat=wl.pbat, out=h_out.array)
        else:
            e.map_se_packed(h_pk.array, h_offs.array, ag=wl.ag, m=wl.m, b=B, out=h_out.array)

    for _ in range(max(1, args.warmup // 2)):
        e2e_packed_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_packed_step()
    torch.cuda.synchronize()
    t_pk = time.perf_counter() - t0
    windows.append((t0, t0 + t_pk))
    same_pk = bool(np.array_equal(h_out.array.view(np.uint8), ascii_out.view(np.uint8)))

    # ---- a ragged batch (reads of 100..150 bases): the offsets array crosses PCIe too and the kernels take every
    #      length from it (`e2e_ragged`; single-end, the first 2 M reads cut to pseudo-random lengths) ----
    ragged = None
    if not pe and world == 1:
        try:
            nr = min(n, 2_000_000)
            lens = (100 + (np.arange(nr, dtype=np.uint64) * np.uint64(2654435761) >> np.uint64(7)) % np.uint64(rl - 99)).astype(np.int64)
            r_offs = PinnedArray((nr + 1,), np.uint64)
            r_offs.array[0] = 0
            r_offs.array[1:] = np.cumsum(lens).astype(np.uint64)
            r_reads = PinnedArray((int(r_offs.array[nr]),), np.uint8)
            keep = (np.arange(rl, dtype=np.int64)[None, :] < lens[:, None])
            r_reads.array[:] = h_reads.array[: nr * rl].reshape(nr, rl)[keep]
            del keep
            r_pk = PinnedArray((int(r_offs.array[nr] >> np.uint64(2)) + nr + 16,), np.uint8)
            wh.pack_reads_2bit(r_reads.array, r_offs.array, out=r_pk.array)
            r_out = PinnedArray((nr,), wl.out_dt)
            r_steps = max(2, min(args.steps, 5))
            times = {}
            for name, call in (("ascii", lambda: e.map_se(r_reads.array, r_offs.array, ag=wl.ag, m=wl.m, b=B, out=r_out.array)),
                               ("packed", lambda: e.map_se_packed(r_pk.array, r_offs.array, ag=wl.ag, m=wl.m, b=B, out=r_out.array))):
                call()
                if name == "ascii":
                    r_first = r_out.array.copy()
                barrier()
                t0 = time.perf_counter()
                for _ in range(r_steps):
                    call()
                torch.cuda.synchronize()
                times[name] = time.perf_counter() - t0
                windows.append((t0, t0 + times[name]))
            t_ra, t_rp = times["ascii"], times["packed"]
            ragged = {"value": nr * world * r_steps / t_ra, "unit": unit, "packed_value": nr * world * r_steps / t_rp,
                      "reads_per_step": nr, "steps": r_steps, "read_lengths": f"100..{rl}, mean {float(lens.mean()):.1f}",
                      "h2d_bytes_per_step": int(r_reads.array.size + 8 * (nr + 1)),
                      "h2d_bytes_per_step_packed": int(r_pk.array.size + 8 * (nr + 1)), "d2h_bytes_per_step": int(16 * nr),
                      "ms_per_step": 1e3 * t_ra / r_steps, "ms_per_step_packed": 1e3 * t_rp / r_steps,
                      "packed_equals_ascii": bool(np.array_equal(r_out.array.view(np.uint8), r_first.view(np.uint8))),
                      "mapped_frac": float((r_first["times"] >= 1).mean()),
                      "what": "walt_engine_map_se / _packed on a batch whose reads differ in length: offsets (8 B per read) travel with the reads"}
            for h in (r_offs, r_reads, r_pk, r_out):
                h.free()
        except Exception as ex:      # an extra leg must not take the line down
            ragged = {"error": str(ex)[-300:]}
    clocks = sampler.stop(windows)
    barrier()

    t_dev, t_e2e, t_pk = max_over_ranks([t_dev, t_e2e, t_pk], dist, dev)
    for h in (h_reads, h_reads2, h_offs, h_out, h_pk, h_pk2):
        if h is not None:
            h.free()
    if rank == 0:
        total = n * world * args.steps
        kernel_s = float(np.mean(step_ms)) / 1e3
        roof = roofline_dict(wl, alg, kernel_s, dstats, args.genome_mb) if alg is not None else None
        line = {"metric": metric, "value": total / t_dev, "unit": unit, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": config_dict(args, {"index_build_s": round(wl.t_index, 1), "hbm_index_bytes": e.hbm_bytes(),
                                             "table_depth": [e.subindex_info(w)["depth"] for w in wl.which],
                                             "group_width": args.group_width,
                                             "rank0_cpus_local_to_gpu": numa_cpus,
                                             "kernel_source_sha16": kernel_source_hash(),
                                             "index_tie_order": {"rule": "std::sort replay (byte-identical to reference makedb)",
                                                                 **wl.build_info}}),
                "clocks": clocks,
                "e2e": {"value": total / t_e2e, "unit": unit, "h2d_bytes_per_step": int(n * rl * (2 if pe else 1)),
                        "d2h_bytes_per_step": int(wl.out_dt.itemsize * n), "ms_per_step": 1e3 * t_e2e / args.steps,
                        "kernel_launches_per_step": launches_e2e, "matches_device_path": same,
                        "input": "ASCII reads as the loader leaves them (the reference's representation at the seam)"},
                "e2e_packed": {"value": total / t_pk, "unit": unit,
                               "h2d_bytes_per_step": int(((n * rl >> 2) + n) * (2 if pe else 1)),
                               "d2h_bytes_per_step": int(wl.out_dt.itemsize * n), "ms_per_step": 1e3 * t_pk / args.steps,
                               "identical_to_e2e_result": same_pk,
                               "input": "2-bit packed reads (walt_pack_reads, packed by the loader outside the timed region)"},
                "e2e_ragged": ragged,
                "gpu_launches": launches_dev,
                "roofline": roof, "cpu_baseline": cpu_baseline, "parity_check": parity}
        if world == 1 and args.workload == "se" and not args.no_cpu:
            wl.close()
            if not args.no_configs:
                line["configs"] = other_configs(args, local, stream, barrier)
                try:
                    line["verify"] = verify_leg(args, local, stream, barrier)
                except Exception as ex:
                    line["verify"] = {"error": str(ex)[-300:]}
            if not args.no_cli:
                # files in -> files out through the walt program, next to the reference program (not a timed step)
                try:
                    line["cli"] = cli_leg(local, args.cli_genome_mb, args.cli_reads, rl)
                except Exception as ex:   # the bench line stands without it
                    line["cli"] = {"error": str(ex)[-300:]}
        emit(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


_RESULT_OUT = None


def emit(text):
    """The one JSON line goes to the process's real stdout; everything else that native code writes
    to fd 1 while the bench runs (NCCL's version banner, for one) has been pointed at stderr."""
    out = _RESULT_OUT or sys.stdout
    out.write(text + "\n")
    out.flush()


def main():
    global _RESULT_OUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genome-mb", type=float, default=0.0, help="default: the configuration's full size (3100; 10 for se_small)")
    ap.add_argument("--workload", default="se", choices=["se_small", "se", "se_ag", "pe", "pe_stress", "cli", "verify"],
                    help="se = configs[1] (the bench line), se_small = configs[0], se_ag = configs[2], pe = configs[3], "
                         "pe_stress = configs[4], cli = the walt program on files against the reference program")
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` block (the other four configurations) of the default run")
    ap.add_argument("--cli-genome-mb", type=float, default=300.0)
    ap.add_argument("--cli-reads", type=int, default=10_000_000)
    ap.add_argument("--makedb-genome-mb", type=float, default=50.0,
                    help="--workload cli: also run makedb against the reference makedb on a genome of this size (0 = skip)")
    ap.add_argument("--makedb-repeats", action="store_true", help="... on the repeat-heavy genome model (tied suffixes)")
    ap.add_argument("--no-cli", action="store_true", help="skip the files-in/files-out leg of the default run")
    ap.add_argument("--reads", type=int, default=0, help="reads (pairs) per GPU; default 10 M reads / 5 M pairs")
    ap.add_argument("--read-len", type=int, default=0, help="default: 150 (100 for se_small)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the oracle/reference side legs")
    ap.add_argument("--no-e2e", action="store_true", help="kernel experiment: print the device timing only")
    ap.add_argument("--group-width", type=int, default=8, help="lanes that own one read (8, 16, 32)")
    ap.add_argument("--table-depth", type=int, default=0, help="prefix-table depth (0 = auto)")
    args = ap.parse_args()
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    full = FULL_SIZE.get(args.workload, FULL_SIZE["se"])
    if args.genome_mb <= 0:
        args.genome_mb = full[0]
    if args.reads <= 0:
        args.reads = full[1]
    if args.read_len <= 0:
        args.read_len = full[2]
    if args.workload == "cli":
        sys.exit(run_cli(args))
    if args.workload == "verify":
        sys.exit(run_verify(args))
    sys.exit(run_reference(args) if args.impl == "reference" else run_ours(args))


if __name__ == "__main__":
    main()
