"""ctypes binding of include/walt_b200.h (libwaltb200.so).

Mirrors the seam the reference exposes to its own drivers -- "given an index pair and a batch
of reads, fill map_results[0..n)" (src/walt/mapping.cpp:486-500, src/walt/paired.cpp:642-699)
-- with the reference's argument meaning: reads are the ACGT-only strings
LoadReadsFromFastqFile produced, `m`/`b`/`top_k`/`frag_range` are -m/-b/-k/-L, results are
BestMatch / CandidatePosition records indexed by read.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

CT00, CT01, GA10, GA11 = 0, 1, 2, 3

BEST_DT = np.dtype([("genome_pos", np.uint32), ("times", np.uint32), ("mismatch", np.uint32),
                    ("strand", "S1"), ("pad", "S3")])
CAND_DT = np.dtype([("genome_pos", np.uint32), ("mismatch", np.uint32), ("strand", "S1"),
                    ("pad", "S3")])
PAIR_DT = np.dtype([("best_times", np.uint32), ("best_i", np.int32), ("best_j", np.int32),
                    ("frag_len", np.int32)])
PE_RESULT_DT = np.dtype([("pair", PAIR_DT), ("c1", CAND_DT), ("c2", CAND_DT), ("single1", BEST_DT),
                         ("single2", BEST_DT)])
STATS_FIELDS = ("n_lookups", "n_candidates", "n_literal", "n_kernel_launches", "n_parked", "n_verify_slots", "verify_ns")

_ROOT = os.path.dirname(os.path.abspath(__file__))
_lib = None


class WaltError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"walt_b200 error {code}: {msg}")
        self.code = code


def lib_path():
    return os.path.join(_ROOT, "lib", "libwaltb200.so")


def load_library():
    """Load libwaltb200.so; raises (no fallback) if it has not been built."""
    global _lib
    if _lib is None:
        p = lib_path()
        if not os.path.exists(p):
            raise WaltError(-1, f"{p} is missing: build it with `make -C walt_b200/csrc` "
                                "(there is no CPU fallback)")
        L = C.CDLL(p)
        L.walt_last_error.restype = C.c_char_p
        L.walt_engine_hbm_bytes.restype = C.c_uint64
        L.walt_engine_hbm_bytes.argtypes = [C.c_void_p]
        L.walt_host_alloc.restype = C.c_void_p
        L.walt_host_alloc.argtypes = [C.c_size_t]
        L.walt_host_free.argtypes = [C.c_void_p]
        L.walt_packed_genome_bytes.restype = C.c_uint64
        L.walt_packed_genome_bytes.argtypes = [C.c_uint64]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def pack_reads(reads):
    """list of bytes / (n, rl) uint8 array -> (uint8 buffer, uint64 offsets[n+1])"""
    if isinstance(reads, np.ndarray) and reads.ndim == 2:
        n, rl = reads.shape
        return np.ascontiguousarray(reads).reshape(-1), np.arange(n + 1, dtype=np.uint64) * np.uint64(rl)
    lens = np.array([len(r) for r in reads], dtype=np.uint64)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    buf = np.frombuffer(b"".join(bytes(r) for r in reads), dtype=np.uint8).copy()
    if buf.size == 0:
        buf = np.zeros(1, np.uint8)
    return buf, offs


def packed_genome_bytes(n_bases):
    return int(load_library().walt_packed_genome_bytes(C.c_uint64(n_bases)))


_synth = None


def synth_library():
    """libwaltsynth.so: the BENCH-ONLY synthetic workload generators (include/walt_synth.h); not part of the product."""
    global _synth
    if _synth is None:
        load_library()
        p = os.path.join(_ROOT, "lib", "libwaltsynth.so")
        if not os.path.exists(p):
            raise WaltError(-1, f"{p} is missing: build it with `make -C walt_b200/bench`")
        _synth = C.CDLL(p)
    return _synth


def synth_genome_device(device, n_bases, seed, d_out, repeats=False):
    S = synth_library()
    fn = S.walt_synth_repeat_genome_device if repeats else S.walt_synth_genome_device
    rc = fn(C.c_int(device), C.c_uint64(n_bases), C.c_uint64(seed), C.c_void_p(d_out))
    if rc:
        raise WaltError(rc, load_library().walt_last_error().decode())


def synth_verify_genome_device(device, n_bases, seed, n_families, rep_pct, div_per_mille, d_out):
    rc = synth_library().walt_synth_verify_genome_device(C.c_int(device), C.c_uint64(n_bases), C.c_uint64(seed),
                                                         C.c_uint32(n_families), C.c_uint32(rep_pct),
                                                         C.c_uint32(div_per_mille), C.c_void_p(d_out))
    if rc:
        raise WaltError(rc, load_library().walt_last_error().decode())


class PinnedArray:
    """numpy view over cudaMallocHost memory (walt_host_alloc)."""

    def __init__(self, shape, dtype):
        L = load_library()
        self.dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * self.dtype.itemsize
        self.ptr = L.walt_host_alloc(max(n, 1))
        if not self.ptr:
            raise WaltError(3, L.walt_last_error().decode())
        buf = (C.c_char * max(n, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            load_library().walt_host_free(C.c_void_p(self.ptr))
            self.ptr = None


class Engine:
    def __init__(self, device=0, handle=None):
        """handle: wrap an engine that somebody else owns (a member of a Group)"""
        self.L = load_library()
        self.owned = handle is None
        if handle is None:
            handle = C.c_void_p()
            self._check(self.L.walt_engine_create(C.byref(handle), C.c_int(device)))
        self.h = handle
        self.device = device

    def _check(self, rc):
        if rc != 0:
            raise WaltError(rc, self.L.walt_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            if self.owned:
                self.L.walt_engine_destroy(self.h)
            self.h = None

    def clone_index_from(self, src):
        """walt_engine_clone_index: every resident sub-index of `src`, device to device"""
        self._check(self.L.walt_engine_clone_index(self.h, src.h))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- index residency ----
    def load_dbindex(self, path, which=(CT00, CT01)):
        mask = 0
        for w in which:
            mask |= 1 << w
        self._check(self.L.walt_engine_load_dbindex(self.h, path.encode(), C.c_uint32(mask)))

    def set_chromosomes(self, lengths, names=None):
        lengths = np.ascontiguousarray(lengths, dtype=np.uint32)
        arr = None
        if names is not None:
            arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
        self._check(self.L.walt_engine_set_chromosomes(self.h, C.c_uint32(lengths.size), _p(lengths), arr))

    def load_subindex(self, which, seq, counter, index):
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        index = np.ascontiguousarray(index, dtype=np.uint32)
        if counter is not None:
            counter = np.ascontiguousarray(counter, dtype=np.uint32)
        self._check(self.L.walt_engine_load_subindex(self.h, C.c_int(which), _p(seq), _p(counter),
                                                     _p(index) if index.size else None,
                                                     C.c_uint32(index.size)))

    def build_from_sequence(self, seq, which=(CT00, CT01, GA10, GA11)):
        """makedb on the device from the concatenated upper-case ACGT genome (host array)."""
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        mask = sum(1 << w for w in which)
        self._check(self.L.walt_engine_build_from_sequence(self.h, _p(seq), C.c_uint32(mask)))

    def build_from_device_genome(self, d_packed, which=(CT00, CT01)):
        mask = sum(1 << w for w in which)
        self._check(self.L.walt_engine_build_from_device_genome(self.h, C.c_void_p(d_packed), C.c_uint32(mask)))

    def export_subindex(self, which, genome_len, want_seq=True, want_counter=True, want_index=True):
        """-> (seq uint8[L] | None, counter uint32[4^12+1] | None, index uint32[index_size] | None)"""
        n = self.subindex_info(which)["index_size"]
        seq = np.zeros(genome_len, np.uint8) if want_seq else None
        counter = np.zeros((1 << 24) + 1, np.uint32) if want_counter else None
        index = np.zeros(max(n, 1), np.uint32) if want_index else None
        got = C.c_uint32()
        self._check(self.L.walt_engine_export_subindex(self.h, C.c_int(which), _p(seq), _p(counter), _p(index),
                                                       C.byref(got)))
        return seq, counter, (index[:n] if index is not None else None)

    def synth_reads_device(self, d_packed, n_reads, read_len, seed, a_rich, d_out):
        self._check(synth_library().walt_synth_reads_device(self.h, C.c_void_p(d_packed), C.c_uint32(n_reads),
                                                            C.c_uint32(read_len), C.c_uint64(seed), C.c_int(int(a_rich)),
                                                            C.c_void_p(d_out)))

    def synth_pairs_device(self, d_packed, n_pairs, read_len, seed, d_out1, d_out2, readthrough_pct=0):
        self._check(synth_library().walt_synth_pairs_device(self.h, C.c_void_p(d_packed), C.c_uint32(n_pairs),
                                                            C.c_uint32(read_len), C.c_uint64(seed),
                                                            C.c_uint32(readthrough_pct), C.c_void_p(d_out1),
                                                            C.c_void_p(d_out2)))

    def synth_verify_reads_device(self, d_packed, n_reads, read_len, seed, genome_seed, n_families, rep_pct, d_out):
        self._check(synth_library().walt_synth_verify_reads_device(self.h, C.c_void_p(d_packed), C.c_uint32(n_reads),
                                                                   C.c_uint32(read_len), C.c_uint64(seed), C.c_uint64(genome_seed),
                                                                   C.c_uint32(n_families), C.c_uint32(rep_pct), C.c_void_p(d_out)))

    def subindex_info(self, which):
        a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
        self._check(self.L.walt_engine_subindex_info(self.h, C.c_int(which), C.byref(a), C.byref(b), C.byref(c)))
        return {"index_size": a.value, "depth": b.value, "n_taint": c.value}

    def hbm_bytes(self):
        return int(self.L.walt_engine_hbm_bytes(self.h))

    def set_search_mode(self, literal):
        self._check(self.L.walt_engine_set_search_mode(self.h, C.c_int(1 if literal else 0)))

    def set_table_depth(self, depth):
        self._check(self.L.walt_engine_set_table_depth(self.h, C.c_int(depth)))

    def set_tie_order(self, ascending_position):
        """False (default): equal suffixes in std::sort's order (byte-identical to reference makedb)."""
        self._check(self.L.walt_engine_set_tie_order(self.h, C.c_int(1 if ascending_position else 0)))

    def last_build_info(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        self._check(self.L.walt_engine_last_build_info(self.h, C.byref(a), C.byref(b)))
        return {"tied_slots": int(a.value), "buckets_replayed": int(b.value)}

    def set_group_width(self, lanes):
        self._check(self.L.walt_engine_set_group_width(self.h, C.c_uint32(lanes)))

    def set_defer(self, on):
        self._check(self.L.walt_engine_set_defer(self.h, C.c_int(int(on))))

    def set_kernel_timing(self, on):
        self._check(self.L.walt_engine_set_kernel_timing(self.h, C.c_int(int(on))))

    def set_chunk_reads(self, n):
        self._check(self.L.walt_engine_set_chunk_reads(self.h, C.c_uint32(n)))

    def stats(self):
        a = (C.c_uint64 * len(STATS_FIELDS))()
        self._check(self.L.walt_engine_last_stats(self.h, a))
        return dict(zip(STATS_FIELDS, (int(x) for x in a)))

    def device_stats(self):
        """work counters of the device-resident calls since the last call (waits for the device)"""
        a = (C.c_uint64 * len(STATS_FIELDS))()
        self._check(self.L.walt_engine_device_stats(self.h, a))
        return dict(zip(STATS_FIELDS, (int(x) for x in a)))

    # ---- mapping ----
    def map_se(self, buf, offs, ag=False, m=6, b=5000, out=None):
        """Both strand passes of mapping.cpp:486-500 -> (BestMatch[n], num_of_short_reads)."""
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        n = offs.size - 1
        if out is None:
            out = np.zeros(n, dtype=BEST_DT)
        short = C.c_uint32()
        self._check(self.L.walt_engine_map_se(self.h, _p(buf), _p(offs), C.c_uint32(n), C.c_int(int(ag)),
                                              C.c_uint32(m), C.c_uint32(b), _p(out), C.byref(short)))
        return out, short.value

    def map_se_packed(self, packed, offs, ag=False, m=6, b=5000, out=None):
        """map_se on a 2-bit packed batch (walt_b200.host.pack_reads_2bit); offs are base offsets."""
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        n = offs.size - 1
        if out is None:
            out = np.zeros(n, dtype=BEST_DT)
        short = C.c_uint32()
        self._check(self.L.walt_engine_map_se_packed(self.h, _p(packed), _p(offs), C.c_uint32(n), C.c_int(int(ag)),
                                                     C.c_uint32(m), C.c_uint32(b), _p(out), C.byref(short)))
        return out, short.value

    def map_se_device(self, d_seqs, d_offs, n, max_read_len, d_out, ag=False, m=6, b=5000, stream=0):
        """Kernel-only path: all pointers are device addresses (ints)."""
        self._check(self.L.walt_engine_map_se_device(self.h, C.c_void_p(d_seqs), C.c_void_p(d_offs),
                                                     C.c_uint32(n), C.c_uint32(max_read_len), C.c_int(int(ag)),
                                                     C.c_uint32(m), C.c_uint32(b), C.c_void_p(d_out),
                                                     C.c_void_p(stream)))

    def map_pe(self, buf1, offs1, buf2, offs2, m=6, b=5000, top_k=50, frag_range=1000, pbat=False):
        """paired.cpp:642-699 for one batch -> dict(ranked1, n1, ranked2, n2, pairs, short1, short2)."""
        buf1 = np.ascontiguousarray(buf1, dtype=np.uint8)
        buf2 = np.ascontiguousarray(buf2, dtype=np.uint8)
        offs1 = np.ascontiguousarray(offs1, dtype=np.uint64)
        offs2 = np.ascontiguousarray(offs2, dtype=np.uint64)
        n = offs1.size - 1
        assert offs2.size - 1 == n
        r1 = np.zeros((n, top_k), dtype=CAND_DT)
        r2 = np.zeros((n, top_k), dtype=CAND_DT)
        n1 = np.zeros(n, dtype=np.uint32)
        n2 = np.zeros(n, dtype=np.uint32)
        pairs = np.zeros(n, dtype=PAIR_DT)
        s1, s2 = C.c_uint32(), C.c_uint32()
        self._check(self.L.walt_engine_map_pe(self.h, _p(buf1), _p(offs1), _p(buf2), _p(offs2), C.c_uint32(n),
                                              C.c_uint32(m), C.c_uint32(b), C.c_uint32(top_k),
                                              C.c_int(frag_range), C.c_int(int(pbat)), _p(r1), _p(n1), _p(r2),
                                              _p(n2), _p(pairs), C.byref(s1), C.byref(s2)))
        return {"ranked1": r1, "n1": n1, "ranked2": r2, "n2": n2, "pairs": pairs, "short1": s1.value,
                "short2": s2.value}

    def map_pe_compact(self, buf1, offs1, buf2, offs2, m=6, b=5000, top_k=50, frag_range=1000, pbat=False, out=None):
        """Same mapping, per-pair summary only (pairing + GetBestMatch4Single on the device)
        -> (walt_pe_result[n], short1, short2)."""
        buf1 = np.ascontiguousarray(buf1, dtype=np.uint8)
        buf2 = np.ascontiguousarray(buf2, dtype=np.uint8)
        offs1 = np.ascontiguousarray(offs1, dtype=np.uint64)
        offs2 = np.ascontiguousarray(offs2, dtype=np.uint64)
        n = offs1.size - 1
        assert offs2.size - 1 == n
        if out is None:
            out = np.zeros(n, dtype=PE_RESULT_DT)
        s1, s2 = C.c_uint32(), C.c_uint32()
        self._check(self.L.walt_engine_map_pe_compact(self.h, _p(buf1), _p(offs1), _p(buf2), _p(offs2), C.c_uint32(n),
                                                      C.c_uint32(m), C.c_uint32(b), C.c_uint32(top_k),
                                                      C.c_int(frag_range), C.c_int(int(pbat)), _p(out),
                                                      C.byref(s1), C.byref(s2)))
        return out, s1.value, s2.value

    def map_pe_compact_packed(self, packed1, offs1, packed2, offs2, m=6, b=5000, top_k=50, frag_range=1000, pbat=False,
                              out=None):
        """map_pe_compact on 2-bit packed mates (walt_b200.host.pack_reads_2bit)."""
        packed1 = np.ascontiguousarray(packed1, dtype=np.uint8)
        packed2 = np.ascontiguousarray(packed2, dtype=np.uint8)
        offs1 = np.ascontiguousarray(offs1, dtype=np.uint64)
        offs2 = np.ascontiguousarray(offs2, dtype=np.uint64)
        n = offs1.size - 1
        assert offs2.size - 1 == n
        if out is None:
            out = np.zeros(n, dtype=PE_RESULT_DT)
        s1, s2 = C.c_uint32(), C.c_uint32()
        self._check(self.L.walt_engine_map_pe_compact_packed(self.h, _p(packed1), _p(offs1), _p(packed2), _p(offs2),
                                                             C.c_uint32(n), C.c_uint32(m), C.c_uint32(b),
                                                             C.c_uint32(top_k), C.c_int(frag_range), C.c_int(int(pbat)),
                                                             _p(out), C.byref(s1), C.byref(s2)))
        return out, s1.value, s2.value

    def map_pe_device(self, d_seqs1, d_offs1, d_seqs2, d_offs2, n, max_read_len, d_out, m=6, b=5000, top_k=50,
                      frag_range=1000, pbat=False, stream=0):
        """Kernel-only paired-end path: device addresses in, walt_pe_result[n] on the device out."""
        self._check(self.L.walt_engine_map_pe_device(self.h, C.c_void_p(d_seqs1), C.c_void_p(d_offs1),
                                                     C.c_void_p(d_seqs2), C.c_void_p(d_offs2), C.c_uint32(n),
                                                     C.c_uint32(max_read_len), C.c_uint32(m), C.c_uint32(b),
                                                     C.c_uint32(top_k), C.c_int(frag_range), C.c_int(int(pbat)),
                                                     C.c_void_p(d_out), C.c_void_p(stream)))


class Group:
    """walt_group: one engine per device behind one handle (include/walt_b200.h, "several GPUs")."""

    def __init__(self, devices):
        self.L = load_library()
        self.L.walt_group_engine.restype = C.c_void_p
        ids = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        rc = self.L.walt_group_create(C.byref(h), ids, C.c_int(len(devices)))
        if rc:
            raise WaltError(rc, self.L.walt_last_error().decode())
        self.h = h
        self.devices = list(devices)
        self.engines = [Engine(d, handle=C.c_void_p(self.L.walt_group_engine(self.h, C.c_int(i)))) for i, d in enumerate(devices)]

    def _check(self, rc):
        if rc != 0:
            raise WaltError(rc, self.L.walt_last_error().decode())

    def load_dbindex(self, path, which=(0, 1)):
        mask = sum(1 << w for w in which)
        self._check(self.L.walt_group_load_dbindex(self.h, path.encode(), C.c_uint32(mask)))

    def map_se_packed(self, packed, offs, ag=False, m=6, b=5000, out=None):
        n = offs.size - 1
        if out is None:
            out = np.zeros(n, dtype=BEST_DT)
        short = C.c_uint32()
        self._check(self.L.walt_group_map_se_packed(self.h, _p(packed), _p(offs), C.c_uint32(n), C.c_int(int(ag)), C.c_uint32(m),
                                                    C.c_uint32(b), _p(out), C.byref(short)))
        return out, short.value

    def map_se(self, buf, offs, ag=False, m=6, b=5000, out=None):
        n = offs.size - 1
        if out is None:
            out = np.zeros(n, dtype=BEST_DT)
        short = C.c_uint32()
        self._check(self.L.walt_group_map_se(self.h, _p(buf), _p(offs), C.c_uint32(n), C.c_int(int(ag)), C.c_uint32(m),
                                             C.c_uint32(b), _p(out), C.byref(short)))
        return out, short.value

    def map_pe_compact_packed(self, p1, o1, p2, o2, m=6, b=5000, top_k=50, frag_range=1000, pbat=False, out=None):
        n = o1.size - 1
        if out is None:
            out = np.zeros(n, dtype=PE_RESULT_DT)
        s1, s2 = C.c_uint32(), C.c_uint32()
        self._check(self.L.walt_group_map_pe_compact_packed(self.h, _p(p1), _p(o1), _p(p2), _p(o2), C.c_uint32(n), C.c_uint32(m),
                                                            C.c_uint32(b), C.c_uint32(top_k), C.c_int(frag_range), C.c_int(int(pbat)),
                                                            _p(out), C.byref(s1), C.byref(s2)))
        return out, s1.value, s2.value

    def close(self):
        if getattr(self, "h", None):
            for e in self.engines:
                e.close()
            self.L.walt_group_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
