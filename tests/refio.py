"""Test-side helpers: .dbindex parser (numpy), ctypes bindings for the C oracle
(oracle/libwalt_oracle.so) and for the compiled reference shim (oracle/_ref/libwaltref.so),
and runners for the reference binaries (oracle/_ref/walt, oracle/_ref/makedb).

TEST INFRASTRUCTURE ONLY -- nothing under walt_b200/ imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
N_KEYS = 1 << 24

SUFFIXES = ("_CT00", "_CT01", "_GA10", "_GA11")


# ----------------------------------------------------------------------------------------
# .dbindex files (layout: SURVEY.md A.1; writer: reference.cpp:302-322,353-379)
# ----------------------------------------------------------------------------------------
@dataclass
class Header:
    names: list
    lengths: np.ndarray
    genome_len: int
    size_of_index: int

    @property
    def start_index(self):
        return np.concatenate([[0], np.cumsum(self.lengths.astype(np.uint64))]).astype(np.uint32)


def read_header(path) -> Header:
    buf = open(path, "rb").read()
    off = 0
    n = int(np.frombuffer(buf, np.uint32, 1, off)[0]); off += 4
    names = []
    for _ in range(n):
        ln = int(np.frombuffer(buf, np.uint32, 1, off)[0]); off += 4
        names.append(buf[off:off + ln].decode()); off += ln
    lengths = np.frombuffer(buf, np.uint32, n, off).copy(); off += 4 * n
    glen = int(np.frombuffer(buf, np.uint32, 1, off)[0]); off += 4
    soi = int(np.frombuffer(buf, np.uint32, 1, off)[0]); off += 4
    assert off == len(buf), "trailing bytes in header"
    return Header(names, lengths, glen, soi)


@dataclass
class SubIndex:
    strand: str
    seq: np.ndarray      # uint8 ASCII, converted
    counter: np.ndarray  # uint32[4^12+1]
    index: np.ndarray    # uint32[index_size]


def read_subindex(path, genome_len) -> SubIndex:
    mm = np.memmap(path, dtype=np.uint8, mode="r")
    strand = chr(mm[0])
    seq = np.array(mm[1:1 + genome_len])
    off = 1 + genome_len
    csize, isize = np.frombuffer(mm[off:off + 8].tobytes(), np.uint32)
    off += 8
    assert csize == N_KEYS
    counter = np.frombuffer(mm[off:off + 4 * (N_KEYS + 1)].tobytes(), np.uint32).copy()
    off += 4 * (N_KEYS + 1)
    index = np.frombuffer(mm[off:off + 4 * int(isize)].tobytes(), np.uint32).copy()
    off += 4 * int(isize)
    assert off == mm.shape[0], "trailing bytes in sub-index"
    return SubIndex(strand, seq, counter, index)


# ----------------------------------------------------------------------------------------
# reference binaries
# ----------------------------------------------------------------------------------------
def have_reference():
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in ("walt", "makedb", "libwaltref.so"))


def ref_makedb(fasta, out_index):
    subprocess.run([os.path.join(REF_DIR, "makedb"), "-c", fasta, "-o", out_index],
                   check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def ref_walt(args, check=True):
    return subprocess.run([os.path.join(REF_DIR, "walt")] + list(args), check=check,
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE)


# ----------------------------------------------------------------------------------------
# ctypes structs shared by oracle and shim
# ----------------------------------------------------------------------------------------
BEST_DT = np.dtype([("genome_pos", np.uint32), ("times", np.uint32), ("mismatch", np.uint32),
                    ("strand", "S1"), ("pad", "S3")])
CAND_DT = np.dtype([("genome_pos", np.uint32), ("mismatch", np.uint32), ("strand", "S1"),
                    ("pad", "S3")])


def init_best(n, max_mismatches):
    """mapping.cpp:486-489"""
    a = np.zeros(n, dtype=BEST_DT)
    a["mismatch"] = max_mismatches
    a["strand"] = b"+"
    return a


def pack_reads(reads):
    """list of bytes / 2-D uint8 array -> (concatenated uint8 buffer, uint64 offsets[n+1])"""
    if isinstance(reads, np.ndarray) and reads.ndim == 2:
        n, rl = reads.shape
        return np.ascontiguousarray(reads).reshape(-1), (np.arange(n + 1, dtype=np.uint64) * rl)
    lens = np.array([len(r) for r in reads], dtype=np.uint64)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    buf = np.frombuffer(b"".join(bytes(r) for r in reads), dtype=np.uint8).copy()
    if buf.size == 0:
        buf = np.zeros(1, np.uint8)
    return buf, offs


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


# ----------------------------------------------------------------------------------------
# C oracle
# ----------------------------------------------------------------------------------------
class WoIndex(C.Structure):
    _fields_ = [("seq", C.c_void_p), ("genome_len", C.c_uint64), ("n_chr", C.c_uint32),
                ("start_index", C.c_void_p), ("counter", C.c_void_p), ("index", C.c_void_p),
                ("index_size", C.c_uint32)]


class WoCounters(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("n_reads", "n_lookups", "sum_log2_bucket", "n_probes",
                                          "n_cand", "n_region_over_b", "n_short")]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class WoChroms(C.Structure):
    _fields_ = [("n_chr", C.c_uint32), ("start_index", C.c_void_p), ("length", C.c_void_p)]


_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        path = os.path.join(ORACLE_DIR, "libwalt_oracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-C", ORACLE_DIR, "libwalt_oracle.so"], check=True)
        L = C.CDLL(path)
        L.wo_seed_repeats.restype = C.c_uint32
        L.wo_nocared_position.restype = C.c_uint32
        L.wo_hash_value.restype = C.c_uint32
        L.wo_chrom_id.restype = C.c_uint32
        L.wo_build_index.restype = C.c_uint32
        L.wo_pe_pair.restype = C.c_uint32
        L.wo_heap_drain.restype = C.c_uint32
        L.wo_clip_adaptor.restype = C.c_size_t
        L.wo_fragment_length.restype = C.c_int
        _oracle = L
    return _oracle


class OracleIndex:
    """Keeps the numpy arrays alive behind a wo_index."""

    def __init__(self, hdr: Header, sub: SubIndex):
        self.hdr, self.sub = hdr, sub
        self.starts = np.ascontiguousarray(hdr.start_index, dtype=np.uint32)
        self.lengths = np.ascontiguousarray(hdr.lengths, dtype=np.uint32)
        self.c = WoIndex(_p(sub.seq).value, len(sub.seq), len(hdr.lengths), _p(self.starts).value,
                         _p(sub.counter).value, _p(sub.index).value if sub.index.size else None,
                         sub.index.size)
        self.chroms = WoChroms(len(hdr.lengths), _p(self.starts).value, _p(self.lengths).value)


def oracle_se_pass(oix: OracleIndex, buf, offs, strand, ag, b, best, counters=None):
    L = oracle_lib()
    L.wo_se_map_batch(C.byref(oix.c), _p(buf), _p(offs), C.c_uint32(len(offs) - 1),
                      C.c_char(strand.encode()), C.c_int(int(ag)), C.c_uint32(b), _p(best),
                      C.byref(counters) if counters is not None else None)


def oracle_se_map(hdr, subs, reads, ag=False, m=6, b=5000, counters=None):
    """Both strand passes of mapping.cpp:486-500. subs = (plus SubIndex, minus SubIndex)."""
    buf, offs = pack_reads(reads)
    best = init_best(len(offs) - 1, m)
    for sub, strand in zip(subs, "+-"):
        oracle_se_pass(OracleIndex(hdr, sub), buf, offs, strand, ag, b, best, counters)
    return best


def oracle_pe_mate(hdr, subs, reads, ag, m=6, b=5000, top_k=50, counters=None):
    """Both strand passes for one mate (paired.cpp:650-671) then the drain
    (paired.cpp:684-692).  -> (ranked CAND_DT[n, top_k], sizes[n])"""
    L = oracle_lib()
    buf, offs = pack_reads(reads)
    n = len(offs) - 1
    cands = np.zeros((n, top_k), dtype=CAND_DT)
    sizes = np.zeros(n, dtype=np.uint32)
    for sub, strand in zip(subs, "+-"):
        oix = OracleIndex(hdr, sub)
        L.wo_pe_map_batch(C.byref(oix.c), _p(buf), _p(offs), C.c_uint32(n),
                          C.c_char(strand.encode()), C.c_int(int(ag)), C.c_uint32(m),
                          C.c_uint32(b), C.c_uint32(top_k), _p(cands), _p(sizes),
                          C.byref(counters) if counters is not None else None)
    ranked = np.zeros((n, top_k), dtype=CAND_DT)

    class Heap(C.Structure):
        _fields_ = [("a", C.c_void_p), ("size", C.c_uint32), ("max_size", C.c_uint32)]

    for j in range(n):
        h = Heap(cands[j].ctypes.data, int(sizes[j]), top_k)
        got = L.wo_heap_drain(C.byref(h), C.c_void_p(ranked[j].ctypes.data))
        assert got == sizes[j]
    return ranked, sizes


# ----------------------------------------------------------------------------------------
# compiled reference via the shim
# ----------------------------------------------------------------------------------------
_ref = None


def ref_lib():
    global _ref
    if _ref is None:
        L = C.CDLL(os.path.join(REF_DIR, "libwaltref.so"))
        L.waltref_index_alloc.restype = C.c_void_p
        L.waltref_index_load.restype = C.c_void_p
        L.waltref_index_sequence.restype = C.c_void_p
        L.waltref_index_counter.restype = C.c_void_p
        L.waltref_index_index.restype = C.c_void_p
        L.waltref_index_genome_len.restype = C.c_uint64
        L.waltref_index_index_size.restype = C.c_uint32
        L.waltref_map_se.restype = C.c_uint32
        L.waltref_map_pe.restype = C.c_uint32
        L.waltref_time_se.restype = C.c_double
        if hasattr(L, "waltref_time_pe"):
            L.waltref_time_pe.restype = C.c_double
        L.waltref_heaps_alloc.restype = C.c_void_p
        _ref = L
    return _ref


class RefIndex:
    def __init__(self, header_path, sub_path):
        self.L = ref_lib()
        self.h = C.c_void_p(self.L.waltref_index_load(header_path.encode(), sub_path.encode()))

    def close(self):
        if self.h:
            self.L.waltref_index_free(self.h)
            self.h = None

    def __del__(self):
        self.close()


def ref_se_map(index_path, reads, ag=False, m=6, b=5000, threads=4):
    L = ref_lib()
    buf, offs = pack_reads(reads)
    n = len(offs) - 1
    best = init_best(n, m)
    short = 0
    sfx = ("_GA10", "_GA11") if ag else ("_CT00", "_CT01")
    for s, strand in zip(sfx, "+-"):
        ri = RefIndex(index_path, index_path + s)
        short += L.waltref_map_se(ri.h, _p(buf), _p(offs), C.c_uint32(n), C.c_char(strand.encode()),
                                  C.c_int(int(ag)), C.c_uint32(b), _p(best), C.c_int(threads))
        ri.close()
    return best, short


def ref_pe_mate(index_path, reads, ag, m=6, b=5000, top_k=50, threads=4):
    L = ref_lib()
    buf, offs = pack_reads(reads)
    n = len(offs) - 1
    heaps = C.c_void_p(L.waltref_heaps_alloc(C.c_uint32(n), C.c_uint32(top_k)))
    sfx = ("_GA10", "_GA11") if ag else ("_CT00", "_CT01")
    for s, strand in zip(sfx, "+-"):
        ri = RefIndex(index_path, index_path + s)
        L.waltref_map_pe(ri.h, heaps, _p(buf), _p(offs), C.c_uint32(n), C.c_char(strand.encode()),
                         C.c_int(int(ag)), C.c_uint32(m), C.c_uint32(b), C.c_int(threads))
        ri.close()
    ranked = np.zeros((n, top_k), dtype=CAND_DT)
    sizes = np.zeros(n, dtype=np.uint32)
    L.waltref_heaps_drain(heaps, C.c_uint32(n), C.c_uint32(top_k), _p(ranked), _p(sizes))
    L.waltref_heaps_free(heaps)
    return ranked, sizes


# ----------------------------------------------------------------------------------------
# in-memory index construction with the oracle's makedb restatement (reference.cpp:79-300)
# ----------------------------------------------------------------------------------------
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGT", b"TGCA"):
    _COMP[_a] = _b


def build_index_with_oracle(chroms):
    """chroms: list of (name, uint8 ACGT array) -> (Header, {suffix: SubIndex}).  N's must
    already be replaced (makedb does it with rand(); tests use N-free genomes here)."""
    L = oracle_lib()
    lengths = np.array([len(s) for _, s in chroms], dtype=np.uint32)
    hdr = Header([n for n, _ in chroms], lengths, int(lengths.sum()), 0)
    starts = np.ascontiguousarray(hdr.start_index, np.uint32)
    fwd = np.concatenate([s for _, s in chroms])
    rev = np.concatenate([_COMP[s[::-1]] for _, s in chroms])   # per chromosome, reference.cpp:131-146
    subs = {}
    for sfx, base, frm, to in (("_CT00", fwd, "C", "T"), ("_CT01", rev, "C", "T"),
                               ("_GA10", fwd, "G", "A"), ("_GA11", rev, "G", "A")):
        seq = base.copy()
        seq[seq == ord(frm)] = ord(to)
        counter = np.zeros(N_KEYS + 1, np.uint32)
        index = np.zeros(max(1, hdr.genome_len), np.uint32)
        n = L.wo_build_index(_p(seq), C.c_uint64(hdr.genome_len), C.c_uint32(len(lengths)), _p(starts),
                             _p(counter), _p(index))
        subs[sfx] = SubIndex("-" if sfx.endswith("1") else "+", seq, counter, index[:n].copy())
        hdr.size_of_index = max(hdr.size_of_index, int(n))
    return hdr, subs
