"""ctypes binding of include/walt_host.h (libwalthost.so): FASTQ batches in, SAM/MR/mapstats out.
Pure host code; the mapping between the two is walt_b200.Engine (GPU)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_ROOT = os.path.dirname(os.path.abspath(__file__))
_lib = None


class HostError(RuntimeError):
    pass


def load_library():
    global _lib
    if _lib is None:
        p = os.path.join(_ROOT, "lib", "libwalthost.so")
        if not os.path.exists(p):
            raise HostError(f"{p} is missing: build it with `make -C walt_b200/host`")
        L = C.CDLL(p)
        L.walt_host_last_error.restype = C.c_char_p
        for f in ("walt_chroms_read", "walt_chroms_create", "walt_fastq_open", "walt_batch_create",
                  "walt_se_writer_open", "walt_pe_writer_open", "walt_batch_seqs", "walt_batch_offsets"):
            getattr(L, f).restype = C.c_void_p
        L.walt_batch_name.restype = C.c_char_p
        L.walt_batch_qual.restype = C.c_char_p
        L.walt_chroms_name.restype = C.c_char_p
        L.walt_chroms_lengths.restype = C.POINTER(C.c_uint32)
        L.walt_fastq_next_batch.restype = C.c_int64
        L.walt_fastq_next_part.restype = C.c_int64
        L.walt_batch_size.restype = C.c_uint32
        L.walt_chroms_count.restype = C.c_uint32
        L.walt_clip_adaptor.restype = C.c_size_t
        L.walt_packed_reads_bytes.restype = C.c_uint64
        L.walt_batch_packed.restype = C.c_void_p
        _lib = L
    return _lib


def _err():
    return HostError(load_library().walt_host_last_error().decode())


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def pack_reads_2bit(buf, offs, out=None):
    """walt_pack_reads: ASCII batch (buffer + base offsets[n+1]) -> its 2-bit packed form (uint8 array)."""
    L = load_library()
    buf = np.ascontiguousarray(buf, np.uint8)
    offs = np.ascontiguousarray(offs, np.uint64)
    n = offs.size - 1
    nbytes = int(L.walt_packed_reads_bytes(_p(offs), C.c_uint32(n)))
    if out is None:
        out = np.zeros(nbytes, np.uint8)
    assert out.size >= nbytes
    if L.walt_pack_reads(_p(buf), _p(offs), C.c_uint32(n), _p(out)) != 0:
        raise _err()
    return out


class Chroms:
    def __init__(self, dbindex_path=None, names=None, lengths=None):
        L = load_library()
        if dbindex_path is not None:
            self.h = L.walt_chroms_read(dbindex_path.encode())
        else:
            lengths = np.ascontiguousarray(lengths, np.uint32)
            arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
            self.h = L.walt_chroms_create(C.c_uint32(len(names)), arr, _p(lengths))
        if not self.h:
            raise _err()
        self.h = C.c_void_p(self.h)
        n = L.walt_chroms_count(self.h)
        self.names = [L.walt_chroms_name(self.h, C.c_uint32(i)).decode() for i in range(n)]
        self.lengths = np.array([L.walt_chroms_lengths(self.h)[i] for i in range(n)], np.uint32)


class Batch:
    def __init__(self):
        self.L = load_library()
        self.h = C.c_void_p(self.L.walt_batch_create())

    def __len__(self):
        return int(self.L.walt_batch_size(self.h))

    def arrays(self):
        """-> (uint8 seqs view, uint64 offsets view) valid until the next load"""
        n = len(self)
        offs = np.ctypeslib.as_array(C.cast(self.L.walt_batch_offsets(self.h), C.POINTER(C.c_uint64)), shape=(n + 1,))
        total = int(offs[n])
        if total == 0:
            return np.zeros(1, np.uint8), offs.copy()
        seqs = np.ctypeslib.as_array(C.cast(self.L.walt_batch_seqs(self.h), C.POINTER(C.c_uint8)), shape=(total,))
        return seqs, offs

    def packed(self):
        """-> the loader's 2-bit form of the batch (uint8 view, walt_pack_reads layout), valid until the next load"""
        n = len(self)
        offs = np.ctypeslib.as_array(C.cast(self.L.walt_batch_offsets(self.h), C.POINTER(C.c_uint64)), shape=(n + 1,))
        nbytes = int(self.L.walt_packed_reads_bytes(_p(offs), C.c_uint32(n)))
        ptr = self.L.walt_batch_packed(self.h)
        if not ptr:
            raise _err()
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(nbytes,))

    def name(self, i):
        return self.L.walt_batch_name(self.h, C.c_uint32(i)).decode()

    def qual(self, i):
        return self.L.walt_batch_qual(self.h, C.c_uint32(i)).decode("latin-1")

    def free(self):
        if self.h:
            self.L.walt_batch_free(self.h)
            self.h = None


class Fastq:
    def __init__(self, path):
        self.L = load_library()
        h = self.L.walt_fastq_open(path.encode())
        if not h:
            raise _err()
        self.h = C.c_void_p(h)

    def next_batch(self, batch, max_reads, adaptor=""):
        n = self.L.walt_fastq_next_batch(self.h, batch.h, C.c_uint32(max_reads), adaptor.encode())
        if n < 0:
            raise _err()
        return int(n)

    def next_part(self, batch, max_reads, adaptor="", restart_rand=False):
        """A batch in pieces: the first part restarts the rand() stream of the N replacement, the others carry it on."""
        n = self.L.walt_fastq_next_part(self.h, batch.h, C.c_uint32(max_reads), adaptor.encode(), C.c_int(1 if restart_rand else 0))
        if n < 0:
            raise _err()
        return int(n)

    def close(self):
        if self.h:
            self.L.walt_fastq_close(self.h)
            self.h = None


class SeWriter:
    def __init__(self, path, chroms, ag=False, ambiguous=False, unmapped=False, sam=False):
        self.L = load_library()
        h = self.L.walt_se_writer_open(path.encode(), chroms.h, C.c_int(int(ag)), C.c_int(int(ambiguous)),
                                       C.c_int(int(unmapped)), C.c_int(int(sam)))
        if not h:
            raise _err()
        self.h = C.c_void_p(h)

    def write(self, batch, results, n_short=0):
        results = np.ascontiguousarray(results)
        if self.L.walt_se_writer_write(self.h, batch.h, _p(results), C.c_uint32(len(results))):
            raise _err()
        self.L.walt_se_writer_add_short(self.h, C.c_uint32(n_short))

    def close(self):
        if self.h and self.L.walt_se_writer_close(self.h):
            raise _err()
        self.h = None


class PeWriter:
    def __init__(self, path, chroms, m=6, top_k=50, frag_range=1000, ambiguous=False, unmapped=False, sam=False,
                 pbat=False):
        self.L = load_library()
        h = self.L.walt_pe_writer_open(path.encode(), chroms.h, C.c_uint32(m), C.c_uint32(top_k), C.c_int(frag_range),
                                       C.c_int(int(ambiguous)), C.c_int(int(unmapped)), C.c_int(int(sam)),
                                       C.c_int(int(pbat)))
        if not h:
            raise _err()
        self.h = C.c_void_p(h)

    def write(self, b1, b2, r, n):
        """r: dict from Engine.map_pe (ranked1, n1, ranked2, n2, pairs, short1, short2)"""
        if self.L.walt_pe_writer_write(self.h, b1.h, b2.h, _p(r["ranked1"]), _p(r["n1"]), _p(r["ranked2"]),
                                       _p(r["n2"]), _p(r["pairs"]), C.c_uint32(n)):
            raise _err()
        self.L.walt_pe_writer_add_short(self.h, C.c_uint32(r["short1"]), C.c_uint32(r["short2"]))

    def close(self):
        if self.h and self.L.walt_pe_writer_close(self.h):
            raise _err()
        self.h = None


def set_threads(n):
    """Host threads of the loader / writers (0 = one per hardware thread)."""
    load_library().walt_host_set_threads(C.c_uint(n))


def set_grain(chunk_bytes=0, block_reads=0):
    """Tuning/test hook: loader task size in bytes and writer task size in reads (0 = default)."""
    load_library().walt_host_set_grain(C.c_uint32(chunk_bytes), C.c_uint32(block_reads))


def write_subindex(path, strand, seq, counter, index):
    """WriteIndex (reference.cpp:302-322): one _CT00/_CT01/_GA10/_GA11 file."""
    L = load_library()
    seq = np.ascontiguousarray(seq, np.uint8)
    counter = np.ascontiguousarray(counter, np.uint32)
    index = np.ascontiguousarray(index, np.uint32)
    if L.walt_write_subindex(path.encode(), C.c_char(strand.encode()), _p(seq), C.c_uint64(seq.size), _p(counter),
                             _p(index), C.c_uint32(index.size)):
        raise _err()


def write_dbindex_header(path, chroms, size_of_index):
    """WriteIndexHeadInfo (reference.cpp:353-379)."""
    if load_library().walt_write_dbindex_header(path.encode(), chroms.h, C.c_uint32(size_of_index)):
        raise _err()
