"""Loader for the committed golden vectors (tests/golden/, made by make_golden.py from the
unmodified reference).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import functools
import os

import numpy as np

import refio

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_CODE = np.zeros(256, dtype=np.uint32)
for _i, _c in enumerate(b"ACGT"):
    _CODE[_c] = _i


def hash_values(seq, positions):
    """getHashValue (util.hpp:175-182) of the 12 cared bases at pos+3i+1, vectorised."""
    positions = positions.astype(np.int64)
    h = np.zeros(len(positions), dtype=np.uint32)
    for i in range(12):
        h = h * 4 + _CODE[seq[positions + 3 * i + 1]]
    return h


def counter_from_index(seq, index):
    """counter[] as CountBucketSize/HashToBucket leave it (reference.cpp:220-228,252-255):
    exclusive prefix sums of the bucket sizes; index[] is grouped by ascending hash."""
    h = hash_values(seq, index)
    assert np.all(np.diff(h.astype(np.int64)) >= 0), "index not grouped by hash value"
    cnt = np.bincount(h, minlength=refio.N_KEYS).astype(np.uint64)
    counter = np.zeros(refio.N_KEYS + 1, dtype=np.uint32)
    counter[1:] = np.cumsum(cnt).astype(np.uint32)
    return counter


@functools.lru_cache(maxsize=None)
def genome():
    """-> (Header, {suffix: SubIndex})"""
    z = np.load(os.path.join(GOLDEN, "genome.npz"))
    lengths = z["lengths"].astype(np.uint32)
    names = [str(x) for x in z["names"]]
    subs = {}
    size_of_index = 0
    for sfx in refio.SUFFIXES:
        seq = z["seq" + sfx]
        index = z["index" + sfx]
        subs[sfx] = refio.SubIndex(chr(z["strand" + sfx][0]), seq, counter_from_index(seq, index), index)
        size_of_index = max(size_of_index, index.size)
    hdr = refio.Header(names, lengths, int(lengths.astype(np.uint64).sum()), size_of_index)
    return hdr, subs


def fasta_bytes():
    return np.load(os.path.join(GOLDEN, "genome.npz"))["fasta"].tobytes()


def write_dbindex(path):
    """Re-materialise the .dbindex files exactly as makedb wrote them
    (reference.cpp:302-322,353-379)."""
    hdr, subs = genome()
    with open(path, "wb") as f:
        f.write(np.uint32(len(hdr.names)).tobytes())
        for n in hdr.names:
            f.write(np.uint32(len(n)).tobytes() + n.encode())
        f.write(hdr.lengths.astype(np.uint32).tobytes())
        f.write(np.uint32(hdr.genome_len).tobytes())
        f.write(np.uint32(hdr.size_of_index).tobytes())
    for sfx, sub in subs.items():
        with open(path + sfx, "wb") as f:
            f.write(sub.strand.encode())
            f.write(sub.seq.tobytes())
            f.write(np.uint32(refio.N_KEYS).tobytes())
            f.write(np.uint32(sub.index.size).tobytes())
            f.write(sub.counter.tobytes())
            f.write(sub.index.tobytes())


def se_pair(ag):
    _, subs = genome()
    return (subs["_GA10"], subs["_GA11"]) if ag else (subs["_CT00"], subs["_CT01"])


def load(name):
    return np.load(os.path.join(GOLDEN, name))
