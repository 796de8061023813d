"""ctypes binding for the CPU fiber harness (tests/emu/libwalt_emu.so). TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PAIR_DT = np.dtype([("best_times", np.uint32), ("best_i", np.int32), ("best_j", np.int32),
                    ("frag_len", np.int32)])
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.run(["make", "-s", "-C", HERE], check=True)
        L = C.CDLL(os.path.join(HERE, "libwalt_emu.so"))
        L.emu_engine_create.restype = C.c_void_p
        L.emu_engine_depth.restype = C.c_uint32
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class EmuEngine:
    def __init__(self, lengths):
        self.L = lib()
        lengths = np.ascontiguousarray(lengths, dtype=np.uint32)
        self.h = C.c_void_p(self.L.emu_engine_create(C.c_uint32(len(lengths)), _p(lengths)))

    def load(self, which, seq, index, force_depth=0):
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        index = np.ascontiguousarray(index, dtype=np.uint32)
        rc = self.L.emu_engine_load_subindex(self.h, C.c_int(which), _p(seq), _p(index),
                                             C.c_uint32(index.size), C.c_int(force_depth))
        return rc

    def depth(self, which):
        return int(self.L.emu_engine_depth(self.h, C.c_int(which)))

    def map_se(self, buf, offs, best_dtype, ag=False, m=6, b=5000, literal=False, threads=8, width=32, packed=False):
        n = len(offs) - 1
        out = np.zeros(n, dtype=best_dtype)
        ctr = np.zeros(3, dtype=np.uint64)
        rc = self.L.emu_map_se(self.h, _p(buf), _p(offs), C.c_uint32(n), C.c_int(int(ag)), C.c_uint32(m),
                               C.c_uint32(b), C.c_int(int(literal)), _p(out), C.c_int(threads), _p(ctr),
                               C.c_uint32(width), C.c_int(int(packed)))
        return rc, out, ctr

    def map_pe_mate(self, buf, offs, cand_dtype, ag, m=6, b=5000, top_k=50, literal=False, threads=8, width=32,
                    logged=False):
        n = len(offs) - 1
        ranked = np.zeros((n, top_k), dtype=cand_dtype)
        sizes = np.zeros(n, dtype=np.uint32)
        rc = self.L.emu_map_pe_mate(self.h, _p(buf), _p(offs), C.c_uint32(n), C.c_int(int(ag)),
                                    C.c_uint32(m), C.c_uint32(b), C.c_uint32(top_k), C.c_int(int(literal)),
                                    _p(ranked), _p(sizes), C.c_int(threads), C.c_uint32(width), C.c_int(int(logged)))
        return rc, ranked, sizes

    def pair(self, r1, n1, offs1, r2, n2, offs2, top_k, m, frag_range):
        n = len(n1)
        out = np.zeros(n, dtype=PAIR_DT)
        self.L.emu_pair(self.h, _p(r1), _p(n1), _p(offs1), _p(r2), _p(n2), _p(offs2), C.c_uint32(n),
                        C.c_uint32(top_k), C.c_uint32(m), C.c_int(frag_range), _p(out))
        return out

    def pair_wide(self, r1, n1, offs1, r2, n2, offs2, top_k, m, frag_range, threads=8):
        n = len(n1)
        out = np.zeros(n, dtype=PAIR_DT)
        rc = self.L.emu_pair_wide(self.h, _p(r1), _p(n1), _p(offs1), _p(r2), _p(n2), _p(offs2), C.c_uint32(n),
                                  C.c_uint32(top_k), C.c_uint32(m), C.c_int(frag_range), _p(out), C.c_int(threads))
        assert rc == 0
        return out

    def close(self):
        if self.h:
            self.L.emu_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()
