"""Host side (libwalthost: FASTQ loader with the glibc rand() N-replacement stream, adaptor
clipping, SAM / MR / side files / mapstats writers) against the reference `walt` outputs in
tests/golden/cli.  The mapping step between loader and writer is done by the C ORACLE here
(test-only stand-in for the GPU), so the whole text path is checked on a CPU-only box."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import goldenio
import refio

CLI = os.path.join(goldenio.GOLDEN, "cli")
CASES = json.load(open(os.path.join(CLI, "cases.json")))


def _opts(args):
    o = {"m": 6, "N": 10_000_000, "b": 5000, "k": 50, "L": 1000, "C": "", "sam": False, "u": False, "a": False,
         "A": False, "r": None, "1": None, "2": None}
    i = 0
    while i < len(args):
        a = args[i][1:]
        if a in ("sam", "u", "a", "A"):
            o[a] = True
            i += 1
        else:
            o[a] = args[i + 1]
            i += 2
    for k in ("m", "N", "b", "k", "L"):
        o[k] = int(o[k])
    return o


def oracle_pairs(hdr, r1, n1, l1, r2, n2, l2, m, frag):
    from walt_b200.engine import PAIR_DT
    L = refio.oracle_lib()
    starts = np.ascontiguousarray(hdr.start_index, np.uint32)
    lengths = np.ascontiguousarray(hdr.lengths, np.uint32)
    chroms = refio.WoChroms(len(lengths), starts.ctypes.data, lengths.ctypes.data)
    out = np.zeros(len(n1), PAIR_DT)
    for j in range(len(n1)):
        bi, bj = C.c_int32(-1), C.c_int32(-1)
        a = np.ascontiguousarray(r1[j]); b = np.ascontiguousarray(r2[j])
        t = L.wo_pe_pair(C.byref(chroms), a.ctypes.data_as(C.c_void_p), C.c_uint32(int(n1[j])), C.c_uint32(int(l1[j])),
                         b.ctypes.data_as(C.c_void_p), C.c_uint32(int(n2[j])), C.c_uint32(int(l2[j])),
                         C.c_uint32(m), C.c_int(frag), C.byref(bi), C.byref(bj))
        out[j] = (t, bi.value, bj.value, 0)
    return out


def run_case_with_oracle(case, outdir):
    from walt_b200 import host
    hdr, _ = goldenio.genome()
    o = _opts(case["args"])
    chroms = host.Chroms(names=hdr.names, lengths=hdr.lengths)
    out = os.path.join(outdir, "out")
    open(out, "w").close()
    open(out + ".mapstats", "w").close()
    if o["r"]:
        fq = host.Fastq(os.path.join(CLI, o["r"]))
        w = host.SeWriter(out, chroms, ag=o["A"], ambiguous=o["a"], unmapped=o["u"], sam=o["sam"])
        b = host.Batch()
        while True:
            n = fq.next_batch(b, o["N"], o["C"])
            if n == 0:
                break
            seqs, offs = b.arrays()
            best = refio.init_best(n, o["m"])
            ctr = refio.WoCounters()
            for sub, strand in zip(goldenio.se_pair(o["A"]), "+-"):
                refio.oracle_se_pass(refio.OracleIndex(hdr, sub), seqs, offs, strand, o["A"], o["b"], best, ctr)
            w.write(b, best, n_short=int(ctr.n_short))
            if n < o["N"]:
                break
        w.close(); fq.close(); b.free()
    else:
        ad = o["C"].split(":")
        ad1, ad2 = (ad[0], ad[-1])
        f1 = host.Fastq(os.path.join(CLI, o["1"])); f2 = host.Fastq(os.path.join(CLI, o["2"]))
        w = host.PeWriter(out, chroms, m=o["m"], top_k=o["k"], frag_range=o["L"], ambiguous=o["a"], unmapped=o["u"],
                          sam=o["sam"])
        b1, b2 = host.Batch(), host.Batch()
        while True:
            n = f1.next_batch(b1, o["N"], ad1)
            if n == 0:
                break
            assert f2.next_batch(b2, o["N"], ad2) == n
            res = {}
            for mate, b, ag in ((1, b1, False), (2, b2, True)):
                seqs, offs = b.arrays()
                reads = [seqs[int(offs[i]):int(offs[i + 1])].tobytes() for i in range(n)]
                ctr = refio.WoCounters()
                ranked, sizes = refio.oracle_pe_mate(hdr, goldenio.se_pair(ag), reads, ag, m=o["m"], b=o["b"],
                                                     top_k=o["k"], counters=ctr)
                res[f"ranked{mate}"], res[f"n{mate}"], res[f"short{mate}"] = ranked, sizes, int(ctr.n_short)
                res[f"len{mate}"] = np.diff(offs.astype(np.int64))
            res["pairs"] = oracle_pairs(hdr, res["ranked1"], res["n1"], res["len1"], res["ranked2"], res["n2"],
                                        res["len2"], o["m"], o["L"])
            w.write(b1, b2, res, n)
            if n < o["N"]:
                break
        w.close(); f1.close(); f2.close(); b1.free(); b2.free()


def compare_dirs(got_dir, case):
    want_dir = os.path.join(CLI, case["name"])
    assert sorted(os.listdir(got_dir)) == case["files"]
    for f in case["files"]:
        got = open(os.path.join(got_dir, f), "rb").read()
        want = open(os.path.join(want_dir, f), "rb").read()
        if got != want:
            gl, wl = got.split(b"\n"), want.split(b"\n")
            for i, (a, b) in enumerate(zip(gl, wl)):
                assert a == b, (case["name"], f, i, a[:300], b[:300])
            assert len(gl) == len(wl), (case["name"], f, len(gl), len(wl))


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_host_text_path_matches_reference(case, tmp_path):
    assert case["returncode"] == 0
    run_case_with_oracle(case, str(tmp_path))
    compare_dirs(str(tmp_path), case)


def test_clip_adaptor_matches_oracle_restatement():
    from walt_b200 import host
    L = host.load_library()
    Lo = refio.oracle_lib()
    rng = np.random.default_rng(3)
    ad = b"AGATCGGAAGAGC"
    for _ in range(2000):
        n = int(rng.integers(14, 120))
        s = bytes(rng.choice(list(b"ACGT"), size=n).tolist())
        if rng.random() < 0.7:
            cut = int(rng.integers(0, n))
            s = (s[:cut] + ad + s)[:n]
        a = C.create_string_buffer(s, n + 1); b = C.create_string_buffer(s, n + 1)
        ra = L.walt_clip_adaptor(ad, a, C.c_size_t(n))
        rb = Lo.wo_clip_adaptor(ad, C.c_size_t(len(ad)), b, C.c_size_t(n))
        assert ra == rb and a.raw == b.raw
