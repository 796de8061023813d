"""The C oracle (oracle/walt_oracle.c) against the golden vectors produced by the unmodified
reference (tests/golden/make_golden.py): this is what PINS the oracle."""
import ctypes as C

import numpy as np
import pytest

import goldenio
import refio


def _cmp_best(got, want):
    for f in ("genome_pos", "times", "mismatch", "strand"):
        assert np.array_equal(got[f], want[f]), f


def _cases(z, prefix="best_"):
    out = []
    for k in z.files:
        if k.startswith(prefix):
            m, b = k[len(prefix):].split("_")
            out.append((k, int(m[1:]), int(b[1:])))
    return out


@pytest.mark.parametrize("name,ag", [("se_ct.npz", False), ("se_ga.npz", True)])
def test_se_oracle_matches_reference(name, ag):
    hdr, _ = goldenio.genome()
    z = goldenio.load(name)
    reads = z["reads"]
    for key, m, b in _cases(z):
        got = refio.oracle_se_map(hdr, goldenio.se_pair(ag), reads, ag=ag, m=m, b=b)
        _cmp_best(got, z[key])


def test_se_edge_oracle_matches_reference():
    hdr, _ = goldenio.genome()
    z = goldenio.load("se_edge.npz")
    buf, offs = z["buf"], z["offs"]
    reads = [buf[int(offs[i]):int(offs[i + 1])].tobytes() for i in range(len(offs) - 1)]
    for ag, pre in ((False, "ct_best_"), (True, "ga_best_")):
        for key, m, b in _cases(z, pre):
            ctr = refio.WoCounters()
            got = refio.oracle_se_map(hdr, goldenio.se_pair(ag), reads, ag=ag, m=m, b=b, counters=ctr)
            _cmp_best(got, z[key])
            short = int(z[key.replace("best", "short")])
            assert ctr.n_short == short


def test_pe_oracle_matches_reference():
    hdr, _ = goldenio.genome()
    z = goldenio.load("pe.npz")
    for m, k in ((6, 50), (8, 3), (4, 2)):
        for mate, ag in ((1, False), (2, True)):
            ranked, sizes = refio.oracle_pe_mate(hdr, goldenio.se_pair(ag), z[f"m{mate}"], ag, m=m, top_k=k)
            assert np.array_equal(sizes, z[f"sizes{mate}_m{m}_k{k}"])
            want = z[f"ranked{mate}_m{m}_k{k}"]
            for f in ("genome_pos", "mismatch", "strand"):
                assert np.array_equal(ranked[f], want[f]), (m, k, mate, f)


def test_counter_rederivation_matches_oracle_builder():
    """counter[] rebuilt from index[] must equal what the oracle's makedb restatement
    produces from the converted sequence, and the oracle's index[] must be the reference
    makedb's byte for byte."""
    hdr, subs = goldenio.genome()
    L = refio.oracle_lib()
    starts = np.ascontiguousarray(hdr.start_index, np.uint32)
    for sfx, sub in subs.items():
        counter = np.zeros(refio.N_KEYS + 1, np.uint32)
        index = np.zeros(hdr.genome_len, np.uint32)
        n = L.wo_build_index(sub.seq.ctypes.data_as(C.c_void_p), C.c_uint64(hdr.genome_len),
                             C.c_uint32(len(hdr.lengths)), starts.ctypes.data_as(C.c_void_p),
                             counter.ctypes.data_as(C.c_void_p), index.ctypes.data_as(C.c_void_p))
        assert n == sub.index.size
        assert np.array_equal(counter, sub.counter)
        # identical, ties included: the oracle replays libstdc++'s std::sort (the golden index came
        # from the reference makedb and has tied suffixes: exact repeats, look-alike chromosome ends)
        assert np.array_equal(index[:n], sub.index)
        cmp = L.wo_bucket_cmp
        a = sub.index[:-1].astype(np.uint32)
        b = sub.index[1:].astype(np.uint32)
        ties = 0
        for i in range(0, a.size, 7):   # a sample of adjacent slots: how many are ties?
            ties += cmp(sub.seq.ctypes.data_as(C.c_void_p), C.c_uint32(len(hdr.lengths)),
                        starts.ctypes.data_as(C.c_void_p), C.c_uint32(int(a[i])), C.c_uint32(int(b[i]))) == 0
        assert ties > 0, "the golden genome no longer has tied suffixes: tie order is untested"
