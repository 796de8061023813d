"""The repeat path of walt_core.cuh -- long fingerprint runs streamed through the warp-wide quad
verification (verify_run_wide), the batch-parallel LogSink -- stepped on the CPU fiber harness
against the C oracle on a genome made of high-copy repeat families (hundreds of candidates per
lookup, -b on and off the limit), for every group width."""
import numpy as np
import pytest

import refio
import synth
from emu import emu


@pytest.fixture(scope="module")
def world():
    # 3 chromosomes; families with up to a few hundred copies and little divergence, so narrowed
    # regions run from a handful to several hundred slots
    chroms = synth.make_repeat_genome([90000, 60000, 30000], seed=23, n_families=5, fam_len=(170, 420),
                                      copies=(60, 260), divergence=0.006, repeat_frac=0.7)
    hdr, subs = refio.build_index_with_oracle(chroms)
    e = emu.EmuEngine(hdr.lengths)
    for w, sfx in enumerate(refio.SUFFIXES):
        assert e.load(w, subs[sfx].seq, subs[sfx].index) == 0
    yield chroms, hdr, subs, e
    e.close()


def _acgt(reads):
    out = reads.copy()
    out[out == ord("N")] = ord("A")
    return out


def _cmp_best(got, want):
    for f in ("genome_pos", "times", "mismatch", "strand"):
        bad = np.nonzero(got[f] != want[f])[0]
        assert bad.size == 0, (f, bad[:5], got[bad[:5]], want[bad[:5]])


@pytest.mark.parametrize("width,literal", [(32, False), (8, False), (32, True), (8, True)])
@pytest.mark.parametrize("rl,m,b", [(150, 6, 5000), (150, 6, 40), (100, 4, 5000), (200, 8, 5000), (45, 3, 5000)])
def test_se_repeats(world, width, literal, rl, m, b):
    chroms, hdr, subs, e = world
    for ag, pair in ((False, ("_CT00", "_CT01")), (True, ("_GA10", "_GA11"))):
        reads = _acgt(synth.simulate_se_reads(chroms, 160, rl, seed=31 + rl + m, a_rich=ag))
        ctr = refio.WoCounters()
        want = refio.oracle_se_map(hdr, tuple(subs[s] for s in pair), reads, ag=ag, m=m, b=b, counters=ctr)
        if b == 5000:
            assert ctr.asdict()["n_cand"] > (20 if rl >= 100 else 5) * len(reads)      # the workload is repeat-bound
        buf, offs = refio.pack_reads(reads)
        # literal: every lookup through literal_index_region (table boundaries / fingerprints / genome)
        rc, got, _ = e.map_se(buf, offs, refio.BEST_DT, ag=ag, m=m, b=b, width=width, literal=literal)
        assert rc == 0
        _cmp_best(got, want)


@pytest.mark.parametrize("width", [32, 8])
@pytest.mark.parametrize("m,k,b", [(8, 50, 5000), (6, 10, 5000), (8, 50, 60), (3, 2, 5000), (15, 20, 5000)])
def test_pe_repeats_logged(world, width, m, k, b):
    chroms, hdr, subs, e = world
    m1, m2 = synth.simulate_pe_reads(chroms, 120, 150, seed=41 + m)
    for reads, ag, pair in ((_acgt(m1), False, ("_CT00", "_CT01")), (_acgt(m2), True, ("_GA10", "_GA11"))):
        want, sizes = refio.oracle_pe_mate(hdr, tuple(subs[s] for s in pair), reads, ag, m=m, b=b, top_k=k)
        buf, offs = refio.pack_reads(reads)
        rc, ranked, got_sizes = e.map_pe_mate(buf, offs, refio.CAND_DT, ag, m=m, b=b, top_k=k, width=width, logged=True)
        assert rc == 0
        assert np.array_equal(got_sizes, sizes)
        for f in ("genome_pos", "mismatch", "strand"):
            assert np.array_equal(ranked[f], want[f]), (m, k, b, f)
        if b == 5000 and k >= 10:
            assert (sizes == k).mean() > 0.1                      # heaps do fill up: the full-heap rule is exercised


@pytest.mark.parametrize("m,k,L", [(8, 50, 1000), (6, 10, 400), (8, 50, 160), (3, 2, 1000), (8, 100, 1000)])
def test_pairing_by_warp_equals_reference_loop(world, m, k, L):
    """pair_candidates_wide (a warp per pair: three reductions over the valid pairs) against the oracle's
    restatement of the MergePairedEndResults loop (paired.cpp:472-513) and the thread-per-pair form,
    on ranked lists full of repeats (ties, duplicates, sums equal to -m)."""
    import ctypes as C
    chroms, hdr, subs, e = world
    m1, m2 = synth.simulate_pe_reads(chroms, 150, 150, seed=77)
    got = {}
    for mate, reads, ag, pair in ((1, _acgt(m1), False, ("_CT00", "_CT01")), (2, _acgt(m2), True, ("_GA10", "_GA11"))):
        ranked, sizes = refio.oracle_pe_mate(hdr, tuple(subs[s] for s in pair), reads, ag, m=m, top_k=k)
        got[mate] = (ranked, sizes, refio.pack_reads(reads)[1])
    a = e.pair(got[1][0], got[1][1], got[1][2], got[2][0], got[2][1], got[2][2], k, m, L)
    b = e.pair_wide(got[1][0], got[1][1], got[1][2], got[2][0], got[2][1], got[2][2], k, m, L)
    assert np.array_equal(a, b), np.nonzero(a != b)[0][:10]
    if k >= 50 and L == 1000:
        assert (a["best_times"] >= 2).sum() > 3 and (a["best_times"] == 1).sum() > 3
    Lo = refio.oracle_lib()
    starts = np.ascontiguousarray(hdr.start_index, np.uint32)
    lengths = np.ascontiguousarray(hdr.lengths, np.uint32)
    ch = refio.WoChroms(len(lengths), starts.ctypes.data, lengths.ctypes.data)
    for j in range(len(b)):
        bi, bj = C.c_int32(-1), C.c_int32(-1)
        r1 = np.ascontiguousarray(got[1][0][j]); r2 = np.ascontiguousarray(got[2][0][j])
        t = Lo.wo_pe_pair(C.byref(ch), r1.ctypes.data_as(C.c_void_p), C.c_uint32(int(got[1][1][j])), C.c_uint32(150),
                          r2.ctypes.data_as(C.c_void_p), C.c_uint32(int(got[2][1][j])), C.c_uint32(150), C.c_uint32(m),
                          C.c_int(L), C.byref(bi), C.byref(bj))
        assert (t, bi.value, bj.value) == (b[j]["best_times"], b[j]["best_i"], b[j]["best_j"]), j
