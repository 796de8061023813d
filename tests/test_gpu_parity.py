"""GPU parity tests proper: the CUDA engine, called through the C ABI, against (a) the golden
vectors produced by the unmodified reference and (b) the C oracle on fresh seeded inputs.
Bit-exact: positions, strands, mismatch counts, times, heap contents and order."""
import ctypes as C
import os

import numpy as np
import pytest

import goldenio
import refio
import synth

pytestmark = pytest.mark.gpu


def _cmp_best(got, want, what=""):
    for f in ("genome_pos", "times", "mismatch", "strand"):
        bad = np.nonzero(got[f] != want[f])[0]
        assert bad.size == 0, (what, f, bad[:5], got[bad[:5]], want[bad[:5]])


def _load_golden_engine(depth=0):
    import walt_b200
    hdr, subs = goldenio.genome()
    e = walt_b200.Engine(0)
    if depth:
        e.set_table_depth(depth)
    e.set_chromosomes(hdr.lengths, hdr.names)
    for w, sfx in enumerate(refio.SUFFIXES):
        e.load_subindex(w, subs[sfx].seq, subs[sfx].counter, subs[sfx].index)
    return e


@pytest.fixture(scope="module")
def engine():
    e = _load_golden_engine()
    yield e
    e.close()


def test_native_library_is_loaded(engine):
    maps = open("/proc/self/maps").read()
    assert "libwaltb200.so" in maps
    assert engine.hbm_bytes() > 0
    info = engine.subindex_info(0)
    assert info["depth"] >= 12 and info["n_taint"] > 0


@pytest.mark.parametrize("literal,width", [(False, 8), (True, 8), (False, 16), (False, 32), (True, 32)])
def test_se_golden(engine, literal, width):
    engine.set_search_mode(literal)
    engine.set_group_width(width)
    try:
        for name, ag in (("se_ct.npz", False), ("se_ga.npz", True)):
            z = goldenio.load(name)
            buf, offs = refio.pack_reads(z["reads"])
            for key in [k for k in z.files if k.startswith("best_")]:
                m, b = (int(x[1:]) for x in key[5:].split("_"))
                out, short = engine.map_se(buf, offs, ag=ag, m=m, b=b)
                _cmp_best(out, z[key], (name, key))
                assert short == int(z[key.replace("best", "short")])
                st = engine.stats()
                assert st["n_kernel_launches"] >= 1 and st["n_lookups"] > 0
                if literal:
                    assert st["n_literal"] == st["n_lookups"]
    finally:
        engine.set_search_mode(False)
        engine.set_group_width(8)


@pytest.mark.parametrize("depth,width", [(0, 8), (12, 8), (14, 16), (17, 32), (12, 32), (15, 8)])
def test_se_edge_golden(depth, width):
    e = _load_golden_engine(depth)
    e.set_group_width(width)
    z = goldenio.load("se_edge.npz")
    for ag, pre in ((False, "ct_best_"), (True, "ga_best_")):
        for key in [k for k in z.files if k.startswith(pre)]:
            m, b = (int(x[1:]) for x in key[len(pre):].split("_"))
            out, short = e.map_se(z["buf"], z["offs"], ag=ag, m=m, b=b)
            _cmp_best(out, z[key], (depth, key))
            assert short == int(z[key.replace("best", "short")])
    e.close()


def test_se_chunked_pipeline_matches_single_chunk(engine):
    z = goldenio.load("se_ct.npz")
    buf, offs = refio.pack_reads(z["reads"])
    engine.set_chunk_reads(257)
    try:
        out, _ = engine.map_se(buf, offs, m=6, b=5000)
    finally:
        engine.set_chunk_reads(0)
    _cmp_best(out, z["best_m6_b5000"])


@pytest.mark.parametrize("width", [8, 16, 32])
def test_se_packed_input_golden(engine, width):
    """walt_engine_map_se_packed (2-bit reads over PCIe) == the reference, equal-length and ragged
    batches, one chunk and many"""
    from walt_b200 import host
    engine.set_group_width(width)
    try:
        for name, ag in (("se_ct.npz", False), ("se_ga.npz", True)):
            z = goldenio.load(name)
            buf, offs = refio.pack_reads(z["reads"])
            packed = host.pack_reads_2bit(buf, offs)
            for key in [k for k in z.files if k.startswith("best_")]:
                m, b = (int(x[1:]) for x in key[5:].split("_"))
                out, short = engine.map_se_packed(packed, offs, ag=ag, m=m, b=b)
                _cmp_best(out, z[key], (name, key))
                assert short == int(z[key.replace("best", "short")])
            engine.set_chunk_reads(257)
            out, _ = engine.map_se_packed(packed, offs, ag=ag, m=6, b=5000)
            engine.set_chunk_reads(0)
            _cmp_best(out, z["best_m6_b5000"], "chunked")
        z = goldenio.load("se_edge.npz")
        packed = host.pack_reads_2bit(z["buf"], z["offs"])
        for chunk in (1 << 18, 7):
            engine.set_chunk_reads(chunk)
            for ag, pre in ((False, "ct_best_"), (True, "ga_best_")):
                for key in [k for k in z.files if k.startswith(pre)]:
                    m, b = (int(x[1:]) for x in key[len(pre):].split("_"))
                    out, short = engine.map_se_packed(packed, z["offs"], ag=ag, m=m, b=b)
                    _cmp_best(out, z[key], (chunk, key))
                    assert short == int(z[key.replace("best", "short")])
    finally:
        engine.set_group_width(8)
        engine.set_chunk_reads(0)


def test_se_empty_and_errors(engine):
    import walt_b200
    out, short = engine.map_se(np.zeros(1, np.uint8), np.zeros(1, np.uint64))
    assert out.size == 0 and short == 0
    buf, offs = refio.pack_reads([b"ACGTN" * 20])
    with pytest.raises(walt_b200.WaltError) as ei:
        engine.map_se(buf, offs)
    assert ei.value.code == 5   # WALT_ENONACGT, util.hpp:117-120
    # engine stays usable afterwards
    z = goldenio.load("se_ct.npz")
    buf, offs = refio.pack_reads(z["reads"][:50])
    out, _ = engine.map_se(buf, offs)
    _cmp_best(out, z["best_m6_b5000"][:50])


@pytest.mark.parametrize("width", [8, 32])
def test_pe_golden(engine, width):
    engine.set_group_width(width)
    hdr, _ = goldenio.genome()
    z = goldenio.load("pe.npz")
    L = refio.oracle_lib()
    starts = np.ascontiguousarray(hdr.start_index, np.uint32)
    lengths = np.ascontiguousarray(hdr.lengths, np.uint32)
    chroms = refio.WoChroms(len(lengths), starts.ctypes.data, lengths.ctypes.data)
    b1, o1 = refio.pack_reads(z["m1"])
    b2, o2 = refio.pack_reads(z["m2"])
    for m, k in ((6, 50), (8, 3), (4, 2)):
        for frag_range in (1000, 250):
            r = engine.map_pe(b1, o1, b2, o2, m=m, top_k=k, frag_range=frag_range)
            for mate in (1, 2):
                assert np.array_equal(r[f"n{mate}"], z[f"sizes{mate}_m{m}_k{k}"])
                want = z[f"ranked{mate}_m{m}_k{k}"]
                for f in ("genome_pos", "mismatch", "strand"):
                    assert np.array_equal(r[f"ranked{mate}"][f], want[f]), (m, k, mate, f)
            pr = r["pairs"]
            for j in range(len(pr)):
                bi, bj = C.c_int32(-1), C.c_int32(-1)
                r1 = np.ascontiguousarray(r["ranked1"][j]); r2 = np.ascontiguousarray(r["ranked2"][j])
                t = L.wo_pe_pair(C.byref(chroms), r1.ctypes.data_as(C.c_void_p), C.c_uint32(int(r["n1"][j])),
                                 C.c_uint32(100), r2.ctypes.data_as(C.c_void_p), C.c_uint32(int(r["n2"][j])),
                                 C.c_uint32(100), C.c_uint32(m), C.c_int(frag_range), C.byref(bi), C.byref(bj))
                assert (t, bi.value, bj.value) == (pr[j]["best_times"], pr[j]["best_i"], pr[j]["best_j"]), j


@pytest.mark.parametrize("m,k", [(6, 100), (16, 50), (15, 70), (2, 300)])
def test_pe_heap_variants_vs_oracle(engine, m, k):
    """The corners of the two-phase paired-end form against the C oracle: top_k above the local-heap
    capacity of the replay kernel (heaps in global scratch), the largest -m the candidate log
    supports (15) and the single-kernel fallback beyond it (16)."""
    hdr, _ = goldenio.genome()
    z = goldenio.load("pe.npz")
    for width in (8, 32):
        engine.set_group_width(width)
        b1, o1 = refio.pack_reads(z["m1"])
        b2, o2 = refio.pack_reads(z["m2"])
        r = engine.map_pe(b1, o1, b2, o2, m=m, top_k=k, frag_range=1000)
        c, _, _ = engine.map_pe_compact(b1, o1, b2, o2, m=m, top_k=k, frag_range=1000)
        assert np.array_equal(c["pair"], r["pairs"])
        for mate, ag in ((1, False), (2, True)):
            ranked, sizes = refio.oracle_pe_mate(hdr, goldenio.se_pair(ag), z[f"m{mate}"], ag, m=m, top_k=k)
            assert np.array_equal(sizes, r[f"n{mate}"]), (m, k, mate)
            for f in ("genome_pos", "mismatch", "strand"):
                assert np.array_equal(ranked[f], r[f"ranked{mate}"][f]), (m, k, mate, f)
    engine.set_group_width(8)


@pytest.fixture(scope="module")
def repeat_world():
    """high-copy repeat families: hundreds of candidates per lookup (tests/test_repeat_emu.py runs the
    same genome on the CPU harness)"""
    import walt_b200
    chroms = synth.make_repeat_genome([90000, 60000, 30000], seed=23, n_families=5, fam_len=(170, 420),
                                      copies=(60, 260), divergence=0.006, repeat_frac=0.7)
    hdr, subs = refio.build_index_with_oracle(chroms)
    e = walt_b200.Engine(0)
    e.set_chromosomes(hdr.lengths, hdr.names)
    for w, sfx in enumerate(refio.SUFFIXES):
        e.load_subindex(w, subs[sfx].seq, subs[sfx].counter, subs[sfx].index)
    yield chroms, hdr, subs, e
    e.close()


def _acgt(reads):
    out = reads.copy()
    out[out == ord("N")] = ord("A")
    return out


@pytest.mark.parametrize("defer,width", [(1, 8), (0, 8), (1, 32), (0, 32), (1, 16)])
def test_se_repeats_vs_oracle(repeat_world, defer, width):
    """parked reads finished by the warp-per-read kernel (quad verification) == every read finished
    by its group == the oracle"""
    chroms, hdr, subs, e = repeat_world
    e.set_defer(defer)
    e.set_group_width(width)
    try:
        for rl, m, b in ((150, 6, 5000), (150, 6, 40), (100, 4, 5000), (200, 8, 5000), (192, 6, 5000), (193, 6, 5000)):
            for ag, pair in ((False, ("_CT00", "_CT01")), (True, ("_GA10", "_GA11"))):
                reads = _acgt(synth.simulate_se_reads(chroms, 3000, rl, seed=31 + rl + m, a_rich=ag))
                ctr = refio.WoCounters()
                want = refio.oracle_se_map(hdr, tuple(subs[s] for s in pair), reads, ag=ag, m=m, b=b, counters=ctr)
                buf, offs = refio.pack_reads(reads)
                got, _ = e.map_se(buf, offs, ag=ag, m=m, b=b)
                _cmp_best(got, want, (defer, width, rl, m, b, ag))
                if b == 5000:
                    assert ctr.asdict()["n_cand"] > 20 * len(reads)
                if defer:
                    st = e.stats()
                    assert st["n_kernel_launches"] == 4 and 0 < st["n_parked"] < len(reads)
    finally:
        e.set_defer(1)
        e.set_group_width(8)


@pytest.mark.parametrize("defer,width", [(1, 8), (0, 8), (1, 32)])
def test_pe_repeats_vs_oracle(repeat_world, defer, width):
    chroms, hdr, subs, e = repeat_world
    e.set_defer(defer)
    e.set_group_width(width)
    try:
        m1, m2 = synth.simulate_pe_reads(chroms, 3000, 150, seed=41)
        m1, m2 = _acgt(m1), _acgt(m2)
        b1, o1 = refio.pack_reads(m1)
        b2, o2 = refio.pack_reads(m2)
        for m, k, b in ((8, 50, 5000), (6, 10, 5000), (8, 50, 60), (3, 2, 5000), (15, 20, 5000), (8, 100, 5000)):
            r = e.map_pe(b1, o1, b2, o2, m=m, b=b, top_k=k, frag_range=1000)
            for mate, reads, ag, pair in ((1, m1, False, ("_CT00", "_CT01")), (2, m2, True, ("_GA10", "_GA11"))):
                want, sizes = refio.oracle_pe_mate(hdr, tuple(subs[s] for s in pair), reads, ag, m=m, b=b, top_k=k)
                assert np.array_equal(r[f"n{mate}"], sizes), (m, k, b, mate)
                for f in ("genome_pos", "mismatch", "strand"):
                    assert np.array_equal(r[f"ranked{mate}"][f], want[f]), (m, k, b, mate, f)
    finally:
        e.set_defer(1)
        e.set_group_width(8)


def test_pe_pbat_is_mate_swap(engine):
    z = goldenio.load("pe.npz")
    b1, o1 = refio.pack_reads(z["m1"])
    b2, o2 = refio.pack_reads(z["m2"])
    a = engine.map_pe(b1, o1, b2, o2, m=6, top_k=10, frag_range=600)
    s = engine.map_pe(b2, o2, b1, o1, m=6, top_k=10, frag_range=600, pbat=True)
    assert np.array_equal(a["ranked1"], s["ranked2"]) and np.array_equal(a["ranked2"], s["ranked1"])
    assert np.array_equal(a["pairs"]["best_i"], s["pairs"]["best_j"])
    assert np.array_equal(a["pairs"]["best_times"], s["pairs"]["best_times"])


def test_se_fresh_genome_vs_oracle():
    """A bigger, freshly seeded case: index built by the oracle's makedb restatement, mapped
    by the oracle and by the engine (default table depth and a forced shallow one)."""
    import walt_b200
    lengths = [400000, 250000, 90000, 1000, 37]
    chroms = synth.make_repeat_genome(lengths, seed=101, n_families=10, fam_len=(200, 1500),
                                      copies=(5, 40), divergence=0.02, repeat_frac=0.3)
    hdr, subs = refio.build_index_with_oracle(chroms)
    reads100 = synth.simulate_se_reads(chroms[:3], 20000, 100, seed=7, n_frac=0.0)
    reads150a = synth.simulate_se_reads(chroms[:3], 10000, 150, seed=8, a_rich=True, n_frac=0.0)
    for depth, width in ((0, 8), (13, 8), (13, 16), (0, 32)):
        e = walt_b200.Engine(0)
        e.set_table_depth(depth)
        e.set_group_width(width)
        e.set_chromosomes(hdr.lengths, hdr.names)
        for w, sfx in enumerate(refio.SUFFIXES):
            e.load_subindex(w, subs[sfx].seq, subs[sfx].counter, subs[sfx].index)
        for reads, ag, m, b in ((reads100, False, 6, 5000), (reads150a, True, 6, 5000), (reads100, False, 8, 20)):
            want = refio.oracle_se_map(hdr, (subs["_GA10"], subs["_GA11"]) if ag else (subs["_CT00"], subs["_CT01"]),
                                       reads, ag=ag, m=m, b=b)
            buf, offs = refio.pack_reads(reads)
            got, _ = e.map_se(buf, offs, ag=ag, m=m, b=b)
            _cmp_best(got, want, (depth, ag, m, b))
            assert (got["times"] == 1).mean() > 0.5
        e.close()


def test_device_makedb_matches_oracle_builder():
    """walt_engine_build_from_sequence (makedb on the GPU) against the oracle's restatement of
    CountBucketSize/HashToBucket/SortHashTableBucket: same converted genomes, same counter[],
    same index[] -- tied suffixes (the genome has exact repeats) in std::sort's order in both: the
    oracle runs its C restatement of libstdc++'s introsort with the genome comparator, the device
    radix-sorts and replays the introsort on class ranks -- then mapping parity on top of it."""
    import walt_b200
    lengths = [300000, 180000, 36, 37, 120, 35, 60000]
    chroms = synth.make_repeat_genome(lengths, seed=303, n_families=8, fam_len=(200, 1200),
                                      copies=(5, 40), divergence=0.0, repeat_frac=0.3)
    hdr, subs = refio.build_index_with_oracle(chroms)
    e = walt_b200.Engine(0)
    e.set_chromosomes(hdr.lengths, hdr.names)
    e.build_from_sequence(np.concatenate([s for _, s in chroms]))
    info = e.last_build_info()
    assert info["tied_slots"] > 0 and info["buckets_replayed"] > 0, info
    by_std_sort = {}
    for w, sfx in enumerate(refio.SUFFIXES):
        seq, counter, index = e.export_subindex(w, hdr.genome_len)
        assert np.array_equal(seq, subs[sfx].seq), sfx
        assert np.array_equal(counter, subs[sfx].counter), sfx
        assert index.size == subs[sfx].index.size
        assert np.array_equal(index, subs[sfx].index), sfx
        by_std_sort[sfx] = (counter, index)
    # tie order 1 = ascending position: same buckets, same classes, ties ascending
    e2 = walt_b200.Engine(0)
    e2.set_chromosomes(hdr.lengths, hdr.names)
    e2.set_tie_order(True)
    e2.build_from_sequence(np.concatenate([s for _, s in chroms]), which=(0,))
    assert e2.last_build_info()["buckets_replayed"] == 0
    _, counter, index = e2.export_subindex(0, hdr.genome_len, want_seq=False)
    ref_counter, ref_index = by_std_sort["_CT00"]
    assert np.array_equal(counter, ref_counter) and not np.array_equal(index, ref_index)
    for k in np.nonzero(np.diff(counter.astype(np.int64)) > 1)[0][:20000]:
        lo, hi = int(counter[k]), int(counter[k + 1])
        assert np.array_equal(np.sort(index[lo:hi]), np.sort(ref_index[lo:hi]))
    e2.close()
    reads = synth.simulate_se_reads([chroms[0], chroms[1], chroms[6]], 5000, 100, seed=9, n_frac=0.0)
    want = refio.oracle_se_map(hdr, (subs["_CT00"], subs["_CT01"]), reads)
    buf, offs = refio.pack_reads(reads)
    got, _ = e.map_se(buf, offs)
    _cmp_best(got, want)
    e.close()


def test_device_makedb_erases_large_buckets():
    """A genome with a > 500000-fold repeated 12-mer: the bucket must vanish as in
    reference.cpp:211-218 (checked against the oracle builder)."""
    import walt_b200
    rng = np.random.default_rng(5)
    n = 1_700_000
    seq = synth._ACGT[rng.integers(0, 4, size=n)].copy()
    seq[100000:1_300_000] = ord("T")            # poly-T: one 12-mer ~1.2M times (C->T and G->A alike)
    chroms = [("chr1", seq)]
    hdr, subs = refio.build_index_with_oracle(chroms)
    assert subs["_CT00"].index.size < n - 1_000_000
    e = walt_b200.Engine(0)
    e.set_chromosomes(hdr.lengths, hdr.names)
    e.build_from_sequence(seq, which=(0, 1))
    for w, sfx in ((0, "_CT00"), (1, "_CT01")):
        _, counter, index = e.export_subindex(w, hdr.genome_len, want_seq=False)
        assert np.array_equal(counter, subs[sfx].counter)
        assert np.array_equal(index, subs[sfx].index)
    e.close()


def _single_best(ranked, n, m):
    """GetBestMatch4Single, paired.cpp:296-318 (python restatement for the test)."""
    pos, times, mm, strand = 0, 0, m, b"+"
    for i in range(int(n) - 1, -1, -1):
        c = ranked[i]
        if c["mismatch"] < mm:
            pos, times, mm, strand = int(c["genome_pos"]), 1, int(c["mismatch"]), c["strand"]
        elif c["mismatch"] == mm:
            if pos == int(c["genome_pos"]):
                continue
            pos, strand, times = int(c["genome_pos"]), c["strand"], times + 1
        else:
            break
    return pos, times, mm, strand


@pytest.mark.parametrize("pbat", [False, True])
def test_pe_compact_and_device_paths_match_ranked(engine, pbat):
    """walt_engine_map_pe_compact / _device return, per pair, exactly what the output stage derives
    from the ranked lists of walt_engine_map_pe (pairing result, the winning candidates, each
    mate's GetBestMatch4Single)."""
    import torch
    from walt_b200.engine import PE_RESULT_DT
    z = goldenio.load("pe.npz")
    b1, o1 = refio.pack_reads(z["m1"])
    b2, o2 = refio.pack_reads(z["m2"])
    for m, k, L in ((6, 50, 1000), (8, 3, 250)):
        full = engine.map_pe(b1, o1, b2, o2, m=m, top_k=k, frag_range=L, pbat=pbat)
        comp, s1, s2 = engine.map_pe_compact(b1, o1, b2, o2, m=m, top_k=k, frag_range=L, pbat=pbat)
        assert (s1, s2) == (full["short1"], full["short2"])
        assert np.array_equal(comp["pair"], full["pairs"])
        for j in range(len(comp)):
            pr = full["pairs"][j]
            if pr["best_times"] >= 1:
                assert comp["c1"][j] == full["ranked1"][j][pr["best_i"]], j
                assert comp["c2"][j] == full["ranked2"][j][pr["best_j"]], j
            else:
                assert comp["c1"][j]["genome_pos"] == 0 and comp["c2"][j]["genome_pos"] == 0
            for mate in (1, 2):
                pos, times, mm, strand = _single_best(full[f"ranked{mate}"][j], full[f"n{mate}"][j], m)
                got = comp[f"single{mate}"][j]
                assert (int(got["genome_pos"]), int(got["times"]), int(got["mismatch"]), got["strand"]) == \
                       (pos, times, mm, strand), (j, mate)
        # 2-bit packed mates over PCIe
        from walt_b200 import host
        for chunk in (1 << 18, 61):
            engine.set_chunk_reads(chunk)
            try:
                pk, p1, p2 = engine.map_pe_compact_packed(host.pack_reads_2bit(b1, o1), o1, host.pack_reads_2bit(b2, o2),
                                                          o2, m=m, top_k=k, frag_range=L, pbat=pbat)
            finally:
                engine.set_chunk_reads(0)
            assert (p1, p2) == (s1, s2) and np.array_equal(pk, comp), chunk
        # device-resident path
        dev = "cuda:0"
        d1 = torch.from_numpy(b1).to(dev); d2 = torch.from_numpy(b2).to(dev)
        do1 = torch.from_numpy(o1.astype(np.int64)).to(dev); do2 = torch.from_numpy(o2.astype(np.int64)).to(dev)
        n = len(o1) - 1
        d_out = torch.zeros(n * PE_RESULT_DT.itemsize, dtype=torch.uint8, device=dev)
        engine.map_pe_device(d1.data_ptr(), do1.data_ptr(), d2.data_ptr(), do2.data_ptr(), n, 100, d_out.data_ptr(),
                             m=m, top_k=k, frag_range=L, pbat=pbat, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        got = d_out.cpu().numpy().view(PE_RESULT_DT)
        assert np.array_equal(got, comp)


def test_group_and_clone_match_single_engine():
    """walt_group (several engines behind one handle) with replicas made by walt_engine_clone_index: every
    shard lands in its slice, results equal one engine's.  The engines share device 0 on a one-GPU box and
    take distinct devices when there are more."""
    import torch
    import walt_b200
    from walt_b200 import host
    hdr, subs = goldenio.genome()
    n_dev = max(1, torch.cuda.device_count())
    devs = [i % n_dev for i in range(3)]
    g = walt_b200.Group(devs)
    e0 = g.engines[0]
    e0.set_chromosomes(hdr.lengths, hdr.names)
    for w, sfx in enumerate(refio.SUFFIXES):
        e0.load_subindex(w, subs[sfx].seq, subs[sfx].counter, subs[sfx].index)
    for e in g.engines[1:]:
        e.clone_index_from(e0)
        assert e.hbm_bytes() == e0.hbm_bytes() and e.subindex_info(3) == e0.subindex_info(3)
    z = goldenio.load("se_ct.npz")
    buf, offs = refio.pack_reads(z["reads"])
    out, short = g.map_se(buf, offs, m=6, b=5000)
    _cmp_best(out, z["best_m6_b5000"], "group ascii")
    assert short == int(z["short_m6_b5000"])
    out, _ = g.map_se_packed(host.pack_reads_2bit(buf, offs), offs, m=6, b=5000)
    _cmp_best(out, z["best_m6_b5000"], "group packed")
    ze = goldenio.load("se_edge.npz")     # ragged lengths
    key = [k for k in ze.files if k.startswith("ct_best_")][0]
    m, b = (int(x[1:]) for x in key[len("ct_best_"):].split("_"))
    out, _ = g.map_se(ze["buf"], ze["offs"], m=m, b=b)
    _cmp_best(out, ze[key], "group ragged")
    zp = goldenio.load("pe.npz")
    b1, o1 = refio.pack_reads(zp["m1"]); b2, o2 = refio.pack_reads(zp["m2"])
    want, _, _ = e0.map_pe_compact(b1, o1, b2, o2, m=6, top_k=50, frag_range=1000)
    got, _, _ = g.map_pe_compact_packed(host.pack_reads_2bit(b1, o1), o1, host.pack_reads_2bit(b2, o2), o2, m=6, top_k=50,
                                        frag_range=1000)
    assert np.array_equal(got, want)
    g.close()
