"""walt_core.cuh (the kernels' lane-level source) stepped on the CPU fiber harness
(tests/emu) against the golden vectors: table search, taint handling, ordered folds, heap and
pairing logic, all without a device."""
import ctypes as C

import numpy as np
import pytest

import goldenio
import refio
from emu import emu


@pytest.fixture(scope="module")
def engine():
    hdr, subs = goldenio.genome()
    e = emu.EmuEngine(hdr.lengths)
    for w, sfx in enumerate(refio.SUFFIXES):
        assert e.load(w, subs[sfx].seq, subs[sfx].index) == 0
    yield e
    e.close()


def _cmp_best(got, want):
    for f in ("genome_pos", "times", "mismatch", "strand"):
        bad = np.nonzero(got[f] != want[f])[0]
        assert bad.size == 0, (f, bad[:5], got[bad[:5]], want[bad[:5]])


@pytest.mark.parametrize("literal,width", [(False, 32), (True, 32), (False, 8), (True, 8), (False, 16)])
def test_se_emu_matches_reference(engine, literal, width):
    for name, ag in (("se_ct.npz", False), ("se_ga.npz", True)):
        z = goldenio.load(name)
        buf, offs = refio.pack_reads(z["reads"])
        for key in [k for k in z.files if k.startswith("best_")]:
            m, b = (int(x[1:]) for x in key[5:].split("_"))
            rc, out, ctr = engine.map_se(buf, offs, refio.BEST_DT, ag=ag, m=m, b=b, literal=literal, width=width)
            assert rc == 0
            _cmp_best(out, z[key])
            if literal:
                assert ctr[2] == ctr[0] > 0


def _pack_numpy(buf, offs):
    """independent statement of the walt_pack_reads layout (include/walt_host.h)"""
    code = np.zeros(256, np.uint8)
    for i, c in enumerate(b"ACGT"):
        code[c] = i
    n = len(offs) - 1
    out = np.zeros((int(offs[n]) >> 2) + n + 16, np.uint8)
    for j in range(n):
        o0, o1 = int(offs[j]), int(offs[j + 1])
        c = code[buf[o0:o1]]
        c = np.concatenate([c, np.zeros((-len(c)) % 4, np.uint8)]).reshape(-1, 4)
        b = (c[:, 0] << 6) | (c[:, 1] << 4) | (c[:, 2] << 2) | c[:, 3]
        out[(o0 >> 2) + j:(o0 >> 2) + j + len(b)] |= b.astype(np.uint8)   # |= : overlap would corrupt
    return out


def test_pack_reads_layout():
    from walt_b200 import host
    z = goldenio.load("se_edge.npz")          # ragged lengths
    got = host.pack_reads_2bit(z["buf"], z["offs"])
    assert np.array_equal(got, _pack_numpy(z["buf"], z["offs"]))
    buf, offs = refio.pack_reads([b"ACGTN" * 20])
    with pytest.raises(host.HostError):
        host.pack_reads_2bit(buf, offs)
    assert host.pack_reads_2bit(np.zeros(1, np.uint8), np.zeros(1, np.uint64)).sum() == 0


@pytest.mark.parametrize("width", [8, 16, 32])
def test_se_emu_packed_input(engine, width):
    """the kernels' PACKED read loader (2-bit reads over PCIe) gives the ASCII loader's results"""
    from walt_b200 import host
    for name, ag in (("se_ct.npz", False), ("se_ga.npz", True)):
        z = goldenio.load(name)
        buf, offs = refio.pack_reads(z["reads"])
        rc, out, _ = engine.map_se(host.pack_reads_2bit(buf, offs), offs, refio.BEST_DT, ag=ag, m=6, b=5000,
                                   width=width, packed=True)
        assert rc == 0
        _cmp_best(out, z["best_m6_b5000"])
    z = goldenio.load("se_edge.npz")
    key = [k for k in z.files if k.startswith("ct_best_")][0]
    m, b = (int(x[1:]) for x in key[len("ct_best_"):].split("_"))
    rc, out, _ = engine.map_se(host.pack_reads_2bit(z["buf"], z["offs"]), z["offs"], refio.BEST_DT, m=m, b=b,
                               width=width, packed=True)
    assert rc == 0
    _cmp_best(out, z[key])


@pytest.mark.parametrize("depth,width", [(0, 32), (12, 32), (13, 8), (16, 8), (0, 8), (12, 16)])
def test_se_edge_emu(depth, width):
    hdr, subs = goldenio.genome()
    e = emu.EmuEngine(hdr.lengths)
    for w, sfx in enumerate(refio.SUFFIXES):
        assert e.load(w, subs[sfx].seq, subs[sfx].index, force_depth=depth) == 0
    z = goldenio.load("se_edge.npz")
    for ag, pre in ((False, "ct_best_"), (True, "ga_best_")):
        for key in [k for k in z.files if k.startswith(pre)]:
            m, b = (int(x[1:]) for x in key[len(pre):].split("_"))
            rc, out, _ = e.map_se(z["buf"], z["offs"], refio.BEST_DT, ag=ag, m=m, b=b, width=width)
            assert rc == 0
            _cmp_best(out, z[key])
    e.close()


@pytest.mark.parametrize("logged", [False, True], ids=["heap-in-kernel", "logged+replay"])
@pytest.mark.parametrize("width", [32, 8])
def test_pe_emu_matches_reference(engine, width, logged):
    hdr, _ = goldenio.genome()
    z = goldenio.load("pe.npz")
    L = refio.oracle_lib()
    starts = np.ascontiguousarray(hdr.start_index, np.uint32)
    lengths = np.ascontiguousarray(hdr.lengths, np.uint32)
    chroms = refio.WoChroms(len(lengths), starts.ctypes.data, lengths.ctypes.data)
    for m, k in ((6, 50), (8, 3), (4, 2)):
        got = {}
        for mate, ag in ((1, False), (2, True)):
            buf, offs = refio.pack_reads(z[f"m{mate}"])
            rc, ranked, sizes = engine.map_pe_mate(buf, offs, refio.CAND_DT, ag, m=m, top_k=k, width=width, logged=logged)
            assert rc == 0
            assert np.array_equal(sizes, z[f"sizes{mate}_m{m}_k{k}"])
            want = z[f"ranked{mate}_m{m}_k{k}"]
            for f in ("genome_pos", "mismatch", "strand"):
                assert np.array_equal(ranked[f], want[f]), (m, k, mate, f)
            got[mate] = (ranked, sizes, offs)
        # pairing vs the oracle's MergePairedEndResults loop
        pr = engine.pair(got[1][0], got[1][1], got[1][2], got[2][0], got[2][1], got[2][2], k, m, 500)
        for j in range(len(pr)):
            bi, bj = C.c_int32(-1), C.c_int32(-1)
            r1 = np.ascontiguousarray(got[1][0][j]); r2 = np.ascontiguousarray(got[2][0][j])
            t = L.wo_pe_pair(C.byref(chroms), r1.ctypes.data_as(C.c_void_p), C.c_uint32(int(got[1][1][j])),
                             C.c_uint32(100), r2.ctypes.data_as(C.c_void_p), C.c_uint32(int(got[2][1][j])),
                             C.c_uint32(100), C.c_uint32(m), C.c_int(500), C.byref(bi), C.byref(bj))
            assert (t, bi.value, bj.value) == (pr[j]["best_times"], pr[j]["best_i"], pr[j]["best_j"]), j
