"""The multi-threaded text path of libwalthost (parallel FASTQ scan with ordered N replacement,
block-parallel SAM/MR formatting with ordered commit) must produce the same bytes as (a) itself
on one thread and (b) a line-by-line Python restatement of the reference loader
(LoadReadsFromFastqFile, mapping.cpp:65-121, with glibc srand/rand through ctypes), whatever the
task grain, thread count and batch size -- including the fgets quirks: 999-byte pieces, the last
character of every piece dropped, empty lines skipped, embedded NULs, CRLF, a missing final
newline, a trailing partial record."""
import ctypes as C
import os

import numpy as np
import pytest

from walt_b200 import host

LIBC = C.CDLL("libc.so.6")


def reference_loader(data: bytes, max_reads: int, adaptor: bytes, clip):
    """-> list of batches, each a list of (name, seq, qual) as the reference would load them."""
    pos, batches = 0, []
    while True:
        LIBC.srand(0)
        recs, line_count, code, lim = [], 0, 0, max_reads * 4
        name = seq = b""
        while line_count < lim and pos < len(data):
            room = min(len(data) - pos, 999)
            nl = data.find(b"\n", pos, pos + room)
            q = nl + 1 if nl >= 0 else pos + room
            piece = data[pos:q]
            pos = q
            z = piece.find(b"\0")
            if z >= 0:
                piece = piece[:z]
            piece = piece[:-1]
            if not piece:
                continue
            if code == 0:
                sp = piece.find(b" ")
                name = piece[1:] if sp <= 0 else piece[1:sp]
            elif code == 1:
                s = bytearray(piece)
                if adaptor:
                    clip(adaptor, s)
                for i, c in enumerate(s):
                    if c not in b"ACGT":
                        s[i] = b"ACGT"[LIBC.rand() % 4]
                seq = bytes(s)
            elif code == 3:
                recs.append((name, seq, piece))
            line_count += 1
            code = (code + 1) % 4
        if not recs:
            break
        batches.append(recs)
        if len(recs) < max_reads:
            break
    return batches


def _clip(adaptor, s):
    L = host.load_library()
    buf = C.create_string_buffer(bytes(s), len(s) + 1)
    L.walt_clip_adaptor(adaptor, buf, C.c_size_t(len(s)))
    s[:] = buf.raw[:len(s)]


def make_fastq(rng, n, quirks=True):
    out = []
    for i in range(n):
        ln = int(rng.integers(14, 160))
        seq = bytes(rng.choice(list(b"ACGT"), size=ln).tolist())
        r = rng.random()
        if quirks:
            if r < 0.15:
                seq = seq[:ln // 2] + b"N" * int(rng.integers(1, 5)) + seq[ln // 2:]
            elif r < 0.2:
                seq = seq.lower()
            elif r < 0.3:
                cut = int(rng.integers(0, ln))
                seq = (seq[:cut] + b"AGATCGGAAGAGCACACGTC" + seq)[:ln]
            elif r < 0.31:
                seq = seq * 9          # > 999 characters: fgets splits the line
        qual = bytes(rng.integers(33, 74, size=len(seq)).astype(np.uint8).tolist())
        name = b"@r%d" % i
        if quirks and rng.random() < 0.3:
            name += b" 1:N:0:%d" % i
        if quirks and rng.random() < 0.01:
            name = b"@ lead%d x" % i
        eol = b"\r\n" if quirks and rng.random() < 0.05 else b"\n"
        rec = name + eol + seq + eol + b"+" + eol + qual + eol
        if quirks and rng.random() < 0.03:
            rec = b"\n" + rec
        if quirks and rng.random() < 0.01:
            rec = rec.replace(b"+" + eol, b"+\0junk" + eol)
        out.append(rec)
    return b"".join(out)


def load_all(path, max_reads, adaptor):
    fq, b, batches = host.Fastq(path), host.Batch(), []
    while True:
        n = fq.next_batch(b, max_reads, adaptor)
        if n == 0:
            break
        seqs, offs = b.arrays()
        L = b.L
        recs = []
        for i in range(n):
            recs.append((L.walt_batch_name(b.h, C.c_uint32(i)), seqs[int(offs[i]):int(offs[i + 1])].tobytes(),
                         L.walt_batch_qual(b.h, C.c_uint32(i))))
        batches.append(recs)
        if n < max_reads:
            break
    fq.close(); b.free()
    return batches


@pytest.fixture(autouse=True)
def _restore():
    yield
    host.set_threads(0)
    host.set_grain(0, 0)


@pytest.mark.parametrize("tail", ["newline", "no_newline", "partial_record", "empty_lines"])
@pytest.mark.parametrize("adaptor", ["", "AGATCGGAAGAGC"])
def test_parallel_loader_matches_reference_loader(tmp_path, tail, adaptor):
    rng = np.random.default_rng(hash((tail, adaptor)) % 2 ** 32)
    data = make_fastq(rng, 1500)
    if tail == "no_newline":
        data = data[:-1]
    elif tail == "partial_record":
        data += b"@last\nACGTNACGT\n+\n"
    elif tail == "empty_lines":
        data += b"\n\n\r\n"
    path = str(tmp_path / "r.fastq")
    open(path, "wb").write(data)
    for max_reads in (10 ** 6, 301, 7):
        want = reference_loader(data, max_reads, adaptor.encode(), _clip)
        for threads, chunk in ((1, 0), (8, 512), (3, 4096)):
            host.set_threads(threads)
            host.set_grain(chunk, 0)
            got = load_all(path, max_reads, adaptor)
            assert len(got) == len(want), (max_reads, threads)
            for gb, wb in zip(got, want):
                assert gb == wb, (max_reads, threads, chunk)


def test_loader_corner_files(tmp_path):
    for i, data in enumerate([b"", b"\n", b"@a\nACGT\n+\nIIII", b"@a\nACGT\n+\n", b"@a b\n" + b"ACGT" * 20 + b"\n+\n" + b"I" * 80 + b"\n"]):
        path = str(tmp_path / f"c{i}.fastq")
        open(path, "wb").write(data)
        for threads in (1, 4):
            host.set_threads(threads)
            host.set_grain(16, 0)
            assert load_all(path, 5, "") == reference_loader(data, 5, b"", _clip)


def _random_results(rng, n, genome_len):
    from walt_b200.engine import BEST_DT
    res = np.zeros(n, BEST_DT)
    res["times"] = rng.choice([0, 1, 1, 1, 2, 5], size=n)
    res["genome_pos"] = rng.integers(0, genome_len - 200, size=n)
    res["mismatch"] = rng.integers(0, 7, size=n)
    res["strand"] = rng.choice([b"+", b"-"], size=n)
    return res


@pytest.mark.parametrize("sam", [False, True])
@pytest.mark.parametrize("ag", [False, True])
def test_parallel_se_writer_is_order_preserving(tmp_path, sam, ag):
    rng = np.random.default_rng(11)
    data = make_fastq(rng, 5000, quirks=False)
    path = str(tmp_path / "r.fastq")
    open(path, "wb").write(data)
    lengths = [400000, 250000, 5000]
    chroms = host.Chroms(names=["chrA", "chrB", "c3"], lengths=lengths)
    res = _random_results(rng, 5000, sum(lengths))
    outs = []
    for threads, block in ((1, 0), (8, 64), (5, 333)):
        host.set_threads(threads)
        host.set_grain(0, block)
        out = str(tmp_path / f"o_{threads}")
        open(out, "w").close(); open(out + ".mapstats", "w").close()
        fq, b = host.Fastq(path), host.Batch()
        assert fq.next_batch(b, 10 ** 6) == 5000
        w = host.SeWriter(out, chroms, ag=ag, ambiguous=True, unmapped=True, sam=sam)
        w.write(b, res, n_short=4)
        w.close(); fq.close(); b.free()
        files = sorted(f for f in os.listdir(tmp_path) if f.startswith(f"o_{threads}"))
        outs.append([open(str(tmp_path / f), "rb").read() for f in files])
        assert len(files) == (2 if sam else 4)
    assert outs[0] == outs[1] == outs[2]
    assert len(outs[0][0]) > 100000


@pytest.mark.parametrize("sam", [False, True])
@pytest.mark.parametrize("pbat", [False, True])
def test_parallel_pe_writer_is_order_preserving(tmp_path, sam, pbat):
    from walt_b200.engine import PE_RESULT_DT
    rng = np.random.default_rng(12)
    n = 4000
    p1, p2 = str(tmp_path / "r_1.fastq"), str(tmp_path / "r_2.fastq")
    open(p1, "wb").write(make_fastq(rng, n, quirks=False))
    open(p2, "wb").write(make_fastq(rng, n, quirks=False))
    lengths = [400000, 250000, 5000]
    chroms = host.Chroms(names=["chrA", "chrB", "c3"], lengths=lengths)
    res = np.zeros(n, PE_RESULT_DT)
    res["pair"]["best_times"] = rng.choice([0, 1, 1, 2], size=n)
    for c in ("c1", "c2"):
        res[c]["genome_pos"] = rng.integers(1000, 390000, size=n)
        res[c]["mismatch"] = rng.integers(0, 4, size=n)
    res["c1"]["strand"] = rng.choice([b"+", b"-"], size=n)
    res["c2"]["strand"] = np.where(res["c1"]["strand"] == b"+", b"-", b"+")
    res["c2"]["genome_pos"] = 400000 - res["c1"]["genome_pos"] - rng.integers(100, 400, size=n)
    for sname in ("single1", "single2"):
        r = _random_results(rng, n, sum(lengths))
        for f in ("genome_pos", "times", "mismatch", "strand"):
            res[sname][f] = r[f]
    outs = []
    for threads, block in ((1, 0), (8, 50)):
        host.set_threads(threads)
        host.set_grain(0, block)
        out = str(tmp_path / f"o_{threads}")
        open(out, "w").close(); open(out + ".mapstats", "w").close()
        f1, f2, b1, b2 = host.Fastq(p1), host.Fastq(p2), host.Batch(), host.Batch()
        assert f1.next_batch(b1, 10 ** 6) == n and f2.next_batch(b2, 10 ** 6) == n
        w = host.PeWriter(out, chroms, ambiguous=True, unmapped=True, sam=sam, pbat=pbat)
        assert w.L.walt_pe_writer_write_compact(w.h, b1.h, b2.h, res.ctypes.data_as(C.c_void_p), C.c_uint32(n)) == 0
        w.close(); f1.close(); f2.close(); b1.free(); b2.free()
        files = sorted(f for f in os.listdir(tmp_path) if f.startswith(f"o_{threads}"))
        outs.append([open(str(tmp_path / f), "rb").read() for f in files])
    assert outs[0] == outs[1]
    assert len(outs[0][0]) > 100000
