"""The batch loop of the `walt` program (walt_b200/host/walt_main.cpp) without a GPU: a batch of -N reads
is loaded, mapped and written in parts, three parts under way at a time, while the engines start beside the
first load.  None of that may show in the outputs: the N replacement follows ONE rand() stream per batch
(srand(0) at its start, mapping.cpp:73), lines stay in input order, the counters add up.  The engine is
replaced by tests/stub/fake_engine.cpp (results = a hash of each read's 2-bit bytes) through
LD_LIBRARY_PATH, so the outputs depend on every base the loader hands over and on nothing else."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from walt_b200 import host
from test_host_parallel import load_all, make_fastq

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WALT = os.path.join(ROOT, "walt_b200", "bin", "walt")
GENOME_LEN = 3_000_000


@pytest.fixture(scope="module")
def fake(tmp_path_factory):
    d = tmp_path_factory.mktemp("fake")
    so = str(d / "libwaltb200.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-o", so, os.path.join(ROOT, "tests", "stub", "fake_engine.cpp")],
                   check=True)
    idx = str(d / "g.dbindex")
    chroms = host.Chroms(names=["chrOnly"], lengths=[GENOME_LEN])
    host.write_dbindex_header(idx, chroms, 1000)
    for sfx in ("_CT00", "_CT01", "_GA10", "_GA11"):
        open(idx + sfx, "wb").close()
    return {"dir": str(d), "index": idx}


def _walt(fake, args, part, threads=0, start_ms=0):
    env = dict(os.environ, LD_LIBRARY_PATH=fake["dir"], WALT_FAKE_GENOME_LEN=str(GENOME_LEN), WALT_PART_READS=str(part))
    if start_ms:
        env["WALT_FAKE_START_MS"] = str(start_ms)
    cmd = [WALT, "-i", fake["index"]] + args + (["-t", str(threads)] if threads else [])
    return subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)


def _outputs(d, stem):
    return {f: open(os.path.join(d, f), "rb").read() for f in sorted(os.listdir(d)) if f.startswith(stem)}


def test_parts_of_a_batch_are_the_batch(tmp_path):
    rng = np.random.default_rng(5)
    data = make_fastq(rng, 1200)
    path = str(tmp_path / "r.fastq")
    open(path, "wb").write(data)
    for adaptor in ("", "AGATCGGAAGAGC"):
        for N in (10 ** 6, 500):
            want = [r for b in load_all(path, N, adaptor) for r in b]
            for P in (1, 7, 64, 499, 500):
                fq, b, got, left = host.Fastq(path), host.Batch(), [], N
                while True:
                    take = min(P, left)
                    n = fq.next_part(b, take, adaptor, restart_rand=left == N)
                    seqs, offs = b.arrays()
                    for i in range(n):
                        got.append((b.L.walt_batch_name(b.h, C.c_uint32(i)), seqs[int(offs[i]):int(offs[i + 1])].tobytes(),
                                    b.L.walt_batch_qual(b.h, C.c_uint32(i))))
                    left -= n
                    if left == 0:
                        left = N
                    if n < take:
                        break
                fq.close(); b.free()
                assert got == want, (adaptor, N, P)


def test_loaders_two_bit_form_is_walt_pack_reads_of_its_reads(tmp_path):
    """The loader packs every read while it copies it and patches the characters the rand() stream replaces:
    the result must be what walt_pack_reads makes of the final ASCII reads, zero bytes between the reads included."""
    rng = np.random.default_rng(41)
    path = str(tmp_path / "r.fastq")
    open(path, "wb").write(make_fastq(rng, 2500))
    for adaptor in ("", "AGATCGGAAGAGC"):
        for threads, chunk in ((1, 0), (8, 256), (3, 4096)):
            host.set_threads(threads)
            host.set_grain(chunk, 0)
            fq, b = host.Fastq(path), host.Batch()
            for part, restart in ((700, True), (1, False), (900, False), (10 ** 6, True)):
                n = fq.next_part(b, part, adaptor, restart_rand=restart)
                assert n > 0
                seqs, offs = b.arrays()
                assert np.isin(seqs[:int(offs[n])], np.frombuffer(b"ACGT", np.uint8)).all()
                want = host.pack_reads_2bit(seqs, offs)
                assert np.array_equal(b.packed(), want), (adaptor, threads, part)
            fq.close(); b.free()
    host.set_threads(0); host.set_grain(0, 0)


@pytest.mark.parametrize("mode", ["mr", "sam", "ag"])
def test_single_end_outputs_do_not_depend_on_the_parts(fake, tmp_path, mode):
    rng = np.random.default_rng(17)
    fq = str(tmp_path / "r.fastq")
    open(fq, "wb").write(make_fastq(rng, 4000))
    flags = {"mr": ["-a", "-u"], "sam": ["-sam", "-a", "-u"], "ag": ["-A", "-a", "-u", "-C", "AGATCGGAAGAGC"]}[mode]
    outs = []
    # -N 1500: three batches (the rand() stream restarts twice); parts that divide a batch, that do not, one read each time
    for k, (part, threads, start_ms) in enumerate([(1 << 20, 1, 0), (1500, 0, 0), (400, 3, 150), (37, 8, 0), (1, 2, 0)]):
        o = str(tmp_path / f"o{k}")
        r = _walt(fake, ["-r", fq, "-o", o, "-N", "1500"] + flags, part, threads, start_ms)
        assert r.returncode == 0, r.stderr
        outs.append({f[len(f"o{k}"):]: v for f, v in _outputs(str(tmp_path), f"o{k}").items()})
    assert len(outs[0]) == (2 if mode == "sam" else 4) and len(outs[0][""]) > 100000
    n_records = sum(len(b) for b in load_all(fq, 1500, ""))      # (the quirks of make_fastq shift a few record frames)
    assert b"total_reads: %d\n" % n_records in outs[0][".mapstats"] and n_records > 3900
    for o in outs[1:]:
        assert o == outs[0]


@pytest.mark.parametrize("mode", ["mr", "sam", "pbat"])
def test_paired_end_outputs_do_not_depend_on_the_parts(fake, tmp_path, mode):
    rng = np.random.default_rng(23)
    f1, f2 = str(tmp_path / "r_1.fastq"), str(tmp_path / "r_2.fastq")
    open(f1, "wb").write(make_fastq(rng, 3000, quirks=False))
    open(f2, "wb").write(make_fastq(rng, 3000, quirks=False))
    flags = {"mr": ["-a", "-u"], "sam": ["-sam", "-a", "-u"], "pbat": ["-P", "-a", "-u", "-C", "AGATCGGAAGAGC:TTTTTTTTTT"]}[mode]
    outs = []
    for k, (part, threads) in enumerate([(1 << 20, 1), (1000, 0), (333, 4), (1, 3)]):
        o = str(tmp_path / f"o{k}")
        r = _walt(fake, ["-1", f1, "-2", f2, "-o", o, "-N", "1000"] + flags, part, threads)
        assert r.returncode == 0, r.stderr
        outs.append({f[len(f"o{k}"):]: v for f, v in _outputs(str(tmp_path), f"o{k}").items()})
    assert b"total_read_pairs: 3000" in outs[0][".mapstats"] and len(outs[0][""]) > 100000
    for o in outs[1:]:
        assert o == outs[0]


def test_unequal_mate_files_end_the_run_like_the_reference(fake, tmp_path):
    """paired.cpp:650-651, 673-677: a batch whose two files hold different numbers of reads ends the program
    with an error and without mapstats -- unless the first file ends exactly at a batch boundary, where
    the second one is not looked at."""
    rng = np.random.default_rng(29)
    recs = [make_fastq(rng, 1, quirks=False) for _ in range(260)]
    def files(n1, n2):
        f1, f2 = str(tmp_path / f"a{n1}_{n2}_1.fastq"), str(tmp_path / f"a{n1}_{n2}_2.fastq")
        open(f1, "wb").write(b"".join(recs[:n1])); open(f2, "wb").write(b"".join(recs[:n2]))
        return f1, f2
    for n1, n2, N, fails in [(200, 205, 100, False),     # file 1 ends at a batch boundary: file 2's extra reads are never seen
                             (200, 205, 150, True),      # ... inside a batch: counts 50 != 55
                             (205, 200, 100, True), (130, 130, 100, False),
                             (100, 130, 1000, True)]:    # inside the batch, at a part boundary (parts of 50)
        f1, f2 = files(n1, n2)
        for part in (1 << 20, 50):
            o = str(tmp_path / f"u{n1}_{n2}_{N}_{part}")
            r = _walt(fake, ["-1", f1, "-2", f2, "-o", o, "-N", str(N)], part)
            assert (r.returncode != 0) == fails, (n1, n2, N, part, r.stderr)
            assert (b"should be the same" in r.stderr) == fails
            assert (os.path.getsize(o + ".mapstats") == 0) == fails


def test_errors_surface_whatever_stage_sees_them(fake, tmp_path):
    fq = str(tmp_path / "r.fastq")
    open(fq, "wb").write(make_fastq(np.random.default_rng(3), 300))
    r = _walt(fake, ["-r", str(tmp_path / "missing.fastq"), "-o", str(tmp_path / "o")], 64, start_ms=200)
    assert r.returncode == 1 and b"cannot open input file" in r.stderr
    r = _walt(fake, ["-r", fq, "-o", str(tmp_path / "no_such_dir" / "o")], 64)
    assert r.returncode == 1
    empty = str(tmp_path / "e.fastq")
    open(empty, "wb").close()
    r = _walt(fake, ["-r", empty, "-o", str(tmp_path / "oe")], 64)
    assert r.returncode == 0 and b"total_reads: 0" in open(str(tmp_path / "oe.mapstats"), "rb").read()
