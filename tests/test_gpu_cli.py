"""The `walt` and `makedb` programs end to end on the GPU against the reference's outputs:
every golden command line (tests/golden/cli/cases.json) must produce byte-identical SAM / MR /
side files / mapstats; -P is checked against its derived oracle (mates swapped); makedb's
files are compared with the reference makedb's on an N-free genome."""
import json
import os
import subprocess

import numpy as np
import pytest

import goldenio
import refio
import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WALT = os.path.join(ROOT, "walt_b200", "bin", "walt")
MAKEDB = os.path.join(ROOT, "walt_b200", "bin", "makedb")
CLI = os.path.join(goldenio.GOLDEN, "cli")
CASES = json.load(open(os.path.join(CLI, "cases.json")))


@pytest.fixture(scope="module")
def dbindex(tmp_path_factory):
    d = tmp_path_factory.mktemp("idx")
    p = str(d / "g.dbindex")
    goldenio.write_dbindex(p)
    return p


def _run(args, cwd=None, env=None):
    return subprocess.run([WALT] + args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600,
                          env=dict(os.environ, **env) if env else None)


def _compare(got_dir, want_dir, files):
    assert sorted(os.listdir(got_dir)) == files
    for f in files:
        got = open(os.path.join(got_dir, f), "rb").read()
        want = open(os.path.join(want_dir, f), "rb").read()
        if got != want:
            gl, wl = got.split(b"\n"), want.split(b"\n")
            for i, (a, b) in enumerate(zip(gl, wl)):
                assert a == b, (f, i, a[:300], b[:300])
            assert len(gl) == len(wl), (f, len(gl), len(wl))


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_walt_cli_matches_reference(case, dbindex, tmp_path):
    args = [a if not a.endswith(".fastq") else os.path.join(CLI, a) for a in case["args"]]
    r = _run(["-i", dbindex, "-o", str(tmp_path / "out"), "-t", "4"] + args)
    assert r.returncode == case["returncode"], r.stderr.decode()[-2000:]
    _compare(str(tmp_path), os.path.join(CLI, case["name"]), case["files"])


@pytest.mark.parametrize("name", ["se_smallN", "se_sam", "pe_smallN", "pe_sam", "pe_clip_k3"])
def test_walt_cli_sharded_over_engines(dbindex, tmp_path, name):
    """-gpus 3: every batch is cut into three contiguous ranges, each mapped by its own engine
    (walt_main.cpp: packed + lo, offs + lo) and written back into its slice of the batch's result
    array.  The program clamps -gpus to the visible devices; WALT_SHARE_DEVICES=1 (a test hook for small
    indexes) lets the shards share devices instead, so the split is exercised on a one-GPU box as well;
    the bytes must not depend on it."""
    case = next(c for c in CASES if c["name"] == name)
    args = [a if not a.endswith(".fastq") else os.path.join(CLI, a) for a in case["args"]]
    r = _run(["-i", dbindex, "-o", str(tmp_path / "out"), "-gpus", "3"] + args, env={"WALT_SHARE_DEVICES": "1"})
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    _compare(str(tmp_path), os.path.join(CLI, case["name"]), case["files"])


def test_walt_cli_gpus_clamped_to_devices(dbindex, tmp_path):
    """-gpus beyond the visible devices is clamped (no second index replica on one device), same bytes"""
    case = next(c for c in CASES if c["name"] == "se_sam")
    args = [a if not a.endswith(".fastq") else os.path.join(CLI, a) for a in case["args"]]
    r = _run(["-i", dbindex, "-o", str(tmp_path / "out"), "-gpus", "64"] + args)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    assert b"device(s) visible, using" in r.stderr
    _compare(str(tmp_path), os.path.join(CLI, case["name"]), case["files"])


@pytest.mark.parametrize("name", ["se_sam", "pe_sam", "pe_clip_k3"])
def test_walt_cli_two_physical_gpus(dbindex, tmp_path, name):
    """-gpus 2 on two physical devices (skipped on a one-GPU box): the second engine gets its index over
    NVLink from the first (walt_engine_clone_index), the bytes are the reference's."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible devices")
    case = next(c for c in CASES if c["name"] == name)
    args = [a if not a.endswith(".fastq") else os.path.join(CLI, a) for a in case["args"]]
    r = _run(["-i", dbindex, "-o", str(tmp_path / "out"), "-gpus", "2"] + args, env={"WALT_TIMING": "1"})
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    assert b"cloned over" in r.stderr
    _compare(str(tmp_path), os.path.join(CLI, case["name"]), case["files"])


def _swap_mate_lines(sam_bytes):
    """reference output of (-1 B -2 A) -> what (-1 A -2 B -P) must print: within every pair of
    lines swap the order and the first/last flag bits."""
    lines = sam_bytes.decode().split("\n")
    head = [l for l in lines if l.startswith("@")]
    body = [l for l in lines if l and not l.startswith("@")]
    assert len(body) % 2 == 0
    out = []
    for i in range(0, len(body), 2):
        pair = []
        for l in (body[i + 1], body[i]):
            f = l.split("\t")
            flag = int(f[1])
            flag = (flag & ~0xC0) | (0x80 if flag & 0x40 else 0) | (0x40 if flag & 0x80 else 0)
            f[1] = str(flag)
            pair.append("\t".join(f))
        out += pair
    return ("\n".join(head + out) + "\n").encode()


@pytest.mark.reference
@pytest.mark.parametrize("clip", [[], ["-C", "AGATCGGAAGAGC:TCGGAAGAGCACA", "-m", "8"]], ids=["plain", "clip_m8"])
def test_pbat_equals_reference_with_mates_swapped(dbindex, tmp_path, clip):
    """-P against its derived oracle; with -C T_adaptor:A_adaptor the T-rich adaptor
    belongs to the C->T mate, which is the SECOND file under PBAT (configs[4]: -P -C ... -m 8)."""
    if not refio.have_reference():
        pytest.skip("oracle/_ref not built")
    f1, f2 = os.path.join(CLI, "pe_reads_1.fastq"), os.path.join(CLI, "pe_reads_2.fastq")
    ref_out = str(tmp_path / "ref.sam")
    refio.ref_walt(["-i", dbindex, "-1", f2, "-2", f1, "-o", ref_out, "-sam", "-u", "-a", "-k", "10", "-L", "500"] + clip)
    r = _run(["-i", dbindex, "-1", f1, "-2", f2, "-P", "-o", str(tmp_path / "our.sam"), "-sam", "-u", "-a", "-k", "10",
              "-L", "500"] + clip)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    want = _swap_mate_lines(open(ref_out, "rb").read())
    got = open(str(tmp_path / "our.sam"), "rb").read()
    assert got == want
    # mapstats: mate1/mate2 blocks swap
    ws = open(ref_out + ".mapstats").read()
    gs = open(str(tmp_path / "our.sam.mapstats")).read()
    i1, i2, i3 = ws.index("mate1:"), ws.index("mate2:"), ws.index("frag_len_distribution:")
    swapped = ws[:i1] + "mate1:" + ws[i2 + 6:i3] + "mate2:" + ws[i1 + 6:i2] + ws[i3:]
    assert gs == swapped
    # single-end -P == -A
    se = os.path.join(CLI, "ga_reads.fastq")
    a = _run(["-i", dbindex, "-r", se, "-A", "-o", str(tmp_path / "a.mr"), "-u", "-a"])
    p = _run(["-i", dbindex, "-r", se, "-P", "-o", str(tmp_path / "p.mr"), "-u", "-a"])
    assert a.returncode == 0 and p.returncode == 0
    for sfx in ("", "_ambiguous", "_unmapped", ".mapstats"):
        assert open(str(tmp_path / "a.mr") + sfx, "rb").read() == open(str(tmp_path / "p.mr") + sfx, "rb").read()


@pytest.mark.reference
@pytest.mark.parametrize("one_output", [False, True])
def test_file_lists_match_reference(dbindex, tmp_path, one_output):
    """Comma-separated read-file lists, single-end and paired-end inputs in one run, one output per
    input or one shared output (walt.cpp:205-233, 255-273): every file both programs write must be
    identical.  The reference is run here, next to ours."""
    if not refio.have_reference():
        pytest.skip("oracle/_ref not built")
    se = ",".join(os.path.join(CLI, f) for f in ("se_reads.fastq", "crlf_reads.fastq"))
    opts = ["-i", dbindex, "-r", se, "-1", os.path.join(CLI, "pe_reads_1.fastq"), "-2", os.path.join(CLI, "pe_reads_2.fastq"),
            "-sam", "-u", "-N", "700"]
    dirs = {}
    for who in ("ref", "ours"):
        d = tmp_path / who
        d.mkdir()
        outs = str(d / "all.sam") if one_output else ",".join(str(d / f"o{i}.sam") for i in range(3))
        if who == "ref":
            refio.ref_walt(opts + ["-o", outs])
        else:
            r = _run(opts + ["-o", outs])
            assert r.returncode == 0, r.stderr.decode()[-2000:]
        dirs[who] = d
    names = sorted(os.listdir(dirs["ref"]))
    assert names == sorted(os.listdir(dirs["ours"])) and len(names) >= (2 if one_output else 6)
    for f in names:
        assert open(dirs["ref"] / f, "rb").read() == open(dirs["ours"] / f, "rb").read(), f


def test_cli_errors(dbindex, tmp_path):
    f = os.path.join(CLI, "se_reads.fastq")
    out = str(tmp_path / "o")
    assert _run(["-i", dbindex + ".nope", "-r", f, "-o", out]).returncode == 1
    assert _run(["-i", dbindex, "-r", f + ".txt", "-o", out]).returncode == 1
    assert _run(["-i", dbindex, "-r", f, "-o", out, "-k", "1"]).returncode == 1
    assert _run(["-i", dbindex, "-r", f, "-o", out, "-N", "100000001"]).returncode == 1
    assert _run(["-i", dbindex, "-r", f]).returncode == 0          # missing -o: message, exit 0
    assert _run(["-i", dbindex, "-r", f, "-o", out, "stray"]).returncode == 0   # leftover: help, exit 0
    # unequal mate files (paired.cpp:673-677)
    f1 = os.path.join(CLI, "pe_reads_1.fastq")
    short = str(tmp_path / "short_2.fastq")
    lines = open(os.path.join(CLI, "pe_reads_2.fastq")).read().split("\n")
    open(short, "w").write("\n".join(lines[:400]) + "\n")
    r = _run(["-i", dbindex, "-1", f1, "-2", short, "-o", out])
    assert r.returncode == 1 and b"should be the same" in r.stderr


@pytest.mark.reference
@pytest.mark.parametrize("divergence", [0.02, 0.0])
def test_makedb_matches_reference_makedb(tmp_path, divergence):
    """All five files of the index byte-identical to the reference makedb's, including the order
    of tied suffixes inside a bucket (exact repeats when divergence is 0, chromosome ends that
    look alike): the device builder replays libstdc++'s std::sort (walt_stdsort.cuh)."""
    if not refio.have_reference():
        pytest.skip("oracle/_ref not built")
    chroms = synth.make_repeat_genome([120000, 70000, 500, 30], seed=41, n_families=5, fam_len=(200, 900),
                                      copies=(3, 20), divergence=divergence, repeat_frac=0.2)
    chroms[1][1][-300:] = chroms[0][1][-300:]      # suffixes cut by a chromosome end tie as well
    chroms[2][1][:] = chroms[0][1][-500:]
    if divergence == 0.0:                          # one large class: 400 copies of a 250-mer
        unit = chroms[0][1][5000:5250].copy()
        for k in range(400):
            chroms[0][1][10000 + 260 * k: 10000 + 260 * k + 250] = unit
    fa = str(tmp_path / "g.fa")
    synth.write_fasta(fa, chroms)
    ours, ref = str(tmp_path / "ours.dbindex"), str(tmp_path / "ref.dbindex")
    r = subprocess.run([MAKEDB, "-c", fa, "-o", ours], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    refio.ref_makedb(fa, ref)
    assert open(ours, "rb").read() == open(ref, "rb").read()
    hdr = refio.read_header(ref)
    L = refio.oracle_lib()
    import ctypes as C
    starts = np.ascontiguousarray(hdr.start_index, np.uint32)
    n_ties = 0
    for sfx in refio.SUFFIXES:
        b = refio.read_subindex(ref + sfx, hdr.genome_len)
        for i in range(0, b.index.size - 1, 5):
            n_ties += L.wo_bucket_cmp(b.seq.ctypes.data_as(C.c_void_p), C.c_uint32(len(hdr.lengths)),
                                      starts.ctypes.data_as(C.c_void_p), C.c_uint32(int(b.index[i])),
                                      C.c_uint32(int(b.index[i + 1]))) == 0
        assert open(ours + sfx, "rb").read() == open(ref + sfx, "rb").read(), sfx
    assert n_ties > 0, "no tied suffixes in this genome: the tie order is untested"
