"""Generate the committed golden vectors from the UNMODIFIED reference built under
oracle/_ref/ (makedb, walt, libwaltref.so).  Run in the build container only:

    python tests/golden/make_golden.py

Outputs (all small, committed):
  genome.npz     chromosome names/lengths + the four sub-indexes as makedb wrote them
                 (converted ASCII sequence, index[]; counter[] is re-derived from index[] by
                 hashing, tests/goldenio.py) + the FASTA text
  se_ct.npz / se_ga.npz / se_edge.npz
                 ACGT-only reads + BestMatch arrays from the reference's SingleEndMapping
                 (through oracle/ref_shim.cpp) for several (-m, -b)
  pe.npz         mate reads + drained TopCandidates of both mates from the reference's
                 PairEndMapping for (-m 6 -k 50) and (-m 8 -k 3)
  cli/           FASTQ inputs (with N, lowercase, CRLF, short reads, adaptors) and the
                 reference `walt` outputs (SAM, MR, side files, mapstats) for a list of
                 command lines (cli/cases.json)

TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refio  # noqa: E402
import synth  # noqa: E402

CLI = os.path.join(HERE, "cli")


def build_genome():
    # uniform background with diverged + exact repeats so ambiguous hits and -b matter;
    # a 150 bp chromosome (every entry tainted) and a 20 bp one (never indexed)
    chroms = synth.make_repeat_genome([60000, 35000, 150, 20], seed=11, n_families=6,
                                      fam_len=(200, 600), copies=(3, 12), divergence=0.01,
                                      repeat_frac=0.25)
    rng = np.random.default_rng(5)
    s0 = chroms[0][1]
    # exact duplicates (forward and reverse-complement) -> ambiguous reads
    s0[30000:30400] = s0[1000:1400]
    chroms[1][1][5000:5300] = synth.revcomp(s0[2000:2300])
    # a poly-T/low-complexity stretch and some N's (makedb replaces them with rand())
    s0[45000:45120] = ord("T")
    chroms[1][1][20000:20010] = ord("N")
    # chromosome ends that look like each other, to exercise the past-the-end probes
    chroms[1][1][-200:] = s0[-200:]
    chroms[2][1][:] = s0[-150:]
    _ = rng
    return chroms


def edge_reads(chroms, rng):
    """ACGT-only reads of awkward lengths / positions."""
    cat = np.concatenate([s for _, s in chroms])
    cat = np.where(cat == ord("N"), ord("A"), cat)
    starts = np.concatenate([[0], np.cumsum([len(s) for _, s in chroms])])
    out = []
    for rl in (37, 38, 39, 40, 70, 71, 72, 100, 101, 142, 143, 149, 150, 151, 152, 153, 200, 250):
        for _ in range(12):
            p = int(rng.integers(0, starts[2] - rl - 1))
            r = cat[p:p + rl].copy()
            if rng.random() < 0.5:
                r = synth.revcomp(r)
            k = int(rng.integers(0, 8))
            for _ in range(k):
                r[int(rng.integers(0, rl))] = synth._ACGT[int(rng.integers(0, 4))]
            out.append(r.tobytes())
        # windows glued to chromosome starts / ends (bounds rules, taint)
        for c in (0, 1):
            cs, ce = int(starts[c]), int(starts[c + 1])
            for p in (cs, cs + 1, cs + 2, ce - rl - 2, ce - rl - 1, ce - rl):
                if p < cs or p + rl > ce:
                    continue
                out.append(cat[p:p + rl].tobytes())
                out.append(synth.revcomp(cat[p:p + rl]).tobytes())
    # reads spanning into the 150 bp chromosome and the tail
    for rl in (38, 60, 100, 149):
        cs = int(starts[2])
        for d in (-40, -10, 0, 1, 2):
            p = cs + d
            if p + rl <= len(cat):
                out.append(cat[p:p + rl].tobytes())
    out.append(b"T" * 100)
    out.append(b"A" * 100)
    out.append(b"ACGT" * 25)
    out.append(b"")
    return out


def se_cases(index, reads, ag, cases):
    res = {}
    for m, b in cases:
        best, short = refio.ref_se_map(index, reads, ag=ag, m=m, b=b, threads=4)
        res[f"best_m{m}_b{b}"] = best
        res[f"short_m{m}_b{b}"] = np.uint32(short)
    return res


def main():
    assert refio.have_reference(), "build oracle/_ref first (make -C oracle)"
    tmp = tempfile.mkdtemp(prefix="walt_golden_")
    try:
        chroms = build_genome()
        fasta = os.path.join(tmp, "genome.fa")
        synth.write_fasta(fasta, chroms)
        index = os.path.join(tmp, "g.dbindex")
        refio.ref_makedb(fasta, index)
        hdr = refio.read_header(index)
        g = {"names": np.array(hdr.names), "lengths": hdr.lengths,
             "fasta": np.frombuffer(open(fasta, "rb").read(), np.uint8)}
        for sfx in refio.SUFFIXES:
            sub = refio.read_subindex(index + sfx, hdr.genome_len)
            g["seq" + sfx] = sub.seq
            g["index" + sfx] = sub.index
            g["strand" + sfx] = np.frombuffer(sub.strand.encode(), np.uint8)
        np.savez_compressed(os.path.join(HERE, "genome.npz"), **g)

        # the chromosomes as makedb saw them (N replaced): take them back from _CT00? No --
        # reads are simulated from the FASTA (N -> random inside synth).
        rng = np.random.default_rng(77)
        cases = [(6, 5000), (0, 5000), (8, 5000), (6, 1), (6, 3)]
        big = [c for c in chroms[:2]]
        r_ct = synth.simulate_se_reads(big, 3000, 100, seed=21, n_frac=0.0)
        np.savez_compressed(os.path.join(HERE, "se_ct.npz"), reads=r_ct,
                            **se_cases(index, r_ct, False, cases))
        r_ga = synth.simulate_se_reads(big, 2000, 150, seed=22, a_rich=True, n_frac=0.0)
        np.savez_compressed(os.path.join(HERE, "se_ga.npz"), reads=r_ga,
                            **se_cases(index, r_ga, True, cases[:3]))
        er = edge_reads(chroms, rng)
        buf, offs = refio.pack_reads(er)
        d = {"buf": buf, "offs": offs}
        for ag in (False, True):
            for k, v in se_cases(index, er, ag, [(6, 5000), (2, 5000), (8, 2)]).items():
                d[("ga_" if ag else "ct_") + k] = v
        np.savez_compressed(os.path.join(HERE, "se_edge.npz"), **d)

        m1, m2 = synth.simulate_pe_reads(big, 1500, 100, seed=23, insert_mean=220, insert_sd=60,
                                         insert_min=100, insert_max=600, n_frac=0.0)
        d = {"m1": m1, "m2": m2}
        for m, k in ((6, 50), (8, 3), (4, 2)):
            for mate, reads, ag in ((1, m1, False), (2, m2, True)):
                ranked, sizes = refio.ref_pe_mate(index, reads, ag, m=m, b=5000, top_k=k, threads=4)
                d[f"ranked{mate}_m{m}_k{k}"] = ranked
                d[f"sizes{mate}_m{m}_k{k}"] = sizes
        np.savez_compressed(os.path.join(HERE, "pe.npz"), **d)

        # ---- CLI-level goldens -------------------------------------------------------------
        shutil.rmtree(CLI, ignore_errors=True)
        os.makedirs(CLI)
        ad = "AGATCGGAAGAGC"
        se = synth.simulate_se_reads(big, 1200, 100, seed=31)
        # lowercase + adaptor read-through + short reads
        se_list = [r.tobytes() for r in se]
        for i in range(0, 1200, 40):
            se_list[i] = se_list[i].lower()
        for i in range(5, 1200, 17):
            cut = 40 + (i * 7) % 55
            se_list[i] = (se_list[i][:cut] + ad.encode() + se_list[i])[:100]
        lens = np.full(1200, 100)
        lens[3::50] = 37
        lens[4::50] = 38
        lens[7::50] = 71
        synth.write_fastq(os.path.join(CLI, "se_reads.fastq"), se_list, prefix="s", lengths=lens)
        ga = synth.simulate_se_reads(big, 800, 150, seed=32, a_rich=True)
        synth.write_fastq(os.path.join(CLI, "ga_reads.fastq"), ga, prefix="g")
        synth.write_fastq(os.path.join(CLI, "crlf_reads.fastq"), [r.tobytes() for r in se[:60]],
                          prefix="c", crlf=True, final_newline=False)
        p1, p2 = synth.simulate_pe_reads(big, 1000, 100, seed=33, insert_mean=200, insert_sd=70,
                                         insert_min=60, insert_max=700, adaptor=ad)
        synth.write_fastq(os.path.join(CLI, "pe_reads_1.fastq"), p1, prefix="p")
        synth.write_fastq(os.path.join(CLI, "pe_reads_2.fastq"), p2, prefix="p")

        cases = [
            {"name": "se_sam", "args": ["-r", "se_reads.fastq", "-sam", "-u", "-a"]},
            {"name": "se_mr", "args": ["-r", "se_reads.fastq", "-u", "-a"]},
            {"name": "se_mr_plain", "args": ["-r", "se_reads.fastq"]},
            {"name": "se_clip", "args": ["-r", "se_reads.fastq", "-sam", "-u", "-a", "-C", ad, "-m", "4"]},
            {"name": "se_smallN", "args": ["-r", "se_reads.fastq", "-sam", "-u", "-a", "-N", "500"]},
            {"name": "se_b1", "args": ["-r", "se_reads.fastq", "-sam", "-a", "-b", "1", "-m", "8"]},
            {"name": "se_ga_sam", "args": ["-r", "ga_reads.fastq", "-sam", "-u", "-a", "-A"]},
            {"name": "se_ga_mr", "args": ["-r", "ga_reads.fastq", "-u", "-a", "-A"]},
            {"name": "se_crlf", "args": ["-r", "crlf_reads.fastq", "-sam", "-u", "-a"]},
            {"name": "pe_sam", "args": ["-1", "pe_reads_1.fastq", "-2", "pe_reads_2.fastq", "-sam", "-u", "-a"]},
            {"name": "pe_mr", "args": ["-1", "pe_reads_1.fastq", "-2", "pe_reads_2.fastq", "-u", "-a"]},
            {"name": "pe_clip_k3", "args": ["-1", "pe_reads_1.fastq", "-2", "pe_reads_2.fastq", "-sam", "-u",
                                            "-a", "-C", ad, "-k", "3", "-L", "400", "-m", "8"]},
            {"name": "pe_smallN", "args": ["-1", "pe_reads_1.fastq", "-2", "pe_reads_2.fastq", "-sam",
                                           "-N", "300"]},
        ]
        for c in cases:
            outdir = os.path.join(CLI, c["name"])
            os.makedirs(outdir)
            args = [a if not a.endswith(".fastq") else os.path.join(CLI, a) for a in c["args"]]
            out = os.path.join(outdir, "out")
            r = refio.ref_walt(["-i", index, "-o", out, "-t", "4"] + args, check=False)
            c["returncode"] = r.returncode
            c["files"] = sorted(os.listdir(outdir))
        json.dump(cases, open(os.path.join(CLI, "cases.json"), "w"), indent=1)
        subprocess.run(["du", "-sh", HERE])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
