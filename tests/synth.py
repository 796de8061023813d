"""Seeded synthetic genomes and simulated bisulfite reads (numpy, vectorised).

Shapes follow SURVEY.md section 8(d): uniform ACGT chromosomes; directional WGBS reads
(strand 50/50, each C kept with p=0.05 else C->T, k substitutions with
k in {0,0,0,1,2,3,5,7}, 1 % fully random reads, 2 % reads with one 'N').
Never touches glibc rand() so the reference's N-replacement stream stays untouched.
"""
from __future__ import annotations

import numpy as np

_COMP = np.zeros(256, dtype=np.uint8)
for a, b in zip(b"ACGTNacgtn", b"TGCANtgcan"):
    _COMP[a] = b
_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def make_genome(lengths, seed, n_frac=0.0):
    """-> list of (name, uint8 ASCII array)."""
    rng = np.random.default_rng(seed)
    chroms = []
    for i, n in enumerate(lengths):
        seq = _ACGT[rng.integers(0, 4, size=n)]
        if n_frac > 0:
            seq = seq.copy()
            seq[rng.random(n) < n_frac] = ord("N")
        chroms.append((f"chr{i + 1}", seq))
    return chroms


def make_repeat_genome(lengths, seed, n_families=20, fam_len=(300, 2000), copies=(5, 60),
                       divergence=0.02, repeat_frac=0.4):
    """Repeat-heavy genome: uniform background with diverged copies of repeat families."""
    rng = np.random.default_rng(seed)
    chroms = make_genome(lengths, seed + 1)
    fams = [_ACGT[rng.integers(0, 4, size=rng.integers(*fam_len))] for _ in range(n_families)]
    for _, seq in chroms:
        budget = int(len(seq) * repeat_frac)
        guard = 0
        while budget > 0 and guard < 10000:
            guard += 1
            f = fams[rng.integers(0, n_families)]
            if len(seq) <= len(f) + 1:
                f = f[: max(1, len(seq) // 2)]
            for _ in range(rng.integers(*copies)):
                p = rng.integers(0, len(seq) - len(f))
                c = f.copy()
                m = rng.random(len(c)) < divergence
                c[m] = _ACGT[rng.integers(0, 4, size=int(m.sum()))]
                seq[p:p + len(c)] = c
                budget -= len(c)
                if budget <= 0:
                    break
    return chroms


def write_fasta(path, chroms, width=60):
    with open(path, "wb") as f:
        for name, seq in chroms:
            f.write(b">" + name.encode() + b"\n")
            n = len(seq)
            full = (n // width) * width
            if full:
                body = seq[:full].reshape(-1, width)
                lines = np.concatenate([body, np.full((body.shape[0], 1), 10, np.uint8)], axis=1)
                f.write(lines.tobytes())
            if n > full:
                f.write(seq[full:].tobytes() + b"\n")


def revcomp(a):
    return _COMP[a[..., ::-1]]


def _sample_windows(chroms, n, rl, rng):
    lens = np.array([len(s) for _, s in chroms], dtype=np.int64)
    ok = lens > rl + 2
    w = np.where(ok, lens, 0).astype(np.float64)
    cid = rng.choice(len(chroms), size=n, p=w / w.sum())
    pos = (rng.random(n) * (lens[cid] - rl - 1)).astype(np.int64)
    cat = np.concatenate([s for _, s in chroms])
    starts = np.concatenate([[0], np.cumsum(lens)])[:-1]
    gpos = starts[cid] + pos
    win = cat[gpos[:, None] + np.arange(rl)[None, :]]
    return win, cid, pos


def _mutate(reads, rng, ks=(0, 0, 0, 1, 2, 3, 5, 7)):
    n, rl = reads.shape
    k = np.array(ks)[rng.integers(0, len(ks), size=n)]
    for t in range(max(ks)):
        rows = np.nonzero(k > t)[0]
        if rows.size == 0:
            break
        cols = rng.integers(0, rl, size=rows.size)
        old = reads[rows, cols]
        new = _ACGT[rng.integers(0, 4, size=rows.size)]
        same = new == old
        new[same] = _ACGT[(np.searchsorted(_ACGT, old[same]) + 1) % 4]
        reads[rows, cols] = new
    return reads


def simulate_se_reads(chroms, n, rl, seed, a_rich=False, random_frac=0.01, n_frac=0.02,
                      conv_rate=0.95, mutate=True):
    """-> (n, rl) uint8 ASCII reads.  T-rich (mate-1 style) unless a_rich (mate-2 style)."""
    rng = np.random.default_rng(seed)
    win, _, _ = _sample_windows(chroms, n, rl, rng)
    win = np.where(win == ord("N"), _ACGT[rng.integers(0, 4, size=win.shape)], win)
    minus = rng.random(n) < 0.5
    frag = win.copy()
    frag[minus] = revcomp(win[minus])
    conv = (frag == ord("C")) & (rng.random(frag.shape) < conv_rate)
    frag[conv] = ord("T")
    if mutate:
        frag = _mutate(frag, rng)
    rnd = rng.random(n) < random_frac
    frag[rnd] = _ACGT[rng.integers(0, 4, size=(int(rnd.sum()), rl))]
    if a_rich:
        frag = revcomp(frag)
    withn = np.nonzero(rng.random(n) < n_frac)[0]
    frag[withn, rng.integers(0, rl, size=withn.size)] = ord("N")
    return np.ascontiguousarray(frag)


def simulate_pe_reads(chroms, n, rl, seed, insert_mean=300, insert_sd=50, insert_min=None,
                      insert_max=1000, random_frac=0.01, n_frac=0.02, conv_rate=0.95,
                      adaptor=None):
    """-> (mate1, mate2) each (n, rl) uint8.  mate1 T-rich from the fragment 5' end, mate2 =
    revcomp of the fragment 3' end.  With `adaptor`, inserts shorter than rl read through."""
    rng = np.random.default_rng(seed)
    lo = insert_min if insert_min is not None else rl
    ins = np.clip(rng.normal(insert_mean, insert_sd, size=n).astype(np.int64), lo, insert_max)
    maxins = int(ins.max())
    win, _, _ = _sample_windows(chroms, n, maxins, rng)
    win = np.where(win == ord("N"), _ACGT[rng.integers(0, 4, size=win.shape)], win)
    minus = rng.random(n) < 0.5
    ad = np.frombuffer((adaptor or "").encode(), dtype=np.uint8)
    m1 = np.empty((n, rl), np.uint8)
    m2 = np.empty((n, rl), np.uint8)
    filler = _ACGT[rng.integers(0, 4, size=(n, rl))]
    for L in np.unique(ins):
        rows = np.nonzero(ins == L)[0]
        frag = win[rows, :L].copy()
        mm = minus[rows]
        frag[mm] = revcomp(frag[mm])
        conv = (frag == ord("C")) & (rng.random(frag.shape) < conv_rate)
        frag[conv] = ord("T")
        rc = revcomp(frag)
        if L >= rl:
            m1[rows] = frag[:, :rl]
            m2[rows] = rc[:, :rl]
        else:
            tail = np.concatenate([ad, filler[0]])[: rl - L] if ad.size else filler[0][: rl - L]
            m1[rows, :L] = frag
            m1[rows, L:] = tail
            m2[rows, :L] = rc
            m2[rows, L:] = tail
    m1 = _mutate(m1, rng, ks=(0, 0, 0, 0, 1, 1, 2, 3))
    m2 = _mutate(m2, rng, ks=(0, 0, 0, 0, 1, 1, 2, 3))
    rnd = rng.random(n) < random_frac
    m1[rnd] = _ACGT[rng.integers(0, 4, size=(int(rnd.sum()), rl))]
    for m in (m1, m2):
        withn = np.nonzero(rng.random(n) < n_frac)[0]
        m[withn, rng.integers(0, rl, size=withn.size)] = ord("N")
    return m1, m2


def _qual(i, n):
    return ((np.int64(i) * 7 + np.arange(n, dtype=np.int64) * 3) % 40 + 33).astype(np.uint8)


def write_fastq(path, reads, prefix="r", lengths=None, crlf=False, final_newline=True):
    """reads: (n, rl) uint8 array or list of bytes.  Names are '<prefix><i> extra' so the
    loader's cut-at-first-space rule (mapping.cpp:88-93) is exercised."""
    eol = b"\r\n" if crlf else b"\n"
    n = len(reads)
    if isinstance(reads, np.ndarray) and reads.ndim == 2 and lengths is None and not crlf \
            and final_newline and n > 0:
        rl = reads.shape[1]
        q = ((np.arange(n, dtype=np.int64)[:, None] * 7 + np.arange(rl, dtype=np.int64)[None, :] * 3)
             % 40 + 33).astype(np.uint8)
        nl = np.full((n, 1), 10, np.uint8)
        body = np.concatenate([reads, nl, np.full((n, 1), ord("+"), np.uint8), nl, q, nl], axis=1)
        with open(path, "wb") as f:
            for i in range(n):
                f.write(b"@%s%d extra\n" % (prefix.encode(), i))
                f.write(body[i].tobytes())
        return
    with open(path, "wb") as f:
        for i in range(n):
            s = reads[i].tobytes() if isinstance(reads[i], np.ndarray) else bytes(reads[i])
            if lengths is not None:
                s = s[: int(lengths[i])]
            q = _qual(i, len(s)).tobytes()
            rec = b"@" + f"{prefix}{i}".encode() + b" extra" + eol + s + eol + b"+" + eol + q
            if i + 1 < n or final_newline:
                rec += eol
            f.write(rec)
