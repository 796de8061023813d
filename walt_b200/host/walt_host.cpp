// walt_host.cpp -- host side of the B200 WALT engine (libwalthost.so, see include/walt_host.h).
//
// Byte-compatible restatements of the reference's text paths around the mapping call:
//   LoadReadsFromFastqFile      src/walt/mapping.cpp:65-121   (+ util.hpp:150-217)
//   ReadIndexHeadInfo / SAMHead src/walt/reference.cpp:381-440
//   StatInfoUpdate, OutputSingleResults, OutputSingleSAM, StatSingleReads::tostring
//                               src/walt/mapping.cpp:47-63,318-419
//   MergePairedEndResults (after its pairing loop), OutputBestPairedResults,
//   GetBestMatch4Single, GetSAMFLAG, OutputPairedSAM, StatPairedReads::tostring
//                               src/walt/paired.cpp:52-95,210-435,515-569
// The pairing loop itself and all mapping run on the GPU (walt_engine.cu); nothing here
// computes alignments.  glibc rand()/srand() are called exactly like the reference does, so
// the N-replacement stream is identical.
#include "../../include/walt_host.h"

#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <fstream>
#include <numeric>
#include <sstream>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;
int fail(const std::string& m) { g_err = m; return -1; }

constexpr size_t LINE_CAP = 1000;      // MAX_LINE_LENGTH, util.hpp:43
constexpr uint32_t MIN_READ_LEN = 38;  // MINIMALREADLEN

inline char comp(char c) {
  switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; }
  return c;
}

// buffered text sink
struct Sink {
  FILE* f = nullptr;
  std::string buf;
  bool open(const std::string& path, const char* mode) {
    f = fopen(path.c_str(), mode);
    buf.reserve(1u << 20);
    return f != nullptr;
  }
  void flush() { if (f && !buf.empty()) { fwrite(buf.data(), 1, buf.size(), f); buf.clear(); } }
  void maybe_flush() { if (buf.size() > (1u << 20) - 4096) flush(); }
  void close() { flush(); if (f) fclose(f); f = nullptr; }
  void str(const char* s, size_t n) { buf.append(s, n); }
  template <size_t N> void lit(const char (&s)[N]) { buf.append(s, N - 1); }
  void str(const std::string& s) { buf.append(s); }
  void ch(char c) { buf.push_back(c); }
  void u32(uint32_t v) { char t[12]; int n = 0; do { t[n++] = (char)('0' + v % 10); v /= 10; } while (v); while (n) buf.push_back(t[--n]); }
  void i32(int32_t v) { if (v < 0) { buf.push_back('-'); u32((uint32_t)(-(int64_t)v)); } else u32((uint32_t)v); }
  void revcomp(const char* s, size_t n) { for (size_t i = n; i-- > 0;) buf.push_back(comp(s[i])); }
  void rev(const char* s, size_t n) { for (size_t i = n; i-- > 0;) buf.push_back(s[i]); }
};

}  // namespace

struct walt_chroms {
  std::vector<std::string> names;
  std::vector<uint32_t> lengths, starts;
  uint32_t genome_len = 0, size_of_index = 0;
  void finish() {
    starts.assign(names.size() + 1, 0);
    for (size_t i = 0; i < names.size(); ++i) starts[i + 1] = starts[i] + lengths[i];
    genome_len = starts.back();
  }
  // getChromID, reference.cpp:43-60
  uint32_t chrom_of(uint32_t pos) const {
    uint32_t l = 0, h = (uint32_t)names.size();
    while (l < h) { uint32_t m = (l + h + 1) >> 1; if (pos >= starts[m]) l = m; else h = m - 1; }
    return l;
  }
};

struct walt_fastq { FILE* f = nullptr; };

struct walt_batch {
  std::string seqs, quals, names;
  std::vector<uint64_t> seq_off, qual_off, name_off;   // n + 1 each; names/quals NUL-terminated copies
  std::vector<uint8_t> packed;                         // 2-bit form of seqs (walt_batch_packed), built on demand
  uint32_t n = 0;
  void clear() { seqs.clear(); quals.clear(); names.clear(); packed.clear(); seq_off.assign(1, 0); qual_off.assign(1, 0); name_off.assign(1, 0); n = 0; }
  const char* seq(uint32_t i) const { return seqs.data() + seq_off[i]; }
  uint32_t seq_len(uint32_t i) const { return (uint32_t)(seq_off[i + 1] - seq_off[i]); }
  const char* qual(uint32_t i) const { return quals.data() + qual_off[i]; }
  uint32_t qual_len(uint32_t i) const { return (uint32_t)(qual_off[i + 1] - qual_off[i] - 1); }
  const char* name(uint32_t i) const { return names.data() + name_off[i]; }
  uint32_t name_len(uint32_t i) const { return (uint32_t)(name_off[i + 1] - name_off[i] - 1); }
};

namespace {

// util.hpp:189-198
size_t similarity(const char* s, size_t len, size_t pos, const char* ad, size_t ad_len) {
  const size_t lim = std::min(std::min(len - pos, ad_len), (size_t)14);
  size_t c = 0;
  for (size_t i = 0; i < lim; ++i) c += s[pos + i] == ad[i];
  return c;
}

struct SingleStats {
  uint32_t total = 0, unique = 0, ambiguous = 0, unmapped = 0, n_short = 0;
  void update(uint32_t times) { ++total; if (times == 0) ++unmapped; else if (times == 1) ++unique; else ++ambiguous; }
  // StatSingleReads::tostring, mapping.cpp:47-63
  std::string text(size_t n_tabs) const {
    std::string t(4 * n_tabs, ' ');
    std::ostringstream o;
    o << t << "total_reads: " << total << std::endl
      << t << "mapped:" << std::endl
      << t << "    unique: " << unique << std::endl
      << t << "    percent_unique: " << (100.0 * (double)unique) / (double)total << std::endl
      << t << "    ambiguous: " << ambiguous << std::endl
      << t << "unmapped: " << unmapped << std::endl
      << t << "min_read_length: " << MIN_READ_LEN << std::endl
      << t << "too_short: " << n_short;
    return o.str();
  }
};

// the three MR destinations + flags of one StatSingleReads (mapping.hpp:55-108)
struct SingleOut {
  bool ambiguous = false, unmapped = false, sam = false;
  Sink famb, funm;
  SingleStats st;
  bool open(const std::string& prefix, bool amb, bool unm, bool is_sam) {
    ambiguous = amb; unmapped = unm; sam = is_sam;
    if (ambiguous && !sam && !famb.open(prefix + "_ambiguous", "w")) return false;
    if (unmapped && !sam && !funm.open(prefix + "_unmapped", "w")) return false;
    return true;
  }
  void close() { famb.close(); funm.close(); }
};

// OutputUniquelyAndAmbiguousMapped, mapping.cpp:329-349 (seq/qual already oriented)
void mr_mapped(Sink& o, const walt_chroms& g, const walt_best& bm, const char* name, uint32_t name_len, const char* seq,
               const char* qual, uint32_t len, uint32_t qlen, bool ag) {
  const uint32_t chr = g.chrom_of(bm.genome_pos);
  uint32_t start = bm.genome_pos - g.starts[chr];
  if (bm.strand == '-') start = g.lengths[chr] - start - len;
  const uint32_t end = start + len;
  char strand = bm.strand;
  if (ag) strand = bm.strand == '+' ? '-' : '+';
  o.str(g.names[chr]); o.ch('\t'); o.u32(start); o.ch('\t'); o.u32(end); o.ch('\t'); o.str(name, name_len); o.ch('\t');
  o.u32(bm.mismatch); o.ch('\t'); o.ch(strand); o.ch('\t');
  if (ag) o.revcomp(seq, len); else o.str(seq, len);
  o.ch('\t');
  if (ag) o.rev(qual, qlen); else o.str(qual, qlen);
  o.ch('\n');
  o.maybe_flush();
}

// OutputSingleResults, mapping.cpp:358-380
void mr_single(Sink& fout, SingleOut& so, const walt_chroms& g, const walt_best& bm, const char* name, uint32_t name_len,
               const char* seq, const char* qual, uint32_t len, uint32_t qlen, bool ag) {
  if (bm.times == 0 && so.unmapped) {
    Sink& o = so.funm;
    o.str(name, name_len); o.ch('\t');
    if (ag) o.revcomp(seq, len); else o.str(seq, len);
    o.ch('\t');
    if (ag) o.rev(qual, qlen); else o.str(qual, qlen);
    o.ch('\n');
    o.maybe_flush();
  } else if (bm.times == 1) {
    mr_mapped(fout, g, bm, name, name_len, seq, qual, len, qlen, ag);
  } else if (bm.times >= 2 && so.ambiguous) {
    mr_mapped(so.famb, g, bm, name, name_len, seq, qual, len, qlen, ag);
  }
}

// OutputSingleSAM, mapping.cpp:382-419
void sam_single(Sink& o, const SingleOut& so, const walt_chroms& g, const walt_best& bm, const char* name,
                uint32_t name_len, const char* seq, const char* qual, uint32_t len, uint32_t qlen) {
  const bool rev = bm.strand == '-';
  int flag = 0;
  flag += bm.times == 0 ? 0x4 : 0;
  flag += rev ? 0x10 : 0;
  flag += bm.times >= 2 ? 0x100 : 0;
  const bool print_unmapped = bm.times == 0 && so.unmapped;
  const bool print_mapped = bm.times == 1 || (bm.times >= 2 && so.ambiguous);
  if (!print_unmapped && !print_mapped) return;
  o.str(name, name_len); o.ch('\t'); o.i32(flag); o.ch('\t');
  if (print_unmapped) {
    o.lit("*\t0\t255\t*\t*\t0\t0\t");
  } else {
    const uint32_t chr = g.chrom_of(bm.genome_pos);
    uint32_t start = bm.genome_pos - g.starts[chr];
    if (rev) start = g.lengths[chr] - start - len;
    o.str(g.names[chr]); o.ch('\t'); o.u32(start + 1); o.lit("\t255\t"); o.u32(len); o.lit("M\t*\t0\t0\t");
  }
  if (rev) o.revcomp(seq, len); else o.str(seq, len);
  o.ch('\t');
  if (rev) o.rev(qual, qlen); else o.str(qual, qlen);
  o.lit("\tNM:i:");
  o.u32(print_unmapped ? 0u : bm.mismatch);
  o.ch('\n');
  o.maybe_flush();
}

void sam_header(Sink& o, const walt_chroms& g) {
  o.lit("@HD\tVN:1.0\n");
  for (size_t i = 0; i < g.names.size(); ++i) {
    o.lit("@SQ\tSN:"); o.str(g.names[i]); o.lit("\tLN:"); o.u32(g.lengths[i]); o.ch('\n');
  }
  o.lit("@PG\tID:WALT\tVN:1.0\tCL:walt\n");   // walt_version "1.0" (util.hpp:41); CL is the literal "walt"
}

}  // namespace

struct walt_se_writer {
  const walt_chroms* g = nullptr;
  std::string path;
  bool ag = false, sam = false;
  Sink out;
  SingleOut so;
};

struct walt_pe_writer {
  const walt_chroms* g = nullptr;
  std::string path;
  uint32_t m = 6, top_k = 50;
  int frag_range = 1000;
  bool sam = false, pbat = false;
  Sink out;
  SingleOut so1, so2;
  uint32_t total_pairs = 0, unique_pairs = 0, ambiguous_pairs = 0, unmapped_pairs = 0;
  std::vector<uint32_t> frag_count;
};

namespace {

// ForwardChromPosition, paired.cpp:98-104
void forward_pos(const walt_chroms& g, uint32_t pos, char strand, uint32_t chr, uint32_t len, uint32_t& s, uint32_t& e) {
  s = pos - g.starts[chr];
  s = strand == '+' ? s : g.lengths[chr] - s - len;
  e = s + len;
}

// GetBestMatch4Single, paired.cpp:296-318
walt_best best_for_single(const walt_cand* r, uint32_t n, uint32_t m) {
  walt_best b; b.genome_pos = 0; b.times = 0; b.mismatch = m; b.strand = '+'; b.pad[0] = b.pad[1] = b.pad[2] = 0;
  for (int i = (int)n - 1; i >= 0; --i) {
    if (r[i].mismatch < b.mismatch) {
      b.genome_pos = r[i].genome_pos; b.times = 1; b.strand = r[i].strand; b.mismatch = r[i].mismatch;
    } else if (r[i].mismatch == b.mismatch) {
      if (b.genome_pos == r[i].genome_pos) continue;
      b.genome_pos = r[i].genome_pos; b.strand = r[i].strand; b.times++;
    } else {
      break;
    }
  }
  return b;
}

// OutputBestPairedResults, paired.cpp:210-294.  Returns the fragment length; writes the MR
// FRAG line unless `sam`.
int best_pair_out(Sink& o, const walt_chroms& g, const walt_cand& r1, const walt_cand& r2, int frag_range, const char* name,
                  uint32_t name_len, const char* seq1, const char* q1, uint32_t len1, const char* seq2, const char* q2,
                  uint32_t len2, bool sam) {
  const uint32_t c1 = g.chrom_of(r1.genome_pos), c2 = g.chrom_of(r2.genome_pos);
  uint32_t s1, e1, s2, e2;
  forward_pos(g, r1.genome_pos, r1.strand, c1, len1, s1, e1);
  forward_pos(g, r2.genome_pos, r2.strand, c2, len2, s2, e2);
  const bool plus = r1.strand == '+';
  const uint32_t os = std::max(s1, s2), oe = std::min(e1, e2);
  const uint32_t one_l = plus ? s1 : std::max(oe, s1);
  const uint32_t one_r = plus ? std::min(os, e1) : e1;
  const uint32_t two_l = plus ? std::max(oe, s2) : s2;
  const uint32_t two_r = plus ? e2 : std::min(os, e2);
  const int len = plus ? (int)(two_r - one_l) : (int)(one_r - two_l);
  if (sam) return len;

  std::string seq((size_t)len, 'N'), scr((size_t)len, 'B');
  if (len > 0 && len <= frag_range) {
    std::string s2r(len2, 'N'), q2r(len2, 'B');
    for (uint32_t i = 0; i < len2; ++i) { s2r[i] = comp(seq2[len2 - 1 - i]); q2r[i] = q2[len2 - 1 - i]; }
    const uint32_t lim_one = one_r - one_l;
    std::copy(seq1, seq1 + lim_one, seq.begin());
    std::copy(q1, q1 + lim_one, scr.begin());
    const uint32_t lim_two = two_r - two_l;
    std::copy(s2r.end() - lim_two, s2r.end(), seq.end() - lim_two);
    std::copy(q2r.end() - lim_two, q2r.end(), scr.end() - lim_two);
    if (os < oe) {
      const int info_one = (int)(len1 - ((uint32_t)std::count(seq1, seq1 + len1, 'N') + r1.mismatch));
      const int info_two = (int)(len2 - ((uint32_t)std::count(s2r.begin(), s2r.end(), 'N') + r2.mismatch));
      if (info_one >= info_two) {
        const uint32_t a = plus ? (os - s1) : (e1 - oe), b = plus ? (oe - s1) : (e1 - os);
        std::copy(seq1 + a, seq1 + b, seq.begin() + lim_one);
        std::copy(q1 + a, q1 + b, scr.begin() + lim_one);
      } else {
        const uint32_t a = plus ? (os - s2) : (e2 - oe), b = plus ? (oe - s2) : (e2 - os);
        std::copy(s2r.begin() + a, s2r.begin() + b, seq.begin() + lim_one);
        std::copy(q2r.begin() + a, q2r.begin() + b, scr.begin() + lim_one);
      }
    }
  }
  const uint32_t start = plus ? s1 : s2;
  o.str(g.names[c1]); o.ch('\t'); o.u32(start); o.ch('\t'); o.u32(start + (uint32_t)len); o.lit("\tFRAG:");
  o.str(name, name_len); o.ch('\t'); o.u32(r1.mismatch + r2.mismatch); o.ch('\t'); o.ch(r1.strand); o.ch('\t');
  o.str(seq); o.ch('\t'); o.str(scr); o.ch('\n');
  o.maybe_flush();
  return len;
}

// GetSAMFLAG, paired.cpp:80-95
int sam_flag(bool paired_mapped, bool unmapped, bool next_unmapped, bool rev, bool next_rev, bool first, bool secondary) {
  return 0x1 + (paired_mapped ? 0x2 : 0) + (unmapped ? 0x4 : 0) + (next_unmapped ? 0x8 : 0) + (rev ? 0x10 : 0) +
         (next_rev ? 0x20 : 0) + (first ? 0x40 : 0x80) + (secondary ? 0x100 : 0);
}

// one line of OutputPairedSAM, paired.cpp:333-435
void sam_mate_line(Sink& o, const SingleOut& so, const walt_chroms& g, const walt_best& bm, uint32_t chr, const char* name,
                   uint32_t name_len, int flag, uint32_t pos, const std::string& rnext, uint32_t pnext, int tlen,
                   const char* seq, const char* qual, uint32_t len, uint32_t qlen, uint32_t nm) {
  const bool print_unmapped = bm.times == 0 && so.unmapped;
  const bool print_mapped = bm.times == 1 || (bm.times >= 2 && so.ambiguous);
  if (!print_unmapped && !print_mapped) return;
  const bool rev = bm.strand == '-';
  o.str(name, name_len); o.ch('\t'); o.i32(flag); o.ch('\t');
  if (print_unmapped) {
    o.lit("*\t"); o.u32(pos); o.lit("\t255\t*\t");
  } else {
    o.str(g.names[chr]); o.ch('\t'); o.u32(pos); o.lit("\t255\t"); o.u32(len); o.lit("M\t");
  }
  o.str(rnext); o.ch('\t'); o.u32(pnext); o.ch('\t'); o.i32(tlen); o.ch('\t');
  if (rev) o.revcomp(seq, len); else o.str(seq, len);
  o.ch('\t');
  if (rev) o.rev(qual, qlen); else o.str(qual, qlen);
  o.lit("\tNM:i:"); o.u32(nm); o.ch('\n');
  o.maybe_flush();
}

}  // namespace

// ==========================================================================================
extern "C" {

const char* walt_host_last_error(void) { return g_err.c_str(); }

// ---- chromosome table ----------------------------------------------------------------------
walt_chroms* walt_chroms_read(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) { fail(std::string("cannot open ") + path + ": " + strerror(errno)); return nullptr; }
  walt_chroms* c = new walt_chroms;
  bool ok = true;
  uint32_t n = 0;
  ok = fread(&n, 4, 1, f) == 1 && n > 0 && n < (1u << 24);
  for (uint32_t i = 0; ok && i < n; ++i) {
    uint32_t ln = 0;
    ok = fread(&ln, 4, 1, f) == 1 && ln <= 255;
    std::string nm(ln, '\0');
    if (ok && ln) ok = fread(&nm[0], 1, ln, f) == ln;
    c->names.push_back(nm);
  }
  c->lengths.resize(n);
  if (ok) ok = fread(c->lengths.data(), 4, n, f) == n;
  uint32_t glen = 0;
  if (ok) ok = fread(&glen, 4, 1, f) == 1;
  if (ok) ok = fread(&c->size_of_index, 4, 1, f) == 1;
  fclose(f);
  if (!ok) { delete c; fail(std::string("bad index head file: ") + path); return nullptr; }
  c->finish();
  return c;
}

walt_chroms* walt_chroms_create(uint32_t n_chr, const char* const* names, const uint32_t* lengths) {
  walt_chroms* c = new walt_chroms;
  for (uint32_t i = 0; i < n_chr; ++i) { c->names.emplace_back(names[i]); c->lengths.push_back(lengths[i]); }
  c->finish();
  return c;
}
void walt_chroms_free(walt_chroms* c) { delete c; }
uint32_t walt_chroms_count(const walt_chroms* c) { return (uint32_t)c->names.size(); }
const uint32_t* walt_chroms_lengths(const walt_chroms* c) { return c->lengths.data(); }
const char* walt_chroms_name(const walt_chroms* c, uint32_t i) { return c->names[i].c_str(); }

// ---- FASTQ ---------------------------------------------------------------------------------
walt_fastq* walt_fastq_open(const char* path) {
  FILE* f = fopen(path, "r");
  if (!f) { fail(std::string("cannot open input file ") + path); return nullptr; }
  static const size_t BUF = 4u << 20;
  setvbuf(f, nullptr, _IOFBF, BUF);
  walt_fastq* q = new walt_fastq; q->f = f;
  return q;
}
void walt_fastq_close(walt_fastq* f) { if (f) { if (f->f) fclose(f->f); delete f; } }
walt_batch* walt_batch_create(void) { walt_batch* b = new walt_batch; b->clear(); return b; }
void walt_batch_free(walt_batch* b) { delete b; }

size_t walt_clip_adaptor(const char* adaptor, char* s, size_t len) {
  // util.hpp:200-217, including its size_t arithmetic
  const size_t ad_len = strlen(adaptor);
  const size_t lim1 = len - 14 + 1;
  for (size_t i = 0; i < lim1; ++i)
    if (similarity(s, len, i, adaptor, ad_len) >= 11) { memset(s + i, 'N', len - i); return len - i; }
  const size_t lim2 = len - 5 + 1;
  for (size_t i = lim1; i < lim2; ++i)
    if (similarity(s, len, i, adaptor, ad_len) >= len - i - 1) { memset(s + i, 'N', len - i); return len - i; }
  return 0;
}

int64_t walt_fastq_next_batch(walt_fastq* fq, walt_batch* b, uint32_t max_reads, const char* adaptor) {
  if (!fq || !b) return fail("bad argument");
  b->clear();
  srand(0);                                   // mapping.cpp:73
  char line[LINE_CAP];
  int code = 0;
  uint64_t line_count = 0;
  const uint64_t lim = (uint64_t)max_reads * 4u;
  const bool clip = adaptor && adaptor[0];
  std::string pend_name;
  bool have_name = false, have_seq = false;
  size_t pend_seq_off = 0;
  while (line_count < lim && fgets(line, (int)LINE_CAP, fq->f)) {
    size_t len = strlen(line);
    if (len) line[--len] = 0;                 // the last character goes, whatever it is
    if (len == 0) continue;                   // empty lines do not advance the record state
    switch (code) {
      case 0: {
        const char* sp = (const char*)memchr(line, ' ', len);
        const size_t end = sp ? (size_t)(sp - line) : len;
        pend_name.assign(line + 1, end > 0 ? end - 1 : 0);
        have_name = true;
        break;
      }
      case 1: {
        if (clip) walt_clip_adaptor(adaptor, line, len);
        pend_seq_off = b->seqs.size();
        for (size_t i = 0; i < len; ++i) {
          char c = line[i];
          if (!(c == 'A' || c == 'C' || c == 'G' || c == 'T')) c = "ACGT"[rand() % 4];   // util.hpp:156-163
          b->seqs.push_back(c);
        }
        have_seq = true;
        break;
      }
      case 2: break;
      case 3: {
        // a record is committed when its quality line arrives (mapping.cpp:104-108)
        if (!have_seq) pend_seq_off = b->seqs.size();
        if (!have_name) pend_name.clear();
        b->seq_off.back() = pend_seq_off;
        b->seq_off.push_back(b->seqs.size());
        b->names.append(pend_name); b->names.push_back('\0'); b->name_off.push_back(b->names.size());
        b->quals.append(line, len); b->quals.push_back('\0'); b->qual_off.push_back(b->quals.size());
        b->n++;
        have_name = have_seq = false;
        break;
      }
    }
    ++line_count;
    if (++code == 4) code = 0;
  }
  // a trailing partial record leaves bases behind the last committed read: drop them
  b->seqs.resize(b->seq_off.back());
  return b->n;
}

uint64_t walt_packed_reads_bytes(const uint64_t* offs, uint32_t n) { return (offs[n] >> 2) + n + 16u; }

int walt_pack_reads(const char* seqs, const uint64_t* offs, uint32_t n, uint8_t* out) {
  if (!offs || !out || (n && !seqs)) return fail("bad argument");
  // A0 C1 G2 T3 (util.hpp:107-121): bits 1..2 of the ASCII letter, with the upper bit folded in
  unsigned bad = 0;
  for (uint32_t j = 0; j < n; ++j) {
    const char* s = seqs + offs[j];
    const uint64_t len = offs[j + 1] - offs[j];
    uint8_t* o = out + (offs[j] >> 2) + j;
    uint64_t i = 0;
    for (; i + 4 <= len; i += 4) {
      unsigned v = 0;
      for (unsigned t = 0; t < 4; ++t) {
        const unsigned c = (unsigned char)s[i + t], x = (c >> 1) & 3u, code = x ^ (x >> 1);
        bad |= (unsigned)("ACGT"[code] != (char)c);
        v = (v << 2) | code;
      }
      *o++ = (uint8_t)v;
    }
    if (i < len) {
      unsigned v = 0, t = 0;
      for (; i + t < len; ++t) {
        const unsigned c = (unsigned char)s[i + t], x = (c >> 1) & 3u, code = x ^ (x >> 1);
        bad |= (unsigned)("ACGT"[code] != (char)c);
        v |= code << (6u - 2u * t);
      }
      *o = (uint8_t)v;
    }
  }
  return bad ? fail("walt_pack_reads: a read holds a byte that is not A/C/G/T") : 0;
}

const uint8_t* walt_batch_packed(walt_batch* b) {
  if (!b) return nullptr;
  if (b->packed.empty()) {
    b->packed.assign(walt_packed_reads_bytes(b->seq_off.data(), b->n), 0);
    if (walt_pack_reads(b->seqs.data(), b->seq_off.data(), b->n, b->packed.data()) != 0) { b->packed.clear(); return nullptr; }
  }
  return b->packed.data();
}

uint32_t walt_batch_size(const walt_batch* b) { return b->n; }
const char* walt_batch_seqs(const walt_batch* b) { return b->seqs.data(); }
const uint64_t* walt_batch_offsets(const walt_batch* b) { return b->seq_off.data(); }
const char* walt_batch_name(const walt_batch* b, uint32_t i) { return b->name(i); }
const char* walt_batch_qual(const walt_batch* b, uint32_t i) { return b->qual(i); }

// ---- single end ------------------------------------------------------------------------------
int walt_write_sam_header(const char* output_path, const walt_chroms* chroms) {
  Sink s;
  if (!s.open(output_path, "a")) return fail(std::string("cannot open input file ") + output_path);
  sam_header(s, *chroms);
  s.close();
  return 0;
}

walt_se_writer* walt_se_writer_open(const char* output_path, const walt_chroms* chroms, int ag, int ambiguous,
                                    int unmapped, int sam) {
  walt_se_writer* w = new walt_se_writer;
  w->g = chroms; w->path = output_path; w->ag = ag != 0; w->sam = sam != 0;
  if (!w->out.open(w->path, "a")) { fail("cannot open input file " + w->path); delete w; return nullptr; }
  if (!w->so.open(w->path, ambiguous != 0, unmapped != 0, w->sam)) {
    fail("cannot open input file " + w->path + "_ambiguous/_unmapped"); w->out.close(); delete w; return nullptr;
  }
  if (w->sam) sam_header(w->out, *chroms);
  return w;
}

int walt_se_writer_write(walt_se_writer* w, const walt_batch* b, const walt_best* res, uint32_t n) {
  if (!w || !b || n > b->n) return fail("bad argument");
  for (uint32_t j = 0; j < n; ++j) {
    w->so.st.update(res[j].times);
    if (w->sam) sam_single(w->out, w->so, *w->g, res[j], b->name(j), b->name_len(j), b->seq(j), b->qual(j), b->seq_len(j), b->qual_len(j));
    else mr_single(w->out, w->so, *w->g, res[j], b->name(j), b->name_len(j), b->seq(j), b->qual(j), b->seq_len(j), b->qual_len(j), w->ag);
  }
  return 0;
}

void walt_se_writer_add_short(walt_se_writer* w, uint32_t n_short) { w->so.st.n_short += n_short; }

int walt_se_writer_close(walt_se_writer* w) {
  if (!w) return 0;
  w->out.close();
  w->so.close();
  std::ofstream ms(w->path + ".mapstats", std::ios::app);
  ms << w->so.st.text(0) << std::endl;
  const bool ok = (bool)ms;
  delete w;
  return ok ? 0 : fail("cannot write mapstats");
}

// ---- paired end --------------------------------------------------------------------------------
walt_pe_writer* walt_pe_writer_open(const char* output_path, const walt_chroms* chroms, uint32_t max_mismatches,
                                    uint32_t top_k, int frag_range, int ambiguous, int unmapped, int sam, int pbat) {
  walt_pe_writer* w = new walt_pe_writer;
  w->g = chroms; w->path = output_path; w->m = max_mismatches; w->top_k = top_k; w->frag_range = frag_range;
  w->sam = sam != 0; w->pbat = pbat != 0;
  w->frag_count.assign((size_t)frag_range + 1, 0);
  if (!w->out.open(w->path, "a")) { fail("cannot open input file " + w->path); delete w; return nullptr; }
  if (!w->so1.open(w->path + "_1", ambiguous != 0, unmapped != 0, w->sam) ||
      !w->so2.open(w->path + "_2", ambiguous != 0, unmapped != 0, w->sam)) {
    fail("cannot open the _1/_2 side files of " + w->path); w->out.close(); delete w; return nullptr;
  }
  if (w->sam) sam_header(w->out, *chroms);
  return w;
}

void walt_pe_writer_add_short(walt_pe_writer* w, uint32_t s1, uint32_t s2) { w->so1.st.n_short += s1; w->so2.st.n_short += s2; }

// One pair through MergePairedEndResults' output half (paired.cpp:515-569).  `r` holds the
// per-file-mate fields (c1/single1 belong to the first file's read).
static void write_pair(walt_pe_writer* w, const walt_batch* b1, const walt_batch* b2, uint32_t j, const walt_pe_result& r) {
  const walt_chroms& g = *w->g;
  // Under PBAT the mate that plays the reference's "mate 1" role (C->T, pairing-loop outer
  // index) is the second file's read; names and first/last flags stay with the files.
  const walt_batch* ba = w->pbat ? b2 : b1;   // role A: C->T mate
  const walt_batch* bb = w->pbat ? b1 : b2;   // role B: G->A mate
  const walt_cand& pa = w->pbat ? r.c2 : r.c1;
  const walt_cand& pb = w->pbat ? r.c1 : r.c2;
  const char* name = b1->name(j);            // QNAME of both lines = first file's read name
  const uint32_t name_len = b1->name_len(j);
  const uint32_t la = ba->seq_len(j), lb = bb->seq_len(j);
  walt_best ma, mb;
  ma.genome_pos = 0; ma.times = 0; ma.mismatch = w->m; ma.strand = '+';
  mb = ma;
  bool paired = false;
  int len = 0;
  if (r.pair.best_times == 1) {
    w->unique_pairs++;
    len = best_pair_out(w->out, g, pa, pb, w->frag_range, name, name_len, ba->seq(j), ba->qual(j), la,
                        bb->seq(j), bb->qual(j), lb, w->sam);
    if (len >= 0 && (size_t)len < w->frag_count.size()) w->frag_count[len]++;
    if (w->sam) {
      paired = true;
      ma.genome_pos = pa.genome_pos; ma.times = 1; ma.strand = pa.strand; ma.mismatch = pa.mismatch;
      mb.genome_pos = pb.genome_pos; mb.times = 1; mb.strand = pb.strand; mb.mismatch = pb.mismatch;
    }
  } else {
    if (r.pair.best_times >= 2) w->ambiguous_pairs++; else w->unmapped_pairs++;
    ma = w->pbat ? r.single2 : r.single1;
    mb = w->pbat ? r.single1 : r.single2;
    SingleOut& soa = w->pbat ? w->so2 : w->so1;
    SingleOut& sob = w->pbat ? w->so1 : w->so2;
    soa.st.update(ma.times);
    sob.st.update(mb.times);
    if (!w->sam) {
      mr_single(w->out, soa, g, ma, name, name_len, ba->seq(j), ba->qual(j), la, ba->qual_len(j), false);
      mr_single(w->out, sob, g, mb, name, name_len, bb->seq(j), bb->qual(j), lb, bb->qual_len(j), true);
    }
  }
  if (w->sam) {
    const uint32_t chra = g.chrom_of(ma.genome_pos), chrb = g.chrom_of(mb.genome_pos);
    uint32_t sa, ea, sb, eb;
    forward_pos(g, ma.genome_pos, ma.strand, chra, la, sa, ea);
    forward_pos(g, mb.genome_pos, mb.strand, chrb, lb, sb, eb);
    uint32_t mma = ma.mismatch, mmb = mb.mismatch;
    if (ma.times == 0) { sa = 0; mma = 0; } else sa += 1;
    if (mb.times == 0) { sb = 0; mmb = 0; } else sb += 1;
    const int tla = ma.strand == '+' ? len : -len, tlb = mb.strand == '+' ? len : -len;
    std::string rna = "=", rnb = "=";
    if (!paired) {
      rna = ma.times == 0 ? "*" : g.names[chra];
      rnb = mb.times == 0 ? "*" : g.names[chrb];
    }
    const bool a_first = !w->pbat;
    const int fa = sam_flag(paired, ma.times == 0, mb.times == 0, ma.strand == '-', mb.strand == '-', a_first, ma.times >= 2);
    const int fb = sam_flag(paired, mb.times == 0, ma.times == 0, mb.strand == '-', ma.strand == '-', !a_first, mb.times >= 2);
    const SingleOut& soa = w->pbat ? w->so2 : w->so1;
    const SingleOut& sob = w->pbat ? w->so1 : w->so2;
    auto line_a = [&]() { sam_mate_line(w->out, soa, g, ma, chra, name, name_len, fa, sa, rnb, sb, tla, ba->seq(j), ba->qual(j), la, ba->qual_len(j), mma); };
    auto line_b = [&]() { sam_mate_line(w->out, sob, g, mb, chrb, name, name_len, fb, sb, rna, sa, tlb, bb->seq(j), bb->qual(j), lb, bb->qual_len(j), mmb); };
    if (a_first) { line_a(); line_b(); } else { line_b(); line_a(); }
  }
}

int walt_pe_writer_write_compact(walt_pe_writer* w, const walt_batch* b1, const walt_batch* b2, const walt_pe_result* res,
                                 uint32_t n) {
  if (!w || !b1 || !b2 || !res || n > b1->n || n > b2->n) return fail("bad argument");
  w->total_pairs += n;
  for (uint32_t j = 0; j < n; ++j) write_pair(w, b1, b2, j, res[j]);
  return 0;
}

int walt_pe_writer_write(walt_pe_writer* w, const walt_batch* b1, const walt_batch* b2, const walt_cand* ranked1,
                         const uint32_t* n_ranked1, const walt_cand* ranked2, const uint32_t* n_ranked2,
                         const walt_pair* pairs, uint32_t n) {
  if (!w || !b1 || !b2 || n > b1->n || n > b2->n) return fail("bad argument");
  w->total_pairs += n;
  for (uint32_t j = 0; j < n; ++j) {
    // the summary the device computes in the compact path, derived here from the ranked lists
    const walt_cand* c1 = ranked1 + (size_t)j * w->top_k;
    const walt_cand* c2 = ranked2 + (size_t)j * w->top_k;
    walt_pe_result r;
    memset(&r, 0, sizeof(r));
    r.pair = pairs[j];
    if (pairs[j].best_times >= 1) { r.c1 = c1[pairs[j].best_i]; r.c2 = c2[pairs[j].best_j]; }
    r.single1 = best_for_single(c1, n_ranked1[j], w->m);
    r.single2 = best_for_single(c2, n_ranked2[j], w->m);
    write_pair(w, b1, b2, j, r);
  }
  return 0;
}

int walt_pe_writer_close(walt_pe_writer* w) {
  if (!w) return 0;
  w->out.close();
  w->so1.close(); w->so2.close();
  // StatPairedReads::tostring, paired.cpp:52-77
  std::ostringstream o;
  o << "pairs:" << std::endl
    << "    total_read_pairs: " << w->total_pairs << std::endl
    << "    mapped:" << std::endl
    << "        unique: " << w->unique_pairs << std::endl
    << "        percent_unique: " << (100.0 * (double)w->unique_pairs) / (double)w->total_pairs << std::endl
    << "        ambiguous: " << w->ambiguous_pairs << std::endl
    << "    unmapped: " << w->unmapped_pairs << std::endl
    << "mate1:" << std::endl << w->so1.st.text(1) << std::endl
    << "mate2:" << std::endl << w->so2.st.text(1) << std::endl;
  o << "frag_len_distribution:" << std::endl;
  double total = 0.0;
  for (size_t i = 0; i < w->frag_count.size(); ++i) {
    o << "    " << i << ": " << w->frag_count[i] << std::endl;
    total += (double)(i * w->frag_count[i]);
  }
  o << "frag_len_mean: " << total / std::accumulate(w->frag_count.begin(), w->frag_count.end(), 0.0);
  std::ofstream ms(w->path + ".mapstats", std::ios::app);
  ms << o.str() << std::endl;
  const bool ok = (bool)ms;
  delete w;
  return ok ? 0 : fail("cannot write mapstats");
}

// ---- makedb output -----------------------------------------------------------------------------
int walt_write_dbindex_header(const char* path, const walt_chroms* c, uint32_t size_of_index) {
  FILE* f = fopen(path, "wb");
  if (!f) return fail(std::string("cannot open ") + path);
  const uint32_t n = (uint32_t)c->names.size();
  fwrite(&n, 4, 1, f);
  for (uint32_t i = 0; i < n; ++i) {
    uint32_t ln = (uint32_t)std::min<size_t>(c->names[i].size(), 255);
    fwrite(&ln, 4, 1, f);
    fwrite(c->names[i].data(), 1, ln, f);
  }
  fwrite(c->lengths.data(), 4, n, f);
  fwrite(&c->genome_len, 4, 1, f);
  fwrite(&size_of_index, 4, 1, f);
  return fclose(f) == 0 ? 0 : fail("write failed");
}

int walt_write_subindex(const char* path, char strand, const char* sequence, uint64_t genome_len,
                        const uint32_t* counter, const uint32_t* index, uint32_t index_size) {
  FILE* f = fopen(path, "wb");
  if (!f) return fail(std::string("cannot open ") + path);
  const uint32_t counter_size = 1u << 24;
  bool ok = fwrite(&strand, 1, 1, f) == 1;
  ok = ok && fwrite(sequence, 1, genome_len, f) == genome_len;
  ok = ok && fwrite(&counter_size, 4, 1, f) == 1 && fwrite(&index_size, 4, 1, f) == 1;
  ok = ok && fwrite(counter, 4, (size_t)counter_size + 1, f) == (size_t)counter_size + 1;
  ok = ok && fwrite(index, 4, index_size, f) == index_size;
  ok = (fclose(f) == 0) && ok;
  return ok ? 0 : fail(std::string("short write: ") + path);
}

}  // extern "C"
