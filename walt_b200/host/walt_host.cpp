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
// computes alignments.  The N-replacement stream is glibc's: random_r on a private state set
// up by initstate_r(0, ..) yields exactly the srand(0); rand() sequence of mapping.cpp:73 and
// util.hpp:156-163 (rand() is random_r on glibc's global state), without sharing that global
// state between the loader threads of the two mate files.
//
// At GPU mapping rates the text path is the whole wall clock, so both
// directions are multi-threaded with an ordered commit: the loader splits the mapped file at
// line boundaries, the writers format blocks of reads on worker threads while the calling
// thread appends finished blocks to the output files in input order.
#include "../../include/walt_host.h"

#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <fstream>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_err;
int fail(const std::string& m) { g_err = m; return -1; }

std::atomic<unsigned> g_threads{0};
std::atomic<uint32_t> g_chunk_bytes{1u << 16};   // smallest piece of a file one loader task scans
std::atomic<uint32_t> g_block_reads{8192};       // reads one writer task formats
unsigned host_threads() {
  unsigned t = g_threads.load();
  if (!t) { t = std::thread::hardware_concurrency(); if (!t) t = 1; }
  return t;
}

// fn(i) for i in [0, n), handed out dynamically to up to host_threads() threads (the caller's included)
template <class F>
void parallel_tasks(size_t n, F&& fn) {
  const size_t T = std::min<size_t>(host_threads(), n);
  if (T <= 1) { for (size_t i = 0; i < n; ++i) fn(i); return; }
  std::atomic<size_t> next{0};
  auto body = [&]() { for (;;) { const size_t i = next.fetch_add(1); if (i >= n) break; fn(i); } };
  std::vector<std::thread> th;
  for (size_t t = 1; t < T; ++t) th.emplace_back(body);
  body();
  for (auto& t : th) t.join();
}

constexpr size_t LINE_CAP = 1000;      // MAX_LINE_LENGTH, util.hpp:43
constexpr uint32_t MIN_READ_LEN = 38;  // MINIMALREADLEN

inline char comp(char c) {
  switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; }
  return c;
}

// complement table (anything but ACGT maps to itself, like complimentBase never sees here)
struct CompTable {
  char t[256];
  CompTable() { for (int i = 0; i < 256; ++i) t[i] = (char)i; t['A'] = 'T'; t['C'] = 'G'; t['G'] = 'C'; t['T'] = 'A'; }
};
const CompTable g_comp;

// Text sink: a raw growable buffer in front of a file (or alone: an in-memory block that a
// worker thread fills and the committing thread appends to the file).  The file offset is kept
// here and every write is a pwrite at a known offset.
struct Sink {
  int fd = -1;
  uint64_t file_off = 0;
  char* buf = nullptr;
  size_t n = 0, cap = 0;
  std::atomic<bool> write_failed{false};
  Sink() = default;
  Sink(const Sink&) = delete;
  Sink& operator=(const Sink&) = delete;
  ~Sink() { if (fd >= 0) ::close(fd); free(buf); }
  bool is_file() const { return fd >= 0; }
  // mode "a": behind whatever the file holds (the outputs are created empty by main, walt.cpp:230-233,
  // and the SAM header / earlier batches are already in them); "w": truncate
  bool open(const std::string& path, const char* mode) {
    const bool append = mode[0] == 'a';
    fd = ::open(path.c_str(), O_WRONLY | O_CREAT | (append ? 0 : O_TRUNC), 0666);
    if (fd < 0) return false;
    const off_t end = append ? lseek(fd, 0, SEEK_END) : 0;
    file_off = end > 0 ? (uint64_t)end : 0;
    return true;
  }
  char* room(size_t k) {   // space for k more bytes
    if (n + k > cap) {
      size_t c = cap ? cap * 2 : (1u << 16);
      while (c < n + k) c *= 2;
      buf = (char*)realloc(buf, c);
      cap = c;
    }
    return buf + n;
  }
  static bool write_all(int fd, const char* p, size_t k, uint64_t off) {
    while (k) {
      const ssize_t w = pwrite(fd, p, k, (off_t)off);
      if (w < 0) { if (errno == EINTR) continue; return false; }
      p += w; k -= (size_t)w; off += (uint64_t)w;
    }
    return true;
  }
  void flush() {
    if (fd >= 0 && n) { if (!write_all(fd, buf, n, file_off)) write_failed = true; file_off += n; n = 0; }
  }
  void maybe_flush() { if (fd >= 0 && n > (1u << 20) - 4096) flush(); }   // no file: an in-memory block
  // A block formatted elsewhere goes behind everything written so far -- in two steps, so that only the
  // first one is ordered: reserve() (blocks in input order, one at a time) gives the block its place in
  // the file, write_reserved() (any thread, any order, concurrently) puts it there.  One writer thread
  // moves about 1 GB/s into the page cache (5 GB/s on the GPU boxes) whatever formats the blocks;
  // writes at known offsets from several threads add up on the file systems that let them.
  Sink* dest = nullptr;    // of a block: the file it was reserved in
  uint64_t at = 0;         // ... and where
  void reserve(Sink& block) {
    block.dest = nullptr;
    if (!block.n) return;
    if (fd < 0) { memcpy(room(block.n), block.buf, block.n); n += block.n; return; }
    flush();
    block.dest = this; block.at = file_off;
    file_off += block.n;
  }
  void write_reserved() {
    if (dest && n && !write_all(dest->fd, buf, n, at)) dest->write_failed = true;
    dest = nullptr;
  }
  void append_block(Sink& block) { reserve(block); block.write_reserved(); }
  bool close() {
    flush();
    if (fd >= 0 && ::close(fd) != 0) write_failed = true;
    fd = -1;
    return !write_failed;
  }
  void str(const char* s, size_t k) { if (k) { memcpy(room(k), s, k); n += k; } }
  template <size_t N> void lit(const char (&s)[N]) { memcpy(room(N - 1), s, N - 1); n += N - 1; }
  void str(const std::string& s) { str(s.data(), s.size()); }
  void ch(char c) { *room(1) = c; ++n; }
  void u32(uint32_t v) {
    char t[12]; int k = 0;
    do { t[k++] = (char)('0' + v % 10); v /= 10; } while (v);
    char* o = room((size_t)k);
    for (int i = 0; i < k; ++i) o[i] = t[k - 1 - i];
    n += (size_t)k;
  }
  void i32(int32_t v) { if (v < 0) { ch('-'); u32((uint32_t)(-(int64_t)v)); } else u32((uint32_t)v); }
  // reversed (and complemented) copies, 16 bytes at a time: half of all output lines are '-' strand
#if defined(__SSE2__)
  static __m128i reversed16(__m128i v) {
    v = _mm_or_si128(_mm_srli_epi16(v, 8), _mm_slli_epi16(v, 8));                  // bytes of every 16-bit lane
    v = _mm_shufflelo_epi16(v, _MM_SHUFFLE(0, 1, 2, 3));
    v = _mm_shufflehi_epi16(v, _MM_SHUFFLE(0, 1, 2, 3));
    return _mm_shuffle_epi32(v, _MM_SHUFFLE(1, 0, 3, 2));
  }
  static __m128i complemented16(__m128i v) {   // A <-> T is ^ 0x15, C <-> G is ^ 0x04, anything else stays
    const __m128i at = _mm_or_si128(_mm_cmpeq_epi8(v, _mm_set1_epi8('A')), _mm_cmpeq_epi8(v, _mm_set1_epi8('T')));
    const __m128i cg = _mm_or_si128(_mm_cmpeq_epi8(v, _mm_set1_epi8('C')), _mm_cmpeq_epi8(v, _mm_set1_epi8('G')));
    return _mm_xor_si128(v, _mm_or_si128(_mm_and_si128(at, _mm_set1_epi8(0x15)), _mm_and_si128(cg, _mm_set1_epi8(0x04))));
  }
#endif
  void revcomp(const char* s, size_t k) {
    char* o = room(k);
    size_t i = 0;
#if defined(__SSE2__)
    for (; i + 16 <= k; i += 16)
      _mm_storeu_si128(reinterpret_cast<__m128i*>(o + i),
                       complemented16(reversed16(_mm_loadu_si128(reinterpret_cast<const __m128i*>(s + k - 16 - i)))));
#endif
    for (; i < k; ++i) o[i] = g_comp.t[(unsigned char)s[k - 1 - i]];
    n += k;
  }
  void rev(const char* s, size_t k) {
    char* o = room(k);
    size_t i = 0;
#if defined(__SSE2__)
    for (; i + 16 <= k; i += 16)
      _mm_storeu_si128(reinterpret_cast<__m128i*>(o + i), reversed16(_mm_loadu_si128(reinterpret_cast<const __m128i*>(s + k - 16 - i))));
#endif
    for (; i < k; ++i) o[i] = s[k - 1 - i];
    n += k;
  }
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

struct walt_fastq {
  int fd = -1;
  const char* data = nullptr;   // the whole file: mmap for regular files, a heap copy otherwise
  size_t size = 0, pos = 0;
  bool mapped = false;
  std::vector<char> heap;
  double bytes_per_line = 90.0; // running estimate, sizes the window scanned for one batch
  struct random_data rd;        // the rand() stream of toACGT (srand(0) at the start of every batch, mapping.cpp:73)
  char rd_state[128];
  walt_fastq() { restart_rand(); }
  void restart_rand() {
    memset(&rd, 0, sizeof rd);
    initstate_r(0, rd_state, sizeof rd_state, &rd);
  }
};

// uninitialised growable byte buffer (std::string would zero-fill gigabytes on resize)
struct RawBuf {
  char* p = nullptr;
  size_t n = 0, cap = 0;
  RawBuf() = default;
  RawBuf(const RawBuf&) = delete;
  RawBuf& operator=(const RawBuf&) = delete;
  ~RawBuf() { free(p); }
  bool size_to(size_t need) {
    if (need > cap) {
      free(p);
      p = (char*)malloc(need + 64);
      cap = p ? need : 0;
      if (!p) { n = 0; return false; }
    }
    n = need;
    return true;
  }
};

// n + 1 offsets, uninitialised like RawBuf (a std::vector would zero 80 MB per array of a 10 M read batch on one thread)
struct OffBuf {
  RawBuf raw;
  OffBuf() { resize(1); (*this)[0] = 0; }
  bool resize(size_t n) { return raw.size_to(n * sizeof(uint64_t)); }
  uint64_t& operator[](size_t i) { return reinterpret_cast<uint64_t*>(raw.p)[i]; }
  const uint64_t& operator[](size_t i) const { return reinterpret_cast<const uint64_t*>(raw.p)[i]; }
  const uint64_t* data() const { return reinterpret_cast<const uint64_t*>(raw.p); }
};

struct walt_batch {
  RawBuf seqs, quals, names, packed;                   // names/quals hold NUL-terminated copies
  OffBuf seq_off, qual_off, name_off;                  // n + 1 each
  bool have_packed = false;                            // 2-bit form of seqs (walt_batch_packed): the loader leaves it behind
  uint32_t n = 0;
  void clear() {
    seqs.n = quals.n = names.n = packed.n = 0; have_packed = false; n = 0;
    seq_off.resize(1); qual_off.resize(1); name_off.resize(1);
    seq_off[0] = qual_off[0] = name_off[0] = 0;
  }
  const char* seq(uint32_t i) const { return seqs.p + seq_off[i]; }
  uint32_t seq_len(uint32_t i) const { return (uint32_t)(seq_off[i + 1] - seq_off[i]); }
  const char* qual(uint32_t i) const { return quals.p + qual_off[i]; }
  uint32_t qual_len(uint32_t i) const { return (uint32_t)(qual_off[i + 1] - qual_off[i] - 1); }
  const char* name(uint32_t i) const { return names.p + name_off[i]; }
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
  bool close() { const bool a = famb.close(), b = funm.close(); return a && b; }
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

namespace {

// Formats reads [0, n) in blocks on worker threads; the blocks get their places in the files in input
// order, and are written there by whichever thread got them their place.
//   make()            -> a fresh block-local output (in-memory sinks, zeroed counters)
//   reset(block)         ... the same from a used one (its buffers keep their pages)
//   fmt(block, lo, hi)   formats reads [lo, hi) into it                           (any thread)
//   reserve(block)       file offsets for its sinks, its counters into the totals (ascending order, one at a time)
//   write(block)         its sinks into the files at those offsets                (any thread)
template <class Block, class Make, class Reset, class Fmt, class Reserve, class Write>
void ordered_blocks(uint32_t n, Make make, Reset reset, Fmt fmt, Reserve reserve, Write write) {
  const uint32_t BLOCK = g_block_reads.load();
  const size_t nb = ((size_t)n + BLOCK - 1) / BLOCK;
  const size_t T = std::min<size_t>(host_threads(), nb);
  if (T <= 1) {
    std::unique_ptr<Block> bl;
    for (size_t k = 0; k < nb; ++k) {
      if (bl) reset(*bl); else bl.reset(make());
      fmt(*bl, (uint32_t)(k * BLOCK), (uint32_t)std::min<size_t>(n, (k + 1) * BLOCK));
      reserve(*bl);
      write(*bl);
    }
    return;
  }
  std::mutex mu;
  std::condition_variable cv_room;
  std::map<size_t, std::unique_ptr<Block>> ready;      // formatted, waiting for the blocks before them
  std::vector<std::unique_ptr<Block>> spare;           // written: their buffers are reused
  size_t next_commit = 0;
  std::atomic<size_t> next_task{0};
  const size_t window = 4 * T;   // blocks formatted ahead of the commit point (bounds memory)
  auto body = [&]() {
    std::vector<std::unique_ptr<Block>> mine;
    for (;;) {
      const size_t k = next_task.fetch_add(1);
      if (k >= nb) return;
      std::unique_ptr<Block> bl;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_room.wait(lk, [&]() { return k < next_commit + window; });
        if (!spare.empty()) { bl = std::move(spare.back()); spare.pop_back(); }
      }
      if (bl) reset(*bl); else bl.reset(make());
      fmt(*bl, (uint32_t)(k * BLOCK), (uint32_t)std::min<size_t>(n, (k + 1) * BLOCK));
      {
        std::lock_guard<std::mutex> lk(mu);
        ready[k] = std::move(bl);
        // the thread that completes the prefix gives every block of it its place ...
        for (auto it = ready.find(next_commit); it != ready.end(); it = ready.find(next_commit)) {
          reserve(*it->second);
          mine.push_back(std::move(it->second));
          ready.erase(it);
          ++next_commit;
        }
      }
      if (mine.empty()) continue;
      cv_room.notify_all();
      for (auto& m : mine) write(*m);                  // ... and writes them, beside the other threads' writes
      {
        std::lock_guard<std::mutex> lk(mu);
        for (auto& m : mine) spare.push_back(std::move(m));
      }
      mine.clear();
    }
  };
  std::vector<std::thread> th;
  for (size_t t = 1; t < T; ++t) th.emplace_back(body);
  body();
  for (auto& t : th) t.join();
}

void add_stats(SingleStats& a, const SingleStats& b) {
  a.total += b.total; a.unique += b.unique; a.ambiguous += b.ambiguous; a.unmapped += b.unmapped; a.n_short += b.n_short;
}
// block-local twin of a SingleOut: same flags, in-memory side files, zero counters
void local_twin(SingleOut& l, const SingleOut& of) { l.ambiguous = of.ambiguous; l.unmapped = of.unmapped; l.sam = of.sam; }
void reset_twin(SingleOut& l) { l.famb.n = l.funm.n = 0; l.st = SingleStats(); }
void reserve_twin(SingleOut& to, SingleOut& l) {
  to.famb.reserve(l.famb); to.funm.reserve(l.funm);
  add_stats(to.st, l.st);
}
void write_twin(SingleOut& l) { l.famb.write_reserved(); l.funm.write_reserved(); }

struct SeBlock { Sink out; SingleOut so; };

// what write_pair touches: the main output, the two mates' side files / counters, pair totals
struct PeOut {
  Sink out;
  SingleOut so1, so2;
  uint32_t unique_pairs = 0, ambiguous_pairs = 0, unmapped_pairs = 0;
  std::vector<uint32_t> frag_count;
};

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
  PeOut o;                   // the files and the running totals
  uint32_t total_pairs = 0;
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
  if (ok) {
    c->lengths.resize(n);
    ok = fread(c->lengths.data(), 4, n, f) == n;
  }
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
void walt_host_set_threads(unsigned n) { g_threads.store(n); }
unsigned walt_host_threads(void) { return host_threads(); }
void walt_host_set_grain(uint32_t chunk_bytes, uint32_t block_reads) {
  g_chunk_bytes.store(chunk_bytes ? chunk_bytes : (1u << 16));
  g_block_reads.store(block_reads ? block_reads : 8192u);
}

walt_fastq* walt_fastq_open(const char* path) {
  const int fd = open(path, O_RDONLY);
  struct stat st;
  if (fd < 0 || fstat(fd, &st) != 0 || S_ISDIR(st.st_mode)) {
    if (fd >= 0) close(fd);
    fail(std::string("cannot open input file ") + path);
    return nullptr;
  }
  walt_fastq* q = new walt_fastq;
  q->fd = fd;
  if (S_ISREG(st.st_mode) && st.st_size > 0) {
    void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m != MAP_FAILED) {
      madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
      q->data = (const char*)m; q->size = (size_t)st.st_size; q->mapped = true;
    }
  }
  if (!q->mapped) {   // pipes, empty files, mmap refused: read it all
    char tmp[1 << 16];
    for (;;) {
      const ssize_t r = read(fd, tmp, sizeof tmp);
      if (r < 0 && errno == EINTR) continue;
      if (r <= 0) break;
      q->heap.insert(q->heap.end(), tmp, tmp + r);
    }
    q->data = q->heap.data(); q->size = q->heap.size();
  }
  return q;
}
void walt_fastq_close(walt_fastq* f) {
  if (!f) return;
  if (f->mapped) munmap((void*)f->data, f->size);
  if (f->fd >= 0) close(f->fd);
  delete f;
}
walt_batch* walt_batch_create(void) { walt_batch* b = new walt_batch; b->clear(); return b; }
void walt_batch_free(walt_batch* b) { delete b; }

size_t walt_clip_adaptor(const char* adaptor, char* s, size_t len) {
  // util.hpp:200-217.  For reads shorter than 13 (5) bases the reference's size_t arithmetic
  // wraps and its loops run off the end of the string (undefined behaviour, in practice a
  // crash); here those loops are simply empty.
  const size_t ad_len = strlen(adaptor);
  if (len < 5) return 0;
  const size_t lim1 = len >= 13 ? len - 14 + 1 : 0;
  for (size_t i = 0; i < lim1; ++i)
    if (similarity(s, len, i, adaptor, ad_len) >= 11) { memset(s + i, 'N', len - i); return len - i; }
  const size_t lim2 = len - 5 + 1;
  for (size_t i = lim1; i < lim2; ++i)
    if (similarity(s, len, i, adaptor, ad_len) >= len - i - 1) { memset(s + i, 'N', len - i); return len - i; }
  return 0;
}

}  // extern "C"

namespace {

// One `fgets(cline, MAX_LINE_LENGTH, fin); cline[strlen(cline) - 1] = 0;` step of the reference
// loader (mapping.cpp:79-80) over memory: at most 999 bytes, ending behind a newline if one
// comes first; strlen stops at an embedded NUL; the last character goes, whatever it is.
struct Piece { const char* s = nullptr; uint32_t len = 0; const char* next = nullptr; };
inline bool next_piece(const char* p, const char* end, Piece& out) {
  if (p >= end) return false;
  const size_t room = std::min<size_t>((size_t)(end - p), LINE_CAP - 1);
  const char* nl = (const char*)memchr(p, '\n', room);
  const char* q = nl ? nl + 1 : p + room;
  size_t slen = (size_t)(q - p);
  const char* z = (const char*)memchr(p, 0, slen);
  if (z) slen = (size_t)(z - p);
  if (slen) --slen;
  out.s = p; out.len = (uint32_t)slen; out.next = q;
  return true;
}
// next piece that survives `if (line.size() == 0) continue;` (mapping.cpp:82-83)
inline bool next_line(const char*& p, const char* end, Piece& out) {
  while (next_piece(p, end, out)) {
    p = out.next;
    if (out.len) return true;
  }
  return false;
}

struct RecRef {            // the four lines of one record inside the file
  uint64_t name;           // offset of the name's first character
  uint32_t seq_delta, qual_delta;   // from `name`
  uint16_t name_len, seq_len, qual_len;
};

// Every surviving piece of [p, end) in file order, emit(start, len): what next_line() returns in a loop, found
// 16 bytes at a time.  The common piece -- a newline within 999 bytes, no NUL in front of it -- never
// leaves the vector loop; anything else goes through next_piece().  `end` is a physical line start or
// the end of the file; 16-byte loads stay in front of `safe_end` (the end of the file).
template <class Emit>
void scan_pieces(const char* p, const char* end, const char* safe_end, Emit&& emit) {
  const char* line = p;
  auto exact = [&]() {
    Piece pc;
    next_piece(line, end, pc);
    if (pc.len) emit(pc.s, pc.len);
    line = pc.next;
  };
#if defined(__SSE2__)
  const __m128i nl = _mm_set1_epi8('\n'), zero = _mm_setzero_si128();
  for (const char* b = p; b < end && b + 16 <= safe_end; b += 16) {
    const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(b));
    unsigned m = (unsigned)_mm_movemask_epi8(_mm_or_si128(_mm_cmpeq_epi8(v, nl), _mm_cmpeq_epi8(v, zero)));
    while (m) {
      const char* q = b + __builtin_ctz(m);
      m &= m - 1;
      if (q >= end) break;
      while (q >= line) {     // (q < line: inside a piece that took the exact path)
        if (*q == '\n' && (size_t)(q - line) < LINE_CAP - 1) {
          if (q != line) emit(line, (uint32_t)(q - line));
          line = q + 1;
        } else {
          exact();              // a NUL, or a piece that fgets cuts off: it may end in front of q
        }
      }
    }
    while (line < end && (size_t)(b + 16 - line) >= LINE_CAP - 1) exact();   // 999 bytes without a newline
  }
#else
  (void)safe_end;
#endif
  while (line < end) exact();
}

// 16 sequence bytes at a time: a bit for every byte that is not A, C, G or T
inline unsigned not_acgt16(const char* s) {
#if defined(__SSE2__)
  const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s));
  const __m128i ok = _mm_or_si128(_mm_or_si128(_mm_cmpeq_epi8(v, _mm_set1_epi8('A')), _mm_cmpeq_epi8(v, _mm_set1_epi8('C'))),
                                  _mm_or_si128(_mm_cmpeq_epi8(v, _mm_set1_epi8('G')), _mm_cmpeq_epi8(v, _mm_set1_epi8('T'))));
  return ~(unsigned)_mm_movemask_epi8(ok) & 0xFFFFu;
#else
  unsigned m = 0;
  for (unsigned t = 0; t < 16; ++t) { const char c = s[t]; if (!(c == 'A' || c == 'C' || c == 'G' || c == 'T')) m |= 1u << t; }
  return m;
#endif
}

struct Mark { uint64_t at, packed_at; uint32_t shift; };   // a character toACGT replaces: where in seqs, where in the 2-bit form

inline uint32_t codes_of4(uint32_t w) {       // see walt_pack_reads
  const uint32_t x = (w >> 1) & 0x03030303u;
  return x ^ ((x >> 1) & 0x01010101u);
}
// the 2-bit form of one read, whatever its bytes are (a byte that is not ACGT leaves two bits that its Mark overwrites)
inline void pack_one(const char* s, uint32_t len, uint8_t* o) {
  uint32_t i = 0;
  for (; i + 4 <= len; i += 4) {
    uint32_t w;
    memcpy(&w, s + i, 4);
    *o++ = (uint8_t)((codes_of4(w) * 0x40100401u) >> 24);
  }
  if (i < len) {
    unsigned v = 0;
    for (uint32_t t = 0; i + t < len; ++t) {
      const unsigned c = (unsigned char)s[i + t], x = (c >> 1) & 3u;
      v |= (x ^ (x >> 1)) << (6u - 2u * t);
    }
    *o = (uint8_t)v;
  }
}

}  // namespace

extern "C" {

int64_t walt_fastq_next_batch(walt_fastq* fq, walt_batch* b, uint32_t max_reads, const char* adaptor) {
  return walt_fastq_next_part(fq, b, max_reads, adaptor, 1);
}

int64_t walt_fastq_next_part(walt_fastq* fq, walt_batch* b, uint32_t max_reads, const char* adaptor, int restart_rand) {
  if (!fq || !b) return fail("bad argument");
  if (restart_rand) fq->restart_rand();
  b->clear();
  const uint64_t lim = (uint64_t)max_reads * 4u;   // lines, mapping.cpp:78
  const bool clip = adaptor && adaptor[0];
  if (lim == 0 || fq->pos >= fq->size) return 0;
  const char* const base = fq->data;
  const char* const file_end = base + fq->size;

  // ---- 1. a window [lo, hi) that holds the batch's lines, cut at physical line starts; every chunk lists its lines ----
  const char* lo = base + fq->pos;
  const char* hi = lo;
  std::vector<const char*> cut;       // chunk boundaries, cut.front() == lo, cut.back() == hi
  std::vector<uint64_t> first_line;   // index of the first surviving line of each chunk (+ total)
  std::vector<std::vector<uint64_t>> lines;   // per chunk: (start - lo) << 16 | length, for every surviving piece
  uint64_t total = 0;
  double want = (double)lim * fq->bytes_per_line * 1.1 + 65536.0;
  for (;;) {
    const char* h = (double)(file_end - lo) <= want ? file_end : lo + (size_t)want;
    if (h < file_end) {
      const char* nl = (const char*)memchr(h, '\n', (size_t)(file_end - h));
      h = nl ? nl + 1 : file_end;
    }
    hi = h;
    const size_t K = std::max<size_t>(1, std::min<size_t>((size_t)host_threads() * 4u, (size_t)(hi - lo) / g_chunk_bytes.load()));
    cut.assign(K + 1, hi);
    cut[0] = lo;
    for (size_t i = 1; i < K; ++i) {
      const char* c = lo + (size_t)(hi - lo) / K * i;
      c = std::max(c, cut[i - 1]);
      if (c > lo && c[-1] != '\n') {
        const char* nl = (const char*)memchr(c, '\n', (size_t)(hi - c));
        c = nl ? nl + 1 : hi;
      }
      cut[i] = c;
    }
    lines.assign(K, std::vector<uint64_t>());
    parallel_tasks(K, [&](size_t i) {
      std::vector<uint64_t>& out = lines[i];
      out.reserve((size_t)((double)(cut[i + 1] - cut[i]) / fq->bytes_per_line * 1.05) + 16u);
      scan_pieces(cut[i], cut[i + 1], file_end, [&](const char* st, uint32_t len) { out.push_back((uint64_t)(st - lo) << 16 | len); });
    });
    first_line.assign(K + 1, 0);
    for (size_t i = 0; i < K; ++i) first_line[i + 1] = first_line[i] + lines[i].size();
    total = first_line[K];
    if (total >= lim || hi == file_end) break;
    want *= 2.0;
  }
  const size_t K = cut.size() - 1;
  const uint64_t take = std::min(total, lim);
  const uint64_t n_rec = take / 4u;   // a trailing partial record is never committed (mapping.cpp:104-108)
  if (total) fq->bytes_per_line = std::max(16.0, (double)(hi - lo) / (double)total);

  // line g of the window (g < total), looked for from chunk c on
  auto line_at = [&](uint64_t g, size_t c) -> uint64_t {
    while (g >= first_line[c + 1]) ++c;
    return lines[c][(size_t)(g - first_line[c])];
  };
  // where the next batch starts: at the first line behind the ones taken (the empty pieces in between are
  // skipped either way), behind the window if it was used up
  fq->pos = take == total ? (size_t)(hi - base) : (size_t)(lo - base) + (size_t)(line_at(take, 0) >> 16);
  if (n_rec == 0) return 0;

  // ---- 2. record table: every chunk emits the records that START in it, and how much room they need ----
  std::vector<std::vector<RecRef>> recs(K);
  struct Room { uint64_t seq = 0, qual = 0, name = 0; };
  std::vector<Room> room(K + 1);
  parallel_tasks(K, [&](size_t i) {
    const uint64_t g0 = first_line[i], g1 = std::min<uint64_t>(first_line[i + 1], n_rec * 4u);
    uint64_t g = g0 + (4u - g0 % 4u) % 4u;   // first record boundary at or behind the chunk start
    if (g >= g1) return;
    std::vector<RecRef>& out = recs[i];
    out.reserve((size_t)((g1 - g + 3u) / 4u));
    Room rm;
    for (; g < g1; g += 4u) {
      const uint64_t l0 = lines[i][(size_t)(g - g0)];
      const bool inside = g + 3u < first_line[i + 1];   // (the last record of a chunk may end in the next ones)
      const uint64_t l1 = inside ? lines[i][(size_t)(g + 1 - g0)] : line_at(g + 1, i);
      const uint64_t l3 = inside ? lines[i][(size_t)(g + 3 - g0)] : line_at(g + 3, i);
      const char* s0 = lo + (l0 >> 16);
      const uint32_t len0 = (uint32_t)(l0 & 0xFFFFu);
      // mapping.cpp:86-92: substr(1) or substr(1, space_pos - 1); a leading space makes the count npos
      const char* sp = (const char*)memchr(s0, ' ', len0);
      const uint32_t name_end = (sp && sp != s0) ? (uint32_t)(sp - s0) : len0;
      RecRef r;
      r.name = (uint64_t)(s0 + 1 - base);
      r.name_len = (uint16_t)(name_end - 1u);
      r.seq_delta = (uint32_t)((l1 >> 16) - (l0 >> 16) - 1u); r.seq_len = (uint16_t)(l1 & 0xFFFFu);
      r.qual_delta = (uint32_t)((l3 >> 16) - (l0 >> 16) - 1u); r.qual_len = (uint16_t)(l3 & 0xFFFFu);
      out.push_back(r);
      rm.seq += r.seq_len; rm.qual += (uint64_t)r.qual_len + 1u; rm.name += (uint64_t)r.name_len + 1u;
    }
    room[i] = rm;
  });
  lines.clear(); lines.shrink_to_fit();

  // ---- 3. where every chunk's records go (a prefix sum over the chunks), then offsets, copies, the 2-bit form
  //         and the characters toACGT will replace, chunk by chunk in parallel ----
  b->n = (uint32_t)n_rec;
  std::vector<size_t> rec_base(K + 1, 0);
  {
    Room at;
    for (size_t i = 0; i < K; ++i) {
      const Room mine = room[i];
      room[i] = at;
      at.seq += mine.seq; at.qual += mine.qual; at.name += mine.name;
      rec_base[i + 1] = rec_base[i] + recs[i].size();
    }
    room[K] = at;
    if (rec_base[K] != n_rec) return fail("internal error: FASTQ record table out of step");
    if (!b->seq_off.resize(n_rec + 1) || !b->qual_off.resize(n_rec + 1) || !b->name_off.resize(n_rec + 1) ||
        !b->seqs.size_to(at.seq) || !b->quals.size_to(at.qual) || !b->names.size_to(at.name) ||
        !b->packed.size_to((at.seq >> 2) + n_rec + 16u))
      return fail("ERROR: could not allocate memory");
    b->seq_off[n_rec] = at.seq; b->qual_off[n_rec] = at.qual; b->name_off[n_rec] = at.name;
  }
  std::vector<std::vector<Mark>> marks(K);
  parallel_tasks(K, [&](size_t i) {
    size_t j = rec_base[i];
    Room at = room[i];
    std::vector<Mark>& mk = marks[i];
    uint8_t* const packed = reinterpret_cast<uint8_t*>(b->packed.p);
    for (const RecRef& r : recs[i]) {
      b->seq_off[j] = at.seq; b->qual_off[j] = at.qual; b->name_off[j] = at.name;
      const char* nm = base + r.name;
      char* dn = b->names.p + at.name;
      memcpy(dn, nm, r.name_len); dn[r.name_len] = 0;
      char* dq = b->quals.p + at.qual;
      memcpy(dq, nm + r.qual_delta, r.qual_len); dq[r.qual_len] = 0;
      char* ds = b->seqs.p + at.seq;
      const uint32_t len = r.seq_len;
      memcpy(ds, nm + r.seq_delta, len);
      if (clip) walt_clip_adaptor(adaptor, ds, len);
      // read j of the 2-bit form: ceil(len / 4) bytes from byte (offset >> 2) + j, zeros up to the next read's
      const uint64_t pk = (at.seq >> 2) + j, pk_next = ((at.seq + len) >> 2) + j + 1u;
      pack_one(ds, len, packed + pk);
      for (uint64_t z = pk + (len + 3u) / 4u; z < pk_next; ++z) packed[z] = 0;
      uint32_t t = 0;
      for (; t + 16u <= len; t += 16u)
        for (unsigned m = not_acgt16(ds + t); m; m &= m - 1) {
          const uint32_t u = t + (uint32_t)__builtin_ctz(m);
          mk.push_back(Mark{at.seq + u, pk + (u >> 2), 6u - 2u * (u & 3u)});
        }
      for (; t < len; ++t) {
        const char c = ds[t];
        if (!(c == 'A' || c == 'C' || c == 'G' || c == 'T')) mk.push_back(Mark{at.seq + t, pk + (t >> 2), 6u - 2u * (t & 3u)});
      }
      at.seq += len; at.qual += (uint64_t)r.qual_len + 1u; at.name += (uint64_t)r.name_len + 1u;
      ++j;
    }
  });
  memset(b->packed.p + b->packed.n - 16, 0, 16);

  // ---- 4. toACGT (util.hpp:156-163) in file order on the batch's rand() stream (mapping.cpp:73) ----
  for (size_t i = 0; i < K; ++i)
    for (const Mark& mk : marks[i]) {
      int32_t r = 0;
      random_r(&fq->rd, &r);
      const unsigned code = (unsigned)(r % 4);               // codes A0 C1 G2 T3 are the indices of "ACGT"
      b->seqs.p[mk.at] = "ACGT"[code];
      uint8_t& byte = reinterpret_cast<uint8_t*>(b->packed.p)[mk.packed_at];
      byte = (uint8_t)((byte & ~(3u << mk.shift)) | (code << mk.shift));
    }
  b->have_packed = true;
  return b->n;
}

uint64_t walt_packed_reads_bytes(const uint64_t* offs, uint32_t n) { return (offs[n] >> 2) + n + 16u; }

// Codes A0 C1 G2 T3 (util.hpp:107-121) are bits 1..2 of the ASCII letter with the upper bit folded
// into the lower: x = (c >> 1) & 3, code = x ^ (x >> 1).  Four letters at a time: the four codes sit
// in the low bits of the four bytes of a word; one multiply gathers them into a byte, first letter
// in the top bits.  Validation is the same arithmetic run backwards ("ACGT"[code] == letter).
static inline uint32_t letters_of4(uint32_t code) {   // "ACGT"[code] per byte: 0x41 + {0, 2, 6, 0x13}
  const uint32_t lo = code & 0x01010101u, hi = (code >> 1) & 0x01010101u;
  return 0x41414141u + lo * 2u + hi * 6u + (lo & hi) * 0x0Bu;
}

int walt_pack_reads(const char* seqs, const uint64_t* offs, uint32_t n, uint8_t* out) {
  if (!offs || !out || (n && !seqs)) return fail("bad argument");
  uint32_t bad = 0;
  for (uint32_t j = 0; j < n; ++j) {
    const char* s = seqs + offs[j];
    const uint64_t len = offs[j + 1] - offs[j];
    uint8_t* o = out + (offs[j] >> 2) + j;
    uint64_t i = 0;
    for (; i + 4 <= len; i += 4) {
      uint32_t w;
      memcpy(&w, s + i, 4);
      const uint32_t code = codes_of4(w);
      bad |= letters_of4(code) ^ w;
      *o++ = (uint8_t)((code * 0x40100401u) >> 24);
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
  if (!b->have_packed) {
    const uint64_t* offs = b->seq_off.data();
    if (!b->packed.size_to(walt_packed_reads_bytes(offs, b->n))) { fail("ERROR: could not allocate memory"); return nullptr; }
    const uint32_t BLOCK = 1u << 16;
    const size_t nb = ((size_t)b->n + BLOCK - 1) / BLOCK;
    std::atomic<int> bad{0};
    // read j lives at byte (offs[j] >> 2) + j whatever range it is packed with: shift the output
    // pointer by the block's first read so that a sub-range call writes the same bytes
    parallel_tasks(nb, [&](size_t k) {
      const uint32_t lo = (uint32_t)k * BLOCK, cnt = std::min<uint32_t>(BLOCK, b->n - lo);
      const uint64_t b0 = (offs[lo] >> 2) + lo, b1 = (offs[lo + cnt] >> 2) + lo + cnt;
      memset(b->packed.p + b0, 0, (size_t)(b1 - b0));   // bytes between two reads stay zero
      if (walt_pack_reads(b->seqs.p, offs + lo, cnt, (uint8_t*)b->packed.p + lo) != 0) bad = 1;
    });
    memset(b->packed.p + b->packed.n - 16, 0, 16);
    if (bad) { fail("walt_pack_reads: a read holds a byte that is not A/C/G/T"); return nullptr; }
    b->have_packed = true;
  }
  return (const uint8_t*)b->packed.p;
}

uint32_t walt_batch_size(const walt_batch* b) { return b->n; }
const char* walt_batch_seqs(const walt_batch* b) { return b->seqs.p; }
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
  ordered_blocks<SeBlock>(
      n,
      [&]() { SeBlock* bl = new SeBlock; local_twin(bl->so, w->so); return bl; },
      [&](SeBlock& bl) { bl.out.n = 0; reset_twin(bl.so); },
      [&](SeBlock& bl, uint32_t lo, uint32_t hi) {
        for (uint32_t j = lo; j < hi; ++j) {
          bl.so.st.update(res[j].times);
          if (w->sam) sam_single(bl.out, bl.so, *w->g, res[j], b->name(j), b->name_len(j), b->seq(j), b->qual(j), b->seq_len(j), b->qual_len(j));
          else mr_single(bl.out, bl.so, *w->g, res[j], b->name(j), b->name_len(j), b->seq(j), b->qual(j), b->seq_len(j), b->qual_len(j), w->ag);
        }
      },
      [&](SeBlock& bl) { w->out.reserve(bl.out); reserve_twin(w->so, bl.so); },
      [&](SeBlock& bl) { bl.out.write_reserved(); write_twin(bl.so); });
  return 0;
}

void walt_se_writer_add_short(walt_se_writer* w, uint32_t n_short) { w->so.st.n_short += n_short; }

int walt_se_writer_close(walt_se_writer* w) {
  if (!w) return 0;
  bool ok = w->out.close();
  ok = w->so.close() && ok;
  std::ofstream ms(w->path + ".mapstats", std::ios::app);
  ms << w->so.st.text(0) << std::endl;
  ok = ok && (bool)ms;
  delete w;
  return ok ? 0 : fail("cannot write the outputs / mapstats");
}

// ---- paired end --------------------------------------------------------------------------------
walt_pe_writer* walt_pe_writer_open(const char* output_path, const walt_chroms* chroms, uint32_t max_mismatches,
                                    uint32_t top_k, int frag_range, int ambiguous, int unmapped, int sam, int pbat) {
  walt_pe_writer* w = new walt_pe_writer;
  w->g = chroms; w->path = output_path; w->m = max_mismatches; w->top_k = top_k; w->frag_range = frag_range;
  w->sam = sam != 0; w->pbat = pbat != 0;
  w->o.frag_count.assign((size_t)frag_range + 1, 0);
  if (!w->o.out.open(w->path, "a")) { fail("cannot open input file " + w->path); delete w; return nullptr; }
  if (!w->o.so1.open(w->path + "_1", ambiguous != 0, unmapped != 0, w->sam) ||
      !w->o.so2.open(w->path + "_2", ambiguous != 0, unmapped != 0, w->sam)) {
    fail("cannot open the _1/_2 side files of " + w->path); w->o.out.close(); delete w; return nullptr;
  }
  if (w->sam) sam_header(w->o.out, *chroms);
  return w;
}

void walt_pe_writer_add_short(walt_pe_writer* w, uint32_t s1, uint32_t s2) { w->o.so1.st.n_short += s1; w->o.so2.st.n_short += s2; }

// One pair through MergePairedEndResults' output half (paired.cpp:515-569).  `r` holds the
// per-file-mate fields (c1/single1 belong to the first file's read).
static void write_pair(const walt_pe_writer* w, PeOut& o, const walt_batch* b1, const walt_batch* b2, uint32_t j, const walt_pe_result& r) {
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
    o.unique_pairs++;
    len = best_pair_out(o.out, g, pa, pb, w->frag_range, name, name_len, ba->seq(j), ba->qual(j), la,
                        bb->seq(j), bb->qual(j), lb, w->sam);
    if (len >= 0 && (size_t)len < o.frag_count.size()) o.frag_count[len]++;
    if (w->sam) {
      paired = true;
      ma.genome_pos = pa.genome_pos; ma.times = 1; ma.strand = pa.strand; ma.mismatch = pa.mismatch;
      mb.genome_pos = pb.genome_pos; mb.times = 1; mb.strand = pb.strand; mb.mismatch = pb.mismatch;
    }
  } else {
    if (r.pair.best_times >= 2) o.ambiguous_pairs++; else o.unmapped_pairs++;
    ma = w->pbat ? r.single2 : r.single1;
    mb = w->pbat ? r.single1 : r.single2;
    SingleOut& soa = w->pbat ? o.so2 : o.so1;
    SingleOut& sob = w->pbat ? o.so1 : o.so2;
    soa.st.update(ma.times);
    sob.st.update(mb.times);
    if (!w->sam) {
      mr_single(o.out, soa, g, ma, name, name_len, ba->seq(j), ba->qual(j), la, ba->qual_len(j), false);
      mr_single(o.out, sob, g, mb, name, name_len, bb->seq(j), bb->qual(j), lb, bb->qual_len(j), true);
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
    const SingleOut& soa = w->pbat ? o.so2 : o.so1;
    const SingleOut& sob = w->pbat ? o.so1 : o.so2;
    auto line_a = [&]() { sam_mate_line(o.out, soa, g, ma, chra, name, name_len, fa, sa, rnb, sb, tla, ba->seq(j), ba->qual(j), la, ba->qual_len(j), mma); };
    auto line_b = [&]() { sam_mate_line(o.out, sob, g, mb, chrb, name, name_len, fb, sb, rna, sa, tlb, bb->seq(j), bb->qual(j), lb, bb->qual_len(j), mmb); };
    if (a_first) { line_a(); line_b(); } else { line_b(); line_a(); }
  }
}

// block-local PeOut and its commit into the writer's files and totals
static PeOut* pe_block(const walt_pe_writer* w) {
  PeOut* bl = new PeOut;
  local_twin(bl->so1, w->o.so1); local_twin(bl->so2, w->o.so2);
  bl->frag_count.assign(w->o.frag_count.size(), 0);
  return bl;
}
static void pe_reset(PeOut& bl) {
  bl.out.n = 0; reset_twin(bl.so1); reset_twin(bl.so2);
  bl.unique_pairs = bl.ambiguous_pairs = bl.unmapped_pairs = 0;
  std::fill(bl.frag_count.begin(), bl.frag_count.end(), 0u);
}
static void pe_reserve(walt_pe_writer* w, PeOut& bl) {
  w->o.out.reserve(bl.out);
  reserve_twin(w->o.so1, bl.so1); reserve_twin(w->o.so2, bl.so2);
  w->o.unique_pairs += bl.unique_pairs; w->o.ambiguous_pairs += bl.ambiguous_pairs; w->o.unmapped_pairs += bl.unmapped_pairs;
  for (size_t i = 0; i < bl.frag_count.size(); ++i) w->o.frag_count[i] += bl.frag_count[i];
}
static void pe_write(PeOut& bl) { bl.out.write_reserved(); write_twin(bl.so1); write_twin(bl.so2); }

int walt_pe_writer_write_compact(walt_pe_writer* w, const walt_batch* b1, const walt_batch* b2, const walt_pe_result* res,
                                 uint32_t n) {
  if (!w || !b1 || !b2 || !res || n > b1->n || n > b2->n) return fail("bad argument");
  w->total_pairs += n;
  ordered_blocks<PeOut>(
      n, [&]() { return pe_block(w); },
      pe_reset,
      [&](PeOut& bl, uint32_t lo, uint32_t hi) { for (uint32_t j = lo; j < hi; ++j) write_pair(w, bl, b1, b2, j, res[j]); },
      [&](PeOut& bl) { pe_reserve(w, bl); }, pe_write);
  return 0;
}

int walt_pe_writer_write(walt_pe_writer* w, const walt_batch* b1, const walt_batch* b2, const walt_cand* ranked1,
                         const uint32_t* n_ranked1, const walt_cand* ranked2, const uint32_t* n_ranked2,
                         const walt_pair* pairs, uint32_t n) {
  if (!w || !b1 || !b2 || n > b1->n || n > b2->n) return fail("bad argument");
  w->total_pairs += n;
  ordered_blocks<PeOut>(
      n, [&]() { return pe_block(w); }, pe_reset,
      [&](PeOut& bl, uint32_t lo, uint32_t hi) {
        for (uint32_t j = lo; j < hi; ++j) {
          // the summary the device computes in the compact path, derived here from the ranked lists
          const walt_cand* c1 = ranked1 + (size_t)j * w->top_k;
          const walt_cand* c2 = ranked2 + (size_t)j * w->top_k;
          walt_pe_result r;
          memset(&r, 0, sizeof(r));
          r.pair = pairs[j];
          if (pairs[j].best_times >= 1) { r.c1 = c1[pairs[j].best_i]; r.c2 = c2[pairs[j].best_j]; }
          r.single1 = best_for_single(c1, n_ranked1[j], w->m);
          r.single2 = best_for_single(c2, n_ranked2[j], w->m);
          write_pair(w, bl, b1, b2, j, r);
        }
      },
      [&](PeOut& bl) { pe_reserve(w, bl); }, pe_write);
  return 0;
}

int walt_pe_writer_close(walt_pe_writer* w) {
  if (!w) return 0;
  bool files_ok = w->o.out.close();
  files_ok = w->o.so1.close() && files_ok;
  files_ok = w->o.so2.close() && files_ok;
  // StatPairedReads::tostring, paired.cpp:52-77
  std::ostringstream o;
  o << "pairs:" << std::endl
    << "    total_read_pairs: " << w->total_pairs << std::endl
    << "    mapped:" << std::endl
    << "        unique: " << w->o.unique_pairs << std::endl
    << "        percent_unique: " << (100.0 * (double)w->o.unique_pairs) / (double)w->total_pairs << std::endl
    << "        ambiguous: " << w->o.ambiguous_pairs << std::endl
    << "    unmapped: " << w->o.unmapped_pairs << std::endl
    << "mate1:" << std::endl << w->o.so1.st.text(1) << std::endl
    << "mate2:" << std::endl << w->o.so2.st.text(1) << std::endl;
  o << "frag_len_distribution:" << std::endl;
  double total = 0.0;
  for (size_t i = 0; i < w->o.frag_count.size(); ++i) {
    o << "    " << i << ": " << w->o.frag_count[i] << std::endl;
    total += (double)(i * w->o.frag_count[i]);
  }
  o << "frag_len_mean: " << total / std::accumulate(w->o.frag_count.begin(), w->o.frag_count.end(), 0.0);
  std::ofstream ms(w->path + ".mapstats", std::ios::app);
  ms << o.str() << std::endl;
  const bool ok = files_ok && (bool)ms;
  delete w;
  return ok ? 0 : fail("cannot write the outputs / mapstats");
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
