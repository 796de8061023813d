// walt_core.cuh -- lane-level mapping logic of the B200 engine.
//
// Everything here is written against a tiny "warp policy" W (lane id, ballot, shfl, sync) so
// that the very same source is (a) the body of the sm_100a kernels in walt_engine.cu, where W
// maps 1:1 onto the hardware warp intrinsics, and (b) steppable on a CPU by the test-only
// fiber harness in tests/emu/ (32 cooperative fibers per emulated warp), which lets the
// "-m 'not gpu'" suite exercise the kernel logic against the oracle without a device.  The
// product library (libwaltb200.so) contains only instantiation (a).
//
// Data model (one strand sub-index, all in HBM):
//   genome : 2 bits/base (A0 C1 G2 T3), 32 bases per 64-bit word, first base in the top bits,
//            so that masked word compares are lexicographic; PAD_BASES zero bases in front
//            and >= TAIL_PAD_WORDS zero words behind.
//   entries: the .dbindex position array, verbatim order (reference.cpp:302-322), each position
//            paired with a 32-bit fingerprint = its seed characters depth..depth+19 as a base-3
//            number (past-the-chromosome-end -> digit 0), so a table range is narrowed to the
//            read's 40-character seed prefix by integer compares, without touching the genome.
//   table  : T[k] = first slot of `index` whose first `depth` seed characters, read as a
//            base-3 number (3-letter alphabet), are >= k.  depth >= 12, so T restricted to
//            12-character prefixes IS the reference's counter[] (reference.cpp:192-229).
//   taint  : positions within 148 bases of a chromosome end.  The reference's per-character
//            binary search (mapping.cpp:166-222) reads past chromosome ends while the bucket
//            sort truncated there (reference.cpp:258-288), so for a lookup whose read matches
//            such an entry's in-chromosome prefix the search is replayed literally.
//
// Reference semantics reproduced here: SingleEndMapping (mapping.cpp:224-316), IndexRegion
// (mapping.cpp:198-222), getChromID (reference.cpp:43-60), the seed tables including their
// two reachable typos (seedpattern.hpp:448-455), PairEndMapping (paired.cpp:106-201) and
// TopCandidates with libstdc++ heap mechanics (paired.hpp:51-74).
#pragma once

#include <stdint.h>

#include <vector>

#if defined(__CUDACC__)
#define WALT_HD __host__ __device__ __forceinline__
#define WALT_HD_NOINLINE inline __host__ __device__ __noinline__
#define WALT_UNROLL _Pragma("unroll")
#define WALT_NO_UNROLL _Pragma("unroll 1")
#else
#define WALT_HD inline
#define WALT_HD_NOINLINE inline
#define WALT_UNROLL
#define WALT_NO_UNROLL
#endif
// pull the line that holds *p towards the SM ahead of a dependent load (no-op on the CPU harness)
#if defined(__CUDA_ARCH__)
#define WALT_PREFETCH(p) asm volatile("prefetch.global.L1 [%0];" ::"l"(p))
#else
#define WALT_PREFETCH(p) ((void)(p))
#endif

namespace waltcore {

constexpr uint32_t KEY_WEIGHT = 12;      // F2SEEDKEYWEIGHT
constexpr uint32_t MIN_READ_LEN = 38;    // MINIMALREADLEN
constexpr uint32_t MIN_SEED_LEN = 36;    // MINIMALSEEDLEN
constexpr uint32_t MAX_REPEATS = 50;     // mapping.cpp:238
constexpr uint32_t PAD_BASES = 32;       // zero bases in front of the packed genome
constexpr uint32_t MAX_READ_LEN = 1024;  // fgets(.., 1000, ..) bounds reads below this
constexpr uint32_t MAX_WORDS = MAX_READ_LEN / 32;
constexpr uint32_t TAIL_PAD_WORDS = MAX_WORDS + 8;
constexpr uint32_t TAINT_SPAN = 148;     // largest probed offset: 3*49+1 (seed_len 50)
constexpr uint32_t MAX_DEPTH = 20;
constexpr uint32_t FP_DIGITS = 20;       // seed characters depth .. depth+19 kept beside each position
constexpr uint32_t N_KEY12 = 531441;     // 3^12 reachable 12-mers of a 3-letter genome
constexpr uint32_t BUCKET_ERASE = 500000;  // reference.cpp:212

// ------------------------------------------------------------------------------------------
// portable bit helpers
// ------------------------------------------------------------------------------------------
WALT_HD uint32_t popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__popcll(x);
#else
  return (uint32_t)__builtin_popcountll(x);
#endif
}
WALT_HD uint32_t popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__popc(x);
#else
  return (uint32_t)__builtin_popcount(x);
#endif
}
WALT_HD uint32_t brev32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __brev(x);
#else
  x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
  x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
  x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
  x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
  return (x >> 16) | (x << 16);
#endif
}
WALT_HD int ffs32(uint32_t x) {  // 1-based index of lowest set bit, 0 if none
#if defined(__CUDA_ARCH__)
  return __ffs((int)x);
#else
  return __builtin_ffs((int)x);
#endif
}
WALT_HD int clz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __clz((int)x);
#else
  return x ? __builtin_clz(x) : 32;
#endif
}
// bit j of x -> bit 2j of the result
WALT_HD uint64_t spread32(uint32_t v) {
  uint64_t x = v;
  x = (x | (x << 16)) & 0x0000FFFF0000FFFFull;
  x = (x | (x << 8)) & 0x00FF00FF00FF00FFull;
  x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0Full;
  x = (x | (x << 2)) & 0x3333333333333333ull;
  x = (x | (x << 1)) & 0x5555555555555555ull;
  return x;
}

// ------------------------------------------------------------------------------------------
// alphabet
// ------------------------------------------------------------------------------------------
// ASCII -> 2-bit code A0 C1 G2 T3 (util.hpp:107-121); valid only for A,C,G,T
WALT_HD uint32_t ascii_code(uint32_t c) {
  uint32_t x = (c >> 1) & 3u;
  return x ^ (x >> 1);
}
WALT_HD bool ascii_is_acgt(uint32_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }
// read conversion (mapping.cpp:142-164): C->T, or G->A under the A/G wildcard
WALT_HD uint32_t convert_code(uint32_t code, bool ag) {
  if (ag) return code == 2u ? 0u : code;
  return code == 1u ? 3u : code;
}
// rank of a converted code inside its 3-letter alphabet: CT {A,G,T} / GA {A,C,T} -> 0,1,2
WALT_HD uint32_t ternary_digit(uint32_t code, bool ag) {
  return ag ? code - (code >> 1) : code - (code != 0u);
}
// base at padded position p (p already includes PAD_BASES)
WALT_HD uint32_t packed_base(const uint64_t* __restrict__ words, uint64_t p) {
  return (uint32_t)(words[p >> 5] >> (62u - 2u * (uint32_t)(p & 31u))) & 3u;
}

// ------------------------------------------------------------------------------------------
// seed geometry
// ------------------------------------------------------------------------------------------
WALT_HD uint32_t seed_repeats(uint32_t read_len) {  // mapping.cpp:236-239
  uint32_t r = (read_len - 2u) / 3u;
  return r < MAX_REPEATS ? r : MAX_REPEATS;
}
// is read position p one of the seed's cared positions for shift s (p = s + 3i + 1, i < spr)
WALT_HD bool is_seed_position(uint32_t p, uint32_t s, uint32_t spr) {
  if (p < s + 1u) return false;
  uint32_t d = p - s - 1u;
  return (d % 3u == 0u) && (d / 3u < spr);
}
// is read position p compared by the verification loops (mapping.cpp:288-304).  Regular rule:
// every position that is not a cared one; F2NOCAREDPOSITION[2] deviates at entries 47 and 95
// (seedpattern.hpp:448-455): 60 instead of 70 (reached when 2*spr+2 > 47) and 141 instead of
// 142 (reached when 2*spr+2 > 95).
WALT_HD bool is_verify_position(uint32_t p, uint32_t s, uint32_t spr, uint32_t read_len) {
  if (p >= read_len) return false;
  bool v = !is_seed_position(p, s, spr);
  if (s == 2u) {
    if (spr >= 23u) { if (p == 70u) v = false; if (p == 60u) v = true; }
    if (spr >= 47u) { if (p == 142u) v = false; if (p == 141u) v = true; }
  }
  return v;
}

// ------------------------------------------------------------------------------------------
// views
// ------------------------------------------------------------------------------------------
struct alignas(8) Entry { uint32_t pos; uint32_t fp; };

struct SubIndexView {
  const uint64_t* genome;      // packed, PAD_BASES in front
  const Entry* entries;        // index_size (+ readable pad)
  const uint32_t* table;       // 3^depth + 1 entries
  const uint32_t* taint_bits;  // N_KEY12 bits
  const uint32_t* taint_rank;  // set bits in front of each 32-bit word of taint_bits
  const uint32_t* taint_start; // per distinct tainted key (rank order): first slot in taint_pos/len, + end
  const uint32_t* taint_pos;   // their genome positions
  const uint32_t* taint_len;   // chromosome end - position
  const uint32_t* taint_slot;  // their slots in `entries` (0xFFFFFFFF: not an index entry)
  uint32_t n_taint;
  uint32_t index_size;
  uint32_t depth;
  uint32_t ag;                 // 1 for the G->A sub-indexes
};

struct ChromView {
  const uint32_t* starts;      // n_chr + 1 prefix sums (reference.cpp:407-410)
  uint32_t n_chr;
  uint32_t genome_len;
};

struct Pow3 { uint32_t v[MAX_DEPTH + 1]; };

// Host helper (engine and test harness): filter bits, rank directory and per-key slot ranges of
// the taint list from its sorted 12-mer keys.
inline void build_taint_directory(const std::vector<uint32_t>& sorted_keys, std::vector<uint32_t>& bits,
                                  std::vector<uint32_t>& rank, std::vector<uint32_t>& start) {
  const uint32_t words = (N_KEY12 + 31u) / 32u;
  bits.assign(words, 0u);
  rank.assign(words, 0u);
  start.clear();
  for (uint32_t i = 0; i < sorted_keys.size(); ++i) {
    const uint32_t k = sorted_keys[i];
    if (i == 0 || k != sorted_keys[i - 1]) start.push_back(i);
    bits[k >> 5] |= 1u << (k & 31u);
  }
  start.push_back((uint32_t)sorted_keys.size());
  uint32_t acc = 0;
  for (uint32_t w = 0; w < words; ++w) { rank[w] = acc; acc += popc32(bits[w]); }
}

// getChromID, reference.cpp:43-60
WALT_HD uint32_t chrom_of(const uint32_t* __restrict__ starts, uint32_t n_chr, uint32_t pos) {
  uint32_t l = 0, h = n_chr;
  while (l < h) {
    uint32_t m = (l + h + 1u) >> 1;
    if (pos >= starts[m]) l = m; else h = m - 1u;
  }
  return l;
}

// Key of index entry `e` over its first `depth` seed characters as a base-3 number, with the
// builder's truncation rule (characters at or beyond the chromosome end sort lowest,
// reference.cpp:272-277) mapped to digit 0.  Monotone non-decreasing along `index`.
WALT_HD uint32_t entry_table_key(const uint64_t* __restrict__ genome, uint32_t e, uint32_t chrom_end,
                                 uint32_t depth, bool ag, const Pow3& p3) {
  uint32_t avail = chrom_end - e;  // characters e .. chrom_end-1 exist
  uint32_t key = 0;
  for (uint32_t i = 0; i < depth; ++i) {
    uint32_t off = 3u * i + 1u;
    uint32_t d = 0;
    if (off < avail) d = ternary_digit(packed_base(genome, (uint64_t)e + PAD_BASES + off), ag);
    key += d * p3.v[depth - 1u - i];
  }
  return key;
}

// Fingerprint of index entry `e`: seed characters depth .. depth+FP_DIGITS-1 as a base-3 number,
// same truncation rule.  Monotone non-decreasing inside a table range.
WALT_HD uint32_t entry_fingerprint(const uint64_t* __restrict__ genome, uint32_t e, uint32_t chrom_end,
                                   uint32_t depth, bool ag, const Pow3& p3) {
  uint32_t avail = chrom_end - e;
  uint32_t fp = 0;
  for (uint32_t i = 0; i < FP_DIGITS; ++i) {
    uint32_t off = 3u * (depth + i) + 1u;
    uint32_t d = 0;
    if (off < avail) d = ternary_digit(packed_base(genome, (uint64_t)e + PAD_BASES + off), ag);
    fp += d * p3.v[FP_DIGITS - 1u - i];
  }
  return fp;
}

// 12-mer key (base 3) of a genome position; used for the taint list
WALT_HD uint32_t entry_key12(const uint64_t* __restrict__ genome, uint32_t e, bool ag, const Pow3& p3) {
  uint32_t key = 0;
  for (uint32_t i = 0; i < KEY_WEIGHT; ++i)
    key += ternary_digit(packed_base(genome, (uint64_t)e + PAD_BASES + 3u * i + 1u), ag) * p3.v[KEY_WEIGHT - 1u - i];
  return key;
}

// ------------------------------------------------------------------------------------------
// per-warp scratch (shared memory on the device)
// ------------------------------------------------------------------------------------------
// layout in 64-bit words: R[nw] | VM[3][nw] | SM[3][nw] | D[32*nw bytes] | C[LOOKUP_LANES*LANE_RUN_CAP]
constexpr uint32_t LOOKUP_LANES = 6;   // 2 strands x 3 seed shifts, one lane each
constexpr uint32_t LANE_RUN_CAP = 4;   // fingerprint-equal slots a single lane verifies itself
struct LaneCand { uint32_t g; uint32_t mm; };   // mm == NO_HIT: empty slot
struct ReadScratch {
  uint64_t* R;    // converted read, packed like the genome (no pad)
  uint64_t* VM;   // verification masks (low bit of each 2-bit field), per shift
  uint64_t* SM;   // seed (cared position) masks, per shift
  uint8_t* D;     // base-3 digit of every converted read base (rank inside the 3-letter alphabet)
  LaneCand* C;    // paired-end: candidates found by the lookup lanes, [lane][LANE_RUN_CAP]
  uint32_t nw;    // stride = words reserved per array
};
WALT_HD uint32_t scratch_words(uint32_t nw) { return 11u * nw + LOOKUP_LANES * LANE_RUN_CAP; }
WALT_HD ReadScratch carve_scratch(uint64_t* base, uint32_t nw) {
  ReadScratch s;
  s.R = base; s.VM = base + nw; s.SM = base + 4u * nw; s.D = reinterpret_cast<uint8_t*>(base + 7u * nw);
  s.C = reinterpret_cast<LaneCand*>(base + 11u * nw); s.nw = nw;
  return s;
}

// Four consecutive read characters p..p+3 in one register, first character in the low byte,
// zero at and beyond read_len.
WALT_HD uint32_t load4_ascii(const char* __restrict__ seq, uint32_t p, uint32_t read_len) {
  if (p >= read_len) return 0u;
  const uint32_t nv = read_len - p;
#if defined(__CUDA_ARCH__)
  // two aligned 32-bit loads + funnel shift; the second word is touched only if it holds read bytes
  const uintptr_t a = reinterpret_cast<uintptr_t>(seq + p);
  const uint32_t* w0 = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
  const uint32_t sh = (uint32_t)(a & 3u);
  const uint32_t lo = w0[0];
  uint32_t hi = 0u;
  if (sh != 0u && 4u - sh < nv) hi = w0[1];
  uint32_t v = __funnelshift_r(lo, hi, 8u * sh);
#else
  uint32_t v = 0u;
  for (uint32_t j = 0; j < 4u && j < nv; ++j) v |= (uint32_t)(uint8_t)seq[p + j] << (8u * j);
#endif
  if (nv < 4u) v &= (1u << (8u * nv)) - 1u;
  return v;
}

// read conversion (mapping.cpp:142-164) on four codes, one per byte: C->T sets bit 1 where bit 0
// is set; G->A clears bit 1 where bit 0 is clear
WALT_HD uint32_t convert4(uint32_t code4, bool ag) {
  const uint32_t b0 = code4 & 0x01010101u;
  return ag ? (code4 & ((b0 << 1) | 0x01010101u)) : (code4 | (b0 << 1));
}
WALT_HD uint32_t keep4(uint32_t n_valid) { return n_valid >= 4u ? 0xFFFFFFFFu : ((1u << (8u * n_valid)) - 1u); }

// The same four characters as 2-bit codes (one per byte), SIMD inside the register.
// *bad is set if a byte in front of read_len is not A/C/G/T.
WALT_HD uint32_t codes4(uint32_t ascii4, uint32_t n_valid, bool ag, bool& bad) {
  const uint32_t x = (ascii4 >> 1) & 0x03030303u;
  const uint32_t code = x ^ ((x >> 1) & 0x01010101u);                 // A0 C1 G2 T3 (util.hpp:107-121)
  // rebuild the letters from the codes; any difference is a non-ACGT byte
  uint32_t letters = 0u;
#if defined(__CUDA_ARCH__)
  const uint32_t sel = (code & 0x3u) | ((code >> 4) & 0x30u) | ((code >> 8) & 0x300u) | ((code >> 12) & 0x3000u);
  letters = __byte_perm(0x54474341u, 0u, sel);
#else
  for (uint32_t j = 0; j < 4u; ++j) letters |= (uint32_t)(uint8_t)"ACGT"[(code >> (8u * j)) & 3u] << (8u * j);
#endif
  const uint32_t keep = keep4(n_valid);
  bad |= ((letters ^ ascii4) & keep) != 0u;
  return convert4(code, ag) & keep;
}
// One byte of a 2-bit packed read (first base in the top bits, see walt_pack_reads in
// include/walt_host.h) -> the same register layout: four codes, first base in the low byte.
WALT_HD uint32_t packed_codes4(uint32_t byte, uint32_t n_valid, bool ag) {
  const uint32_t code = ((byte >> 6) & 3u) | (((byte >> 4) & 3u) << 8) | (((byte >> 2) & 3u) << 16) | ((byte & 3u) << 24);
  return convert4(code, ag) & keep4(n_valid);
}
// rank of each converted code inside its 3-letter alphabet (== ternary_digit, four at a time)
WALT_HD uint32_t digits4(uint32_t code4, bool ag) {
  return ag ? (((code4 + 0x01010101u) >> 1) & 0x03030303u)
            : (((code4 >> 1) & 0x01010101u) + (code4 & 0x01010101u));
}

// Pack + convert one read into sc.R and its base-3 digits into sc.D (group-cooperative).  The
// read arrives as ASCII (PACKED = false; returns false, uniformly, if a byte is not A/C/G/T) or
// already 2 bits per base, unconverted (PACKED = true: ceil(read_len / 4) bytes, first base in the
// top bits of each byte).  A group of W::WIDTH lanes owns the read: every lane turns 4 consecutive
// characters into one byte of the packed words (no cross-lane traffic).
template <class W, bool PACKED>
WALT_HD bool load_read(W& w, const char* __restrict__ seq, uint32_t read_len, bool ag, ReadScratch& sc) {
  const uint32_t lane = w.lane();
  const uint32_t nw = (read_len + 31u) >> 5;
  bool bad = false;
  {
    // every lane turns 4 consecutive characters into one byte of the packed words (no cross-lane traffic)
    uint8_t* rb = reinterpret_cast<uint8_t*>(sc.R);
    uint32_t* dw = reinterpret_cast<uint32_t*>(sc.D);
    for (uint32_t p = 4u * lane; p < 32u * nw; p += 4u * W::WIDTH) {
      const uint32_t nv = p < read_len ? read_len - p : 0u;
      uint32_t c4;
      if (PACKED) c4 = nv ? packed_codes4((uint8_t)seq[p >> 2], nv, ag) : 0u;
      else c4 = codes4(load4_ascii(seq, p, read_len), nv, ag, bad);
      // bytes c0..c3 (first character lowest) -> c0<<6 | c1<<4 | c2<<2 | c3
      rb[(p >> 5) * 8u + (7u - ((p & 31u) >> 2))] = (uint8_t)((c4 * 0x40100401u) >> 24);   // little-endian u64: first base on top
      dw[p >> 2] = digits4(c4, ag);
    }
  }
  bool any_bad = PACKED ? false : w.ballot(bad) != 0u;
  w.sync();
  return !any_bad;
}

// Verification / seed masks for every shift (depend on read_len only).
template <class W>
WALT_HD_NOINLINE void build_masks(W& w, uint32_t read_len, ReadScratch& sc) {   // once per read length: out of the hot loop
  const uint32_t lane = w.lane();
  const uint32_t nw = (read_len + 31u) >> 5;
  const uint32_t spr = seed_repeats(read_len);
  for (uint32_t s = 0; s < 3u; ++s) {
    for (uint32_t k = lane; k < nw; k += W::WIDTH) {
      uint64_t vm = 0, sm = 0;
      for (uint32_t j = 0; j < 32u; ++j) {
        const uint32_t p = 32u * k + j;
        vm = (vm << 2) | (is_verify_position(p, s, spr, read_len) ? 1ull : 0ull);
        sm = (sm << 2) | (is_seed_position(p, s, spr) ? 1ull : 0ull);
      }
      sc.VM[s * sc.nw + k] = vm;
      sc.SM[s * sc.nw + k] = sm;
    }
  }
  w.sync();
}

// ------------------------------------------------------------------------------------------
// candidate window compare (one lane, one candidate)
// ------------------------------------------------------------------------------------------
struct WindowResult {
  uint32_t mismatches;  // over the verification mask
  bool seed_equal;      // all cared positions equal
};

// genome window aligned with read position 0 at padded base position gp
WALT_HD WindowResult compare_window(const uint64_t* __restrict__ genome, uint64_t gp,
                                    const uint64_t* R, const uint64_t* VM, const uint64_t* SM,
                                    uint32_t nw) {
  const uint64_t* g = genome + (gp >> 5);
  const uint32_t sh = 2u * (uint32_t)(gp & 31u);
  uint64_t cur = g[0];
  uint32_t mm = 0;
  uint64_t sdiff = 0;
  for (uint32_t k = 0; k < nw; ++k) {
    uint64_t nxt = g[k + 1];
    uint64_t gw = sh ? ((cur << sh) | (nxt >> (64u - sh))) : cur;
    uint64_t x = gw ^ R[k];
    uint64_t d = x | (x >> 1);
    mm += popc64(d & VM[k]);
    sdiff |= d & SM[k];
    cur = nxt;
  }
  WindowResult r;
  r.mismatches = mm;
  r.seed_equal = (sdiff == 0);
  return r;
}

// three-way compare of the entry's seed characters against the read's: <0 entry smaller
WALT_HD int compare_seed(const uint64_t* __restrict__ genome, uint64_t gp, const uint64_t* R,
                         const uint64_t* SM, uint32_t nws) {
  const uint64_t* g = genome + (gp >> 5);
  const uint32_t sh = 2u * (uint32_t)(gp & 31u);
  uint64_t cur = g[0];
  for (uint32_t k = 0; k < nws; ++k) {
    uint64_t nxt = g[k + 1];
    uint64_t gw = sh ? ((cur << sh) | (nxt >> (64u - sh))) : cur;
    uint64_t m3 = SM[k] * 3ull;
    uint64_t a = gw & m3, b = R[k] & m3;
    if (a != b) return a < b ? -1 : 1;
    cur = nxt;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
// literal IndexRegion (mapping.cpp:166-222), executed redundantly by every lane
// ------------------------------------------------------------------------------------------
// genome character as the reference's raw byte read sees it, as an ordered code:
// 0 past the end of the sequence (below 'A'), else 1 + 2-bit code (A<C<G<T).
WALT_HD uint32_t literal_char(const uint64_t* __restrict__ genome, uint64_t pos, uint32_t genome_len) {
  if (pos >= genome_len) return 0u;
  return 1u + packed_base(genome, pos + PAD_BASES);
}

// LowerBound / UpperBound of mapping.cpp:166-196 over slots [low, high] for the probe value `ch`;
// probe(slot) is the genome character the reference reads for that slot, as an ordered code.
template <class Probe>
WALT_HD uint32_t literal_lower(uint32_t low, uint32_t high, uint32_t ch, Probe probe) {
  while (low < high) {
    const uint32_t mid = low + (high - low) / 2u;
    if (probe(mid) >= ch) high = mid; else low = mid + 1u;
  }
  return low;
}
template <class Probe>
WALT_HD uint32_t literal_upper(uint32_t low, uint32_t high, uint32_t ch, Probe probe) {
  while (low < high) {
    const uint32_t mid = low + (high - low + 1u) / 2u;
    if (probe(mid) <= ch) low = mid; else high = mid - 1u;
  }
  return low;
}
// one character of IndexRegion (mapping.cpp:198-222): false if the search fails here
template <class Probe>
WALT_HD bool literal_step(uint32_t& l, uint32_t& u, uint32_t ch, Probe probe) {
  l = literal_lower(l, u, ch, probe);
  u = literal_upper(l, u, ch, probe);
  return !(l == u && ch != probe(l));
}

// Slots [t0, t1) of the taint list that hold positions whose 12-mer key is key12 (its filter bit
// is known to be set).
WALT_HD void taint_slots(const SubIndexView& ix, uint32_t key12, uint32_t& t0, uint32_t& t1) {
  const uint32_t word = key12 >> 5, bit = key12 & 31u;
  const uint32_t idx = ix.taint_rank[word] + popc32(ix.taint_bits[word] & ((1u << bit) - 1u));
  t0 = ix.taint_start[idx]; t1 = ix.taint_start[idx + 1u];
}
// inverse of ternary_digit: the converted code of a base-3 digit (CT {A,G,T}, GA {A,C,T})
WALT_HD uint32_t digit_code(uint32_t d, bool ag) { return ag ? (d == 2u ? 3u : d) : (d == 0u ? 0u : d + 1u); }

// IndexRegion (mapping.cpp:198-222) replayed literally over the 12-mer bucket [first, second) of
// key12: same probes, same decisions, same result -- but most probes are answered without touching
// the entry or the genome.  Every probe asks for seed character p of the entry in some slot.  All
// entries but the tainted ones (whose probed characters may lie past their chromosome's end: the
// bucket's slice of the taint list, with their slots) were sorted by exactly the characters the
// reference reads, so for such a slot
//   * p < depth: inside the table range of the read's first p characters the character is given
//     by where the slot lies between the four table boundaries of the range's three sub-ranges
//     (four independent loads per CHARACTER instead of two dependent ones per probe);
//   * depth <= p < depth + FP_DIGITS: it is a digit of the entry's fingerprint (one load);
// everything else (tainted slots, slots outside the range, later characters) is read as the
// reference reads it.  On a bucket of 2^18 entries this turns ~400 dependent pairs of cache
// misses into about a dozen.
// bail: give up (return false, first/second untouched) once the region still holds more than `bail`
// slots past the table's characters -- a repeat, whose long search is not a single lane's job.
WALT_HD bool literal_index_region(const SubIndexView& ix, uint32_t genome_len, const Pow3& p3, const uint64_t* R,
                                  uint32_t seed_i, uint32_t seed_len, uint32_t key12, uint32_t& first,
                                  uint32_t& second, uint32_t bail = 0xFFFFFFFFu) {
  const bool ag = ix.ag != 0u;
  uint32_t t0 = 0u, t1 = 0u;
  if ((ix.taint_bits[key12 >> 5] >> (key12 & 31u)) & 1u) taint_slots(ix, key12, t0, t1);
  auto tainted = [&](uint32_t slot) {
    for (uint32_t t = t0; t < t1; ++t) if (ix.taint_slot[t] == slot) return true;
    return false;
  };
  uint32_t l = first, u = second - 1u;
  uint32_t prefix = key12;
  bool ok = true;
  for (uint32_t p = KEY_WEIGHT; p < seed_len && ok; ++p) {
    const uint32_t cp = 3u * p + 1u;
    const uint32_t code = packed_base(R, seed_i + cp);
    const uint32_t ch = 1u + code;
    auto read_it = [&](uint32_t slot) { return literal_char(ix.genome, (uint64_t)ix.entries[slot].pos + cp, genome_len); };
    if (p < ix.depth) {
      const uint32_t s = p3.v[ix.depth - 1u - p];
      const uint32_t b0 = ix.table[3u * prefix * s], b1 = ix.table[(3u * prefix + 1u) * s],
                     b2 = ix.table[(3u * prefix + 2u) * s], b3 = ix.table[(3u * prefix + 3u) * s];
      ok = literal_step(l, u, ch, [&](uint32_t slot) {
        if (slot < b0 || slot >= b3 || tainted(slot)) return read_it(slot);
        return 1u + digit_code((slot >= b1 ? 1u : 0u) + (slot >= b2 ? 1u : 0u), ag);
      });
      prefix = 3u * prefix + ternary_digit(code, ag);
    } else if (p < ix.depth + FP_DIGITS) {
      if (u - l > bail && u >= l) return false;
      const uint32_t div = p3.v[FP_DIGITS - 1u - (p - ix.depth)];
      ok = literal_step(l, u, ch, [&](uint32_t slot) {
        if (tainted(slot)) return read_it(slot);
        return 1u + digit_code((ix.entries[slot].fp / div) % 3u, ag);
      });
    } else {
      ok = literal_step(l, u, ch, read_it);
    }
  }
  if (!ok || l > u) { first = 1u; second = 0u; return true; }
  first = l; second = u;
  return true;
}

// ------------------------------------------------------------------------------------------
// sinks: what happens to each verified candidate, in reference order
// ------------------------------------------------------------------------------------------
// SE: BestMatch fold, mapping.cpp:306-313 (state is warp-uniform)
struct BestState {
  uint32_t pos, times, mm;
  uint32_t strand;  // '+' or '-'
};

// lane-local work counters (summed over lanes by the kernel epilogue)
struct Counters {
  uint32_t lookups, candidates, literal;
};
constexpr uint32_t NO_HIT = 0xFFFFFFFFu;

template <class W>
struct BestSink {
  BestState st;
  WALT_HD bool stop_before_shift(uint32_t seed_i) const {  // mapping.cpp:250-256
    return (st.mm == 0u && seed_i >= 1u) || (st.mm == 1u && seed_i >= 2u);
  }
  // can a candidate with this many mismatches still change the state?
  WALT_HD bool may_take(uint32_t mm) const { return mm <= st.mm; }
  // candidates of lanes (ascending) with valid set, in index order; g are distinct per lookup
  WALT_HD void consume(W& w, bool valid, uint32_t mm, uint32_t g, uint32_t strand) {
    uint32_t below = w.ballot(valid && mm < st.mm);
    uint32_t equal = w.ballot(valid && mm == st.mm);
    if (below) {
      // the minimum decides; every earlier improvement is overwritten by the first minimum
      uint32_t mn = mm;
      if (!(valid && mm < st.mm)) mn = 0xFFFFFFFFu;
      mn = w.reduce_min(mn);
      uint32_t e = w.ballot(valid && mm == mn);
      int last = 31 - clz32(e);
      st.pos = w.shfl(g, last);
      st.times = popc32(e);
      st.mm = mn;
      st.strand = strand;
    } else if (equal) {
      int first = ffs32(equal) - 1;
      int last = 31 - clz32(equal);
      uint32_t g_first = w.shfl(g, first);
      uint32_t g_last = w.shfl(g, last);
      uint32_t acc = popc32(equal) - (g_first == st.pos ? 1u : 0u);
      if (acc) { st.pos = g_last; st.strand = strand; st.times += acc; }
    }
  }
  // the same fold for one lookup summarised by a single lane: `cnt` candidates (distinct
  // positions, index order) share the lookup's minimum `mn`; first/last are their positions
  WALT_HD void apply(uint32_t mn, uint32_t cnt, uint32_t g_first, uint32_t g_last, uint32_t strand) {
    if (mn < st.mm) {
      st.pos = g_last; st.times = cnt; st.mm = mn; st.strand = strand;
    } else if (mn == st.mm) {
      const uint32_t acc = cnt - (g_first == st.pos ? 1u : 0u);
      if (acc) { st.pos = g_last; st.strand = strand; st.times += acc; }
    }
  }
};

// PE: TopCandidates (paired.hpp:51-74) with libstdc++'s heap (stl_heap.h:135-147,224-267).
// The heap lives in memory owned by the warp; lane 0 mutates it.
struct HeapEntry { uint32_t pos; uint32_t mm_strand; };  // mm in the low 16 bits, strand bit 16
WALT_HD uint32_t he_mm(const HeapEntry& e) { return e.mm_strand & 0xFFFFu; }

WALT_HD void heap_sift_up(HeapEntry* a, uint32_t hole, uint32_t top, HeapEntry v) {
  while (hole > top) {
    uint32_t parent = (hole - 1u) >> 1;
    if (!(he_mm(a[parent]) < he_mm(v))) break;
    a[hole] = a[parent];
    hole = parent;
  }
  a[hole] = v;
}
WALT_HD void heap_adjust(HeapEntry* a, uint32_t hole, uint32_t len, HeapEntry v) {
  const uint32_t top = hole;
  uint32_t child = hole;
  while (len >= 2u && child < (len - 1u) / 2u) {
    child = 2u * (child + 1u);
    if (he_mm(a[child]) < he_mm(a[child - 1u])) --child;
    a[hole] = a[child];
    hole = child;
  }
  if ((len & 1u) == 0u && len >= 2u && child == (len - 2u) / 2u) {
    child = 2u * (child + 1u);
    a[hole] = a[child - 1u];
    hole = child - 1u;
  }
  heap_sift_up(a, hole, top, v);
}
WALT_HD void heap_push_bounded(HeapEntry* a, uint32_t& size, uint32_t cap, HeapEntry v) {
  if (size < cap) {
    a[size] = v; ++size;
    heap_sift_up(a, size - 1u, 0u, v);
  } else if (he_mm(v) < he_mm(a[0])) {
    uint32_t n = size;
    if (n > 1u) { HeapEntry last = a[n - 1u]; a[n - 1u] = a[0]; heap_adjust(a, 0u, n - 1u, last); }
    a[n - 1u] = v;
    heap_sift_up(a, n - 1u, 0u, v);
  }
}
WALT_HD HeapEntry heap_pop(HeapEntry* a, uint32_t& size) {
  HeapEntry top = a[0];
  uint32_t n = size;
  if (n > 1u) { HeapEntry last = a[n - 1u]; a[n - 1u] = a[0]; heap_adjust(a, 0u, n - 1u, last); }
  size = n - 1u;
  return top;
}

template <class W>
struct HeapSink {
  HeapEntry* heap;     // capacity cap, visible to the whole warp
  uint32_t size;       // warp-uniform copy
  uint32_t cap;
  uint32_t top_mm;     // warp-uniform copy of heap[0].mm (valid if size > 0)
  uint32_t max_mm;
  WALT_HD bool stop_before_shift(uint32_t seed_i) const {  // paired.cpp:127-137
    bool full = size > 0u && size >= cap;
    return full && ((top_mm == 0u && seed_i >= 1u) || (top_mm == 1u && seed_i >= 2u));
  }
  // Can a candidate with this many mismatches still enter the heap?  Once the heap is full a push
  // needs mm < heap[0].mm (paired.hpp:60-67), and heap[0].mm only falls, so a candidate turned
  // away by the (warp-uniform, possibly stale) top_mm would be turned away by the heap as well.
  WALT_HD bool may_take(uint32_t mm) const { return mm <= max_mm && (size < cap || mm < top_mm); }
  WALT_HD void consume(W& w, bool valid, uint32_t mm, uint32_t g, uint32_t strand) {
    uint32_t take = w.ballot(valid && may_take(mm));
    if (!take) return;
    // pushes must happen one by one in lane order; lane 0 owns the heap
    while (take) {
      int src = ffs32(take) - 1;
      take &= take - 1u;
      uint32_t cg = w.shfl(g, src);
      uint32_t cm = w.shfl(mm, src);
      if (w.lane() == 0u) {
        HeapEntry v; v.pos = cg; v.mm_strand = cm | (strand == '-' ? 0x10000u : 0u);
        heap_push_bounded(heap, size, cap, v);
      }
    }
    w.sync();
    size = w.shfl(size, 0);
    uint32_t t = 0;
    if (w.lane() == 0u && size) t = he_mm(heap[0]);
    top_mm = w.shfl(t, 0);
  }
  // n candidates (all with mm <= max_mm, index order) found by one lookup lane
  WALT_HD void push_list(W& w, const LaneCand* c, uint32_t n, uint32_t strand) {
    if (w.lane() == 0u) {
      for (uint32_t k = 0; k < n; ++k) {
        HeapEntry v; v.pos = c[k].g; v.mm_strand = c[k].mm | (strand == '-' ? 0x10000u : 0u);
        heap_push_bounded(heap, size, cap, v);
      }
    }
    w.sync();
    size = w.shfl(size, 0);
    uint32_t t = 0;
    if (w.lane() == 0u && size) t = he_mm(heap[0]);
    top_mm = w.shfl(t, 0);
  }
};

// PE, two-phase form.  Heap maintenance is serial per read (one lane works, the rest of the warp
// idles) and on repeat-heavy input it was ~40 % of all issued instructions at ~1.2 active threads.
// The mapping kernel therefore only LOGS the candidates that change the heap, in order, and a
// second kernel replays every read's log through the libstdc++ heap with one THREAD per read.
// What the mapping pass needs from the heap -- is it full, what is its worst mismatch count
// (paired.hpp:60-67 for the push rule, paired.cpp:127-137 for the shift exits) -- depends only on
// the multiset of kept mismatch counts, which evolves without the heap's shape: while not full
// every candidate is kept; once full a candidate with mm < max(kept) replaces one entry with
// mm == max(kept).  `hist` (max_mm + 1 counters in memory owned by the group) is that multiset,
// so the log holds exactly the pushes that std::priority_queue would have executed.
constexpr uint32_t LOG_MAX_MM = 15;    // the logged form needs (max_mm + 1) * top_k log slots per read
WALT_HD uint32_t pe_log_slots(uint32_t top_k, uint32_t max_mm) { return top_k * (max_mm + 1u); }

template <class W>
struct LogSink {
  HeapEntry* log;      // this read's event list (global memory), pe_log_slots entries
  uint32_t* hist;      // kept entries per mismatch count, [0, max_mm]; lane 0 writes, all read
  uint32_t n_log;      // group-uniform
  uint32_t size;       // group-uniform, exact
  uint32_t cap;
  uint32_t top_mm;     // group-uniform, exact: the heap's heap[0].mm (valid if size > 0)
  uint32_t max_mm;
  WALT_HD void reset(W& w) {
    for (uint32_t i = w.lane(); i <= max_mm; i += W::WIDTH) hist[i] = 0u;
    n_log = 0u; size = 0u; top_mm = 0u;
    w.sync();
  }
  WALT_HD bool stop_before_shift(uint32_t seed_i) const {  // paired.cpp:127-137
    bool full = size > 0u && size >= cap;
    return full && ((top_mm == 0u && seed_i >= 1u) || (top_mm == 1u && seed_i >= 2u));
  }
  WALT_HD bool may_take(uint32_t mm) const { return mm <= max_mm && (size < cap || mm < top_mm); }
  // Candidates of lanes (ascending = index order) with valid set.  The reference pushes them one by
  // one (paired.hpp:60-67); here whole prefixes of the batch are taken per step, which is the same
  // thing because the state only changes in two ways:
  //  * not full: every candidate with mm <= max_mm is kept, so the first `cap - size` of them are
  //    taken together;
  //  * full: a candidate is kept iff mm < top_mm, and each take removes one entry with mm == top_mm.
  //    top_mm therefore stays put for the next hist[top_mm] takes: the first hist[top_mm] lanes
  //    with mm < top_mm are taken together, every lane in front of the last of them that was not
  //    taken has been turned away for good (top_mm only falls), and only then is top_mm lowered.
  // Log slots, histogram counters and top_mm end up exactly as after the one-by-one loop.
  WALT_HD void consume(W& w, bool valid, uint32_t mm, uint32_t g, uint32_t strand) {
    uint32_t pending = w.ballot(valid && may_take(mm));   // a superset: the state only gets stricter
    if (!pending) return;
    const uint32_t lane = w.lane();
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t sbit = strand == '-' ? 0x10000u : 0u;
    while (pending) {
      const bool mine = ((pending >> lane) & 1u) != 0u;
      const bool was_full = size >= cap;
      uint32_t T;
      if (!was_full) {
        T = w.ballot(mine && popc32(pending & lt) < cap - size);
      } else {
        const uint32_t P = w.ballot(mine && mm < top_mm);
        if (!P) break;
        T = w.ballot(((P >> lane) & 1u) != 0u && popc32(P & lt) < hist[top_mm]);   // hist[top_mm] >= 1
      }
      const bool in_t = ((T >> lane) & 1u) != 0u;
      const uint32_t n_t = popc32(T);
      const int last = 31 - clz32(T);
      if (in_t) {
        HeapEntry v; v.pos = g; v.mm_strand = mm | sbit;
        log[n_log + popc32(T & lt)] = v;
      }
      uint32_t mx = 0u;   // largest mismatch count among the taken
      if (n_t == 1u) {
        mx = w.shfl(mm, last);
        if (lane == 0u) hist[mx]++;
      } else {
        for (uint32_t v = 0; v <= max_mm; ++v) {
          const uint32_t c = popc32(w.ballot(in_t && mm == v));
          if (c) { mx = v; if (lane == 0u) hist[v] += c; }
        }
      }
      n_log += n_t;
      if (!was_full) {
        top_mm = (size == 0u || mx > top_mm) ? mx : top_mm;
        size += n_t;
      } else {
        if (lane == 0u) hist[top_mm] -= n_t;
        w.sync();
        uint32_t t = top_mm;
        while (t > 0u && hist[t] == 0u) --t;   // something below went in, so a counter at or below is set
        top_mm = t;
      }
      w.sync();                                // the next step's writer must not run ahead of these reads
      pending &= ~((2u << last) - 1u);         // lanes up to the last taken one are settled
    }
  }
  // n candidates (all with mm <= max_mm, index order) found by one lookup lane
  WALT_HD void push_list(W& w, const LaneCand* c, uint32_t n, uint32_t strand) {
    const uint32_t lane = w.lane();
    LaneCand v; v.g = 0u; v.mm = 0u;
    if (lane < n) v = c[lane];
    consume(w, lane < n, v.mm, v.g, strand);
  }
};

// Second phase: one read's log through TopCandidates (paired.hpp:51-74).  Every logged event is a
// push the reference executes; heap_push_bounded re-derives the same decision.
WALT_HD void replay_heap_log(const HeapEntry* log, uint32_t n_log, uint32_t cap, HeapEntry* heap, uint32_t& size) {
  size = 0u;
  for (uint32_t i = 0; i < n_log; ++i) heap_push_bounded(heap, size, cap, log[i]);
}

// ------------------------------------------------------------------------------------------
// one seed lookup: table -> region -> candidates, fed to the sink in reference order
// ------------------------------------------------------------------------------------------
struct MapConfig {
  uint32_t b;            // -b
  uint32_t literal_all;  // test hook: replay IndexRegion literally for every lookup
};


// Is there a tainted entry in this 12-mer bucket whose in-chromosome seed characters all
// match the read (so that the reference's search would look at its out-of-chromosome bytes)?
// Group-cooperative version (the lanes split the characters).
template <class W>
WALT_HD bool lookup_is_affected(W& w, const SubIndexView& ix, const uint64_t* R, uint32_t seed_i,
                                uint32_t seed_len, uint32_t key12) {
  uint32_t t0, t1;
  taint_slots(ix, key12, t0, t1);
  const uint32_t lane = w.lane();
  bool affected = false;
  for (uint32_t t = t0; t < t1; ++t) {
    uint32_t avail = ix.taint_len[t];
    uint32_t n_in = (avail + 1u) / 3u;  // seed characters with 3i+1 < avail
    if (n_in >= seed_len) continue;     // every probed character is inside the chromosome
    uint32_t e = ix.taint_pos[t];
    bool differ = false;
    for (uint32_t i = KEY_WEIGHT + lane; i < n_in; i += W::WIDTH) {
      uint32_t off = 3u * i + 1u;
      differ |= packed_base(ix.genome, (uint64_t)e + PAD_BASES + off) != packed_base(R, seed_i + off);
    }
    if (w.ballot(differ) == 0u) affected = true;
  }
  return affected;
}
// The same test by a single lane.
WALT_HD_NOINLINE bool lane_is_affected(const SubIndexView& ix, const uint64_t* R, uint32_t seed_i, uint32_t seed_len,
                              uint32_t key12) {
  uint32_t t0, t1;
  taint_slots(ix, key12, t0, t1);
  for (uint32_t t = t0; t < t1; ++t) {
    const uint32_t n_in = (ix.taint_len[t] + 1u) / 3u;
    if (n_in >= seed_len) continue;
    const uint32_t e = ix.taint_pos[t];
    bool differ = false;
    for (uint32_t i = KEY_WEIGHT; i < n_in && !differ; ++i) {
      const uint32_t off = 3u * i + 1u;
      differ = packed_base(ix.genome, (uint64_t)e + PAD_BASES + off) != packed_base(R, seed_i + off);
    }
    if (!differ) return true;
  }
  return false;
}

// (WD+1)-ary narrowing of [l, h) to at most WD slots around the first slot whose value is not
// `below` (pred(slot) true for a prefix of the range).
template <class W, class Pred>
WALT_HD void kary_narrow(W& w, uint32_t& l, uint32_t& h, Pred below) {
  constexpr uint32_t WD = W::WIDTH;
  const uint32_t lane = w.lane();
  while (h - l > WD) {
    const uint32_t p = l + (uint32_t)(((uint64_t)(lane + 1u) * (h - l)) / (WD + 1u));
    const uint32_t cnt = popc32(w.ballot(below(p)));
    const uint32_t nl = cnt ? w.shfl(p, (int)cnt - 1) + 1u : l;
    const uint32_t nh = cnt < WD ? w.shfl(p, (int)(cnt & (WD - 1u))) : h;
    l = nl; h = nh;
  }
}

// ------------------------------------------------------------------------------------------
// warp-wide verification of a run of index slots (repeats)
// ------------------------------------------------------------------------------------------
// Four lanes (a quad) own one candidate: lane q loads the q-th aligned 16 bytes (64 bases) of the
// 64-byte block that holds the candidate's window, so one 128-bit load instruction of the warp
// fetches eight whole windows, each from a single cache line (a lane-per-candidate loop touches 32
// lines per instruction, and the L1 serves about one line per two cycles).  A window of up to
// WIDE_MAX_READ bases starting anywhere in the first 64 bases of the block fits its 256 bases.
// Lane q compares read bases [64q, 64q + 64): its genome bases straddle its own and the next
// lane's 16 bytes (four shuffles), are aligned with two select levels and four funnel shifts, and
// the read words / masks of block q stay in registers for the whole run.
constexpr uint32_t WIDE_MAX_READ = 192;

struct Quad32 { uint32_t b[4]; };   // 64 bases in base order, 16 per word, first base on top

WALT_HD Quad32 load_genome16(const uint64_t* __restrict__ genome, uint64_t unit16) {
  Quad32 r;
#if defined(__CUDA_ARCH__)
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(genome) + unit16);   // little-endian u64 pairs
  r.b[0] = v.y; r.b[1] = v.x; r.b[2] = v.w; r.b[3] = v.z;
#else
  const uint64_t w0 = genome[2u * unit16], w1 = genome[2u * unit16 + 1u];
  r.b[0] = (uint32_t)(w0 >> 32); r.b[1] = (uint32_t)w0; r.b[2] = (uint32_t)(w1 >> 32); r.b[3] = (uint32_t)w1;
#endif
  return r;
}
// high word of (hi:lo) << n, n in [0, 31]
WALT_HD uint32_t funnel_left(uint32_t hi, uint32_t lo, uint32_t n) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_l(lo, hi, n);
#else
  return n ? ((hi << n) | (lo >> (32u - n))) : hi;
#endif
}
// 64-base block q of a packed array of nw words (zero beyond it), as four words in base order
WALT_HD Quad32 read_block(const uint64_t* a, uint32_t q, uint32_t nw) {
  const uint64_t w0 = 2u * q < nw ? a[2u * q] : 0ull, w1 = 2u * q + 1u < nw ? a[2u * q + 1u] : 0ull;
  Quad32 r;
  r.b[0] = (uint32_t)(w0 >> 32); r.b[1] = (uint32_t)w0; r.b[2] = (uint32_t)(w1 >> 32); r.b[3] = (uint32_t)w1;
  return r;
}

// One block of 32 consecutive slots in flight: lane L holds slot base + L; quad c verifies slots
// base + 4c + u in step u (u = 0..3), so after the four steps lane 4c + u picks step u's result and
// again holds slot base + L -- lane order is slot order, which is what the sinks consume.
struct WideBlock {
  Entry en;          // this lane's slot (zero past the end)
  Quad32 g[4];       // this lane's 16 bytes of the four candidates of its quad
  uint32_t sh;       // their alignments inside the 64-base unit, 8 bits each
};

template <class W>
WALT_HD Entry wide_load_entry(W& w, const SubIndexView& ix, uint32_t base, uint32_t last_excl) {
  const uint32_t slot = base + w.lane();
  Entry en; en.pos = 0u; en.fp = 0u;
  if (slot < last_excl && base < last_excl) en = ix.entries[slot];
  return en;
}
template <class W>
WALT_HD void wide_load_genome(W& w, const SubIndexView& ix, uint32_t seed_i, WideBlock& blk) {
  const uint32_t lane = w.lane(), q = lane & 3u, c4 = lane & ~3u;
  blk.sh = 0u;
  WALT_UNROLL
  for (uint32_t u = 0; u < 4u; ++u) {
    const uint32_t e = w.shfl(blk.en.pos, (int)(c4 + u));       // slot 0 of a finished run: genome start, harmless
    const uint64_t gp = (uint64_t)e + PAD_BASES - seed_i;
    blk.g[u] = load_genome16(ix.genome, (gp >> 6) + q);
    blk.sh |= ((uint32_t)gp & 63u) << (8u * u);
  }
}
// mismatches of the lane's own slot over the verification mask (low 16 bits) + 0x10000 per quad lane
// that saw a difference on a cared position (0 in the upper half <=> seed-equal)
template <class W>
WALT_HD uint32_t wide_compare(W& w, const WideBlock& blk, const Quad32& Rq, const Quad32& VMq, const Quad32& SMq) {
  const uint32_t lane = w.lane(), q = lane & 3u;
  uint32_t mine = 0u;
  WALT_UNROLL
  for (uint32_t u = 0; u < 4u; ++u) {
    uint32_t B[8];
    WALT_UNROLL
    for (uint32_t i = 0; i < 4u; ++i) { B[i] = blk.g[u].b[i]; B[4u + i] = w.shfl(blk.g[u].b[i], (int)((lane + 1u) & 31u)); }
    const uint32_t s = (blk.sh >> (8u * u)) & 63u;   // bases; 2s bits = (s >> 4) words + (2s & 31) bits
    const uint32_t off = s >> 4, bit = (2u * s) & 31u;
    uint32_t t[6], v[5];
    WALT_UNROLL
    for (uint32_t i = 0; i < 6u; ++i) t[i] = (off & 2u) ? B[i + 2u] : B[i];
    WALT_UNROLL
    for (uint32_t i = 0; i < 5u; ++i) v[i] = (off & 1u) ? t[i + 1u] : t[i];
    uint32_t mm = 0u, sd = 0u;
    WALT_UNROLL
    for (uint32_t i = 0; i < 4u; ++i) {
      const uint32_t x = funnel_left(v[i], v[i + 1u], bit) ^ Rq.b[i];
      const uint32_t d = x | (x >> 1);
      mm += popc32(d & VMq.b[i]);
      sd |= d & SMq.b[i];
    }
    mm += sd ? 0x10000u : 0u;
    mm += w.shfl(mm, (int)(lane ^ 1u));
    mm += w.shfl(mm, (int)(lane ^ 2u));
    mine = q == u ? mm : mine;
  }
  return mine;
}

// Feed the candidates of slots [first, last_excl) to the sink in slot order.  by_fp: the candidates
// are the leading slots whose fingerprint lies in [fp_lo, fp_lo + fp_span] (a prefix of the range:
// fingerprints are sorted inside a table range) and, of those, the seed-equal ones -- the
// reference's narrowed region; otherwise every slot is a candidate.  W::WIDTH == 32 and
// read_len <= WIDE_MAX_READ.  The entries of the next block are loaded while the windows of the
// current one are in flight; other warps of the SM cover the rest of the latency.
template <class W, class Sink>
WALT_HD void verify_run_wide(W& w, const SubIndexView& ix, const ChromView& cv, const ReadScratch& sc, uint32_t read_len,
                             uint32_t seed_i, uint32_t strand, uint32_t first, uint32_t last_excl, bool by_fp,
                             uint32_t fp_lo, uint32_t fp_span, Sink& sink, Counters& ctr) {
  const uint32_t lane = w.lane(), q = lane & 3u;
  const uint32_t nw = (read_len + 31u) >> 5;
  const Quad32 Rq = read_block(sc.R, q, nw);
  const Quad32 VMq = read_block(sc.VM + seed_i * sc.nw, q, nw);
  Quad32 SMq = read_block(sc.SM + seed_i * sc.nw, q, nw);
  if (!by_fp) { SMq.b[0] = SMq.b[1] = SMq.b[2] = SMq.b[3] = 0u; }
  WideBlock blk;
  blk.en = wide_load_entry(w, ix, first, last_excl);
  for (uint32_t base = first; base < last_excl; base += 32u) {
    const uint32_t ok = w.ballot(base + lane < last_excl && (!by_fp || blk.en.fp - fp_lo <= fp_span));   // unsigned: also rejects fp < fp_lo
    if (!ok) break;
    if (!((ok >> lane) & 1u)) blk.en.pos = 0u;
    wide_load_genome(w, ix, seed_i, blk);
    const uint32_t e = blk.en.pos;
    const bool more = ok == 0xFFFFFFFFu;   // the run goes on only if every slot of this block was in it
    if (more) blk.en = wide_load_entry(w, ix, base + 32u, last_excl);
    const uint32_t r = wide_compare(w, blk, Rq, VMq, SMq);
    const uint32_t mm = r & 0xFFFFu;
    bool valid = ((ok >> lane) & 1u) != 0u && (r >> 16) == 0u;
    if (valid) ctr.candidates++;
    valid = valid && sink.may_take(mm);
    uint32_t g = 0u;
    if (valid) {   // bounds, mapping.cpp:281-286
      const uint32_t chr = chrom_of(cv.starts, cv.n_chr, e);
      g = e - seed_i;
      valid = (e - cv.starts[chr] >= seed_i) && !(g + read_len >= cv.starts[chr + 1u]);
    }
    sink.consume(w, valid, mm, g, strand);
    if (!more) break;
  }
}

// Exact region [first, last_excl) of one lookup -> sink, in slot order.
template <class W, class Sink>
WALT_HD void verify_region(W& w, const SubIndexView& ix, const ChromView& cv, const ReadScratch& sc, uint32_t read_len,
                           uint32_t seed_i, uint32_t strand, uint32_t first, uint32_t last_excl, Sink& sink, Counters& ctr) {
  constexpr uint32_t WD = W::WIDTH;
  const uint32_t lane = w.lane();
  const uint32_t nw = (read_len + 31u) >> 5;
  const uint64_t* R = sc.R;
  const uint64_t* VM = sc.VM + seed_i * sc.nw;
  const uint64_t* SM = sc.SM + seed_i * sc.nw;
  // Exact region [first, last_excl): every slot is a candidate.  Long regions (repeats) are a
  // chain of dependent gathers -- entry, then genome window -- so every lane keeps SLOTS_AHEAD
  // of them in flight: the entries of a block are loaded together and their windows prefetched
  // before the first is compared.  Candidates reach the sink in slot order, WD at a time.
#ifndef WALT_SLOTS_AHEAD
#define WALT_SLOTS_AHEAD 4
#endif
  if (WD == 32u && read_len <= WIDE_MAX_READ) {
    verify_run_wide(w, ix, cv, sc, read_len, seed_i, strand, first, last_excl, false, 0u, 0u, sink, ctr);
    return;
  }
  constexpr uint32_t SLOTS_AHEAD = WALT_SLOTS_AHEAD;
  for (uint32_t base = first; base < last_excl; base += WD * SLOTS_AHEAD) {
    uint32_t e[SLOTS_AHEAD];
    WALT_UNROLL
    for (uint32_t u = 0; u < SLOTS_AHEAD; ++u) {
      const uint32_t slot = base + u * WD + lane;
      e[u] = slot < last_excl ? ix.entries[slot].pos : 0u;   // 64 readable pad entries behind index[]
    }
    WALT_UNROLL
    for (uint32_t u = 0; u < SLOTS_AHEAD; ++u) {
      if (base + u * WD + lane < last_excl) {
        const uint64_t* g0 = ix.genome + (((uint64_t)e[u] + PAD_BASES - seed_i) >> 5);
        WALT_PREFETCH(g0);
        WALT_PREFETCH(g0 + nw);
      }
    }
#ifndef WALT_ROLLED_COMPARE
#define WALT_ROLLED_COMPARE 1
#endif
#if WALT_ROLLED_COMPARE
    // One copy of the compare / bounds / sink code instead of four.  With the body unrolled the
    // repeat path was bound by instruction fetch (stalled_no_instruction 19.7 cycles per issue:
    // the four groups of a warp run this loop out of step, each fetching its own stream through a
    // ~25 KB body); rolled, configs[4] went from 215 to 149 ms per 5 M pairs.
    WALT_NO_UNROLL
    for (uint32_t u = 0; u < SLOTS_AHEAD; ++u) {
      if (base + u * WD >= last_excl) break;                 // uniform
      uint32_t eu = e[0];   // e[] lives in registers: select, do not index
      WALT_UNROLL
      for (uint32_t q = 1; q < SLOTS_AHEAD; ++q) eu = u == q ? e[q] : eu;
      bool valid = base + u * WD + lane < last_excl;
      uint32_t g = 0, mm = 0;
      if (valid) {
        WindowResult r = compare_window(ix.genome, (uint64_t)eu + PAD_BASES - seed_i, R, VM, SM, nw);
        mm = r.mismatches;
        ctr.candidates++;
        valid = sink.may_take(mm);
      }
      if (valid) {
        const uint32_t chr = chrom_of(cv.starts, cv.n_chr, eu);
        g = eu - seed_i;
        valid = (eu - cv.starts[chr] >= seed_i) && !(g + read_len >= cv.starts[chr + 1u]);
      }
      sink.consume(w, valid, mm, g, strand);
    }
#else
    WALT_UNROLL
    for (uint32_t u = 0; u < SLOTS_AHEAD; ++u) {
      if (base + u * WD >= last_excl) break;                 // uniform
      bool valid = base + u * WD + lane < last_excl;
      uint32_t g = 0, mm = 0;
      if (valid) {
        WindowResult r = compare_window(ix.genome, (uint64_t)e[u] + PAD_BASES - seed_i, R, VM, SM, nw);
        mm = r.mismatches;
        ctr.candidates++;
        valid = sink.may_take(mm);
      }
      if (valid) {
        // bounds, mapping.cpp:281-286 (all uint32 arithmetic, as the reference)
        const uint32_t chr = chrom_of(cv.starts, cv.n_chr, e[u]);
        g = e[u] - seed_i;
        valid = (e[u] - cv.starts[chr] >= seed_i) && !(g + read_len >= cv.starts[chr + 1u]);
      }
      sink.consume(w, valid, mm, g, strand);
    }
#endif
  }
}

template <class W, class Sink>
WALT_HD void seed_lookup(W& w, const SubIndexView& ix, const ChromView& cv, const Pow3& p3,
                         const MapConfig& cfg, const ReadScratch& sc, uint32_t read_len,
                         uint32_t seed_i, uint32_t strand, Sink& sink, Counters& ctr) {
  constexpr uint32_t WD = W::WIDTH;
  const uint32_t lane = w.lane();
  const uint32_t spr = seed_repeats(read_len);
  const uint32_t seed_len = spr;
  const uint32_t nw = (read_len + 31u) >> 5;
  const uint32_t nws = (3u * spr + seed_i + 31u) >> 5;  // words that hold seed positions
  const bool ag = ix.ag != 0u;
  const uint64_t* R = sc.R;
  const uint64_t* VM = sc.VM + seed_i * sc.nw;
  const uint64_t* SM = sc.SM + seed_i * sc.nw;

  // base-3 numbers of the read's seed characters: table key over the first min(depth, seed_len),
  // 12-mer key, fingerprint over the next min(FP_DIGITS, seed_len - depth)
  const uint32_t n_pref = ix.depth < seed_len ? ix.depth : seed_len;
  const uint32_t n_fp = seed_len > ix.depth ? (seed_len - ix.depth < FP_DIGITS ? seed_len - ix.depth : FP_DIGITS) : 0u;
  uint32_t part = 0, part12 = 0, partfp = 0;
  for (uint32_t i = lane; i < n_pref + n_fp; i += WD) {
    const uint32_t d = ternary_digit(packed_base(R, seed_i + 3u * i + 1u), ag);
    if (i < n_pref) part += d * p3.v[ix.depth - 1u - i];
    else partfp += d * p3.v[FP_DIGITS - 1u - (i - n_pref)];
    if (i < KEY_WEIGHT) part12 += d * p3.v[KEY_WEIGHT - 1u - i];
  }
  const uint32_t lo_key = w.reduce_add(part);
  const uint32_t key12 = w.reduce_add(part12);
  const uint32_t hi_key = lo_key + p3.v[ix.depth - n_pref];
  const uint32_t fp_lo = w.reduce_add(partfp);
  const uint32_t fp_span = p3.v[FP_DIGITS - n_fp] - 1u;   // fingerprints in [fp_lo, fp_lo + fp_span] match

  // the table range of the read's seed prefix; an empty 12-mer bucket (reference:
  // counter[h] == counter[h+1] -> continue, mapping.cpp:268-272) has an empty range too
  const uint32_t lo = ix.table[lo_key], hi = ix.table[hi_key];
  const uint32_t tainted = (ix.taint_bits[key12 >> 5] >> (key12 & 31u)) & 1u;

  bool literal = cfg.literal_all != 0u;
  if (!literal && tainted) literal = lookup_is_affected(w, ix, R, seed_i, seed_len, key12);

  uint32_t first, last_excl;  // candidate slots [first, last_excl)
  if (literal) {
    const uint32_t k12_span = p3.v[ix.depth - KEY_WEIGHT];
    const uint32_t bucket_lo = ix.table[key12 * k12_span];
    const uint32_t bucket_hi = ix.table[(key12 + 1u) * k12_span];
    if (bucket_lo == bucket_hi) return;
    if (lane == 0u) { ctr.lookups++; ctr.literal++; }
    uint32_t f = bucket_lo, s = bucket_hi;
    literal_index_region(ix, cv.genome_len, p3, R, seed_i, seed_len, key12, f, s);
    if (s - f + 1u > cfg.b) return;      // mapping.cpp:275-277 (u32 arithmetic; (1,0) -> 0)
    if (f > s) return;                   // failed search: empty candidate loop
    first = f; last_excl = s + 1u;
  } else {
    if (lo == hi) return;
    if (lane == 0u) ctr.lookups++;
    // 1. narrow to the slots whose fingerprint matches (integer compares on the entry array)
    uint32_t l = lo, h = hi;
    kary_narrow(w, l, h, [&](uint32_t p) { return ix.entries[p].fp < fp_lo; });
    Entry en; en.pos = 0u; en.fp = 0u;
    const bool in_chunk = l + lane < h;
    if (in_chunk) en = ix.entries[l + lane];
    const uint32_t below = w.ballot(in_chunk && en.fp < fp_lo);
    const uint32_t match = w.ballot(in_chunk && en.fp >= fp_lo && en.fp - fp_lo <= fp_span);
    const uint32_t f0 = l + popc32(below);
    const uint32_t n_match = popc32(match);
    // the run of matches starts at f0; it is complete if it stops before the chunk does, or
    // the chunk reaches the end of the table range
    const bool complete = (f0 + n_match < h) || (h == hi);
    if (WD == 32u && read_len <= WIDE_MAX_READ && !(complete && n_match > cfg.b)) {
      // Whole warp on one lookup (the kernels that take over parked reads): the fingerprint run
      // that starts at f0 is streamed through the quad verification, eight windows per load
      // instruction; seed equality is checked per candidate, so no equal-range search on the
      // genome is needed.  The narrowed region (what -b applies to, mapping.cpp:275-277) is a
      // subset of the run: a run of at most b slots passes the filter whatever it holds.
      if (complete && n_match == 0u) return;
      bool longer = false;   // does the run reach past b slots?
      if (!complete && (uint64_t)f0 + cfg.b < hi) longer = ix.entries[f0 + cfg.b].fp - fp_lo <= fp_span;
      if (!longer) {
        verify_run_wide(w, ix, cv, sc, read_len, seed_i, strand, f0, complete ? f0 + n_match : hi, true, fp_lo, fp_span,
                        sink, ctr);
        return;
      }
    } else if (complete) {
      if (n_match == 0u) return;
      // 2a. common case: the lanes already hold the candidates -- one genome window each
      const bool cand = ((match >> lane) & 1u) != 0u;
      uint32_t g = 0, mm = 0;
      bool seed_eq = false, valid = false;
      if (cand) {
        const uint32_t e = en.pos;
        WindowResult r = compare_window(ix.genome, (uint64_t)e + PAD_BASES - seed_i, R, VM, SM, nw);
        mm = r.mismatches;
        seed_eq = r.seed_equal;
        const uint32_t chr = chrom_of(cv.starts, cv.n_chr, e);   // bounds, mapping.cpp:281-286
        g = e - seed_i;
        valid = seed_eq && (e - cv.starts[chr] >= seed_i) && !(g + read_len >= cv.starts[chr + 1u]);
      }
      // the reference's narrowed region = the seed-equal slots; -b applies to its size
      if (popc32(w.ballot(seed_eq)) > cfg.b) return;
      if (valid) ctr.candidates++;
      sink.consume(w, valid, mm, g, strand);
      return;
    }
    // 2b. long runs (repeats): finish the fingerprint range, then equal-range on the genome
    uint32_t f1;
    {
      uint32_t l2 = f0 + n_match, h2 = hi;
      kary_narrow(w, l2, h2, [&](uint32_t p) { const uint32_t f = ix.entries[p].fp; return f < fp_lo || f - fp_lo <= fp_span; });
      bool in = false;
      if (l2 + lane < h2) { const uint32_t f = ix.entries[l2 + lane].fp; in = f < fp_lo || f - fp_lo <= fp_span; }
      f1 = l2 + popc32(w.ballot(in));
    }
    l = f0; h = f1;
    kary_narrow(w, l, h, [&](uint32_t p) {
      return compare_seed(ix.genome, (uint64_t)ix.entries[p].pos + PAD_BASES - seed_i, R, SM, nws) < 0; });
    {
      bool lt = false;
      if (l + lane < h)
        lt = compare_seed(ix.genome, (uint64_t)ix.entries[l + lane].pos + PAD_BASES - seed_i, R, SM, nws) < 0;
      first = l + popc32(w.ballot(lt));
    }
    l = first; h = f1;
    kary_narrow(w, l, h, [&](uint32_t p) {
      return compare_seed(ix.genome, (uint64_t)ix.entries[p].pos + PAD_BASES - seed_i, R, SM, nws) <= 0; });
    {
      bool le = false;
      if (l + lane < h)
        le = compare_seed(ix.genome, (uint64_t)ix.entries[l + lane].pos + PAD_BASES - seed_i, R, SM, nws) <= 0;
      last_excl = l + popc32(w.ballot(le));
    }
    if (last_excl <= first) return;
    if (last_excl - first > cfg.b) return;
  }

  verify_region(w, ix, cv, sc, read_len, seed_i, strand, first, last_excl, sink, ctr);
}

// Out-of-line copy of seed_lookup for the read loops: the cooperative replay is the rare path
// (tainted buckets, long repeat runs, the literal test mode), and keeping it out of the hot
// loop keeps that loop inside the instruction cache.  Callers hand in copies of their state so
// that nothing of the hot path has its address taken.
template <class W, class Sink>
WALT_HD_NOINLINE void seed_lookup_cold(W& w, const SubIndexView& ix, const ChromView& cv, const Pow3& p3,
                                       const MapConfig& cfg, const ReadScratch& sc, uint32_t read_len,
                                       uint32_t seed_i, uint32_t strand, Sink& sink, Counters& ctr) {
  seed_lookup(w, ix, cv, p3, cfg, sc, read_len, seed_i, strand, sink, ctr);
}
template <class W, class Sink>
WALT_HD void replay_lookup(W& w, const SubIndexView& ix, const ChromView& cv, const Pow3& p3,
                           const MapConfig& cfg, const ReadScratch& sc, uint32_t read_len,
                           uint32_t seed_i, uint32_t strand, Sink& sink, Counters& ctr) {
  W wc = w;
  Sink tmp = sink;
  Counters c{0u, 0u, 0u};
  ReadScratch sc2 = sc;
  seed_lookup_cold(wc, ix, cv, p3, cfg, sc2, read_len, seed_i, strand, tmp, c);
  w = wc;
  sink = tmp;
  ctr.lookups += c.lookups; ctr.candidates += c.candidates; ctr.literal += c.literal;
}

// ------------------------------------------------------------------------------------------
// lane-per-lookup fast path
// ------------------------------------------------------------------------------------------
// The six lookups of a read (2 strand sub-indexes x 3 seed shifts) are independent until their
// candidates are folded, so six lanes of the group run them concurrently -- one dependent chain
// table -> entries -> genome window per lane instead of six in sequence -- and only the fold
// (with the reference's early exits, mapping.cpp:248-257 / paired.cpp:127-137) is ordered.
// A lane gives up when the lookup needs the group: more than LANE_RUN_CAP candidate slots (repeats).
// A lookup that needs the literal IndexRegion replay (tainted bucket) is replayed by its lane.
// `emit(g, mm)` receives the verified candidates in index order; `discard()` is called if the
// lookup turns out to be filtered by -b after some were emitted.
// The literal IndexRegion replay for one lookup by one lane: 0 if the 12-mer bucket is empty, 1 with
// the inclusive region ((1, 0) for a failed search), 2 if the region stays long (a repeat).
constexpr uint32_t LANE_LITERAL_BAIL = 1024;
WALT_HD_NOINLINE uint32_t lane_literal_region(const SubIndexView& ix, uint32_t genome_len, const Pow3& p3, const uint64_t* R,
                                              uint32_t seed_i, uint32_t seed_len, uint32_t key12, uint32_t& f, uint32_t& t) {
  const uint32_t k12_span = p3.v[ix.depth - KEY_WEIGHT];
  const uint32_t bucket_lo = ix.table[key12 * k12_span], bucket_hi = ix.table[(key12 + 1u) * k12_span];
  if (bucket_lo == bucket_hi) return 0u;
  f = bucket_lo; t = bucket_hi;
  return literal_index_region(ix, genome_len, p3, R, seed_i, seed_len, key12, f, t, LANE_LITERAL_BAIL) ? 1u : 2u;
}

// LANE_GROUP: the whole lookup is the group's job (a literal region of more than LANE_RUN_CAP slots).  LANE_RUN: more than LANE_RUN_CAP
// fingerprint-equal slots; `run` then says where the run starts and where its table range ends, so
// that a whole warp can stream it (run_lookup) without searching again.
enum LaneResult : uint32_t { LANE_DONE = 0u, LANE_GROUP = 1u, LANE_RUN = 2u };
struct LaneRun { uint32_t f0, hi, fp_lo; };
WALT_HD uint32_t lookup_fp_span(const SubIndexView& ix, uint32_t read_len, const Pow3& p3) {
  const uint32_t seed_len = seed_repeats(read_len);
  const uint32_t n_fp = seed_len > ix.depth ? (seed_len - ix.depth < FP_DIGITS ? seed_len - ix.depth : FP_DIGITS) : 0u;
  return p3.v[FP_DIGITS - n_fp] - 1u;
}

template <class Emit, class Discard>
WALT_HD LaneResult lane_lookup(const SubIndexView& ix, const ChromView& cv, const Pow3& p3, const MapConfig& cfg,
                               const ReadScratch& sc, uint32_t read_len, uint32_t seed_i, Emit emit, Discard discard,
                               Counters& ctr, LaneRun& run) {
  const uint32_t seed_len = seed_repeats(read_len);
  const uint32_t nw = (read_len + 31u) >> 5;
  const uint32_t n_pref = ix.depth < seed_len ? ix.depth : seed_len;
  const uint32_t n_fp = seed_len > ix.depth ? (seed_len - ix.depth < FP_DIGITS ? seed_len - ix.depth : FP_DIGITS) : 0u;
  // base-3 numbers of the read's seed characters (Horner over the digit bytes)
  const uint8_t* D = sc.D + seed_i + 1u;
  uint32_t key = 0u, key12 = 0u, fp = 0u;
  for (uint32_t i = 0; i < KEY_WEIGHT; ++i) key = key * 3u + D[3u * i];
  key12 = key;
  for (uint32_t i = KEY_WEIGHT; i < n_pref; ++i) key = key * 3u + D[3u * i];
  for (uint32_t i = n_pref; i < n_pref + n_fp; ++i) fp = fp * 3u + D[3u * i];
  const uint32_t lo_key = key * p3.v[ix.depth - n_pref];
  const uint32_t hi_key = lo_key + p3.v[ix.depth - n_pref];
  const uint32_t fp_lo = fp * p3.v[FP_DIGITS - n_fp];
  const uint32_t fp_span = p3.v[FP_DIGITS - n_fp] - 1u;   // fingerprints in [fp_lo, fp_lo + fp_span] match

  const uint32_t lo = ix.table[lo_key], hi = ix.table[hi_key];
  // a tainted bucket needs the literal replay only if the read agrees with a tainted position
  // on every seed character inside the chromosome (rare) ...
  bool literal = false;
  uint32_t l = lo, h = hi;
  Entry en[LANE_RUN_CAP];
  uint32_t match = 0u;
  if (((ix.taint_bits[key12 >> 5] >> (key12 & 31u)) & 1u) && lane_is_affected(ix, sc.R, seed_i, seed_len, key12)) {
    // ... which this lane does on its own (a chain of dependent loads that the other warps of the SM
    // cover); only a literal region longer than LANE_RUN_CAP goes to the group
    uint32_t f, t;
    const uint32_t how = lane_literal_region(ix, cv.genome_len, p3, sc.R, seed_i, seed_len, key12, f, t);
    if (how == 0u) return LANE_DONE;   // empty bucket
    if (how == 2u || (f <= t && t - f + 1u <= cfg.b && t - f + 1u > LANE_RUN_CAP)) return LANE_GROUP;
    ctr.lookups++; ctr.literal++;
    if (t - f + 1u > cfg.b || f > t) return LANE_DONE;      // mapping.cpp:275-277 (u32 arithmetic; (1,0) -> 0); failed search
    literal = true;
    l = f;
    match = (1u << (t - f + 1u)) - 1u;                      // every slot of the region is a candidate
    WALT_UNROLL
    for (uint32_t k = 0; k < LANE_RUN_CAP; ++k) en[k] = ix.entries[l + k];
  }
  if (!literal) {
    if (lo == hi) return LANE_DONE;
    ctr.lookups++;
    // first slot of [lo, hi) whose fingerprint is >= fp_lo: bisect down to a LANE_RUN_CAP window
    // (the window [l, l + CAP) must hold slot h unless h is the end of the range)
    while (h - l > LANE_RUN_CAP || (h - l == LANE_RUN_CAP && h < hi)) {
      const uint32_t mid = l + ((h - l) >> 1);
      if (ix.entries[mid].fp < fp_lo) l = mid + 1u; else h = mid;
    }
    WALT_UNROLL
    for (uint32_t k = 0; k < LANE_RUN_CAP; ++k) en[k] = ix.entries[l + k];   // 64 readable pad entries behind index[]
    WALT_UNROLL
    for (uint32_t k = 0; k < LANE_RUN_CAP; ++k)
      if (l + k < hi && en[k].fp - fp_lo <= fp_span) match |= 1u << k;      // unsigned: also rejects fp < fp_lo
    if (match == 0u) return LANE_DONE;
    // fingerprints are sorted inside a table range, so the matches are one run; it is complete
    // unless it touches the end of the window and the range goes on
    if ((match >> (LANE_RUN_CAP - 1u)) && l + LANE_RUN_CAP < hi) {
      // ... then the window is moved onto the run's first slot (a lone hit in the window's last slot
      // is the common case on T-rich keys, whose table ranges hold many slots), and only a run that
      // fills it AND goes on is somebody else's job
      const uint32_t f0 = l + (uint32_t)ffs32(match) - 1u;
      if (f0 != l) {
        l = f0;
        WALT_UNROLL
        for (uint32_t k = 0; k < LANE_RUN_CAP; ++k) en[k] = ix.entries[l + k];
        match = 0u;
        WALT_UNROLL
        for (uint32_t k = 0; k < LANE_RUN_CAP; ++k)
          if (l + k < hi && en[k].fp - fp_lo <= fp_span) match |= 1u << k;
      }
      if (match == (1u << LANE_RUN_CAP) - 1u && l + LANE_RUN_CAP < hi && ix.entries[l + LANE_RUN_CAP].fp - fp_lo <= fp_span) {
        run.f0 = l; run.hi = hi; run.fp_lo = fp_lo;
        WALT_PREFETCH(ix.entries + run.f0 + LANE_RUN_CAP + 16u);   // the streaming pass starts here
        return LANE_RUN;
      }
    }
  }
  const uint64_t* R = sc.R;
  const uint64_t* VM = sc.VM + seed_i * sc.nw;
  const uint64_t* SM = sc.SM + seed_i * sc.nw;
  uint32_t n_region = 0u;   // the reference's narrowed region = seed-equal slots (-b applies to it)
  WALT_NO_UNROLL
  while (match) {   // lanes walk their own set bits, so the warp makes max-over-lanes passes
    const uint32_t k = (uint32_t)ffs32(match) - 1u;
    match &= match - 1u;
    uint32_t e = 0u;   // en[k].pos without dynamic register indexing
    WALT_UNROLL
    for (uint32_t q = 0; q < LANE_RUN_CAP; ++q) if (q == k) e = en[q].pos;
    const WindowResult r = compare_window(ix.genome, (uint64_t)e + PAD_BASES - seed_i, R, VM, SM, nw);
    if (!literal) {
      if (!r.seed_equal) continue;
      ++n_region;
    }
    const uint32_t chr = chrom_of(cv.starts, cv.n_chr, e);   // bounds, mapping.cpp:281-286
    const uint32_t g = e - seed_i;
    if ((e - cv.starts[chr] >= seed_i) && !(g + read_len >= cv.starts[chr + 1u])) {
      emit(g, r.mismatches);
      ctr.candidates++;
    }
  }
  if (n_region > cfg.b) discard();   // mapping.cpp:275-277
  return LANE_DONE;
}

// A long fingerprint run found by lookup lane j (LANE_RUN), fed to the sink.  A whole warp streams
// it through the quad verification straight from the run's first slot; narrower groups (and runs
// that may exceed -b, whose exact region size matters) take the general cooperative lookup.
template <class W, class Sink>
WALT_HD void run_lookup(W& w, const SubIndexView& ix, const ChromView& cv, const Pow3& p3, const MapConfig& cfg,
                        const ReadScratch& sc, uint32_t read_len, uint32_t seed_i, uint32_t strand, const LaneRun& mine,
                        uint32_t j, Sink& sink, Counters& ctr) {
  if (W::WIDTH == 32u && read_len <= WIDE_MAX_READ) {
    const uint32_t f0 = w.shfl(mine.f0, (int)j), hi = w.shfl(mine.hi, (int)j), fp_lo = w.shfl(mine.fp_lo, (int)j);
    const uint32_t fp_span = lookup_fp_span(ix, read_len, p3);
    bool longer = false;   // does the run reach past b slots?  (the narrowed region is a subset of the run)
    if ((uint64_t)f0 + cfg.b < hi) longer = ix.entries[f0 + cfg.b].fp - fp_lo <= fp_span;
    if (!longer) {
      verify_run_wide(w, ix, cv, sc, read_len, seed_i, strand, f0, hi, true, fp_lo, fp_span, sink, ctr);
      return;
    }
  }
  replay_lookup(w, ix, cv, p3, cfg, sc, read_len, seed_i, strand, sink, ctr);
}

// The long runs of a read are streamed one after the other (the fold is ordered), each a chain
// entries -> windows of dependent cache misses.  Before the fold, the first block of EVERY run is
// started: its entries are loaded by the 32 lanes and the lines of their windows prefetched, so
// that the six chains overlap and the ordered pass finds its first blocks on chip.
template <class W>
WALT_HD void prefetch_runs(W& w, const SubIndexView* ix2, uint32_t run_mask, const LaneRun& mine) {
#if defined(__CUDA_ARCH__)
  if (W::WIDTH != 32u) return;
  const uint32_t lane = w.lane();
  uint32_t e[LOOKUP_LANES];
  WALT_UNROLL
  for (uint32_t j = 0; j < LOOKUP_LANES; ++j) {
    e[j] = 0xFFFFFFFFu;
    if ((run_mask >> j) & 1u) {
      const uint32_t slot = w.shfl(mine.f0, (int)j) + lane, hi = w.shfl(mine.hi, (int)j);
      if (slot < hi) e[j] = ix2[j / 3u].entries[slot].pos;
    }
  }
  WALT_UNROLL
  for (uint32_t j = 0; j < LOOKUP_LANES; ++j)
    if (e[j] != 0xFFFFFFFFu)
      WALT_PREFETCH(reinterpret_cast<const char*>(ix2[j / 3u].genome) + ((((uint64_t)e[j] + PAD_BASES - j % 3u) >> 6) << 4));
#else
  (void)w; (void)ix2; (void)run_mask; (void)mine;
#endif
}

// ------------------------------------------------------------------------------------------
// parked reads: what the lane phase found, flat verification of their runs, ordered fold
// ------------------------------------------------------------------------------------------
// A read that a sub-warp group cannot finish (its fold reaches a long run) is PARKED: its lanes
// write down what they found -- per lookup either the verified candidates or the run (first slot,
// length) -- and reserve one 32-candidate block after the other of a launch-wide scratch for the
// runs.  A second kernel then verifies ALL blocks of ALL parked reads, a warp per block, with no
// order and no state (verify_kernel, walt_engine.cu: eight windows per copy instruction, thousands of
// blocks in flight), leaving one byte per slot: the mismatch count, or MM_NONE.  A third kernel folds each
// parked read in reference order from its record and those bytes (fold_parked_*), touching the
// index only for the few candidates a sink still takes.
constexpr uint32_t PARK_WORDS = 64;          // words per parked read: [0] read, [1] blocks, [2 + 10 j ..) lookup j
constexpr uint32_t PARK_LOOKUP_WORDS = 10;   // [0] kind | n << 8, then the data of the kind
constexpr uint32_t PK_NONE = 0u, PK_CANDS = 1u, PK_RUN = 2u, PK_GROUP = 3u;
constexpr uint32_t MM_NONE = 0xFFu;

// no parking: every read is finished where it is (CPU harness, MAP_ALL kernels)
struct NoPark {
  static constexpr bool ENABLED = false;
  template <class W> WALT_HD uint32_t* claim(W&, uint32_t, bool, uint32_t&, uint32_t&) { return nullptr; }
  WALT_HD uint32_t* block_desc(uint32_t) { return nullptr; }
};

// Length of the fingerprint run that starts at run.f0, by one lane: doubling steps, then bisection.
// Returns b + 1 as soon as the run is known to be longer than b.
WALT_HD uint32_t run_length(const SubIndexView& ix, const LaneRun& run, uint32_t fp_span, uint32_t b) {
  const uint64_t limit64 = (uint64_t)run.f0 + b + 1u;
  const uint32_t limit = limit64 < run.hi ? (uint32_t)limit64 : run.hi;   // slots [f0, limit) matter
  auto in_run = [&](uint32_t slot) { return ix.entries[slot].fp - run.fp_lo <= fp_span; };
  uint32_t good = run.f0 + LANE_RUN_CAP;      // slots up to here are known to match (lane_lookup)
  uint32_t step = LANE_RUN_CAP;
  uint32_t bad = limit;                       // first slot known not to match (or the limit)
  while (good + step < limit) {
    if (in_run(good + step)) { good += step; step *= 2u; } else { bad = good + step; break; }
  }
  // invariant: good matches, bad does not (or is the limit): first non-matching slot in (good, bad]
  uint32_t lo = good + 1u, hi = bad;
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (in_run(mid)) lo = mid + 1u; else hi = mid;
  }
  return lo - run.f0;
}

// The narrowed region inside a fingerprint run longer than -b, by one lane: the run is sorted by the
// seed characters, so the slots whose seed characters all equal the read's are a stretch [first,
// first + n) of it, found by two bisections (compare_seed).  The run's end comes from bisection too.
WALT_HD_NOINLINE uint32_t exact_region(const SubIndexView& ix, const LaneRun& run, uint32_t fp_span, const ReadScratch& sc,
                                       uint32_t read_len, uint32_t seed_i, uint32_t& first) {
  const uint32_t nws = (3u * seed_repeats(read_len) + seed_i + 31u) >> 5;
  const uint64_t* SM = sc.SM + seed_i * sc.nw;
  uint32_t lo = run.f0, hi = run.hi;     // end of the run: first slot of [f0, hi) whose fingerprint is beyond
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (ix.entries[mid].fp - run.fp_lo <= fp_span) lo = mid + 1u; else hi = mid;
  }
  const uint32_t end = lo;
  auto cmp = [&](uint32_t slot) { return compare_seed(ix.genome, (uint64_t)ix.entries[slot].pos + PAD_BASES - seed_i, sc.R, SM, nws); };
  lo = run.f0; hi = end;                 // first slot that is not smaller than the read
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (cmp(mid) < 0) lo = mid + 1u; else hi = mid;
  }
  first = lo;
  hi = end;                              // first slot that is larger
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (cmp(mid) <= 0) lo = mid + 1u; else hi = mid;
  }
  return lo - first;
}

// A parked run in the ordered fold: its bytes, 32 at a time; the entry is read only for the slots a
// sink still takes.
template <class W, class Sink>
WALT_HD void fold_run(W& w, const SubIndexView& ix, const ChromView& cv, uint32_t read_len, uint32_t seed_i,
                      uint32_t strand, uint32_t f0, uint32_t count, const uint8_t* bytes, Sink& sink) {
  const uint32_t lane = w.lane();
  for (uint32_t base = 0; base < count; base += 32u) {
    const uint32_t mm = base + lane < count ? bytes[base + lane] : MM_NONE;
    bool valid = mm != MM_NONE && sink.may_take(mm);
    uint32_t g = 0u;
    if (valid) {   // bounds, mapping.cpp:281-286
      const uint32_t e = ix.entries[f0 + base + lane].pos;
      const uint32_t chr = chrom_of(cv.starts, cv.n_chr, e);
      g = e - seed_i;
      valid = (e - cv.starts[chr] >= seed_i) && !(g + read_len >= cv.starts[chr + 1u]);
    }
    sink.consume(w, valid, mm, g, strand);
  }
}

// The lanes of a group write the record of their read (see above).  `res`, `run`: what the lane's
// lookup came to; write_cands(rec) stores a lane's verified candidates and returns their number.
// Reads that the flat path cannot take (longer than WIDE_MAX_READ, no room left) go on the list of
// the warp-per-read kernel instead (park.claim sees to that).
template <class W, class Park, class WriteCands>
WALT_HD void park_read(W& w, const SubIndexView* ix2, const Pow3& p3, const MapConfig& cfg, const ReadScratch& sc,
                       uint32_t read_len, LaneResult res, const LaneRun& run, Park& park, WriteCands write_cands) {
  const uint32_t lane = w.lane();
  const bool legacy = read_len > WIDE_MAX_READ;
  uint32_t kind = PK_NONE, count = 0u, first = run.f0;
  bool skip = false;
  if (lane < LOOKUP_LANES && !legacy) {
    if (res == LANE_RUN) {
      count = run_length(ix2[lane / 3u], run, lookup_fp_span(ix2[lane / 3u], read_len, p3), cfg.b);
      kind = PK_RUN;
      if (count > cfg.b) {
        // the run may exceed -b, so the size of the narrowed region (the seed-equal slots, a contiguous
        // stretch of the run) matters: this lane finds it by bisection on the seed characters
        first = run.f0;
        count = exact_region(ix2[lane / 3u], run, lookup_fp_span(ix2[lane / 3u], read_len, p3), sc, read_len, lane % 3u, first);
        if (count > cfg.b || count == 0u) { kind = PK_NONE; count = 0u; skip = true; }   // mapping.cpp:275-277: the lookup is dropped
      }
    } else if (res == LANE_GROUP) {
      kind = PK_GROUP;
    }
  }
  // a lookup that needs the cooperative replay (a literal region, a run that may exceed -b) sends the
  // whole read to the warp-per-read kernel: the fold kernels stay free of that code
  const bool whole = legacy || w.ballot(kind == PK_GROUP) != 0u;
  const uint32_t blocks = (count + 31u) >> 5;
  uint32_t before = 0u, total = 0u;
  for (uint32_t k = 0; k < LOOKUP_LANES; ++k) {
    const uint32_t bk = w.shfl(blocks, (int)k);
    if (k < lane) before += bk;
    total += bk;
  }
  uint32_t blk0 = 0u, t = 0u;
  uint32_t* rec = park.claim(w, total, whole, blk0, t);    // group-uniform; NULL: the read went on the other list
  if (rec) {
    if (lane == 0u) rec[1] = total;                        // rec[0] (the read's number) is claim()'s
    if (lane < LOOKUP_LANES) {
      uint32_t* mine = rec + 2u + PARK_LOOKUP_WORDS * lane;
      if (kind == PK_RUN) {
        mine[1] = first; mine[2] = count; mine[3] = blk0 + before;
        for (uint32_t k = 0; k < blocks; ++k) {
          uint32_t* d = park.block_desc(blk0 + before + k);
          d[0] = t; d[1] = lane | (k << 3);
        }
      } else if (kind == PK_NONE && !skip) {
        const uint32_t n = write_cands(mine);
        if (n) kind = PK_CANDS | (n << 8);
      }
      mine[0] = kind;
    }
  }
  w.sync();
}

// ... and the fold of a parked read from its record (cands(rec, strand): the verified candidates of a
// lookup go to the sink as in map_read_*).
template <class W, class Sink, class Cands>
WALT_HD void fold_parked(W& w, const SubIndexView* ix2, const ChromView& cv, uint32_t read_len, const uint32_t* rec,
                         const uint8_t* bytes, Sink& sink, Cands cands) {
  for (uint32_t j = 0; j < LOOKUP_LANES; ++j) {
    const uint32_t* mine = rec + 2u + PARK_LOOKUP_WORDS * j;
    const uint32_t kind = mine[0] & 0xFFu;
    if (kind == PK_NONE) continue;
    const uint32_t s = j >= 3u ? 1u : 0u, seed_i = j - 3u * s;
    const uint32_t strand = s ? '-' : '+';
    if (sink.stop_before_shift(seed_i)) continue;
    if (kind == PK_CANDS) cands(mine, strand);
    else if (kind == PK_RUN) fold_run(w, ix2[s], cv, read_len, seed_i, strand, mine[1], mine[2], bytes + 32u * (size_t)mine[3], sink);
  }
}

// ------------------------------------------------------------------------------------------
// whole reads
// ------------------------------------------------------------------------------------------
// What became of a read: mapped; holds a non-ACGT byte; or PARKed -- its ordered fold reached a
// lookup that needs the whole group (repeats, tainted buckets) and the caller asked for such reads
// to be handed to the warp-per-read kernel instead (nothing of the read's result is valid then).
enum MapStatus : uint32_t { MAP_OK = 0u, MAP_BAD = 1u, MAP_PARKED = 2u, MAP_PARKED_LIT = 3u };   // _LIT: some lookup of the read is a literal one

// SingleEndMapping for both strand passes of one read (mapping.cpp:486-500 order: all shifts
// on the '+' sub-index, then all shifts on the '-' sub-index, state carried across).
template <class W, bool PACKED = false, class Park = NoPark>
WALT_HD MapStatus map_read_se(W& w, const SubIndexView* ix2, const ChromView& cv, const Pow3& p3,
                         const MapConfig& cfg, const char* seq, uint32_t read_len, bool ag,
                         uint32_t max_mismatches, ReadScratch& sc, uint32_t& cached_len,
                         BestState& out, Counters& ctr, Park park = Park()) {
  constexpr bool PARK = Park::ENABLED;
  static_assert(W::WIDTH >= LOOKUP_LANES, "a group needs one lane per lookup");
  BestSink<W> sink;
  sink.st.pos = 0u; sink.st.times = 0u; sink.st.mm = max_mismatches; sink.st.strand = '+';
  out = sink.st;
  const Counters ctr_in = ctr;
  if (read_len < MIN_READ_LEN) return MAP_OK;
  if (!load_read<W, PACKED>(w, seq, read_len, ag, sc)) return MAP_BAD;
  if (cached_len != read_len) { build_masks(w, read_len, sc); cached_len = read_len; }
  // every lookup lane runs its own lookup and summarises it: minimum, how many candidates
  // share it, the first and the last of them (all the ordered fold needs, see BestSink::apply)
  const uint32_t lane = w.lane();
  LaneResult res = LANE_DONE;
  LaneRun run; run.f0 = run.hi = run.fp_lo = 0u;
  uint32_t mn = NO_HIT, cnt = 0u, g_first = 0u, g_last = 0u;
  if (lane < LOOKUP_LANES) {
    if (cfg.literal_all) {
      res = LANE_GROUP;
    } else {
      res = lane_lookup(ix2[lane / 3u], cv, p3, cfg, sc, read_len, lane % 3u,
                        [&](uint32_t g, uint32_t mmc) {
                          if (mmc < mn) { mn = mmc; cnt = 1u; g_first = g_last = g; }
                          else if (mmc == mn) { ++cnt; g_last = g; }
                        },
                        [&]() { mn = NO_HIT; cnt = 0u; }, ctr, run);
    }
  }
  const uint32_t group_mask = w.ballot(res == LANE_GROUP);
  const uint32_t run_mask = w.ballot(res == LANE_RUN);
  const uint32_t hit_mask = w.ballot(mn != NO_HIT);
  // ordered fold over the lookups that can change the state.  Skipping a lookup where the
  // reference would have left the shift loop (mapping.cpp:250-256) is the same as its break: the
  // state is untouched by a skip, so the later shifts of that strand are skipped as well.
  if (!PARK && run_mask && read_len <= WIDE_MAX_READ) prefetch_runs(w, ix2, run_mask, run);
  uint32_t todo = (group_mask | run_mask | hit_mask) & ((1u << LOOKUP_LANES) - 1u);
  while (todo) {
    const uint32_t j = (uint32_t)ffs32(todo) - 1u;
    todo &= todo - 1u;
    const uint32_t s = j >= 3u ? 1u : 0u, seed_i = j - 3u * s;
    const uint32_t strand = s ? '-' : '+';
    if (sink.stop_before_shift(seed_i)) continue;
    if (((group_mask | run_mask) >> j) & 1u) {
      if (PARK) {   // the kernels that take the read over count its work
        ctr = ctr_in;
        park_read(w, ix2, p3, cfg, sc, read_len, res, run, park, [&](uint32_t* rec) {
          if (mn == NO_HIT) return 0u;
          rec[1] = mn; rec[2] = cnt; rec[3] = g_first; rec[4] = g_last;
          return 1u;
        });
        return MAP_PARKED;
      }
      if ((run_mask >> j) & 1u) run_lookup(w, ix2[s], cv, p3, cfg, sc, read_len, seed_i, strand, run, j, sink, ctr);
      else replay_lookup(w, ix2[s], cv, p3, cfg, sc, read_len, seed_i, strand, sink, ctr);
    } else {
      sink.apply(w.shfl(mn, (int)j), w.shfl(cnt, (int)j), w.shfl(g_first, (int)j), w.shfl(g_last, (int)j), strand);
    }
  }
  out = sink.st;
  w.sync();  // scratch is reused by the next read
  return MAP_OK;
}

// PairEndMapping for both strand passes of one mate (paired.cpp:650-671) into `sink` (reset by the
// caller; its state persists across the two passes).
template <class W, bool PACKED, class Park, class SinkT>
WALT_HD MapStatus map_read_pe_into(W& w, const SubIndexView* ix2, const ChromView& cv, const Pow3& p3,
                              const MapConfig& cfg, const char* seq, uint32_t read_len, bool ag,
                              uint32_t max_mismatches, ReadScratch& sc, uint32_t& cached_len, SinkT& sink,
                              Counters& ctr, Park park = Park()) {
  constexpr bool PARK = Park::ENABLED;
  static_assert(W::WIDTH >= LOOKUP_LANES, "a group needs one lane per lookup");
  const Counters ctr_in = ctr;
  if (read_len < MIN_READ_LEN) return MAP_OK;
  if (!load_read<W, PACKED>(w, seq, read_len, ag, sc)) return MAP_BAD;
  if (cached_len != read_len) { build_masks(w, read_len, sc); cached_len = read_len; }
  // lookup lanes leave their candidates (index order) in the group's scratch; they reach the
  // sink in reference order
  const uint32_t lane = w.lane();
  LaneResult res = LANE_DONE;
  LaneRun run; run.f0 = run.hi = run.fp_lo = 0u;
  uint32_t n_mine = 0u;
  if (lane < LOOKUP_LANES) {
    if (cfg.literal_all) {
      res = LANE_GROUP;
    } else {
      LaneCand* mine = sc.C + lane * LANE_RUN_CAP;
      res = lane_lookup(ix2[lane / 3u], cv, p3, cfg, sc, read_len, lane % 3u,
                        [&](uint32_t g, uint32_t mmc) {
                          if (mmc > max_mismatches) return;          // paired.cpp:191-193
                          LaneCand c; c.g = g; c.mm = mmc;
                          mine[n_mine] = c;
                          ++n_mine;
                        },
                        [&]() { n_mine = 0u; }, ctr, run);
    }
  }
  const uint32_t group_mask = w.ballot(res == LANE_GROUP);
  const uint32_t run_mask = w.ballot(res == LANE_RUN);
  const uint32_t hit_mask = w.ballot(n_mine != 0u);
  w.sync();
  if (!PARK && run_mask && read_len <= WIDE_MAX_READ) prefetch_runs(w, ix2, run_mask, run);
  uint32_t todo = (group_mask | run_mask | hit_mask) & ((1u << LOOKUP_LANES) - 1u);
  while (todo) {   // see map_read_se: a skip is the reference's break (paired.cpp:127-137)
    const uint32_t j = (uint32_t)ffs32(todo) - 1u;
    todo &= todo - 1u;
    const uint32_t s = j >= 3u ? 1u : 0u, seed_i = j - 3u * s;
    const uint32_t strand = s ? '-' : '+';
    if (sink.stop_before_shift(seed_i)) continue;
    if (((group_mask | run_mask) >> j) & 1u) {
      if (PARK) {
        ctr = ctr_in;
        park_read(w, ix2, p3, cfg, sc, read_len, res, run, park, [&](uint32_t* rec) {
          const LaneCand* c = sc.C + lane * LANE_RUN_CAP;
          for (uint32_t k = 0; k < n_mine; ++k) { rec[1u + 2u * k] = c[k].g; rec[2u + 2u * k] = c[k].mm; }
          return n_mine;
        });
        return MAP_PARKED;
      }
      if ((run_mask >> j) & 1u) run_lookup(w, ix2[s], cv, p3, cfg, sc, read_len, seed_i, strand, run, j, sink, ctr);
      else replay_lookup(w, ix2[s], cv, p3, cfg, sc, read_len, seed_i, strand, sink, ctr);
    } else {
      const uint32_t nj = w.shfl(n_mine, (int)j);
      sink.push_list(w, sc.C + j * LANE_RUN_CAP, nj, strand);
    }
  }
  w.sync();
  return MAP_OK;
}

// ... with the heap maintained in place by lane 0.  On return heap[0..size) is the libstdc++ heap array.
template <class W, bool PACKED = false>
WALT_HD MapStatus map_read_pe(W& w, const SubIndexView* ix2, const ChromView& cv, const Pow3& p3,
                         const MapConfig& cfg, const char* seq, uint32_t read_len, bool ag,
                         uint32_t max_mismatches, uint32_t top_k, ReadScratch& sc,
                         uint32_t& cached_len, HeapEntry* heap, uint32_t& heap_size, Counters& ctr) {
  HeapSink<W> sink;
  sink.heap = heap; sink.size = 0u; sink.cap = top_k; sink.top_mm = 0u; sink.max_mm = max_mismatches;
  heap_size = 0u;
  const MapStatus st = map_read_pe_into<W, PACKED, NoPark>(w, ix2, cv, p3, cfg, seq, read_len, ag, max_mismatches, sc, cached_len, sink, ctr);
  heap_size = sink.size;
  return st;
}

// ... logging the heap-changing candidates for replay_heap_log (max_mismatches <= LOG_MAX_MM; `log`
// holds pe_log_slots(top_k, max_mismatches) entries, `hist` max_mismatches + 1 counters).
template <class W, bool PACKED = false, class Park = NoPark>
WALT_HD MapStatus map_read_pe_logged(W& w, const SubIndexView* ix2, const ChromView& cv, const Pow3& p3,
                                const MapConfig& cfg, const char* seq, uint32_t read_len, bool ag,
                                uint32_t max_mismatches, uint32_t top_k, ReadScratch& sc,
                                uint32_t& cached_len, HeapEntry* log, uint32_t* hist, uint32_t& n_log, Counters& ctr,
                                Park park = Park()) {
  LogSink<W> sink;
  sink.log = log; sink.hist = hist; sink.cap = top_k; sink.max_mm = max_mismatches;
  sink.reset(w);
  const MapStatus st = map_read_pe_into<W, PACKED, Park>(w, ix2, cv, p3, cfg, seq, read_len, ag, max_mismatches, sc, cached_len, sink, ctr, park);
  n_log = sink.n_log;
  return st;
}

// ------------------------------------------------------------------------------------------
// pairing (MergePairedEndResults loop, paired.cpp:472-513) -- one lane per pair
// ------------------------------------------------------------------------------------------
struct RankedCand { uint32_t pos; uint32_t mm; uint32_t strand; };

WALT_HD void forward_position(const uint32_t* __restrict__ starts, uint32_t pos, uint32_t strand,
                              uint32_t chr, uint32_t read_len, uint32_t& s, uint32_t& e) {
  // ForwardChromPosition, paired.cpp:98-104 (length[chr] = starts[chr+1]-starts[chr])
  uint32_t v = pos - starts[chr];
  if (strand != '+') v = (starts[chr + 1u] - starts[chr]) - v - read_len;
  s = v; e = v + read_len;
}

struct PairResult { uint32_t best_times; int32_t best_i, best_j; int32_t frag; };

// The chromosome and forward start of every mate-2 candidate are pure functions of the candidate;
// the reference recomputes them inside the O(k^2) loop (paired.cpp:486-499), here they are
// computed once per candidate when the list is short enough for a local array (top_k <= 64).
constexpr uint32_t PAIR_PRECOMPUTE = 64;

template <class GetCand>
WALT_HD PairResult pair_candidates(const ChromView& cv, GetCand get1, uint32_t n1, uint32_t len1,
                                   GetCand get2, uint32_t n2, uint32_t len2,
                                   uint32_t max_mismatches, int32_t frag_range) {
  PairResult r; r.best_times = 0u; r.best_i = -1; r.best_j = -1; r.frag = 0;
  uint32_t min_mm = max_mismatches;
  uint64_t best_pos = 0;
  uint32_t chr2[PAIR_PRECOMPUTE], start2[PAIR_PRECOMPUTE];
  const bool pre = n2 <= PAIR_PRECOMPUTE && n1 > 1u;
  if (pre) {
    for (uint32_t j = 0; j < n2; ++j) {
      const RankedCand bb = get2(j);
      uint32_t s2, e2;
      chr2[j] = chrom_of(cv.starts, cv.n_chr, bb.pos);
      forward_position(cv.starts, bb.pos, bb.strand, chr2[j], len2, s2, e2);
      start2[j] = s2;
    }
  }
  for (int32_t i = (int32_t)n1 - 1; i >= 0; --i) {
    const RankedCand a = get1((uint32_t)i);
    const uint32_t c1 = chrom_of(cv.starts, cv.n_chr, a.pos);
    uint32_t s1, e1;
    forward_position(cv.starts, a.pos, a.strand, c1, len1, s1, e1);
    for (int32_t j = (int32_t)n2 - 1; j >= 0; --j) {
      const RankedCand bb = get2((uint32_t)j);
      if (a.strand == bb.strand) continue;
      const uint32_t sum = a.mm + bb.mm;
      if (sum > min_mm) break;
      const uint32_t c2 = pre ? chr2[j] : chrom_of(cv.starts, cv.n_chr, bb.pos);
      if (c1 != c2) continue;
      uint32_t s2, e2;
      if (pre) { s2 = start2[j]; e2 = s2 + len2; }
      else forward_position(cv.starts, bb.pos, bb.strand, c2, len2, s2, e2);
      const int32_t frag = a.strand == '+' ? (int32_t)(e2 - s1) : (int32_t)(e1 - s2);
      if (frag <= 0 || frag > frag_range) continue;
      const uint64_t cur = ((uint64_t)a.pos << 32) + bb.pos;
      if (sum < min_mm) {
        r.best_i = i; r.best_j = j; r.best_times = 1u; min_mm = sum; best_pos = cur; r.frag = frag;
      } else if (sum == min_mm && cur != best_pos) {
        r.best_i = i; r.best_j = j; r.best_times++; r.frag = frag;
      }
    }
  }
  return r;
}

// The same pairing by a whole warp, for pairs whose k1 x k2 loop is long (both mates in repeats): a
// thread-per-pair loop leaves the other 31 lanes of the warp idle for thousands of iterations.
// The reference's loop is order dependent only through three facts, each a reduction over the
// VALID pairs (opposite strands, same chromosome, 0 < fragment <= L, mismatch sum <= m; the
// loop's `break` only skips pairs whose sum exceeds the running minimum):
//   m*    the smallest mismatch sum;
//   F     the first valid pair with sum m* in loop order -- its positions become best_pos, unless
//         m* == m: then no `<` ever fired and best_pos stays 0 (paired.cpp:502-511);
//   count / L   the valid pairs with sum m* whose positions differ from best_pos, and the last of
//         them in loop order: best_times = count (+ 1 for F if m* < m), the reported pair is L
//         (or F if nothing was counted).
// Lanes own mate-1 candidates; mate 2's chromosome / forward start / strand / mismatches sit in
// `sm` (3 words per candidate, owned by the warp).  Needs W::WIDTH == 32, max_mismatches <= 127.
constexpr uint32_t PAIR_WIDE_MIN = 256;   // k1 * k2 above which the warp takes the pair
template <class W, class GetCand>
WALT_HD PairResult pair_candidates_wide(W& w, const ChromView& cv, GetCand get1, uint32_t n1, uint32_t len1,
                                        GetCand get2, uint32_t n2, uint32_t len2, uint32_t max_mismatches,
                                        int32_t frag_range, uint32_t* sm) {
  PairResult r; r.best_times = 0u; r.best_i = -1; r.best_j = -1; r.frag = 0;
  const uint32_t lane = w.lane();
  uint32_t* start2 = sm; uint32_t* key2 = sm + n2; uint32_t* pos2 = sm + 2u * n2;
  for (uint32_t j = lane; j < n2; j += 32u) {
    const RankedCand bb = get2(j);
    const uint32_t c2 = chrom_of(cv.starts, cv.n_chr, bb.pos);
    uint32_t s2, e2;
    forward_position(cv.starts, bb.pos, bb.strand, c2, len2, s2, e2);
    start2[j] = s2; key2[j] = (c2 << 8) | (bb.strand == '+' ? 0x80u : 0u) | bb.mm; pos2[j] = bb.pos;
  }
  w.sync();
  const uint32_t rounds = (n1 + 31u) / 32u;
  // visit(i, j, sum, frag) for every valid pair of this lane's candidates with sum <= limit (== limit if exact)
  auto scan = [&](uint32_t limit, bool exact, auto visit) {
    for (uint32_t k = 0; k < rounds; ++k) {
      const uint32_t i = 32u * k + lane;
      if (i >= n1) continue;
      const RankedCand a = get1(i);
      if (a.mm > limit) continue;
      const uint32_t c1 = chrom_of(cv.starts, cv.n_chr, a.pos);
      uint32_t s1, e1;
      forward_position(cv.starts, a.pos, a.strand, c1, len1, s1, e1);
      const uint32_t want = (c1 << 8) | (a.strand == '+' ? 0u : 0x80u);   // same chromosome, the other strand
      for (int32_t j = (int32_t)n2 - 1; j >= 0; --j) {
        const uint32_t kj = key2[j];
        const uint32_t sum = a.mm + (kj & 0x7Fu);
        if (sum > limit) break;                       // mismatches only grow towards slot 0
        if ((kj & ~0x7Fu) != want || (exact && sum != limit)) continue;
        const uint32_t s2 = start2[j];
        const int32_t frag = a.strand == '+' ? (int32_t)(s2 + len2 - s1) : (int32_t)(e1 - s2);
        if (frag <= 0 || frag > frag_range) continue;
        visit(i, (uint32_t)j, sum, frag);
      }
    }
  };
  uint32_t mn = NO_HIT;
  scan(max_mismatches, false, [&](uint32_t, uint32_t, uint32_t sum, int32_t) { mn = sum < mn ? sum : mn; });
  const uint32_t best = w.reduce_min(mn);
  if (best == NO_HIT) { w.sync(); return r; }
  uint32_t first = NO_HIT;
  scan(best, true, [&](uint32_t i, uint32_t j, uint32_t, int32_t) {
    const uint32_t order = (n1 - 1u - i) * n2 + (n2 - 1u - j);
    first = order < first ? order : first;
  });
  const uint32_t F = w.reduce_min(first);
  uint64_t best_pos = 0;
  if (best < max_mismatches) best_pos = ((uint64_t)get1(n1 - 1u - F / n2).pos << 32) + pos2[n2 - 1u - F % n2];
  uint32_t cnt = 0u, last_inv = NO_HIT;   // last_inv = ~(largest counted order)
  scan(best, true, [&](uint32_t i, uint32_t j, uint32_t, int32_t) {
    if ((((uint64_t)get1(i).pos << 32) + pos2[j]) == best_pos) return;
    const uint32_t order = (n1 - 1u - i) * n2 + (n2 - 1u - j);
    ++cnt;
    last_inv = ~order < last_inv ? ~order : last_inv;
  });
  const uint32_t total = w.reduce_add(cnt);
  const uint32_t L = ~w.reduce_min(last_inv);
  w.sync();   // `sm` is reused by the warp's next pair
  r.best_times = total + (best < max_mismatches ? 1u : 0u);
  if (r.best_times == 0u) return r;
  const uint32_t win = total ? L : F;
  r.best_i = (int32_t)(n1 - 1u - win / n2); r.best_j = (int32_t)(n2 - 1u - win % n2);
  {  // fragment length of the reported pair (GetFragmentLength, paired.cpp:320-331)
    const RankedCand a = get1((uint32_t)r.best_i), bb = get2((uint32_t)r.best_j);
    const uint32_t c1 = chrom_of(cv.starts, cv.n_chr, a.pos), c2 = chrom_of(cv.starts, cv.n_chr, bb.pos);
    uint32_t s1, e1, s2, e2;
    forward_position(cv.starts, a.pos, a.strand, c1, len1, s1, e1);
    forward_position(cv.starts, bb.pos, bb.strand, c2, len2, s2, e2);
    r.frag = a.strand == '+' ? (int32_t)(e2 - s1) : (int32_t)(e1 - s2);
  }
  return r;
}

// GetBestMatch4Single (paired.cpp:296-318): the single-end style fold over one mate's ranked
// list, best (last slot) to worst, stopping at the first worse candidate.
template <class GetCand>
WALT_HD BestState single_best(GetCand get, uint32_t n, uint32_t max_mismatches) {
  BestState b; b.pos = 0u; b.times = 0u; b.mm = max_mismatches; b.strand = '+';
  for (int32_t i = (int32_t)n - 1; i >= 0; --i) {
    const RankedCand c = get((uint32_t)i);
    if (c.mm < b.mm) {
      b.pos = c.pos; b.times = 1u; b.strand = c.strand; b.mm = c.mm;
    } else if (c.mm == b.mm) {
      if (b.pos == c.pos) continue;
      b.pos = c.pos; b.strand = c.strand; b.times++;
    } else {
      break;
    }
  }
  return b;
}

}  // namespace waltcore
