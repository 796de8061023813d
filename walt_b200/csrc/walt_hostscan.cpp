// walt_hostscan.cpp -- the one pass the host makes over a chunk's read offsets before its copies
// are queued: longest read (sizes the kernel's shared memory), reads below the 38-base minimum
// (StatSingleReads::num_of_short_reads, mapping.cpp:230-232) and the common length if all reads
// share one (then no offsets cross PCIe).  At a gigaread per second this loop is on the critical
// path of the batch pipeline, so it is an AVX2 loop where the CPU has AVX2 (runtime check) and a
// branch-free scalar loop elsewhere.  Plain C++ (no CUDA) so that the host compiler sees it.
#include "walt_hostscan.h"

#include <immintrin.h>

namespace waltb200 {

namespace {

constexpr uint32_t kMinReadLen = 38;   // MINIMALREADLEN (seedpattern.hpp:359)

struct Acc { uint64_t wide; uint32_t mx, ns, df; };

inline void scalar_range(const uint64_t* o, uint32_t lo, uint32_t hi, uint32_t first, Acc& a) {
  for (uint32_t i = lo; i < hi; ++i) {
    const uint64_t l = o[i + 1] - o[i];
    const uint32_t l32 = (uint32_t)l;
    a.wide |= l >> 32;
    a.mx = l32 > a.mx ? l32 : a.mx;
    a.ns += l32 < kMinReadLen ? 1u : 0u;
    a.df |= l32 ^ first;
  }
}

__attribute__((target("avx2"))) void avx2_range(const uint64_t* o, uint32_t n, uint32_t first, Acc& a) {
  __m256i all = _mm256_setzero_si256(), mx = _mm256_setzero_si256(), ns = _mm256_setzero_si256(), df = _mm256_setzero_si256();
  const __m256i f = _mm256_set1_epi64x((long long)first), thr = _mm256_set1_epi64x((long long)kMinReadLen);
  uint32_t i = 0;
  for (; i + 4 <= n; i += 4) {
    const __m256i lo = _mm256_loadu_si256((const __m256i*)(o + i));
    const __m256i hi = _mm256_loadu_si256((const __m256i*)(o + i + 1));
    const __m256i l = _mm256_sub_epi64(hi, lo);
    all = _mm256_or_si256(all, l);
    mx = _mm256_max_epu32(mx, l);                             // valid when no length needs 64 bits (checked below)
    ns = _mm256_sub_epi64(ns, _mm256_cmpgt_epi64(thr, l));    // mask is -1 per short read
    df = _mm256_or_si256(df, _mm256_xor_si256(l, f));
  }
  alignas(32) uint64_t t[4];
  _mm256_store_si256((__m256i*)t, all);
  a.wide |= (t[0] | t[1] | t[2] | t[3]) >> 32;
  _mm256_store_si256((__m256i*)t, mx);
  for (int k = 0; k < 4; ++k) a.mx = (uint32_t)t[k] > a.mx ? (uint32_t)t[k] : a.mx;
  _mm256_store_si256((__m256i*)t, ns);
  a.ns += (uint32_t)(t[0] + t[1] + t[2] + t[3]);
  _mm256_store_si256((__m256i*)t, df);
  a.df |= (uint32_t)(t[0] | t[1] | t[2] | t[3]) | (uint32_t)((t[0] | t[1] | t[2] | t[3]) >> 32);
  scalar_range(o, i, n, first, a);
}

}  // namespace

ChunkScan scan_chunk(const uint64_t* offs, uint32_t r0, uint32_t cn) {
  ChunkScan c{0u, 0u, 0u};
  if (cn == 0) return c;
  const uint64_t* o = offs + r0;
  const uint64_t first64 = o[1] - o[0];
  const uint32_t first = (uint32_t)first64;
  Acc a{first64 >> 32, 0u, 0u, 0u};
  static const bool have_avx2 = __builtin_cpu_supports("avx2");
  if (have_avx2) avx2_range(o, cn, first, a);
  else scalar_range(o, 0, cn, first, a);
  if (a.wide) { c.max_len = 0xFFFFFFFFu; return c; }
  c.max_len = a.mx;
  c.n_short = a.ns;
  if (a.df == 0u && first > 0u) c.uniform_len = first;
  return c;
}

}  // namespace waltb200
