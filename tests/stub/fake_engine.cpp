// fake_engine.cpp -- TEST DOUBLE, not a mapper.  It exports the few libwaltb200 entry points the `walt`
// program binds (include/walt_b200.h) and answers every read with values derived from a hash of its
// 2-bit bytes.  tests/test_host_pipeline.py puts it in front of the real library (LD_LIBRARY_PATH) to run
// the program's batch loop -- parts of a batch loaded, "mapped" and written side by side, the rand()
// stream of the N replacement carried from part to part -- on a machine without a GPU, where the only
// question is whether the outputs depend on how the work was cut.  Nothing outside tests/ may use it.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "../../include/walt_b200.h"

struct walt_group { int n; uint32_t genome_len; };

static uint64_t hash_read(const uint8_t* packed, const uint64_t* offs, uint32_t j, uint32_t* len_out) {
  const uint64_t len = offs[j + 1] - offs[j];
  const uint8_t* p = packed + (offs[j] >> 2) + j;
  uint64_t h = 1469598103934665603ull ^ len;
  for (uint64_t i = 0; i < (len + 3) / 4; ++i) { h ^= p[i]; h *= 1099511628211ull; }
  h ^= h >> 29; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 32;
  *len_out = (uint32_t)len;
  return h;
}

static walt_best fake_best(uint64_t h, uint32_t len, uint32_t m, uint32_t genome_len) {
  walt_best r;
  memset(&r, 0, sizeof r);
  r.strand = '+'; r.mismatch = m;
  if (len < 38) return r;                                   // too short: unmapped (and counted by the caller)
  const uint32_t k = (uint32_t)(h & 15u);
  r.times = k == 0 ? 0u : k < 3 ? 2u + k : 1u;
  if (r.times == 0) return r;
  r.genome_pos = (uint32_t)((h >> 8) % (genome_len - 400u));
  r.mismatch = (uint32_t)((h >> 44) % (m + 1u));
  r.strand = ((h >> 60) & 1u) ? '+' : '-';
  return r;
}

extern "C" {

const char* walt_last_error(void) { return "fake engine"; }
int walt_device_count(void) { return 1; }

int walt_group_create(walt_group** out, const int*, int n_devices) {
  walt_group* g = new walt_group;
  g->n = n_devices;
  const char* v = getenv("WALT_FAKE_GENOME_LEN");
  g->genome_len = v ? (uint32_t)strtoul(v, nullptr, 10) : 1000000u;
  *out = g;
  return WALT_OK;
}
void walt_group_destroy(walt_group* g) { delete g; }

int walt_group_load_dbindex(walt_group*, const char*, uint32_t) {
  const char* v = getenv("WALT_FAKE_START_MS");             // an engine that takes a while to start
  if (v) usleep((useconds_t)strtoul(v, nullptr, 10) * 1000u);
  return WALT_OK;
}

int walt_group_map_se_packed(walt_group* g, const uint8_t* packed, const uint64_t* offs, uint32_t n, int, uint32_t m, uint32_t,
                             walt_best* out, uint32_t* n_short) {
  uint32_t s = 0;
  for (uint32_t j = 0; j < n; ++j) {
    uint32_t len;
    const uint64_t h = hash_read(packed, offs, j, &len);
    if (len < 38) ++s;
    out[j] = fake_best(h, len, m, g->genome_len);
  }
  if (n_short) *n_short = s;
  return WALT_OK;
}

int walt_group_map_pe_compact_packed(walt_group* g, const uint8_t* packed1, const uint64_t* offs1, const uint8_t* packed2,
                                     const uint64_t* offs2, uint32_t n, uint32_t m, uint32_t, uint32_t, int frag_range, int pbat,
                                     walt_pe_result* out, uint32_t* n_short1, uint32_t* n_short2) {
  uint32_t s1 = 0, s2 = 0;
  for (uint32_t j = 0; j < n; ++j) {
    uint32_t l1, l2;
    const uint64_t h1 = hash_read(packed1, offs1, j, &l1), h2 = hash_read(packed2, offs2, j, &l2);
    if (l1 < 38) ++s1;
    if (l2 < 38) ++s2;
    walt_pe_result r;
    memset(&r, 0, sizeof r);
    r.single1 = fake_best(h1, l1, m, g->genome_len);
    r.single2 = fake_best(h2 ^ 0x9e3779b97f4a7c15ull, l2, m, g->genome_len);
    const uint32_t k = (uint32_t)((h1 ^ h2) >> 17) & 7u;
    r.pair.best_i = r.pair.best_j = -1;
    if (l1 >= 38 && l2 >= 38 && k < 6) {
      r.pair.best_times = k < 4 ? 1u : 2u + k;
      r.pair.best_i = (int32_t)(h1 % 5u); r.pair.best_j = (int32_t)(h2 % 5u);
      const uint32_t pos = (uint32_t)((h1 >> 8) % (g->genome_len - 2000u));
      const uint32_t gap = (uint32_t)((h2 >> 8) % (uint32_t)(frag_range > 400 ? frag_range - 300 : 100));
      // the C->T mate (the second file's under -P) lies upstream, so that the writer sees a fragment it can lay out
      walt_cand& up = pbat ? r.c2 : r.c1;
      walt_cand& down = pbat ? r.c1 : r.c2;
      up.genome_pos = pos; up.mismatch = (uint32_t)((h1 >> 44) % (m + 1u)); up.strand = '+';
      down.genome_pos = pos + gap; down.mismatch = (uint32_t)((h2 >> 44) % (m + 1u)); down.strand = '+';
      r.pair.frag_len = (int32_t)(gap + (pbat ? l1 : l2));
    }
    out[j] = r;
  }
  if (n_short1) *n_short1 = s1;
  if (n_short2) *n_short2 = s2;
  return WALT_OK;
}

}  // extern "C"
