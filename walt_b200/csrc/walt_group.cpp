// walt_group.cpp -- several GPUs behind one handle of the C ABI (include/walt_b200.h, "several GPUs").
//
// Reads shard naturally: every device holds a full index replica (read from the files
// once, replicated device to device), a batch is cut into contiguous ranges, one host thread drives
// each engine and every range lands in its slice of the caller's result array.  No collective.
#include <string>
#include <thread>
#include <vector>

#include "../../include/walt_b200.h"

namespace waltb200 {
int fail(int code, const std::string& msg);
}
using waltb200::fail;

struct walt_group {
  std::vector<walt_engine*> e;
};

namespace {
// contiguous split of [0, n) over g workers
inline uint32_t cut(uint32_t n, uint32_t g, uint32_t i) { return (uint32_t)((uint64_t)n * i / g); }

// run fn(i, lo, hi) for every non-empty shard on its own thread; first error wins
template <class F>
int sharded(walt_group* g, uint32_t n, F fn) {
  const uint32_t k = (uint32_t)g->e.size();
  std::vector<int> rc(k, 0);
  std::vector<std::string> err(k);
  if (k == 1) {
    rc[0] = n ? fn(0u, 0u, n) : 0;
    return rc[0];
  }
  std::vector<std::thread> th;
  for (uint32_t i = 0; i < k; ++i)
    th.emplace_back([&, i]() {
      const uint32_t lo = cut(n, k, i), hi = cut(n, k, i + 1);
      if (hi == lo) return;
      rc[i] = fn(i, lo, hi);
      if (rc[i]) err[i] = walt_last_error();     // the error text is per thread
    });
  for (auto& t : th) t.join();
  for (uint32_t i = 0; i < k; ++i)
    if (rc[i]) return fail(rc[i], err[i]);
  return WALT_OK;
}
}  // namespace

extern "C" {

int walt_group_create(walt_group** out, const int* device_ids, int n_devices) {
  if (!out || n_devices < 1 || !device_ids) return fail(WALT_EINVAL, "bad argument");
  *out = nullptr;
  walt_group* g = new walt_group;
  g->e.assign((size_t)n_devices, nullptr);
  // one thread per device: a CUDA context takes about a second to come up, eight of them in a row eight
  std::vector<int> rcs((size_t)n_devices, 0);
  std::vector<std::string> err((size_t)n_devices);
  std::vector<std::thread> th;
  for (int i = 0; i < n_devices; ++i)
    th.emplace_back([&, i]() {
      rcs[i] = walt_engine_create(&g->e[i], device_ids[i]);
      if (rcs[i]) err[i] = walt_last_error();
    });
  for (auto& t : th) t.join();
  for (int i = 0; i < n_devices; ++i)
    if (rcs[i]) {
      const int rc = rcs[i];
      const std::string msg = err[i];
      walt_group_destroy(g);
      return fail(rc, msg);
    }
  *out = g;
  return WALT_OK;
}

void walt_group_destroy(walt_group* g) {
  if (!g) return;
  for (auto* e : g->e) walt_engine_destroy(e);
  delete g;
}

int walt_group_size(const walt_group* g) { return g ? (int)g->e.size() : 0; }
walt_engine* walt_group_engine(walt_group* g, int i) { return g && i >= 0 && i < (int)g->e.size() ? g->e[i] : nullptr; }

int walt_group_load_dbindex(walt_group* g, const char* dbindex_path, uint32_t which_mask) {
  if (!g || g->e.empty()) return fail(WALT_EINVAL, "bad argument");
  int rc = walt_engine_load_dbindex(g->e[0], dbindex_path, which_mask);
  if (rc) return rc;
  // every other device pulls its replica from the first one, all at once (NVSwitch: full bandwidth to every peer)
  const uint32_t k = (uint32_t)g->e.size();
  std::vector<int> rcs(k, 0);
  std::vector<std::string> err(k);
  std::vector<std::thread> th;
  for (uint32_t i = 1; i < k; ++i)
    th.emplace_back([&, i]() {
      rcs[i] = walt_engine_clone_index(g->e[i], g->e[0]);
      if (rcs[i]) err[i] = walt_last_error();
    });
  for (auto& t : th) t.join();
  for (uint32_t i = 1; i < k; ++i)
    if (rcs[i]) return fail(rcs[i], err[i]);
  return WALT_OK;
}

int walt_group_map_se(walt_group* g, const char* seqs, const uint64_t* offs, uint32_t n, int ag_wildcard,
                      uint32_t max_mismatches, uint32_t b, walt_best* out, uint32_t* n_short) {
  if (!g || g->e.empty() || !offs) return fail(WALT_EINVAL, "bad argument");
  std::vector<uint32_t> sh(g->e.size(), 0);
  // ASCII reads: read j starts at byte offs[j]; a shard passes its own slice of offs (absolute offsets)
  const int rc = sharded(g, n, [&](uint32_t i, uint32_t lo, uint32_t hi) {
    return walt_engine_map_se(g->e[i], seqs, offs + lo, hi - lo, ag_wildcard, max_mismatches, b, out + lo, &sh[i]);
  });
  if (n_short) { *n_short = 0; for (uint32_t v : sh) *n_short += v; }
  return rc;
}

int walt_group_map_se_packed(walt_group* g, const uint8_t* packed, const uint64_t* offs, uint32_t n, int ag_wildcard,
                             uint32_t max_mismatches, uint32_t b, walt_best* out, uint32_t* n_short) {
  if (!g || g->e.empty() || !offs) return fail(WALT_EINVAL, "bad argument");
  std::vector<uint32_t> sh(g->e.size(), 0);
  // read j of the batch lives at byte (offs[j] >> 2) + j of `packed`: a shard that starts at read lo passes
  // the buffer shifted by lo so that its own j = 0 lands on the same byte
  const int rc = sharded(g, n, [&](uint32_t i, uint32_t lo, uint32_t hi) {
    return walt_engine_map_se_packed(g->e[i], packed + lo, offs + lo, hi - lo, ag_wildcard, max_mismatches, b, out + lo, &sh[i]);
  });
  if (n_short) { *n_short = 0; for (uint32_t v : sh) *n_short += v; }
  return rc;
}

int walt_group_map_pe_compact(walt_group* g, const char* seqs1, const uint64_t* offs1, const char* seqs2,
                              const uint64_t* offs2, uint32_t n, uint32_t max_mismatches, uint32_t b, uint32_t top_k,
                              int frag_range, int pbat, walt_pe_result* out, uint32_t* n_short1, uint32_t* n_short2) {
  if (!g || g->e.empty() || !offs1 || !offs2) return fail(WALT_EINVAL, "bad argument");
  std::vector<uint32_t> s1(g->e.size(), 0), s2(g->e.size(), 0);
  const int rc = sharded(g, n, [&](uint32_t i, uint32_t lo, uint32_t hi) {
    return walt_engine_map_pe_compact(g->e[i], seqs1, offs1 + lo, seqs2, offs2 + lo, hi - lo, max_mismatches, b, top_k, frag_range,
                                      pbat, out + lo, &s1[i], &s2[i]);
  });
  if (n_short1) { *n_short1 = 0; for (uint32_t v : s1) *n_short1 += v; }
  if (n_short2) { *n_short2 = 0; for (uint32_t v : s2) *n_short2 += v; }
  return rc;
}

int walt_group_map_pe_compact_packed(walt_group* g, const uint8_t* packed1, const uint64_t* offs1, const uint8_t* packed2,
                                     const uint64_t* offs2, uint32_t n, uint32_t max_mismatches, uint32_t b, uint32_t top_k,
                                     int frag_range, int pbat, walt_pe_result* out, uint32_t* n_short1, uint32_t* n_short2) {
  if (!g || g->e.empty() || !offs1 || !offs2) return fail(WALT_EINVAL, "bad argument");
  std::vector<uint32_t> s1(g->e.size(), 0), s2(g->e.size(), 0);
  const int rc = sharded(g, n, [&](uint32_t i, uint32_t lo, uint32_t hi) {
    return walt_engine_map_pe_compact_packed(g->e[i], packed1 + lo, offs1 + lo, packed2 + lo, offs2 + lo, hi - lo, max_mismatches, b,
                                             top_k, frag_range, pbat, out + lo, &s1[i], &s2[i]);
  });
  if (n_short1) { *n_short1 = 0; for (uint32_t v : s1) *n_short1 += v; }
  if (n_short2) { *n_short2 = 0; for (uint32_t v : s2) *n_short2 += v; }
  return rc;
}

}  // extern "C"
