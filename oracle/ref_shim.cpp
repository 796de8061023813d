/* oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * A thin C ABI over the UNMODIFIED reference object code (compiled by oracle/Makefile from
 * /root/reference/src/walt/{reference,mapping,paired}.cpp into oracle/_ref/libwaltref.so).
 * Nothing of the algorithm lives here: every function below only marshals arrays into the
 * reference's own `Genome` / `HashTable` / `BestMatch` / `TopCandidates` types
 * (src/walt/reference.hpp:44-92, src/walt/mapping.hpp:39-52, src/walt/paired.hpp:35-74) and
 * calls the reference's own functions in the same order as its batch drivers
 * (src/walt/mapping.cpp:486-500, src/walt/paired.cpp:650-671).
 *
 * Used by tests/ (as the checker) and by bench.py's cpu_baseline / --impl reference leg
 * (as the thing timed on the host cores).  Never loaded by the product path.
 */
#include "paired.hpp"
#include "mapping.hpp"
#include "reference.hpp"

#include <omp.h>
#include <cstring>
#include <string>
#include <vector>

/* defined in src/walt/mapping.cpp:224 / src/walt/paired.cpp:106 (no header entries for the
 * per-read functions) */
void SingleEndMapping(const std::string& org_read, const Genome& genome,
                      const HashTable& hash_table, const char& strand,
                      const bool& AG_WILDCARD, const uint32_t& b,
                      BestMatch& best_match, StatSingleReads& stat_single_reads);
void PairEndMapping(const std::string& org_read, const Genome& genome,
                    const HashTable& hash_table, const char& strand,
                    const bool& AG_WILDCARD, const uint32_t& max_mismatches,
                    const uint32_t& b, TopCandidates& top_match,
                    StatSingleReads& stat_single_reads);

namespace {
struct RefIndex {
  Genome genome;
  HashTable hash_table;
};
}  // namespace

extern "C" {

struct waltref_best {   /* == BestMatch, fixed layout for ctypes */
  uint32_t genome_pos;
  uint32_t times;
  uint32_t mismatch;
  char strand;
  char pad[3];
};

struct waltref_cand {   /* == CandidatePosition */
  uint32_t genome_pos;
  uint32_t mismatch;
  char strand;
  char pad[3];
};

/* Allocate a sub-index with the given chromosome table; sequence/counter/index are sized
 * like ProcessSingledEndReads does (mapping.cpp:437-440). */
void* waltref_index_alloc(uint32_t n_chr, const uint32_t* lengths, uint32_t index_size) {
  RefIndex* r = new RefIndex;
  Genome& g = r->genome;
  g.num_of_chroms = n_chr;
  g.name.resize(n_chr);
  g.length.assign(lengths, lengths + n_chr);
  g.start_index.resize(n_chr + 1);
  g.start_index[0] = 0;
  for (uint32_t i = 1; i <= n_chr; ++i)
    g.start_index[i] = g.start_index[i - 1] + g.length[i - 1];
  g.length_of_genome = g.start_index[n_chr];
  g.sequence.resize(g.length_of_genome);
  r->hash_table.counter_size = power(4, F2SEEDKEYWEIGHT);
  r->hash_table.counter.resize(r->hash_table.counter_size + 1);
  r->hash_table.index_size = index_size;
  r->hash_table.index.resize(index_size);
  return r;
}

/* Load header + one sub-index with the reference's own readers (reference.cpp:324-417). */
void* waltref_index_load(const char* header_path, const char* subindex_path) {
  RefIndex* r = new RefIndex;
  uint32_t size_of_index = 0;
  ReadIndexHeadInfo(header_path, r->genome, size_of_index);
  r->genome.sequence.resize(r->genome.length_of_genome);
  r->hash_table.counter.resize(power(4, F2SEEDKEYWEIGHT) + 1);
  r->hash_table.index.resize(size_of_index);
  ReadIndex(subindex_path, r->genome, r->hash_table);
  return r;
}

void waltref_index_free(void* h) { delete static_cast<RefIndex*>(h); }

/* Raw buffers so a caller can fill them in place (e.g. a device->host copy). */
char* waltref_index_sequence(void* h) { return &static_cast<RefIndex*>(h)->genome.sequence[0]; }
uint32_t* waltref_index_counter(void* h) { return &static_cast<RefIndex*>(h)->hash_table.counter[0]; }
uint32_t* waltref_index_index(void* h) { return &static_cast<RefIndex*>(h)->hash_table.index[0]; }
uint64_t waltref_index_genome_len(void* h) { return static_cast<RefIndex*>(h)->genome.length_of_genome; }
uint32_t waltref_index_index_size(void* h) { return static_cast<RefIndex*>(h)->hash_table.index_size; }
void waltref_index_set_index_size(void* h, uint32_t n) { static_cast<RefIndex*>(h)->hash_table.index_size = n; }
void waltref_index_set_strand(void* h, char s) { static_cast<RefIndex*>(h)->genome.strand = s; }

/* One strand pass of the SE hot loop (mapping.cpp:494-499) over reads given as one
 * concatenated ASCII buffer.  `io` must be pre-initialised by the caller exactly like
 * mapping.cpp:486-489 before the '+' pass and is carried into the '-' pass.
 * Returns stat.num_of_short_reads accumulated by this pass. */
uint32_t waltref_map_se(void* h, const char* seqs, const uint64_t* offs, uint32_t n,
                        char strand, int ag_wildcard, uint32_t b, waltref_best* io,
                        int num_threads) {
  RefIndex* r = static_cast<RefIndex*>(h);
  std::vector<std::string> read_seqs(n);
  for (uint32_t j = 0; j < n; ++j)
    read_seqs[j].assign(seqs + offs[j], seqs + offs[j + 1]);
  std::vector<BestMatch> res(n);
  for (uint32_t j = 0; j < n; ++j)
    res[j] = BestMatch(io[j].genome_pos, io[j].times, io[j].strand, io[j].mismatch);
  StatSingleReads stat(false, false, "", true);
  const bool AG = ag_wildcard != 0;
  omp_set_dynamic(0);
  omp_set_num_threads(num_threads);
#pragma omp parallel for
  for (uint32_t j = 0; j < n; ++j)
    SingleEndMapping(read_seqs[j], r->genome, r->hash_table, strand, AG, b, res[j], stat);
  for (uint32_t j = 0; j < n; ++j) {
    io[j].genome_pos = res[j].genome_pos;
    io[j].times = res[j].times;
    io[j].mismatch = res[j].mismatch;
    io[j].strand = res[j].strand;
  }
  return stat.num_of_short_reads;
}

/* Same, but only the timed OpenMP loop is inside [t0,t1): returns seconds of wall time of
 * the loop itself (string construction excluded, as in the reference where the loader has
 * already produced std::string reads). */
double waltref_time_se(void* h, const char* seqs, const uint64_t* offs, uint32_t n,
                       char strand, int ag_wildcard, uint32_t b, waltref_best* io,
                       int num_threads) {
  RefIndex* r = static_cast<RefIndex*>(h);
  std::vector<std::string> read_seqs(n);
  for (uint32_t j = 0; j < n; ++j)
    read_seqs[j].assign(seqs + offs[j], seqs + offs[j + 1]);
  std::vector<BestMatch> res(n);
  for (uint32_t j = 0; j < n; ++j)
    res[j] = BestMatch(io[j].genome_pos, io[j].times, io[j].strand, io[j].mismatch);
  StatSingleReads stat(false, false, "", true);
  const bool AG = ag_wildcard != 0;
  omp_set_dynamic(0);
  omp_set_num_threads(num_threads);
  const double t0 = omp_get_wtime();
#pragma omp parallel for
  for (uint32_t j = 0; j < n; ++j)
    SingleEndMapping(read_seqs[j], r->genome, r->hash_table, strand, AG, b, res[j], stat);
  const double t1 = omp_get_wtime();
  for (uint32_t j = 0; j < n; ++j) {
    io[j].genome_pos = res[j].genome_pos;
    io[j].times = res[j].times;
    io[j].mismatch = res[j].mismatch;
    io[j].strand = res[j].strand;
  }
  return t1 - t0;
}

/* PE: one strand pass of the per-mate hot loop (paired.cpp:664-669).  The heaps live in an
 * opaque array that persists across the two strand passes of a mate. */
void* waltref_heaps_alloc(uint32_t n, uint32_t top_k) {
  std::vector<TopCandidates>* v = new std::vector<TopCandidates>(n);
  for (uint32_t j = 0; j < n; ++j) {
    (*v)[j].Clear();
    (*v)[j].SetSize(top_k);
  }
  return v;
}
void waltref_heaps_free(void* hp) { delete static_cast<std::vector<TopCandidates>*>(hp); }

uint32_t waltref_map_pe(void* h, void* hp, const char* seqs, const uint64_t* offs, uint32_t n,
                        char strand, int ag_wildcard, uint32_t max_mismatches, uint32_t b,
                        int num_threads) {
  RefIndex* r = static_cast<RefIndex*>(h);
  std::vector<TopCandidates>& heaps = *static_cast<std::vector<TopCandidates>*>(hp);
  std::vector<std::string> read_seqs(n);
  for (uint32_t j = 0; j < n; ++j)
    read_seqs[j].assign(seqs + offs[j], seqs + offs[j + 1]);
  StatSingleReads stat(false, false, "", true);
  const bool AG = ag_wildcard != 0;
  omp_set_dynamic(0);
  omp_set_num_threads(num_threads);
#pragma omp parallel for
  for (uint32_t j = 0; j < n; ++j)
    PairEndMapping(read_seqs[j], r->genome, r->hash_table, strand, AG, max_mismatches, b,
                   heaps[j], stat);
  return stat.num_of_short_reads;
}

/* Same pass, returning the wall time of the OpenMP loop alone (see waltref_time_se). */
double waltref_time_pe(void* h, void* hp, const char* seqs, const uint64_t* offs, uint32_t n,
                       char strand, int ag_wildcard, uint32_t max_mismatches, uint32_t b,
                       int num_threads) {
  RefIndex* r = static_cast<RefIndex*>(h);
  std::vector<TopCandidates>& heaps = *static_cast<std::vector<TopCandidates>*>(hp);
  std::vector<std::string> read_seqs(n);
  for (uint32_t j = 0; j < n; ++j)
    read_seqs[j].assign(seqs + offs[j], seqs + offs[j + 1]);
  StatSingleReads stat(false, false, "", true);
  const bool AG = ag_wildcard != 0;
  omp_set_dynamic(0);
  omp_set_num_threads(num_threads);
  const double t0 = omp_get_wtime();
#pragma omp parallel for
  for (uint32_t j = 0; j < n; ++j)
    PairEndMapping(read_seqs[j], r->genome, r->hash_table, strand, AG, max_mismatches, b,
                   heaps[j], stat);
  return omp_get_wtime() - t0;
}

/* Drain every heap like paired.cpp:684-692: out[j*top_k + 0] is the worst, the last one the
 * best; n_out[j] = number of entries. */
void waltref_heaps_drain(void* hp, uint32_t n, uint32_t top_k, waltref_cand* out,
                         uint32_t* n_out) {
  std::vector<TopCandidates>& heaps = *static_cast<std::vector<TopCandidates>*>(hp);
  for (uint32_t j = 0; j < n; ++j) {
    uint32_t c = 0;
    while (!heaps[j].Empty()) {
      CandidatePosition p = heaps[j].Top();
      heaps[j].Pop();
      out[(size_t)j * top_k + c].genome_pos = p.genome_pos;
      out[(size_t)j * top_k + c].mismatch = p.mismatch;
      out[(size_t)j * top_k + c].strand = p.strand;
      ++c;
    }
    n_out[j] = c;
  }
}

}  /* extern "C" */
