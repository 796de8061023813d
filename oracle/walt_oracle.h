/* oracle/walt_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference WALT mapping path.  It exists to CHECK the CUDA
 * engine (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg) and is never linked
 * into, imported by or called from the product (walt_b200/).  Parity of this restatement is
 * PINNED against the compiled, unmodified reference (oracle/_ref/libwaltref.so, walt, makedb)
 * by tests/test_oracle_vs_reference.py and against the committed golden vectors in
 * tests/golden/ (generated from the reference binaries by tests/golden/make_golden.py).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/).
 */
#ifndef WALT_ORACLE_H_
#define WALT_ORACLE_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* == BestMatch, src/walt/mapping.hpp:39-52 (fixed layout for ctypes) */
typedef struct {
  uint32_t genome_pos;
  uint32_t times;
  uint32_t mismatch;
  char strand;
  char pad[3];
} wo_best;

/* == CandidatePosition, src/walt/paired.hpp:35-46 */
typedef struct {
  uint32_t genome_pos;
  uint32_t mismatch;
  char strand;
  char pad[3];
} wo_cand;

/* One strand sub-index as it sits in a .dbindex_XXnn file (src/walt/reference.cpp:302-351):
 * converted ASCII genome, chromosome starts (prefix sums of the header lengths,
 * reference.cpp:407-410), counter[4^12+1], index[index_size]. */
typedef struct {
  const char* seq;
  uint64_t genome_len;
  uint32_t n_chr;
  const uint32_t* start_index; /* n_chr + 1 */
  const uint32_t* counter;     /* 16777217 */
  const uint32_t* index;
  uint32_t index_size;
} wo_index;

/* Work counters that feed the algorithmic-bytes model of SURVEY.md section 8(d). */
typedef struct {
  uint64_t n_reads;        /* reads with rl >= 38 that entered the seed loop (per strand pass) */
  uint64_t n_lookups;      /* non-empty bucket lookups performed after the early exits        */
  uint64_t sum_log2_bucket;/* sum over lookups of ceil(log2(B_i + 1))                          */
  uint64_t n_probes;       /* binary-search probes actually executed (index+genome loads)      */
  uint64_t n_cand;         /* candidates that passed the bounds checks and were verified       */
  uint64_t n_region_over_b;/* lookups dropped by the -b filter                                 */
  uint64_t n_short;        /* reads shorter than 38 (counted once per strand pass)             */
} wo_counters;

/* seed geometry, src/walt/mapping.cpp:236-239 + src/walt/seedpattern.hpp:355-456 */
uint32_t wo_seed_repeats(uint32_t read_len);
/* fills mask[0..read_len) with 1 where SingleEndMapping compares (mapping.cpp:288-304) */
void wo_verify_positions(uint32_t seed_i, uint32_t read_len, uint8_t* mask);
uint32_t wo_nocared_position(uint32_t seed_i, uint32_t q);

uint32_t wo_hash_value(const char* seed);                           /* util.hpp:175-182 */
uint32_t wo_chrom_id(const uint32_t* starts, uint32_t n_starts, uint32_t pos); /* reference.cpp:43-60 */

/* IndexRegion (mapping.cpp:198-222); region in/out as (first, second) */
void wo_index_region(const wo_index* ix, const char* read_seed, uint32_t seed_len,
                     uint32_t* first, uint32_t* second, wo_counters* ctr);

/* SingleEndMapping (mapping.cpp:224-316) for one read and one strand pass */
void wo_se_map_read(const wo_index* ix, const char* read, uint32_t read_len, char strand,
                    int ag_wildcard, uint32_t b, wo_best* io, wo_counters* ctr);

/* the strand loop body of ProcessSingledEndReads (mapping.cpp:494-499) for a batch */
void wo_se_map_batch(const wo_index* ix, const char* seqs, const uint64_t* offs, uint32_t n,
                     char strand, int ag_wildcard, uint32_t b, wo_best* io, wo_counters* ctr);

/* ---- paired-end ---- */
/* TopCandidates (paired.hpp:51-74) with libstdc++ heap mechanics
 * (/usr/include/c++/13/bits/stl_heap.h:135-147,224-267). */
typedef struct {
  wo_cand* a;      /* capacity max_size */
  uint32_t size;
  uint32_t max_size;
} wo_heap;

void wo_heap_push(wo_heap* h, wo_cand c);   /* TopCandidates::Push */
wo_cand wo_heap_pop(wo_heap* h);            /* Top() + Pop()        */

/* PairEndMapping (paired.cpp:106-201) for one read and one strand pass */
void wo_pe_map_read(const wo_index* ix, const char* read, uint32_t read_len, char strand,
                    int ag_wildcard, uint32_t max_mismatches, uint32_t b, wo_heap* heap,
                    wo_counters* ctr);

/* batch variant: heaps laid out as cands[j*top_k ..], sizes[j] (persist across strands) */
void wo_pe_map_batch(const wo_index* ix, const char* seqs, const uint64_t* offs, uint32_t n,
                     char strand, int ag_wildcard, uint32_t max_mismatches, uint32_t b,
                     uint32_t top_k, wo_cand* cands, uint32_t* sizes, wo_counters* ctr);

/* drain (paired.cpp:684-692): ranked[0] worst ... ranked[n-1] best; returns n */
uint32_t wo_heap_drain(wo_heap* h, wo_cand* ranked);

/* pairing loop of MergePairedEndResults (paired.cpp:472-513).  Returns best_times; the
 * winning pair indices go to best_i/best_j (-1 if none) and its fragment length to frag. */
typedef struct {
  uint32_t n_chr;
  const uint32_t* start_index; /* n_chr + 1 */
  const uint32_t* length;      /* n_chr */
} wo_chroms;

uint32_t wo_pe_pair(const wo_chroms* g, const wo_cand* r1, uint32_t n1, uint32_t len1,
                    const wo_cand* r2, uint32_t n2, uint32_t len2, uint32_t max_mismatches,
                    int frag_range, int32_t* best_i, int32_t* best_j);

/* GetBestMatch4Single (paired.cpp:296-318) */
void wo_best_for_single(const wo_cand* ranked, uint32_t n, uint32_t max_mismatches, wo_best* out);

/* Everything MergePairedEndResults derives per pair from the two ranked lists (pairing loop,
 * the winning candidates, each mate's GetBestMatch4Single), in the fixed 72-byte layout of
 * walt_pe_result (include/walt_b200.h), for n pairs whose lists are top_k apart.  Threads: OpenMP. */
typedef struct {
  uint32_t best_times; int32_t best_i, best_j, frag_len;
  wo_cand c1, c2;
  wo_best single1, single2;
} wo_pe_result;
void wo_pe_result_batch(const wo_chroms* g, const wo_cand* ranked1, const uint32_t* n1, const uint64_t* offs1,
                        const wo_cand* ranked2, const uint32_t* n2, const uint64_t* offs2, uint32_t n, uint32_t top_k,
                        uint32_t max_mismatches, int frag_range, wo_pe_result* out);

/* GetFragmentLength (paired.cpp:320-331) */
int wo_fragment_length(const wo_chroms* g, const wo_cand* r1, uint32_t len1, const wo_cand* r2,
                       uint32_t len2, uint32_t chr1, uint32_t chr2);

/* ---- index builder (makedb), reference.cpp:192-300 ---- */
/* Build counter/index for an already converted ASCII genome. counter must hold 4^12+1
 * entries; index must hold genome_len entries; returns index_size.  Every bucket is sorted by
 * a restatement of libstdc++'s std::sort (wo_std_sort_bucket), so suffixes that compare equal
 * come out in the order the reference's makedb writes them. */
uint32_t wo_build_index(const char* seq, uint64_t genome_len, uint32_t n_chr,
                        const uint32_t* start_index, uint32_t* counter, uint32_t* index);
/* std::sort(v, v + n, SortHashTableBucketCMP(genome)) as GCC 13 libstdc++ executes it */
void wo_std_sort_bucket(const char* seq, uint32_t n_chr, const uint32_t* start_index, uint32_t* v, uint32_t n);
/* comparator of SortHashTableBucketCMP (reference.cpp:258-288): <0, 0, >0 */
int wo_bucket_cmp(const char* seq, uint32_t n_chr, const uint32_t* start_index, uint32_t p1,
                  uint32_t p2);

/* adaptor clipping, util.hpp:189-217; returns number of clipped characters */
size_t wo_clip_adaptor(const char* adaptor, size_t adaptor_len, char* s, size_t len);

#ifdef __cplusplus
}
#endif
#endif /* WALT_ORACLE_H_ */
