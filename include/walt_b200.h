/* walt_b200.h -- C ABI of the B200-native WALT mapping engine (libwaltb200.so).
 *
 * This is the drop-in boundary for the reference's mapping hot path.  The reference has no
 * FFI of its own; the seam is the strand-loop body of its two batch drivers
 *     ReadIndex(...) + #pragma omp parallel for { SingleEndMapping | PairEndMapping }
 *     (src/walt/mapping.cpp:491-500, src/walt/paired.cpp:660-671)
 * plus the index reader that feeds it (src/walt/reference.cpp:324-417).  Each entry point
 * below names the reference interface it replaces.  Plain pointers and sizes only; no C++
 * or torch types cross this boundary.  All functions return 0 on success and a non-zero
 * WALT_E* code on failure; walt_last_error() returns a message for the calling thread.
 * There is no CPU fallback: every mapping call runs on the CUDA device or fails.
 *
 * Conventions: caller owns all host buffers; calls are synchronous (internally the batch is
 * pipelined over CUDA streams, three chunks in flight); one engine per GPU per process; an engine is
 * thread-compatible, not thread-safe.  Results are indexed by input read order.
 */
#ifndef WALT_B200_H_
#define WALT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WALT_OK          0
#define WALT_EINVAL      1   /* bad argument                                              */
#define WALT_EIO         2   /* file missing / short read (reference: FREAD_CHECK, exit)  */
#define WALT_ECUDA       3   /* CUDA runtime error, including "no device"                 */
#define WALT_EFORMAT     4   /* index content violates the makedb invariants              */
#define WALT_ENONACGT    5   /* read contains a non-ACGT byte (reference: util.hpp:117-120
                                prints "[ERROR: NON-ACGT NUCLEOTIDE]" and exits)          */
#define WALT_ENOTLOADED  6   /* needed sub-index not resident                             */

/* Sub-index ids, in makedb order (src/walt/makedb.cpp:144-155). */
#define WALT_CT00 0   /* C->T, forward genome            */
#define WALT_CT01 1   /* C->T, reverse-complement genome */
#define WALT_GA10 2   /* G->A, forward genome            */
#define WALT_GA11 3   /* G->A, reverse-complement genome */

typedef struct walt_engine walt_engine;

/* == BestMatch (src/walt/mapping.hpp:39-52) with a fixed 16-byte layout. */
typedef struct {
  uint32_t genome_pos;
  uint32_t times;
  uint32_t mismatch;
  char strand;
  char pad[3];
} walt_best;

/* == CandidatePosition (src/walt/paired.hpp:35-46) with a fixed 12-byte layout. */
typedef struct {
  uint32_t genome_pos;
  uint32_t mismatch;
  char strand;
  char pad[3];
} walt_cand;

/* Per-pair result of the pairing step (MergePairedEndResults, paired.cpp:472-513). */
typedef struct {
  uint32_t best_times;     /* 1 = unique pair, >=2 ambiguous, 0 none                     */
  int32_t best_i, best_j;  /* indices into the ranked lists of mate 1 / mate 2, or -1    */
  int32_t frag_len;        /* GetFragmentLength of the winning pair (valid if times>=1)  */
} walt_pair;

/* Everything the output stage needs for one pair, without the ranked lists: the pairing result,
 * the two candidates of the winning pair (zero unless best_times >= 1) and each mate's own
 * single-end resolution (GetBestMatch4Single, paired.cpp:296-318; what MergePairedEndResults
 * reports for pairs that are not uniquely paired). */
typedef struct {
  walt_pair pair;
  walt_cand c1, c2;
  walt_best single1, single2;
} walt_pe_result;

/* Work counters of the last mapping call (device-side, for the roofline model). */
typedef struct {
  uint64_t n_lookups;      /* table lookups with a non-empty 12-mer bucket               */
  uint64_t n_candidates;   /* candidate windows compared against the genome              */
  uint64_t n_literal;      /* lookups that took the literal IndexRegion emulation        */
  uint64_t n_kernel_launches;
  uint64_t n_parked;       /* reads finished by the kernels behind the first (walt_engine_set_defer) */
  uint64_t n_verify_slots; /* index slots verified by verify_kernel (runs of parked reads)        */
  uint64_t verify_ns;      /* device time of its launches (walt_engine_set_kernel_timing)         */
} walt_stats;

const char* walt_last_error(void);

/* ---- engine lifetime ----------------------------------------------------------------- */
/* CUDA devices visible to the process (0 if there is none or the driver cannot be reached). */
int walt_device_count(void);
int walt_engine_create(walt_engine** out, int device);
void walt_engine_destroy(walt_engine* e);

/* ---- index residency ------------------------------------------------------------------
 * Replaces ReadIndexHeadInfo + the per-batch, per-strand ReadIndex calls
 * (reference.cpp:324-417; call sites mapping.cpp:437,492 and paired.cpp:578,661): the header
 * and the requested sub-indexes are read ONCE, packed to 2 bits/base and kept in HBM
 * together with the derived lookup table. `which_mask` has bit WALT_CT00.. set for every
 * sub-index to load (SE without -A: CT00|CT01; SE -A: GA10|GA11; PE: all four). */
int walt_engine_load_dbindex(walt_engine* e, const char* dbindex_path, uint32_t which_mask);

/* Same residency from caller memory (one sub-index per call): what ReadIndex would have
 * put into Genome/HashTable.  `sequence` is the converted ASCII genome of that sub-index. */
int walt_engine_set_chromosomes(walt_engine* e, uint32_t n_chr, const uint32_t* lengths,
                                const char* const* names /* may be NULL */);
int walt_engine_load_subindex(walt_engine* e, int which, const char* sequence,
                              const uint32_t* counter /* 4^12+1 */, const uint32_t* index,
                              uint32_t index_size);

/* Index replication between GPUs: everything `src` holds resident (packed genomes, entry arrays, prefix
 * tables, taint lists, chromosome table) is copied into `dst` device to device (cudaMemcpyPeer: NVLink /
 * NVSwitch when the devices are peers) -- no file is read again and nothing is rebuilt.  Replaces the
 * per-GPU repetition of ReadIndex (reference.cpp:324-351, call site mapping.cpp:492). */
int walt_engine_clone_index(walt_engine* dst, const walt_engine* src);

/* Chromosome table as read from the header (names are NUL-terminated, owned by the engine). */
int walt_engine_chromosomes(const walt_engine* e, uint32_t* n_chr, const uint32_t** lengths,
                            const uint32_t** start_index, const char* const** names);
uint64_t walt_engine_hbm_bytes(const walt_engine* e);
/* Residency facts of one sub-index: entries of index[], depth of the base-3 prefix table that
 * supersedes counter[], number of positions within 148 bases of a chromosome end. */
int walt_engine_subindex_info(const walt_engine* e, int which, uint32_t* index_size, uint32_t* depth,
                              uint32_t* n_taint);

/* ---- single-end mapping ---------------------------------------------------------------
 * Replaces both strand passes of mapping.cpp:486-500 for one batch: out[j] is what
 * map_results[j] holds after the '-' pass.  `seqs` is the concatenation of the reads as
 * LoadReadsFromFastqFile left them (ACGT only), read j = seqs[offs[j] .. offs[j+1]).
 * *n_short = 2 * #reads shorter than 38 (StatSingleReads::num_of_short_reads is bumped once
 * per strand pass, mapping.cpp:230-232). */
int walt_engine_map_se(walt_engine* e, const char* seqs, const uint64_t* offs, uint32_t n,
                       int ag_wildcard, uint32_t max_mismatches, uint32_t b, walt_best* out,
                       uint32_t* n_short);

/* The same call on a 2-bit packed batch (layout: walt_pack_reads, include/walt_host.h; `offs`
 * are still base offsets).  A quarter of the read bytes cross PCIe; the C->T / G->A conversion
 * still happens on the device. */
int walt_engine_map_se_packed(walt_engine* e, const uint8_t* packed, const uint64_t* offs, uint32_t n,
                              int ag_wildcard, uint32_t max_mismatches, uint32_t b, walt_best* out,
                              uint32_t* n_short);

/* Same computation with every buffer already resident in device memory (kernel-only
 * timing; no host<->device copies). */
int walt_engine_map_se_device(walt_engine* e, const void* d_seqs, const void* d_offs, uint32_t n,
                              uint32_t max_read_len, int ag_wildcard, uint32_t max_mismatches,
                              uint32_t b, void* d_out, void* cuda_stream);

/* ---- paired-end mapping ---------------------------------------------------------------
 * Replaces, for one batch, the mate/strand loops of paired.cpp:642-672 (PairEndMapping into
 * TopCandidates), the heap drain of paired.cpp:684-692 and the pairing loop of
 * MergePairedEndResults (paired.cpp:472-513).  ranked1/ranked2 hold top_k slots per pair,
 * worst first (ranked[j*top_k + 0]) to best (ranked[j*top_k + n_ranked[j]-1]).
 * pbat != 0 swaps the bisulfite roles of the mates (mate 1 G->A, mate 2 C->T). */
int walt_engine_map_pe(walt_engine* e, const char* seqs1, const uint64_t* offs1, const char* seqs2,
                       const uint64_t* offs2, uint32_t n, uint32_t max_mismatches, uint32_t b,
                       uint32_t top_k, int frag_range, int pbat, walt_cand* ranked1,
                       uint32_t* n_ranked1, walt_cand* ranked2, uint32_t* n_ranked2,
                       walt_pair* pairs, uint32_t* n_short1, uint32_t* n_short2);

/* The same mapping with the per-pair summary only (72 bytes per pair cross PCIe instead of
 * 2 * top_k * 12): pairing loop and GetBestMatch4Single run on the device. */
int walt_engine_map_pe_compact(walt_engine* e, const char* seqs1, const uint64_t* offs1, const char* seqs2,
                               const uint64_t* offs2, uint32_t n, uint32_t max_mismatches, uint32_t b,
                               uint32_t top_k, int frag_range, int pbat, walt_pe_result* out,
                               uint32_t* n_short1, uint32_t* n_short2);
/* ... on 2-bit packed mates (walt_pack_reads layout) ... */
int walt_engine_map_pe_compact_packed(walt_engine* e, const uint8_t* packed1, const uint64_t* offs1,
                                      const uint8_t* packed2, const uint64_t* offs2, uint32_t n,
                                      uint32_t max_mismatches, uint32_t b, uint32_t top_k, int frag_range, int pbat,
                                      walt_pe_result* out, uint32_t* n_short1, uint32_t* n_short2);
/* ... and with every buffer already resident in device memory (kernel-only timing). */
int walt_engine_map_pe_device(walt_engine* e, const void* d_seqs1, const void* d_offs1, const void* d_seqs2,
                              const void* d_offs2, uint32_t n, uint32_t max_read_len, uint32_t max_mismatches,
                              uint32_t b, uint32_t top_k, int frag_range, int pbat, void* d_out /* walt_pe_result[n] */,
                              void* cuda_stream);

int walt_engine_last_stats(const walt_engine* e, walt_stats* out);
/* After walt_engine_map_*_device calls: waits for the device, then returns (and clears) the work
 * counters accumulated since the last call of this function. */
int walt_engine_device_stats(walt_engine* e, walt_stats* out);

/* Test hook: 0 = table-driven search (default), 1 = literal IndexRegion emulation for every
 * lookup (the in-repo device oracle; same results, slower). */
int walt_engine_set_search_mode(walt_engine* e, int mode);
/* Tuning/test hooks: prefix-table depth for sub-indexes loaded afterwards (0 = auto, else
 * 12..20) and the number of reads per in-flight host chunk (0 = automatic: 2^18 for ASCII batches,
 * 2^19 for packed ones; three chunks in flight). */
int walt_engine_set_table_depth(walt_engine* e, int depth);
int walt_engine_set_chunk_reads(walt_engine* e, uint32_t n);
/* Lanes of a warp that cooperate on one read: 8 (default; four reads per warp), 16 or 32. */
int walt_engine_set_group_width(walt_engine* e, uint32_t lanes);
/* 1 (default): a read whose ordered fold reaches a lookup that needs more than one lane (a long run of
 * equal seeds = repeats, or a bucket next to a chromosome end) is parked by the mapping kernel and
 * finished by a second kernel with a whole warp per read; 0: every read is finished by its group. */
int walt_engine_set_defer(walt_engine* e, int on);
/* 1: CUDA events around the verify_kernel launches of the device-resident calls; their sum comes back
 * in walt_stats.verify_ns from walt_engine_device_stats (how bench.py times verification alone). */
int walt_engine_set_kernel_timing(walt_engine* e, int on);

/* ---- several GPUs behind one handle --------------------------------------------------------
 * A group owns one engine per listed device (reads shard naturally, every GPU
 * holds a full index replica, there is no collective on the hot path).  walt_group_load_dbindex reads
 * the files ONCE (into the first device) and replicates with walt_engine_clone_index.  The map calls
 * cut the batch into n_devices contiguous ranges, one host thread drives each engine, and every range
 * lands in its slice of the caller's result array: results are in input order whatever the split.
 * Arguments as in the walt_engine_map_* calls of the same name. */
typedef struct walt_group walt_group;
int walt_group_create(walt_group** out, const int* device_ids, int n_devices);
void walt_group_destroy(walt_group* g);
int walt_group_size(const walt_group* g);
walt_engine* walt_group_engine(walt_group* g, int i);     /* for the per-engine setters and queries */
int walt_group_load_dbindex(walt_group* g, const char* dbindex_path, uint32_t which_mask);
int walt_group_map_se(walt_group* g, const char* seqs, const uint64_t* offs, uint32_t n, int ag_wildcard,
                      uint32_t max_mismatches, uint32_t b, walt_best* out, uint32_t* n_short);
int walt_group_map_se_packed(walt_group* g, const uint8_t* packed, const uint64_t* offs, uint32_t n, int ag_wildcard,
                             uint32_t max_mismatches, uint32_t b, walt_best* out, uint32_t* n_short);
int walt_group_map_pe_compact(walt_group* g, const char* seqs1, const uint64_t* offs1, const char* seqs2,
                              const uint64_t* offs2, uint32_t n, uint32_t max_mismatches, uint32_t b, uint32_t top_k,
                              int frag_range, int pbat, walt_pe_result* out, uint32_t* n_short1, uint32_t* n_short2);
int walt_group_map_pe_compact_packed(walt_group* g, const uint8_t* packed1, const uint64_t* offs1, const uint8_t* packed2,
                                     const uint64_t* offs2, uint32_t n, uint32_t max_mismatches, uint32_t b, uint32_t top_k,
                                     int frag_range, int pbat, walt_pe_result* out, uint32_t* n_short1, uint32_t* n_short2);

/* ---- pinned host memory for batch buffers --------------------------------------------- */
void* walt_host_alloc(size_t bytes);
void walt_host_free(void* p);

/* ---- index construction on the device (makedb) -------------------------------------------
 * Replaces BuildIndex (makedb.cpp:46-85; reference.cpp:131-300): reverse complement,
 * C->T / G->A conversion, bucket counting with the ">= 500000" erasure, bucketing and bucket
 * sorting all run on the GPU and leave the requested sub-indexes resident in the engine.
 * The forward genome is 2 bits/base (A0 C1 G2 T3), 32 bases per 64-bit word, first base in
 * the top bits, one zero word in front; a buffer holds walt_packed_genome_bytes(n) bytes. */
uint64_t walt_packed_genome_bytes(uint64_t n_bases);
int walt_pack_genome_device(int device, const char* sequence /* upper-case ACGT */, uint64_t n_bases,
                            void* d_packed_out);
int walt_engine_build_from_device_genome(walt_engine* e, const void* d_packed_genome, uint32_t which_mask);
int walt_engine_build_from_sequence(walt_engine* e, const char* sequence, uint32_t which_mask);

/* Order of suffixes that compare equal under SortHashTableBucketCMP (reference.cpp:258-288)
 * in indexes built afterwards: 0 (default) = the order the reference's std::sort leaves them in
 * (reference.cpp:296-298; libstdc++ introsort replayed per bucket, so the .dbindex files are
 * byte-identical to the reference makedb's), 1 = ascending position (skips the replay). */
int walt_engine_set_tie_order(walt_engine* e, int mode);
/* Of the last sub-index built: slots that tie with their predecessor, buckets replayed. */
int walt_engine_last_build_info(const walt_engine* e, uint64_t* n_tied_slots, uint64_t* n_buckets_replayed);

/* Export a resident sub-index to host arrays in the reference's in-memory form (what ReadIndex
 * fills, reference.cpp:324-351): converted ASCII genome, counter[4^12+1], index[index_size].
 * Any pointer may be NULL. */
int walt_engine_export_subindex(walt_engine* e, int which, char* sequence, uint32_t* counter,
                                uint32_t* index, uint32_t* index_size);

#ifdef __cplusplus
}
#endif
#endif /* WALT_B200_H_ */
