/* walt_host.h -- C ABI of the host side of the B200 WALT engine (libwalthost.so).
 *
 * Pure host code (no CUDA): the byte-compatible FASTQ batch loader, adaptor clipping, the
 * .dbindex header reader and the SAM / MR / .mapstats writers that sit on either side of the
 * mapping call.  `walt` (walt_b200/host/walt_main.cpp) is loader -> walt_engine_map_* ->
 * writer; the library form exists so the same code can be driven from tests without a GPU.
 * Each entry point cites the reference function it replaces (paths under src/walt/).
 */
#ifndef WALT_HOST_H_
#define WALT_HOST_H_

#include <stddef.h>
#include <stdint.h>

#include "walt_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

const char* walt_host_last_error(void);

/* Host threads used by the loader and the writers (the reference's -t sets its OpenMP mapping
 * threads, walt.cpp:150; here the mapping is on the GPU and -t sizes the text path instead).
 * 0 = one per hardware thread (default).  Output bytes do not depend on it. */
void walt_host_set_threads(unsigned n);
unsigned walt_host_threads(void);
/* Tuning/test hook: the smallest piece of a file one loader task scans (default 64 KiB) and the
 * number of reads one writer task formats (default 8192); 0 restores a default. */
void walt_host_set_grain(uint32_t chunk_bytes, uint32_t block_reads);

/* ---- chromosome table: ReadIndexHeadInfo (reference.cpp:381-417) -------------------------- */
typedef struct walt_chroms walt_chroms;
walt_chroms* walt_chroms_read(const char* dbindex_path);
walt_chroms* walt_chroms_create(uint32_t n_chr, const char* const* names, const uint32_t* lengths);
void walt_chroms_free(walt_chroms* c);
uint32_t walt_chroms_count(const walt_chroms* c);
const uint32_t* walt_chroms_lengths(const walt_chroms* c);
const char* walt_chroms_name(const walt_chroms* c, uint32_t i);

/* ---- FASTQ batches: LoadReadsFromFastqFile (mapping.cpp:65-121) --------------------------- */
typedef struct walt_fastq walt_fastq;     /* an open reads file                               */
typedef struct walt_batch walt_batch;     /* names / sequences / qualities of one batch       */
walt_fastq* walt_fastq_open(const char* path);
void walt_fastq_close(walt_fastq* f);
walt_batch* walt_batch_create(void);
void walt_batch_free(walt_batch* b);
/* Loads up to max_reads records (srand(0) first, adaptor clipped if non-empty, every
 * non-ACGT character replaced by "ACGT"[rand() % 4] in file order).  Returns the number of
 * reads loaded, or -1 on error. */
int64_t walt_fastq_next_batch(walt_fastq* f, walt_batch* b, uint32_t max_reads, const char* adaptor);
/* The same in pieces: a batch of N reads loaded as consecutive parts is the same batch if the first part
 * restarts the rand() stream (restart_rand != 0) and the others carry it on -- the parts can then be
 * mapped and written while the next ones are loaded (walt_main.cpp). */
int64_t walt_fastq_next_part(walt_fastq* f, walt_batch* b, uint32_t max_reads, const char* adaptor, int restart_rand);
uint32_t walt_batch_size(const walt_batch* b);
/* concatenated sequences + offsets[n+1]: the layout walt_engine_map_se/pe take */
const char* walt_batch_seqs(const walt_batch* b);
const uint64_t* walt_batch_offsets(const walt_batch* b);
const char* walt_batch_name(const walt_batch* b, uint32_t i);
const char* walt_batch_qual(const walt_batch* b, uint32_t i);

/* 2-bit form of a batch for walt_engine_map_se_packed / walt_engine_map_pe_compact_packed (a
 * quarter of the bytes cross PCIe).  Codes A0 C1 G2 T3 (util.hpp:107-121), unconverted, four
 * bases per byte with the first base in the top bits; read j occupies ceil(len_j / 4) bytes
 * starting at byte (offs[j] >> 2) + j, so the same offsets array addresses both forms.  `out`
 * holds walt_packed_reads_bytes(offs, n) bytes.  Returns -1 if a byte is not A/C/G/T (the
 * loader never leaves one behind, mapping.cpp:101-104). */
uint64_t walt_packed_reads_bytes(const uint64_t* offs, uint32_t n);
int walt_pack_reads(const char* seqs, const uint64_t* offs, uint32_t n, uint8_t* out);
/* The loader's own packed copy of the batch (built on first use, owned by the batch). */
const uint8_t* walt_batch_packed(walt_batch* b);

/* clip_adaptor_from_read (util.hpp:189-217); returns the number of clipped characters */
size_t walt_clip_adaptor(const char* adaptor, char* s, size_t len);

/* ---- single-end output: mapping.cpp:318-419, 503-522 ------------------------------------- */
typedef struct walt_se_writer walt_se_writer;
walt_se_writer* walt_se_writer_open(const char* output_path, const walt_chroms* chroms, int ag_wildcard,
                                    int ambiguous, int unmapped, int sam);
/* StatInfoUpdate + OutputSingleSAM | OutputSingleResults for reads [0, n) of the batch */
int walt_se_writer_write(walt_se_writer* w, const walt_batch* b, const walt_best* results, uint32_t n);
/* adds to StatSingleReads::num_of_short_reads */
void walt_se_writer_add_short(walt_se_writer* w, uint32_t n_short);
/* closes the outputs and appends StatSingleReads::tostring to <output>.mapstats */
int walt_se_writer_close(walt_se_writer* w);

/* ---- paired-end output: MergePairedEndResults after the pairing loop (paired.cpp:515-569),
 *      OutputBestPairedResults, GetBestMatch4Single, OutputPairedSAM, mapstats ------------- */
typedef struct walt_pe_writer walt_pe_writer;
walt_pe_writer* walt_pe_writer_open(const char* output_path, const walt_chroms* chroms, uint32_t max_mismatches,
                                    uint32_t top_k, int frag_range, int ambiguous, int unmapped, int sam,
                                    int pbat);
int walt_pe_writer_write(walt_pe_writer* w, const walt_batch* b1, const walt_batch* b2, const walt_cand* ranked1,
                         const uint32_t* n_ranked1, const walt_cand* ranked2, const uint32_t* n_ranked2,
                         const walt_pair* pairs, uint32_t n);
/* Same output from the engine's compact per-pair summary (walt_engine_map_pe_compact). */
int walt_pe_writer_write_compact(walt_pe_writer* w, const walt_batch* b1, const walt_batch* b2,
                                 const walt_pe_result* results, uint32_t n);
void walt_pe_writer_add_short(walt_pe_writer* w, uint32_t n_short1, uint32_t n_short2);
int walt_pe_writer_close(walt_pe_writer* w);

/* SAMHead (reference.cpp:430-440) appended to `output_path` */
int walt_write_sam_header(const char* output_path, const walt_chroms* chroms);

/* ---- makedb output: WriteIndex / WriteIndexHeadInfo (reference.cpp:302-322,353-379) ------- */
int walt_write_dbindex_header(const char* path, const walt_chroms* chroms, uint32_t size_of_index);
int walt_write_subindex(const char* path, char strand, const char* sequence, uint64_t genome_len,
                        const uint32_t* counter, const uint32_t* index, uint32_t index_size);

#ifdef __cplusplus
}
#endif
#endif /* WALT_HOST_H_ */
