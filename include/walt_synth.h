/* walt_synth.h -- BENCH / TEST ONLY (libwaltsynth.so): seeded synthetic genomes and reads generated on
 * the device for bench.py.  Not part of the product ABI (include/walt_b200.h). */
#ifndef WALT_SYNTH_H_
#define WALT_SYNTH_H_

#include "walt_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* SURVEY.md 8(d) shapes */
int walt_synth_genome_device(int device, uint64_t n_bases, uint64_t seed, void* d_packed_out);
int walt_synth_reads_device(walt_engine* e, const void* d_packed_genome, uint32_t n_reads, uint32_t read_len,
                            uint64_t seed, int a_rich, void* d_seqs_out /* n_reads*read_len ASCII */);
/* Directional paired-end library: fragments N(300,50) clipped to [read_len, 1000]; mate 1 is the
 * T-rich 5' end, mate 2 the reverse complement of the 3' end (A-rich).  `readthrough_pct` percent
 * of the pairs have an insert shorter than the read, followed by random bases (what the loader's
 * adaptor clipping + N replacement leaves, mapping.cpp:92-103). */
int walt_synth_pairs_device(walt_engine* e, const void* d_packed_genome, uint32_t n_pairs, uint32_t read_len,
                            uint64_t seed, uint32_t readthrough_pct, void* d_seqs1_out, void* d_seqs2_out);
/* Repeat-heavy genome (configs[4]): ~40 % of the bases are copies of 2000 repeat families. */
int walt_synth_repeat_genome_device(int device, uint64_t n_bases, uint64_t seed, void* d_packed_out);

/* Verification-bound workload: rep_pct % of the genome's 512-base tiles start with a copy of one of
 * n_families 400-base families (div_per_mille substitutions per 1000 bases); reads come from inside copies. */
int walt_synth_verify_genome_device(int device, uint64_t n_bases, uint64_t seed, uint32_t n_families, uint32_t rep_pct,
                                    uint32_t div_per_mille, void* d_packed_out);
int walt_synth_verify_reads_device(walt_engine* e, const void* d_packed_genome, uint32_t n_reads, uint32_t read_len,
                                   uint64_t seed, uint64_t genome_seed, uint32_t n_families, uint32_t rep_pct, void* d_seqs_out);

#ifdef __cplusplus
}
#endif
#endif /* WALT_SYNTH_H_ */
