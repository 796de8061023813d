/* oracle/walt_oracle.c -- TEST INFRASTRUCTURE ONLY (see walt_oracle.h).
 *
 * CPU restatement of the reference WALT mapping path in plain C.  Written from the
 * fold-form description of the algorithm; each function cites the reference lines it
 * follows (paths relative to /root/reference/).  Parity of this file against the compiled
 * reference is pinned by tests/test_oracle_vs_reference.py.
 */
#include "walt_oracle.h"

#include <stdlib.h>
#include <string.h>

enum { KEY_WEIGHT = 12, PATTERN_LEN = 3, MIN_READ_LEN = 38, MIN_SEED_LEN = 36, MAX_REPEATS = 50 };

/* ------------------------------------------------------------------------------------ */
/* seed geometry                                                                        */
/* ------------------------------------------------------------------------------------ */

/* cared position of seed character p inside a seed that starts at offset 0:
 * F2CAREDPOSITION[p] = 3p + 1 (src/walt/seedpattern.hpp:424-430). */
static inline uint32_t cared_position(uint32_t p) { return PATTERN_LEN * p + 1; }

/* src/walt/mapping.cpp:236-239 */
uint32_t wo_seed_repeats(uint32_t read_len) {
  uint32_t r = (read_len - PATTERN_LEN + 1) / PATTERN_LEN;
  return r < MAX_REPEATS ? r : MAX_REPEATS;
}

/* F2NOCAREDPOSITION[seed_i][q] (src/walt/seedpattern.hpp:431-455).  The table is the q-th
 * position (ascending) that is NOT of the form seed_i + 3p + 1, with two deviations in
 * row 2 that are reachable for reads >= 71 / >= 143 bases: entry 47 holds 60 (regular: 70)
 * and entry 95 holds 141 (regular: 142).  Entries q >= 102 are never read because
 * q < 2 * repeats + seed_i <= 102 (mapping.cpp:290-293). */
uint32_t wo_nocared_position(uint32_t seed_i, uint32_t q) {
  if (seed_i == 2) {
    if (q == 47) return 60;
    if (q == 95) return 141;
  }
  /* positions below seed_i are all "no-care"; afterwards two of every three */
  if (q < seed_i + 1) return q;
  uint32_t k = q - (seed_i + 1);           /* index among the positions after seed_i+1 ...  */
  /* after the first cared position (seed_i + 1) the no-care positions come in pairs:
   * seed_i+2, seed_i+3, seed_i+5, seed_i+6, ... */
  return seed_i + 2 + (k / 2) * 3 + (k % 2);
}

/* positions compared by the two verification loops, mapping.cpp:288-304 */
void wo_verify_positions(uint32_t seed_i, uint32_t read_len, uint8_t* mask) {
  memset(mask, 0, read_len);
  uint32_t rep = wo_seed_repeats(read_len);
  uint32_t n_nocared = rep * 2 + seed_i;
  for (uint32_t q = 0; q < n_nocared; ++q) mask[wo_nocared_position(seed_i, q)] += 1;
  for (uint32_t p = rep * PATTERN_LEN + seed_i; p < read_len; ++p) mask[p] += 1;
}

/* util.hpp:107-121 */
static inline uint32_t base_bits(char c) {
  switch (c) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    default:  return 3; /* 'T'; anything else cannot occur after toACGT */
  }
}

/* util.hpp:175-182 */
uint32_t wo_hash_value(const char* seed) {
  uint32_t h = 0;
  for (uint32_t i = 0; i < KEY_WEIGHT; ++i) h = (h << 2) + base_bits(seed[cared_position(i)]);
  return h;
}

/* reference.cpp:43-60: last m with starts[m] <= pos, searched over all n_starts entries */
uint32_t wo_chrom_id(const uint32_t* starts, uint32_t n_starts, uint32_t pos) {
  uint32_t l = 0, h = n_starts - 1;
  while (l < h) {
    uint32_t m = (l + h + 1) / 2;
    if (pos >= starts[m]) l = m; else h = m - 1;
  }
  return l;
}

/* genome byte as the reference's unchecked vector read sees it: past the end of the
 * allocation the de-facto value is 0. */
static inline char genome_at(const wo_index* ix, uint64_t pos) {
  return pos < ix->genome_len ? ix->seq[pos] : 0;
}

/* ------------------------------------------------------------------------------------ */
/* IndexRegion, mapping.cpp:166-222                                                     */
/* ------------------------------------------------------------------------------------ */
static uint32_t lower_bound_step(const wo_index* ix, uint32_t low, uint32_t high, char ch,
                                 uint32_t cmp_pos, wo_counters* ctr) {
  while (low < high) {
    uint32_t mid = low + (high - low) / 2;
    char c = genome_at(ix, (uint64_t)ix->index[mid] + cmp_pos);
    if (ctr) ctr->n_probes++;
    if (c >= ch) high = mid; else low = mid + 1;
  }
  return low;
}

static uint32_t upper_bound_step(const wo_index* ix, uint32_t low, uint32_t high, char ch,
                                 uint32_t cmp_pos, wo_counters* ctr) {
  while (low < high) {
    uint32_t mid = low + (high - low + 1) / 2;
    char c = genome_at(ix, (uint64_t)ix->index[mid] + cmp_pos);
    if (ctr) ctr->n_probes++;
    if (c <= ch) low = mid; else high = mid - 1;
  }
  return low;
}

void wo_index_region(const wo_index* ix, const char* read_seed, uint32_t seed_len,
                     uint32_t* first, uint32_t* second, wo_counters* ctr) {
  uint32_t l = *first, u = *second - 1;
  for (uint32_t p = KEY_WEIGHT; p < seed_len; ++p) {
    uint32_t cp = cared_position(p);
    l = lower_bound_step(ix, l, u, read_seed[cp], cp, ctr);
    u = upper_bound_step(ix, l, u, read_seed[cp], cp, ctr);
    if (l == u && read_seed[cp] != genome_at(ix, (uint64_t)ix->index[l] + cp)) {
      *first = 1; *second = 0;
      return;
    }
  }
  if (l > u) { *first = 1; *second = 0; return; }
  *first = l; *second = u;
}

/* ------------------------------------------------------------------------------------ */
/* shared front half of SingleEndMapping / PairEndMapping                               */
/* ------------------------------------------------------------------------------------ */
static void convert_read(const char* in, uint32_t n, int ag, char* out) {
  /* mapping.cpp:142-164 */
  for (uint32_t i = 0; i < n; ++i) {
    char c = in[i];
    if (ag) out[i] = (c == 'G') ? 'A' : c; else out[i] = (c == 'C') ? 'T' : c;
  }
}

static uint32_t ceil_log2_u64(uint64_t v) { /* ceil(log2(v)) for v >= 1 */
  uint32_t r = 0;
  while (((uint64_t)1 << r) < v) ++r;
  return r;
}

/* Masked mismatch count at one candidate, mapping.cpp:288-304 / paired.cpp:174-190.
 * `limit` is the running bound of the two loops (count stops once it exceeds limit). */
static uint32_t count_mismatches(const wo_index* ix, const char* read, uint32_t read_len,
                                 uint32_t genome_pos, uint32_t seed_i, uint32_t rep,
                                 uint32_t limit) {
  uint32_t mm = 0;
  uint32_t n_nocared = rep * 2 + seed_i;
  for (uint32_t q = 0; q < n_nocared && mm <= limit; ++q) {
    uint32_t pos = wo_nocared_position(seed_i, q);
    if (genome_at(ix, (uint64_t)genome_pos + pos) != read[pos]) ++mm;
  }
  for (uint32_t p = rep * PATTERN_LEN + seed_i; p < read_len && mm <= limit; ++p)
    if (genome_at(ix, (uint64_t)genome_pos + p) != read[p]) ++mm;
  return mm;
}

/* ------------------------------------------------------------------------------------ */
/* SingleEndMapping, mapping.cpp:224-316                                                */
/* ------------------------------------------------------------------------------------ */
void wo_se_map_read(const wo_index* ix, const char* org_read, uint32_t read_len, char strand,
                    int ag, uint32_t b, wo_best* best, wo_counters* ctr) {
  if (read_len < MIN_READ_LEN) { if (ctr) ctr->n_short++; return; }
  if (ctr) ctr->n_reads++;
  uint32_t rep = wo_seed_repeats(read_len);
  uint32_t seed_len = rep;
  char stackbuf[1024];
  char* read = read_len <= sizeof(stackbuf) ? stackbuf : (char*)malloc(read_len);
  convert_read(org_read, read_len, ag, read);

  for (uint32_t seed_i = 0; seed_i < PATTERN_LEN; ++seed_i) {
    if (best->mismatch == 0 && seed_i) break;       /* mapping.cpp:250-251 */
    if (best->mismatch == 1 && seed_i >= 2) break;  /* mapping.cpp:255-256 */
    const char* read_seed = read + seed_i;
    uint32_t h = wo_hash_value(read_seed);
    uint32_t first = ix->counter[h], second = ix->counter[h + 1];
    if (first == second) continue;
    if (ctr) {
      ctr->n_lookups++;
      ctr->sum_log2_bucket += ceil_log2_u64((uint64_t)(second - first) + 1);
    }
    wo_index_region(ix, read_seed, seed_len, &first, &second, ctr);
    if (second - first + 1 > b) { if (ctr) ctr->n_region_over_b++; continue; }

    for (uint32_t j = first; j <= second; ++j) {
      uint32_t gp = ix->index[j];
      uint32_t chr = wo_chrom_id(ix->start_index, ix->n_chr + 1, gp);
      if (gp - ix->start_index[chr] < seed_i) continue;
      gp -= seed_i;
      if (gp + read_len >= ix->start_index[chr + 1]) continue;
      if (ctr) ctr->n_cand++;
      uint32_t mm = count_mismatches(ix, read, read_len, gp, seed_i, rep, best->mismatch);
      if (mm < best->mismatch) {
        best->genome_pos = gp; best->times = 1; best->strand = strand; best->mismatch = mm;
      } else if (mm == best->mismatch && best->genome_pos != gp) {
        best->genome_pos = gp; best->strand = strand; best->times++;
      }
    }
  }
  if (read != stackbuf) free(read);
}

void wo_se_map_batch(const wo_index* ix, const char* seqs, const uint64_t* offs, uint32_t n,
                     char strand, int ag, uint32_t b, wo_best* io, wo_counters* ctr) {
  for (uint32_t j = 0; j < n; ++j)
    wo_se_map_read(ix, seqs + offs[j], (uint32_t)(offs[j + 1] - offs[j]), strand, ag, b, &io[j], ctr);
}

/* ------------------------------------------------------------------------------------ */
/* TopCandidates with libstdc++ heap mechanics                                          */
/* ------------------------------------------------------------------------------------ */
/* std::__push_heap (stl_heap.h:135-147) with comp = operator< on mismatch */
static void sift_up(wo_cand* a, uint32_t hole, uint32_t top, wo_cand v) {
  while (hole > top) {
    uint32_t parent = (hole - 1) / 2;
    if (!(a[parent].mismatch < v.mismatch)) break;
    a[hole] = a[parent];
    hole = parent;
  }
  a[hole] = v;
}

/* std::__adjust_heap (stl_heap.h:224-249) */
static void adjust_heap(wo_cand* a, uint32_t hole, uint32_t len, wo_cand v) {
  const uint32_t top = hole;
  uint32_t child = hole;
  while (len >= 2 && child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (a[child].mismatch < a[child - 1].mismatch) --child;
    a[hole] = a[child];
    hole = child;
  }
  if ((len & 1) == 0 && len >= 2 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    a[hole] = a[child - 1];
    hole = child - 1;
  }
  sift_up(a, hole, top, v);
}

static void heap_push_raw(wo_heap* h, wo_cand c) { /* priority_queue::push */
  h->a[h->size] = c;
  h->size++;
  sift_up(h->a, h->size - 1, 0, c);
}

wo_cand wo_heap_pop(wo_heap* h) { /* top() then priority_queue::pop -> std::__pop_heap */
  wo_cand top = h->a[0];
  uint32_t n = h->size;
  if (n > 1) {
    wo_cand v = h->a[n - 1];
    h->a[n - 1] = h->a[0];
    adjust_heap(h->a, 0, n - 1, v);
  }
  h->size = n - 1;
  return top;
}

void wo_heap_push(wo_heap* h, wo_cand c) { /* paired.hpp:63-70 */
  if (h->size < h->max_size) {
    heap_push_raw(h, c);
  } else if (c.mismatch < h->a[0].mismatch) {
    (void)wo_heap_pop(h);
    heap_push_raw(h, c);
  }
}

uint32_t wo_heap_drain(wo_heap* h, wo_cand* ranked) {
  uint32_t n = 0;
  while (h->size) ranked[n++] = wo_heap_pop(h);
  return n;
}

/* ------------------------------------------------------------------------------------ */
/* PairEndMapping, paired.cpp:106-201                                                   */
/* ------------------------------------------------------------------------------------ */
void wo_pe_map_read(const wo_index* ix, const char* org_read, uint32_t read_len, char strand,
                    int ag, uint32_t max_mismatches, uint32_t b, wo_heap* heap, wo_counters* ctr) {
  if (read_len < MIN_READ_LEN) { if (ctr) ctr->n_short++; return; }
  if (ctr) ctr->n_reads++;
  uint32_t rep = wo_seed_repeats(read_len);
  uint32_t seed_len = rep;
  char stackbuf[1024];
  char* read = read_len <= sizeof(stackbuf) ? stackbuf : (char*)malloc(read_len);
  convert_read(org_read, read_len, ag, read);

  uint32_t cur_max = max_mismatches;
  for (uint32_t seed_i = 0; seed_i < PATTERN_LEN; ++seed_i) {
    int full = heap->size > 0 && heap->size >= heap->max_size;
    if (full && heap->a[0].mismatch == 0 && seed_i) break;       /* paired.cpp:129-131 */
    if (full && heap->a[0].mismatch == 1 && seed_i >= 2) break;  /* paired.cpp:135-137 */
    const char* read_seed = read + seed_i;
    uint32_t h = wo_hash_value(read_seed);
    uint32_t first = ix->counter[h], second = ix->counter[h + 1];
    if (first == second) continue;
    if (ctr) {
      ctr->n_lookups++;
      ctr->sum_log2_bucket += ceil_log2_u64((uint64_t)(second - first) + 1);
    }
    wo_index_region(ix, read_seed, seed_len, &first, &second, ctr);
    if (second - first + 1 > b) { if (ctr) ctr->n_region_over_b++; continue; }

    for (uint32_t j = first; j <= second; ++j) {
      uint32_t gp = ix->index[j];
      uint32_t chr = wo_chrom_id(ix->start_index, ix->n_chr + 1, gp);
      if (gp - ix->start_index[chr] < seed_i) continue;
      gp -= seed_i;
      if (gp + read_len >= ix->start_index[chr + 1]) continue;
      if (ctr) ctr->n_cand++;
      uint32_t mm = count_mismatches(ix, read, read_len, gp, seed_i, rep, cur_max);
      if (mm > max_mismatches) continue;
      wo_cand c; memset(&c, 0, sizeof c);
      c.genome_pos = gp; c.mismatch = mm; c.strand = strand;
      wo_heap_push(heap, c);
      if (heap->size >= heap->max_size) cur_max = heap->a[0].mismatch;
    }
  }
  if (read != stackbuf) free(read);
}

void wo_pe_map_batch(const wo_index* ix, const char* seqs, const uint64_t* offs, uint32_t n,
                     char strand, int ag, uint32_t max_mismatches, uint32_t b, uint32_t top_k,
                     wo_cand* cands, uint32_t* sizes, wo_counters* ctr) {
  for (uint32_t j = 0; j < n; ++j) {
    wo_heap h; h.a = cands + (size_t)j * top_k; h.size = sizes[j]; h.max_size = top_k;
    wo_pe_map_read(ix, seqs + offs[j], (uint32_t)(offs[j + 1] - offs[j]), strand, ag,
                   max_mismatches, b, &h, ctr);
    sizes[j] = h.size;
  }
}

/* ------------------------------------------------------------------------------------ */
/* pairing, paired.cpp:98-104, 296-331, 472-513                                         */
/* ------------------------------------------------------------------------------------ */
static void forward_position(const wo_chroms* g, uint32_t genome_pos, char strand, uint32_t chr,
                             uint32_t read_len, uint32_t* s, uint32_t* e) {
  uint32_t v = genome_pos - g->start_index[chr];
  if (strand != '+') v = g->length[chr] - v - read_len;
  *s = v; *e = v + read_len;
}

int wo_fragment_length(const wo_chroms* g, const wo_cand* r1, uint32_t len1, const wo_cand* r2,
                       uint32_t len2, uint32_t chr1, uint32_t chr2) {
  uint32_t s1, e1, s2, e2;
  forward_position(g, r1->genome_pos, r1->strand, chr1, len1, &s1, &e1);
  forward_position(g, r2->genome_pos, r2->strand, chr2, len2, &s2, &e2);
  return r1->strand == '+' ? (int)(e2 - s1) : (int)(e1 - s2);
}

uint32_t wo_pe_pair(const wo_chroms* g, const wo_cand* R1, uint32_t n1, uint32_t len1,
                    const wo_cand* R2, uint32_t n2, uint32_t len2, uint32_t max_mismatches,
                    int frag_range, int32_t* best_i, int32_t* best_j) {
  uint32_t min_mm = max_mismatches, best_times = 0;
  uint64_t best_pos = 0;
  *best_i = -1; *best_j = -1;
  for (int i = (int)n1 - 1; i >= 0; --i) {
    for (int j = (int)n2 - 1; j >= 0; --j) {
      const wo_cand* r1 = &R1[i];
      const wo_cand* r2 = &R2[j];
      if (r1->strand == r2->strand) continue;
      uint32_t s = r1->mismatch + r2->mismatch;
      if (s > min_mm) break;
      uint32_t c1 = wo_chrom_id(g->start_index, g->n_chr + 1, r1->genome_pos);
      uint32_t c2 = wo_chrom_id(g->start_index, g->n_chr + 1, r2->genome_pos);
      if (c1 != c2) continue;
      int frag = wo_fragment_length(g, r1, len1, r2, len2, c1, c2);
      if (frag <= 0 || frag > frag_range) continue;
      uint64_t cur = ((uint64_t)r1->genome_pos << 32) + r2->genome_pos;
      if (s < min_mm) {
        *best_i = i; *best_j = j; best_times = 1; min_mm = s; best_pos = cur;
      } else if (s == min_mm && cur != best_pos) {
        *best_i = i; *best_j = j; best_times++;
      }
    }
  }
  return best_times;
}

void wo_best_for_single(const wo_cand* ranked, uint32_t n, uint32_t max_mismatches, wo_best* out) {
  memset(out, 0, sizeof *out);
  out->genome_pos = 0; out->times = 0; out->strand = '+'; out->mismatch = max_mismatches;
  for (int i = (int)n - 1; i >= 0; --i) {
    const wo_cand* r = &ranked[i];
    if (r->mismatch < out->mismatch) {
      out->genome_pos = r->genome_pos; out->times = 1; out->strand = r->strand;
      out->mismatch = r->mismatch;
    } else if (r->mismatch == out->mismatch) {
      if (out->genome_pos == r->genome_pos) continue;
      out->genome_pos = r->genome_pos; out->strand = r->strand; out->times++;
    } else {
      break;
    }
  }
}

/* per pair: what MergePairedEndResults (paired.cpp:438-570) takes from the ranked lists */
void wo_pe_result_batch(const wo_chroms* g, const wo_cand* ranked1, const uint32_t* n1, const uint64_t* offs1,
                        const wo_cand* ranked2, const uint32_t* n2, const uint64_t* offs2, uint32_t n, uint32_t top_k,
                        uint32_t max_mismatches, int frag_range, wo_pe_result* out) {
#pragma omp parallel for schedule(dynamic, 256)
  for (long p = 0; p < (long)n; ++p) {
    const wo_cand* r1 = ranked1 + (size_t)p * top_k;
    const wo_cand* r2 = ranked2 + (size_t)p * top_k;
    const uint32_t len1 = (uint32_t)(offs1[p + 1] - offs1[p]), len2 = (uint32_t)(offs2[p + 1] - offs2[p]);
    wo_pe_result* o = &out[p];
    memset(o, 0, sizeof *o);
    int32_t bi, bj;
    o->best_times = wo_pe_pair(g, r1, n1[p], len1, r2, n2[p], len2, max_mismatches, frag_range, &bi, &bj);
    o->best_i = bi; o->best_j = bj;
    if (o->best_times >= 1) {
      o->c1 = r1[bi]; o->c2 = r2[bj];
      memset(o->c1.pad, 0, sizeof o->c1.pad); memset(o->c2.pad, 0, sizeof o->c2.pad);
      const uint32_t c1 = wo_chrom_id(g->start_index, g->n_chr + 1, r1[bi].genome_pos);
      const uint32_t c2 = wo_chrom_id(g->start_index, g->n_chr + 1, r2[bj].genome_pos);
      o->frag_len = wo_fragment_length(g, &r1[bi], len1, &r2[bj], len2, c1, c2);
    }
    wo_best_for_single(r1, n1[p], max_mismatches, &o->single1);
    wo_best_for_single(r2, n2[p], max_mismatches, &o->single2);
  }
}

/* ------------------------------------------------------------------------------------ */
/* index builder, reference.cpp:192-300                                                 */
/* ------------------------------------------------------------------------------------ */
int wo_bucket_cmp(const char* seq, uint32_t n_chr, const uint32_t* start_index, uint32_t p1,
                  uint32_t p2) {
  uint32_t c1 = wo_chrom_id(start_index, n_chr + 1, p1);
  uint32_t c2 = wo_chrom_id(start_index, n_chr + 1, p2);
  uint32_t l1 = start_index[c1 + 1] - p1, l2 = start_index[c2 + 1] - p2;
  for (uint32_t j = KEY_WEIGHT; j < 60; ++j) {
    uint32_t cp = cared_position(j);
    if (cp >= l2) return (cp >= l1) ? 0 : 1;   /* "not less"; equal if both ran out */
    if (cp >= l1) return -1;
    char a = seq[(uint64_t)p1 + cp], b = seq[(uint64_t)p2 + cp];
    if (a < b) return -1;
    if (a > b) return 1;
  }
  return 0;
}

/* std::sort as GCC 13's libstdc++ implements it (bits/stl_algo.h:1941-1952 __sort ->
 * __introsort_loop 1918-1936 with median-of-three __unguarded_partition_pivot 1871-1900,
 * heap-sort fallback 1905-1913 + bits/stl_heap.h:135-147,224-267,340-362,419-427, then
 * __final_insertion_sort 1792-1866 with threshold 16).  The reference sorts each bucket with it
 * (reference.cpp:296-298); the sort is unstable, so the order of suffixes that compare equal
 * in the .dbindex file is a product of this exact sequence of compares and swaps applied to the
 * bucket's initial ascending-position arrangement.  Restated here so that the oracle's index
 * is the reference's index byte for byte (third-party algorithm). */
typedef struct { const char* seq; const uint32_t* starts; uint32_t n_chr; } sort_ctx;
static int sl(const sort_ctx* c, uint32_t a, uint32_t b) {   /* SortHashTableBucketCMP(a, b) */
  return wo_bucket_cmp(c->seq, c->n_chr, c->starts, a, b) < 0;
}
static void ss_swap(uint32_t* v, long a, long b) { uint32_t t = v[a]; v[a] = v[b]; v[b] = t; }

static void ss_push_heap(const sort_ctx* c, uint32_t* v, long hole, long top, uint32_t value) {
  long parent = (hole - 1) / 2;
  while (hole > top && sl(c, v[parent], value)) { v[hole] = v[parent]; hole = parent; parent = (hole - 1) / 2; }
  v[hole] = value;
}
static void ss_adjust_heap(const sort_ctx* c, uint32_t* v, long hole, long len, uint32_t value) {
  const long top = hole;
  long child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (sl(c, v[child], v[child - 1])) child--;
    v[hole] = v[child]; hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    v[hole] = v[child - 1]; hole = child - 1;
  }
  ss_push_heap(c, v, hole, top, value);
}
static void ss_heap_sort(const sort_ctx* c, uint32_t* v, long len) {   /* __partial_sort(f, l, l) */
  if (len >= 2)
    for (long parent = (len - 2) / 2;; --parent) {
      ss_adjust_heap(c, v, parent, len, v[parent]);
      if (parent == 0) break;
    }
  for (long last = len; last > 1;) {   /* __sort_heap: __pop_heap(first, last, last) */
    --last;
    uint32_t value = v[last];
    v[last] = v[0];
    ss_adjust_heap(c, v, 0, last, value);
  }
}
static void ss_linear_insert(const sort_ctx* c, uint32_t* v, long last) {   /* unguarded */
  uint32_t val = v[last];
  long next = last - 1;
  while (sl(c, val, v[next])) { v[last] = v[next]; last = next; --next; }
  v[last] = val;
}
static void ss_insertion_sort(const sort_ctx* c, uint32_t* v, long first, long last) {
  if (first == last) return;
  for (long i = first + 1; i != last; ++i) {
    if (sl(c, v[i], v[first])) {
      uint32_t val = v[i];
      memmove(v + first + 1, v + first, sizeof(uint32_t) * (size_t)(i - first));
      v[first] = val;
    } else {
      ss_linear_insert(c, v, i);
    }
  }
}
static void ss_introsort_loop(const sort_ctx* c, uint32_t* v, long first, long last, long depth_limit) {
  while (last - first > 16) {
    if (depth_limit == 0) { ss_heap_sort(c, v + first, last - first); return; }
    --depth_limit;
    /* __move_median_to_first(first, first + 1, mid, last - 1) */
    long a = first + 1, b = first + (last - first) / 2, d = last - 1;
    if (sl(c, v[a], v[b])) {
      if (sl(c, v[b], v[d])) ss_swap(v, first, b);
      else if (sl(c, v[a], v[d])) ss_swap(v, first, d);
      else ss_swap(v, first, a);
    } else if (sl(c, v[a], v[d])) ss_swap(v, first, a);
    else if (sl(c, v[b], v[d])) ss_swap(v, first, d);
    else ss_swap(v, first, b);
    /* __unguarded_partition(first + 1, last, pivot = first) */
    long lo = first + 1, hi = last;
    for (;;) {
      while (sl(c, v[lo], v[first])) ++lo;
      --hi;
      while (sl(c, v[first], v[hi])) --hi;
      if (!(lo < hi)) break;
      ss_swap(v, lo, hi);
      ++lo;
    }
    ss_introsort_loop(c, v, lo, last, depth_limit);
    last = lo;
  }
}
void wo_std_sort_bucket(const char* seq, uint32_t n_chr, const uint32_t* start_index, uint32_t* v, uint32_t n) {
  if (n == 0) return;
  sort_ctx c = {seq, start_index, n_chr};
  long lg = 0;
  for (uint32_t x = n; x > 1; x >>= 1) ++lg;
  ss_introsort_loop(&c, v, 0, (long)n, 2 * lg);
  if (n > 16) {
    ss_insertion_sort(&c, v, 0, 16);
    for (long i = 16; i != (long)n; ++i) ss_linear_insert(&c, v, i);
  } else {
    ss_insertion_sort(&c, v, 0, (long)n);
  }
}

uint32_t wo_build_index(const char* seq, uint64_t genome_len, uint32_t n_chr,
                        const uint32_t* start_index, uint32_t* counter, uint32_t* index) {
  (void)genome_len;
  const uint32_t n_keys = 1u << (2 * KEY_WEIGHT);
  memset(counter, 0, sizeof(uint32_t) * ((size_t)n_keys + 1));
  /* CountBucketSize, reference.cpp:192-229 */
  for (uint32_t c = 0; c < n_chr; ++c) {
    if (start_index[c + 1] - start_index[c] < MIN_SEED_LEN) continue;
    uint32_t end = start_index[c + 1] - MIN_SEED_LEN;
    for (uint32_t j = start_index[c]; j < end; ++j) counter[wo_hash_value(seq + j)]++;
  }
  uint8_t* erased = (uint8_t*)calloc(n_keys, 1);
  for (uint32_t k = 0; k < n_keys; ++k)
    if (counter[k] >= 500000) { counter[k] = 0; erased[k] = 1; }
  /* exclusive prefix sums; counter[n_keys] = index_size */
  uint32_t run = 0;
  for (uint32_t k = 0; k < n_keys; ++k) { uint32_t c = counter[k]; counter[k] = run; run += c; }
  counter[n_keys] = run;
  uint32_t index_size = run;
  /* HashToBucket, reference.cpp:231-256 (positions ascending inside a bucket) */
  uint32_t* cursor = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n_keys);
  memcpy(cursor, counter, sizeof(uint32_t) * (size_t)n_keys);
  for (uint32_t c = 0; c < n_chr; ++c) {
    if (start_index[c + 1] - start_index[c] < MIN_SEED_LEN) continue;
    uint32_t end = start_index[c + 1] - MIN_SEED_LEN;
    for (uint32_t j = start_index[c]; j < end; ++j) {
      uint32_t h = wo_hash_value(seq + j);
      if (erased[h]) continue;
      index[cursor[h]++] = j;
    }
  }
  free(cursor); free(erased);
  /* SortHashTableBucket, reference.cpp:290-300 */
  for (uint32_t k = 0; k < n_keys; ++k) {
    uint32_t n = counter[k + 1] - counter[k];
    if (n > 1) wo_std_sort_bucket(seq, n_chr, start_index, index + counter[k], n);
  }
  return index_size;
}

/* ------------------------------------------------------------------------------------ */
/* adaptor clipping, util.hpp:189-217                                                   */
/* ------------------------------------------------------------------------------------ */
static size_t head_similarity(const char* s, size_t len, size_t pos, const char* ad, size_t ad_len) {
  size_t lim = len - pos;
  if (ad_len < lim) lim = ad_len;
  if (14 < lim) lim = 14;
  size_t c = 0;
  for (size_t i = 0; i < lim; ++i) c += (s[pos + i] == ad[i]);
  return c;
}

size_t wo_clip_adaptor(const char* ad, size_t ad_len, char* s, size_t len) {
  /* size_t arithmetic wraps for len < 14 exactly like the reference (UB there; callers of
   * the oracle keep len >= 14) */
  size_t lim1 = len - 14 + 1;
  for (size_t i = 0; i < lim1; ++i)
    if (head_similarity(s, len, i, ad, ad_len) >= 11) { memset(s + i, 'N', len - i); return len - i; }
  size_t lim2 = len - 5 + 1;
  for (size_t i = lim1; i < lim2; ++i)
    if (head_similarity(s, len, i, ad, ad_len) >= len - i - 1) { memset(s + i, 'N', len - i); return len - i; }
  return 0;
}
