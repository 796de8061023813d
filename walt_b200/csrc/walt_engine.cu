// walt_engine.cu -- B200 (sm_100a) mapping engine behind the C ABI of include/walt_b200.h.
//
// Replaces the strand-loop bodies of the reference's batch drivers
//   ReadIndex + #pragma omp parallel for { SingleEndMapping }   src/walt/mapping.cpp:486-500
//   ReadIndex + #pragma omp parallel for { PairEndMapping }     src/walt/paired.cpp:642-672
// and the pairing loop of MergePairedEndResults                 src/walt/paired.cpp:472-513
// with CUDA kernels over an index that is loaded ONCE into HBM.  The per-lane logic lives in
// walt_core.cuh (shared with the CPU fiber harness of tests/emu); this file owns device
// memory, index residency (2-bit genome, position array, base-3 prefix table, taint list),
// the kernels' warp policy and launch geometry, and the double-buffered host<->device batch
// pipeline.  There is no CPU fallback: without a CUDA device every entry point fails.
#include "walt_engine.cuh"
#include "walt_hostscan.h"

#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <tuple>

using namespace waltcore;

namespace waltb200 {

static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
int fail(int code, const std::string& msg) { g_error = msg; return code; }

// ------------------------------------------------------------------------------------------
// hardware warp policy for walt_core.cuh
// ------------------------------------------------------------------------------------------
// A group of WD lanes (8, 16 or 32) of a hardware warp owns one read; the groups of a warp run
// independently (Volta+ independent thread scheduling), every collective names only the group's
// lanes.  Ballots and shuffle sources are group-relative.
template <uint32_t WD>
struct HwGroup {
  static constexpr uint32_t WIDTH = WD;
  uint32_t base;    // first lane of the group inside the warp
  uint32_t gmask;   // the group's lanes
  __device__ __forceinline__ HwGroup() {
    const uint32_t l = threadIdx.x & 31u;
    base = l & ~(WD - 1u);
    gmask = WD == 32u ? 0xFFFFFFFFu : (((1u << WD) - 1u) << base);
  }
  __device__ __forceinline__ uint32_t lane() const { return (threadIdx.x & 31u) - base; }
  __device__ __forceinline__ uint32_t ballot(bool p) const { return (__ballot_sync(gmask, p) & gmask) >> base; }
  __device__ __forceinline__ uint32_t shfl(uint32_t v, int src) const { return __shfl_sync(gmask, v, (int)base + src); }
  __device__ __forceinline__ uint32_t reduce_add(uint32_t v) const { return __reduce_add_sync(gmask, v); }
  __device__ __forceinline__ uint32_t reduce_min(uint32_t v) const { return __reduce_min_sync(gmask, v); }
  __device__ __forceinline__ void sync() const { __syncwarp(gmask); }
};

SubIndexView DeviceSubIndex::view(int which) const {
  SubIndexView v;
  v.genome = genome; v.entries = entries; v.table = table; v.taint_bits = taint_bits;
  // one allocation: filter bits | rank per word | start per distinct key (+1) | pos | len | slot
  const size_t words = (N_KEY12 + 31u) / 32u;
  v.taint_rank = taint_bits + words; v.taint_start = taint_bits + 2 * words;
  v.taint_pos = v.taint_start + n_taint_keys + 1u; v.taint_len = v.taint_pos + n_taint; v.taint_slot = v.taint_len + n_taint;
  v.n_taint = n_taint; v.index_size = index_size; v.depth = depth; v.ag = which >= 2 ? 1u : 0u;
  return v;
}

void DeviceSubIndex::release() {
  cudaFree(genome); cudaFree(index); cudaFree(entries); cudaFree(table); cudaFree(taint_bits);
  *this = DeviceSubIndex();
}

ChromView chrom_view(const walt_engine* e) {
  ChromView c; c.starts = e->d_starts; c.n_chr = e->n_chr; c.genome_len = e->genome_len;
  return c;
}

int ensure_device(walt_engine* e) {
  WALT_CUDA_TRY(cudaSetDevice(e->device));
  return WALT_OK;
}

uint32_t choose_depth(uint32_t index_size) {
  uint32_t d = KEY_WEIGHT;
  uint64_t p = N_KEY12;
  while (d < MAX_DEPTH && p < index_size) { p *= 3; ++d; }
  return d;
}

// ------------------------------------------------------------------------------------------
// index residency kernels
// ------------------------------------------------------------------------------------------
// 32 ASCII bases -> one packed word.  `first_base` is a multiple of 32; genome word index is
// (first_base + PAD_BASES) / 32 + w.  bad[0] counts bytes outside the sub-index's alphabet.
__global__ void pack_ascii_kernel(const uint8_t* __restrict__ ascii, uint64_t n_bases, uint64_t* __restrict__ words,
                                  uint32_t forbidden_code, uint32_t* __restrict__ bad) {
  const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t b0 = w * 32u;
  if (b0 >= n_bases) return;
  uint64_t out = 0;
  uint32_t nbad = 0;
  const uint32_t cnt = (uint32_t)min((uint64_t)32, n_bases - b0);
  if (cnt == 32u) {
    const uint4* p = reinterpret_cast<const uint4*>(ascii + b0);
    uint4 q[2] = {p[0], p[1]};
    const uint32_t* u = reinterpret_cast<const uint32_t*>(q);
#pragma unroll
    for (uint32_t i = 0; i < 8; ++i) {
#pragma unroll
      for (uint32_t j = 0; j < 4; ++j) {
        uint32_t c = (u[i] >> (8u * j)) & 0xFFu;
        uint32_t code = ascii_code(c);
        nbad += (!ascii_is_acgt(c) || code == forbidden_code) ? 1u : 0u;
        out = (out << 2) | code;
      }
    }
  } else {
    for (uint32_t i = 0; i < 32u; ++i) {
      uint32_t code = 0;
      if (i < cnt) {
        uint32_t c = ascii[b0 + i];
        code = ascii_code(c);
        nbad += (!ascii_is_acgt(c) || code == forbidden_code) ? 1u : 0u;
      }
      out = (out << 2) | code;
    }
  }
  words[w] = out;
  if (nbad) atomicAdd(bad, nbad);
}

// per slot: table key (first `depth` seed characters) and the {position, fingerprint} entry
// ... and, for the few positions on the taint list, the slot they sit in (literal_index_region)
__global__ void table_keys_kernel(const uint64_t* __restrict__ genome, const uint32_t* __restrict__ index,
                                  uint32_t index_size, uint32_t depth, uint32_t ag, ChromView cv, Pow3 p3,
                                  uint32_t* __restrict__ keys, Entry* __restrict__ entries, SubIndexView tv,
                                  uint32_t* __restrict__ taint_slot) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= index_size) return;
  const uint32_t e = index[i];
  const uint32_t chr = chrom_of(cv.starts, cv.n_chr, e);
  const uint32_t ce = cv.starts[chr + 1u];
  keys[i] = entry_table_key(genome, e, ce, depth, ag != 0u, p3);
  Entry en; en.pos = e; en.fp = entry_fingerprint(genome, e, ce, depth, ag != 0u, p3);
  entries[i] = en;
  if (ce - e <= TAINT_SPAN) {
    const uint32_t key12 = entry_key12(genome, e, ag != 0u, p3);
    if ((tv.taint_bits[key12 >> 5] >> (key12 & 31u)) & 1u) {
      uint32_t t0, t1;
      taint_slots(tv, key12, t0, t1);
      for (uint32_t t = t0; t < t1; ++t)
        if (tv.taint_pos[t] == e) taint_slot[t] = (uint32_t)i;
    }
  }
}

// table[k] = first slot whose key >= k, for k in [0, n_keys]; slot i owns (key[i-1], key[i]]
__global__ void table_fill_kernel(const uint32_t* __restrict__ keys, uint32_t index_size, uint32_t n_keys,
                                  uint32_t* __restrict__ table, uint32_t* __restrict__ unsorted) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > index_size) return;
  const uint32_t cur = i < index_size ? keys[i] : n_keys;
  uint64_t from = 0;
  if (i > 0) {
    const uint32_t prev = keys[i - 1];
    if (cur < prev) { atomicAdd(unsorted, 1u); return; }
    from = (uint64_t)prev + 1u;
  }
  for (uint64_t k = from; k <= cur; ++k) table[k] = (uint32_t)i;
}

int pack_ascii_device(const uint8_t* d_ascii, uint64_t n, uint64_t* d_words, uint32_t forbidden, uint32_t* d_bad) {
  const uint64_t words = (n + 31u) / 32u;
  const uint32_t T = 256;
  pack_ascii_kernel<<<(uint32_t)((words + T - 1) / T), T>>>(d_ascii, n, d_words, forbidden, d_bad);
  WALT_CUDA_TRY(cudaGetLastError());
  return WALT_OK;
}

int alloc_packed_genome(walt_engine* e, DeviceSubIndex& s) {
  s.genome_words = ((uint64_t)e->genome_len + PAD_BASES + 31u) / 32u + TAIL_PAD_WORDS;
  WALT_CUDA_TRY(cudaMalloc(&s.genome, s.genome_words * 8u));
  WALT_CUDA_TRY(cudaMemset(s.genome, 0, s.genome_words * 8u));
  s.bytes += s.genome_words * 8u;
  return WALT_OK;
}

int finalize_subindex(walt_engine* e, int which) {
  DeviceSubIndex& s = e->sub[which];
  s.depth = e->force_depth > 0 ? (uint32_t)e->force_depth : choose_depth(s.index_size);
  if (s.depth < KEY_WEIGHT || s.depth > MAX_DEPTH) return fail(WALT_EINVAL, "table depth out of range");
  if (e->force_depth == 0) {
    // keep the table within what is left of HBM after the entry array and the key scratch;
    // a shallower table only means one more narrowing step on the fingerprints
    size_t free_b = 0, total_b = 0;
    WALT_CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    const size_t fixed = (size_t)s.index_size * 12u + (256u << 20);
    while (s.depth > KEY_WEIGHT && ((size_t)e->pow3.v[s.depth] + 1u) * 4u + fixed > free_b) --s.depth;
  }
  const uint32_t n_keys = e->pow3.v[s.depth];
  WALT_CUDA_TRY(cudaMalloc(&s.table, ((size_t)n_keys + 1u) * 4u));
  s.bytes += ((size_t)n_keys + 1u) * 4u;
  WALT_CUDA_TRY(cudaMalloc(&s.entries, ((size_t)s.index_size + 64u) * sizeof(Entry)));
  WALT_CUDA_TRY(cudaMemset(s.entries + s.index_size, 0, 64u * sizeof(Entry)));
  s.bytes += ((size_t)s.index_size + 64u) * sizeof(Entry);
  // taint list first (host; <= 112 positions per chromosome): view() needs its pointers
  {
    std::vector<std::tuple<uint32_t, uint32_t, uint32_t>> t;
    std::vector<uint64_t> win;
    for (uint32_t c = 0; c < e->n_chr; ++c) {
      const uint32_t cs = e->starts[c], ce = e->starts[c + 1];
      if (ce - cs < MIN_SEED_LEN + 1u) continue;
      const uint32_t span = std::min(ce - cs, TAINT_SPAN);
      // packed words covering bases [ce - span, ce + 64)
      const uint64_t p0 = (uint64_t)(ce - span) + PAD_BASES;
      const uint64_t w0 = p0 >> 5, w1 = std::min<uint64_t>(s.genome_words, (((uint64_t)ce + PAD_BASES + 64u) >> 5) + 1u);
      win.assign(w1 - w0, 0);
      WALT_CUDA_TRY(cudaMemcpy(win.data(), s.genome + w0, (w1 - w0) * 8u, cudaMemcpyDeviceToHost));
      // entry_key12 indexes from the genome base; hand it a shifted pointer
      const uint64_t* base = win.data() - w0;
      for (uint32_t d = MIN_SEED_LEN + 1u; d <= span; ++d) {
        const uint32_t pos = ce - d;
        t.emplace_back(entry_key12(base, pos, which >= 2, e->pow3), pos, d);
      }
    }
    std::sort(t.begin(), t.end());
    s.n_taint = (uint32_t)t.size();
    std::vector<uint32_t> keys12(s.n_taint), bits, rank, start;
    for (uint32_t i = 0; i < s.n_taint; ++i) keys12[i] = std::get<0>(t[i]);
    build_taint_directory(keys12, bits, rank, start);
    s.n_taint_keys = (uint32_t)start.size() - 1u;
    std::vector<uint32_t> flat(bits);
    flat.insert(flat.end(), rank.begin(), rank.end());
    flat.insert(flat.end(), start.begin(), start.end());
    for (uint32_t i = 0; i < s.n_taint; ++i) flat.push_back(std::get<1>(t[i]));
    for (uint32_t i = 0; i < s.n_taint; ++i) flat.push_back(std::get<2>(t[i]));
    flat.insert(flat.end(), (size_t)s.n_taint + 1u, 0xFFFFFFFFu);   // slots: filled by table_keys_kernel
    s.taint_words = flat.size();
    WALT_CUDA_TRY(cudaMalloc(&s.taint_bits, flat.size() * 4u));
    WALT_CUDA_TRY(cudaMemcpy(s.taint_bits, flat.data(), flat.size() * 4u, cudaMemcpyHostToDevice));
    s.bytes += flat.size() * 4u;
  }
  uint32_t* keys = nullptr;
  WALT_CUDA_TRY(cudaMalloc(&keys, ((size_t)s.index_size + 1u) * 4u));
  const uint32_t T = 256;
  if (s.index_size)
    table_keys_kernel<<<(uint32_t)(((uint64_t)s.index_size + T - 1) / T), T>>>(
        s.genome, s.index, s.index_size, s.depth, which >= 2 ? 1u : 0u, chrom_view(e), e->pow3, keys, s.entries,
        s.view(which), const_cast<uint32_t*>(s.view(which).taint_slot));
  WALT_CUDA_TRY(cudaMemset(e->d_flags + 2, 0, 4));
  table_fill_kernel<<<(uint32_t)(((uint64_t)s.index_size + 1u + T - 1) / T), T>>>(keys, s.index_size, n_keys, s.table,
                                                                                 e->d_flags + 2);
  uint32_t unsorted = 0;
  cudaError_t ce = cudaMemcpy(&unsorted, e->d_flags + 2, 4, cudaMemcpyDeviceToHost);
  cudaFree(keys);
  cudaFree(s.index); s.index = nullptr;   // positions now live in the entry array
  if (ce != cudaSuccess) return fail(WALT_ECUDA, std::string("table build: ") + cudaGetErrorString(ce));
  if (unsorted)
    return fail(WALT_EFORMAT, "index[] is not in makedb order (" + std::to_string(unsorted) + " inversions)");
  s.loaded = true;
  return WALT_OK;
}

// ------------------------------------------------------------------------------------------
// mapping kernels: one warp owns a read
// ------------------------------------------------------------------------------------------
constexpr uint32_t WARPS_PER_BLOCK = 8;
constexpr uint32_t BLOCK_THREADS = WARPS_PER_BLOCK * 32;
constexpr uint32_t MIN_BLOCKS_PER_SM = 4;   // caps the kernels at 64 registers (the out-of-line replay may spill)

// device buffers of a launch that parks reads (see "parked reads" in walt_core.cuh)
struct ParkView {
  uint32_t* recs;       // PARK_WORDS words per parked read with a record (first rec_cap of them)
  uint32_t* desc;       // two words per verification block: record number, lookup | block << 3
  uint8_t* bytes;       // 32 per verification block: what verify_kernel leaves for the fold
  uint32_t* list;       // read numbers of the parked reads WITHOUT a record (MAP_TAKE maps them)
  uint32_t rec_cap, cap_blocks;
};

struct SeArgs {
  SubIndexView ix[2];   // '+' then '-' sub-index
  ChromView cv;
  Pow3 p3;
  MapConfig cfg;
  const char* seqs;
  const uint64_t* offs;   // n + 1, absolute; seqs[0] is byte `seq_base` (unused if uniform_len)
  uint64_t seq_base;
  uint32_t uniform_len;   // != 0: every read has this length, read r starts at offset seq_base + r * uniform_len
  uint32_t read_base;     // packed input: read r of this launch is read `read_base + r` of the buffer
  uint32_t n;
  uint32_t nw_max;        // scratch stride (words) for the longest read
  uint32_t ag;
  uint32_t max_mismatches;
  walt_best* out;
  uint32_t* flags;        // [0] non-ACGT
  uint32_t* queue;        // QUEUE_WORDS words, zeroed before launch: [0] work-queue head, [1] parked reads with a record,
                          // [2] queue head of MAP_TAKE, [3] reads on its list, [5] verification blocks reserved,
                          // [6] / [7] queue heads of verify_kernel / the fold kernels
  ParkView park;          // where parked reads go (MAP_PARK writes, the kernels behind it read)
  unsigned long long* counters;  // optional
};

// Kernel modes.  MAP_ALL: every read is finished where it is.  MAP_PARK: a read whose ordered fold
// reaches a lookup that needs more than one lane (a long fingerprint run = repeats) is parked, so that the
// sub-warp groups never leave the converged fast path: its lanes write a record of what they found and
// reserve verification blocks for its runs (DevPark, park_read).  verify_kernel then verifies every block,
// a warp per block and no order; se_fold_kernel / pe_fold_kernel fold every parked read from its record.
// MAP_TAKE (a whole warp per read, from scratch) maps the parked reads that got no record (no room, reads
// longer than WIDE_MAX_READ) -- and all of them when the flat path is switched off.
enum : int { MAP_ALL = 0, MAP_PARK = 1, MAP_TAKE = 2 };
constexpr uint32_t TAKE_BLOCKS_PER_SM = 2;   // the take-over kernels get 128 registers

// The groups of a warp take consecutive reads with one queue ticket and walk the read loop
// together, so the warp stays converged through the common phases (pack, keys, table and entry
// loads, fold) and only splits where the data makes it (candidate windows, cooperative replays).
#ifndef WALT_TICKET_ROUNDS
#define WALT_TICKET_ROUNDS 1
#endif
// One queue ticket covers TICKET_ROUNDS rounds of the warp's groups.  Measured on configs[1]: 4 rounds
// per ticket (a quarter of the same-address atomics) is 3 % SLOWER than 1 (8.94 vs 8.67 ms per 10 M
// reads): the queue is not a bottleneck and the extra loop state costs registers.
constexpr uint32_t TICKET_ROUNDS = WALT_TICKET_ROUNDS;

template <uint32_t WD>
__device__ __forceinline__ uint32_t next_ticket(uint32_t* queue) {
  constexpr uint32_t GROUPS = 32u / WD;
  uint32_t r = 0;
  if ((threadIdx.x & 31u) == 0u) r = atomicAdd(queue, GROUPS * TICKET_ROUNDS);
  return __shfl_sync(0xFFFFFFFFu, r, 0);
}

// Where read r of a launch starts and how long it is.  ASCII: byte offs[r] - seq_base of `seqs`.
// 2-bit packed (walt_pack_reads, include/walt_host.h): read j of a batch occupies ceil(len / 4)
// bytes from byte (offs[j] >> 2) + j, so one offsets array addresses both forms.
template <bool PACKED, class Args>
__device__ __forceinline__ const char* read_at(const Args& a, uint32_t r, uint32_t& len) {
  uint64_t o0;
  if (a.uniform_len) {
    len = a.uniform_len; o0 = a.seq_base + (uint64_t)r * len;
  } else {
    o0 = a.offs[r]; len = (uint32_t)(a.offs[r + 1] - o0);
  }
  if (PACKED) return a.seqs + ((o0 >> 2) - (a.seq_base >> 2) + r + a.read_base);
  return a.seqs + (o0 - a.seq_base);
}

// Work counters and the non-ACGT flag leave the kernel through ONE set of global atomics per
// CTA: every group adds into shared memory, thread 0 publishes after the barrier.  (One set per
// group would be ~57 000 same-address atomics arriving together in the tail of every launch --
// a fixed ~0.2 ms that dominated 262 144-read chunk launches.)
struct BlockTally { unsigned long long c[3]; unsigned int bad; };

__device__ __forceinline__ void tally_init(BlockTally& t) {
  if (threadIdx.x == 0) { t.c[0] = t.c[1] = t.c[2] = 0ull; t.bad = 0u; }
  __syncthreads();
}

template <uint32_t WD>
__device__ __forceinline__ void flush_counters(const HwGroup<WD>& w, const Counters& ctr, bool bad, BlockTally& t,
                                               uint32_t* flags, unsigned long long* counters) {
  const uint32_t a = w.reduce_add(ctr.lookups), b = w.reduce_add(ctr.candidates), c = w.reduce_add(ctr.literal);
  if (w.lane() == 0) {
    if (bad) atomicOr(&t.bad, 1u);
    if (a) atomicAdd(&t.c[0], (unsigned long long)a);
    if (b) atomicAdd(&t.c[1], (unsigned long long)b);
    if (c) atomicAdd(&t.c[2], (unsigned long long)c);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (t.bad) atomicOr(flags, 1u);
    if (counters) {
      if (t.c[0]) atomicAdd(counters + 0, t.c[0]);
      if (t.c[1]) atomicAdd(counters + 1, t.c[1]);
      if (t.c[2]) atomicAdd(counters + 2, t.c[2]);
    }
  }
}

// Parking policy of the MAP_PARK kernels (walt_core.cuh: park_read).  claim(): a record and `total`
// verification blocks for read r, or -- no room, or a read the flat path cannot take -- a place on
// the MAP_TAKE list.
struct DevPark {
  static constexpr bool ENABLED = true;
  ParkView v;
  uint32_t* queue;
  uint32_t r;
  template <class W>
  __device__ __forceinline__ uint32_t* claim(W& w, uint32_t total, bool legacy, uint32_t& blk0, uint32_t& t) {
    uint32_t at = 0xFFFFFFFFu, b0 = 0u;
    if (w.lane() == 0u) {
      if (!legacy && v.rec_cap) {
        at = atomicAdd(queue + 1, 1u);
        if (at < v.rec_cap) {
          b0 = atomicAdd(queue + 5, total);
          if ((uint64_t)b0 + total > v.cap_blocks) {   // no room for its blocks: a dead record, dead descriptors
            v.recs[(size_t)at * PARK_WORDS] = 0xFFFFFFFFu; at = 0xFFFFFFFFu;
            for (uint32_t blk = b0; blk < v.cap_blocks && blk - b0 < total; ++blk) v.desc[2u * (size_t)blk] = 0xFFFFFFFFu;
          }
        } else {
          at = 0xFFFFFFFFu;
        }
      }
      if (at == 0xFFFFFFFFu) v.list[atomicAdd(queue + 3, 1u)] = r;
      else v.recs[(size_t)at * PARK_WORDS] = r;
    }
    at = w.shfl(at, 0); blk0 = w.shfl(b0, 0); t = at;
    return at == 0xFFFFFFFFu ? nullptr : v.recs + (size_t)at * PARK_WORDS;
  }
  __device__ __forceinline__ uint32_t* block_desc(uint32_t blk) const { return v.desc + 2u * (size_t)blk; }
};

template <uint32_t WD, bool PACKED, int MODE>
__global__ void __launch_bounds__(BLOCK_THREADS, MODE >= MAP_TAKE ? TAKE_BLOCKS_PER_SM : MIN_BLOCKS_PER_SM)
se_map_kernel(const __grid_constant__ SeArgs a) {
  static_assert(MODE < MAP_TAKE || WD == 32u, "parked reads are taken over by whole warps");
  extern __shared__ uint64_t smem[];
  __shared__ BlockTally tally;
  tally_init(tally);
  HwGroup<WD> w;
  const uint32_t lane = w.lane();
  const uint32_t group_in_block = threadIdx.x / WD;
  ReadScratch sc = carve_scratch(smem + (size_t)group_in_block * scratch_words(a.nw_max), a.nw_max);
  uint32_t cached_len = 0;
  Counters ctr{0u, 0u, 0u};
  bool bad = false;
  uint32_t n = a.n;
  if (MODE == MAP_TAKE) {
    n = *reinterpret_cast<volatile const uint32_t*>(a.queue + 3);   // the parked reads without a record
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.counters && n) atomicAdd(a.counters + 3, (unsigned long long)n);
  }
  uint32_t* const queue = MODE == MAP_TAKE ? a.queue + 2 : a.queue;
  for (uint32_t round = TICKET_ROUNDS, base = 0;; ++round) {
    if (round == TICKET_ROUNDS) { base = next_ticket<WD>(queue); round = 0; }
    const uint32_t first = base + round * (32u / WD);
    if (first >= n) break;                              // warp-uniform: past the batch
    uint32_t r = first + (threadIdx.x & 31u) / WD;
    if (r < n) {
      if (MODE == MAP_TAKE) r = a.park.list[r];
      uint32_t len;
      const char* seq = read_at<PACKED>(a, r, len);
      BestState st;
      MapStatus ms;
      if (MODE == MAP_PARK)
        ms = map_read_se<HwGroup<WD>, PACKED, DevPark>(w, a.ix, a.cv, a.p3, a.cfg, seq, len, a.ag != 0u, a.max_mismatches, sc,
                                                        cached_len, st, ctr, DevPark{a.park, a.queue, r});
      else
        ms = map_read_se<HwGroup<WD>, PACKED>(w, a.ix, a.cv, a.p3, a.cfg, seq, len, a.ag != 0u, a.max_mismatches, sc,
                                               cached_len, st, ctr);
      bad |= ms == MAP_BAD;
      if (lane == 0 && ms < MAP_PARKED) {
        uint4 o;
        o.x = st.pos; o.y = st.times; o.z = st.mm; o.w = st.strand & 0xFFu;
        *reinterpret_cast<uint4*>(a.out + r) = o;
      }
    }
    __syncwarp();
  }
  flush_counters(w, ctr, bad, tally, a.flags, a.counters);
}

struct PeArgs {
  SubIndexView ix[2];
  ChromView cv;
  Pow3 p3;
  MapConfig cfg;
  const char* seqs;
  const uint64_t* offs;
  uint64_t seq_base;
  uint32_t uniform_len;   // see SeArgs
  uint32_t read_base;
  uint32_t n;
  uint32_t nw_max;
  uint32_t ag;
  uint32_t max_mismatches;
  uint32_t top_k;
  walt_cand* ranked;      // n * top_k, worst first
  uint32_t* n_ranked;
  HeapEntry* log;         // two-phase form: log_slots heap-changing candidates per read, in order
  uint32_t* n_log;
  uint32_t log_slots;
  uint32_t zero_fill;     // the ranked lists travel to the host: define (zero) their unused slots
  uint32_t* flags;
  uint32_t* queue;        // see SeArgs
  ParkView park;
  unsigned long long* counters;
};

// Two-phase form, first phase: PairEndMapping (paired.cpp:106-201) for one mate batch with the
// heap-changing candidates logged (LogSink, walt_core.cuh); pe_heap_kernel finishes the job.
template <uint32_t WD, bool PACKED, int MODE>
__global__ void __launch_bounds__(BLOCK_THREADS, MODE >= MAP_TAKE ? TAKE_BLOCKS_PER_SM : MIN_BLOCKS_PER_SM)
pe_log_kernel(const __grid_constant__ PeArgs a) {
  static_assert(MODE < MAP_TAKE || WD == 32u, "parked reads are taken over by whole warps");
  extern __shared__ uint64_t smem[];
  __shared__ BlockTally tally;
  tally_init(tally);
  HwGroup<WD> w;
  const uint32_t lane = w.lane();
  const uint32_t group_in_block = threadIdx.x / WD;
  const uint32_t per_group = scratch_words(a.nw_max) + (LOG_MAX_MM + 2u) / 2u;   // + the mismatch histogram
  uint64_t* mine = smem + (size_t)group_in_block * per_group;
  ReadScratch sc = carve_scratch(mine, a.nw_max);
  uint32_t* hist = reinterpret_cast<uint32_t*>(mine + scratch_words(a.nw_max));
  uint32_t cached_len = 0;
  Counters ctr{0u, 0u, 0u};
  bool bad = false;
  uint32_t n = a.n;
  if (MODE == MAP_TAKE) {
    n = *reinterpret_cast<volatile const uint32_t*>(a.queue + 3);   // the parked reads without a record
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.counters && n) atomicAdd(a.counters + 3, (unsigned long long)n);
  }
  uint32_t* const queue = MODE == MAP_TAKE ? a.queue + 2 : a.queue;
  for (uint32_t round = TICKET_ROUNDS, base = 0;; ++round) {
    if (round == TICKET_ROUNDS) { base = next_ticket<WD>(queue); round = 0; }
    const uint32_t first = base + round * (32u / WD);
    if (first >= n) break;
    uint32_t r = first + (threadIdx.x & 31u) / WD;
    if (r < n) {
      if (MODE == MAP_TAKE) r = a.park.list[r];
      uint32_t len;
      const char* seq = read_at<PACKED>(a, r, len);
      uint32_t n_log = 0;
      MapStatus ms;
      if (MODE == MAP_PARK)
        ms = map_read_pe_logged<HwGroup<WD>, PACKED, DevPark>(w, a.ix, a.cv, a.p3, a.cfg, seq, len, a.ag != 0u, a.max_mismatches,
                                                               a.top_k, sc, cached_len, a.log + (size_t)r * a.log_slots, hist, n_log,
                                                               ctr, DevPark{a.park, a.queue, r});
      else
        ms = map_read_pe_logged<HwGroup<WD>, PACKED>(w, a.ix, a.cv, a.p3, a.cfg, seq, len, a.ag != 0u, a.max_mismatches, a.top_k,
                                                      sc, cached_len, a.log + (size_t)r * a.log_slots, hist, n_log, ctr);
      bad |= ms == MAP_BAD;
      if (lane == 0 && ms < MAP_PARKED) a.n_log[r] = n_log;
    }
    __syncwarp();
  }
  flush_counters(w, ctr, bad, tally, a.flags, a.counters);
}

// ------------------------------------------------------------------------------------------
// parked reads: flat verification and fold
// ------------------------------------------------------------------------------------------
// A warp per verification block, blocks handed out by tickets: no order, no state, one gather per
// candidate -- the part of the path that runs at memory speed.  `bench.py --workload verify` times it.
constexpr uint32_t VERIFY_TICKET = 16;   // consecutive blocks per ticket (usually of one run: the read is set up once)
constexpr uint32_t VERIFY_WORDS = 12;    // 32-bit words (16 bases each) of a read of up to WIDE_MAX_READ bases
constexpr uint32_t VERIFY_ROW = 20;      // words per candidate in a window tile: 16 + 4 of padding (conflict-free 128-bit rows)
constexpr uint32_t VERIFY_WARP_WORDS = 3u * VERIFY_WORDS + 2u * 32u * VERIFY_ROW;
// per-warp shared memory of verify_kernel: the read scratch, the read / verification mask / seed mask of the
// current lookup as 32-bit words in base order, and two window tiles (one being filled, one being compared)
static size_t verify_smem_bytes(uint32_t nw_max) {
  return (size_t)WARPS_PER_BLOCK * (scratch_words(nw_max) * 8u + VERIFY_WARP_WORDS * 4u);
}

template <int CA>
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gmem) {
  if (CA) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gmem));
  else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// The windows are FETCHED four lanes per candidate -- eight whole 64-byte blocks per instruction, each from
// one cache line (a lane-per-candidate load touches 32 lines per instruction and the L1 serves about one line
// per two cycles) -- by asynchronous copies straight into a 32 x 16-word tile in shared memory, and COMPARED
// one lane per candidate from that tile (alignment selects, funnel shifts and mask loads once per candidate
// instead of once per 16 bases).  Two tiles per warp: the windows of block i + 1 are on their way while block i
// is compared, the entries of block i + 2 while those are issued.
template <bool PACKED, class Args, int CA = 1>
__global__ void __launch_bounds__(BLOCK_THREADS, MIN_BLOCKS_PER_SM)
verify_kernel(const __grid_constant__ Args a) {
  extern __shared__ uint64_t smem[];
  __shared__ BlockTally tally;
  tally_init(tally);
  HwGroup<32> w;
  const uint32_t lane = w.lane(), q = lane & 3u, c4 = lane & ~3u;
  const uint32_t warp = threadIdx.x / 32u;
  ReadScratch sc = carve_scratch(smem + (size_t)warp * scratch_words(a.nw_max), a.nw_max);
  uint32_t* const masks = reinterpret_cast<uint32_t*>(smem + (size_t)WARPS_PER_BLOCK * scratch_words(a.nw_max)) +
                          (size_t)warp * VERIFY_WARP_WORDS;
  uint32_t* const tiles = masks + 3u * VERIFY_WORDS;
  const uint32_t tiles_sa = (uint32_t)__cvta_generic_to_shared(tiles);
  uint32_t cached_len = 0, slots = 0;
  Counters ctr{0u, 0u, 0u};
  uint32_t n_blocks = *reinterpret_cast<volatile const uint32_t*>(a.queue + 5);
  if (n_blocks > a.park.cap_blocks) n_blocks = a.park.cap_blocks;
  for (;;) {
    uint32_t base = 0;
    if (lane == 0u) base = atomicAdd(a.queue + 6, VERIFY_TICKET);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (base >= n_blocks) break;
    // lane i describes block base + i: record, lookup, first slot, end of the run (dead: nothing to do)
    uint32_t d_t = 0xFFFFFFFFu, d_j = 0u, d_r = 0u, d_first = 0u, d_end = 0u, d_len = 0u;
    if (lane < VERIFY_TICKET && base + lane < n_blocks) {
      d_t = a.park.desc[2u * (size_t)(base + lane)];
      if (d_t != 0xFFFFFFFFu) {                           // else: reserved by a read that found no room
        const uint32_t word = a.park.desc[2u * (size_t)(base + lane) + 1u];
        const uint32_t* rec = a.park.recs + (size_t)d_t * PARK_WORDS;
        const uint32_t* mine = rec + 2u + PARK_LOOKUP_WORDS * (word & 7u);
        d_j = word & 7u; d_r = rec[0]; d_first = mine[1] + 32u * (word >> 3); d_end = mine[1] + mine[2];
        read_at<PACKED>(a, d_r, d_len);
      }
    }
    // this lane's entry of block i (0 for the slots past the end of a run and for dead blocks)
    auto entry_of = [&](uint32_t i) -> uint32_t {
      if (i >= VERIFY_TICKET) return 0u;
      const uint32_t first = __shfl_sync(0xFFFFFFFFu, d_first, (int)i), end = __shfl_sync(0xFFFFFFFFu, d_end, (int)i);
      const uint32_t j = __shfl_sync(0xFFFFFFFFu, d_j, (int)i);
      return first + lane < end ? a.ix[j / 3u].entries[first + lane].pos : 0u;
    };
    // the windows of block i, straight into tile i & 1: step u fetches the 64-byte block of slot 4c + u, 16 bytes per lane
    auto fetch = [&](uint32_t i, uint32_t e) {
      if (i < VERIFY_TICKET) {
        const uint32_t j = __shfl_sync(0xFFFFFFFFu, d_j, (int)i), rl = __shfl_sync(0xFFFFFFFFu, d_len, (int)i);
        const uint64_t* genome = a.ix[j / 3u].genome;
        const uint32_t sa = tiles_sa + ((i & 1u) * 32u * VERIFY_ROW + 4u * q) * 4u;
#pragma unroll
        for (uint32_t u = 0; u < 4u; ++u) {
          const uint64_t gp = (uint64_t)__shfl_sync(0xFFFFFFFFu, e, (int)(c4 + u)) + PAD_BASES - j % 3u;
          // the fourth 16 bytes only if the window reaches them: every other block starts in the middle of a
          // 32-byte sector, and its last quarter is a third sector
          if (q < 3u || ((uint32_t)gp & 63u) + rl > 192u)
            cp_async16<CA>(sa + (c4 + u) * VERIFY_ROW * 4u, reinterpret_cast<const uint4*>(genome) + ((gp >> 6) + q));
        }
      }
      cp_async_commit();
    };
    uint32_t cur_t = 0xFFFFFFFFu, cur_j = 0xFFFFFFFFu, len = 0u;
    uint32_t e_cur = entry_of(0), e_next = entry_of(1);
    fetch(0, e_cur);
    for (uint32_t i = 0; i < VERIFY_TICKET; ++i) {
      fetch(i + 1u, e_next);                               // an empty group past the ticket's end
      const uint32_t e_after = entry_of(i + 2u);
      const uint32_t t = __shfl_sync(0xFFFFFFFFu, d_t, (int)i);
      cp_async_wait<1>();                                 // block i's windows have landed (block i + 1's may still fly)
      __syncwarp();
      if (t != 0xFFFFFFFFu) {
        const uint32_t j = __shfl_sync(0xFFFFFFFFu, d_j, (int)i), first = __shfl_sync(0xFFFFFFFFu, d_first, (int)i),
                       end = __shfl_sync(0xFFFFFFFFu, d_end, (int)i);
        const uint32_t seed_i = j % 3u;
        const bool live = first + lane < end;
        if (t != cur_t) {                                 // another read: pack it
          const char* seq = read_at<PACKED>(a, __shfl_sync(0xFFFFFFFFu, d_r, (int)i), len);
          __syncwarp();
          load_read<HwGroup<32>, PACKED>(w, seq, len, a.ag != 0u, sc);
          if (cached_len != len) { build_masks(w, len, sc); cached_len = len; }
        }
        if (t != cur_t || j != cur_j) {                   // another lookup: read and masks as 32-bit words in base order
          const uint32_t nw = (len + 31u) >> 5;
          __syncwarp();
          if (lane < VERIFY_WORDS) {
            const uint32_t k = lane >> 1, sh = (lane & 1u) ? 0u : 32u;
            masks[lane] = k < nw ? (uint32_t)(sc.R[k] >> sh) : 0u;
            masks[VERIFY_WORDS + lane] = k < nw ? (uint32_t)(sc.VM[seed_i * sc.nw + k] >> sh) : 0u;
            masks[2u * VERIFY_WORDS + lane] = k < nw ? (uint32_t)(sc.SM[seed_i * sc.nw + k] >> sh) : 0u;
          }
          cur_t = t; cur_j = j;
          __syncwarp();
        }
        const uint32_t* tile = tiles + (i & 1u) * 32u * VERIFY_ROW;
        uint32_t W[16];
#pragma unroll
        for (uint32_t k = 0; k < 4u; ++k) {               // memory order of the packed words: u64 pairs, low half first
          const uint4 v = *reinterpret_cast<const uint4*>(tile + lane * VERIFY_ROW + 4u * k);
          W[4u * k] = v.y; W[4u * k + 1u] = v.x; W[4u * k + 2u] = v.w; W[4u * k + 3u] = v.z;
        }
        // align the window with read position 0: 2 s bits = (s >> 4) words + (2 s & 31) bits
        const uint32_t s = (uint32_t)(((uint64_t)e_cur + PAD_BASES - seed_i) & 63u);
        const uint32_t off = s >> 4, bit = (2u * s) & 31u;
        uint32_t tt[15], v[13];
#pragma unroll
        for (uint32_t k = 0; k < 15u; ++k) tt[k] = (off & 2u) ? (k + 2u < 16u ? W[k + 2u] : 0u) : W[k];
#pragma unroll
        for (uint32_t k = 0; k < 13u; ++k) v[k] = (off & 1u) ? tt[k + 1u] : tt[k];
        uint32_t mm = 0u, sd = 0u;
#pragma unroll
        for (uint32_t k4 = 0; k4 < VERIFY_WORDS; k4 += 4u) {
          const uint4 r4 = *reinterpret_cast<const uint4*>(masks + k4);
          const uint4 vm4 = *reinterpret_cast<const uint4*>(masks + VERIFY_WORDS + k4);
          const uint4 sm4 = *reinterpret_cast<const uint4*>(masks + 2u * VERIFY_WORDS + k4);
          const uint32_t rr[4] = {r4.x, r4.y, r4.z, r4.w}, vv[4] = {vm4.x, vm4.y, vm4.z, vm4.w}, ss[4] = {sm4.x, sm4.y, sm4.z, sm4.w};
#pragma unroll
          for (uint32_t k = 0; k < 4u; ++k) {
            const uint32_t x = __funnelshift_l(v[k4 + k + 1u], v[k4 + k], bit) ^ rr[k];
            const uint32_t d = x | (x >> 1);
            mm += (uint32_t)__popc(d & vv[k]);
            sd |= d & ss[k];
          }
        }
        const bool cand = live && sd == 0u;
        if (cand) ctr.candidates++;
        if (live) ++slots;
        a.park.bytes[32u * (size_t)(base + i) + lane] = (uint8_t)(cand ? mm : MM_NONE);
      }
      __syncwarp();                                       // tile i & 1 is refilled two blocks on
      e_cur = e_next; e_next = e_after;
    }
    cp_async_wait<0>();
  }
  slots = __reduce_add_sync(0xFFFFFFFFu, slots);
  if (lane == 0u && slots && a.counters) atomicAdd(a.counters + 4, (unsigned long long)slots);
  flush_counters(w, ctr, false, tally, a.flags, a.counters);
}

// the bytes of all runs of a parked read, on their way before the ordered pass asks for them
__device__ __forceinline__ void prefetch_run_bytes(const uint32_t* rec, const uint8_t* bytes, uint32_t lane) {
  for (uint32_t j = 0; j < LOOKUP_LANES; ++j) {
    const uint32_t* mine = rec + 2u + PARK_LOOKUP_WORDS * j;
    if ((mine[0] & 0xFFu) != PK_RUN) continue;
    const uint8_t* p = bytes + 32u * (size_t)mine[3];
    for (uint32_t off = 128u * lane; off < mine[2]; off += 128u * 32u) WALT_PREFETCH(p + off);
  }
}

// A warp per parked read with a record: the ordered fold (fold_parked, walt_core.cuh).
template <bool PACKED>
__global__ void __launch_bounds__(BLOCK_THREADS, MIN_BLOCKS_PER_SM)
se_fold_kernel(const __grid_constant__ SeArgs a) {
  __shared__ BlockTally tally;
  tally_init(tally);
  HwGroup<32> w;
  const uint32_t lane = w.lane();
  Counters ctr{0u, 0u, 0u};
  uint32_t n = *reinterpret_cast<volatile const uint32_t*>(a.queue + 1);
  if (n > a.park.rec_cap) n = a.park.rec_cap;
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.counters && n) atomicAdd(a.counters + 3, (unsigned long long)n);
  for (;;) {
    const uint32_t t = next_ticket<32>(a.queue + 7);
    if (t >= n) break;
    const uint32_t* rec = a.park.recs + (size_t)t * PARK_WORDS;
    const uint32_t r = rec[0];
    if (r == 0xFFFFFFFFu) continue;                       // a dead record: the read is on MAP_TAKE's list
    uint32_t len;
    read_at<PACKED>(a, r, len);
    prefetch_run_bytes(rec, a.park.bytes, lane);
    BestSink<HwGroup<32>> sink;
    sink.st.pos = 0u; sink.st.times = 0u; sink.st.mm = a.max_mismatches; sink.st.strand = '+';
    fold_parked(w, a.ix, a.cv, len, rec, a.park.bytes, sink,
                [&](const uint32_t* m, uint32_t strand) { sink.apply(m[1], m[2], m[3], m[4], strand); });
    if (lane == 0) {
      uint4 o;
      o.x = sink.st.pos; o.y = sink.st.times; o.z = sink.st.mm; o.w = sink.st.strand & 0xFFu;
      *reinterpret_cast<uint4*>(a.out + r) = o;
    }
    __syncwarp();
  }
  flush_counters(w, ctr, false, tally, a.flags, a.counters);
}

template <bool PACKED>
__global__ void __launch_bounds__(BLOCK_THREADS, MIN_BLOCKS_PER_SM)
pe_fold_kernel(const __grid_constant__ PeArgs a) {
  __shared__ BlockTally tally;
  __shared__ uint32_t hists[WARPS_PER_BLOCK][LOG_MAX_MM + 1u];
  tally_init(tally);
  HwGroup<32> w;
  const uint32_t lane = w.lane();
  uint32_t* hist = hists[threadIdx.x / 32u];
  Counters ctr{0u, 0u, 0u};
  uint32_t n = *reinterpret_cast<volatile const uint32_t*>(a.queue + 1);
  if (n > a.park.rec_cap) n = a.park.rec_cap;
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.counters && n) atomicAdd(a.counters + 3, (unsigned long long)n);
  for (;;) {
    const uint32_t t = next_ticket<32>(a.queue + 7);
    if (t >= n) break;
    const uint32_t* rec = a.park.recs + (size_t)t * PARK_WORDS;
    const uint32_t r = rec[0];
    if (r == 0xFFFFFFFFu) continue;
    uint32_t len;
    read_at<PACKED>(a, r, len);
    prefetch_run_bytes(rec, a.park.bytes, lane);
    LogSink<HwGroup<32>> sink;
    sink.log = a.log + (size_t)r * a.log_slots; sink.hist = hist; sink.cap = a.top_k; sink.max_mm = a.max_mismatches;
    sink.reset(w);
    fold_parked(w, a.ix, a.cv, len, rec, a.park.bytes, sink,
                [&](const uint32_t* m, uint32_t strand) {
                  const uint32_t k = m[0] >> 8;
                  uint32_t g = 0u, mm = 0u;
                  if (lane < k) { g = m[1u + 2u * lane]; mm = m[2u + 2u * lane]; }
                  sink.consume(w, lane < k, mm, g, strand);
                });
    if (lane == 0) a.n_log[r] = sink.n_log;
    __syncwarp();
  }
  flush_counters(w, ctr, false, tally, a.flags, a.counters);
}

// Second phase: one THREAD per read replays its log through TopCandidates' heap (paired.hpp:51-74,
// libstdc++ mechanics) and drains it (paired.cpp:684-692).  Thread t < n serves mate 1's read t,
// thread n + t mate 2's.  LOCAL_CAP > 0: the heap is a local array (top_k <= LOCAL_CAP), else it
// lives in `heaps` (top_k entries per thread).
struct HeapArgs {
  const HeapEntry* log[2];
  const uint32_t* n_log[2];
  walt_cand* ranked[2];
  uint32_t* n_ranked[2];
  HeapEntry* heaps;
  const uint32_t* list[2];       // parked reads of the two mate launches: MAP_TAKE's lists (NULL: nothing was parked) ...
  const uint32_t* recs[2];       // ... and the records of the others
  const uint32_t* queue[2];      // their queue blocks: [1] records, [3] reads on the list
  uint32_t rec_cap;
  uint32_t n, top_k, log_slots;
  uint32_t zero_fill;     // see PeArgs
};

// Work items: [0, 2n) every read of both mates whose log is short (reads that were not parked
// log at most LOOKUP_LANES * LANE_RUN_CAP events); then, per mate, n places for the records and n for
// MAP_TAKE's list: the parked reads, i.e. the long logs, next to each other -- so that the lanes of a
// warp replay logs of similar length.
constexpr uint32_t HEAP_BLOCK = 128;
constexpr uint32_t SHORT_LOG = LOOKUP_LANES * LANE_RUN_CAP;
// one thread's heap inside a block-wide array: slot i of thread t at [i * HEAP_BLOCK + t] (8-byte
// entries: a warp's accesses never conflict, whatever slots its lanes are at)
struct StridedHeap {
  HeapEntry* base;
  __device__ __forceinline__ HeapEntry& operator[](uint32_t i) const { return base[i * HEAP_BLOCK]; }
};
template <class H>
__device__ __forceinline__ void sift_up_g(H a, uint32_t hole, uint32_t top, HeapEntry v) {
  while (hole > top) {
    const uint32_t parent = (hole - 1u) >> 1;
    const HeapEntry pe = a[parent];
    if (!(he_mm(pe) < he_mm(v))) break;
    a[hole] = pe;
    hole = parent;
  }
  a[hole] = v;
}
template <class H>
__device__ __forceinline__ void adjust_g(H a, uint32_t hole, uint32_t len, HeapEntry v) {   // stl_heap.h:224-249
  const uint32_t top = hole;
  uint32_t child = hole;
  while (len >= 2u && child < (len - 1u) / 2u) {
    child = 2u * (child + 1u);
    HeapEntry c = a[child];
    const HeapEntry d = a[child - 1u];
    if (he_mm(c) < he_mm(d)) { --child; c = d; }
    a[hole] = c;
    hole = child;
  }
  if ((len & 1u) == 0u && len >= 2u && child == (len - 2u) / 2u) {
    child = 2u * (child + 1u);
    a[hole] = a[child - 1u];
    hole = child - 1u;
  }
  sift_up_g(a, hole, top, v);
}
template <class H>
__device__ __forceinline__ void push_bounded_g(H a, uint32_t& size, uint32_t cap, HeapEntry v) {   // paired.hpp:60-67
  if (size < cap) {
    ++size;
    sift_up_g(a, size - 1u, 0u, v);
  } else if (he_mm(v) < he_mm(a[0])) {
    const uint32_t n = size;
    if (n > 1u) { const HeapEntry last = a[n - 1u]; a[n - 1u] = a[0]; adjust_g(a, 0u, n - 1u, last); }
    sift_up_g(a, n - 1u, 0u, v);
  }
}
template <class H>
__device__ __forceinline__ HeapEntry pop_g(H a, uint32_t& size) {
  const HeapEntry top = a[0];
  const uint32_t n = size;
  if (n > 1u) { const HeapEntry last = a[n - 1u]; a[n - 1u] = a[0]; adjust_g(a, 0u, n - 1u, last); }
  size = n - 1u;
  return top;
}

template <bool SMEM>
__global__ void __launch_bounds__(HEAP_BLOCK)
pe_heap_kernel(const __grid_constant__ HeapArgs a) {
  extern __shared__ HeapEntry heap_sm[];   // SMEM: top_k * HEAP_BLOCK entries
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t mate, r;
  if (t < 2u * a.n) {
    mate = t >= a.n ? 1u : 0u; r = t - mate * a.n;
    if (a.list[0] && a.n_log[mate][r] > SHORT_LOG) return;        // a parked read: below
  } else {
    if (!a.list[0] || t >= 6u * a.n) return;
    uint32_t u = t - 2u * a.n;
    mate = u >= 2u * a.n ? 1u : 0u; u -= mate * 2u * a.n;
    if (u < a.n) {
      if (u >= a.queue[mate][1] || u >= a.rec_cap) return;
      r = a.recs[mate][(size_t)u * PARK_WORDS];
      if (r == 0xFFFFFFFFu) return;
    } else {
      u -= a.n;
      if (u >= a.queue[mate][3]) return;
      r = a.list[mate][u];
    }
    if (a.n_log[mate][r] <= SHORT_LOG) return;                    // done above
  }
  const HeapEntry* log = a.log[mate] + (size_t)r * a.log_slots;
  const uint32_t n_log = a.n_log[mate][r];
  walt_cand* dst = a.ranked[mate] + (size_t)r * a.top_k;
  uint32_t size = 0, c = 0;
  auto run = [&](auto heap) {
    for (uint32_t i = 0; i < n_log; ++i) push_bounded_g(heap, size, a.top_k, log[i]);
    while (size) {
      const HeapEntry e = pop_g(heap, size);
      walt_cand o;
      o.genome_pos = e.pos; o.mismatch = he_mm(e); o.strand = (e.mm_strand & 0x10000u) ? '-' : '+';
      o.pad[0] = o.pad[1] = o.pad[2] = 0;
      dst[c++] = o;
    }
  };
  if (SMEM) run(StridedHeap{heap_sm + threadIdx.x});
  else run(a.heaps + ((size_t)mate * a.n + r) * a.top_k);
  a.n_ranked[mate][r] = c;
  if (a.zero_fill) {   // whole-array compares and copies on the host side are deterministic
    uint32_t* z = reinterpret_cast<uint32_t*>(dst + c);
    for (uint32_t i = 0; i < (a.top_k - c) * 3u; ++i) z[i] = 0u;
  }
}

// Single-kernel form (max_mismatches > LOG_MAX_MM, or WALT_PE_LOGGED=0): PairEndMapping
// (paired.cpp:106-201) for one mate batch + the heap drain (paired.cpp:684-692), heap kept by lane 0
template <uint32_t WD, bool PACKED>
__global__ void __launch_bounds__(BLOCK_THREADS, MIN_BLOCKS_PER_SM)
pe_map_kernel(const __grid_constant__ PeArgs a) {
  extern __shared__ uint64_t smem[];
  __shared__ BlockTally tally;
  tally_init(tally);
  HwGroup<WD> w;
  const uint32_t lane = w.lane();
  const uint32_t group_in_block = threadIdx.x / WD;
  const uint32_t per_group = scratch_words(a.nw_max) + a.top_k + 1u;  // HeapEntry is 8 bytes
  uint64_t* mine = smem + (size_t)group_in_block * per_group;
  ReadScratch sc = carve_scratch(mine, a.nw_max);
  HeapEntry* heap = reinterpret_cast<HeapEntry*>(mine + scratch_words(a.nw_max));
  uint32_t cached_len = 0;
  Counters ctr{0u, 0u, 0u};
  bool bad = false;
  for (uint32_t round = TICKET_ROUNDS, base = 0;; ++round) {
    if (round == TICKET_ROUNDS) { base = next_ticket<WD>(a.queue); round = 0; }
    const uint32_t first = base + round * (32u / WD);
    if (first >= a.n) break;
    const uint32_t r = first + (threadIdx.x & 31u) / WD;
    if (r < a.n) {
      uint32_t len;
      const char* seq = read_at<PACKED>(a, r, len);
      uint32_t hsize = 0;
      const MapStatus ms = map_read_pe<HwGroup<WD>, PACKED>(w, a.ix, a.cv, a.p3, a.cfg, seq, len, a.ag != 0u,
                                                          a.max_mismatches, a.top_k, sc, cached_len, heap, hsize, ctr);
      bad |= ms == MAP_BAD;
      if (lane == 0) {
        walt_cand* dst = a.ranked + (size_t)r * a.top_k;
        uint32_t c = 0, sz = hsize;
        while (sz) {
          HeapEntry t = heap_pop(heap, sz);
          walt_cand o;
          o.genome_pos = t.pos; o.mismatch = he_mm(t); o.strand = (t.mm_strand & 0x10000u) ? '-' : '+';
          o.pad[0] = o.pad[1] = o.pad[2] = 0;
          dst[c++] = o;
        }
        a.n_ranked[r] = c;
      }
      w.sync();
      if (a.zero_fill) {  // unused slots are defined (zero) so whole-array compares and copies are deterministic
        const uint32_t used = w.shfl(hsize, 0);
        uint32_t* z = reinterpret_cast<uint32_t*>(a.ranked + (size_t)r * a.top_k + used);
        for (uint32_t i = lane; i < (a.top_k - used) * 3u; i += WD) z[i] = 0u;
      }
    }
    __syncwarp();
  }
  flush_counters(w, ctr, bad, tally, a.flags, a.counters);
}

struct GetRanked {
  const walt_cand* p;
  __device__ __forceinline__ RankedCand operator()(uint32_t i) const {
    RankedCand c; c.pos = p[i].genome_pos; c.mm = p[i].mismatch; c.strand = (uint8_t)p[i].strand;
    return c;
  }
};

// pairing loop of MergePairedEndResults (paired.cpp:472-513) and, for the compact result, the
// per-mate GetBestMatch4Single (paired.cpp:296-318): one thread per pair.  `swap` hands every
// per-mate field back to the other mate (PBAT runs the directional protocol on exchanged mates).
struct PairArgs {
  ChromView cv;
  const walt_cand* r1; const uint32_t* n1; const uint64_t* offs1; uint32_t ulen1;
  const walt_cand* r2; const uint32_t* n2; const uint64_t* offs2; uint32_t ulen2;
  uint32_t n, top_k, max_mismatches;
  int32_t frag_range;
  uint32_t swap;
  uint32_t wide;             // long lists are paired by the whole warp (needs max_mismatches <= 127)
  walt_pair* pairs;          // may be NULL
  walt_pe_result* compact;   // may be NULL
};

__device__ __forceinline__ walt_best to_best(const BestState& b) {
  walt_best o; o.genome_pos = b.pos; o.times = b.times; o.mismatch = b.mm; o.strand = (char)b.strand;
  o.pad[0] = o.pad[1] = o.pad[2] = 0;
  return o;
}

// Pairs with short lists are paired by their own thread; the pairs of a warp whose k1 x k2 loop is
// long (both mates in repeats) are then taken one after the other by the whole warp
// (pair_candidates_wide): a lone lane in a 2500-iteration loop costs the warp as much as 32 of them.
constexpr uint32_t PAIR_BLOCK = 128;
__global__ void __launch_bounds__(PAIR_BLOCK)
pair_kernel(const __grid_constant__ PairArgs a) {
  extern __shared__ uint32_t pair_sm[];   // 3 * top_k words per warp (wide path only)
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = p < a.n;
  uint32_t len1 = 0, len2 = 0, k1 = 0, k2 = 0;
  const walt_cand* c1 = a.r1 + (size_t)(live ? p : 0u) * a.top_k;
  const walt_cand* c2 = a.r2 + (size_t)(live ? p : 0u) * a.top_k;
  if (live) {
    len1 = a.ulen1 ? a.ulen1 : (uint32_t)(a.offs1[p + 1] - a.offs1[p]);
    len2 = a.ulen2 ? a.ulen2 : (uint32_t)(a.offs2[p + 1] - a.offs2[p]);
    k1 = a.n1[p]; k2 = a.n2[p];
  }
  const bool wide = a.wide && k1 * k2 > PAIR_WIDE_MIN;
  PairResult r; r.best_times = 0u; r.best_i = -1; r.best_j = -1; r.frag = 0;
  if (live && !wide)
    r = pair_candidates(a.cv, GetRanked{c1}, k1, len1, GetRanked{c2}, k2, len2, a.max_mismatches, a.frag_range);
  uint32_t todo = __ballot_sync(0xFFFFFFFFu, wide);
  if (todo) {
    HwGroup<32> w;
    uint32_t* sm = pair_sm + (threadIdx.x / 32u) * 3u * a.top_k;
    const uint32_t l = threadIdx.x & 31u;
    while (todo) {
      const int src = __ffs((int)todo) - 1;
      todo &= todo - 1u;
      const uint32_t q = __shfl_sync(0xFFFFFFFFu, p, src);
      const PairResult t = pair_candidates_wide(w, a.cv, GetRanked{a.r1 + (size_t)q * a.top_k}, __shfl_sync(0xFFFFFFFFu, k1, src),
                                                __shfl_sync(0xFFFFFFFFu, len1, src), GetRanked{a.r2 + (size_t)q * a.top_k},
                                                __shfl_sync(0xFFFFFFFFu, k2, src), __shfl_sync(0xFFFFFFFFu, len2, src),
                                                a.max_mismatches, a.frag_range, sm);
      if ((int)l == src) r = t;
    }
  }
  if (!live) return;
  walt_pair o; o.best_times = r.best_times; o.frag_len = r.frag;
  o.best_i = a.swap ? r.best_j : r.best_i; o.best_j = a.swap ? r.best_i : r.best_j;
  if (a.pairs) a.pairs[p] = o;
  if (a.compact) {
    walt_pe_result out;
    out.pair = o;
    walt_cand z; z.genome_pos = 0; z.mismatch = 0; z.strand = 0; z.pad[0] = z.pad[1] = z.pad[2] = 0;
    walt_cand w1 = z, w2 = z;
    if (r.best_times >= 1u) { w1 = c1[r.best_i]; w2 = c2[r.best_j]; }
    const walt_best s1 = to_best(single_best(GetRanked{c1}, k1, a.max_mismatches));
    const walt_best s2 = to_best(single_best(GetRanked{c2}, k2, a.max_mismatches));
    out.c1 = a.swap ? w2 : w1; out.c2 = a.swap ? w1 : w2;
    out.single1 = a.swap ? s2 : s1; out.single2 = a.swap ? s1 : s2;
    a.compact[p] = out;
  }
}

// ------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------
static size_t se_smem_bytes(uint32_t nw_max, uint32_t wd) { return (size_t)(BLOCK_THREADS / wd) * scratch_words(nw_max) * 8u; }
static size_t pe_smem_bytes(uint32_t nw_max, uint32_t top_k, uint32_t wd) {
  return (size_t)(BLOCK_THREADS / wd) * (scratch_words(nw_max) + top_k + 1u) * 8u;
}

static size_t pe_log_smem_bytes(uint32_t nw_max, uint32_t wd) {
  return (size_t)(BLOCK_THREADS / wd) * (scratch_words(nw_max) + (LOG_MAX_MM + 2u) / 2u) * 8u;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is state of the function on its device, shared by every engine
// (and host thread) of the process that runs on that device: it is only ever raised, under a lock, so a
// launch of one engine never finds the limit lowered by another engine's smaller batch.
static int allow_dynamic_smem(const void* kernel, size_t smem) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, size_t> allowed;
  int dev = 0;
  WALT_CUDA_TRY(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  size_t& have = allowed[{dev, kernel}];
  if (smem <= have) return WALT_OK;
  WALT_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  have = smem;
  return WALT_OK;
}

template <class K>
static int grid_for(walt_engine* e, K kernel, size_t smem, uint32_t n, uint32_t wd, uint32_t* grid, uint32_t share = 1) {
  int per_sm = 0;
  if (int rc = allow_dynamic_smem((const void*)kernel, smem)) return rc;
  WALT_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, (int)BLOCK_THREADS, smem));
  if (per_sm < 1) return fail(WALT_ECUDA, "mapping kernel does not fit on an SM");
  // persistent: a multiple of the SM count.  share > 1: this launch is one of several chunk kernels in
  // flight (host batches) and takes per_sm / share CTA slots per SM, so that the next chunk's kernel
  // runs beside it and each launch's ramp-up and tail overlap the other's steady state.
  per_sm = std::max<int>(1, per_sm / (int)std::max<uint32_t>(1u, share));
  uint64_t g = (uint64_t)per_sm * (uint64_t)e->sm_count;
  const uint32_t per_block = BLOCK_THREADS / wd;
  const uint64_t need = ((uint64_t)n + per_block - 1) / per_block;
  if (need < g) g = need ? need : 1;
  *grid = (uint32_t)g;
  return WALT_OK;
}

static int check_pair(walt_engine* e, int ag) {
  const int a = ag ? WALT_GA10 : WALT_CT00;
  if (!e->sub[a].loaded || !e->sub[a + 1].loaded)
    return fail(WALT_ENOTLOADED, ag ? "G->A sub-indexes (_GA10/_GA11) are not resident"
                                    : "C->T sub-indexes (_CT00/_CT01) are not resident");
  return WALT_OK;
}

// one batch (or chunk) of reads resident on the device
struct ReadSrc {
  const char* d_seqs;       // ASCII, or 2-bit packed (walt_pack_reads layout)
  const uint64_t* d_offs;   // n + 1 base offsets; may be NULL if uniform_len
  uint64_t seq_base;        // offs value of the first read of d_seqs
  uint32_t uniform_len;     // != 0: all reads have this length (stride addressing)
  uint32_t read_base;       // packed: index inside the packed buffer of the launch's first read
  uint32_t max_len;
  bool packed;
};

// queue block i of the engine: QUEUE_WORDS words (see SeArgs::queue)
// 0: device-resident SE; 1..N_SLOTS: SE host chunks; then two (one per mate) for device-resident PE
// and two per slot for PE host chunks
constexpr uint32_t N_COUNTERS = 5;   // lookups, candidates, literal lookups, parked reads, slots verified by verify_kernel
constexpr uint32_t QUEUE_WORDS = 8;
static uint32_t* queue_block(walt_engine* e, uint32_t i) { return e->d_flags + 16u + QUEUE_WORDS * i; }
constexpr uint32_t QB_SE_DEVICE = 0, QB_SE_SLOT = 1, QB_PE_DEVICE = 1 + N_SLOTS, QB_PE_SLOT = 3 + N_SLOTS;
constexpr uint32_t N_FLAG_WORDS = 16u + 8u * (QB_PE_SLOT + 2u * N_SLOTS);

template <class Args>
static void fill_common(walt_engine* e, Args& a, const ReadSrc& src, uint32_t n, int ag, uint32_t m, uint32_t b,
                        uint32_t* d_queue) {
  const int base = ag ? WALT_GA10 : WALT_CT00;
  a.ix[0] = e->sub[base].view(base); a.ix[1] = e->sub[base + 1].view(base + 1);
  a.cv = chrom_view(e); a.p3 = e->pow3;
  a.cfg.b = b; a.cfg.literal_all = e->search_mode == 1 ? 1u : 0u;
  a.seqs = src.d_seqs; a.offs = src.d_offs; a.seq_base = src.seq_base; a.n = n; a.uniform_len = src.uniform_len;
  a.read_base = src.read_base;
  a.nw_max = std::max<uint32_t>(1u, (src.max_len + 31u) / 32u);
  a.ag = ag ? 1u : 0u; a.max_mismatches = m; a.flags = e->d_flags; a.queue = d_queue; a.park = ParkView{nullptr, nullptr, nullptr, nullptr, 0u, 0u};
  a.counters = e->d_counters;
}

// grid of a kernel that takes over parked reads: persistent, a whole warp per read
template <class K>
static int take_grid(walt_engine* e, K kernel, size_t smem, uint32_t* grid) {
  int per_sm = 0;
  if (int rc = allow_dynamic_smem((const void*)kernel, smem)) return rc;
  WALT_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, (int)BLOCK_THREADS, smem));
  if (per_sm < 1) return fail(WALT_ECUDA, "take-over kernel does not fit on an SM");
  *grid = (uint32_t)per_sm * (uint32_t)e->sm_count;
  return WALT_OK;
}

// The kernels behind one that parks reads: verification of every reserved block, the fold of the reads
// with a record, MAP_TAKE for the rest.  All persistent; each finds out on the device how much there is to do.
template <class Args>
static int launch_behind_park(walt_engine* e, const Args& a, void (*verify)(Args), void (*fold)(Args), void (*take)(Args),
                              size_t smem_v, size_t smem_f, size_t smem_t, cudaStream_t st) {
  int rc;
  uint32_t grid = 0;
  if (a.park.rec_cap) {
    if ((rc = take_grid(e, verify, smem_v, &grid))) return rc;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (e->time_kernels) {
      WALT_CUDA_TRY(cudaEventCreate(&ev0)); WALT_CUDA_TRY(cudaEventCreate(&ev1));
      WALT_CUDA_TRY(cudaEventRecord(ev0, st));
    }
    verify<<<grid, BLOCK_THREADS, smem_v, st>>>(a);
    WALT_CUDA_TRY(cudaGetLastError());
    if (e->time_kernels) {
      WALT_CUDA_TRY(cudaEventRecord(ev1, st));
      e->verify_events.emplace_back(ev0, ev1);
    }
    if ((rc = take_grid(e, fold, smem_f, &grid))) return rc;
    fold<<<grid, BLOCK_THREADS, smem_f, st>>>(a);
    WALT_CUDA_TRY(cudaGetLastError());
    e->stats.n_kernel_launches += 2;
  }
  if ((rc = take_grid(e, take, smem_t, &grid))) return rc;
  take<<<grid, BLOCK_THREADS, smem_t, st>>>(a);
  WALT_CUDA_TRY(cudaGetLastError());
  e->stats.n_kernel_launches++;
  return WALT_OK;
}

static ParkView park_view(const ParkBuf* pk) {
  return ParkView{pk->recs, pk->desc, pk->bytes, pk->list, pk->rec_cap, pk->cap_blocks};
}

// pk != NULL (reserve_park): reads that need more than their lookup lanes are parked by the first kernel
// and finished by the kernels behind it
static int launch_se(walt_engine* e, const ReadSrc& src, uint32_t n, int ag, uint32_t m, uint32_t b, walt_best* d_out,
                     uint32_t* d_queue, const ParkBuf* pk, cudaStream_t st, uint32_t share = 1) {
  if (src.max_len > MAX_READ_LEN) return fail(WALT_EINVAL, "read longer than 1024 bases");
  SeArgs a;
  fill_common(e, a, src, n, ag, m, b, d_queue);
  a.out = d_out;
  const bool park = e->defer && pk != nullptr;
  if (park) a.park = park_view(pk);
  const uint32_t wd = e->group_width;
  const size_t smem = se_smem_bytes(a.nw_max, wd);
  uint32_t grid = 0;
  void (*kernel)(SeArgs);
  if (park)
    kernel = src.packed ? (wd == 8u ? se_map_kernel<8, true, MAP_PARK> : wd == 16u ? se_map_kernel<16, true, MAP_PARK> : se_map_kernel<32, true, MAP_PARK>)
                        : (wd == 8u ? se_map_kernel<8, false, MAP_PARK> : wd == 16u ? se_map_kernel<16, false, MAP_PARK> : se_map_kernel<32, false, MAP_PARK>);
  else
    kernel = src.packed ? (wd == 8u ? se_map_kernel<8, true, MAP_ALL> : wd == 16u ? se_map_kernel<16, true, MAP_ALL> : se_map_kernel<32, true, MAP_ALL>)
                        : (wd == 8u ? se_map_kernel<8, false, MAP_ALL> : wd == 16u ? se_map_kernel<16, false, MAP_ALL> : se_map_kernel<32, false, MAP_ALL>);
  int rc = grid_for(e, kernel, smem, n, wd, &grid, share);
  if (rc) return rc;
  WALT_CUDA_TRY(cudaMemsetAsync(d_queue, 0, 4 * QUEUE_WORDS, st));
  kernel<<<grid, BLOCK_THREADS, smem, st>>>(a);
  WALT_CUDA_TRY(cudaGetLastError());
  e->stats.n_kernel_launches++;
  if (park)
    return launch_behind_park(e, a, e->verify_cg ? (src.packed ? verify_kernel<true, SeArgs, 0> : verify_kernel<false, SeArgs, 0>)
                                                 : (src.packed ? verify_kernel<true, SeArgs, 1> : verify_kernel<false, SeArgs, 1>),
                              src.packed ? se_fold_kernel<true> : se_fold_kernel<false>,
                              src.packed ? se_map_kernel<32, true, MAP_TAKE> : se_map_kernel<32, false, MAP_TAKE>,
                              verify_smem_bytes(a.nw_max), 0, se_smem_bytes(a.nw_max, 32u), st);
  return WALT_OK;
}

// first phase of the two-phase form for one mate
static int launch_pe_log(walt_engine* e, const ReadSrc& src, uint32_t n, int ag, uint32_t m, uint32_t b, uint32_t top_k,
                         HeapEntry* d_log, uint32_t* d_nlog, uint32_t* d_queue, const ParkBuf* pk, cudaStream_t st) {
  // (zero_fill belongs to the second phase here)
  if (src.max_len > MAX_READ_LEN) return fail(WALT_EINVAL, "read longer than 1024 bases");
  PeArgs a;
  fill_common(e, a, src, n, ag, m, b, d_queue);
  a.top_k = top_k; a.ranked = nullptr; a.n_ranked = nullptr;
  a.log = d_log; a.n_log = d_nlog; a.log_slots = pe_log_slots(top_k, m); a.zero_fill = 0u;
  const bool park = e->defer && pk != nullptr;
  if (park) a.park = park_view(pk);
  const uint32_t wd = e->group_width;
  const size_t smem = pe_log_smem_bytes(a.nw_max, wd);
  uint32_t grid = 0;
  void (*kernel)(PeArgs);
  if (park)
    kernel = src.packed ? (wd == 8u ? pe_log_kernel<8, true, MAP_PARK> : wd == 16u ? pe_log_kernel<16, true, MAP_PARK> : pe_log_kernel<32, true, MAP_PARK>)
                        : (wd == 8u ? pe_log_kernel<8, false, MAP_PARK> : wd == 16u ? pe_log_kernel<16, false, MAP_PARK> : pe_log_kernel<32, false, MAP_PARK>);
  else
    kernel = src.packed ? (wd == 8u ? pe_log_kernel<8, true, MAP_ALL> : wd == 16u ? pe_log_kernel<16, true, MAP_ALL> : pe_log_kernel<32, true, MAP_ALL>)
                        : (wd == 8u ? pe_log_kernel<8, false, MAP_ALL> : wd == 16u ? pe_log_kernel<16, false, MAP_ALL> : pe_log_kernel<32, false, MAP_ALL>);
  int rc = grid_for(e, kernel, smem, n, wd, &grid);
  if (rc) return rc;
  WALT_CUDA_TRY(cudaMemsetAsync(d_queue, 0, 4 * QUEUE_WORDS, st));
  kernel<<<grid, BLOCK_THREADS, smem, st>>>(a);
  WALT_CUDA_TRY(cudaGetLastError());
  e->stats.n_kernel_launches++;
  if (park)
    return launch_behind_park(e, a, e->verify_cg ? (src.packed ? verify_kernel<true, PeArgs, 0> : verify_kernel<false, PeArgs, 0>)
                                                 : (src.packed ? verify_kernel<true, PeArgs, 1> : verify_kernel<false, PeArgs, 1>),
                              src.packed ? pe_fold_kernel<true> : pe_fold_kernel<false>,
                              src.packed ? pe_log_kernel<32, true, MAP_TAKE> : pe_log_kernel<32, false, MAP_TAKE>,
                              verify_smem_bytes(a.nw_max), 0, pe_log_smem_bytes(a.nw_max, 32u), st);
  return WALT_OK;
}

static int launch_pe_mate(walt_engine* e, const ReadSrc& src, uint32_t n, int ag, uint32_t m, uint32_t b, uint32_t top_k,
                          walt_cand* d_ranked, uint32_t* d_nranked, uint32_t* d_queue, cudaStream_t st, bool zero_fill) {
  if (src.max_len > MAX_READ_LEN) return fail(WALT_EINVAL, "read longer than 1024 bases");
  PeArgs a;
  fill_common(e, a, src, n, ag, m, b, d_queue);
  a.top_k = top_k; a.ranked = d_ranked; a.n_ranked = d_nranked;
  a.log = nullptr; a.n_log = nullptr; a.log_slots = 0; a.zero_fill = zero_fill ? 1u : 0u;
  const uint32_t wd = e->group_width;
  const size_t smem = pe_smem_bytes(a.nw_max, top_k, wd);
  uint32_t grid = 0;
  auto kernel = src.packed ? (wd == 8u ? pe_map_kernel<8, true> : wd == 16u ? pe_map_kernel<16, true> : pe_map_kernel<32, true>)
                           : (wd == 8u ? pe_map_kernel<8, false> : wd == 16u ? pe_map_kernel<16, false> : pe_map_kernel<32, false>);
  int rc = grid_for(e, kernel, smem, n, wd, &grid);
  if (rc) return rc;
  WALT_CUDA_TRY(cudaMemsetAsync(d_queue, 0, 4, st));
  kernel<<<grid, BLOCK_THREADS, smem, st>>>(a);
  WALT_CUDA_TRY(cudaGetLastError());
  e->stats.n_kernel_launches++;
  return WALT_OK;
}

template <class T>
static int reserve(T** p, size_t* cap, size_t need) {
  if (need <= *cap) return WALT_OK;
  if (*p) cudaFree(*p);
  *p = nullptr; *cap = 0;
  size_t want = need + need / 8 + 256;
  WALT_CUDA_TRY(cudaMalloc(p, want * sizeof(T)));
  *cap = want;
  return WALT_OK;
}
static int reserve_bytes(void** p, size_t* cap, size_t need) {
  return reserve(reinterpret_cast<char**>(p), cap, need);
}

// room for the parked reads of a launch over n reads
static int reserve_park(walt_engine* e, ParkBuf* pk, uint32_t n) {
  int rc;
  if ((rc = reserve(&pk->list, &pk->list_cap, (size_t)n))) return rc;
  pk->rec_cap = e->flat_verify ? std::min<uint32_t>(n, REC_CAP_MAX) : 0u;
  pk->cap_blocks = e->flat_verify ? (uint32_t)std::min<uint64_t>(BLOCK_CAP_MAX, std::max<uint64_t>(1u << 22, 8ull * n)) : 0u;
  if (!pk->rec_cap) return WALT_OK;
  if ((rc = reserve(&pk->recs, &pk->recs_words, (size_t)pk->rec_cap * PARK_WORDS))) return rc;
  if ((rc = reserve(&pk->desc, &pk->desc_words, (size_t)pk->cap_blocks * 2u))) return rc;
  return reserve(&pk->bytes, &pk->bytes_cap, (size_t)pk->cap_blocks * 32u);
}

static int fetch_status(walt_engine* e) {
  uint32_t f = 0;
  WALT_CUDA_TRY(cudaMemcpy(&f, e->d_flags, 4, cudaMemcpyDeviceToHost));
  unsigned long long c[N_COUNTERS];
  WALT_CUDA_TRY(cudaMemcpy(c, e->d_counters, sizeof(c), cudaMemcpyDeviceToHost));
  e->stats.n_lookups = c[0]; e->stats.n_candidates = c[1]; e->stats.n_literal = c[2]; e->stats.n_parked = c[3];
  e->stats.n_verify_slots = c[4];
  // device time of the verify_kernel launches since the last fetch (walt_engine_set_kernel_timing)
  double ms = 0.0;
  for (auto& ev : e->verify_events) {
    float t = 0.f;
    if (cudaEventSynchronize(ev.second) == cudaSuccess && cudaEventElapsedTime(&t, ev.first, ev.second) == cudaSuccess) ms += t;
    cudaEventDestroy(ev.first); cudaEventDestroy(ev.second);
  }
  e->verify_events.clear();
  e->stats.verify_ns = (uint64_t)(ms * 1e6);
  if (f & 1u) {
    cudaMemset(e->d_flags, 0, 4);
    return fail(WALT_ENONACGT, "[ERROR: NON-ACGT NUCLEOTIDE] in a read handed to the mapping engine");
  }
  return WALT_OK;
}

// scan_chunk (one pass over a chunk's read offsets) lives in walt_hostscan.cpp.
//
// One core streams the offsets at ~5 GB/s: 80 MB for ten million reads is 15 ms, longer than the GPU
// needs to map them.  AheadScans runs the per-chunk scans of a batch on a few helper threads, ahead
// of the loop that queues the chunks; get(k) waits for chunk k's result.
struct AheadScans {
  std::vector<ChunkScan> res[2];
  std::unique_ptr<std::atomic<int>[]> ready;
  std::vector<std::thread> th;
  AheadScans(const uint64_t* offs1, const uint64_t* offs2, uint32_t n, uint32_t chunk) {
    const uint32_t nk = (n + chunk - 1) / chunk;
    res[0].resize(nk); res[1].resize(nk);
    ready.reset(new std::atomic<int>[nk ? nk : 1]);
    for (uint32_t k = 0; k < nk; ++k) ready[k].store(0, std::memory_order_relaxed);
    auto one = [=](uint32_t k) {
      const uint32_t r0 = k * chunk, cn = std::min<uint32_t>(chunk, n - r0);
      res[0][k] = scan_chunk(offs1, r0, cn);
      if (offs2) res[1][k] = scan_chunk(offs2, r0, cn);
      ready[k].store(1, std::memory_order_release);
    };
    const uint32_t T = nk >= 8 ? 4u : 0u;   // small batches: scanned inline by get()
    for (uint32_t t = 0; t < T; ++t)
      th.emplace_back([=]() { for (uint32_t k = t; k < nk; k += T) one(k); });
    if (!T) inline_one = one;
  }
  std::function<void(uint32_t)> inline_one;
  void get(uint32_t k, ChunkScan* a, ChunkScan* b) {
    if (inline_one) inline_one(k);
    while (!ready[k].load(std::memory_order_acquire)) std::this_thread::yield();
    *a = res[0][k];
    if (b) *b = res[1][k];
  }
  ~AheadScans() { for (auto& t : th) t.join(); }
};

}  // namespace waltb200

using namespace waltb200;

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

const char* walt_last_error(void) { return g_error.c_str(); }

int walt_device_count(void) {
  int n = 0;
  return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

int walt_engine_create(walt_engine** out, int device) {
  if (!out) return fail(WALT_EINVAL, "out is NULL");
  *out = nullptr;
  int n_dev = 0;
  cudaError_t ce = cudaGetDeviceCount(&n_dev);
  if (ce != cudaSuccess || n_dev == 0)
    return fail(WALT_ECUDA, std::string("no CUDA device: ") + cudaGetErrorString(ce) +
                                " (the engine has no CPU fallback)");
  if (device < 0 || device >= n_dev) return fail(WALT_EINVAL, "device ordinal out of range");
  std::unique_ptr<walt_engine> e(new walt_engine);
  e->device = device;
  WALT_CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  WALT_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  e->sm_count = prop.multiProcessorCount;
  // undocumented tuning knobs for experiments (defaults are what bench.py measures)
  if (const char* v = getenv("WALT_MIN_BLOCKS")) e->min_blocks = atoi(v);
  if (const char* v = getenv("WALT_PE_SIDE")) e->pe_side = atoi(v);
  if (const char* v = getenv("WALT_PE_LOGGED")) e->pe_logged = atoi(v);
  if (const char* v = getenv("WALT_DEFER")) e->defer = atoi(v);
  if (const char* v = getenv("WALT_FLAT")) e->flat_verify = atoi(v);
  if (const char* v = getenv("WALT_VERIFY_CG")) e->verify_cg = atoi(v);
  if (const char* v = getenv("WALT_PAIR_WIDE")) e->pair_wide = atoi(v);
  if (const char* v = getenv("WALT_HEAP_SMEM")) e->heap_smem = atoi(v);
  if (const char* v = getenv("WALT_CHUNK_SHARE")) e->chunk_share = (uint32_t)std::max(1, atoi(v));
  if (const char* v = getenv("WALT_SE_SLOTS")) e->se_slots = (uint32_t)std::min<int>(N_SLOTS, std::max(1, atoi(v)));
  if (const char* v = getenv("WALT_PE_SLOTS")) e->pe_slots = (uint32_t)std::min<int>(N_SLOTS, std::max(1, atoi(v)));
  if (const char* v = getenv("WALT_L2_FETCH")) WALT_CUDA_TRY(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(v)));
  uint32_t p = 1;
  for (uint32_t i = 0; i <= MAX_DEPTH; ++i) { e->pow3.v[i] = p; p *= 3u; }
  WALT_CUDA_TRY(cudaMalloc(&e->d_flags, N_FLAG_WORDS * 4));
  WALT_CUDA_TRY(cudaMemset(e->d_flags, 0, N_FLAG_WORDS * 4));
  WALT_CUDA_TRY(cudaMalloc(&e->d_counters, N_COUNTERS * 8));
  WALT_CUDA_TRY(cudaMemset(e->d_counters, 0, N_COUNTERS * 8));
  for (auto& s : e->slot) {
    WALT_CUDA_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    WALT_CUDA_TRY(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
  }
  WALT_CUDA_TRY(cudaStreamCreateWithFlags(&e->side_stream, cudaStreamNonBlocking));
  WALT_CUDA_TRY(cudaEventCreateWithFlags(&e->fork, cudaEventDisableTiming));
  WALT_CUDA_TRY(cudaEventCreateWithFlags(&e->join, cudaEventDisableTiming));
  *out = e.release();
  return WALT_OK;
}

void walt_engine_destroy(walt_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  for (auto& s : e->sub) s.release();
  for (auto& s : e->slot) {
    cudaFree(s.d_seqs); cudaFree(s.d_offs); cudaFree(s.d_out); cudaFree(s.d_seqs2); cudaFree(s.d_offs2);
    cudaFree(s.d_pe);
    for (auto& p : s.park) p.release();
    if (s.stream) cudaStreamDestroy(s.stream);
    if (s.done) cudaEventDestroy(s.done);
  }
  if (e->side_stream) cudaStreamDestroy(e->side_stream);
  if (e->fork) cudaEventDestroy(e->fork);
  if (e->join) cudaEventDestroy(e->join);
  cudaFree(e->d_starts); cudaFree(e->d_flags); cudaFree(e->d_counters);
  for (auto& p : e->dev_park) p.release();
  delete e;
}

int walt_engine_set_chromosomes(walt_engine* e, uint32_t n_chr, const uint32_t* lengths, const char* const* names) {
  if (!e || !lengths || n_chr == 0) return fail(WALT_EINVAL, "bad chromosome table");
  int rc = ensure_device(e);
  if (rc) return rc;
  uint64_t total = 0;
  for (uint32_t i = 0; i < n_chr; ++i) total += lengths[i];
  if (total >= 0xFFFFFFFFull - MAX_READ_LEN) return fail(WALT_EFORMAT, "genome does not fit 32-bit offsets");
  e->n_chr = n_chr;
  e->lengths.assign(lengths, lengths + n_chr);
  e->starts.assign(n_chr + 1, 0);
  for (uint32_t i = 0; i < n_chr; ++i) e->starts[i + 1] = e->starts[i] + lengths[i];   // reference.cpp:407-410
  e->genome_len = e->starts[n_chr];
  e->names.clear();
  for (uint32_t i = 0; i < n_chr; ++i) e->names.emplace_back(names && names[i] ? names[i] : "");
  e->name_ptrs.clear();
  for (auto& s : e->names) e->name_ptrs.push_back(s.c_str());
  cudaFree(e->d_starts); e->d_starts = nullptr;
  WALT_CUDA_TRY(cudaMalloc(&e->d_starts, (n_chr + 1) * 4u));
  WALT_CUDA_TRY(cudaMemcpy(e->d_starts, e->starts.data(), (n_chr + 1) * 4u, cudaMemcpyHostToDevice));
  for (auto& s : e->sub) s.release();
  return WALT_OK;
}

// pack `n` ASCII bases starting at genome base `first` (multiple of 32) from host memory
static int upload_genome_chunk(walt_engine* e, DeviceSubIndex& s, int which, const char* host, uint64_t first,
                               uint64_t n, uint8_t* d_stage) {
  WALT_CUDA_TRY(cudaMemcpy(d_stage, host, n, cudaMemcpyHostToDevice));
  const uint64_t words = (n + 31u) / 32u;
  const uint32_t T = 256;
  pack_ascii_kernel<<<(uint32_t)((words + T - 1) / T), T>>>(d_stage, n, s.genome + ((first + PAD_BASES) >> 5),
                                                          which >= 2 ? 2u : 1u, e->d_flags + 3);
  WALT_CUDA_TRY(cudaGetLastError());
  return WALT_OK;
}

static const size_t STAGE_BYTES = 64u << 20;

int walt_engine_load_subindex(walt_engine* e, int which, const char* sequence, const uint32_t* counter,
                              const uint32_t* index, uint32_t index_size) {
  if (!e || which < 0 || which > 3 || !sequence || (!index && index_size)) return fail(WALT_EINVAL, "bad argument");
  if (!e->d_starts) return fail(WALT_EINVAL, "set the chromosome table first");
  int rc = ensure_device(e);
  if (rc) return rc;
  if (counter && counter[1u << 24] != index_size)
    return fail(WALT_EFORMAT, "counter[4^12] != index_size (reference.cpp:252-255)");
  DeviceSubIndex& s = e->sub[which];
  s.release();
  if ((rc = alloc_packed_genome(e, s))) return rc;
  uint8_t* d_stage = nullptr;
  WALT_CUDA_TRY(cudaMalloc(&d_stage, STAGE_BYTES));
  if (cudaMemset(e->d_flags + 3, 0, 4) != cudaSuccess) { cudaFree(d_stage); return fail(WALT_ECUDA, "cudaMemset"); }
  for (uint64_t off = 0; off < e->genome_len; off += STAGE_BYTES) {
    const uint64_t n = std::min<uint64_t>(STAGE_BYTES, e->genome_len - off);
    if ((rc = upload_genome_chunk(e, s, which, sequence + off, off, n, d_stage))) { cudaFree(d_stage); return rc; }
  }
  uint32_t bad = 0;
  const cudaError_t ce = cudaMemcpy(&bad, e->d_flags + 3, 4, cudaMemcpyDeviceToHost);
  cudaFree(d_stage);
  if (ce != cudaSuccess) return fail(WALT_ECUDA, std::string("genome check: ") + cudaGetErrorString(ce));
  if (bad) return fail(WALT_EFORMAT, std::to_string(bad) + " genome bytes outside the sub-index's 3-letter alphabet");
  s.index_size = index_size;
  WALT_CUDA_TRY(cudaMalloc(&s.index, ((size_t)index_size + 64u) * 4u));
  WALT_CUDA_TRY(cudaMemset(s.index + index_size, 0, 64u * 4u));
  if (index_size) WALT_CUDA_TRY(cudaMemcpy(s.index, index, (size_t)index_size * 4u, cudaMemcpyHostToDevice));
  return finalize_subindex(e, which);
}

static int read_exact(FILE* f, void* dst, size_t n, const char* what) {
  if (fread(dst, 1, n, f) != n) return fail(WALT_EIO, std::string("short read: ") + what);
  return WALT_OK;
}

extern "C++" {
// Pinned staging ring for the one-time index upload.  A few reader threads pread pieces of a
// sub-index file into the ring in parallel (a single thread copies out of the page cache at a few
// GB/s; the PCIe link takes ten times that), the calling thread consumes the pieces in file order.
struct StageRing {
  static constexpr size_t PIECE = 32u << 20;
  static constexpr int SLOTS = 8, READERS = 6;
  char* slot[SLOTS] = {};
  ~StageRing() { for (char* p : slot) if (p) cudaFreeHost(p); }
  int init() {
    for (int i = 0; i < SLOTS; ++i) WALT_CUDA_TRY(cudaMallocHost(&slot[i], PIECE));
    return WALT_OK;
  }
  // consume(host, offset_in_range, n) is called for consecutive pieces of [off, off + len) of fd;
  // it must be done with `host` when it returns.
  template <class F>
  int stream(int fd, uint64_t off, uint64_t len, const char* what, F consume) {
    const uint64_t np = (len + PIECE - 1) / PIECE;
    if (np == 0) return WALT_OK;
    std::mutex mu;
    std::condition_variable cv;
    std::vector<int> state(np, 0);          // 0 not read, 1 ready, 2 consumed, -1 read error
    uint64_t consumed = 0;                  // pieces [0, consumed) are done with their slots
    bool abort = false;
    std::atomic<uint64_t> next{0};
    auto reader = [&]() {
      for (;;) {
        const uint64_t k = next.fetch_add(1);
        if (k >= np) return;
        {
          std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [&]() { return abort || k < consumed + SLOTS; });
          if (abort) return;
        }
        const uint64_t p0 = k * PIECE, n = std::min<uint64_t>(PIECE, len - p0);
        char* dst = slot[k % SLOTS];
        uint64_t got = 0;
        bool ok = true;
        while (got < n) {
          const ssize_t r = pread(fd, dst + got, n - got, (off_t)(off + p0 + got));
          if (r < 0 && errno == EINTR) continue;
          if (r <= 0) { ok = false; break; }
          got += (uint64_t)r;
        }
        {
          std::lock_guard<std::mutex> lk(mu);
          state[k] = ok ? 1 : -1;
        }
        cv.notify_all();
      }
    };
    std::vector<std::thread> th;
    const int nr = (int)std::min<uint64_t>(READERS, np);
    for (int i = 0; i < nr; ++i) th.emplace_back(reader);
    int rc = WALT_OK;
    for (uint64_t k = 0; k < np && !rc; ++k) {
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&]() { return state[k] != 0; });
        if (state[k] < 0) rc = fail(WALT_EIO, std::string("short read: ") + what);
      }
      if (!rc) rc = consume(slot[k % SLOTS], k * PIECE, std::min<uint64_t>(PIECE, len - k * PIECE));
      {
        std::lock_guard<std::mutex> lk(mu);
        consumed = k + 1;
        if (rc) abort = true;
      }
      cv.notify_all();
    }
    for (auto& t : th) t.join();
    return rc;
  }
};
}  // extern "C++"

// ReadIndexHeadInfo (reference.cpp:381-417) + ReadIndex (reference.cpp:324-351), once.
int walt_engine_load_dbindex(walt_engine* e, const char* path, uint32_t which_mask) {
  if (!e || !path) return fail(WALT_EINVAL, "bad argument");
  int rc = ensure_device(e);
  if (rc) return rc;
  FILE* f = fopen(path, "rb");
  if (!f) return fail(WALT_EIO, std::string("cannot open ") + path + ": " + strerror(errno));
  uint32_t n_chr = 0;
  std::vector<std::string> names;
  std::vector<uint32_t> lengths;
  uint32_t genome_len = 0, size_of_index = 0;
  rc = read_exact(f, &n_chr, 4, "chromosome count");
  if (!rc && (n_chr == 0 || n_chr > (1u << 24))) rc = fail(WALT_EFORMAT, "implausible chromosome count");
  for (uint32_t i = 0; !rc && i < n_chr; ++i) {
    uint32_t ln = 0;
    rc = read_exact(f, &ln, 4, "name length");
    if (!rc && ln > 255) rc = fail(WALT_EFORMAT, "chromosome name longer than 255");
    std::string nm(ln, '\0');
    if (!rc && ln) rc = read_exact(f, &nm[0], ln, "chromosome name");
    names.push_back(nm);
  }
  if (!rc) {
    lengths.resize(n_chr);
    rc = read_exact(f, lengths.data(), 4u * n_chr, "chromosome lengths");
  }
  if (!rc) rc = read_exact(f, &genome_len, 4, "genome length");
  if (!rc) rc = read_exact(f, &size_of_index, 4, "index size");
  fclose(f);
  if (rc) return rc;
  std::vector<const char*> np;
  for (auto& s : names) np.push_back(s.c_str());
  if ((rc = walt_engine_set_chromosomes(e, n_chr, lengths.data(), np.data()))) return rc;
  if (e->genome_len != genome_len) return fail(WALT_EFORMAT, "length_of_genome != sum of chromosome lengths");

  static const char* SFX[4] = {"_CT00", "_CT01", "_GA10", "_GA11"};
  uint8_t* d_stage = nullptr;
  StageRing ring;
  if ((rc = ring.init())) return rc;
  WALT_CUDA_TRY(cudaMalloc(&d_stage, StageRing::PIECE));
  for (int which = 0; which < 4 && !rc; ++which) {
    if (!((which_mask >> which) & 1u)) continue;
    const std::string sp = std::string(path) + SFX[which];
    const int fd = open(sp.c_str(), O_RDONLY);
    if (fd < 0) { rc = fail(WALT_EIO, "cannot open " + sp + ": " + strerror(errno)); break; }
    auto pread_exact = [&](void* dst, size_t n, uint64_t at, const char* what) {
      size_t got = 0;
      while (got < n) {
        const ssize_t r = pread(fd, (char*)dst + got, n - got, (off_t)(at + got));
        if (r < 0 && errno == EINTR) continue;
        if (r <= 0) return fail(WALT_EIO, std::string("short read: ") + what);
        got += (size_t)r;
      }
      return (int)WALT_OK;
    };
    // file layout (reference.cpp:302-322): strand byte, genome, counter_size, index_size, counter[], index[]
    DeviceSubIndex& s = e->sub[which];
    s.release();
    char strand = 0;
    rc = pread_exact(&strand, 1, 0, "strand byte");
    if (!rc && strand != ((which & 1) ? '-' : '+')) rc = fail(WALT_EFORMAT, sp + ": wrong strand byte");
    if (!rc) rc = alloc_packed_genome(e, s);
    if (!rc && cudaMemset(e->d_flags + 3, 0, 4) != cudaSuccess) rc = fail(WALT_ECUDA, "memset");
    if (!rc)
      rc = ring.stream(fd, 1, genome_len, "genome sequence", [&](const char* host, uint64_t off, uint64_t n) {
        return upload_genome_chunk(e, s, which, host, off, n, d_stage);
      });
    uint32_t hdr[2] = {0, 0};
    const uint64_t at_hdr = 1u + (uint64_t)genome_len;
    if (!rc) rc = pread_exact(hdr, 8, at_hdr, "counter/index sizes");
    if (!rc && hdr[0] != (1u << 24)) rc = fail(WALT_EFORMAT, sp + ": counter_size != 4^12");
    if (!rc && hdr[1] > size_of_index) rc = fail(WALT_EFORMAT, sp + ": index_size exceeds the header's size_of_index");
    // counter[] is superseded by the base-3 prefix table; only its last entry is checked
    uint32_t last = 0;
    if (!rc) rc = pread_exact(&last, 4, at_hdr + 8u + (uint64_t)4 * (1u << 24), "counter[4^12]");
    if (!rc && last != hdr[1]) rc = fail(WALT_EFORMAT, sp + ": counter[4^12] != index_size");
    if (!rc) {
      s.index_size = hdr[1];
      if (cudaMalloc(&s.index, ((size_t)s.index_size + 64u) * 4u) != cudaSuccess) rc = fail(WALT_ECUDA, "cudaMalloc(index)");
      else cudaMemset(s.index + s.index_size, 0, 64u * 4u);
    }
    if (!rc)
      rc = ring.stream(fd, at_hdr + 8u + (uint64_t)4 * ((1u << 24) + 1u), (uint64_t)s.index_size * 4u, "index[]",
                       [&](const char* host, uint64_t off, uint64_t n) {
                         if (cudaMemcpy((char*)s.index + off, host, n, cudaMemcpyHostToDevice) != cudaSuccess)
                           return fail(WALT_ECUDA, "index upload");
                         return (int)WALT_OK;
                       });
    close(fd);
    if (!rc) {
      uint32_t bad = 0;
      cudaMemcpy(&bad, e->d_flags + 3, 4, cudaMemcpyDeviceToHost);
      if (bad) rc = fail(WALT_EFORMAT, sp + ": genome bytes outside the 3-letter alphabet");
    }
    if (!rc) rc = finalize_subindex(e, which);
  }
  cudaFree(d_stage);
  return rc;
}

int walt_engine_clone_index(walt_engine* dst, const walt_engine* src) {
  if (!dst || !src || dst == src) return fail(WALT_EINVAL, "bad argument");
  if (!src->d_starts) return fail(WALT_ENOTLOADED, "the source engine holds no index");
  int rc;
  std::vector<const char*> names;
  for (auto& n : src->names) names.push_back(n.c_str());
  if ((rc = walt_engine_set_chromosomes(dst, src->n_chr, src->lengths.data(), names.data()))) return rc;   // also releases dst's sub-indexes
  if (dst->device != src->device) {
    int can = 0;
    cudaDeviceCanAccessPeer(&can, dst->device, src->device);
    if (can) {
      const cudaError_t ce = cudaDeviceEnablePeerAccess(src->device, 0);
      if (ce != cudaSuccess && ce != cudaErrorPeerAccessAlreadyEnabled) return fail(WALT_ECUDA, std::string("peer access: ") + cudaGetErrorString(ce));
      cudaGetLastError();
    }
  }
  auto pull = [&](void** d, const void* s, size_t bytes) -> int {
    *d = nullptr;
    if (!bytes || !s) return WALT_OK;
    WALT_CUDA_TRY(cudaMalloc(d, bytes));
    WALT_CUDA_TRY(cudaMemcpyPeer(*d, dst->device, s, src->device, bytes));
    return WALT_OK;
  };
  for (int which = 0; which < 4; ++which) {
    const DeviceSubIndex& a = src->sub[which];
    if (!a.loaded) continue;
    DeviceSubIndex& d = dst->sub[which];
    d.genome_words = a.genome_words; d.index_size = a.index_size; d.depth = a.depth; d.n_taint = a.n_taint;
    d.n_taint_keys = a.n_taint_keys; d.taint_words = a.taint_words; d.bytes = a.bytes;
    if ((rc = pull((void**)&d.genome, a.genome, a.genome_words * 8u)) ||
        (rc = pull((void**)&d.entries, a.entries, ((size_t)a.index_size + 64u) * sizeof(Entry))) ||
        (rc = pull((void**)&d.table, a.table, ((size_t)dst->pow3.v[a.depth] + 1u) * 4u)) ||
        (rc = pull((void**)&d.taint_bits, a.taint_bits, a.taint_words * 4u))) {
      d.release();
      return rc;
    }
    d.loaded = true;
  }
  WALT_CUDA_TRY(cudaDeviceSynchronize());
  return WALT_OK;
}

int walt_engine_chromosomes(const walt_engine* e, uint32_t* n_chr, const uint32_t** lengths,
                            const uint32_t** start_index, const char* const** names) {
  if (!e) return fail(WALT_EINVAL, "engine is NULL");
  if (n_chr) *n_chr = e->n_chr;
  if (lengths) *lengths = e->lengths.data();
  if (start_index) *start_index = e->starts.data();
  if (names) *names = e->name_ptrs.data();
  return WALT_OK;
}

uint64_t walt_engine_hbm_bytes(const walt_engine* e) {
  uint64_t t = 0;
  if (e) for (auto& s : e->sub) t += s.bytes;
  return t;
}

int walt_engine_subindex_info(const walt_engine* e, int which, uint32_t* index_size, uint32_t* depth,
                              uint32_t* n_taint) {
  if (!e || which < 0 || which > 3) return fail(WALT_EINVAL, "bad argument");
  if (!e->sub[which].loaded) return fail(WALT_ENOTLOADED, "sub-index not resident");
  if (index_size) *index_size = e->sub[which].index_size;
  if (depth) *depth = e->sub[which].depth;
  if (n_taint) *n_taint = e->sub[which].n_taint;
  return WALT_OK;
}

int walt_engine_set_search_mode(walt_engine* e, int mode) {
  if (!e || (mode != 0 && mode != 1)) return fail(WALT_EINVAL, "search mode must be 0 or 1");
  e->search_mode = mode;
  return WALT_OK;
}

int walt_engine_set_table_depth(walt_engine* e, int depth) {
  if (!e || (depth != 0 && (depth < (int)KEY_WEIGHT || depth > (int)MAX_DEPTH)))
    return fail(WALT_EINVAL, "table depth must be 0 (auto) or 12..20");
  e->force_depth = depth;
  return WALT_OK;
}

int walt_engine_set_tie_order(walt_engine* e, int mode) {
  if (!e || (mode != 0 && mode != 1)) return fail(WALT_EINVAL, "tie order must be 0 (std::sort) or 1 (ascending position)");
  e->tie_order = mode;
  return WALT_OK;
}

int walt_engine_last_build_info(const walt_engine* e, uint64_t* n_tied_slots, uint64_t* n_buckets_replayed) {
  if (!e) return fail(WALT_EINVAL, "bad argument");
  if (n_tied_slots) *n_tied_slots = e->last_build_ties;
  if (n_buckets_replayed) *n_buckets_replayed = e->last_build_replayed;
  return WALT_OK;
}

int walt_engine_set_group_width(walt_engine* e, uint32_t lanes) {
  if (!e || (lanes != 8u && lanes != 16u && lanes != 32u)) return fail(WALT_EINVAL, "group width must be 8, 16 or 32");
  e->group_width = lanes;
  return WALT_OK;
}

int walt_engine_set_defer(walt_engine* e, int on) {
  if (!e || (on != 0 && on != 1)) return fail(WALT_EINVAL, "defer must be 0 or 1");
  e->defer = on;
  return WALT_OK;
}

int walt_engine_set_kernel_timing(walt_engine* e, int on) {
  if (!e) return fail(WALT_EINVAL, "bad argument");
  e->time_kernels = on != 0;
  return WALT_OK;
}

int walt_engine_set_chunk_reads(walt_engine* e, uint32_t n) {
  if (!e) return fail(WALT_EINVAL, "bad argument");
  e->chunk_reads = n;   // 0 = automatic
  return WALT_OK;
}

int walt_engine_device_stats(walt_engine* e, walt_stats* out) {
  if (!e || !out) return fail(WALT_EINVAL, "bad argument");
  int rc = ensure_device(e);
  if (rc) return rc;
  WALT_CUDA_TRY(cudaDeviceSynchronize());
  const uint64_t launches = e->stats.n_kernel_launches;
  if ((rc = fetch_status(e))) return rc;
  WALT_CUDA_TRY(cudaMemset(e->d_counters, 0, N_COUNTERS * 8));
  *out = e->stats;
  out->n_kernel_launches = launches;
  return WALT_OK;
}

int walt_engine_last_stats(const walt_engine* e, walt_stats* out) {
  if (!e || !out) return fail(WALT_EINVAL, "bad argument");
  *out = e->stats;
  return WALT_OK;
}

// ---- single end ---------------------------------------------------------------------------
int walt_engine_map_se_device(walt_engine* e, const void* d_seqs, const void* d_offs, uint32_t n,
                              uint32_t max_read_len, int ag_wildcard, uint32_t max_mismatches, uint32_t b,
                              void* d_out, void* cuda_stream) {
  if (!e || (n && (!d_seqs || !d_offs || !d_out))) return fail(WALT_EINVAL, "bad argument");
  int rc = ensure_device(e);
  if (rc) return rc;
  if ((rc = check_pair(e, ag_wildcard))) return rc;
  if (n == 0) return WALT_OK;
  e->stats.n_kernel_launches = 0;
  const ReadSrc src{(const char*)d_seqs, (const uint64_t*)d_offs, 0, 0, 0, max_read_len, false};
  if (e->defer && (rc = reserve_park(e, &e->dev_park[0], n))) return rc;
  return launch_se(e, src, n, ag_wildcard, max_mismatches, b, (walt_best*)d_out, queue_block(e, QB_SE_DEVICE),
                   e->defer ? &e->dev_park[0] : nullptr, (cudaStream_t)cuda_stream);
}

// Bytes [*b0, *b1) of a batch buffer that hold reads [r0, r0 + cn): ASCII, or the 2-bit packed
// form where read j starts at byte (offs[j] >> 2) + j.
static void chunk_bytes(const uint64_t* offs, uint32_t r0, uint32_t cn, bool packed, uint64_t* b0, uint64_t* b1) {
  if (packed) { *b0 = (offs[r0] >> 2) + r0; *b1 = (offs[r0 + cn] >> 2) + r0 + cn; }
  else { *b0 = offs[r0]; *b1 = offs[r0 + cn]; }
}

// Every copy and kernel the host-batch loops have queued is waited for before an error goes back to the
// caller: the chunks in flight still read the caller's input and write its output buffers.
static void drain_slots(walt_engine* e) {
  for (auto& s : e->slot) cudaStreamSynchronize(s.stream);
  cudaStreamSynchronize(e->side_stream);
}

static int map_se_host_loop(walt_engine* e, const char* seqs, const uint64_t* offs, uint32_t n, bool packed, int ag_wildcard,
                            uint32_t max_mismatches, uint32_t b, walt_best* out, uint32_t* n_short);
static int map_se_host(walt_engine* e, const char* seqs, const uint64_t* offs, uint32_t n, bool packed, int ag_wildcard,
                       uint32_t max_mismatches, uint32_t b, walt_best* out, uint32_t* n_short) {
  const int rc = map_se_host_loop(e, seqs, offs, n, packed, ag_wildcard, max_mismatches, b, out, n_short);
  if (rc && e) { const std::string msg = g_error; drain_slots(e); g_error = msg; }
  return rc;
}

static int map_se_host_loop(walt_engine* e, const char* seqs, const uint64_t* offs, uint32_t n, bool packed, int ag_wildcard,
                            uint32_t max_mismatches, uint32_t b, walt_best* out, uint32_t* n_short) {
  if (!e || !offs || (n && (!seqs || !out))) return fail(WALT_EINVAL, "bad argument");
  int rc = ensure_device(e);
  if (rc) return rc;
  if ((rc = check_pair(e, ag_wildcard))) return rc;
  e->stats = walt_stats{};
  WALT_CUDA_TRY(cudaMemsetAsync(e->d_counters, 0, N_COUNTERS * 8, e->slot[0].stream));
  WALT_CUDA_TRY(cudaStreamSynchronize(e->slot[0].stream));
  uint32_t k = 0, total_short = 0;
  // chunk size: ASCII batches are PCIe-bound and like a short pipeline fill; packed batches are
  // kernel-bound and like fewer launches (measured: 2^18 resp. 2^19 reads)
  const uint32_t chunk = e->chunk_reads ? e->chunk_reads : (packed ? 1u << 19 : 1u << 18);
  AheadScans scans(offs, nullptr, n, chunk);
  for (uint32_t r0 = 0; r0 < n; r0 += chunk, ++k) {
    const uint32_t cn = std::min<uint32_t>(chunk, n - r0);
    ChunkScan sc;
    scans.get(k, &sc, nullptr);
    if (sc.max_len > MAX_READ_LEN) {
      for (auto& s : e->slot) cudaStreamSynchronize(s.stream);
      return fail(WALT_EINVAL, "read longer than 1024 bases");
    }
    total_short += sc.n_short;
    BatchSlot& s = e->slot[k % e->se_slots];
    WALT_CUDA_TRY(cudaEventSynchronize(s.done));
    uint64_t sb, se;
    chunk_bytes(offs, r0, cn, packed, &sb, &se);
    if ((rc = reserve(&s.d_seqs, &s.seqs_cap, (size_t)(se - sb) + 16u))) return rc;
    if ((rc = reserve_bytes(&s.d_out, &s.out_cap, (size_t)cn * sizeof(walt_best)))) return rc;
    if (se > sb) WALT_CUDA_TRY(cudaMemcpyAsync(s.d_seqs, seqs + sb, se - sb, cudaMemcpyHostToDevice, s.stream));
    if (!sc.uniform_len) {   // equal-length reads are addressed by stride: no offsets cross PCIe
      if ((rc = reserve(&s.d_offs, &s.offs_cap, (size_t)cn + 1u))) return rc;
      WALT_CUDA_TRY(cudaMemcpyAsync(s.d_offs, offs + r0, ((size_t)cn + 1u) * 8u, cudaMemcpyHostToDevice, s.stream));
    }
    const ReadSrc src{s.d_seqs, s.d_offs, offs[r0], sc.uniform_len, 0, sc.max_len, packed};
    if (e->defer && (rc = reserve_park(e, &s.park[0], cn))) return rc;
    if ((rc = launch_se(e, src, cn, ag_wildcard, max_mismatches, b, (walt_best*)s.d_out, queue_block(e, QB_SE_SLOT + k % e->se_slots),
                        e->defer ? &s.park[0] : nullptr, s.stream, e->chunk_share)))
      return rc;
    WALT_CUDA_TRY(cudaMemcpyAsync(out + r0, s.d_out, (size_t)cn * sizeof(walt_best), cudaMemcpyDeviceToHost, s.stream));
    WALT_CUDA_TRY(cudaEventRecord(s.done, s.stream));
  }
  if (n_short) *n_short = 2u * total_short;   // once per strand pass (mapping.cpp:230-232)
  for (auto& s : e->slot) WALT_CUDA_TRY(cudaStreamSynchronize(s.stream));
  return fetch_status(e);
}

int walt_engine_map_se(walt_engine* e, const char* seqs, const uint64_t* offs, uint32_t n, int ag_wildcard,
                       uint32_t max_mismatches, uint32_t b, walt_best* out, uint32_t* n_short) {
  return map_se_host(e, seqs, offs, n, false, ag_wildcard, max_mismatches, b, out, n_short);
}

int walt_engine_map_se_packed(walt_engine* e, const uint8_t* packed, const uint64_t* offs, uint32_t n, int ag_wildcard,
                              uint32_t max_mismatches, uint32_t b, walt_best* out, uint32_t* n_short) {
  return map_se_host(e, (const char*)packed, offs, n, true, ag_wildcard, max_mismatches, b, out, n_short);
}

// ---- paired end ---------------------------------------------------------------------------
// device layout of one chunk's paired-end scratch inside a slot's d_pe allocation
// pe_heap_kernel keeps its heaps in shared memory while two blocks fit on an SM (top_k <= 100), else in global scratch
constexpr uint32_t HEAP_SMEM_TOPK = 100;
static bool heaps_in_smem(const walt_engine* e, uint32_t top_k) { return e->heap_smem && top_k <= HEAP_SMEM_TOPK; }

static bool pe_two_phase(const walt_engine* e, uint32_t m) { return e->pe_logged && m <= LOG_MAX_MM; }

struct PeScratch {
  walt_pair* pairs; walt_pe_result* compact; walt_cand* r1; walt_cand* r2; uint32_t* n1; uint32_t* n2;
  // two-phase form: per-mate logs and their lengths, heaps when they do not fit shared memory
  HeapEntry* log1; HeapEntry* log2; uint32_t* nlog1; uint32_t* nlog2; HeapEntry* heaps;
};
// bytes of scratch one pair needs (sizes the chunks)
static size_t pe_pair_bytes(uint32_t top_k, uint32_t m, bool two_phase, bool gheaps) {
  size_t b = sizeof(walt_pe_result) + sizeof(walt_pair) + 2u * (size_t)top_k * sizeof(walt_cand) + 2u * 4u;
  if (two_phase) {
    b += 2u * (size_t)pe_log_slots(top_k, m) * sizeof(HeapEntry) + 2u * 4u;
    if (gheaps) b += 2u * (size_t)top_k * sizeof(HeapEntry);
  }
  return b;
}
static size_t pe_scratch_bytes(uint32_t cn, uint32_t top_k, uint32_t m, bool two_phase, bool gheaps) {
  return (size_t)cn * pe_pair_bytes(top_k, m, two_phase, gheaps) + 512u;
}
static PeScratch carve_pe(void* base, uint32_t cn, uint32_t top_k, uint32_t m, bool two_phase, bool gheaps) {
  PeScratch p;
  char* c = (char*)base;
  auto take = [&](size_t bytes) { char* at = c; c += (bytes + 15u) & ~(size_t)15u; return at; };
  p.compact = (walt_pe_result*)take((size_t)cn * sizeof(walt_pe_result));
  p.pairs = (walt_pair*)take((size_t)cn * sizeof(walt_pair));
  p.r1 = (walt_cand*)take((size_t)cn * top_k * sizeof(walt_cand));
  p.r2 = (walt_cand*)take((size_t)cn * top_k * sizeof(walt_cand));
  p.n1 = (uint32_t*)take((size_t)cn * 4u);
  p.n2 = (uint32_t*)take((size_t)cn * 4u);
  p.log1 = p.log2 = p.heaps = nullptr; p.nlog1 = p.nlog2 = nullptr;
  if (two_phase) {
    const size_t slots = pe_log_slots(top_k, m);
    p.log1 = (HeapEntry*)take((size_t)cn * slots * sizeof(HeapEntry));
    p.log2 = (HeapEntry*)take((size_t)cn * slots * sizeof(HeapEntry));
    p.nlog1 = (uint32_t*)take((size_t)cn * 4u);
    p.nlog2 = (uint32_t*)take((size_t)cn * 4u);
    if (gheaps) p.heaps = (HeapEntry*)take(2u * (size_t)cn * top_k * sizeof(HeapEntry));
  }
  return p;
}

// both mate kernels + the pairing kernel for one chunk resident on the device
static int launch_pe_chunk(walt_engine* e, const ReadSrc& m1, const ReadSrc& m2, uint32_t cn, uint32_t m, uint32_t b,
                           uint32_t top_k, int frag_range, int swap, const PeScratch& ps, bool want_pairs,
                           walt_pe_result* d_compact, uint32_t* q, const ParkBuf* pk, cudaStream_t st) {
  int rc;
  // mate 1: C->T against _CT00/_CT01; mate 2: G->A against _GA10/_GA11 (paired.cpp:642-672).  The
  // two mate kernels are independent: the second runs on a side stream so that its blocks fill
  // the SMs the first one's tail (a few repeat-heavy reads) leaves idle.
  cudaStream_t st2 = e->pe_side ? e->side_stream : st;
  if (e->pe_side) {
    WALT_CUDA_TRY(cudaEventRecord(e->fork, st));
    WALT_CUDA_TRY(cudaStreamWaitEvent(st2, e->fork, 0));
  }
  const bool two_phase = ps.log1 != nullptr;
  if (two_phase) {
    if ((rc = launch_pe_log(e, m1, cn, 0, m, b, top_k, ps.log1, ps.nlog1, q, pk ? pk : nullptr, st))) return rc;
    if ((rc = launch_pe_log(e, m2, cn, 1, m, b, top_k, ps.log2, ps.nlog2, q + QUEUE_WORDS, pk ? pk + 1 : nullptr, st2))) return rc;
  } else {
    if ((rc = launch_pe_mate(e, m1, cn, 0, m, b, top_k, ps.r1, ps.n1, q, st, want_pairs))) return rc;
    if ((rc = launch_pe_mate(e, m2, cn, 1, m, b, top_k, ps.r2, ps.n2, q + QUEUE_WORDS, st2, want_pairs))) return rc;
  }
  if (e->pe_side) {
    WALT_CUDA_TRY(cudaEventRecord(e->join, st2));
    WALT_CUDA_TRY(cudaStreamWaitEvent(st, e->join, 0));
  }
  if (two_phase) {
    HeapArgs h;
    h.log[0] = ps.log1; h.log[1] = ps.log2; h.n_log[0] = ps.nlog1; h.n_log[1] = ps.nlog2;
    h.ranked[0] = ps.r1; h.ranked[1] = ps.r2; h.n_ranked[0] = ps.n1; h.n_ranked[1] = ps.n2;
    h.heaps = ps.heaps; h.n = cn; h.top_k = top_k; h.log_slots = pe_log_slots(top_k, m);
    h.zero_fill = want_pairs ? 1u : 0u;   // only walt_engine_map_pe hands the lists themselves to the host
    const bool parked = pk != nullptr && e->defer;
    for (int i = 0; i < 2; ++i) {
      h.list[i] = parked ? pk[i].list : nullptr;
      h.recs[i] = parked ? pk[i].recs : nullptr;
      h.queue[i] = q + QUEUE_WORDS * i;
    }
    // SMEM: heaps of top_k entries per thread in shared memory (local memory heaps of 16 resident blocks thrash the L1)
    const bool smem = ps.heaps == nullptr;
    h.rec_cap = parked ? pk[0].rec_cap : 0u;
    const uint32_t blocks = (2u * cn + (parked ? 4u * cn : 0u) + HEAP_BLOCK - 1u) / HEAP_BLOCK;
    if (smem) {
      const size_t bytes = (size_t)top_k * HEAP_BLOCK * sizeof(HeapEntry);
      if (int rc = allow_dynamic_smem((const void*)pe_heap_kernel<true>, bytes)) return rc;
      pe_heap_kernel<true><<<blocks, HEAP_BLOCK, bytes, st>>>(h);
    } else {
      pe_heap_kernel<false><<<blocks, HEAP_BLOCK, 0, st>>>(h);
    }
    WALT_CUDA_TRY(cudaGetLastError());
    e->stats.n_kernel_launches++;
  }
  const uint64_t* d_offs1 = m1.d_offs; const uint64_t* d_offs2 = m2.d_offs;
  const uint32_t ulen1 = m1.uniform_len, ulen2 = m2.uniform_len;
  PairArgs a;
  a.cv = chrom_view(e);
  a.r1 = ps.r1; a.n1 = ps.n1; a.offs1 = d_offs1; a.ulen1 = ulen1;
  a.r2 = ps.r2; a.n2 = ps.n2; a.offs2 = d_offs2; a.ulen2 = ulen2;
  a.n = cn; a.top_k = top_k; a.max_mismatches = m; a.frag_range = frag_range; a.swap = swap ? 1u : 0u;
  a.pairs = want_pairs ? ps.pairs : nullptr; a.compact = d_compact;
  a.wide = (e->pair_wide && m <= 127u) ? 1u : 0u;
  pair_kernel<<<(cn + PAIR_BLOCK - 1u) / PAIR_BLOCK, PAIR_BLOCK, (PAIR_BLOCK / 32u) * 3u * top_k * 4u, st>>>(a);
  WALT_CUDA_TRY(cudaGetLastError());
  e->stats.n_kernel_launches++;
  return WALT_OK;
}

// Host batch, chunked and pipelined.  Either the full ranked lists (ranked1 != NULL) or the
// compact per-pair summary (compact != NULL) travel back.  Under PBAT the caller has already
// exchanged the mates; `swap` makes the pairing kernel hand the per-mate fields back.
static int map_pe_host_loop(walt_engine* e, const char* seqs1, const uint64_t* offs1, const char* seqs2, const uint64_t* offs2,
                            uint32_t n, bool packed, uint32_t m, uint32_t b, uint32_t top_k, int frag_range, int swap,
                            walt_cand* ranked1, uint32_t* n_ranked1, walt_cand* ranked2, uint32_t* n_ranked2, walt_pair* pairs,
                            walt_pe_result* compact, uint32_t* n_short1, uint32_t* n_short2);
static int map_pe_host(walt_engine* e, const char* seqs1, const uint64_t* offs1, const char* seqs2, const uint64_t* offs2,
                       uint32_t n, bool packed, uint32_t m, uint32_t b, uint32_t top_k, int frag_range, int swap,
                       walt_cand* ranked1, uint32_t* n_ranked1, walt_cand* ranked2, uint32_t* n_ranked2, walt_pair* pairs,
                       walt_pe_result* compact, uint32_t* n_short1, uint32_t* n_short2) {
  const int rc = map_pe_host_loop(e, seqs1, offs1, seqs2, offs2, n, packed, m, b, top_k, frag_range, swap, ranked1, n_ranked1,
                                  ranked2, n_ranked2, pairs, compact, n_short1, n_short2);
  if (rc) { const std::string msg = g_error; drain_slots(e); g_error = msg; }   // see map_se_host
  return rc;
}

static int map_pe_host_loop(walt_engine* e, const char* seqs1, const uint64_t* offs1, const char* seqs2, const uint64_t* offs2,
                            uint32_t n, bool packed, uint32_t m, uint32_t b, uint32_t top_k, int frag_range, int swap,
                            walt_cand* ranked1, uint32_t* n_ranked1, walt_cand* ranked2, uint32_t* n_ranked2, walt_pair* pairs,
                            walt_pe_result* compact, uint32_t* n_short1, uint32_t* n_short2) {
  int rc;
  e->stats = walt_stats{};
  WALT_CUDA_TRY(cudaMemsetAsync(e->d_counters, 0, N_COUNTERS * 8, e->slot[0].stream));
  WALT_CUDA_TRY(cudaStreamSynchronize(e->slot[0].stream));
  // chunk so that a slot's scratch (ranked lists, and the candidate logs of the two-phase form)
  // stays below ~1.5 GiB
  const bool two_phase = pe_two_phase(e, m);
  const bool gheaps = two_phase && !heaps_in_smem(e, top_k);
  const uint32_t chunk = std::max<uint32_t>(1024u, (uint32_t)std::min<uint64_t>(e->chunk_reads ? e->chunk_reads : (1u << 18),
                                                                                (3ull << 29) / pe_pair_bytes(top_k, m, two_phase, gheaps)));
  uint32_t k = 0, short1 = 0, short2 = 0;
  AheadScans scans(offs1, offs2, n, chunk);
  for (uint32_t r0 = 0; r0 < n; r0 += chunk, ++k) {
    const uint32_t cn = std::min<uint32_t>(chunk, n - r0);
    ChunkScan s1, s2;
    scans.get(k, &s1, &s2);
    if (s1.max_len > MAX_READ_LEN || s2.max_len > MAX_READ_LEN) {
      for (auto& s : e->slot) cudaStreamSynchronize(s.stream);
      return fail(WALT_EINVAL, "read longer than 1024 bases");
    }
    short1 += s1.n_short; short2 += s2.n_short;
    BatchSlot& s = e->slot[k % e->pe_slots];
    WALT_CUDA_TRY(cudaEventSynchronize(s.done));
    uint64_t sb1, se1, sb2, se2;
    chunk_bytes(offs1, r0, cn, packed, &sb1, &se1);
    chunk_bytes(offs2, r0, cn, packed, &sb2, &se2);
    if ((rc = reserve(&s.d_seqs, &s.seqs_cap, (size_t)(se1 - sb1) + 16u))) return rc;
    if ((rc = reserve(&s.d_seqs2, &s.seqs2_cap, (size_t)(se2 - sb2) + 16u))) return rc;
    if ((rc = reserve_bytes(&s.d_pe, &s.pe_cap, pe_scratch_bytes(cn, top_k, m, two_phase, gheaps)))) return rc;
    const PeScratch ps = carve_pe(s.d_pe, cn, top_k, m, two_phase, gheaps);
    if (se1 > sb1) WALT_CUDA_TRY(cudaMemcpyAsync(s.d_seqs, seqs1 + sb1, se1 - sb1, cudaMemcpyHostToDevice, s.stream));
    if (se2 > sb2) WALT_CUDA_TRY(cudaMemcpyAsync(s.d_seqs2, seqs2 + sb2, se2 - sb2, cudaMemcpyHostToDevice, s.stream));
    if (!s1.uniform_len) {
      if ((rc = reserve(&s.d_offs, &s.offs_cap, (size_t)cn + 1u))) return rc;
      WALT_CUDA_TRY(cudaMemcpyAsync(s.d_offs, offs1 + r0, ((size_t)cn + 1u) * 8u, cudaMemcpyHostToDevice, s.stream));
    }
    if (!s2.uniform_len) {
      if ((rc = reserve(&s.d_offs2, &s.offs2_cap, (size_t)cn + 1u))) return rc;
      WALT_CUDA_TRY(cudaMemcpyAsync(s.d_offs2, offs2 + r0, ((size_t)cn + 1u) * 8u, cudaMemcpyHostToDevice, s.stream));
    }
    uint32_t* q = queue_block(e, QB_PE_SLOT + 2u * (k % e->pe_slots));
    const ReadSrc m1{s.d_seqs, s.d_offs, offs1[r0], s1.uniform_len, 0, s1.max_len, packed};
    const ReadSrc m2{s.d_seqs2, s.d_offs2, offs2[r0], s2.uniform_len, 0, s2.max_len, packed};
    if (e->defer && two_phase)
      for (int i = 0; i < 2; ++i)
        if ((rc = reserve_park(e, &s.park[i], cn))) return rc;
    if ((rc = launch_pe_chunk(e, m1, m2, cn, m, b, top_k, frag_range, swap, ps, pairs != nullptr,
                              compact ? ps.compact : nullptr, q, e->defer && two_phase ? s.park : nullptr, s.stream)))
      return rc;
    if (ranked1) {
      const size_t rk = (size_t)cn * top_k * sizeof(walt_cand);
      WALT_CUDA_TRY(cudaMemcpyAsync(ranked1 + (size_t)r0 * top_k, ps.r1, rk, cudaMemcpyDeviceToHost, s.stream));
      WALT_CUDA_TRY(cudaMemcpyAsync(ranked2 + (size_t)r0 * top_k, ps.r2, rk, cudaMemcpyDeviceToHost, s.stream));
      WALT_CUDA_TRY(cudaMemcpyAsync(n_ranked1 + r0, ps.n1, (size_t)cn * 4u, cudaMemcpyDeviceToHost, s.stream));
      WALT_CUDA_TRY(cudaMemcpyAsync(n_ranked2 + r0, ps.n2, (size_t)cn * 4u, cudaMemcpyDeviceToHost, s.stream));
    }
    if (pairs) WALT_CUDA_TRY(cudaMemcpyAsync(pairs + r0, ps.pairs, (size_t)cn * sizeof(walt_pair), cudaMemcpyDeviceToHost, s.stream));
    if (compact)
      WALT_CUDA_TRY(cudaMemcpyAsync(compact + r0, ps.compact, (size_t)cn * sizeof(walt_pe_result), cudaMemcpyDeviceToHost, s.stream));
    WALT_CUDA_TRY(cudaEventRecord(s.done, s.stream));
  }
  if (n_short1) *n_short1 = 2u * short1;   // once per strand pass (paired.cpp:112-115)
  if (n_short2) *n_short2 = 2u * short2;
  for (auto& s : e->slot) WALT_CUDA_TRY(cudaStreamSynchronize(s.stream));
  return fetch_status(e);
}

static int check_pe_args(walt_engine* e, uint32_t top_k) {
  if (top_k < 2 || top_k > 300) return fail(WALT_EINVAL, "-k must be in [2, 300] (walt.cpp:245-249)");
  int rc = ensure_device(e);
  if (rc) return rc;
  if ((rc = check_pair(e, 0)) || (rc = check_pair(e, 1))) return rc;
  return WALT_OK;
}

int walt_engine_map_pe(walt_engine* e, const char* seqs1, const uint64_t* offs1, const char* seqs2,
                       const uint64_t* offs2, uint32_t n, uint32_t max_mismatches, uint32_t b, uint32_t top_k,
                       int frag_range, int pbat, walt_cand* ranked1, uint32_t* n_ranked1, walt_cand* ranked2,
                       uint32_t* n_ranked2, walt_pair* pairs, uint32_t* n_short1, uint32_t* n_short2) {
  if (!e || !offs1 || !offs2 || (n && (!seqs1 || !seqs2 || !ranked1 || !ranked2 || !n_ranked1 || !n_ranked2 || !pairs)))
    return fail(WALT_EINVAL, "bad argument");
  int rc = check_pe_args(e, top_k);
  if (rc) return rc;
  // PBAT swaps the bisulfite roles of the mates: run the directional protocol with the mates
  // exchanged and hand every per-mate result back to its owner.
  if (pbat)
    return map_pe_host(e, seqs2, offs2, seqs1, offs1, n, false, max_mismatches, b, top_k, frag_range, 1, ranked2, n_ranked2,
                       ranked1, n_ranked1, pairs, nullptr, n_short2, n_short1);
  return map_pe_host(e, seqs1, offs1, seqs2, offs2, n, false, max_mismatches, b, top_k, frag_range, 0, ranked1, n_ranked1,
                     ranked2, n_ranked2, pairs, nullptr, n_short1, n_short2);
}

static int map_pe_compact(walt_engine* e, const char* seqs1, const uint64_t* offs1, const char* seqs2, const uint64_t* offs2,
                          uint32_t n, bool packed, uint32_t max_mismatches, uint32_t b, uint32_t top_k, int frag_range,
                          int pbat, walt_pe_result* out, uint32_t* n_short1, uint32_t* n_short2) {
  if (!e || !offs1 || !offs2 || (n && (!seqs1 || !seqs2 || !out))) return fail(WALT_EINVAL, "bad argument");
  int rc = check_pe_args(e, top_k);
  if (rc) return rc;
  if (pbat)
    return map_pe_host(e, seqs2, offs2, seqs1, offs1, n, packed, max_mismatches, b, top_k, frag_range, 1, nullptr, nullptr,
                       nullptr, nullptr, nullptr, out, n_short2, n_short1);
  return map_pe_host(e, seqs1, offs1, seqs2, offs2, n, packed, max_mismatches, b, top_k, frag_range, 0, nullptr, nullptr,
                     nullptr, nullptr, nullptr, out, n_short1, n_short2);
}

int walt_engine_map_pe_compact(walt_engine* e, const char* seqs1, const uint64_t* offs1, const char* seqs2,
                               const uint64_t* offs2, uint32_t n, uint32_t max_mismatches, uint32_t b, uint32_t top_k,
                               int frag_range, int pbat, walt_pe_result* out, uint32_t* n_short1, uint32_t* n_short2) {
  return map_pe_compact(e, seqs1, offs1, seqs2, offs2, n, false, max_mismatches, b, top_k, frag_range, pbat, out, n_short1,
                        n_short2);
}

int walt_engine_map_pe_compact_packed(walt_engine* e, const uint8_t* packed1, const uint64_t* offs1, const uint8_t* packed2,
                                      const uint64_t* offs2, uint32_t n, uint32_t max_mismatches, uint32_t b,
                                      uint32_t top_k, int frag_range, int pbat, walt_pe_result* out, uint32_t* n_short1,
                                      uint32_t* n_short2) {
  return map_pe_compact(e, (const char*)packed1, offs1, (const char*)packed2, offs2, n, true, max_mismatches, b, top_k,
                        frag_range, pbat, out, n_short1, n_short2);
}

int walt_engine_map_pe_device(walt_engine* e, const void* d_seqs1, const void* d_offs1, const void* d_seqs2,
                              const void* d_offs2, uint32_t n, uint32_t max_read_len, uint32_t max_mismatches, uint32_t b,
                              uint32_t top_k, int frag_range, int pbat, void* d_out, void* cuda_stream) {
  if (!e || (n && (!d_seqs1 || !d_offs1 || !d_seqs2 || !d_offs2 || !d_out))) return fail(WALT_EINVAL, "bad argument");
  int rc = check_pe_args(e, top_k);
  if (rc) return rc;
  if (max_read_len > MAX_READ_LEN) return fail(WALT_EINVAL, "read longer than 1024 bases");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  e->stats.n_kernel_launches = 0;
  const char* s1 = (const char*)(pbat ? d_seqs2 : d_seqs1);
  const char* s2 = (const char*)(pbat ? d_seqs1 : d_seqs2);
  const uint64_t* o1 = (const uint64_t*)(pbat ? d_offs2 : d_offs1);
  const uint64_t* o2 = (const uint64_t*)(pbat ? d_offs1 : d_offs2);
  // ranked lists live in engine scratch, one chunk at a time (stream-ordered reuse); the chunk is
  // as large as a quarter of the free HBM allows (1..8 GiB): every launch ends in a tail of a few
  // repeat-heavy reads, so fewer, larger launches are faster
  BatchSlot& s = e->slot[0];
  WALT_CUDA_TRY(cudaStreamSynchronize(s.stream));
  size_t free_b = 0, total_b = 0;
  WALT_CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
  const uint64_t budget = std::max<uint64_t>(s.pe_cap, std::min<uint64_t>(8ull << 30, std::max<uint64_t>(1ull << 30, free_b / 4)));
  const bool two_phase = pe_two_phase(e, max_mismatches);
  const bool gheaps = two_phase && !heaps_in_smem(e, top_k);
  uint32_t chunk = std::max<uint32_t>(1024u, (uint32_t)std::min<uint64_t>(n, budget / pe_pair_bytes(top_k, max_mismatches, two_phase, gheaps)));
  if (n > chunk) {   // equal chunks: no short last launch
    const uint32_t k = (n + chunk - 1u) / chunk;
    chunk = (n + k - 1u) / k;
  }
  if (getenv("WALT_DEBUG"))
    fprintf(stderr, "[walt debug] map_pe_device: free %.2f GB, budget %.2f GB, chunk %u pairs, two_phase %d\n", free_b / 1e9,
            budget / 1e9, chunk, (int)two_phase);
  if ((rc = reserve_bytes(&s.d_pe, &s.pe_cap, pe_scratch_bytes(chunk, top_k, max_mismatches, two_phase, gheaps)))) return rc;
  if (e->defer && two_phase)
    for (int i = 0; i < 2; ++i)
      if ((rc = reserve_park(e, &e->dev_park[i], chunk))) return rc;
  for (uint32_t r0 = 0; r0 < n; r0 += chunk) {
    const uint32_t cn = std::min<uint32_t>(chunk, n - r0);
    const PeScratch ps = carve_pe(s.d_pe, cn, top_k, max_mismatches, two_phase, gheaps);
    // absolute offsets: read r of the chunk is global read r0 + r, addressed from the buffer start
    const ReadSrc m1{s1, o1 + r0, 0, 0, r0, max_read_len, false}, m2{s2, o2 + r0, 0, 0, r0, max_read_len, false};
    if ((rc = launch_pe_chunk(e, m1, m2, cn, max_mismatches, b, top_k, frag_range, pbat, ps, false,
                              (walt_pe_result*)d_out + r0, queue_block(e, QB_PE_DEVICE), e->defer && two_phase ? e->dev_park : nullptr,
                              st)))
      return rc;
  }
  return WALT_OK;
}

// ---- pinned host memory ------------------------------------------------------------------------
void* walt_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { set_error("cudaMallocHost failed"); return nullptr; }
  return p;
}
void walt_host_free(void* p) { if (p) cudaFreeHost(p); }

}  // extern "C"
