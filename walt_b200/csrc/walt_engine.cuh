// walt_engine.cuh -- internal declarations shared by walt_engine.cu / walt_builder.cu.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/walt_b200.h"
#include "walt_core.cuh"

namespace waltb200 {

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define WALT_CUDA_TRY(expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return ::waltb200::fail(WALT_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

// one strand sub-index resident in HBM
struct DeviceSubIndex {
  bool loaded = false;
  uint64_t* genome = nullptr;     // packed words incl. front/tail pad
  uint64_t genome_words = 0;
  uint32_t* index = nullptr;      // staging only: positions as loaded/sorted, freed by finalize
  waltcore::Entry* entries = nullptr;   // {position, fingerprint} per slot (+ 64 zero entries)
  uint32_t index_size = 0;
  uint32_t* table = nullptr;
  uint32_t depth = 0;
  uint32_t* taint_bits = nullptr; // one allocation: filter bits | rank | start | pos | len
  uint32_t n_taint = 0;           // tainted positions
  uint32_t n_taint_keys = 0;      // distinct 12-mer keys among them
  size_t taint_words = 0;         // words of the taint allocation
  uint64_t bytes = 0;
  waltcore::SubIndexView view(int which) const;
  void release();
};

constexpr uint32_t N_SLOTS = 8;   // most host chunks in flight (copy in / map / copy out); see walt_engine::se_slots, pe_slots
constexpr uint32_t REC_CAP_MAX = 1u << 21;     // parked reads of one launch that get a record (the rest: MAP_TAKE)
constexpr uint32_t BLOCK_CAP_MAX = 1u << 23;   // 32-candidate verification blocks of one launch (256 M candidates)

// device buffers of a launch that parks reads (ParkView, walt_engine.cu)
struct ParkBuf {
  uint32_t* list = nullptr;   size_t list_cap = 0;
  uint32_t* recs = nullptr;   size_t recs_words = 0;
  uint32_t* desc = nullptr;   size_t desc_words = 0;
  uint8_t* bytes = nullptr;   size_t bytes_cap = 0;
  uint32_t rec_cap = 0, cap_blocks = 0;
  void release() {
    cudaFree(list); cudaFree(recs); cudaFree(desc); cudaFree(bytes);
    *this = ParkBuf();
  }
};

// one in-flight chunk of a host batch
struct BatchSlot {
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  char* d_seqs = nullptr;      size_t seqs_cap = 0;
  uint64_t* d_offs = nullptr;  size_t offs_cap = 0;   // entries
  void* d_out = nullptr;       size_t out_cap = 0;    // bytes
  // paired-end: second mate + heaps' drained lists
  char* d_seqs2 = nullptr;     size_t seqs2_cap = 0;
  uint64_t* d_offs2 = nullptr; size_t offs2_cap = 0;
  void* d_pe = nullptr;        size_t pe_cap = 0;     // bytes
  ParkBuf park[2];             // one per mate
};

}  // namespace waltb200

struct walt_engine {
  int device = 0;
  int sm_count = 0;
  uint32_t n_chr = 0;
  uint32_t genome_len = 0;
  std::vector<uint32_t> lengths, starts;
  std::vector<std::string> names;
  std::vector<const char*> name_ptrs;
  uint32_t* d_starts = nullptr;
  waltb200::DeviceSubIndex sub[4];
  waltcore::Pow3 pow3;
  int search_mode = 0;
  int tie_order = 0;          // 0: std::sort's order among equal suffixes, 1: ascending position
  unsigned long long last_build_ties = 0, last_build_replayed = 0;
  int force_depth = 0;
  int min_blocks = 4;         // resident CTAs per SM the SE kernel is compiled for (register cap)
  uint32_t group_width = 8;   // lanes that own one read (8, 16 or 32)
  uint32_t chunk_reads = 0;   // reads per in-flight host chunk; 0 = automatic
  waltb200::BatchSlot slot[waltb200::N_SLOTS];
  cudaStream_t side_stream = nullptr;        // second mate kernel of a paired-end chunk
  cudaEvent_t fork = nullptr, join = nullptr;
  uint32_t se_slots = 6, pe_slots = 3;       // chunks in flight of the single-end / paired-end host batches: the kernels behind
                                             // a parking kernel wait for SM room, so a chunk's results leave late -- more chunks
                                             // in flight keep the copy engines busy meanwhile
  uint32_t chunk_share = 1;                  // chunk kernels of a host batch that share the SMs (see grid_for)
  int pe_logged = 1;                         // 1: two-phase paired-end (candidate log + per-thread heap replay)
  int pe_side = 1;                           // 0: both mate kernels on the caller's stream
  int defer = 1;                             // 1: reads that need their whole group (repeats) are parked by the mapping
                                             // kernels and finished by a warp-per-read kernel
  waltb200::ParkBuf dev_park[2];             // ... of the device-resident calls
  int flat_verify = 1;                       // 1: parked reads get records, their runs are verified by verify_kernel (0: MAP_TAKE maps them all)
  int pair_wide = 1;                         // 1: pairs with long lists are paired by a whole warp
  int heap_smem = 1;                         // 1: pe_heap_kernel keeps its heaps in shared memory when they fit
  uint32_t* d_flags = nullptr;               // [0] non-ACGT flag, [2..3] index-build scratch, from [16] the queue
                                             // blocks of the launches in flight (queue_block, walt_engine.cu)
  int verify_cg = 1;                         // verify_kernel's window copies bypass the L1 (cp.async.cg; measured faster than .ca)
  int time_kernels = 0;                      // 1: CUDA events around every verify_kernel launch (walt_stats.verify_ns)
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> verify_events;
  unsigned long long* d_counters = nullptr;  // lookups, candidates, literal, parked, verified slots
  walt_stats stats{};
};

namespace waltb200 {
waltcore::ChromView chrom_view(const walt_engine* e);
int ensure_device(walt_engine* e);
// after genome+index are resident: choose depth, build table and taint list
int finalize_subindex(walt_engine* e, int which);
int alloc_packed_genome(walt_engine* e, DeviceSubIndex& s);
uint32_t choose_depth(uint32_t index_size);
// 32 ASCII bases per packed word; `forbidden` = 2-bit code that must not occur (4 = none)
int pack_ascii_device(const uint8_t* d_ascii, uint64_t n, uint64_t* d_words, uint32_t forbidden, uint32_t* d_bad);
}  // namespace waltb200
