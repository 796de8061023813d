// walt_builder.cu -- index construction on the device (the reference's makedb) and the
// (the synthetic workload generators of bench.py live in ../bench/walt_synth.cu)
//
// Replaces BuildIndex (src/walt/makedb.cpp:46-85): ReverseComplementGenome + C2T/G2A
// (reference.cpp:131-162), CountBucketSize incl. the ">= 500000" bucket erasure
// (reference.cpp:192-229), HashToBucket (reference.cpp:231-256) and SortHashTableBucket
// (reference.cpp:258-300).  The reference fills 4^12 buckets and std::sorts each with a
// comparator over the genome characters at offsets F2CAREDPOSITION[12..59]; since buckets are
// themselves ordered by the 12 key characters, the whole position array is simply the indexed
// positions ordered by their 60 spaced characters (offsets 1,4,..,178), a character at or
// beyond the chromosome end comparing lowest.  Here that is four stable LSD radix-sort passes
// (cub::DeviceRadixSort) over 15-digit base-4 keys recomputed from the packed genome.  Ties
// (fully equal 60-character keys) come out of that by ascending position, which is not what the
// reference's unstable std::sort leaves behind; buckets that contain ties are therefore
// finished by replaying libstdc++'s std::sort on the class ranks the radix sort established
// (walt_stdsort.cuh), one thread per bucket, so the index is the reference's byte for byte.
#include "walt_engine.cuh"
#include "walt_stdsort.cuh"

#include <cub/cub.cuh>

#include <algorithm>

using namespace waltcore;

namespace waltb200 {

constexpr uint32_t N_SORT_CHARS = 60;     // F2CAREDPOSITION_SIZE (seedpattern.hpp:355-456)
constexpr uint32_t DIGITS_PER_PASS = 15;  // 30-bit keys

// ---- genome derivation ---------------------------------------------------------------------
// dst word w of the sub-index genome `which` from the forward, unconverted packed genome.
__global__ void derive_genome_kernel(const uint64_t* __restrict__ fwd, uint64_t* __restrict__ dst, ChromView cv,
                                     uint32_t which) {
  const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t n_words = ((uint64_t)cv.genome_len + PAD_BASES + 31u) / 32u;
  if (w >= n_words) return;
  const bool rev = (which & 1u) != 0u, ag = which >= 2u;
  uint64_t out = 0;
  uint32_t chr = 0xFFFFFFFFu, cs = 0, ce = 0;
  for (uint32_t i = 0; i < 32u; ++i) {
    const uint64_t pp = w * 32u + i;       // padded position
    uint32_t code = 0;
    if (pp >= PAD_BASES && pp - PAD_BASES < cv.genome_len) {
      const uint32_t p = (uint32_t)(pp - PAD_BASES);
      if (chr == 0xFFFFFFFFu || p >= ce) {
        chr = chrom_of(cv.starts, cv.n_chr, p); cs = cv.starts[chr]; ce = cv.starts[chr + 1u];
      }
      if (rev) code = 3u - packed_base(fwd, (uint64_t)(cs + (ce - 1u - p)) + PAD_BASES);
      else code = packed_base(fwd, pp);
      code = convert_code(code, ag);
    }
    out = (out << 2) | code;
  }
  dst[w] = out;
}

// ---- bucket statistics ---------------------------------------------------------------------
__device__ __forceinline__ bool is_indexed(const ChromView& cv, uint32_t p, uint32_t* chrom_end) {
  const uint32_t chr = chrom_of(cv.starts, cv.n_chr, p);
  const uint32_t cs = cv.starts[chr], ce = cv.starts[chr + 1u];
  *chrom_end = ce;
  return (ce - cs >= MIN_SEED_LEN) && (p < ce - MIN_SEED_LEN);   // reference.cpp:199-204
}

__global__ void key12_hist_kernel(const uint64_t* __restrict__ genome, ChromView cv, Pow3 p3, uint32_t ag,
                                  uint32_t* __restrict__ hist) {
  const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= cv.genome_len) return;
  uint32_t ce;
  if (!is_indexed(cv, (uint32_t)p, &ce)) return;
  atomicAdd(hist + entry_key12(genome, (uint32_t)p, ag != 0u, p3), 1u);
}

__global__ void erase_large_kernel(uint32_t* __restrict__ hist, uint32_t* __restrict__ n_erased) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= N_KEY12) return;
  const bool big = hist[k] >= BUCKET_ERASE;            // reference.cpp:211-218
  if (big) atomicAdd(n_erased, 1u);
  hist[k] = big ? 1u : 0u;                             // becomes the "erased" flag
}

struct KeepPosition {
  const uint64_t* genome; ChromView cv; Pow3 p3; uint32_t ag; const uint32_t* erased;
  __device__ bool operator()(uint32_t p) const {
    uint32_t ce;
    if (!is_indexed(cv, p, &ce)) return false;
    return erased[entry_key12(genome, p, ag != 0u, p3)] == 0u;
  }
};

// ---- sort keys -----------------------------------------------------------------------------
// 15 base-4 digits (first digit most significant) for sort characters [d0, d0+15): 0 = at or
// beyond the chromosome end, else 1 + rank of the converted letter.
__global__ void sort_keys_kernel(const uint64_t* __restrict__ genome, ChromView cv, uint32_t ag,
                                 const uint32_t* __restrict__ pos, uint64_t n, uint32_t d0, uint32_t* __restrict__ keys) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t e = pos[i];
  const uint32_t chr = chrom_of(cv.starts, cv.n_chr, e);
  const uint32_t avail = cv.starts[chr + 1u] - e;
  uint32_t key = 0;
  uint64_t widx = ~0ull, word = 0;
  for (uint32_t d = d0; d < d0 + DIGITS_PER_PASS; ++d) {
    const uint32_t off = 3u * d + 1u;
    uint32_t digit = 0;
    if (off < avail) {
      const uint64_t pp = (uint64_t)e + PAD_BASES + off;
      if ((pp >> 5) != widx) { widx = pp >> 5; word = genome[widx]; }
      digit = 1u + ternary_digit((uint32_t)(word >> (62u - 2u * (uint32_t)(pp & 31u))) & 3u, ag != 0u);
    }
    key = (key << 2) | digit;
  }
  keys[i] = key;
}

// ---- tie order (walt_stdsort.cuh) ------------------------------------------------------------
// head[i] = 1 when sorted slot i starts a new equivalence class of SortHashTableBucketCMP
// (reference.cpp:258-288): its 60 sort digits differ from slot i-1's.  An inclusive scan of head
// is the class rank.  n_ties counts the slots that continue a class.
__global__ void tie_head_kernel(const uint64_t* __restrict__ genome, ChromView cv, uint32_t ag,
                                const uint32_t* __restrict__ pos, uint64_t n, uint32_t* __restrict__ head,
                                unsigned long long* __restrict__ n_ties) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t differs = 1u;
  if (i > 0) {
    const uint32_t a = pos[i - 1u], b = pos[i];
    const uint32_t avail_a = cv.starts[chrom_of(cv.starts, cv.n_chr, a) + 1u] - a;
    const uint32_t avail_b = cv.starts[chrom_of(cv.starts, cv.n_chr, b) + 1u] - b;
    differs = 0u;
    for (uint32_t d = 0; d < N_SORT_CHARS && !differs; ++d) {
      const uint32_t off = 3u * d + 1u;
      const uint32_t da = off < avail_a ? 1u + ternary_digit(packed_base(genome, (uint64_t)a + PAD_BASES + off), ag != 0u) : 0u;
      const uint32_t db = off < avail_b ? 1u + ternary_digit(packed_base(genome, (uint64_t)b + PAD_BASES + off), ag != 0u) : 0u;
      differs = da != db ? 1u : 0u;
    }
    if (!differs) atomicAdd(n_ties, 1ull);
  }
  head[i] = differs;
}

// counts[k] = size of bucket k (0 if erased) -> an exclusive scan gives the bucket's first slot
__global__ void kept_counts_kernel(const uint32_t* __restrict__ raw_hist, uint32_t* __restrict__ counts) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > N_KEY12) return;
  const uint32_t c = k < N_KEY12 ? raw_hist[k] : 0u;
  counts[k] = c >= BUCKET_ERASE ? 0u : c;
}

// One thread per 12-mer bucket: a bucket whose class ranks are not all distinct becomes the root
// task of a std::sort replay.  (A bucket with distinct ranks has one sorted order: left alone.)
__global__ void tie_seed_kernel(const uint32_t* __restrict__ starts, const uint32_t* __restrict__ cls,
                                waltsort::SortTask* __restrict__ seeds, unsigned long long* __restrict__ n_seeds) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= N_KEY12) return;
  const uint64_t s = starts[k], e = starts[k + 1u];
  if (e - s < 2u) return;
  if ((uint64_t)(cls[e - 1u] - cls[s]) + 1u == e - s) return;
  const unsigned long long at = atomicAdd(n_seeds, 1ull);
  seeds[at] = waltsort::SortTask{(uint32_t)s, (uint32_t)e, waltsort::depth_limit_for(e - s)};
}

// One level of the task tree (walt_stdsort.cuh): one thread per task, children appended to the
// next level's list.  BY_POS = true first brings every tie bucket back to the arrangement
// HashToBucket produced (ascending position, reference.cpp:231-256; positions are distinct, so
// any sort gives it), BY_POS = false is std::sort's exact sequence of swaps with "class rank less"
// standing in for the genome comparator.
template <bool BY_POS>
__global__ void sort_level_kernel(const waltsort::SortTask* __restrict__ in, unsigned long long n_in,
                                  uint32_t* __restrict__ pos, uint32_t* __restrict__ cls,
                                  waltsort::SortTask* __restrict__ out, unsigned long long* __restrict__ n_out) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_in) return;
  const waltsort::PairSeq seq{pos, cls};
  waltsort::SortTask kids[2];
  int k;
  if (BY_POS) k = waltsort::sort_task_step(seq, in[i], waltsort::ByPos(), kids);
  else k = waltsort::sort_task_step(seq, in[i], waltsort::ByClass(), kids);
  if (k) {
    const unsigned long long at = atomicAdd(n_out, (unsigned long long)k);
    for (int j = 0; j < k; ++j) out[at + j] = kids[j];
  }
}

// ---- export helpers --------------------------------------------------------------------------
__global__ void unpack_ascii_kernel(const uint64_t* __restrict__ genome, uint64_t first, uint64_t n, char* __restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = "ACGT"[packed_base(genome, first + i + PAD_BASES)];
}

// reference hash value (util.hpp:175-182) of every index entry -> bucket sizes
__global__ void hash_hist_kernel(const uint64_t* __restrict__ genome, const Entry* __restrict__ entries, uint32_t n,
                                 uint32_t* __restrict__ counter) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t e = entries[i].pos;
  uint32_t h = 0;
  for (uint32_t k = 0; k < KEY_WEIGHT; ++k) h = (h << 2) | packed_base(genome, (uint64_t)e + PAD_BASES + 3u * k + 1u);
  atomicAdd(counter + h + 1u, 1u);   // shifted by one: an inclusive scan then yields bucket starts
}

static uint32_t blocks_for(uint64_t n, uint32_t t) { return (uint32_t)((n + t - 1) / t); }

// Build sub-index `which` of `e` from the forward packed genome (device).
static int build_subindex_device(walt_engine* e, int which, const uint64_t* d_fwd) {
  DeviceSubIndex& s = e->sub[which];
  s.release();
  int rc = alloc_packed_genome(e, s);
  if (rc) return rc;
  const ChromView cv = chrom_view(e);
  const uint32_t T = 256;
  const uint32_t ag = which >= 2 ? 1u : 0u;
  const uint64_t n_words = ((uint64_t)e->genome_len + PAD_BASES + 31u) / 32u;
  derive_genome_kernel<<<blocks_for(n_words, T), T>>>(d_fwd, s.genome, cv, (uint32_t)which);
  WALT_CUDA_TRY(cudaGetLastError());

  // bucket sizes over the reachable 3^12 keys, erase the extremal ones
  uint32_t* d_hist = nullptr;
  WALT_CUDA_TRY(cudaMalloc(&d_hist, (N_KEY12 + 2u) * 4u));
  WALT_CUDA_TRY(cudaMemset(d_hist, 0, (N_KEY12 + 2u) * 4u));
  key12_hist_kernel<<<blocks_for(e->genome_len, T), T>>>(s.genome, cv, e->pow3, ag, d_hist);
  uint32_t* d_starts = nullptr;   // bucket sizes now, first slots later (tie replay)
  WALT_CUDA_TRY(cudaMalloc(&d_starts, (N_KEY12 + 1u) * 4u));
  kept_counts_kernel<<<blocks_for(N_KEY12 + 1u, T), T>>>(d_hist, d_starts);
  erase_large_kernel<<<blocks_for(N_KEY12, T), T>>>(d_hist, d_hist + N_KEY12 + 1u);
  WALT_CUDA_TRY(cudaGetLastError());

  // kept positions in ascending order (HashToBucket visits them in this order)
  uint64_t cap = 0;
  for (uint32_t c = 0; c < e->n_chr; ++c)
    if (e->lengths[c] >= MIN_SEED_LEN) cap += e->lengths[c] - MIN_SEED_LEN;
  uint32_t *vals[2] = {nullptr, nullptr}, *keys[2] = {nullptr, nullptr};
  unsigned long long* d_count = nullptr;
  void* d_temp = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_hist); cudaFree(d_starts); cudaFree(keys[0]); cudaFree(keys[1]); cudaFree(d_count); cudaFree(d_temp);
  };
  for (int i = 0; i < 2; ++i) {
    if (cudaMalloc(&vals[i], (cap + 64u) * 4u) != cudaSuccess || cudaMalloc(&keys[i], (cap + 64u) * 4u) != cudaSuccess) {
      cleanup(); cudaFree(vals[0]); cudaFree(vals[1]);
      return fail(WALT_ECUDA, "out of device memory while building the index");
    }
  }
  cudaMalloc(&d_count, 8);
  KeepPosition keep{s.genome, cv, e->pow3, ag, d_hist};
  uint64_t n = 0;
  {
    // chunks of 2^30 positions keep every CUB offset type comfortable
    const uint64_t CH = 1ull << 30;
    size_t temp_bytes = 0;
    cub::CountingInputIterator<uint32_t> it0(0u);
    cub::DeviceSelect::If(nullptr, temp_bytes, it0, vals[0], d_count, (::cuda::std::int64_t)CH, keep);
    if (cudaMalloc(&d_temp, temp_bytes) != cudaSuccess) { cleanup(); cudaFree(vals[0]); cudaFree(vals[1]); return fail(WALT_ECUDA, "cudaMalloc(select temp)"); }
    for (uint64_t p0 = 0; p0 < e->genome_len; p0 += CH) {
      const uint64_t cnt = std::min<uint64_t>(CH, e->genome_len - p0);
      cub::CountingInputIterator<uint32_t> it((uint32_t)p0);
      cub::DeviceSelect::If(d_temp, temp_bytes, it, vals[0] + n, d_count, (::cuda::std::int64_t)cnt, keep);
      unsigned long long got = 0;
      if (cudaMemcpy(&got, d_count, 8, cudaMemcpyDeviceToHost) != cudaSuccess) { cleanup(); cudaFree(vals[0]); cudaFree(vals[1]); return fail(WALT_ECUDA, "select"); }
      n += got;
    }
    cudaFree(d_temp); d_temp = nullptr;
  }
  // four stable LSD passes, least significant 15 characters first
  {
    cub::DoubleBuffer<uint32_t> dk(keys[0], keys[1]);
    cub::DoubleBuffer<uint32_t> dv(vals[0], vals[1]);
    size_t temp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, dk, dv, (uint64_t)n, 0, 2 * DIGITS_PER_PASS);
    if (cudaMalloc(&d_temp, temp_bytes ? temp_bytes : 1) != cudaSuccess) { cleanup(); cudaFree(vals[0]); cudaFree(vals[1]); return fail(WALT_ECUDA, "cudaMalloc(sort temp)"); }
    for (int pass = (int)(N_SORT_CHARS / DIGITS_PER_PASS) - 1; pass >= 0 && n > 0; --pass) {
      sort_keys_kernel<<<blocks_for(n, T), T>>>(s.genome, cv, ag, dv.Current(), n, (uint32_t)pass * DIGITS_PER_PASS, dk.Current());
      cudaError_t ce = cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, dk, dv, (uint64_t)n, 0, 2 * DIGITS_PER_PASS);
      if (ce != cudaSuccess) { cleanup(); cudaFree(vals[0]); cudaFree(vals[1]); return fail(WALT_ECUDA, std::string("radix sort: ") + cudaGetErrorString(ce)); }
    }
    // tie order: class ranks from the sorted order, then std::sort replayed on the tie buckets
    e->last_build_ties = 0; e->last_build_replayed = 0;
    uint32_t* final_index = dv.Current();
    if (n > 1 && e->tie_order == 0) {
      auto bail = [&](const char* what, cudaError_t ce) {
        cleanup(); cudaFree(vals[0]); cudaFree(vals[1]);
        return fail(WALT_ECUDA, std::string(what) + ": " + cudaGetErrorString(ce));
      };
      uint32_t* cls = dk.Current();          // head flags, then class ranks, in sorted order
      uint32_t* work_pos = dk.Alternate();   // the array the replay permutes
      unsigned long long* d_ctr = nullptr;   // [0] tied slots, [1] seeds, [2] next level's tasks
      if (cudaMalloc(&d_ctr, 32) != cudaSuccess) return bail("cudaMalloc", cudaGetLastError());
      cudaMemset(d_ctr, 0, 32);
      tie_head_kernel<<<blocks_for(n, T), T>>>(s.genome, cv, ag, dv.Current(), n, cls, d_ctr);
      unsigned long long ctr[3] = {0, 0, 0};
      cudaError_t ce = cudaMemcpy(ctr, d_ctr, 24, cudaMemcpyDeviceToHost);
      if (ce != cudaSuccess) { cudaFree(d_ctr); return bail("tie detection", ce); }
      if (ctr[0] > 0) {
        size_t need = 0, need2 = 0;
        cub::DeviceScan::InclusiveSum(nullptr, need, cls, cls, (::cuda::std::int64_t)n);
        cub::DeviceScan::ExclusiveSum(nullptr, need2, d_starts, d_starts, (int)(N_KEY12 + 1u));
        need = std::max(need, need2);
        if (need > temp_bytes) {
          cudaFree(d_temp); d_temp = nullptr; temp_bytes = need;
          if (cudaMalloc(&d_temp, temp_bytes) != cudaSuccess) { cudaFree(d_ctr); return bail("cudaMalloc(scan temp)", cudaGetLastError()); }
        }
        size_t tb = temp_bytes;
        cub::DeviceScan::InclusiveSum(d_temp, tb, cls, cls, (::cuda::std::int64_t)n);
        tb = temp_bytes;
        cub::DeviceScan::ExclusiveSum(d_temp, tb, d_starts, d_starts, (int)(N_KEY12 + 1u));
        waltsort::SortTask* seeds = nullptr;
        if (cudaMalloc(&seeds, (size_t)N_KEY12 * sizeof(waltsort::SortTask)) != cudaSuccess) { cudaFree(d_ctr); return bail("cudaMalloc(seeds)", cudaGetLastError()); }
        tie_seed_kernel<<<blocks_for(N_KEY12, T), T>>>(d_starts, cls, seeds, d_ctr + 1);
        cudaMemcpy(ctr, d_ctr, 24, cudaMemcpyDeviceToHost);
        // the replay permutes a copy; buckets without ties keep what the radix sort left
        cudaMemcpyAsync(work_pos, dv.Current(), (size_t)n * 4u, cudaMemcpyDeviceToDevice);
        // two task lists inside the free value buffer: a level holds at most n / 17 tasks
        waltsort::SortTask* lists[2];
        lists[0] = (waltsort::SortTask*)dv.Alternate();
        lists[1] = lists[0] + ((size_t)n / 17u + 2u);
        auto run = [&](bool by_pos) -> cudaError_t {
          const waltsort::SortTask* in = seeds;
          unsigned long long n_in = ctr[1];
          int flip = 0;
          while (n_in > 0) {
            cudaMemsetAsync(d_ctr + 2, 0, 8);
            const uint32_t TB = 64;
            if (by_pos) sort_level_kernel<true><<<blocks_for(n_in, TB), TB>>>(in, n_in, work_pos, cls, lists[flip], d_ctr + 2);
            else sort_level_kernel<false><<<blocks_for(n_in, TB), TB>>>(in, n_in, work_pos, cls, lists[flip], d_ctr + 2);
            unsigned long long n_out = 0;
            const cudaError_t rc2 = cudaMemcpy(&n_out, d_ctr + 2, 8, cudaMemcpyDeviceToHost);
            if (rc2 != cudaSuccess) return rc2;
            in = lists[flip]; n_in = n_out; flip ^= 1;
          }
          return cudaSuccess;
        };
        ce = run(true);
        if (ce == cudaSuccess) ce = run(false);
        cudaFree(seeds);
        if (ce != cudaSuccess) { cudaFree(d_ctr); return bail("tie replay", ce); }
        final_index = work_pos;
      }
      cudaFree(d_ctr);
      e->last_build_ties = ctr[0]; e->last_build_replayed = ctr[1];
    }
    cudaError_t ce = cudaDeviceSynchronize();
    if (ce != cudaSuccess) { cleanup(); cudaFree(vals[0]); cudaFree(vals[1]); return fail(WALT_ECUDA, std::string("index build: ") + cudaGetErrorString(ce)); }
    // keep the buffer that holds the finished index, free the other three
    s.index = final_index;
    for (int i = 0; i < 2; ++i) {
      if (vals[i] != final_index) cudaFree(vals[i]);
      if (keys[i] == final_index) keys[i] = nullptr;
    }
  }
  cleanup();
  s.index_size = (uint32_t)n;
  WALT_CUDA_TRY(cudaMemset(s.index + n, 0, 64u * 4u));
  return finalize_subindex(e, which);
}

}  // namespace waltb200

using namespace waltb200;

extern "C" {

int walt_engine_build_from_device_genome(walt_engine* e, const void* d_packed_genome, uint32_t which_mask) {
  if (!e || !d_packed_genome) return fail(WALT_EINVAL, "bad argument");
  if (!e->d_starts) return fail(WALT_EINVAL, "set the chromosome table first");
  int rc = ensure_device(e);
  if (rc) return rc;
  for (int which = 0; which < 4; ++which) {
    if (!((which_mask >> which) & 1u)) continue;
    if ((rc = build_subindex_device(e, which, (const uint64_t*)d_packed_genome))) return rc;
  }
  return WALT_OK;
}

uint64_t walt_packed_genome_bytes(uint64_t n_bases) { return ((n_bases + PAD_BASES + 31u) / 32u + TAIL_PAD_WORDS) * 8u; }

// ASCII upper-case ACGT genome (host) -> packed forward genome (device buffer of
// walt_packed_genome_bytes(n_bases) bytes)
int walt_pack_genome_device(int device, const char* sequence, uint64_t n_bases, void* d_packed_out);

int walt_engine_build_from_sequence(walt_engine* e, const char* sequence, uint32_t which_mask) {
  if (!e || !sequence) return fail(WALT_EINVAL, "bad argument");
  if (!e->d_starts) return fail(WALT_EINVAL, "set the chromosome table first");
  int rc = ensure_device(e);
  if (rc) return rc;
  void* d_fwd = nullptr;
  WALT_CUDA_TRY(cudaMalloc(&d_fwd, walt_packed_genome_bytes(e->genome_len)));
  rc = walt_pack_genome_device(e->device, sequence, e->genome_len, d_fwd);
  if (!rc) rc = walt_engine_build_from_device_genome(e, d_fwd, which_mask);
  cudaFree(d_fwd);
  return rc;
}

int walt_engine_export_subindex(walt_engine* e, int which, char* sequence, uint32_t* counter, uint32_t* index,
                                uint32_t* index_size) {
  if (!e || which < 0 || which > 3) return fail(WALT_EINVAL, "bad argument");
  const DeviceSubIndex& s = e->sub[which];
  if (!s.loaded) return fail(WALT_ENOTLOADED, "sub-index not resident");
  int rc = ensure_device(e);
  if (rc) return rc;
  const uint32_t T = 256;
  if (index_size) *index_size = s.index_size;
  if (sequence) {
    const uint64_t CH = 256u << 20;
    char* d = nullptr;
    WALT_CUDA_TRY(cudaMalloc(&d, CH));
    for (uint64_t off = 0; off < e->genome_len; off += CH) {
      const uint64_t n = std::min<uint64_t>(CH, e->genome_len - off);
      unpack_ascii_kernel<<<blocks_for(n, T), T>>>(s.genome, off, n, d);
      cudaError_t ce = cudaMemcpy(sequence + off, d, n, cudaMemcpyDeviceToHost);
      if (ce != cudaSuccess) { cudaFree(d); return fail(WALT_ECUDA, cudaGetErrorString(ce)); }
    }
    cudaFree(d);
  }
  if (counter) {
    const size_t NC = (1u << 24) + 1u;
    uint32_t* d = nullptr;
    WALT_CUDA_TRY(cudaMalloc(&d, NC * 4u));
    WALT_CUDA_TRY(cudaMemset(d, 0, NC * 4u));
    if (s.index_size) hash_hist_kernel<<<blocks_for(s.index_size, T), T>>>(s.genome, s.entries, s.index_size, d);
    size_t tb = 0;
    void* tmp = nullptr;
    cub::DeviceScan::InclusiveSum(nullptr, tb, d, d, (int)NC);
    WALT_CUDA_TRY(cudaMalloc(&tmp, tb));
    cub::DeviceScan::InclusiveSum(tmp, tb, d, d, (int)NC);
    cudaError_t ce = cudaMemcpy(counter, d, NC * 4u, cudaMemcpyDeviceToHost);
    cudaFree(tmp); cudaFree(d);
    if (ce != cudaSuccess) return fail(WALT_ECUDA, cudaGetErrorString(ce));
  }
  if (index && s.index_size)   // strided copy: the position half of every {position, fingerprint} entry
    WALT_CUDA_TRY(cudaMemcpy2D(index, 4, s.entries, sizeof(Entry), 4, s.index_size, cudaMemcpyDeviceToHost));
  return WALT_OK;
}

// pack kernel lives in walt_engine.cu (forbidden_code 4 = accept all four letters)

int walt_pack_genome_device(int device, const char* sequence, uint64_t n_bases, void* d_packed_out) {
  if (!sequence || !d_packed_out) return fail(WALT_EINVAL, "bad argument");
  WALT_CUDA_TRY(cudaSetDevice(device));
  WALT_CUDA_TRY(cudaMemset(d_packed_out, 0, walt_packed_genome_bytes(n_bases)));
  const uint64_t CH = 64u << 20;
  uint8_t* d_stage = nullptr;
  uint32_t* d_bad = nullptr;
  WALT_CUDA_TRY(cudaMalloc(&d_stage, CH));
  WALT_CUDA_TRY(cudaMalloc(&d_bad, 4));
  WALT_CUDA_TRY(cudaMemset(d_bad, 0, 4));
  int rc = WALT_OK;
  for (uint64_t off = 0; off < n_bases && !rc; off += CH) {
    const uint64_t n = std::min<uint64_t>(CH, n_bases - off);
    if (cudaMemcpy(d_stage, sequence + off, n, cudaMemcpyHostToDevice) != cudaSuccess) { rc = fail(WALT_ECUDA, "genome upload"); break; }
    rc = pack_ascii_device(d_stage, n, (uint64_t*)d_packed_out + ((off + PAD_BASES) >> 5), 4u, d_bad);
  }
  uint32_t bad = 0;
  cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost);
  cudaFree(d_stage); cudaFree(d_bad);
  if (!rc && bad) rc = fail(WALT_ENONACGT, std::to_string(bad) + " genome bytes are not A/C/G/T");
  return rc;
}

}  // extern "C"
