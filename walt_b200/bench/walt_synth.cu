// walt_synth.cu -- BENCH / TEST ONLY: seeded synthetic genomes and reads generated on the device
// (SURVEY.md 8(d) shapes) for bench.py.  Built into libwaltsynth.so, which links the product
// library (libwaltb200.so) for the engine's chromosome table; nothing of the product links this.
#include "../csrc/walt_engine.cuh"
#include "../../include/walt_synth.h"

using namespace waltcore;

namespace waltb200 {

__device__ __forceinline__ uint64_t mix64(uint64_t x) {   // splitmix64 finaliser
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// ---- synthetic workloads ---------------------------------------------------------------------
__global__ void synth_genome_kernel(uint64_t* __restrict__ words, uint64_t n_bases, uint64_t seed) {
  const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t n_words = (n_bases + PAD_BASES + 31u) / 32u;
  if (w >= n_words) return;
  uint64_t v = w == 0 ? 0ull : mix64(seed * 0x100000001B3ull + w);   // PAD_BASES == 32: word 0 is the pad
  const uint64_t end = n_bases + PAD_BASES;
  if ((w + 1u) * 32u > end) {
    const uint32_t keep = (uint32_t)(end - w * 32u);
    v = keep ? (v >> (64u - 2u * keep)) << (64u - 2u * keep) : 0ull;
  }
  words[w] = v;
}

// Repeat-heavy genome: the genome is a sequence of 4096-base tiles; 75 %
// of the tiles start with a copy of one of 2000 repeat families (consensus length 300..4096,
// copy number skewed over ~2.5 decades, 0..3 % per-base divergence per family), the rest of
// the tile and the other tiles are i.i.d. uniform.  Everything is a hash of (seed, tile, position),
// so the kernel is stateless.
__global__ void synth_repeat_genome_kernel(uint64_t* __restrict__ words, uint64_t n_bases, uint64_t seed) {
  const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t n_words = (n_bases + PAD_BASES + 31u) / 32u;
  if (w >= n_words) return;
  uint64_t v = 0;
  if (w > 0) {
    const uint64_t b0 = (w - 1u) * 32u;                 // genome base of the word's first character
    const uint64_t tile = b0 >> 12;
    const uint32_t in_tile = (uint32_t)(b0 & 4095u);
    const uint64_t th = mix64(seed * 0x9E3779B97F4A7C15ull + tile * 2u + 1u);
    const bool rep = (th % 100u) < 75u;
    // family: skewed pick (cube of a uniform) so that a few families own thousands of copies
    const double u = (double)((th >> 20) & 0xFFFFFu) / 1048576.0;
    const uint32_t fam = (uint32_t)(u * u * u * 2000.0);
    const uint64_t fh = mix64(seed ^ (0xC2B2AE3D27D4EB4Full * (fam + 1ull)));
    const uint32_t flen = 300u + (uint32_t)(fh % 3797u);
    const uint32_t div_pm = (uint32_t)((fh >> 32) % 31u);   // divergence in per mille, 0..30
    for (uint32_t i = 0; i < 32u; ++i) {
      const uint32_t q = in_tile + i;
      uint32_t base;
      if (rep && q < flen) {
        base = (uint32_t)(mix64(fh + 0x632BE59BD9B4E019ull * (q + 1ull)) & 3u);
        const uint64_t dh = mix64(th ^ (0xD6E8FEB86659FD93ull * (q + 1ull)));
        if ((dh % 1000u) < div_pm) base = (base + 1u + (uint32_t)((dh >> 20) % 3u)) & 3u;
      } else {
        base = (uint32_t)(mix64(seed * 0x100000001B3ull + b0 + i) & 3u);
      }
      v = (v << 2) | base;
    }
  }
  const uint64_t end = n_bases + PAD_BASES;
  if ((w + 1u) * 32u > end) {
    const uint32_t keep = (uint32_t)(end - w * 32u);
    v = keep ? (v >> (64u - 2u * keep)) << (64u - 2u * keep) : 0ull;
  }
  words[w] = v;
}

// SURVEY 8(d) read model: window uniform over the genome (not crossing a chromosome end),
// strand 50/50, each C kept with p = 0.05 else C->T, k substitutions with k uniform over
// {0,0,0,1,2,3,5,7}, 1 % fully random reads; a_rich = reverse complement of that read.
__global__ void synth_reads_kernel(const uint64_t* __restrict__ fwd, ChromView cv, uint32_t n_reads, uint32_t rl,
                                   uint64_t seed, uint32_t a_rich, char* __restrict__ out) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  uint64_t s = mix64(seed ^ (0xD1B54A32D192ED03ull * (r + 1ull)));
  auto next = [&]() { s = mix64(s); return s; };
  char* dst = out + (size_t)r * rl;
  uint8_t buf[MAX_READ_LEN];
  if (next() % 100u == 0u) {
    for (uint32_t i = 0; i < rl; ++i) buf[i] = (uint8_t)(next() & 3u);
  } else {
    uint32_t p, cs, ce;
    for (;;) {
      p = (uint32_t)(next() % cv.genome_len);
      const uint32_t chr = chrom_of(cv.starts, cv.n_chr, p);
      cs = cv.starts[chr]; ce = cv.starts[chr + 1u];
      if ((uint64_t)p + rl + 1u < ce) break;
      (void)cs;
    }
    const bool minus = (next() & 1u) != 0u;
    for (uint32_t i = 0; i < rl; ++i) {
      uint32_t code = minus ? 3u - packed_base(fwd, (uint64_t)p + (rl - 1u - i) + PAD_BASES)
                            : packed_base(fwd, (uint64_t)p + i + PAD_BASES);
      if (code == 1u && (next() % 100u) < 95u) code = 3u;
      buf[i] = (uint8_t)code;
    }
    const uint32_t ks[8] = {0, 0, 0, 1, 2, 3, 5, 7};
    const uint32_t k = ks[next() & 7u];
    for (uint32_t t = 0; t < k; ++t) {
      const uint32_t at = (uint32_t)(next() % rl);
      buf[at] = (uint8_t)((buf[at] + 1u + (uint32_t)(next() % 3u)) & 3u);
    }
  }
  for (uint32_t i = 0; i < rl; ++i) {
    const uint32_t code = a_rich ? 3u - buf[rl - 1u - i] : buf[i];
    dst[i] = "ACGT"[code];
  }
}

// Directional paired-end library: a fragment of length ~N(300,50) clipped
// to [rl, 1000] is taken from either genome strand and bisulfite-converted (95 % of C -> T); mate 1
// is its first rl bases, mate 2 the reverse complement of its last rl bases; each mate then gets
// k substitutions, k uniform over {0,0,0,1,2,3,5,7}; 1 % of the pairs are random sequence.
__global__ void synth_pairs_kernel(const uint64_t* __restrict__ fwd, ChromView cv, uint32_t n_pairs, uint32_t rl,
                                   uint64_t seed, uint32_t readthrough_pct, char* __restrict__ out1,
                                   char* __restrict__ out2) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_pairs) return;
  uint64_t s = mix64(seed ^ (0x9E3779B97F4A7C15ull * (r + 1ull)));
  auto next = [&]() { s = mix64(s); return s; };
  char* d1 = out1 + (size_t)r * rl;
  char* d2 = out2 + (size_t)r * rl;
  if (next() % 100u == 0u) {
    for (uint32_t i = 0; i < rl; ++i) { d1[i] = "ACGT"[next() & 3u]; d2[i] = "ACGT"[next() & 3u]; }
    return;
  }
  // fragment length: sum of 12 uniforms ~ N(6, 1)
  uint32_t acc = 0;
  for (int i = 0; i < 12; ++i) acc += (uint32_t)(next() & 0xFFFFu);
  const float z = (float)acc / 65536.0f - 6.0f;
  int32_t fl = (int32_t)(300.0f + 50.0f * z);
  fl = fl < (int32_t)rl ? (int32_t)rl : (fl > 1000 ? 1000 : fl);
  // adaptor read-through: the insert is shorter than the read; behind it the loader's clipping
  // leaves 'N's which become random bases (mapping.cpp:101-103) -- here random from the start
  if ((next() % 100u) < readthrough_pct) fl = (int32_t)(rl / 3u + (uint32_t)(next() % (rl - rl / 3u)));
  const uint32_t ins = (uint32_t)fl < rl ? (uint32_t)fl : rl;   // bases of each mate that come from the fragment
  uint32_t p;
  for (;;) {
    p = (uint32_t)(next() % cv.genome_len);
    const uint32_t chr = chrom_of(cv.starts, cv.n_chr, p);
    if ((uint64_t)p + (uint32_t)fl + 1u < cv.starts[chr + 1u]) break;
  }
  const bool minus = (next() & 1u) != 0u;
  // base i of the fragment (5' -> 3' on its own strand), converted
  auto frag_base = [&](uint32_t i) -> uint32_t {
    uint32_t code = minus ? 3u - packed_base(fwd, (uint64_t)p + ((uint32_t)fl - 1u - i) + PAD_BASES)
                          : packed_base(fwd, (uint64_t)p + i + PAD_BASES);
    // conversion decided by a hash of (pair, fragment position) so that overlapping mates agree
    if (code == 1u && (mix64(s ^ (0xA24BAED4963EE407ull * (i + 1ull))) % 100u) < 95u) code = 3u;
    return code;
  };
  const uint32_t ks[8] = {0, 0, 0, 1, 2, 3, 5, 7};
  uint8_t buf[MAX_READ_LEN];
  for (int mate = 0; mate < 2; ++mate) {
    uint64_t t = mix64(s ^ (mate ? 0x1234567ull : 0x7654321ull));
    auto tn = [&]() { t = mix64(t); return t; };
    for (uint32_t i = 0; i < rl; ++i)
      buf[i] = i < ins ? (uint8_t)(mate == 0 ? frag_base(i) : 3u - frag_base((uint32_t)fl - 1u - i)) : (uint8_t)(tn() & 3u);
    const uint32_t k = ks[tn() & 7u];
    for (uint32_t q = 0; q < k; ++q) {
      const uint32_t at = (uint32_t)(tn() % rl);
      buf[at] = (uint8_t)((buf[at] + 1u + (uint32_t)(tn() % 3u)) & 3u);
    }
    char* d = mate ? d2 : d1;
    for (uint32_t i = 0; i < rl; ++i) d[i] = "ACGT"[buf[i]];
  }
}


// Verification-bound workload (bench.py --workload verify): the genome is a sequence of 512-base tiles;
// rep_pct % of them start with a copy of one of n_fam families (consensus of VERIFY_FAM_LEN bases, each
// base of a copy substituted with probability div_pm / 1000), everything else is i.i.d. uniform.  With
// n_tiles * rep_pct / 100 / n_fam ~ 5000 copies per family and little divergence every seed lookup of a
// read taken from a copy meets thousands of candidates.
constexpr uint32_t VERIFY_TILE = 512, VERIFY_FAM_LEN = 400;
__device__ __forceinline__ bool verify_tile_family(uint64_t seed, uint64_t tile, uint32_t n_fam, uint32_t rep_pct, uint32_t& fam) {
  const uint64_t th = mix64(seed * 0x9E3779B97F4A7C15ull + tile * 2u + 1u);
  fam = (uint32_t)((th >> 20) % n_fam);
  return (th % 100u) < rep_pct;
}
__global__ void synth_verify_genome_kernel(uint64_t* __restrict__ words, uint64_t n_bases, uint64_t seed, uint32_t n_fam,
                                           uint32_t rep_pct, uint32_t div_pm) {
  const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t n_words = (n_bases + PAD_BASES + 31u) / 32u;
  if (w >= n_words) return;
  uint64_t v = 0;
  if (w > 0) {
    const uint64_t b0 = (w - 1u) * 32u;
    const uint64_t tile = b0 / VERIFY_TILE;
    const uint32_t in_tile = (uint32_t)(b0 % VERIFY_TILE);
    uint32_t fam;
    const bool rep = verify_tile_family(seed, tile, n_fam, rep_pct, fam);
    const uint64_t fh = mix64(seed ^ (0xC2B2AE3D27D4EB4Full * (fam + 1ull)));
    for (uint32_t i = 0; i < 32u; ++i) {
      const uint32_t q = in_tile + i;
      uint32_t base;
      if (rep && q < VERIFY_FAM_LEN) {
        base = (uint32_t)(mix64(fh + 0x632BE59BD9B4E019ull * (q + 1ull)) & 3u);
        const uint64_t dh = mix64((seed + tile) ^ (0xD6E8FEB86659FD93ull * (q + 1ull)));
        if ((dh % 1000u) < div_pm) base = (base + 1u + (uint32_t)((dh >> 20) % 3u)) & 3u;
      } else {
        base = (uint32_t)(mix64(seed * 0x100000001B3ull + b0 + i) & 3u);
      }
      v = (v << 2) | base;
    }
  }
  const uint64_t end = n_bases + PAD_BASES;
  if ((w + 1u) * 32u > end) {
    const uint32_t keep = (uint32_t)(end - w * 32u);
    v = keep ? (v >> (64u - 2u * keep)) << (64u - 2u * keep) : 0ull;
  }
  words[w] = v;
}

// reads from inside the copies (directional bisulfite model of synth_reads_kernel, k in {0,0,1,2} substitutions)
__global__ void synth_verify_reads_kernel(const uint64_t* __restrict__ fwd, ChromView cv, uint32_t n_reads, uint32_t rl,
                                          uint64_t seed, uint64_t gseed, uint32_t n_fam, uint32_t rep_pct, char* __restrict__ out) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  uint64_t s = mix64(seed ^ (0xD1B54A32D192ED03ull * (r + 1ull)));
  auto next = [&]() { s = mix64(s); return s; };
  char* dst = out + (size_t)r * rl;
  const uint64_t n_tiles = cv.genome_len / VERIFY_TILE;
  uint32_t p = 0;
  for (int tries = 0; tries < 256; ++tries) {
    const uint64_t tile = next() % n_tiles;
    uint32_t fam;
    if (!verify_tile_family(gseed, tile, n_fam, rep_pct, fam)) continue;
    p = (uint32_t)(tile * VERIFY_TILE + next() % (VERIFY_FAM_LEN - rl + 1u));
    const uint32_t chr = chrom_of(cv.starts, cv.n_chr, p);
    if (chrom_of(cv.starts, cv.n_chr, p + rl + 1u) == chr) break;   // not across a chromosome end
  }
  const bool minus = (next() & 1u) != 0u;
  const uint32_t ks[4] = {0, 0, 1, 2};
  const uint32_t k = ks[next() & 3u];
  uint32_t at[2] = {(uint32_t)(next() % rl), (uint32_t)(next() % rl)};
  for (uint32_t i = 0; i < rl; ++i) {
    uint32_t code = minus ? 3u - packed_base(fwd, (uint64_t)p + (rl - 1u - i) + PAD_BASES)
                          : packed_base(fwd, (uint64_t)p + i + PAD_BASES);
    if (code == 1u && (next() % 100u) < 95u) code = 3u;
    for (uint32_t t = 0; t < k; ++t) if (at[t] == i) code = (code + 1u + (uint32_t)(next() % 3u)) & 3u;
    dst[i] = "ACGT"[code];
  }
}

static uint32_t blocks_for(uint64_t n, uint32_t t) { return (uint32_t)((n + t - 1) / t); }

}  // namespace waltb200

using namespace waltb200;

extern "C" {

int walt_synth_genome_device(int device, uint64_t n_bases, uint64_t seed, void* d_packed_out) {
  if (!d_packed_out) return fail(WALT_EINVAL, "bad argument");
  WALT_CUDA_TRY(cudaSetDevice(device));
  WALT_CUDA_TRY(cudaMemset(d_packed_out, 0, walt_packed_genome_bytes(n_bases)));
  const uint64_t n_words = (n_bases + PAD_BASES + 31u) / 32u;
  synth_genome_kernel<<<blocks_for(n_words, 256), 256>>>((uint64_t*)d_packed_out, n_bases, seed);
  WALT_CUDA_TRY(cudaGetLastError());
  WALT_CUDA_TRY(cudaDeviceSynchronize());
  return WALT_OK;
}

int walt_synth_repeat_genome_device(int device, uint64_t n_bases, uint64_t seed, void* d_packed_out) {
  if (!d_packed_out) return fail(WALT_EINVAL, "bad argument");
  WALT_CUDA_TRY(cudaSetDevice(device));
  WALT_CUDA_TRY(cudaMemset(d_packed_out, 0, walt_packed_genome_bytes(n_bases)));
  const uint64_t n_words = (n_bases + PAD_BASES + 31u) / 32u;
  synth_repeat_genome_kernel<<<blocks_for(n_words, 256), 256>>>((uint64_t*)d_packed_out, n_bases, seed);
  WALT_CUDA_TRY(cudaGetLastError());
  WALT_CUDA_TRY(cudaDeviceSynchronize());
  return WALT_OK;
}

int walt_synth_reads_device(walt_engine* e, const void* d_packed_genome, uint32_t n_reads, uint32_t read_len,
                            uint64_t seed, int a_rich, void* d_seqs_out) {
  if (!e || !d_packed_genome || !d_seqs_out || read_len == 0 || read_len > MAX_READ_LEN)
    return fail(WALT_EINVAL, "bad argument");
  if (!e->d_starts) return fail(WALT_EINVAL, "set the chromosome table first");
  int rc = ensure_device(e);
  if (rc) return rc;
  bool fits = false;
  for (uint32_t c = 0; c < e->n_chr; ++c) fits |= e->lengths[c] > read_len + 2u;
  if (!fits) return fail(WALT_EINVAL, "no chromosome is longer than the read length");
  if (n_reads)
    synth_reads_kernel<<<blocks_for(n_reads, 64), 64>>>((const uint64_t*)d_packed_genome, chrom_view(e), n_reads,
                                                        read_len, seed, a_rich ? 1u : 0u, (char*)d_seqs_out);
  WALT_CUDA_TRY(cudaGetLastError());
  WALT_CUDA_TRY(cudaDeviceSynchronize());
  return WALT_OK;
}

int walt_synth_pairs_device(walt_engine* e, const void* d_packed_genome, uint32_t n_pairs, uint32_t read_len,
                            uint64_t seed, uint32_t readthrough_pct, void* d_seqs1_out, void* d_seqs2_out) {
  if (!e || !d_packed_genome || !d_seqs1_out || !d_seqs2_out || read_len == 0 || read_len > MAX_READ_LEN)
    return fail(WALT_EINVAL, "bad argument");
  int rc = ensure_device(e);
  if (rc) return rc;
  if (!e->d_starts) return fail(WALT_EINVAL, "set the chromosome table first");
  if (n_pairs)
    synth_pairs_kernel<<<blocks_for(n_pairs, 64), 64>>>((const uint64_t*)d_packed_genome, chrom_view(e), n_pairs, read_len,
                                                        seed, readthrough_pct, (char*)d_seqs1_out, (char*)d_seqs2_out);
  WALT_CUDA_TRY(cudaGetLastError());
  WALT_CUDA_TRY(cudaDeviceSynchronize());
  return WALT_OK;
}


int walt_synth_verify_genome_device(int device, uint64_t n_bases, uint64_t seed, uint32_t n_families, uint32_t rep_pct,
                                    uint32_t div_per_mille, void* d_packed_out) {
  if (!d_packed_out || !n_families || rep_pct > 100u) return fail(WALT_EINVAL, "bad argument");
  WALT_CUDA_TRY(cudaSetDevice(device));
  WALT_CUDA_TRY(cudaMemset(d_packed_out, 0, walt_packed_genome_bytes(n_bases)));
  const uint64_t n_words = (n_bases + PAD_BASES + 31u) / 32u;
  synth_verify_genome_kernel<<<blocks_for(n_words, 256), 256>>>((uint64_t*)d_packed_out, n_bases, seed, n_families, rep_pct,
                                                               div_per_mille);
  WALT_CUDA_TRY(cudaGetLastError());
  WALT_CUDA_TRY(cudaDeviceSynchronize());
  return WALT_OK;
}

int walt_synth_verify_reads_device(walt_engine* e, const void* d_packed_genome, uint32_t n_reads, uint32_t read_len,
                                   uint64_t seed, uint64_t genome_seed, uint32_t n_families, uint32_t rep_pct, void* d_seqs_out) {
  if (!e || !d_packed_genome || !d_seqs_out || read_len == 0 || read_len > VERIFY_FAM_LEN || !n_families)
    return fail(WALT_EINVAL, "bad argument");
  if (!e->d_starts) return fail(WALT_EINVAL, "set the chromosome table first");
  int rc = ensure_device(e);
  if (rc) return rc;
  if (n_reads)
    synth_verify_reads_kernel<<<blocks_for(n_reads, 64), 64>>>((const uint64_t*)d_packed_genome, chrom_view(e), n_reads, read_len,
                                                               seed, genome_seed, n_families, rep_pct, (char*)d_seqs_out);
  WALT_CUDA_TRY(cudaGetLastError());
  WALT_CUDA_TRY(cudaDeviceSynchronize());
  return WALT_OK;
}

}  // extern "C"
