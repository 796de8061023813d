// tools/gather_probe.cu -- measures what the B200 memory system delivers for random gathers:
// every thread issues ILP independent loads of `bytes` contiguous bytes at hashed addresses of a
// `gb`-GB array.  Prints accesses/s and GB/s for 8-, 32-, 64- and 128-byte accesses.  This is
// the ceiling the mapping kernels' dependent gathers run against (roofline section).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

template <int WORDS, int ILP>
__global__ void gather(const uint64_t* __restrict__ a, uint64_t n_slots, uint64_t per_thread, uint64_t* out) {
  uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t s = mix64(tid * 0x12345ull + 1);
  uint64_t acc = 0;
  for (uint64_t it = 0; it < per_thread; it += ILP) {
    uint64_t idx[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) { s = mix64(s); idx[j] = (s % n_slots) * WORDS; }
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
      if (WORDS == 1) acc += a[idx[j]];
      else if (WORDS == 4) { const ulonglong2* p = (const ulonglong2*)(a + idx[j]); ulonglong2 x = p[0], y = p[1]; acc += x.x + x.y + y.x + y.y; }
      else { const ulonglong2* p = (const ulonglong2*)(a + idx[j]);
#pragma unroll
             for (int w = 0; w < WORDS / 2; ++w) { ulonglong2 x = p[w]; acc += x.x + x.y; } }
    }
  }
  if (acc == 0x1234567) out[0] = acc;
}

template <int WORDS, int ILP>
void run(const uint64_t* d, uint64_t total_words, uint64_t* d_out, int sms, const char* name) {
  const uint64_t n_slots = total_words / WORDS;
  const int threads = 256, blocks = sms * 8;
  const uint64_t per_thread = 2048;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  gather<WORDS, ILP><<<blocks, threads>>>(d, n_slots, 256, d_out);
  cudaEventRecord(e0);
  gather<WORDS, ILP><<<blocks, threads>>>(d, n_slots, per_thread, d_out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  const double acc = (double)blocks * threads * per_thread;
  printf("{\"probe\": \"random_gather\", \"access_bytes\": %d, \"ilp\": %d, \"accesses_per_s\": %.4g, \"useful_gb_per_s\": %.1f, \"ms\": %.2f, \"what\": \"%s\"}\n",
         WORDS * 8, ILP, acc / (ms * 1e-3), acc * WORDS * 8 / (ms * 1e-3) / 1e9, ms, name);
}

int main(int argc, char** argv) {
  const double gb = argc > 1 ? atof(argv[1]) : 16.0;
  const uint64_t total_words = (uint64_t)(gb * 1e9 / 8) & ~(uint64_t)15;
  uint64_t *d = nullptr, *d_out = nullptr;
  if (cudaMalloc(&d, total_words * 8) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMalloc(&d_out, 8);
  cudaMemset(d, 1, total_words * 8);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  run<1, 4>(d, total_words, d_out, p.multiProcessorCount, "8 B per access");
  run<1, 8>(d, total_words, d_out, p.multiProcessorCount, "8 B per access");
  run<4, 4>(d, total_words, d_out, p.multiProcessorCount, "32 B (one sector) per access");
  run<8, 4>(d, total_words, d_out, p.multiProcessorCount, "64 B per access");
  run<16, 4>(d, total_words, d_out, p.multiProcessorCount, "128 B (one line) per access");
  return 0;
}
