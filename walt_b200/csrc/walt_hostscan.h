// walt_hostscan.h -- see walt_hostscan.cpp
#pragma once
#include <stdint.h>

namespace waltb200 {

// max_len = 0xFFFFFFFF flags a length that does not fit 32 bits (the caller rejects the batch)
struct ChunkScan { uint32_t max_len, n_short, uniform_len; };
ChunkScan scan_chunk(const uint64_t* offs, uint32_t r0, uint32_t cn);

}  // namespace waltb200
