// svb_sort.cuh — radix sort entry points used by the re-bin (particles: 64-bit bin key + 32-bit
// index) and by block activation (64-bit candidate tile keys).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cub/device/device_radix_sort.cuh>

namespace svb {

inline size_t sort_temp_bytes(size_t n) {
  size_t a = 0, b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n, 0, 64);
  cub::DeviceRadixSort::SortKeys(nullptr, b, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)n, 0, 64);
  return (a > b ? a : b) + 256;
}
// returns 0 or a cudaError_t
inline int sort_pairs_u64(void* tmp, size_t tmp_bytes, const unsigned long long* kin, unsigned long long* kout, const uint32_t* vin, uint32_t* vout, uint32_t n, int end_bit,
                          cudaStream_t s, uint64_t* launches) {
  size_t need = tmp_bytes;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, need, kin, kout, vin, vout, (int)n, 0, end_bit, s);
  if (launches) *launches += (uint64_t)((end_bit + 7) / 8 + 1);
  return (int)e;
}
inline int sort_keys_u64(void* tmp, size_t tmp_bytes, const unsigned long long* kin, unsigned long long* kout, uint32_t n, int end_bit, cudaStream_t s, uint64_t* launches) {
  size_t need = tmp_bytes;
  cudaError_t e = cub::DeviceRadixSort::SortKeys(tmp, need, kin, kout, (int)n, 0, end_bit, s);
  if (launches) *launches += (uint64_t)((end_bit + 7) / 8 + 1);
  return (int)e;
}

}  // namespace svb
