// micro-benchmark: what does the particle LAYOUT cost a G2P-shaped pass (gather a row through src_of, write it to its slot)?
//   A: struct of arrays by 4-byte word (the layout of svb_device.cuh): 30 gathered LDG.32 + 34 STG.32 per particle
//   B: the same words packed in groups of four (array of float4 per group): 6 gathered LDG.128 + 9 STG.128 per particle
// src_of is "nearly sorted" like the real inverse map: tiles of 512 rows land in another order, rows are shuffled inside 64-row
// neighbourhoods, run starts are not 32-aligned.  The per-thread arithmetic is a short dependent chain on the loaded words, so that
// the loads stay ahead of their uses like in k_g2p.  Both kernels run at 5 CTAs of 128 threads per SM (G2P's residency; 44 KB of dynamic
// shared memory per CTA enforce it) and unrestricted.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o profiles/bin/gather_bench profiles/gather_bench.cu && profiles/bin/gather_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int NW = 34;       // words per particle (layout A)
constexpr int NG = 9;        // float4 groups per particle (layout B: 36 words)
constexpr int READ_W = 30;   // words G2P reads (x, F, the carried words; not v, C)
constexpr int READ_G = 6;    // groups G2P reads with the words ordered by use

__global__ void k_make_src(uint32_t* src_of, uint32_t n, uint32_t n_tiles) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t tile = i >> 9, in_tile = i & 511;
  const uint32_t t2 = (uint32_t)(((unsigned long long)tile * 7919ull) % n_tiles);
  const uint32_t r = (in_tile & ~63u) | ((in_tile * 37u + 11u) & 63u);
  src_of[i] = t2 * 512 + r;
}

template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_words(const float* __restrict__ src, float* __restrict__ dst, const uint32_t* __restrict__ src_of, size_t cap, uint32_t n, uint32_t first) {
  for (uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t si = src_of[i];
    float w[READ_W];
#pragma unroll
    for (int q = 0; q < READ_W; ++q) w[q] = src[(size_t)q * cap + si];
    float a = w[0], b = w[1], c = w[2];
#pragma unroll
    for (int q = 3; q < 12; ++q) { a = fmaf(a, w[q], b); b = fmaf(b, w[q], c); c = fmaf(c, w[q], a); }
#pragma unroll
    for (int q = 0; q < READ_W; ++q) dst[(size_t)q * cap + i] = q < 3 ? (q == 0 ? a : q == 1 ? b : c) : w[q];
#pragma unroll
    for (int q = READ_W; q < NW; ++q) dst[(size_t)q * cap + i] = a + (float)q;
  }
}

template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_groups(const float4* __restrict__ src, float4* __restrict__ dst, const uint32_t* __restrict__ src_of, size_t cap, uint32_t n, uint32_t first) {
  for (uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t si = src_of[i];
    float4 g[READ_G];
#pragma unroll
    for (int q = 0; q < READ_G; ++q) g[q] = src[(size_t)q * cap + si];
    float a = g[0].x, b = g[0].y, c = g[0].z;
    const float ws[9] = {g[0].w, g[1].x, g[1].y, g[1].z, g[1].w, g[2].x, g[2].y, g[2].z, g[2].w};
#pragma unroll
    for (int q = 0; q < 9; ++q) { a = fmaf(a, ws[q], b); b = fmaf(b, ws[q], c); c = fmaf(c, ws[q], a); }
    g[0].x = a; g[0].y = b; g[0].z = c;
#pragma unroll
    for (int q = 0; q < READ_G; ++q) dst[(size_t)q * cap + i] = g[q];
#pragma unroll
    for (int q = READ_G; q < NG; ++q) dst[(size_t)q * cap + i] = make_float4(a + q, b, c, a);
  }
}

template <class F>
float time_ms(F f, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < reps; ++r) f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps;
}

int main() {
  for (uint32_t n : {1024000u, 8192000u}) {
    const uint32_t n_tiles = n / 512;
    const size_t cap = n;
    float *a_src, *a_dst; float4 *b_src, *b_dst; uint32_t* src_of;
    cudaMalloc(&a_src, cap * NW * 4); cudaMalloc(&a_dst, cap * NW * 4);
    cudaMalloc(&b_src, cap * NG * 16); cudaMalloc(&b_dst, cap * NG * 16);
    cudaMalloc(&src_of, (size_t)n * 4);
    cudaMemset(a_src, 0, cap * NW * 4); cudaMemset(b_src, 0, cap * NG * 16);
    k_make_src<<<(n + 255) / 256, 256>>>(src_of, n, n_tiles);
    const uint32_t first = 13;   // runs do not start on a 32-row boundary
    const double bytes_a = (double)n * (READ_W + NW + 1) * 4, bytes_b = (double)n * ((READ_G + NG) * 16 + 4);
    const int reps = n > 2000000 ? 10 : 40;
    float ms;
    ms = time_ms([&] { k_words<5><<<148 * 5, 128, 44 * 1024>>>(a_src, a_dst, src_of, cap, n, first); }, reps);
    printf("n=%u  A words  5 CTA/SM : %8.1f us  %6.0f GB/s (useful bytes)\n", n, ms * 1e3, bytes_a / ms * 1e-6);
    ms = time_ms([&] { k_words<1><<<148 * 16, 128>>>(a_src, a_dst, src_of, cap, n, first); }, reps);
    printf("n=%u  A words  16 CTA/SM: %8.1f us  %6.0f GB/s\n", n, ms * 1e3, bytes_a / ms * 1e-6);
    ms = time_ms([&] { k_groups<5><<<148 * 5, 128, 44 * 1024>>>(b_src, b_dst, src_of, cap, n, first); }, reps);
    printf("n=%u  B groups 5 CTA/SM : %8.1f us  %6.0f GB/s (useful bytes; %.0f with the 2 pad words)\n", n, ms * 1e3, bytes_a / ms * 1e-6, bytes_b / ms * 1e-6);
    ms = time_ms([&] { k_groups<1><<<148 * 16, 128>>>(b_src, b_dst, src_of, cap, n, first); }, reps);
    printf("n=%u  B groups 16 CTA/SM: %8.1f us  %6.0f GB/s\n", n, ms * 1e3, bytes_a / ms * 1e-6);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    cudaFree(a_src); cudaFree(a_dst); cudaFree(b_src); cudaFree(b_dst); cudaFree(src_of);
  }
  return 0;
}
