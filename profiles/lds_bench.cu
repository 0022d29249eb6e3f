// micro-benchmark: cost of broadcast shared-memory loads by width (is a uniform LDS.128 four wavefronts?)
#include <cstdio>
#include <cuda_runtime.h>
template <int W, int MODE>  // W = floats per load (1,2,4); MODE 0 = all lanes same address, 1 = 3 distinct addresses, 2 = lane-strided conflict-free
__global__ void k(float* out, int iters, long long* cycles) {
  __shared__ __align__(16) float sm[32 * 36 * 4];
  for (int i = threadIdx.x; i < 32 * 36 * 4; i += blockDim.x) sm[i] = (float)i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  int off = MODE == 0 ? 0 : MODE == 1 ? (lane % 3) * 8 : lane * 36;
  float acc = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float* p = sm + off + u * 36 * 4 % (32*36*3);
      if (W == 1) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(p))); acc += v; }
      if (W == 2) { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p))); acc += v.x + v.y; }
      if (W == 4) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((unsigned)__cvta_generic_to_shared(p))); acc += v.x + v.y + v.z + v.w; }
    }
    off = (off + 4) & 1023;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int W, int MODE> void run(const char* name) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  k<W, MODE><<<148, 1024>>>(out, iters, cyc); cudaDeviceSynchronize();
  k<W, MODE><<<148, 1024>>>(out, iters, cyc); cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  // 32 warps per SM, each 8*iters loads
  printf("%-28s cycles per warp-load (SM-wide): %.2f\n", name, (double)c / (32.0 * 8 * iters));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<1, 0>("LDS.32 uniform"); run<2, 0>("LDS.64 uniform"); run<4, 0>("LDS.128 uniform");
  run<1, 1>("LDS.32 3 addresses"); run<2, 1>("LDS.64 3 addresses"); run<4, 1>("LDS.128 3 addresses");
  run<1, 2>("LDS.32 strided"); run<2, 2>("LDS.64 strided"); run<4, 2>("LDS.128 strided");
  return 0;
}
