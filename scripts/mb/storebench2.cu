// storebench2.cu -- which WRITE PATTERN does HBM like?  Mechanism held fixed (plain STG.128 from registers, one 512-byte
// contiguous piece per warp instruction); what varies is how the pieces of one CTA are laid out in the address space:
//   each CTA owns R consecutive rows of a row-major [M, N] fp32 matrix and walks its N columns in visits of C bytes per row;
//   within a visit the CTA's W warps take (row, 512-byte segment) units round-robin.
//   C = 512: 128 columns per visit (the epilogue's per-step footprint) ... C = 4N: complete rows, one after the other.
// A linear fill (every CTA writes one contiguous slab) is the reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o storebench2 storebench2.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void __launch_bounds__(1024, 1)
pattern_kernel(float* Y, int N, int R, int C, int tiles_per_cta, int evict_first) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
  const int segs_per_visit = C / 512;              // 512-byte segments per row and visit
  const int visits = N * 4 / C;
  const float4 v = make_float4(lane, warp, blockIdx.x, 1.f);
  for (int t = 0; t < tiles_per_cta; ++t) {
    const size_t row0 = ((size_t)t * gridDim.x + blockIdx.x) * R;
    for (int vi = 0; vi < visits; ++vi) {
      const int units = R * segs_per_visit;
      for (int u = warp; u < units; u += W) {
        const int r = u / segs_per_visit, s = u % segs_per_visit;
        float4* dst = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(Y + (row0 + r) * (size_t)N) + (size_t)vi * C + (size_t)s * 512) + lane;
        if (evict_first) __stcs(dst, v); else *dst = v;
      }
    }
  }
}

__global__ void __launch_bounds__(1024, 1) fill_kernel(float4* Y, size_t n4) {
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  const size_t per_cta = n4 / gridDim.x;
  float4* base = Y + (size_t)blockIdx.x * per_cta;
  for (size_t i = threadIdx.x; i < per_cta; i += blockDim.x) base[i] = v;
}
__global__ void __launch_bounds__(1024, 1) fill_strided_kernel(float4* Y, size_t n4) {  // grid-stride: all CTAs sweep the buffer together
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) Y[i] = v;
}

int main() {
  float *Y, *flush;
  const size_t cap = (size_t)148 * 128 * 8 * 3072 * 4;  // 1.86 GB
  CK(cudaMalloc(&Y, cap));
  CK(cudaMalloc(&flush, 256u << 20));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  auto timeit = [&](auto launch, double bytes, const char* what) {
    float best = 1e9;
    for (int i = 0; i < 4; ++i) {
      cudaMemsetAsync(flush, 1, 256u << 20);
      cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    printf("%-90s : %6.0f GB/s  (%.1f us)\n", what, bytes / best / 1e6, best * 1e3);
    return 0;
  };
  char name[256];
  for (int W : {8, 16, 32}) {
    const size_t n4 = (size_t)148 * 112 * 2304 / 4;  // ~ the qkv output of M = 16576
    snprintf(name, sizeof name, "linear fill, contiguous slab per CTA, 152 MB            warps=%2d", W);
    timeit([&] { fill_kernel<<<148, W * 32>>>((float4*)Y, n4); }, (double)n4 * 16, name);
    snprintf(name, sizeof name, "linear fill, grid-stride, 152 MB                        warps=%2d", W);
    timeit([&] { fill_strided_kernel<<<148, W * 32>>>((float4*)Y, n4); }, (double)n4 * 16, name);
    snprintf(name, sizeof name, "linear fill, contiguous slab per CTA, 611 MB            warps=%2d", W);
    timeit([&] { fill_kernel<<<148, W * 32>>>((float4*)Y, n4 * 4); }, (double)n4 * 64, name);
  }
  for (int N : {768, 2304, 3072})
    for (int tiles : {1, 4})
      for (int R : {112, 32, 16})
        for (int W : {8, 16})
          for (int C : {512, 1024, N * 4})
            for (int ef : {0}) {
              if ((N * 4) % C) continue;
              if (tiles == 4 && W == 16 && C == 1024) continue;
              const double bytes = (double)148 * tiles * R * N * 4;
              snprintf(name, sizeof name, "N=%4d tiles/CTA=%d rows/tile=%3d  bytes/row/visit=%5d  warps=%2d evict_first=%d  (%.0f MB)", N, tiles, R, C, W, ef, bytes / 1e6);
              timeit([&] { pattern_kernel<<<148, W * 32>>>(Y, N, R, C, tiles, ef); }, bytes, name);
            }
  CK(cudaGetLastError());
  return 0;
}
