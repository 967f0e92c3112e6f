// microbench.cu -- ceilings for the fused kernel's two memory phases with ITS CTA shape (1 CTA / SM).
//  (1) read phase : W warps per CTA, each lane keeps U independent 128-bit loads in flight over rows of a [M,K] fp32 matrix
//  (2) write phase: 4 warps per CTA issuing TMA stores of 32x32 fp32 tiles (SW128) vs plain st.global.v4 (row-strided / coalesced)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// each CTA owns rows [rows_per_cta*b, +rows_per_cta); per k-block (128 floats) warp w reads its share of rows, U loads in flight per lane
template <int U>
__global__ void __launch_bounds__(1024, 1) read_kernel(const float* A, int M, int K, int rows_per_cta, float* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int row0 = blockIdx.x * rows_per_cta;
  float acc = 0.f;
  const int rpw = rows_per_cta / nw;  // rows per warp
  for (int kb = 0; kb < K / 128; ++kb) {
    for (int r0 = 0; r0 < rpw; r0 += U) {
      float4 x[U];
#pragma unroll
      for (int i = 0; i < U; ++i) {
        int row = row0 + warp * rpw + r0 + i;
        x[i] = (row < M && r0 + i < rpw) ? ldg_stream(reinterpret_cast<const float4*>(A + (size_t)row * K + kb * 128 + lane * 4)) : make_float4(0, 0, 0, 0);
      }
#pragma unroll
      for (int i = 0; i < U; ++i) acc += x[i].x + x[i].y + x[i].z + x[i].w;
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

// pattern 1: each warp streams its own rows contiguously (row-major: all of K for row r, then row r+1)
// pattern 2: the CTA's [rows_per_cta x K] block is one contiguous region; warps read 512 B chunks round-robin
template <int U>
__global__ void __launch_bounds__(1024, 1) read_kernel_pat(const float* A, int M, int K, int rows_per_cta, int pattern, float* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const size_t row0 = (size_t)blockIdx.x * rows_per_cta;
  const int kbs = K / 128;
  float acc = 0.f;
  const int rpw = rows_per_cta / nw;
  const int total = rpw * kbs;  // 512 B chunks per warp
  for (int j0 = 0; j0 < total; j0 += U) {
    float4 x[U];
#pragma unroll
    for (int i = 0; i < U; ++i) {
      int j = j0 + i;
      size_t off;
      if (pattern == 1) { int r = j / kbs, kb = j % kbs; off = (row0 + warp * rpw + r) * (size_t)K + kb * 128; }
      else { size_t c = (size_t)j * nw + warp; off = row0 * (size_t)K + c * 128; }
      bool ok = j < total && (off + 128) <= (size_t)M * K;
      x[i] = ok ? ldg_stream(reinterpret_cast<const float4*>(A + off + lane * 4)) : make_float4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < U; ++i) acc += x[i].x + x[i].y + x[i].z + x[i].w;
  }
  if (acc == 123.456f) out[0] = acc;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// TMA store phase: each of 4 warps stores 32-row x 32-col tiles; CTA owns 128 rows x N columns
__global__ void __launch_bounds__(128, 1) tma_store_kernel(const __grid_constant__ CUtensorMap tmap, int N, int bufs) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* mine = smem + warp * bufs * 4096;
  int n = 0;
  for (int c0 = 0; c0 < N; c0 += 32, ++n) {
    if (n >= bufs) {
      if (lane == 0) { if (bufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); else if (bufs == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); else asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory"); }
      __syncwarp();
    }
    uint8_t* tile = mine + (n % bufs) * 4096;
    for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(tile + lane * 128 + ((j * 16) ^ ((lane & 7) << 4))) = make_float4(c0, j, lane, warp);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tmap), "r"(smem_u32(tile)), "r"(c0), "r"((int)(blockIdx.x * 128 + warp * 32)) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// plain stores: mode 0 = thread-per-row 16B (row-strided, v1 epilogue), mode 1 = warp writes 512 contiguous bytes of one row
__global__ void __launch_bounds__(128, 1) st_kernel(float* Y, int N, int mode) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t row_base = (size_t)blockIdx.x * 128 + warp * 32;
  if (mode == 0) {
    float* yrow = Y + (row_base + lane) * N;
    for (int c = 0; c < N; c += 4) __stcs(reinterpret_cast<float4*>(yrow + c), make_float4(c, lane, warp, 0));
  } else {
    for (int r = 0; r < 32; ++r) {
      float* yrow = Y + (row_base + r) * N;
      for (int c = lane * 4; c < N; c += 128) __stcs(reinterpret_cast<float4*>(yrow + c), make_float4(c, lane, warp, 0));
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <class F>
float time_ms(F f, int reps = 5) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); CK(cudaDeviceSynchronize());
  float best = 1e9;
  for (int i = 0; i < reps; ++i) { cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
  return best;
}

int main() {
  const int M = 16384;
  float *A, *Y, *out, *flush;
  CK(cudaMalloc(&A, (size_t)M * 3072 * 4)); CK(cudaMalloc(&Y, (size_t)M * 3072 * 4)); CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&flush, 256u << 20));
  CK(cudaMemset(A, 0, (size_t)M * 3072 * 4));
  auto fl = [&]() { cudaMemsetAsync(flush, 1, 256u << 20); };
  printf("== read phase (fp32 [16384,K]), GB/s\n");
  for (int K : {768}) {
    for (int ctas : {128}) {
      int rpc = (M + ctas - 1) / ctas; rpc = (rpc + 31) / 32 * 32;
      for (int warps : {8, 16, 32}) {
        double bytes = (double)M * K * 4;
        float t4 = time_ms([&]() { fl(); }, 2);
        (void)t4;
        auto run = [&](int U) {
          cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
          float best = 1e9;
          for (int i = 0; i < 4; ++i) {
            fl(); cudaEventRecord(a);
            if (U == 4) read_kernel<4><<<ctas, warps * 32>>>(A, M, K, rpc, out);
            if (U == 8) read_kernel<8><<<ctas, warps * 32>>>(A, M, K, rpc, out);
            if (U == 16) read_kernel<16><<<ctas, warps * 32>>>(A, M, K, rpc, out);
            cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
          }
          return bytes / best / 1e6;
        };
        for (int U : {4, 8, 16}) {
          if ((rpc / warps) % U != 0 && (rpc / warps) < U) continue;
          printf("K=%d ctas=%d rows/cta=%d warps=%d U=%d inflight/SM=%dKB : %.0f GB/s\n", K, ctas, rpc, warps, U, warps * 32 * U * 16 / 1024, run(U));
        }
      }
    }
  }
  printf("== read phase patterns (16 warps, U=8), us and GB/s incl. launch overhead\n");
  for (int K : {768, 3072}) for (int ctas : {128, 147}) {
    int rpc = ctas == 128 ? 128 : 112;
    double bytes = (double)M * K * 4;
    for (int pat : {0, 1, 2}) {
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); float best = 1e9;
      for (int i = 0; i < 5; ++i) {
        fl(); cudaEventRecord(a);
        if (pat == 0) read_kernel<8><<<ctas, 512>>>(A, M, K, rpc, out); else read_kernel_pat<8><<<ctas, 512>>>(A, M, K, rpc, pat, out);
        cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
      }
      printf("K=%d ctas=%d rows/cta=%d pattern=%d : %.1f us  %.0f GB/s\n", K, ctas, rpc, pat, best * 1e3, bytes / best / 1e6);
    }
  }
  {  // empty-kernel launch overhead with the same event bracket
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); float best = 1e9;
    for (int i = 0; i < 5; ++i) { fl(); cudaEventRecord(a); read_kernel<4><<<128, 512>>>(A, 0, 128, 128, out); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    printf("empty-ish kernel: %.1f us\n", best * 1e3);
  }
  // TMA store
  void* sym = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)sym;
  printf("== write phase, GB/s\n");
  for (int N : {768, 3072}) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M}; cuuint64_t strides[1] = {(cuuint64_t)N * 4}; cuuint32_t box[2] = {32, 32}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, Y, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    double bytes = (double)M * N * 4;
    CK(cudaFuncSetAttribute(tma_store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    for (int bufs : {1, 2, 4}) {
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); float best = 1e9;
      for (int i = 0; i < 4; ++i) { fl(); cudaEventRecord(a); tma_store_kernel<<<128, 128, 4 * bufs * 4096>>>(tm, N, bufs); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
      printf("N=%d TMA store 32x32 tiles bufs=%d : %.0f GB/s (%.1f us)\n", N, bufs, bytes / best / 1e6, best * 1e3);
    }
    for (int mode : {0, 1}) {
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); float best = 1e9;
      for (int i = 0; i < 4; ++i) { fl(); cudaEventRecord(a); st_kernel<<<128, 128>>>(Y, N, mode); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
      printf("N=%d st.global.v4 mode=%d (%s) : %.0f GB/s (%.1f us)\n", N, mode, mode ? "warp-coalesced rows" : "thread-per-row", bytes / best / 1e6, best * 1e3);
    }
  }
  return 0;
}
