// storebench.cu -- what the store path of one SM sustains in the shape the fused kernel's epilogue uses:
// E epilogue warps per CTA, each cycling through T staging tiles of 32 rows x 128 B (SWIZZLE_128B) that leave through
// cp.async.bulk.tensor stores (mode 0), through one 128-byte cp.async.bulk per row issued by every lane (mode 1), or
// through plain st.global.v4 from registers (mode 2: thread = row, 8 x 16 B per 32-column group; mode 3: the warp
// transposes through shared memory and writes 512 contiguous bytes per row).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o storebench storebench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// CTA owns `rows` rows (multiple of 32) x N columns; warp w: quarter = w % 4 (32 rows), column slice = w / 4
__global__ void __launch_bounds__(1024, 1)
store_kernel(const __grid_constant__ CUtensorMap tmap, float* Y, int N, int rows, int E, int T, int mode, int do_sts, int hint, int interleave, const __grid_constant__ CUtensorMap tmap3, int J) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= E) return;
  uint64_t policy = 0;
  if (hint == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  if (hint == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
  if (hint == 3) asm volatile("createpolicy.fractional.L2::evict_unchanged.b64 %0, 1.0;" : "=l"(policy));
  const int q = warp & 3, slices = E >> 2, slice = warp >> 2;
  uint8_t* mine = smem + (size_t)warp * T * 4096;
  const int cols_per_slice = N / slices;
  const int row0 = blockIdx.x * rows + q * 32;
  int n = 0;
  for (int rb = 0; rb < rows; rb += 128)
    for (int g = 0; g < cols_per_slice / 32; ++g, ++n) {
      const int c0 = interleave ? (g * slices + slice) * 32 : slice * cols_per_slice + g * 32;
      uint8_t* tile = mine + (n % T) * 4096;
      float4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = make_float4(c0 + j, lane, warp, n);
      if (mode == 4) {
        // one store = 32 rows x J adjacent 32-column groups: shared layout [32 rows][J][128 B], 128B-swizzled per unit
        if ((g % J) == 0) {
          if (n / J >= T) {
            if (lane == 0) { if (T == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
            __syncwarp();
          }
        }
        uint8_t* big = smem + (size_t)warp * T * J * 4096 + (size_t)((n / J) % T) * J * 4096;
        const int jj = g % J;
        const uint32_t unit = (uint32_t)(lane * J + jj);
#pragma unroll
        for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(big + unit * 128 + ((j ^ (unit & 7)) << 4)) = v[j];
        if (jj == J - 1) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(&tmap3), "r"(smem_u32(big)),
                         "r"(0), "r"((c0 >> 5) - (J - 1)), "r"(row0 + rb)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      } else if (mode == 5) {
        // staged rows of J x 128 B, then PLAIN coalesced stores: the warp re-reads one row at a time (J = 4: 512 contiguous bytes,
        // one STG.128 per lane) -- the access pattern of a fill kernel
        uint8_t* big = smem + (size_t)warp * T * J * 4096 + (size_t)((n / J) % T) * J * 4096;
        const int jj = g % J;
        const uint32_t rowb = (uint32_t)J * 128u;
#pragma unroll
        for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(big + lane * rowb + ((((uint32_t)(jj * 8 + j)) ^ ((uint32_t)lane & 7u)) << 4)) = v[j];
        if (jj == J - 1) {
          __syncwarp();
          const int cbase = c0 - (J - 1) * 32;
          if (J == 4) {
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
              const float4 o = *reinterpret_cast<const float4*>(big + r * rowb + ((((uint32_t)lane) ^ ((uint32_t)r & 7u)) << 4));
              float4* dst = reinterpret_cast<float4*>(Y + (size_t)(row0 + rb + r) * N + cbase) + lane;
              if (hint) __stcs(dst, o); else *dst = o;
            }
          } else if (J == 2) {
#pragma unroll 8
            for (int r2 = 0; r2 < 32; r2 += 2) {
              const int r = r2 + (lane >> 4), ch = lane & 15;
              const float4 o = *reinterpret_cast<const float4*>(big + r * rowb + ((((uint32_t)ch) ^ ((uint32_t)r & 7u)) << 4));
              float4* dst = reinterpret_cast<float4*>(Y + (size_t)(row0 + rb + r) * N + cbase) + ch;
              if (hint) __stcs(dst, o); else *dst = o;
            }
          } else if (J == 8) {
#pragma unroll 4
            for (int r = 0; r < 32; ++r) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const uint32_t ch = (uint32_t)(h * 32 + lane);
                const float4 o = *reinterpret_cast<const float4*>(big + r * rowb + ((ch ^ ((uint32_t)r & 7u)) << 4));
                float4* dst = reinterpret_cast<float4*>(Y + (size_t)(row0 + rb + r) * N + cbase) + ch;
                if (hint) __stcs(dst, o); else *dst = o;
              }
            }
          }
          __syncwarp();
        }
      } else if (mode <= 1) {
        if (n >= T) {
          if (lane == 0 || mode == 1) {
            if (T == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            else if (T == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else if (T == 4) asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 7;" ::: "memory");
          }
          __syncwarp();
        }
        if (do_sts) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t off = mode == 0 ? (uint32_t)((j * 16) ^ ((lane & 7) << 4)) : (uint32_t)(j * 16);
            *reinterpret_cast<float4*>(tile + lane * 128 + off) = v[j];
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (mode == 0) {
          if (lane == 0) {
            if (hint == 0)
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tmap), "r"(smem_u32(tile)),
                           "r"(c0), "r"(row0 + rb)
                           : "memory");
            else
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;" ::"l"(&tmap),
                           "r"(smem_u32(tile)), "r"(c0), "r"(row0 + rb), "l"(policy)
                           : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        } else {
          float* g = Y + (size_t)(row0 + rb + lane) * N + c0;
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 128;" ::"l"(g), "r"(smem_u32(tile + lane * 128)) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      } else if (mode == 2) {
        float* g = Y + (size_t)(row0 + rb + lane) * N + c0;
#pragma unroll
        for (int j = 0; j < 8; ++j) __stcs(reinterpret_cast<float4*>(g) + j, v[j]);
      } else {
        // transpose through shared memory: lane writes its row swizzled, then the warp re-reads row by row
#pragma unroll
        for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(tile + lane * 128 + ((j * 16) ^ ((lane & 7) << 4))) = v[j];
        __syncwarp();
#pragma unroll
        for (int r4 = 0; r4 < 32; r4 += 4) {
          const int r = r4 + (lane >> 3), ch = lane & 7;
          const float4 o = *reinterpret_cast<const float4*>(tile + r * 128 + ((ch * 16) ^ ((r & 7) << 4)));
          __stcs(reinterpret_cast<float4*>(Y + (size_t)(row0 + rb + r) * N + c0) + ch, o);
        }
        __syncwarp();
      }
    }
  if ((mode <= 1 || mode == 4) && (lane == 0 || mode == 1)) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}


// mode 6: E producer warps (quarter q = w % 4, slice = w / 4 of 2) write 32 rows x 128 B halves of [32 rows][256 B] staging tiles (two per
// quarter); E drainer warps (q, b) re-read their tile two rows at a time and write it with plain STG.128 (256 contiguous bytes per
// row); full / empty mbarriers per tile.  The store engine is the SM's own LSU instead of the TMA unit.
__device__ __forceinline__ void mb_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mb_arrive(uint64_t* b) { asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t ph) {
  asm volatile("{\n\t.reg .pred p;\n\tW6: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D6;\n\tbra W6;\n\tD6:\n\t}" ::"r"(smem_u32(b)), "r"(ph) : "memory");
}
__global__ void __launch_bounds__(512, 1)
drain_kernel(float* Y, int N, int rows, int hint, int split_rows) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[8], empty[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) { mb_init(&full[i], 2); mb_init(&empty[i], split_rows ? 2 : 1); } }
  __syncthreads();
  const int steps = N / 64;
  if (warp < 8) {
    const int q = warp & 3, slice = warp >> 2;
    uint32_t n = 0;
    for (int rb = 0; rb < rows; rb += 128)
      for (int g = 0; g < steps; ++g, ++n) {
        const int b = n & 1;
        uint8_t* tile = smem + (q * 2 + b) * 8192;
        if (n >= 2) mb_wait(&empty[q * 2 + b], ((n >> 1) - 1) & 1);
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = make_float4(g + j, lane, warp, n);
#pragma unroll
        for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(tile + lane * 256 + slice * 128 + ((j ^ (lane & 7)) << 4)) = v[j];
        __syncwarp();
        if (lane == 0) mb_arrive(&full[q * 2 + b]);
      }
  } else {
    const int d = warp - 8, q = d & 3, b = d >> 2;
    uint32_t n = 0;
    for (int rb = 0; rb < rows; rb += 128)
      for (int g = 0; g < steps; ++g, ++n) {
        const int row0 = blockIdx.x * rows + rb + q * 32;
        if (!split_rows) {
          if ((int)(n & 1) != b) continue;
          const uint8_t* tile = smem + (q * 2 + b) * 8192;
          mb_wait(&full[q * 2 + b], (n >> 1) & 1);
          const int ch = lane & 15;
#pragma unroll 8
          for (int i = 0; i < 16; ++i) {
            const int r = 2 * i + (lane >> 4);
            const float4 o = *reinterpret_cast<const float4*>(tile + r * 256 + (ch >> 3) * 128 + (((ch & 7) ^ (r & 7)) << 4));
            float4* dst = reinterpret_cast<float4*>(Y + (size_t)(row0 + r) * N + g * 64) + ch;
            if (hint) __stcs(dst, o); else *dst = o;
          }
          __syncwarp();
          if (lane == 0) mb_arrive(&empty[q * 2 + b]);
        } else {
          // both drainers of a quarter take half of the rows of EVERY tile
          const int bb = n & 1;
          const uint8_t* tile = smem + (q * 2 + bb) * 8192;
          mb_wait(&full[q * 2 + bb], (n >> 1) & 1);
          const int ch = lane & 15;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = b * 16 + 2 * i + (lane >> 4);
            const float4 o = *reinterpret_cast<const float4*>(tile + r * 256 + (ch >> 3) * 128 + (((ch & 7) ^ (r & 7)) << 4));
            float4* dst = reinterpret_cast<float4*>(Y + (size_t)(row0 + r) * N + g * 64) + ch;
            if (hint) __stcs(dst, o); else *dst = o;
          }
          __syncwarp();
          if (lane == 0) mb_arrive(&empty[q * 2 + bb]);
        }
      }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int M = 148 * 512;  // 148 CTAs x 512 rows: runs of 100+ us, the launch overhead drops out
  float *Y, *flush;
  CK(cudaMalloc(&Y, (size_t)M * 2304 * 4));
  CK(cudaMalloc(&flush, 256u << 20));
  void* sym = nullptr; cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qr));
  EncodeTiledFn enc = (EncodeTiledFn)sym;
  CK(cudaFuncSetAttribute(store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  const int ctas = 148, rows = 512;
  for (int N : {768, 2304}) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M}; cuuint64_t strides[1] = {(cuuint64_t)N * 4}; cuuint32_t box[2] = {32, 32}; cuuint32_t es[2] = {1, 1};
    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, Y, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
    const double bytes = (double)ctas * rows * N * 4;

    for (int hint = 0; hint < 2; ++hint)
      for (int split = 0; split < 2; ++split) {
        CK(cudaFuncSetAttribute(drain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        float best = 1e9;
        for (int i = 0; i < 4; ++i) {
          cudaMemsetAsync(flush, 1, 256u << 20);
          cudaEventRecord(a);
          drain_kernel<<<ctas, 512, 64 * 1024>>>(Y, N, rows, hint, split);
          cudaEventRecord(b);
          CK(cudaEventSynchronize(b));
          float ms; cudaEventElapsedTime(&ms, a, b);
          if (ms < best) best = ms;
        }
        CK(cudaGetLastError());
        printf("N=%4d decoupled 8 producers -> [32 x 256 B] tiles -> 8 drainer warps STG.128  hint=%d split_rows=%d : %6.0f GB/s  %5.1f B/clk/SM @1.9GHz  (%.1f us incl. launch)\n",
               N, hint, split, bytes / best / 1e6, bytes / best / 1e6 / 148 / 1.9, best * 1e3);
      }
    struct Cfg { int mode, E, T, sts, hint, il, J; };
    const Cfg cfgs[] = {{0, 8, 2, 1, 0, 0, 1}, {0, 8, 2, 1, 0, 1, 1}, {0, 4, 4, 1, 0, 0, 1}, {4, 8, 1, 1, 0, 0, 2}, {5, 8, 1, 1, 0, 0, 2}, {5, 8, 1, 1, 0, 0, 4}, {5, 16, 1, 1, 0, 0, 2}};
    for (const Cfg& c : cfgs) {
      if (N % (32 * (c.E / 4) * c.J) != 0 || (size_t)c.E * c.T * c.J * 4096 > 220 * 1024) continue;
      CUtensorMap tm3;
      {
        cuuint64_t d3[3] = {32, (cuuint64_t)(N / 32), (cuuint64_t)M}; cuuint64_t s3[2] = {128, (cuuint64_t)N * 4}; cuuint32_t b3[3] = {32, (cuuint32_t)c.J, 32}; cuuint32_t e3[3] = {1, 1, 1};
        CUresult r3 = enc(&tm3, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, Y, d3, s3, b3, e3, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r3 != CUDA_SUCCESS) { printf("3D encode failed %d (J=%d)\n", (int)r3, c.J); continue; }
      }
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      float best = 1e9;
      for (int i = 0; i < 4; ++i) {
        cudaMemsetAsync(flush, 1, 256u << 20);
        cudaEventRecord(a);
        store_kernel<<<ctas, c.E * 32, (size_t)c.E * c.T * c.J * 4096>>>(tm, Y, N, rows, c.E, c.T, c.mode, c.sts, c.hint, c.il, tm3, c.J);
        cudaEventRecord(b);
        CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
      }
      CK(cudaGetLastError());
      const char* names[] = {"TMA tensor store 32x128B", "cp.async.bulk 128 B per lane", "st.global.v4 thread=row", "smem transpose + st.global.v4 rows", "TMA 3D store 32 rows x J x 128B", "staged + plain STG.128, J x 128 B rows"};
      printf("N=%4d %-36s warps=%2d tiles/warp=%d sts=%d hint=%d interleaved=%d J=%d : %6.0f GB/s  %5.1f B/clk/SM @1.9GHz  (%.1f us incl. launch)\n", N, names[c.mode], c.E, c.T, c.sts, c.hint, c.il, c.J,
             bytes / best / 1e6, bytes / best / 1e6 / 148 / 1.9, best * 1e3);
    }
  }
  return 0;
}
