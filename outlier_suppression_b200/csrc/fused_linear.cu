// fused_linear.cu -- K6: activation fake-quant + per-channel weight fake-quant + Linear in ONE
// persistent, warp-specialised tcgen05 kernel (sm_100a).
//
//   Y[m,n] = s_a * w_scale[n] * ( sum_k qa[m,k] * wc[n,k]  -  Zc * rowsum[n] ) + bias[n]
//
//   qa  = clamp(rint(A/s_a) + Z, qmin, qmax) - qmin   in [0, 255]   (u8, produced IN-KERNEL from fp32 A)
//   wc  = weight bins (s8, packed once per weight version by osq_pack_weight_s8)
//   Zc  = Z - qmin,  rowsum[n] = sum_k wc[n,k]
//
// The contraction is an exact u8 x s8 -> s32 `tcgen05.mma kind::i8` with the accumulator in TMEM.
//
// Data flow per CTA (one 128-row block of A at a time, all of N for that block):
//
//   warps 8-15        A path: 128-bit streaming loads of fp32 A straight into registers (16 rows x 512 B
//                     in flight per warp, software pipelined in two halves) -> integer bins -> A ring in
//                     the UMMA K-major SW128 shared-memory layout                    (a_full / a_empty)
//   warp 1  (1 lane)  TMA: s8 weight tiles [BN rows x 128 k], SW128 -> W ring        (w_full / w_empty)
//   warp 2  (1 lane)  tcgen05.mma  D[tmem] (+)= A[smem] * W[smem]^T, commit -> w_empty/a_empty/acc_full
//   warps 4-7         epilogue: tcgen05.ld -> zero-point correction, scales, bias -> swizzled smem tile
//                     -> TMA store (full 128-byte lines) of fp32 Y
//
// When all of K fits in the A ring (K/128 <= 8: BERT-base 768, BART 1024) the converted A block stays
// RESIDENT in shared memory and is reused for every N chunk: each activation element is read from HBM
// once and quantised once.  Otherwise (K = 3072/4096) pass 0 converts A and also spills the bins (1 B per
// element) to a caller-provided code cache that stays in L2; the remaining N chunks re-load the bins
// by TMA directly in the UMMA layout -- fp32 A is still read from HBM exactly once.
#include <cuda.h>

#include "common.cuh"

namespace osq {

constexpr int kBM = 128;            // rows of A per CTA tile (UMMA M)
constexpr int kBNMax = 256;         // columns per accumulator stage (UMMA N)
constexpr int kStageK = 128;        // k elements (= bytes, u8/s8) per smem stage row: one 128B swizzle row
constexpr int kUmmaK = 32;          // k per tcgen05.mma kind::i8
constexpr int kAStageBytes = kBM * kStageK;          // 16 KB
constexpr int kWStageBytes = kBNMax * kStageK;       // 32 KB
constexpr int kAccStages = 2;
constexpr int kTmemCols = kAccStages * kBNMax;       // 512
constexpr int kNumThreads = 512;
constexpr int kConvWarp0 = 8, kNumConvWarps = 8;
constexpr int kEpiWarp0 = 4, kNumEpiWarps = 4;
constexpr int kRowsPerConvWarp = kBM / kNumConvWarps;  // 16
constexpr int kHalf = kRowsPerConvWarp / 2;            // 8 rows per software-pipeline half
constexpr int kOutTileBytes = 32 * 128;                // 32 rows x 32 fp32 columns, SW128
constexpr int kMaxAStages = 8, kMaxWStages = 4;
constexpr float kMagic = 12582912.f;  // 1.5 * 2^23: fp32 ulp is 1 in [2^23, 2^24)

struct FusedParams {
  int M, K, N;
  int KB;          // K / 128
  int NC;          // number of N chunks
  int BN;          // chunk width (<= 256, multiple of 16)
  int n_mblocks;
  int a_stages, w_stages, out_bufs;
  int resident;    // converted A block stays in smem for all N chunks
  int cached;      // streaming mode with a code cache: passes >= 1 TMA-load bins instead of re-converting
  const float* A;
  const float* a_scale;
  const void* a_zp;
  int a_zp_is_int32;
  float g;
  float qmin, qmax;
  const float* w_scale;
  const int32_t* w_rowsum;
  const float* bias;
  uint8_t* a_codes;  // optional [M, K]: bins side output / code cache
  uint32_t codes_box_bytes;  // bytes one code-cache TMA box delivers
};

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]^T,  u8 x s8 -> s32
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on `bar` once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);  // start address
  d |= (uint64_t)1 << 16;                  // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;        // stride byte offset: 8 rows * 128 B
  d |= (uint64_t)1 << 46;                  // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                  // SWIZZLE_128B
  return d;
}
// cute::UMMA::InstrDescriptor for kind::i8: D=s32, A=u8, B=s8, both K-major, M=128
__host__ __device__ inline uint32_t make_idesc_i8(int n) {
  return (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}

// ------------------------------------------------------------------------------------------
// fp32 -> bin conversion.  Bit-exact with clamp(rint(x / s) + Z, qmin, qmax) (true IEEE division):
// u = fma(x, 1/s, magic + Zc) rounds x/s (approximately) to the integer grid; the residual
// e = fma(x, 1/s, -(u - magic - Zc)) tells how close x/s is to a rounding tie.  Only when
// |e| > 0.4999 (probability 2e-4) can the reciprocal's <=2 ulp error change the bin, and only then
// the exact division is evaluated.  For |x/s| >= 400 the bin saturates on both paths.
// ------------------------------------------------------------------------------------------
struct ConvParam {
  float s, rinv, mz, lo, hi, zc, span;
};

__device__ __forceinline__ uint32_t quant_bin(float x, const ConvParam& c) {
  float u = fmaf(x, c.rinv, c.mz);
  float nf = __fsub_rn(u, c.mz);
  float e = fmaf(x, c.rinv, -nf);
  if (!(fabsf(e) <= 0.4999f)) {  // near a tie, huge, or NaN: exact path
    float t = __fdiv_rn(x, c.s);
    float v = __fadd_rn(rintf(t), c.zc);
    v = fminf(fmaxf(v, 0.f), c.span);
    u = __fadd_rn(v, kMagic);
  }
  u = fminf(fmaxf(u, c.lo), c.hi);
  return __float_as_uint(u);  // low byte = bin - qmin
}

__device__ __forceinline__ uint32_t pack4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}

struct Smem {
  uint64_t a_full[kMaxAStages], a_empty[kMaxAStages];
  uint64_t w_full[kMaxWStages], w_empty[kMaxWStages];
  uint64_t acc_full[kAccStages], acc_empty[kAccStages];
  uint64_t codes_ready;
  uint32_t tmem_base;
  uint32_t pad;
  alignas(16) float c1[kAccStages][kBNMax];      // s_a * w_scale[n]
  alignas(16) int32_t zr[kAccStages][kBNMax];    // Zc * rowsum[n]
  alignas(16) float bias[kAccStages][kBNMax];
};

__global__ void __launch_bounds__(kNumThreads, 1)
fused_fq_linear_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_y,
                       const __grid_constant__ CUtensorMap tmap_codes, const FusedParams p) {
  // dynamic shared memory, 1024B aligned by the attribute (SWIZZLE_128B tiles need it):
  // [A ring][W ring][out staging][Smem bookkeeping]
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* a_ring = smem_raw;
  uint8_t* w_ring = a_ring + (size_t)p.a_stages * kAStageBytes;
  uint8_t* o_ring = w_ring + (size_t)p.w_stages * kWStageBytes;
  Smem& sm = *reinterpret_cast<Smem*>(o_ring + (size_t)kNumEpiWarps * p.out_bufs * kOutTileBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_y);
    if (p.cached) tma_prefetch_desc(&tmap_codes);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.a_stages; ++i) { mbar_init(&sm.a_full[i], kNumConvWarps); mbar_init(&sm.a_empty[i], 1); }
    for (int i = 0; i < p.w_stages; ++i) { mbar_init(&sm.w_full[i], 1); mbar_init(&sm.w_empty[i], 1); }
    for (int i = 0; i < kAccStages; ++i) { mbar_init(&sm.acc_full[i], 1); mbar_init(&sm.acc_empty[i], kNumEpiWarps); }
    mbar_init(&sm.codes_ready, kNumConvWarps);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&sm.tmem_base, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;

  const int n_my_blocks = (p.n_mblocks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int a_passes = p.resident ? 1 : p.NC;  // how many times the A ring is filled per m-block

  if (warp == 1) {
    // ===================== TMA producer: packed weight tiles =====================
    if (lane == 0) {
      uint32_t pw = 0;
      const uint32_t w_bytes = (uint32_t)p.BN * kStageK;
      for (int it = 0; it < n_my_blocks; ++it)
        for (int nc = 0; nc < p.NC; ++nc)
          for (int kb = 0; kb < p.KB; ++kb, ++pw) {
            const int ws = pw % p.w_stages;
            mbar_wait(&sm.w_empty[ws], ((pw / p.w_stages) & 1) ^ 1);
            mbar_arrive_expect_tx(&sm.w_full[ws], w_bytes);
            tma_load_2d(w_ring + (size_t)ws * kWStageBytes, &tmap_w, &sm.w_full[ws], kb * kStageK, nc * p.BN);
          }
    }
  } else if (warp == 2) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_i8(p.BN);
      uint32_t cw = 0, cacc = 0;
      for (int it = 0; it < n_my_blocks; ++it)
        for (int nc = 0; nc < p.NC; ++nc, ++cacc) {
          const int as_ = cacc & 1;
          mbar_wait(&sm.acc_empty[as_], ((cacc >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)as_ * kBNMax;
          for (int kb = 0; kb < p.KB; ++kb, ++cw) {
            const uint32_t ca = p.resident ? (uint32_t)(it * p.KB + kb) : (uint32_t)((it * p.NC + nc) * p.KB + kb);
            const int a_st = ca % p.a_stages;
            const int ws = cw % p.w_stages;
            mbar_wait(&sm.a_full[a_st], (ca / p.a_stages) & 1);
            mbar_wait(&sm.w_full[ws], (cw / p.w_stages) & 1);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(a_ring + (size_t)a_st * kAStageBytes);
            const uint32_t w_addr = smem_u32(w_ring + (size_t)ws * kWStageBytes);
#pragma unroll
            for (int k = 0; k < kStageK / kUmmaK; ++k)
              umma_i8(d_tmem, make_smem_desc(a_addr + k * kUmmaK), make_smem_desc(w_addr + k * kUmmaK), idesc,
                      (kb | k) != 0);
            umma_commit(&sm.w_empty[ws]);
            if (!p.resident || nc == p.NC - 1) umma_commit(&sm.a_empty[a_st]);
          }
          umma_commit(&sm.acc_full[as_]);
        }
    }
  } else if (warp >= kEpiWarp0 && warp < kEpiWarp0 + kNumEpiWarps) {
    // ===================== epilogue =====================
    const int wg = warp - kEpiWarp0;  // == warp % 4: TMEM lanes [32*wg, 32*wg+32)
    const int et = threadIdx.x - kEpiWarp0 * 32;
    const QParam qp = load_qparam(p.a_scale, p.a_zp, p.a_zp_is_int32, p.g, p.qmin, p.qmax, false);
    const float s_a = qp.s;
    const int zc = (int)(rintf(qp.z) - p.qmin);
    uint8_t* my_out = o_ring + (size_t)wg * p.out_bufs * kOutTileBytes;
    const uint32_t sw = ((uint32_t)lane & 7) << 4;  // 128B swizzle phase of this thread's staging row
    uint32_t cacc = 0, n_stores = 0;
    for (int it = 0; it < n_my_blocks; ++it) {
      const int mb = blockIdx.x + it * gridDim.x;
      const int row0 = mb * kBM + wg * 32;
      for (int nc = 0; nc < p.NC; ++nc, ++cacc) {
        const int as_ = cacc & 1;
        const int n0 = nc * p.BN;
        // stage the per-column constants of this chunk (double buffered with the accumulator stage)
        for (int c = et; c < p.BN; c += kNumEpiWarps * 32) {
          const int n = n0 + c;
          const bool ok = n < p.N;
          sm.c1[as_][c] = ok ? __fmul_rn(s_a, __ldg(p.w_scale + n)) : 0.f;
          sm.zr[as_][c] = ok ? zc * __ldg(p.w_rowsum + n) : 0;
          sm.bias[as_][c] = (ok && p.bias != nullptr) ? __ldg(p.bias + n) : 0.f;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kNumEpiWarps * 32) : "memory");
        mbar_wait(&sm.acc_full[as_], (cacc >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(wg * 32) << 16) + (uint32_t)as_ * kBNMax;
        for (int c0 = 0; c0 < p.BN && n0 + c0 < p.N; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + c0, v);
          tmem_ld_wait();
          // the staging tile must have been read out by its previous TMA store
          if (n_stores >= (uint32_t)p.out_bufs) {
            if (lane == 0) {
              if (p.out_bufs == 2) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
            }
            __syncwarp();
          }
          uint8_t* tile = my_out + (size_t)(n_stores % p.out_bufs) * kOutTileBytes;
          uint8_t* trow = tile + lane * 128;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 c1 = *reinterpret_cast<const float4*>(&sm.c1[as_][c0 + j]);
            const int4 zr = *reinterpret_cast<const int4*>(&sm.zr[as_][c0 + j]);
            const float4 bi = *reinterpret_cast<const float4*>(&sm.bias[as_][c0 + j]);
            float4 o;
            o.x = fmaf((float)((int)v[j + 0] - zr.x), c1.x, bi.x);
            o.y = fmaf((float)((int)v[j + 1] - zr.y), c1.y, bi.y);
            o.z = fmaf((float)((int)v[j + 2] - zr.z), c1.z, bi.z);
            o.w = fmaf((float)((int)v[j + 3] - zr.w), c1.w, bi.w);
            *reinterpret_cast<float4*>(trow + ((uint32_t)(j << 2) ^ sw)) = o;  // chunk (j/4) ^ (row & 7)
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmap_y, tile, n0 + c0, row0);  // rows >= M / columns >= N are clipped by the TMA unit
            tma_store_commit();
          }
          ++n_stores;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.acc_empty[as_]);
      }
    }
    if (lane == 0) tma_store_wait_all();
  } else if (warp >= kConvWarp0) {
    // ===================== A path: fp32 -> bins in the UMMA smem layout =====================
    const int cw_ = warp - kConvWarp0;
    const QParam qp = load_qparam(p.a_scale, p.a_zp, p.a_zp_is_int32, p.g, p.qmin, p.qmax,
                                  blockIdx.x == 0 && cw_ == 0 && lane == 0);
    ConvParam cp;
    cp.s = qp.s;
    cp.rinv = __frcp_rn(qp.s);
    cp.zc = rintf(qp.z) - p.qmin;
    cp.span = p.qmax - p.qmin;
    cp.mz = kMagic + cp.zc;
    cp.lo = kMagic;
    cp.hi = kMagic + cp.span;
    const int r_base = cw_ * kRowsPerConvWarp;  // this warp's 16 rows of the 128-row block
    // byte offset of this lane's 4 bins inside a swizzled 128 B stage row (before the row term)
    const uint32_t lane_chunk = (uint32_t)lane >> 2, lane_in = ((uint32_t)lane & 3) << 2;

    float4 xa[kHalf], xb[kHalf];
    auto load_half = [&](float4* x, int mb, int kb, int half) {
      const float* base = p.A + (size_t)kb * kStageK + lane * 4;
#pragma unroll
      for (int i = 0; i < kHalf; ++i) {
        const int grow = mb * kBM + r_base + half * kHalf + i;
        x[i] = (grow < p.M) ? ldg_stream(reinterpret_cast<const float4*>(base + (size_t)grow * p.K))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto convert_half = [&](const float4* x, uint8_t* a_tile, int mb, int kb, int half) {
#pragma unroll
      for (int i = 0; i < kHalf; ++i) {
        const int r = r_base + half * kHalf + i;
        const uint32_t word = pack4(quant_bin(x[i].x, cp), quant_bin(x[i].y, cp), quant_bin(x[i].z, cp), quant_bin(x[i].w, cp));
        const uint32_t off = (uint32_t)r * kStageK + ((lane_chunk ^ ((uint32_t)r & 7)) << 4) + lane_in;
        *reinterpret_cast<uint32_t*>(a_tile + off) = word;
        if (p.a_codes != nullptr) {
          const int grow = mb * kBM + r;
          if (grow < p.M) *reinterpret_cast<uint32_t*>(p.a_codes + (size_t)grow * p.K + kb * kStageK + lane * 4) = word;
        }
      }
    };

    uint32_t pa = 0;
    if (n_my_blocks > 0) {
      load_half(xa, blockIdx.x, 0, 0);
      load_half(xb, blockIdx.x, 0, 1);
    }
    for (int it = 0; it < n_my_blocks; ++it) {
      const int mb = blockIdx.x + it * gridDim.x;
      for (int pass = 0; pass < a_passes; ++pass) {
        if (pass == 0 || !p.cached) {
          // ---- conversion pass: registers -> bins -> smem (next k-block's loads are issued in between) ----
          for (int kb = 0; kb < p.KB; ++kb, ++pa) {
            const int a_st = pa % p.a_stages;
            mbar_wait(&sm.a_empty[a_st], ((pa / p.a_stages) & 1) ^ 1);
            uint8_t* a_tile = a_ring + (size_t)a_st * kAStageBytes;
            // what to prefetch next: the following k-block of this pass, the first of the next
            // conversion pass, or the first k-block of this CTA's next m-block
            int nmb = mb, nkb = kb + 1;
            bool more = true;
            if (nkb == p.KB) {
              nkb = 0;
              const bool next_pass_converts = (pass + 1 < a_passes) && !p.cached;
              if (!next_pass_converts) { nmb = mb + gridDim.x; more = (it + 1 < n_my_blocks); }
            }
            convert_half(xa, a_tile, mb, kb, 0);
            if (more) load_half(xa, nmb, nkb, 0);
            convert_half(xb, a_tile, mb, kb, 1);
            if (more) load_half(xb, nmb, nkb, 1);
            fence_proxy_async_smem();  // generic-proxy smem stores -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.a_full[a_st]);
          }
          if (p.cached && pass == 0) {
            // bins of this m-block are in the code cache: publish them to the async proxy (TMA) of this CTA
            __threadfence();
            fence_proxy_async_all();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.codes_ready);
          }
        } else {
          // ---- cached pass: bins come back from the (L2 resident) code cache by TMA, already in UMMA layout ----
          if (pass == 1) mbar_wait(&sm.codes_ready, it & 1);
          for (int kb = 0; kb < p.KB; ++kb, ++pa) {
            const int a_st = pa % p.a_stages;
            // every warp paces itself on a_empty so that at most one arrival per warp lands in a phase
            mbar_wait(&sm.a_empty[a_st], ((pa / p.a_stages) & 1) ^ 1);
            if (lane == 0) {
              if (cw_ == 0) {
                mbar_arrive_expect_tx(&sm.a_full[a_st], p.codes_box_bytes);
                tma_load_2d(a_ring + (size_t)a_st * kAStageBytes, &tmap_codes, &sm.a_full[a_st], kb * kStageK, mb * kBM);
              } else {
                mbar_arrive(&sm.a_full[a_st]);  // keeps the arrival count of a_full uniform across pass kinds
              }
            }
            __syncwarp();
          }
        }
      }
    }
  }

  // ===================== teardown =====================
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------
// weight packing: bins (q - zp) as s8 + per-row sums
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_weight_s8_kernel(const float* __restrict__ w, int64_t N, int64_t K, const float* __restrict__ scale,
                      const int32_t* __restrict__ zp, float qmin, float qmax, int8_t* __restrict__ codes,
                      int32_t* __restrict__ rowsum) {
  __shared__ int ssum[8];
  for (int64_t n = blockIdx.x; n < N; n += gridDim.x) {
    const float s = scale[n];
    const float z = (float)zp[n];
    int acc = 0;
    for (int64_t k = threadIdx.x; k < K; k += blockDim.x) {
      float q;
      fq_elem(w[n * K + k], s, z, qmin, qmax, q);
      int c = (q != q) ? 0 : (int)(q - z);
      codes[n * K + k] = (int8_t)c;
      acc += c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) ssum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += ssum[i];
      rowsum[n] = t;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = (EncodeTiledFn)sym;
  }
  return fn;
}

static int make_map_2d(CUtensorMap* map, CUtensorMapDataType dt, int elem_bytes, const void* ptr, uint64_t inner,
                       uint64_t outer, uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle sw) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return OSQ_ECUDA;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {inner * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu outer=%llu box=%ux%u)", (int)r,
              (unsigned long long)inner, (unsigned long long)outer, box_inner, box_outer);
    return OSQ_ECUDA;
  }
  return OSQ_OK;
}

}  // namespace osq

extern "C" {

int osq_pack_weight_s8(const float* w, int64_t N, int64_t K, const float* scale, const int32_t* zp, int qmin, int qmax,
                       int8_t* codes, int32_t* rowsum, void* stream) {
  using namespace osq;
  OSQ_CHECK_ARG(w && scale && zp && codes && rowsum, "osq_pack_weight_s8: null pointer");
  OSQ_CHECK_ARG(N > 0 && K > 0, "osq_pack_weight_s8: empty weight");
  OSQ_CHECK_ARG(qmax - qmin <= 255 && qmin < qmax, "osq_pack_weight_s8: more than 8 bits");
  int sms = sm_count();
  if (sms <= 0) { set_error("no CUDA device"); return OSQ_ECUDA; }
  int64_t g = N < (int64_t)sms * 8 ? N : (int64_t)sms * 8;
  pack_weight_s8_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(w, N, K, scale, zp, (float)qmin, (float)qmax, codes, rowsum);
  OSQ_LAUNCH_CHECK();
  return OSQ_OK;
}

int osq_fused_fq_linear(const osq_fused_linear_t* a, void* stream) {
  using namespace osq;
  OSQ_CHECK_ARG(a != nullptr, "osq_fused_fq_linear: null args");
  OSQ_CHECK_ARG(a->A && a->a_scale && a->a_zp && a->w_codes && a->w_scale && a->w_rowsum && a->Y,
                "osq_fused_fq_linear: null pointer");
  OSQ_CHECK_ARG(a->M >= 1 && a->M < (1ll << 31) - 256, "osq_fused_fq_linear: M out of range");
  OSQ_CHECK_ARG(a->K >= kStageK && a->K % kStageK == 0 && a->K <= (1 << 20), "osq_fused_fq_linear: K must be a multiple of 128");
  OSQ_CHECK_ARG(a->N >= 16 && a->N % 16 == 0 && a->N <= (1 << 20), "osq_fused_fq_linear: N must be a multiple of 16");
  OSQ_CHECK_ARG(a->a_qmax - a->a_qmin <= 255 && a->a_qmin < a->a_qmax, "osq_fused_fq_linear: activation bits > 8");
  OSQ_CHECK_ARG(a->mma_kind == 0 || a->mma_kind == 1, "osq_fused_fq_linear: mma_kind %d not built", a->mma_kind);
  OSQ_CHECK_ARG((((uintptr_t)a->A) & 15) == 0 && (((uintptr_t)a->Y) & 15) == 0 && (((uintptr_t)a->w_codes) & 15) == 0,
                "osq_fused_fq_linear: A, Y and w_codes must be 16-byte aligned");
  OSQ_CHECK_ARG(!(a->lsq_grad_factor > 0.f && a->a_zp_is_int32), "osq_fused_fq_linear: LSQ+ needs a float zero_point");
  OSQ_CHECK_ARG(a->a_codes == nullptr || (((uintptr_t)a->a_codes) & 15) == 0, "osq_fused_fq_linear: a_codes must be 16-byte aligned");

  int dev = 0, cc_major = 0, sms = sm_count();
  OSQ_CUDA(cudaGetDevice(&dev));
  OSQ_CUDA(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  if (cc_major != 10) {
    set_error("osq_fused_fq_linear needs an sm_100 device (found compute capability major %d)", cc_major);
    return OSQ_EARCH;
  }

  FusedParams p;
  p.M = (int)a->M; p.K = (int)a->K; p.N = (int)a->N;
  p.KB = p.K / kStageK;
  p.BN = p.N < kBNMax ? p.N : kBNMax;
  p.NC = (p.N + p.BN - 1) / p.BN;
  p.n_mblocks = (p.M + kBM - 1) / kBM;
  // shared-memory plan (227 KB / CTA): [A ring][W ring][out staging][bookkeeping]
  const int budget = 227 * 1024 - (int)sizeof(Smem);
  const int out1 = kNumEpiWarps * kOutTileBytes;  // one 4 KB staging tile per epilogue warp
  if (p.KB <= kMaxAStages && p.KB * kAStageBytes + 2 * kWStageBytes + out1 <= budget) {
    p.resident = 1;
    p.a_stages = p.KB;
  } else {
    p.resident = 0;
    p.a_stages = 4;
  }
  p.cached = (!p.resident && p.NC > 1 && a->a_codes != nullptr) ? 1 : 0;
  int rest = budget - p.a_stages * kAStageBytes - 2 * kWStageBytes - out1;
  p.w_stages = 2;
  p.out_bufs = 1;
  if (rest >= out1) { p.out_bufs = 2; rest -= out1; }
  while (p.w_stages < kMaxWStages && rest >= kWStageBytes) { ++p.w_stages; rest -= kWStageBytes; }
  if (!p.resident)
    while (p.a_stages < kMaxAStages && rest >= kAStageBytes) { ++p.a_stages; rest -= kAStageBytes; }
  const size_t smem_bytes = (size_t)p.a_stages * kAStageBytes + (size_t)p.w_stages * kWStageBytes +
                            (size_t)p.out_bufs * out1 + sizeof(Smem);

  p.A = a->A;
  p.a_scale = a->a_scale; p.a_zp = a->a_zp; p.a_zp_is_int32 = a->a_zp_is_int32; p.g = a->lsq_grad_factor;
  p.qmin = (float)a->a_qmin; p.qmax = (float)a->a_qmax;
  p.w_scale = a->w_scale; p.w_rowsum = a->w_rowsum; p.bias = a->bias; p.a_codes = a->a_codes;
  p.codes_box_bytes = (uint32_t)(p.M < kBM ? p.M : kBM) * kStageK;

  CUtensorMap map_w, map_y, map_c;
  if (int rc = make_map_2d(&map_w, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, a->w_codes, (uint64_t)p.K, (uint64_t)p.N, kStageK,
                           (uint32_t)p.BN, CU_TENSOR_MAP_SWIZZLE_128B))
    return rc;
  if (int rc = make_map_2d(&map_y, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a->Y, (uint64_t)p.N, (uint64_t)p.M, 32,
                           p.M < 32 ? (uint32_t)p.M : 32u, CU_TENSOR_MAP_SWIZZLE_128B))
    return rc;
  if (p.cached) {
    if (int rc = make_map_2d(&map_c, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, a->a_codes, (uint64_t)p.K, (uint64_t)p.M, kStageK,
                             p.M < kBM ? (uint32_t)p.M : (uint32_t)kBM, CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  } else {
    map_c = map_w;  // unused by the kernel
  }

  static bool attr_set[64] = {false};
  if (!attr_set[dev & 63]) {
    OSQ_CUDA(cudaFuncSetAttribute(fused_fq_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set[dev & 63] = true;
  }
  int grid = p.n_mblocks < sms ? p.n_mblocks : sms;
  fused_fq_linear_kernel<<<grid, kNumThreads, smem_bytes, (cudaStream_t)stream>>>(map_w, map_y, map_c, p);
  OSQ_LAUNCH_CHECK();
  return OSQ_OK;
}

}  // extern "C"
