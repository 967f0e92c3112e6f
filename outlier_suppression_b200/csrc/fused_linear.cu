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
//   warps 4-19  (16)  workers, phase A: fp32 A arrives by TMA in one 4 KB landing slot per warp (its 8 rows x one
//                     k-block); slot -> registers -> integer bins -> A ring in the UMMA K-major SW128
//                     shared-memory layout; the slot is re-armed as soon as it is in registers
//                                                                                  (x_full, a_full / a_empty)
//                     (variant `_ldg`: 128-bit streaming loads straight into registers instead of the slots)
//   warp 0  (1 lane)  TMA: s8 weight tiles [BN rows x 128 k], SW128 -> W ring      (w_full / w_empty)
//   warp 1  (1 lane)  tcgen05.mma  D[tmem] (+)= A[smem] * W[smem]^T, commit -> w_empty/a_empty/acc_full
//   warp 2            TMEM allocation; (1 lane) TMA re-load of cached bins for the later sweeps when K > 1024
//   warp 3            idle (optional L2 prefetcher of the `_ldg` variant)
//   warps 4-11 (8)    workers, phase B (epilogue): tcgen05.ld -> zero-point correction, scales, bias ->
//                     swizzled smem tile (the landing slots, two per warp) -> TMA store of fp32 Y
//
// When all of K fits in the A ring (K/128 <= 8: BERT-base 768, BART 1024) the converted A block stays
// RESIDENT in shared memory and is reused for every N chunk: each activation element is read from HBM
// once and quantised once.  Otherwise (K = 3072/4096) the first sweep converts A, feeds every TMEM accumulator
// stage per k-block and spills the bins (1 B per element) to a caller-provided code cache that stays in L2; the
// remaining N chunks re-load the bins by TMA directly in the UMMA layout -- fp32 A is still read from HBM once.
// Variant `_pair` runs two CTAs as one tcgen05 cta_group::2 (M = 256): see DESIGN.md section 5.
#include <cuda.h>
#include <type_traits>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace osq {

constexpr int kBM = 128;            // rows of A per CTA tile (UMMA M)
constexpr int kStageK = 128;        // k elements (= bytes, u8/s8) per smem stage row: one 128B swizzle row
constexpr int kUmmaK = 32;          // k per tcgen05.mma kind::i8
constexpr int kTmemCols = 512;                       // all of TMEM: acc stages x BN (2 x 256, 2 x 192, 4 x 128)
constexpr int kWorkerWarp0 = 4, kNumWorkers = 16;    // warps 4..19 convert A (phase A); warps 4..11 also run the epilogue
constexpr int kNumEpiWarps = 8;                      // workers 0..7: four TMEM lane quarters x two column slices of a chunk (kEW = 16: all workers, four slices)
constexpr int kNumThreads = (kWorkerWarp0 + kNumWorkers) * 32;  // 640
constexpr int kRowsPerWorker = kBM / kNumWorkers;    // 8
constexpr int kOutTileBytes = 32 * 128;              // TMA-store staging tile: 32 rows x 32 fp32 columns, SW128
constexpr int kXSlotBytes = kRowsPerWorker * kStageK * 4;  // fp32 landing slot of one worker: 8 rows x 128 floats = 4 KB
constexpr int kMaxAStages = 8, kMaxWStages = 6, kMaxAccStages = 4;

struct FusedParams {
  int M, K, N;
  int KB;             // K / 128
  int NC;             // number of N chunks
  int BN;             // chunk width chosen by the host plan: 256 (CTA pairs, streamed A), 192 (single CTAs, resident A), or N when N is smaller
  int acc_stages;     // TMEM accumulator stages (512 / BN, at most 4)
  int n_mblocks;
  int csz;            // 1, or 2 = CTA pair: tcgen05 cta_group::2 (M = 256 over two SMs, each CTA stages half of every W tile)
  int mb_count;       // CTAs along M (= grid / nsplit); a CTA's row tiles are mb_lane, mb_lane + mb_count, ...
  int nsplit, cps;    // N is split over `nsplit` groups of CTAs, `cps` chunks each (few rows: keeps row tiles >= 64 and still fills the SMs)
  int n_iters;        // tiles per CTA (identical for every CTA; out-of-range tiles are phantoms that only keep the W protocol alive)
  int rows_per_tile;  // valid rows per CTA tile (<= 128, multiple of 16): chosen so the tile count fills all SMs
  int a_stages, w_stages, out_bufs;
  int a_stage_bytes;  // rows_per_tile * 128
  int a_passes;       // sweeps over K per m-block: 1 (resident A), ceil(NC / cpp) (streamed A)
  int cpp;            // N chunks accumulated concurrently in TMEM per sweep (streamed A with a code cache: all acc stages)
  int x_tma;          // fp32 A arrives by TMA into per-worker landing slots (else: 128-bit loads into registers)
  int alias_xo;       // the TMA-store staging tiles share the landing slots' memory
  int w_stage_bytes;  // (BN / csz) * 128: a CTA of a pair stages half of the tile's rows
  int resident;       // converted A block stays in smem for all N chunks
  int cached;         // streaming mode with a code cache: passes >= 1 TMA-load bins instead of re-converting
  int codes_in;       // A is NULL: a_codes already holds the activation bins (written by the upstream fake-quant kernel);
                      // every sweep TMA-loads them, nothing is converted
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
  float* Y;          // [M, N] output (the drain epilogue writes it with plain stores; the other variants go through tmap_y)
  int x_depth;       // fp32 landing slots per worker (1, or 2 on streamed plans: k-block kb lands in slot kb & 1)
  int pre_l2;        // k-blocks of this CTA's fp32 tile prefetched into L2 before the wait for the previous grid (0 = off)
  int lsu_mod;       // plain tile stores: every lsu_mod-th step of a warp is written by the LSU instead of the TMA unit (0 = never)
  int epi16;         // epilogue variant: all sixteen workers (four column slices), single staging tile per warp
  int drain;         // epilogue variant: staged tiles drained by workers 8..15 with 128-bit stores instead of TMA tensor stores
  uint32_t codes_box_bytes;  // bytes one code-cache TMA box delivers
  int pdl;                   // launched with programmatic stream serialization
  int prefetch;              // L2 prefetch distance of the fp32 activation in k-blocks (0 = off)
  int store3d;               // epilogue: the two column slices of a lane quarter share one 8 KB staging tile and leave as ONE 3-D tensor store
                             // of 32 rows x 2 x 128 B (256 contiguous bytes per row); tmap_y / tmap_y16 are then 3-D maps [M][N/32][32]
  // output stage (kEpi variants): Y is replaced by fq(act(Y)) of the NEXT activation quantizer, whose bins also leave as u8
  int out_act;               // 0 = none, 1 = GELU (erf form, the op order of ATen's CUDA kernel)
  const float* out_scale;    // device [1]
  const void* out_zp;        // device [1]
  int out_zp_is_int32;
  float out_g;               // LSQ+ grad factor of the output quantizer (0: FixedFakeQuantize)
  float out_qmin, out_qmax;
  uint8_t* out_bins;         // [M, N] u8 (bin - out_qmin) or NULL
  int more_sites;            // multi-site launch: another site follows in this kernel (barriers are invalidated at teardown, TMEM is kept)
  int first_site;            // this site allocates TMEM (always 1 for a single-site launch)
  int dbg;                   // profiling experiments (OSQ_FUSED_DBG): 1 = W tile pinned, 2 = no Y stores, 4 = A rows pinned
  long long* trace;          // optional debug timeline: CTA 0 clock64 stamps [0,1024), per-CTA globaltimer start/end [1024, 1024+2*grid)
};

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_inval(uint64_t* bar) {
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_n(uint64_t* bar, uint32_t n) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(n) : "memory");
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
// same primitives on precomputed 32-bit shared addresses (the single-thread issue loops must stay lean:
// their own instruction latency, not the tensor core, was the bottleneck with generic addressing)
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_u32(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void umma_commit_u32(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// ---- CTA pair (cta_group::2): the leader CTA issues the MMAs for both SMs; barriers the leader waits on collect
// arrivals from both CTAs, barriers both CTAs wait on are signalled by multicast commits.  Waits and remote arrives
// use the default (CTA-scope) semantics, as CUTLASS' 2-SM pipelines do: the data these barriers guard never leaves
// the SM that wrote it (each tensor core reads its own CTA's shared memory).
__device__ __forceinline__ void umma_commit_pair_u32(uint32_t bar) {  // arrives on `bar` at this offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {  // shared::cta address -> shared::cluster address in CTA `rank`
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Remote arrive on the leader's barrier.  Default semantics (release at CTA scope), as in CUTLASS' 2-SM pipelines: what
// the arrive orders are this CTA's own shared-memory writes (already made visible to the async proxy by the
// fence.proxy.async in front of it), which only this SM's tensor core reads.  `.release.cluster` here cost 1.7 K
// cycles per k-block in the conversion loop (5.0 K instead of 3.3 K) and made the pair slower than two single CTAs.
__device__ __forceinline__ void mbar_arrive_cluster_u32(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster_u32(uint32_t cluster_bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_n_cluster_u32(uint32_t cluster_bar, uint32_t n) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_bar), "r"(n) : "memory");
}
__device__ __forceinline__ void tma_load_2d_u32(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// the same load with an L2 eviction-priority hint (createpolicy): fp32 A is touched once, it should not displace the code cache
__device__ __forceinline__ void tma_load_2d_hint_u32(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ bool hint_a_any(const FusedParams& p) { return !(p.dbg & 512); }
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap* map, const void* smem_src, int c0, int c1, uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;" ::"l"(map),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void stg_hint(float4* dst, const float4 v, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(policy) : "memory");
}
// pair mode: lands in this CTA's shared memory, completes transaction bytes on the LEADER's barrier (cluster address)
__device__ __forceinline__ void tma_load_2d_pair_u32(uint32_t smem_dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_dst),
      "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
// programmatic dependent launch: let the next kernel of the stream start its prologue / weight stream
// while this one drains, and order this kernel's first dependent access after the previous grid
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait_prior_grids() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
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
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// named barrier of the two epilogue warps (64 threads) of lane quarter q; immediate ids keep ptxas from reserving all 16
__device__ __forceinline__ void pair_barrier(int q) {
  switch (q) {
    case 0: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 5, 64;" ::: "memory"); break;
  }
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
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

__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {  // one warp of EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// pair: D (256 x N: rows 0..127 in this CTA's TMEM, 128..255 in the peer's) (+)= A (each CTA's own 128 rows) * B^T
// (each CTA holds N/2 rows of the B tile at the same shared-memory offset)
__device__ __forceinline__ void umma_i8_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
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
// cute::UMMA::InstrDescriptor for kind::i8: D=s32, A=u8, B=s8, both K-major; M = 128 (one CTA) or 256 (CTA pair)
__host__ __device__ inline uint32_t make_idesc_i8(int n, int m) {
  return (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ long long gtimer() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#ifdef OSQ_ENABLE_TRACE
#define OSQ_TRACE(slot) do { if (p.trace != nullptr && blockIdx.x == 0) p.trace[(slot)] = clock64(); } while (0)
#else
#define OSQ_TRACE(slot) do { } while (0)
#endif

struct Smem {
  uint64_t a_full[kMaxAStages], a_empty[kMaxAStages];
  uint64_t w_full[kMaxWStages], w_empty[kMaxWStages];
  uint64_t acc_full[kMaxAccStages], acc_empty[kMaxAccStages];
  uint64_t codes_ready, passes_issued, tmem_ready;
  uint64_t x_full[2 * kNumWorkers];  // per-worker fp32 landing slot(s) filled (TMA complete_tx); second set: x_depth = 2
  uint64_t d_full[8], d_empty[8];  // drain epilogue: staging tile (lane quarter q, buffer b) staged by both column slices / drained by both drainers
  uint32_t tmem_base;
  volatile uint32_t converted;  // k-blocks worker 0 has converted so far (paces the L2 prefetcher)
};
// after Smem: per-column epilogue constants of the current chunk (2 * BN floats, padded to 32 columns):
//   y = acc * c1[n] + c0[n],  c1 = s_a * w_scale[n],  c0 = bias[n] - Zc * rowsum[n] * c1
// (measured: an exact integer zero-point correction + magic-number int->float instead of I2F is not faster --
//  the epilogue is bound by shared-memory bandwidth, which the MMA operand reads share -- and a third constant
//  array costs 3 %)
static_assert(sizeof(Smem) % 16 == 0, "constants must stay 16-byte aligned");

// One body, three entry points (below): <registers path>, <TMA landing slots>, <TMA landing slots + CTA pair>.
// Only the pair variant contains cta_group::2 instructions: the driver refuses to launch a kernel that uses them
// without a cluster of two.
// GELU(x) = (x * 0.5) * (1 + erf(x / sqrt(2))) with the operation order of ATen's CUDA kernel (ActivationGeluKernel.cu,
// approximate = 'none'): bit-identical to torch.nn.functional.gelu on the same device
__device__ __forceinline__ float gelu_erf(float x) {
  return __fmul_rn(__fmul_rn(x, 0.5f), __fadd_rn(1.0f, erff(__fmul_rn(x, 0.70710678118654752440f))));
}
// output stage on four adjacent columns: optional activation, then the next quantizer's fake-quant (util_quant.py:11-15 /
// :48-55 exactly as K1 computes it); returns the dequantised values and the four bins (q - qmin) packed in one word
__device__ __forceinline__ float4 out_stage4(float4 o, const QParam& oq, float rinv, float qmin, float qmax, int act, uint32_t& bins) {
  if (act == 1) { o.x = gelu_erf(o.x); o.y = gelu_erf(o.y); o.z = gelu_erf(o.z); o.w = gelu_erf(o.w); }
  float q0, q1, q2, q3;
  bool k0, k1, k2, k3;
  float4 r;
  // division-free fast path (exactly K1's): the group is redone with the IEEE division when any element is near a rounding tie
  r.x = fq_elem_fast(o.x, oq.s, rinv, oq.z, qmin, qmax, q0, k0);
  r.y = fq_elem_fast(o.y, oq.s, rinv, oq.z, qmin, qmax, q1, k1);
  r.z = fq_elem_fast(o.z, oq.s, rinv, oq.z, qmin, qmax, q2, k2);
  r.w = fq_elem_fast(o.w, oq.s, rinv, oq.z, qmin, qmax, q3, k3);
  if (k0 | k1 | k2 | k3) {
    r.x = fq_elem(o.x, oq.s, oq.z, qmin, qmax, q0);
    r.y = fq_elem(o.y, oq.s, oq.z, qmin, qmax, q1);
    r.z = fq_elem(o.z, oq.s, oq.z, qmin, qmax, q2);
    r.w = fq_elem(o.w, oq.s, oq.z, qmin, qmax, q3);
  }
  bins = (uint32_t)__float2int_rn(q0 - qmin) | ((uint32_t)__float2int_rn(q1 - qmin) << 8) | ((uint32_t)__float2int_rn(q2 - qmin) << 16) |
         ((uint32_t)__float2int_rn(q3 - qmin) << 24);
  return r;
}

template <bool kXTma, bool kPair, bool kS3d, bool kEpi = false, bool kDrain = false, int kEW = kNumEpiWarps>
__device__ __forceinline__ void fused_fq_linear_body(const CUtensorMap& tmap_w, const CUtensorMap& tmap_y, const CUtensorMap& tmap_y16,
                                                     const CUtensorMap& tmap_codes, const CUtensorMap& tmap_a, const FusedParams& p,
                                                     uint32_t& tmem_keep) {
  // multi-site launches allocate TMEM once (the allocation permit is relinquished right after) and carry the address from
  // site to site in `tmem_keep`; p.first_site / p.more_sites say where this site sits in the list
  // dynamic shared memory, 1024B aligned by the attribute (SWIZZLE_128B tiles need it):
  // [A ring][W ring][fp32 landing slots][TMA-store staging tiles (may alias the slots)][Smem bookkeeping]
  static_assert(kEW == 8 || (kEW == 16 && !kS3d && !kDrain), "sixteen epilogue warps: plain 32 x 128 B tile stores only");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* a_ring = smem_raw;
  uint8_t* w_ring = a_ring + (size_t)p.a_stages * p.a_stage_bytes;
  uint8_t* x_ring = w_ring + (size_t)p.w_stages * p.w_stage_bytes;  // 16 fp32 landing slots of 4 KB (x_tma only)
  const int x_slots = kNumWorkers * (p.x_depth > 1 ? 2 : 1);
  uint8_t* o_ring = p.alias_xo ? x_ring : x_ring + (kXTma ? x_slots * kXSlotBytes : 0);
  Smem& sm = *reinterpret_cast<Smem*>(p.alias_xo ? x_ring + x_slots * kXSlotBytes
                                                 : o_ring + (size_t)kEW * p.out_bufs * kOutTileBytes);
  // per-column constants; the arrays are padded to a multiple of 32 columns (the epilogue reads 32-column groups)
  const int bn_pad = (p.BN + 31) & ~31;
  float* sm_c1 = reinterpret_cast<float*>(&sm + 1);  // s_a * w_scale[n]
  float* sm_c0 = sm_c1 + bn_pad;                     // bias[n] - Zc * rowsum[n] * c1[n]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr bool pair = kPair;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_y);
    tma_prefetch_desc(&tmap_y16);
    if (p.cached || p.codes_in) tma_prefetch_desc(&tmap_codes);
    if (p.prefetch) tma_prefetch_desc(&tmap_a);
    sm.converted = 0;
  }
  if (warp == 1 && lane == 0) {
    // pair mode: the leader's a_full / acc_empty collect the arrivals of both CTAs (the peer's copies stay unused);
    // a_empty / w_empty / acc_full of each CTA receive the leader's multicast commits
    for (int i = 0; i < p.a_stages; ++i) { mbar_init(&sm.a_full[i], kNumWorkers * p.csz); mbar_init(&sm.a_empty[i], 1); }
    for (int i = 0; i < p.w_stages; ++i) { mbar_init(&sm.w_full[i], 1); mbar_init(&sm.w_empty[i], 1); }
    for (int i = 0; i < p.acc_stages; ++i) { mbar_init(&sm.acc_full[i], 1); mbar_init(&sm.acc_empty[i], kEW * p.csz); }
    mbar_init(&sm.codes_ready, kNumWorkers);
    for (int i = 0; i < 2 * kNumWorkers; ++i) mbar_init(&sm.x_full[i], 1);
    mbar_init(&sm.passes_issued, 1);
    mbar_init(&sm.tmem_ready, 1);
    for (int i = 0; i < 8; ++i) { mbar_init(&sm.d_full[i], 2); mbar_init(&sm.d_empty[i], 2); }
    fence_barrier_init();
  }
  if constexpr (pair) {
    if (warp == 2) {
      if (p.first_site) tmem_alloc_pair(&sm.tmem_base, kTmemCols);
      else if (lane == 0) sm.tmem_base = tmem_keep;
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // the peer's barriers are initialised before any remote arrive / multicast commit
    tc_fence_after();
  } else {
    // only the barrier initialisation gates the producers: the TMEM allocation runs behind it and is published
    // through its own mbarrier to the two consumers of the address (MMA issuer, epilogue warps)
    __syncthreads();
    if (warp == 2) {
      if (p.first_site) tmem_alloc(&sm.tmem_base, kTmemCols);
      else if (lane == 0) sm.tmem_base = tmem_keep;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.tmem_ready);
    }
  }
  auto tmem_address = [&]() -> uint32_t {
    if constexpr (!pair) {
      mbar_wait(&sm.tmem_ready, 0);
      tc_fence_after();
    }
    return *reinterpret_cast<volatile uint32_t*>(&sm.tmem_base);
  };
  const uint32_t crank = (p.csz > 1) ? cluster_ctarank() : 0u;
  if (threadIdx.x == 0) OSQ_TRACE(1020);
  if (p.trace != nullptr && threadIdx.x == 0) p.trace[1024 + 2 * blockIdx.x] = gtimer();

  const int n_my_blocks = p.n_iters;
  // this CTA's place in the (row tile, column split) grid and its chunk range [nc0, nc0 + ncn)
  const int mb_lane = (int)blockIdx.x % p.mb_count;
  const int nc0 = ((int)blockIdx.x / p.mb_count) * p.cps;
  const int ncn = min(p.cps, p.NC - nc0);
  if (p.pdl && threadIdx.x == 0) pdl_launch_dependents();
  const int a_passes = p.resident ? 1 : (ncn + p.cpp - 1) / p.cpp;  // how many times the A ring is filled per m-block

  if (warp == 0) {
    // ===================== TMA producer: packed weight tiles =====================
    if (lane == 0) {
      // the packed weights may have been written by the kernel right before this one (first call after a
      // re-pack): like every other global access they are ordered after the previous grid
      // Only a kernel that calls griddepcontrol.launch_dependents early can still be running here, and no such kernel
      // writes packed weights (osq_pack_weight_s8 never triggers early, so this grid starts after it has completed and
      // flushed): the weight stream does not wait for the previous grid.  OSQ_FUSED_DBG=256 restores the wait.
      if (p.pdl && (p.dbg & 256)) pdl_wait_prior_grids();
      const uint32_t w_bytes = (uint32_t)p.w_stage_bytes;
      const uint32_t w_base = smem_u32(w_ring), full0 = smem_u32(&sm.w_full[0]), empty0 = smem_u32(&sm.w_empty[0]);
      const int slice = p.BN / p.csz;  // rows of the tile this CTA stages (pair: half, the MMA reads both halves)
      const uint32_t leader_full0 = pair ? mapa_u32(full0, 0) : full0;
      uint32_t ws = 0, wph = 0;        // ring stage and its phase parity
      // tile order = the MMA issuer's: resident A: chunk-major; streamed A: per sweep, k-block-major over its chunks
      const int n_tiles = ncn * p.KB;
      for (int it = 0; it < n_my_blocks; ++it)
        for (int t = 0, nc = nc0, kb = 0, c_lo = nc0, j = 0; t < n_tiles; ++t) {
          {
            mbar_wait_u32(empty0 + ws * 8, wph ^ 1);
            if (!pair) {
              mbar_arrive_expect_tx_u32(full0 + ws * 8, w_bytes);
              tma_load_2d_u32(w_base + ws * w_bytes, &tmap_w, full0 + ws * 8, (p.dbg & 1) ? 0 : kb * kStageK, (p.dbg & 1) ? 0 : nc * p.BN);
            } else {
              // the leader's barrier expects both halves; each CTA's TMA completes its bytes there
              if (crank == 0) mbar_arrive_expect_tx_u32(full0 + ws * 8, 2 * w_bytes);
              tma_load_2d_pair_u32(w_base + ws * w_bytes, &tmap_w, leader_full0 + ws * 8, kb * kStageK, nc * p.BN + (int)crank * slice);
            }
#ifdef OSQ_ENABLE_TRACE
            { const int ps = (it * p.NC + nc) * p.KB + kb; if (ps >= 36 && ps < 76) OSQ_TRACE(1600 + ps - 36); }
#endif
            if (++ws == (uint32_t)p.w_stages) { ws = 0; wph ^= 1; }
          }
          if (p.resident) {
            if (++kb == p.KB) { kb = 0; ++nc; }
          } else {
            const int n_c = min(p.cpp, nc0 + ncn - c_lo);
            if (++j == n_c) { j = 0; if (++kb == p.KB) { kb = 0; c_lo += n_c; } }
            nc = c_lo + j;
          }
        }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // One thread; everything it touches is a precomputed 32-bit shared address or a running counter --
    // no divisions, no generic->shared conversions inside the loop.
    if (lane == 0 && crank == 0) {
      const uint32_t idesc = make_idesc_i8(p.BN, pair ? 256 : 128);
      const uint32_t tmem_base = tmem_address();
      const uint32_t a_base = smem_u32(a_ring), w_base = smem_u32(w_ring);
      const uint32_t a_full0 = smem_u32(&sm.a_full[0]), a_empty0 = smem_u32(&sm.a_empty[0]);
      const uint32_t w_full0 = smem_u32(&sm.w_full[0]), w_empty0 = smem_u32(&sm.w_empty[0]);
      const uint32_t acc_full0 = smem_u32(&sm.acc_full[0]), acc_empty0 = smem_u32(&sm.acc_empty[0]);
      const uint64_t desc_hi = make_smem_desc(0);  // everything but the 14-bit start address
      const uint32_t w_bytes = (uint32_t)p.w_stage_bytes;
      uint32_t ws = 0, wph = 0, as_ = 0, aph = 0, st_a = 0, sph_a = 0;
      const uint32_t n_acc = (uint32_t)p.acc_stages;
      for (int it = 0; it < n_my_blocks; ++it) {
        if (p.resident) {
          // resident A: chunk-major, the converted block is filled once per m-block and released after its last chunk
          for (int nc = 0; nc < ncn; ++nc) {
            mbar_wait_u32(acc_empty0 + as_ * 8, aph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + as_ * (uint32_t)p.BN;
            st_a = 0; sph_a = (uint32_t)it & 1;
            const bool wait_a = nc == 0, free_a = nc == ncn - 1;
            for (int kb = 0; kb < p.KB; ++kb) {
              if (wait_a) mbar_wait_u32(a_full0 + st_a * 8, sph_a);
#ifdef OSQ_ENABLE_TRACE
              const int tslot = (it * p.NC + nc) * p.KB + kb;
              if (tslot >= 36 && tslot < 72) OSQ_TRACE(1400 + 3 * (tslot - 36));
#endif
              mbar_wait_u32(w_full0 + ws * 8, wph);
#ifdef OSQ_ENABLE_TRACE
              if (tslot >= 36 && tslot < 72) OSQ_TRACE(1401 + 3 * (tslot - 36));
#endif
              tc_fence_after();
              const uint64_t da = desc_hi | (uint64_t)((a_base + st_a * (uint32_t)p.a_stage_bytes) >> 4);
              const uint64_t db = desc_hi | (uint64_t)((w_base + ws * w_bytes) >> 4);
              if (!pair) {
#pragma unroll
                for (int k = 0; k < kStageK / kUmmaK; ++k)
                  umma_i8(d_tmem, da + (uint64_t)(k * (kUmmaK >> 4)), db + (uint64_t)(k * (kUmmaK >> 4)), idesc, (kb | k) != 0);
                umma_commit_u32(w_empty0 + ws * 8);
                if (free_a) umma_commit_u32(a_empty0 + st_a * 8);
              } else {
#pragma unroll
                for (int k = 0; k < kStageK / kUmmaK; ++k)
                  umma_i8_pair(d_tmem, da + (uint64_t)(k * (kUmmaK >> 4)), db + (uint64_t)(k * (kUmmaK >> 4)), idesc, (kb | k) != 0);
                umma_commit_pair_u32(w_empty0 + ws * 8);
                if (free_a) umma_commit_pair_u32(a_empty0 + st_a * 8);
              }
#ifdef OSQ_ENABLE_TRACE
              if (tslot >= 36 && tslot < 72) OSQ_TRACE(1402 + 3 * (tslot - 36));
#endif
              if (++ws == (uint32_t)p.w_stages) { ws = 0; wph ^= 1; }
              ++st_a;
            }
            if (pair) umma_commit_pair_u32(acc_full0 + as_ * 8); else umma_commit_u32(acc_full0 + as_ * 8);
            if (++as_ == n_acc) { as_ = 0; aph ^= 1; }
          }
        } else {
          // streamed A: every sweep over K feeds `cpp` accumulators at once (k-block-major), so the bins pass
          // through the ring ceil(NC / cpp) times instead of NC times
          for (int c_lo = 0; c_lo < ncn; c_lo += p.cpp) {
            const int n_c = min(p.cpp, ncn - c_lo);
            for (int j = 0; j < n_c; ++j) {  // the sweep's accumulator stages must have been drained
              uint32_t s = as_ + (uint32_t)j, ph = aph;
              if (s >= n_acc) { s -= n_acc; ph ^= 1; }
              mbar_wait_u32(acc_empty0 + s * 8, ph ^ 1);
            }
            tc_fence_after();
            for (int kb = 0; kb < p.KB; ++kb) {
              mbar_wait_u32(a_full0 + st_a * 8, sph_a);
              const uint64_t da = desc_hi | (uint64_t)((a_base + st_a * (uint32_t)p.a_stage_bytes) >> 4);
              for (int j = 0; j < n_c; ++j) {
                mbar_wait_u32(w_full0 + ws * 8, wph);
                tc_fence_after();
                uint32_t s = as_ + (uint32_t)j;
                if (s >= n_acc) s -= n_acc;
                const uint32_t d_tmem = tmem_base + s * (uint32_t)p.BN;
                const uint64_t db = desc_hi | (uint64_t)((w_base + ws * w_bytes) >> 4);
                if (!pair) {
#pragma unroll
                  for (int k = 0; k < kStageK / kUmmaK; ++k)
                    umma_i8(d_tmem, da + (uint64_t)(k * (kUmmaK >> 4)), db + (uint64_t)(k * (kUmmaK >> 4)), idesc, (kb | k) != 0);
                  umma_commit_u32(w_empty0 + ws * 8);
                } else {
#pragma unroll
                  for (int k = 0; k < kStageK / kUmmaK; ++k)
                    umma_i8_pair(d_tmem, da + (uint64_t)(k * (kUmmaK >> 4)), db + (uint64_t)(k * (kUmmaK >> 4)), idesc, (kb | k) != 0);
                  umma_commit_pair_u32(w_empty0 + ws * 8);
                }
                if (++ws == (uint32_t)p.w_stages) { ws = 0; wph ^= 1; }
              }
              if (pair) umma_commit_pair_u32(a_empty0 + st_a * 8); else umma_commit_u32(a_empty0 + st_a * 8);
              if (++st_a == (uint32_t)p.a_stages) { st_a = 0; sph_a ^= 1; }
            }
            for (int j = 0; j < n_c; ++j) {
              if (pair) umma_commit_pair_u32(acc_full0 + as_ * 8); else umma_commit_u32(acc_full0 + as_ * 8);
              if (++as_ == n_acc) { as_ = 0; aph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===================== TMA producer: activation bins =====================
    // code cache (K > 1024): the sweeps after the first re-load the bins the workers spilled;
    // bins-in mode: the upstream fake-quant kernel wrote them, every sweep (resident A: the only one) loads them
    if (lane == 0 && (p.cached || p.codes_in)) {
      if (p.pdl) pdl_wait_prior_grids();
      for (int it = 0; it < n_my_blocks; ++it) {
        const int mb = mb_lane + it * p.mb_count;
        if (!p.codes_in) mbar_wait(&sm.codes_ready, it & 1);  // every worker has published this block's bins
        for (int pass = p.codes_in ? 0 : 1; pass < a_passes; ++pass)
          for (int kb = 0; kb < p.KB; ++kb) {
            const uint32_t pa = (uint32_t)((it * a_passes + pass) * p.KB + kb);
            const int a_st = pa % p.a_stages;
            mbar_wait(&sm.a_empty[a_st], ((pa / p.a_stages) & 1) ^ 1);
            if (!pair) {
              mbar_arrive_expect_tx(&sm.a_full[a_st], p.codes_box_bytes);
              mbar_arrive_n(&sm.a_full[a_st], kNumWorkers - 1);  // a_full always counts kNumWorkers arrivals
              tma_load_2d(a_ring + (size_t)a_st * p.a_stage_bytes, &tmap_codes, &sm.a_full[a_st], kb * kStageK, mb * p.rows_per_tile);
            } else {
              // the leader's a_full counts kNumWorkers arrivals per CTA and the bytes of both CTAs' boxes
              const uint32_t lbar = mapa_u32(smem_u32(&sm.a_full[a_st]), 0);
              mbar_arrive_expect_tx_cluster_u32(lbar, p.codes_box_bytes);
              mbar_arrive_n_cluster_u32(lbar, kNumWorkers - 1);
              tma_load_2d_pair_u32(smem_u32(a_ring + (size_t)a_st * p.a_stage_bytes), &tmap_codes, lbar, kb * kStageK, mb * p.rows_per_tile);
            }
          }
        mbar_arrive(&sm.passes_issued);  // workers may start filling the ring for the next m-block
      }
    }
  } else if (warp == 3) {
    // ===================== L2 prefetcher for the fp32 activation =====================
    // The conversion warps can keep only 8 x 512 B per warp in flight (registers), too little to cover DRAM
    // latency at this CTA's share of the HBM bandwidth.  One thread pulls the tile's [rows x 128] fp32 boxes into
    // L2 a bounded distance ahead (in the order the workers consume them), so their loads see L2 latency.
    if constexpr (kXTma) {
      // Programmatic dependent launch: this CTA starts as soon as its SM is free, up to several microseconds before the last CTA
      // of the previous grid has finished.  Until then nothing may be READ INTO THE SM (A may still be written), but pulling this
      // CTA's fp32 tile into L2 is safe at any time -- L2 is the point of coherence: a line prefetched early is simply updated
      // by a later write -- and turns the first DRAM round trips of the read phase into L2 hits.
      if (lane == 0 && p.pre_l2 && p.pdl && !p.codes_in && mb_lane < p.n_mblocks) {
        const int row0 = mb_lane * p.rows_per_tile;
        const int rows = min(p.rows_per_tile, p.M - row0);
        const int kb_n = min(p.KB, p.pre_l2);
        for (int kb = 0; kb < kb_n; ++kb)
          for (int r = 0; r < rows; r += kRowsPerWorker) tma_prefetch_l2_2d(&tmap_a, kb * kStageK, row0 + r);
      }
    }
    if (lane == 0 && p.prefetch > 0) {
      if (p.pdl) pdl_wait_prior_grids();
      const int passes_per_block = (p.resident || p.cached) ? 1 : ncn;  // fp32 A is re-read per chunk only without a cache
      uint32_t done = 0;
      for (int it = 0; it < n_my_blocks; ++it) {
        const int mb = mb_lane + it * p.mb_count;
        if (mb >= p.n_mblocks) break;
        for (int pass = 0; pass < passes_per_block; ++pass)
          for (int kb = 0; kb < p.KB; ++kb, ++done) {
            while (done >= sm.converted + (uint32_t)p.prefetch) __nanosleep(64);
            tma_prefetch_l2_2d(&tmap_a, kb * kStageK, mb * p.rows_per_tile);
          }
      }
    }
  } else if (warp >= kWorkerWarp0) {
    // ===================== workers: A path (all 16) + epilogue (first 8) =====================
    const int w = warp - kWorkerWarp0;
    // A, the quantisation parameters, Y and the code cache may be produced / still be read by the previous
    // kernel of the stream: every global access is ordered after it; only the prologue above overlaps
    if (p.pdl) pdl_wait_prior_grids();
    if constexpr (kXTma) {
      // the first landing-slot fill of the first tile goes out before the quantisation parameters are even read
      const int row0 = mb_lane * p.rows_per_tile + w * kRowsPerWorker;
      if (lane == 0 && !p.codes_in && w * kRowsPerWorker < p.rows_per_tile && row0 < p.M) {
        const uint32_t bar = smem_u32(&sm.x_full[w]);
        const uint32_t box_bytes = (uint32_t)(p.M < kRowsPerWorker ? p.M : kRowsPerWorker) * (uint32_t)(kStageK * 4);
        mbar_arrive_expect_tx_u32(bar, box_bytes);
        tma_load_2d_u32(smem_u32(x_ring + (size_t)w * kXSlotBytes), &tmap_a, bar, 0, row0);
        if (p.x_depth > 1 && p.KB > 1) {   // second landing slot of this worker: k-block 1
          const uint32_t bar1 = smem_u32(&sm.x_full[w + kNumWorkers]);
          mbar_arrive_expect_tx_u32(bar1, box_bytes);
          tma_load_2d_u32(smem_u32(x_ring + (size_t)(w + kNumWorkers) * kXSlotBytes), &tmap_a, bar1, kStageK, row0);
        }
      }
    }
    // streamed A with a code cache: the bins must survive in L2 until the later sweeps re-load them, so everything that is
    // touched once (fp32 A in, Y out) is marked evict-first.  OSQ_FUSED_DBG=512 turns the hints off.
    const bool hint = ((p.cached && !p.codes_in) || (p.dbg & 1024)) && !(p.dbg & 512);
    // fp32 A is read exactly once in every plan: its loads are evict-first everywhere (+0.5 % on the resident sites; marking Y
    // evict-first as well made single sites 1-4 % faster but the interleaved step and the module chain, whose next kernel
    // reads Y, slower -- so Y keeps the default policy unless the code cache needs the room)
    const bool hint_a = !(p.dbg & 512);
    const uint64_t pol_once = hint_a_any(p) ? l2_policy_evict_first() : 0ull;
    const QParam qp = load_qparam(p.a_scale, p.a_zp, p.a_zp_is_int32, p.g, p.qmin, p.qmax,
                                  blockIdx.x == 0 && w == 0 && lane == 0);
    ConvParam cp;
    cp.s = qp.s;
    cp.rinv = __frcp_rn(qp.s);
    cp.zc = rintf(qp.z) - p.qmin;
    cp.span = p.qmax - p.qmin;
    cp.mz = kMagic + cp.zc;
    cp.lo = kMagic;
    cp.hi = kMagic + cp.span;
    const int r_base = w * kRowsPerWorker;  // this warp's 8 rows of the 128-row block
    const uint32_t lc16 = ((uint32_t)lane >> 2) << 4, lane_in = ((uint32_t)lane & 3) << 2;
    uint32_t cacc = 0, n_stores = 0;
    uint32_t acc_st = 0, acc_ph = 0;  // accumulator stage of the next chunk and its phase parity (running: no divisions per chunk)

    float4 x[kRowsPerWorker];
    // one conversion pass over all k-blocks of m-block `mb`; ring positions start at pa0.
    // Row i of the next k-block is re-issued right after row i of the current one is consumed, so every
    // register stays in flight for a whole iteration (iteration time = max(latency, convert), not the sum).
    // All addressing is strength-reduced to one base pointer per array + compile-time row offsets.
    // A-ring position of the conversion stream (running counters: no divisions in the per-k-block path)
    uint32_t c_st = 0, c_ph = 0;
    auto convert_pass = [&](int mb, uint32_t pa0) {
      const int row_first = mb * p.rows_per_tile + r_base;
      const int nvalid = (r_base < p.rows_per_tile) ? min(kRowsPerWorker, p.M - row_first) : 0;  // warp uniform
      const size_t rs = (size_t)p.K;
      const float* aptr = p.A + ((p.dbg & 4) ? (size_t)0 : (size_t)row_first * rs) + lane * 4;
      uint8_t* cptr = (p.a_codes != nullptr) ? p.a_codes + (size_t)row_first * rs + lane * 4 : nullptr;
      const bool full = nvalid == kRowsPerWorker;
      if (full) {
#pragma unroll
        for (int i = 0; i < kRowsPerWorker; ++i) x[i] = ldg_stream(reinterpret_cast<const float4*>(aptr + i * rs));
      }
      for (int kb = 0; kb < p.KB; ++kb) {
        const uint32_t pa = pa0 + kb;
        const uint32_t a_st = c_st;
        if (pa >= (uint32_t)p.a_stages) mbar_wait(&sm.a_empty[a_st], c_ph ^ 1);  // first fill: ring is empty
        uint8_t* st = a_ring + a_st * (uint32_t)p.a_stage_bytes + r_base * kStageK + lane_in;
        if (full) {
          const bool more = kb + 1 < p.KB;
          const float* nxt = aptr + (size_t)(kb + 1) * kStageK;
          uint32_t risky_mask = 0;
          // four straight-line variants (reload next k-block? write bins to the code cache?) keep every
          // predicate out of the unrolled body: predicated-off instructions still cost issue slots
          auto body = [&](auto reload, auto codes) {
#pragma unroll
            for (int i = 0; i < kRowsPerWorker; ++i) {  // r & 7 == i & 7 because r_base is a multiple of 8
              bool risky;
              uint32_t word;
              if (p.dbg & 32) { word = __float_as_uint(x[i].x) ^ __float_as_uint(x[i].w); risky = false; }
              else word = quant_bin4_fast(x[i], cp, risky);
              risky_mask |= (uint32_t)risky << i;
              if (decltype(reload)::value) x[i] = ldg_stream(reinterpret_cast<const float4*>(nxt + i * rs));
              *reinterpret_cast<uint32_t*>(st + i * kStageK + (lc16 ^ (uint32_t)((i & 7) << 4))) = word;
              if (decltype(codes)::value) *reinterpret_cast<uint32_t*>(cptr + i * rs + (size_t)kb * kStageK) = word;
            }
          };
          using T = std::true_type; using F = std::false_type;
          const bool reload = more && !(p.dbg & 16);
          if (cptr == nullptr) { if (reload) body(T{}, F{}); else body(F{}, F{}); }
          else                 { if (reload) body(T{}, T{}); else body(F{}, T{}); }
          // rare exact fix-up (true division): re-read the float4 (L2 hit) and overwrite its bins
          while (risky_mask != 0) {
            const int i = __ffs(risky_mask) - 1;
            risky_mask &= risky_mask - 1;
            const float4 v = __ldg(reinterpret_cast<const float4*>(aptr + i * rs + (size_t)kb * kStageK));
            const uint32_t word = quant_bin4_exact(v.x, v.y, v.z, v.w, cp.s, cp.zc, cp.span);
            *reinterpret_cast<uint32_t*>(st + i * kStageK + (lc16 ^ (uint32_t)((i & 7) << 4))) = word;
            if (cptr != nullptr) *reinterpret_cast<uint32_t*>(cptr + i * rs + (size_t)kb * kStageK) = word;
          }
        } else {
          // ragged tail of the last tile (or a warp past rows_per_tile): simple predicated path
#pragma unroll 1
          for (int i = 0; i < kRowsPerWorker; ++i) {
            const bool ok = i < nvalid;
            const float4 v = ok ? ldg_stream(reinterpret_cast<const float4*>(aptr + i * rs + (size_t)kb * kStageK))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
            const uint32_t word = quant_bin4(v, cp);
            if (r_base < p.rows_per_tile) *reinterpret_cast<uint32_t*>(st + i * kStageK + (lc16 ^ (uint32_t)((i & 7) << 4))) = word;
            if (ok && cptr != nullptr) *reinterpret_cast<uint32_t*>(cptr + i * rs + (size_t)kb * kStageK) = word;
          }
        }
        if (!(p.dbg & 8)) fence_proxy_async_smem();  // generic-proxy smem stores -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.a_full[a_st]);
        if (w == 0 && lane == 0) sm.converted = sm.converted + 1;
        if (w == 0 && lane == 0 && pa < 250) OSQ_TRACE(pa);
        if (++c_st == (uint32_t)p.a_stages) { c_st = 0; c_ph ^= 1; }
      }
    };
    // The same pass with the fp32 rows arriving by TMA: each worker owns one 4 KB landing slot (its 8 rows x one
    // k-block).  Slot -> registers (8 LDS.128 per lane), then lane 0 immediately re-arms the slot with the next
    // k-block, so 16 x 4 KB stay in flight per SM no matter what the warps are doing, without going through the L1
    // (whose capacity bounds plain loads in flight once 227 KB are carved out as shared memory).  Rows past M are
    // zero-filled by the TMA unit: no ragged path.
    // x_depth = 2 (streamed plans with room for it): two slots per worker, k-block kb lands in slot kb & 1 and is re-armed with
    // kb + 2 -- twice the bytes in flight per SM (the read phase is bound by bytes in flight x DRAM latency, not by the conversion)
    uint32_t x_phv[2] = {0u, 0u};
    const uint32_t leader_a_full0 = pair ? mapa_u32(smem_u32(&sm.a_full[0]), 0) : 0u;
    auto convert_pass_tma = [&](int mb, uint32_t pa0, bool first_issued) {
      const uint32_t x_bar0 = smem_u32(&sm.x_full[w]);
      uint8_t* const x_slot0 = x_ring + (size_t)w * kXSlotBytes;
      const int xd = p.x_depth > 1 ? 2 : 1;
      const uint32_t x_box_bytes = (uint32_t)(p.M < kRowsPerWorker ? p.M : kRowsPerWorker) * (uint32_t)(kStageK * 4);
      const int row_first = mb * p.rows_per_tile + r_base;
      const bool active = r_base < p.rows_per_tile && row_first < p.M;  // warp uniform
      const int nvalid = active ? min(kRowsPerWorker, p.M - row_first) : 0;
      const size_t rs = (size_t)p.K;
      uint8_t* cptr = (p.a_codes != nullptr && active) ? p.a_codes + (size_t)row_first * rs + lane * 4 : nullptr;
      if (active && lane == 0 && !first_issued) {
        mbar_arrive_expect_tx_u32(x_bar0, x_box_bytes);
        tma_load_2d_u32(smem_u32(x_slot0), &tmap_a, x_bar0, 0, row_first);
        if (xd > 1 && p.KB > 1) {
          mbar_arrive_expect_tx_u32(x_bar0 + kNumWorkers * 8, x_box_bytes);
          tma_load_2d_u32(smem_u32(x_slot0 + kNumWorkers * kXSlotBytes), &tmap_a, x_bar0 + kNumWorkers * 8, kStageK, row_first);
        }
      }
      for (int kb = 0; kb < p.KB; ++kb) {
        const uint32_t pa = pa0 + kb;
        const uint32_t a_st = c_st;
        if (pa >= (uint32_t)p.a_stages) mbar_wait(&sm.a_empty[a_st], c_ph ^ 1);  // first fill: ring is empty
        if (active) {
          uint8_t* st = a_ring + a_st * (uint32_t)p.a_stage_bytes + r_base * kStageK + lane_in;
          const int xs = (xd > 1) ? (kb & 1) : 0;
          const uint32_t x_bar = x_bar0 + (uint32_t)xs * (kNumWorkers * 8);
          uint8_t* const x_slot = x_slot0 + (size_t)xs * (kNumWorkers * kXSlotBytes);
          mbar_wait_u32(x_bar, x_phv[xs]);
          x_phv[xs] ^= 1;
#pragma unroll
          for (int i = 0; i < kRowsPerWorker; ++i)
            x[i] = *reinterpret_cast<const float4*>(x_slot + i * (kStageK * 4) + lane * 16);
          if (kb + xd < p.KB) {
            // the slot's contents are in registers: order these generic-proxy reads before the async-proxy refill
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              mbar_arrive_expect_tx_u32(x_bar, x_box_bytes);
              if (hint_a) tma_load_2d_hint_u32(smem_u32(x_slot), &tmap_a, x_bar, (kb + xd) * kStageK, row_first, pol_once);
              else tma_load_2d_u32(smem_u32(x_slot), &tmap_a, x_bar, (kb + xd) * kStageK, row_first);
            }
          }
          uint32_t risky_mask = 0;
          auto body = [&](auto codes) {
#pragma unroll
            for (int i = 0; i < kRowsPerWorker; ++i) {
              bool risky;
              const uint32_t word = quant_bin4_fast(x[i], cp, risky);
              risky_mask |= (uint32_t)risky << i;
              *reinterpret_cast<uint32_t*>(st + i * kStageK + (lc16 ^ (uint32_t)((i & 7) << 4))) = word;
              if (decltype(codes)::value) { if (i < nvalid) *reinterpret_cast<uint32_t*>(cptr + i * rs + (size_t)kb * kStageK) = word; }
            }
          };
          if (cptr == nullptr) body(std::false_type{}); else body(std::true_type{});
          if (risky_mask != 0) {  // rare exact fix-up (true division) of the float4 groups near a rounding tie
#pragma unroll
            for (int i = 0; i < kRowsPerWorker; ++i)
              if (risky_mask & (1u << i)) {
                const uint32_t word = quant_bin4_exact(x[i].x, x[i].y, x[i].z, x[i].w, cp.s, cp.zc, cp.span);
                *reinterpret_cast<uint32_t*>(st + i * kStageK + (lc16 ^ (uint32_t)((i & 7) << 4))) = word;
                if (cptr != nullptr && i < nvalid) *reinterpret_cast<uint32_t*>(cptr + i * rs + (size_t)kb * kStageK) = word;
              }
          }
        }
        fence_proxy_async_smem();  // generic-proxy smem stores -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) { if (pair) mbar_arrive_cluster_u32(leader_a_full0 + a_st * 8); else mbar_arrive(&sm.a_full[a_st]); }
        if (w == 0 && lane == 0 && pa < 250) OSQ_TRACE(pa);
        if (++c_st == (uint32_t)p.a_stages) { c_st = 0; c_ph ^= 1; }
      }
    };
    // ring uses that are filled by the TMA thread (cached passes) advance the same counters
    auto skip_ring = [&](uint32_t n) {
      const uint32_t tot = c_st + n;
      c_ph ^= (tot / (uint32_t)p.a_stages) & 1u;
      c_st = tot % (uint32_t)p.a_stages;
    };

    // ---- epilogue (workers 0..7): TMEM -> registers (thread = row) -> y = acc*c1 + c0 -> swizzled staging
    //      tile -> TMA store.  The TMA unit writes full 128-byte lines and bypasses the (1 KB) L1.
    const int q = w & 3;                 // TMEM lane quarter (= warp % 4, a hardware rule)
    const int slice = w >> 2;            // column slice of the chunk (epilogue warps: 0 or 1)
    const float s_a = qp.s;
    const float zcf = cp.zc;
    QParam oq = QParam{1.f, 0.f};
    if constexpr (kEpi) {
      if (w < kEW)
        oq = load_qparam(p.out_scale, p.out_zp, p.out_zp_is_int32, p.out_g, p.out_qmin, p.out_qmax,
                         blockIdx.x == 0 && w == 0 && lane == 0);
    }
    const float oq_rinv = __frcp_rn(oq.s);
    uint8_t* my_tiles = o_ring + (size_t)w * p.out_bufs * kOutTileBytes;
    const uint32_t sw = ((uint32_t)lane & 7) << 4;  // 128B swizzle phase of this thread's staging row
    const int et = threadIdx.x - kWorkerWarp0 * 32;  // 0..255 among the epilogue threads
    // per-column constants: each epilogue thread owns column `et` of the chunk (BN <= 256 = epilogue threads);
    // they are fetched into registers early (L2 latency hides under the previous chunk) and published to
    // shared memory between two barriers
    // The three loads are issued early and only CONSUMED at publish time (one chunk later): using them right away
    // would park the warp on an L2 round trip (1-2 K cycles under store traffic) in front of every chunk.
    float raw_ws = 0.f, raw_b = 0.f;
    int raw_rs = 0;
    auto fetch_consts = [&](int nc) {
      const int n = nc * p.BN + et;
      raw_ws = 0.f; raw_b = 0.f; raw_rs = 0;
      if (et < p.BN && n < p.N) {
        raw_ws = __ldg(p.w_scale + n);
        raw_rs = __ldg(p.w_rowsum + n);
        if (p.bias != nullptr) raw_b = __ldg(p.bias + n);
      }
    };
    auto publish_consts = [&]() {  // all epilogue warps; previous chunk's readers are done (barrier before)
      if (et < p.BN) {
        const float pc1 = __fmul_rn(s_a, raw_ws);
        sm_c1[et] = pc1;
        sm_c0[et] = fmaf(-zcf * (float)raw_rs, pc1, raw_b);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kEW * 32) : "memory");
    };
    int obufs = p.out_bufs;  // store tiles this warp may cycle through in the current m-block
    uint32_t tmem_base = 0;   // fetched right before the first epilogue chunk
    auto epilogue_chunk = [&](int mb, int nc) {
      const int as_ = (int)acc_st;
      const int n0 = nc * p.BN;
      if (nc + 1 < nc0 + ncn) fetch_consts(nc + 1);
      mbar_wait(&sm.acc_full[as_], acc_ph);
      tc_fence_after();
      if (w == 0 && lane == 0 && cacc < 60) OSQ_TRACE(512 + cacc * 4);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as_ * p.BN);
      const int row0 = mb * p.rows_per_tile + q * 32;
      // the two column slices of a lane quarter take alternating 32-column groups, so a row's two 128-byte pieces
      // leave the SM close together in time (measured +6 % on the store path, scripts/mb/storebench.cu)
      const int n_cols = min(p.BN, p.N - n0);
      // rows_per_tile is a multiple of 16: the tile's last lane quarter may own only 16 rows, which go out
      // through the 16-row box so that the neighbouring tile's rows are never touched
      const int rows_q = min(32, p.rows_per_tile - q * 32);
      const bool any_rows = (rows_q > 0) && (row0 < p.M);
      const CUtensorMap* ymap = (rows_q == 32) ? &tmap_y : &tmap_y16;
      if constexpr (kDrain) {
        // Drain epilogue: the store engine is the SM's own LSU.  The two column slices of lane quarter q stage the two 128-byte
        // halves of a [32 rows][256 B] tile (two tiles per quarter, the landing slots); the two drainer warps of the quarter
        // (workers 8..15, idle since the conversion ended) re-read it two rows at a time and write 256 contiguous bytes per row
        // with plain 128-bit stores.  Measured in isolation (scripts/mb/storebench.cu): 21.3 B/clk/SM against 17.8-18.9 for
        // the 32 x 128 B tensor stores.
        for (int c0 = slice * 32; c0 < n_cols; c0 += 64) {
          uint32_t v[32];
          tmem_ld32(taddr + c0, v);
          tmem_ld_wait();
          if (!any_rows) continue;  // uniform over the quarter's producers and drainers
          const uint32_t tb = (uint32_t)q * 2u + (n_stores & 1u);
          if (n_stores >= 2u) mbar_wait(&sm.d_empty[tb], ((n_stores >> 1) - 1u) & 1u);
          uint8_t* trow = x_ring + (size_t)tb * (2 * kXSlotBytes) + lane * 256 + slice * 128;
          float4 c1n = *reinterpret_cast<const float4*>(sm_c1 + c0);
          float4 k0n = *reinterpret_cast<const float4*>(sm_c0 + c0);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 c1 = c1n, k0 = k0n;
            if (j + 4 < 32) {
              c1n = *reinterpret_cast<const float4*>(sm_c1 + c0 + j + 4);
              k0n = *reinterpret_cast<const float4*>(sm_c0 + c0 + j + 4);
            }
            float4 o;
            o.x = fmaf((float)(int)v[j + 0], c1.x, k0.x);
            o.y = fmaf((float)(int)v[j + 1], c1.y, k0.y);
            o.z = fmaf((float)(int)v[j + 2], c1.z, k0.z);
            o.w = fmaf((float)(int)v[j + 3], c1.w, k0.w);
            *reinterpret_cast<float4*>(trow + ((uint32_t)(j << 2) ^ sw)) = o;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.d_full[tb]);
          ++n_stores;
        }
      } else
      if constexpr (kS3d) {
        // Pair store: at every step the two slices of this lane quarter cover 64 adjacent columns.  Both write their 32 x 128 B
        // half into ONE 8 KB staging tile laid out [32 rows][2][128 B] (the box of the 3-D map [M][N/32][32]); slice 0 then
        // issues a single tensor store that writes 256 contiguous bytes per row (measured +10-12 % over 128-byte rows,
        // profiles/r01_storebench_3d.txt).  Two tiles per quarter (the landing slots of both warps): store i overlaps step
        // i + 1.  One 64-thread named barrier per step couples the two warps.
        const uint32_t unit = (uint32_t)lane * 2u + (uint32_t)slice;      // 128-byte unit of this thread's row half
        const uint32_t usw = unit & 7u;                                  // SWIZZLE_128B phase of that unit
        const uint32_t bit = ((uint32_t)lane >> 2) & 1u;                 // lanes l and l + 4 share usw: they swap even / odd chunks
        const bool issuer = slice == 0;
        for (int cb = 0; cb < n_cols; cb += 64) {
          const int c0 = cb + slice * 32;
          const bool active = c0 < n_cols;                               // last chunk of an N that is an odd multiple of 32
          uint32_t v[32];
          tmem_ld32(taddr + c0, v);   // (an inactive slice reads columns of the stage nobody stores: harmless)
          tmem_ld_wait();
          if (!any_rows) continue;  // uniform over both warps of the quarter
          const uint32_t kbuf = n_stores & (uint32_t)(obufs - 1);
          uint8_t* tile = p.alias_xo ? x_ring + (size_t)(kbuf * kEW + 2u * (uint32_t)q) * kXSlotBytes
                                     : o_ring + (size_t)(kbuf * kEW + 2u * (uint32_t)q) * kOutTileBytes;
          if (obufs == 1) {  // single buffer: its previous store must have been read before anyone writes
            if (issuer && lane == 0) tma_store_wait_read<0>();
            __syncwarp();
            pair_barrier(q);
          } else if (p.dbg & 64) {  // experiment: free THIS tile at the top (store i - 2), leave store i - 1 in flight
            if (issuer && lane == 0) tma_store_wait_read<1>();
            __syncwarp();
            pair_barrier(q);
          }
          if (active) {
            uint8_t* trow = tile + unit * 128u;
#ifndef OSQ_S3D_NOSWAP
            // per 8 columns: first the 4-column group (bit), then its neighbour, so that the eight lanes of a store phase
            // (lanes l and l + 4 share the swizzle phase) hit eight different bank groups.  The group's constants are
            // fetched with a lane-dependent address instead of being selected in registers.
            const float* pc1 = sm_c1 + c0 + 4 * (int)bit;
            const float* pk0 = sm_c0 + c0 + 4 * (int)bit;
            const int flip = 4 - 8 * (int)bit;  // offset from the first group to the second: +4 or -4 columns
            float4 c1n = *reinterpret_cast<const float4*>(pc1);
            float4 k0n = *reinterpret_cast<const float4*>(pk0);
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float4 c1 = c1n, k0 = k0n;
              c1n = *reinterpret_cast<const float4*>(pc1 + j + flip);
              k0n = *reinterpret_cast<const float4*>(pk0 + j + flip);
              const uint32_t ce = (uint32_t)(j >> 2);
              float4 o;
              o.x = fmaf((float)(int)(bit ? v[j + 4] : v[j + 0]), c1.x, k0.x);
              o.y = fmaf((float)(int)(bit ? v[j + 5] : v[j + 1]), c1.y, k0.y);
              o.z = fmaf((float)(int)(bit ? v[j + 6] : v[j + 2]), c1.z, k0.z);
              o.w = fmaf((float)(int)(bit ? v[j + 7] : v[j + 3]), c1.w, k0.w);
              *reinterpret_cast<float4*>(trow + (((ce ^ bit) ^ usw) << 4)) = o;
              c1 = c1n; k0 = k0n;
              if (j + 8 < 32) {
                c1n = *reinterpret_cast<const float4*>(pc1 + j + 8);
                k0n = *reinterpret_cast<const float4*>(pk0 + j + 8);
              }
              o.x = fmaf((float)(int)(bit ? v[j + 0] : v[j + 4]), c1.x, k0.x);
              o.y = fmaf((float)(int)(bit ? v[j + 1] : v[j + 5]), c1.y, k0.y);
              o.z = fmaf((float)(int)(bit ? v[j + 2] : v[j + 6]), c1.z, k0.z);
              o.w = fmaf((float)(int)(bit ? v[j + 3] : v[j + 7]), c1.w, k0.w);
              *reinterpret_cast<float4*>(trow + ((((ce + 1u) ^ bit) ^ usw) << 4)) = o;
            }
#else
            // (lanes l and l + 4 share the swizzle phase of their units: a 2-way bank conflict per store phase)
            (void)bit;
            float4 c1n = *reinterpret_cast<const float4*>(sm_c1 + c0);
            float4 k0n = *reinterpret_cast<const float4*>(sm_c0 + c0);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 c1 = c1n, k0 = k0n;
              if (j + 4 < 32) {
                c1n = *reinterpret_cast<const float4*>(sm_c1 + c0 + j + 4);
                k0n = *reinterpret_cast<const float4*>(sm_c0 + c0 + j + 4);
              }
              float4 o;
              o.x = fmaf((float)(int)v[j + 0], c1.x, k0.x);
              o.y = fmaf((float)(int)v[j + 1], c1.y, k0.y);
              o.z = fmaf((float)(int)v[j + 2], c1.z, k0.z);
              o.w = fmaf((float)(int)v[j + 3], c1.w, k0.w);
              *reinterpret_cast<float4*>(trow + ((uint32_t)(j << 2) ^ (usw << 4))) = o;
            }
#endif
          }
          fence_proxy_async_smem();
          if (obufs == 2 && issuer && !(p.dbg & 64)) {  // frees the OTHER tile for the next step (its store was issued a whole step ago)
            if (lane == 0) tma_store_wait_read<0>();
            __syncwarp();
          }
          pair_barrier(q);
          if (issuer && lane == 0 && !(p.dbg & 2)) {
            tma_store_3d(ymap, tile, 0, (n0 + cb) >> 5, row0);   // rows >= M / column groups >= N/32 are clipped by the TMA unit
            tma_store_commit();
          }
          ++n_stores;
        }
      } else {
      for (int c0 = slice * 32; c0 < n_cols; c0 += (kEW / 4) * 32) {
        uint32_t v[32];
#ifdef OSQ_ENABLE_TRACE
        const bool probe = (w == 0 && lane == 0 && cacc == 2);
        const int pslot = 900 + (c0 >> 6) * 6;
        if (probe) OSQ_TRACE(pslot + 0);
#endif
        tmem_ld32(taddr + c0, v);
        tmem_ld_wait();
#ifdef OSQ_ENABLE_TRACE
        if (probe) OSQ_TRACE(pslot + 1);
#endif
        if (!any_rows) continue;  // warp uniform: phantom tile / quarter past the tile's rows
        if (n_stores >= (uint32_t)obufs) {  // the staging tile must have been read out by its previous TMA store
          if (lane == 0) { if (obufs == 2 && p.lsu_mod == 0) tma_store_wait_read<1>(); else tma_store_wait_read<0>(); }  // (mixed engines: bulk groups no longer map to tiles)
          __syncwarp();
        }
#ifdef OSQ_ENABLE_TRACE
        if (probe) OSQ_TRACE(pslot + 2);
#endif
        // aliased mode: this warp's own landing slot, plus (last tile only) the slot of worker w + 8, whose owner is
        // through with it once the tile's last MMA has been committed
        uint8_t* tile = p.alias_xo ? x_ring + (size_t)((kEW == 16 ? 0u : ((n_stores & 1u) & (uint32_t)(obufs - 1)) * 8u) + (uint32_t)w) * kXSlotBytes
                                   : my_tiles + (size_t)(n_stores % (uint32_t)obufs) * kOutTileBytes;
        uint8_t* trow = tile + lane * 128;
        // The constants of the NEXT four columns are loaded before the current four are stored: the compiler
        // cannot move a shared-memory load above an earlier shared-memory store (possible alias), so without this
        // every iteration would expose a full LDS latency.
        // the constants of the NEXT four columns are loaded before the current four are stored: the compiler cannot
        // move a shared-memory load above an earlier shared-memory store (possible alias)
        {
          float4 c1n = *reinterpret_cast<const float4*>(sm_c1 + c0);
          float4 k0n = *reinterpret_cast<const float4*>(sm_c0 + c0);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 c1 = c1n, k0 = k0n;
            if (j + 4 < 32) {
              c1n = *reinterpret_cast<const float4*>(sm_c1 + c0 + j + 4);
              k0n = *reinterpret_cast<const float4*>(sm_c0 + c0 + j + 4);
            }
            float4 o;
            o.x = fmaf((float)(int)v[j + 0], c1.x, k0.x);
            o.y = fmaf((float)(int)v[j + 1], c1.y, k0.y);
            o.z = fmaf((float)(int)v[j + 2], c1.z, k0.z);
            o.w = fmaf((float)(int)v[j + 3], c1.w, k0.w);
            if constexpr (kEpi) {
              uint32_t bw;
              o = out_stage4(o, oq, oq_rinv, p.out_qmin, p.out_qmax, p.out_act, bw);
              v[j >> 2] = bw;   // v[j .. j + 3] are consumed: the accumulator registers double as the bins' staging
            }
            *reinterpret_cast<float4*>(trow + ((uint32_t)(j << 2) ^ sw)) = o;  // 16B chunk (j/4) ^ (row & 7): conflict free
          }
          if constexpr (kEpi) {
            // this thread's row, 32 bins = 32 contiguous bytes = one full sector: straight from registers
            const int row = row0 + lane;
            if (p.out_bins != nullptr && lane < rows_q && row < p.M) {
              uint4* dst = reinterpret_cast<uint4*>(p.out_bins + (size_t)row * (size_t)p.N + (size_t)(n0 + c0));
              dst[0] = make_uint4(v[0], v[1], v[2], v[3]);
              if (n0 + c0 + 32 <= p.N) dst[1] = make_uint4(v[4], v[5], v[6], v[7]);   // N % 32 == 16: the last group is half
            }
          }
        }
#ifdef OSQ_ENABLE_TRACE
        if (probe) OSQ_TRACE(pslot + 3);
#endif
        // Two store engines: the TMA unit sustains ~18 B/clk/SM on these 32 x 128 B boxes, less than the chip can write.  Every
        // `lsu_mod`-th step of a warp therefore leaves through the LSU instead: the warp re-reads its tile four rows at a time and
        // writes 4 x 128 contiguous bytes per 128-bit store instruction.  (kEpi keeps the TMA path: its registers hold the bins.)
        const bool via_lsu = !kEpi && p.lsu_mod > 0 && (n_stores % (uint32_t)p.lsu_mod) == (uint32_t)p.lsu_mod - 1u;
        if (via_lsu) {
          __syncwarp();
          const int cvalid = min(32, n_cols - c0);         // N % 32 == 16: the last group is half
          const int ch = lane & 7;
          float* yrow = p.Y + (size_t)(row0 + (lane >> 3)) * (size_t)p.N + (size_t)(n0 + c0 + ch * 4);
          const int r_lim = min(rows_q, p.M - row0);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = 4 * i + (lane >> 3);
            const float4 o = *reinterpret_cast<const float4*>(tile + r * 128 + (((uint32_t)ch ^ ((uint32_t)r & 7u)) << 4));
            if (r < r_lim && ch * 4 < cvalid && !(p.dbg & 2)) {
              float4* dst = reinterpret_cast<float4*>(yrow + (size_t)(4 * i) * (size_t)p.N);
              if (hint) stg_hint(dst, o, pol_once); else *dst = o;
            }
          }
          __syncwarp();   // the tile is in registers: the next step may overwrite it
        } else {
        fence_proxy_async_smem();
        __syncwarp();
#ifdef OSQ_ENABLE_TRACE
        if (probe) OSQ_TRACE(pslot + 4);
#endif
        if (lane == 0 && !(p.dbg & 2)) {
          if (hint) tma_store_2d_hint(ymap, tile, n0 + c0, row0, pol_once);
          else tma_store_2d(ymap, tile, n0 + c0, row0);  // rows >= M / columns >= N are clipped by the TMA unit
          tma_store_commit();
        }
        }
#ifdef OSQ_ENABLE_TRACE
        if (probe) OSQ_TRACE(pslot + 5);
#endif
        ++n_stores;
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (pair) mbar_arrive_cluster_u32(mapa_u32(smem_u32(&sm.acc_empty[as_]), 0)); else mbar_arrive(&sm.acc_empty[as_]); }
      if (w == 0 && lane == 0 && cacc < 60) OSQ_TRACE(512 + cacc * 4 + 1);
      ++cacc;
      if (++acc_st == (uint32_t)p.acc_stages) { acc_st = 0; acc_ph ^= 1; }
      asm volatile("bar.sync 1, %0;" ::"n"(kEW * 32) : "memory");  // every reader of this chunk's constants is done
      if (w == 0 && lane == 0 && cacc <= 60) OSQ_TRACE(512 + (cacc - 1) * 4 + 2);
      if (nc + 1 < nc0 + ncn) publish_consts();
      if (w == 0 && lane == 0 && cacc <= 60) OSQ_TRACE(512 + (cacc - 1) * 4 + 3);
    };

    // ---- drain epilogue, consumer side (workers 8..15): quarter dq, row half dh of every staged tile
    auto drain_chunk = [&](int mb, int nc) {
      const int dq = (w - kEW) & 3, dh = (w - kEW) >> 2;
      const int n0 = nc * p.BN;
      const int n_cols = min(p.BN, p.N - n0);
      const int row0 = mb * p.rows_per_tile + dq * 32;
      const int rows_q = min(min(32, p.rows_per_tile - dq * 32), p.M - row0);   // rows of this quarter that exist
      if (rows_q <= 0) return;                                                  // the producers skip it as well
      const int ch = lane & 15;
      const int r_first = dh * 16 + (lane >> 4);
      float* ybase = p.Y + (size_t)(row0 + r_first) * (size_t)p.N + (size_t)n0 + (size_t)(ch * 4);
      const uint32_t toff = (uint32_t)(ch >> 3) * 128u;
      for (int cb = 0; cb < n_cols; cb += 64) {
        const uint32_t tb = (uint32_t)dq * 2u + (n_stores & 1u);
        mbar_wait(&sm.d_full[tb], (n_stores >> 1) & 1u);
        const uint8_t* tile = x_ring + (size_t)tb * (2 * kXSlotBytes);
        float4 o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = r_first + 2 * i;
          o[i] = *reinterpret_cast<const float4*>(tile + r * 256 + toff + ((((uint32_t)ch & 7u) ^ ((uint32_t)r & 7u)) << 4));
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.d_empty[tb]);   // the tile is in registers: the producers may refill it
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = r_first + 2 * i;
          if (r < rows_q && !(p.dbg & 2)) {
            float4* dst = reinterpret_cast<float4*>(ybase + (size_t)(2 * i) * (size_t)p.N + cb);
            if (hint) stg_hint(dst, o[i], pol_once); else *dst = o[i];
          }
        }
        ++n_stores;
      }
    };

    for (int it = 0; it < n_my_blocks; ++it) {
      const int mb = mb_lane + it * p.mb_count;
      const uint32_t pa_block = (uint32_t)it * (uint32_t)(a_passes * p.KB);
      // cached mode: the TMA thread must have issued every re-load pass of the previous block before this
      // warp runs ahead on the same ring (two producers may never be more than one ring cycle apart)
      if (w < kEW) fetch_consts(nc0);
      if (!p.codes_in) {
      if (p.cached && it > 0) mbar_wait(&sm.passes_issued, (it - 1) & 1);
      if (!kDrain && p.alias_xo && it > 0 && w < kEW) {  // this warp's landing slot was its store tile: reads must be done
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
        // pair stores: the slot may have been staged by the other slice's warp and stored by slice 0's
        if constexpr (kS3d) asm volatile("bar.sync 1, %0;" ::"n"(kEW * 32) : "memory");
      }
      if constexpr (kXTma) convert_pass_tma(mb, pa_block, it == 0); else convert_pass(mb, pa_block);
      }
      if (p.cached && !p.codes_in) {
        // bins of this m-block are in the code cache: publish them to the async proxy (TMA) of this CTA
        __threadfence();
        fence_proxy_async_all();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.codes_ready);
        skip_ring((uint32_t)(a_passes - 1) * (uint32_t)p.KB);  // N chunks >= 1 are filled by the TMA thread
      }
      if (!kDrain && p.alias_xo && !p.codes_in) { obufs = (it == n_my_blocks - 1 && kEW == 8) ? p.out_bufs : 1; n_stores = 0; }
      if (kEW == 16) obufs = 1;   // every worker stages in its own landing slot
      if (w < kEW && it == 0) tmem_base = tmem_address();
      if (w < kEW) publish_consts();  // chunk 0 constants
      for (int lc = 0; lc < ncn; ++lc) {
        // no code cache and K too large for residency: re-convert A for the next N chunk first
        if (!p.resident && !p.cached && !p.codes_in && lc + 1 < ncn) {
          if constexpr (kXTma) convert_pass_tma(mb, pa_block + (uint32_t)(lc + 1) * p.KB, false); else convert_pass(mb, pa_block + (uint32_t)(lc + 1) * p.KB);
        }
        if (w < kEW) epilogue_chunk(mb, nc0 + lc);
        else if constexpr (kDrain) drain_chunk(mb, nc0 + lc);
      }
      // drain epilogue: the staging tiles are the landing slots of ALL workers; the next tile's fills wait for the last drain
      if constexpr (kDrain) { if (it + 1 < n_my_blocks) asm volatile("bar.sync 6, %0;" ::"n"(kNumWorkers * 32) : "memory"); }
    }
    // the staging tiles must have been READ before the CTA's shared memory goes away; the writes themselves are complete
    // and visible at grid completion (CUTLASS' TMA epilogues end on the same .read wait).  OSQ_FUSED_DBG=128: full wait.
    if (!kDrain && w < kEW && lane == 0) { if (p.dbg & 128) tma_store_wait_all(); else tma_store_wait_read<0>(); }
    if (w == 0 && lane == 0) OSQ_TRACE(1021);
  }

  // ===================== teardown =====================
  tc_fence_before();
  __syncthreads();
  if (p.csz > 1) cluster_sync_all();  // no CTA exits while the pair's MMAs / commits / remote arrives may still touch it
  if (warp == 2 && !p.more_sites) {
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&sm.tmem_base);
    if (pair) tmem_dealloc_pair(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
  }
  if (p.trace != nullptr && threadIdx.x == 0) p.trace[1024 + 2 * blockIdx.x + 1] = gtimer();
  if (p.more_sites) {
    tmem_keep = *reinterpret_cast<volatile uint32_t*>(&sm.tmem_base);   // every thread: the next site's Smem sits elsewhere
    // multi-site launch: the next site re-initialises every barrier at the same addresses (possibly with other counts)
    if (warp == 1 && lane == 0) {
      for (int i = 0; i < p.a_stages; ++i) { mbar_inval(&sm.a_full[i]); mbar_inval(&sm.a_empty[i]); }
      for (int i = 0; i < p.w_stages; ++i) { mbar_inval(&sm.w_full[i]); mbar_inval(&sm.w_empty[i]); }
      for (int i = 0; i < p.acc_stages; ++i) { mbar_inval(&sm.acc_full[i]); mbar_inval(&sm.acc_empty[i]); }
      mbar_inval(&sm.codes_ready);
      for (int i = 0; i < 2 * kNumWorkers; ++i) mbar_inval(&sm.x_full[i]);
      mbar_inval(&sm.passes_issued);
      mbar_inval(&sm.tmem_ready);
      for (int i = 0; i < 8; ++i) { mbar_inval(&sm.d_full[i]); mbar_inval(&sm.d_empty[i]); }
    }
    __syncthreads();   // also: TMEM is deallocated and every warp has left this site's shared memory
    if (p.csz > 1) cluster_sync_all();
  }
}

#define OSQ_FUSED_ENTRY(name, XTMA, PAIR, S3D)                                                                         \
  __global__ void __launch_bounds__(kNumThreads, 1)                                                                    \
  name(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_y,                         \
       const __grid_constant__ CUtensorMap tmap_y16, const __grid_constant__ CUtensorMap tmap_codes,                   \
       const __grid_constant__ CUtensorMap tmap_a, const FusedParams p) {                                              \
    uint32_t tmem_keep = 0;                                                                                            \
    fused_fq_linear_body<XTMA, PAIR, S3D>(tmap_w, tmap_y, tmap_y16, tmap_codes, tmap_a, p, tmem_keep);                 \
  }
OSQ_FUSED_ENTRY(fused_fq_linear_kernel_ldg, false, false, false)
OSQ_FUSED_ENTRY(fused_fq_linear_kernel, true, false, false)
OSQ_FUSED_ENTRY(fused_fq_linear_kernel_pair, true, true, false)
// all sixteen workers run the epilogue (four column slices per lane quarter, one staging tile = the warp's own landing slot)
#define OSQ_FUSED_ENTRY_E16(name, PAIR)                                                                                \
  __global__ void __launch_bounds__(kNumThreads, 1)                                                                    \
  name(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_y,                         \
       const __grid_constant__ CUtensorMap tmap_y16, const __grid_constant__ CUtensorMap tmap_codes,                   \
       const __grid_constant__ CUtensorMap tmap_a, const FusedParams p) {                                              \
    uint32_t tmem_keep = 0;                                                                                            \
    fused_fq_linear_body<true, PAIR, false, false, false, 16>(tmap_w, tmap_y, tmap_y16, tmap_codes, tmap_a, p, tmem_keep); \
  }
OSQ_FUSED_ENTRY_E16(fused_fq_linear_kernel_e16, false)
OSQ_FUSED_ENTRY_E16(fused_fq_linear_kernel_pair_e16, true)
// drain epilogue (plain 128-bit stores of staged 256-byte rows by the idle workers)
#define OSQ_FUSED_ENTRY_DRAIN(name, PAIR)                                                                              \
  __global__ void __launch_bounds__(kNumThreads, 1)                                                                    \
  name(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_y,                         \
       const __grid_constant__ CUtensorMap tmap_y16, const __grid_constant__ CUtensorMap tmap_codes,                   \
       const __grid_constant__ CUtensorMap tmap_a, const FusedParams p) {                                              \
    uint32_t tmem_keep = 0;                                                                                            \
    fused_fq_linear_body<true, PAIR, false, false, true>(tmap_w, tmap_y, tmap_y16, tmap_codes, tmap_a, p, tmem_keep);  \
  }
OSQ_FUSED_ENTRY_DRAIN(fused_fq_linear_kernel_dr, false)
OSQ_FUSED_ENTRY_DRAIN(fused_fq_linear_kernel_pair_dr, true)
// epilogue with pair-shared staging tiles and 256-byte-row 3-D tensor stores (N % 32 == 0, chunks of 64 k columns)
OSQ_FUSED_ENTRY(fused_fq_linear_kernel_ldg_s3, false, false, true)
OSQ_FUSED_ENTRY(fused_fq_linear_kernel_s3, true, false, true)
OSQ_FUSED_ENTRY(fused_fq_linear_kernel_pair_s3, true, true, true)
// output stage: Y leaves as fq(act(Y)) of the next activation quantizer, plus its u8 bins
#define OSQ_FUSED_ENTRY_EPI(name, PAIR)                                                                                \
  __global__ void __launch_bounds__(kNumThreads, 1)                                                                    \
  name(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_y,                         \
       const __grid_constant__ CUtensorMap tmap_y16, const __grid_constant__ CUtensorMap tmap_codes,                   \
       const __grid_constant__ CUtensorMap tmap_a, const FusedParams p) {                                              \
    uint32_t tmem_keep = 0;                                                                                            \
    fused_fq_linear_body<true, PAIR, false, true>(tmap_w, tmap_y, tmap_y16, tmap_codes, tmap_a, p, tmem_keep);         \
  }
OSQ_FUSED_ENTRY_EPI(fused_fq_linear_kernel_epi, false)
OSQ_FUSED_ENTRY_EPI(fused_fq_linear_kernel_pair_epi, true)
// the output stage is ALU-bound (erf, exact division): with all sixteen workers in the epilogue it runs twice as fast
#define OSQ_FUSED_ENTRY_EPI16(name, PAIR)                                                                              \
  __global__ void __launch_bounds__(kNumThreads, 1)                                                                    \
  name(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_y,                         \
       const __grid_constant__ CUtensorMap tmap_y16, const __grid_constant__ CUtensorMap tmap_codes,                   \
       const __grid_constant__ CUtensorMap tmap_a, const FusedParams p) {                                              \
    uint32_t tmem_keep = 0;                                                                                            \
    fused_fq_linear_body<true, PAIR, false, true, false, 16>(tmap_w, tmap_y, tmap_y16, tmap_codes, tmap_a, p, tmem_keep); \
  }
OSQ_FUSED_ENTRY_EPI16(fused_fq_linear_kernel_epi_e16, false)
OSQ_FUSED_ENTRY_EPI16(fused_fq_linear_kernel_pair_epi_e16, true)

// Multi-site launch: one persistent grid walks a short list of INDEPENDENT sites (no site's input is another site's
// output).  A CTA moves on to the next site as soon as its own tiles of the current one are stored, so the launch gap,
// the wait for the slowest CTA and the first-load latency are paid once per list instead of once per site, and the tail of
// one site's write phase overlaps the head of the next site's read phase.
constexpr int kMaxSites = 4;
struct FusedSite {
  CUtensorMap w, y, y16, codes, a;
  FusedParams p;
};
struct FusedSites {
  FusedSite s[kMaxSites];
  int n;
};
// One inlined copy of the body per list position: every copy reads its site's parameters at compile-time constant-bank
// offsets, exactly like a single-site launch (a runtime-indexed site costs registers in every hot loop: measured 10 % slower).
#define OSQ_FUSED_SITE(i)                                                                                              \
  if (sites.n > (i))                                                                                                   \
    fused_fq_linear_body<true, PAIR_, false>(sites.s[i].w, sites.s[i].y, sites.s[i].y16, sites.s[i].codes,             \
                                             sites.s[i].a, sites.s[i].p, tmem_keep);
#define OSQ_FUSED_MULTI_ENTRY(name, PAIR)                                                                              \
  __global__ void __launch_bounds__(kNumThreads, 1) name(const __grid_constant__ FusedSites sites) {                   \
    constexpr bool PAIR_ = PAIR;                                                                                       \
    uint32_t tmem_keep = 0;                                                                                            \
    OSQ_FUSED_SITE(0) OSQ_FUSED_SITE(1) OSQ_FUSED_SITE(2) OSQ_FUSED_SITE(3)                                            \
  }
OSQ_FUSED_MULTI_ENTRY(fused_fq_linear_multi_kernel, false)
OSQ_FUSED_MULTI_ENTRY(fused_fq_linear_multi_kernel_pair, true)

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
      c = c < -128 ? -128 : (c > 127 ? 127 : c);  // contract: zp == 0 for 8-bit ranges; saturate, never wrap
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

// Y [M, N] fp32 viewed as [M][N/32][32]: one box = `rows` rows x 2 adjacent 32-column groups (256 contiguous bytes per row)
static int make_map_3d_y(CUtensorMap* map, const void* y, uint64_t n, uint64_t m, uint32_t rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return OSQ_ECUDA;
  }
  cuuint64_t dims[3] = {32, n / 32, m};
  cuuint64_t strides[2] = {128, n * 4};
  cuuint32_t box[3] = {32, 2, rows};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(y), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (3-D Y map) failed with CUresult %d (N=%llu M=%llu)", (int)r, (unsigned long long)n, (unsigned long long)m);
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
  // q - zp must fit the s8 operand: any zp in [qmin, qmax] is safe up to 7 bits; an 8-bit range only with the
  // symmetric layout [-128, 127] (zp == 0).  Asymmetric 8-bit weights ([0, 255]) are NOT representable.
  OSQ_CHECK_ARG(qmax - qmin <= 127 || (qmin >= -128 && qmax <= 127),
                "osq_pack_weight_s8: q - zp does not fit int8 (asymmetric 8-bit weights are not supported by the fused path)");
  int sms = sm_count();
  if (sms <= 0) { set_error("no CUDA device"); return OSQ_ECUDA; }
  int64_t g = N < (int64_t)sms * 8 ? N : (int64_t)sms * 8;
  pack_weight_s8_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(w, N, K, scale, zp, (float)qmin, (float)qmax, codes, rowsum);
  OSQ_LAUNCH_CHECK();
  return OSQ_OK;
}

}  // extern "C" (closed here: the planner and launcher below are internal)

namespace osq {
struct FusedPlan {
  FusedParams p;
  CUtensorMap map_w, map_y, map_y16, map_c, map_a;
  int grid;
  size_t smem;
};

static int plan_fused(const osq_fused_linear_t* a, FusedPlan* out) {
  OSQ_CHECK_ARG(a != nullptr, "osq_fused_fq_linear: null args");
  OSQ_CHECK_ARG((a->A || a->a_codes) && a->a_scale && a->a_zp && a->w_codes && a->w_scale && a->w_rowsum && a->Y,
                "osq_fused_fq_linear: null pointer");
  OSQ_CHECK_ARG(a->M >= 1 && a->M < (1ll << 31) - 256, "osq_fused_fq_linear: M out of range");
  OSQ_CHECK_ARG(a->K >= kStageK && a->K % kStageK == 0 && a->K <= 32768, "osq_fused_fq_linear: K must be a multiple of 128, at most 32768 (int32 accumulators)");
  OSQ_CHECK_ARG(a->N >= 16 && a->N % 16 == 0 && a->N <= (1 << 20), "osq_fused_fq_linear: N must be a multiple of 16");
  OSQ_CHECK_ARG(a->a_qmax - a->a_qmin <= 255 && a->a_qmin < a->a_qmax, "osq_fused_fq_linear: activation bits > 8");
  OSQ_CHECK_ARG(a->mma_kind == 0 || a->mma_kind == 1, "osq_fused_fq_linear: mma_kind %d not built", a->mma_kind);
  OSQ_CHECK_ARG((((uintptr_t)a->A) & 15) == 0 && (((uintptr_t)a->Y) & 15) == 0 && (((uintptr_t)a->w_codes) & 15) == 0,
                "osq_fused_fq_linear: A, Y and w_codes must be 16-byte aligned");
  OSQ_CHECK_ARG(!(a->lsq_grad_factor > 0.f && a->a_zp_is_int32), "osq_fused_fq_linear: LSQ+ needs a float zero_point");
  OSQ_CHECK_ARG(a->a_codes == nullptr || (((uintptr_t)a->a_codes) & 15) == 0, "osq_fused_fq_linear: a_codes must be 16-byte aligned");
  if (a->out_scale != nullptr) {
    OSQ_CHECK_ARG(a->out_zp != nullptr && a->out_qmin < a->out_qmax && a->out_qmax - a->out_qmin <= 255, "osq_fused_fq_linear: bad output quantizer");
    OSQ_CHECK_ARG(a->out_act == 0 || a->out_act == 1, "osq_fused_fq_linear: out_act must be 0 (none) or 1 (GELU)");
    OSQ_CHECK_ARG(!(a->out_lsq_grad_factor > 0.f && a->out_zp_is_int32), "osq_fused_fq_linear: an LSQ+ output quantizer needs a float zero_point");
    OSQ_CHECK_ARG(a->out_bins == nullptr || (((uintptr_t)a->out_bins) & 15) == 0, "osq_fused_fq_linear: out_bins must be 16-byte aligned");
  } else {
    OSQ_CHECK_ARG(a->out_act == 0 && a->out_bins == nullptr, "osq_fused_fq_linear: out_act / out_bins need an output quantizer (out_scale)");
  }

  int dev = 0, cc_major = 0, sms = sm_count();
  OSQ_CUDA(cudaGetDevice(&dev));
  OSQ_CUDA(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  if (cc_major != 10) {
    set_error("osq_fused_fq_linear needs an sm_100 device (found compute capability major %d)", cc_major);
    return OSQ_EARCH;
  }

  FusedParams p;
  memset(&p, 0, sizeof(p));
  p.M = (int)a->M; p.K = (int)a->K; p.N = (int)a->N;
  p.KB = p.K / kStageK;
  p.A = a->A;
  p.a_scale = a->a_scale; p.a_zp = a->a_zp; p.a_zp_is_int32 = a->a_zp_is_int32; p.g = a->lsq_grad_factor;
  p.qmin = (float)a->a_qmin; p.qmax = (float)a->a_qmax;
  p.w_scale = a->w_scale; p.w_rowsum = a->w_rowsum; p.bias = a->bias; p.a_codes = a->a_codes; p.Y = a->Y;
  p.trace = (long long*)a->debug_trace;
  p.out_act = a->out_act; p.out_scale = a->out_scale; p.out_zp = a->out_zp; p.out_zp_is_int32 = a->out_zp_is_int32;
  p.out_g = a->out_lsq_grad_factor; p.out_qmin = (float)a->out_qmin; p.out_qmax = (float)a->out_qmax; p.out_bins = a->out_bins;
  static int env_dbg = -1, env_ob = -1, env_kb = -1, env_csz = -1, env_pdl = -1, env_pf = -1, env_xtma = -1, env_bn = -1;
  if (env_dbg < 0) {
    auto geti = [](const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; };
    env_dbg = geti("OSQ_FUSED_DBG", 0);
    env_ob = geti("OSQ_FUSED_OUTBUFS", 0);
    env_kb = geti("OSQ_FUSED_SMEM_KB", 227);
    if (env_kb < 100 || env_kb > 227) env_kb = 227;
    env_csz = geti("OSQ_FUSED_CLUSTER", 2);  // CTA pair whenever the plan allows it (resident A, >= 2 tiles, fp32-in)
    env_pdl = geti("OSQ_FUSED_PDL", 1);
    env_pf = geti("OSQ_FUSED_PREFETCH", 0);
    env_xtma = geti("OSQ_FUSED_XTMA", 1);
    env_bn = geti("OSQ_FUSED_BN", 0);
  }
  p.dbg = env_dbg;
  p.pdl = env_pdl ? 1 : 0;

  static bool attr_set[64] = {false};
  if (!attr_set[dev & 63]) {
    OSQ_CUDA(cudaFuncSetAttribute(fused_fq_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    OSQ_CUDA(cudaFuncSetAttribute(fused_fq_linear_kernel_ldg, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    OSQ_CUDA(cudaFuncSetAttribute(fused_fq_linear_kernel_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    OSQ_CUDA(cudaFuncSetAttribute(fused_fq_linear_kernel_s3, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    OSQ_CUDA(cudaFuncSetAttribute(fused_fq_linear_kernel_ldg_s3, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    OSQ_CUDA(cudaFuncSetAttribute(fused_fq_linear_kernel_pair_s3, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    OSQ_CUDA(cudaFuncSetAttribute(fused_fq_linear_kernel_e16, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    OSQ_CUDA(cudaFuncSetAttribute(fused_fq_linear_kernel_pair_e16, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    OSQ_CUDA(cudaFuncSetAttribute(fused_fq_linear_kernel_dr, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    OSQ_CUDA(cudaFuncSetAttribute(fused_fq_linear_kernel_pair_dr, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    OSQ_CUDA(cudaFuncSetAttribute(fused_fq_linear_kernel_epi, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    OSQ_CUDA(cudaFuncSetAttribute(fused_fq_linear_kernel_epi_e16, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    OSQ_CUDA(cudaFuncSetAttribute(fused_fq_linear_kernel_pair_epi_e16, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    OSQ_CUDA(cudaFuncSetAttribute(fused_fq_linear_kernel_pair_epi, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    OSQ_CUDA(cudaFuncSetAttribute(fused_fq_linear_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    OSQ_CUDA(cudaFuncSetAttribute(fused_fq_linear_multi_kernel_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set[dev & 63] = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(kNumThreads);
  cfg.stream = nullptr;   // the occupancy query below does not depend on the stream
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = p.pdl ? 2 : 1;

  // ---- shared-memory plan (<= 227 KB / CTA):
  //   [A ring: a_stages x (rows_per_tile x 128 B)] [W ring: w_stages x (BN / csz x 128 B)] [X: 16 x 4 KB fp32 landing slots]
  //   [O: 8 x out_bufs x 4 KB TMA-store staging, aliased onto X when space is short] [barriers + 2 BN constants]
  // A stages hold only the tile's valid rows: the MMA (M = 128) reads past them into the next stage / the W ring,
  // which only produces accumulator rows nobody stores.
  // CTA pair (csz = 2, cta_group::2): two CTAs on neighbouring SMs run their two row tiles against ONE copy of every
  // W tile (each stages half of its rows), which halves the W traffic into and out of shared memory -- the
  // resource that bounds the epilogue phase.  Tried first; needs resident A, the TMA landing slots and >= 2 tiles.
  const int x_bytes = kNumWorkers * kXSlotBytes;          // 64 KB
  const int out1 = kNumEpiWarps * kOutTileBytes;          // 32 KB per buffer set
  const int total = env_kb * 1024 - (int)sizeof(Smem);
  struct Plan { int ok, bn, resident, cached, a_stages, w_stages, out_bufs, x_tma, alias, score, x_depth; };
  Plan best; memset(&best, 0, sizeof(best));
  static int max_ctas[64][3] = {{0}};
  int grid = 0, nsplit_max = 1;
  bool split_n = false;
  for (int csz = env_csz == 2 ? 2 : 1; csz >= 1 && !best.ok; --csz) {
    p.csz = csz;
    attr[0].val.clusterDim.x = (unsigned)csz;
    // how many CTAs can be co-resident (1 CTA / SM)
    if (max_ctas[dev & 63][csz] == 0) {
      if (csz > 1) {
        int n_clusters = 0;
        cfg.gridDim = dim3((unsigned)(sms / csz * csz));
        cfg.dynamicSmemBytes = 227 * 1024;
        OSQ_CUDA(cudaOccupancyMaxActiveClusters(&n_clusters, fused_fq_linear_kernel_pair, &cfg));
        if (n_clusters <= 0) n_clusters = sms / csz;  // the query is only a hint; the launch itself reports a real problem
        max_ctas[dev & 63][csz] = n_clusters * csz;
      } else {
        max_ctas[dev & 63][csz] = sms;
      }
      if (max_ctas[dev & 63][csz] <= 0) { if (csz > 1) continue; set_error("osq_fused_fq_linear: no resident CTA"); return OSQ_ECUDA; }
    }
    int G = max_ctas[dev & 63][csz];
    {  // experiment knob: leave SMs free for a second stream (two half-batches out of phase)
      static int env_cap = -1;
      if (env_cap < 0) { const char* e = getenv("OSQ_FUSED_MAX_CTAS"); env_cap = e ? atoi(e) : 0; }
      if (env_cap > 0 && env_cap < G) G = env_cap / csz * csz;
    }
    // rows per CTA tile: the 128-row MMA tile is filled with as many rows as make the tile count a multiple of
    // the resident CTA count (M = 16384 on 148 SMs: 147 tiles of 112 rows instead of 128 tiles of 128 rows)
    {
      const int64_t waves = (p.M + (int64_t)kBM * G - 1) / ((int64_t)kBM * G);
      int64_t rpt = (p.M + waves * G - 1) / (waves * G);
      rpt = (rpt + 15) / 16 * 16;
      if (rpt > kBM) rpt = kBM;
      if (rpt < 16) rpt = 16;
      // Few rows (M < 64 x SMs): shrinking the row tile further would leave every CTA streaming all of W for a
      // handful of rows, with one TMEM lane quarter busy in the epilogue.  Keep 64-row tiles and split N across
      // groups of CTAs instead (each group converts the same rows, which come from L2 after the first touch).
      static int env_nsplit = -1;
      if (env_nsplit < 0) { const char* e = getenv("OSQ_FUSED_NSPLIT"); env_nsplit = e ? atoi(e) : 1; }
      split_n = env_nsplit != 0 && rpt < 64 && p.N > 256;
      // (decode-sized launches, M <= 64 -- BART generation runs every decoder Linear at batch x beams rows: ONE row tile holding
      //  all rows, N spread over as many CTAs as there are chunks; without this a 24-row launch ran on two CTAs: 20-60 us)
      if (split_n) rpt = p.M > 64 ? 64 : (p.M + 15) / 16 * 16;
      p.rows_per_tile = (int)rpt;
    }
    p.n_mblocks = (p.M + p.rows_per_tile - 1) / p.rows_per_tile;
    if (csz > 1 && p.n_mblocks < 2) continue;
    grid = p.n_mblocks < G ? p.n_mblocks : G;
    grid = (grid + csz - 1) / csz * csz;
    p.n_iters = (p.n_mblocks + grid - 1) / grid;
    p.mb_count = grid;
    p.nsplit = 1;
    nsplit_max = split_n ? G / grid : 1;
    p.a_stage_bytes = p.rows_per_tile * kStageK;

    // one row tile and a column split: narrow chunks spread N over more CTAs (the launch is bound by the W stream per CTA)
    const bool narrow = split_n && p.n_mblocks == 1 && p.N % 128 == 0 && p.N >= 512;
    static int env_xdepth = -1;
    if (env_xdepth < 0) { const char* e = getenv("OSQ_FUSED_XDEPTH"); env_xdepth = e ? atoi(e) : 1; }
    auto make_plan = [&](int bn, int x_tma) {
      Plan pl; memset(&pl, 0, sizeof(pl));
      pl.bn = bn; pl.x_tma = x_tma; pl.x_depth = 1;
      const int nc = (p.N + bn - 1) / bn;
      const int w_stage = (bn / csz * kStageK + 1023) / 1024 * 1024;
      const int budget = total - 2 * ((bn + 31) & ~31) * (int)sizeof(float);
      const int xo_min = x_tma ? x_bytes : out1;            // aliased X/O region, or one set of store tiles
      pl.resident = (p.KB <= kMaxAStages && p.KB * p.a_stage_bytes + 2 * w_stage + xo_min <= budget) ? 1 : 0;
      if (csz > 1 && !x_tma) return pl;
      pl.cached = (!pl.resident && (nc > 1 || a->A == nullptr) && a->a_codes != nullptr) ? 1 : 0;
      // a single decode-sized row tile: re-converting its few rows per chunk is cheaper than tying four chunks to one CTA
      if (narrow && a->A != nullptr) pl.cached = 0;
      if (csz > 1 && !(pl.resident || pl.cached)) return pl;  // the pair needs resident A or the code-cache sweeps
      pl.a_stages = pl.resident ? p.KB : 4;
      // X and O can share memory only when conversion and epilogue never interleave inside a tile
      const bool can_alias = pl.resident || pl.cached || nc == 1;
      int rest = budget - pl.a_stages * p.a_stage_bytes - 2 * w_stage;
      pl.w_stages = 2;
      // streamed fp32-in plans: a second landing slot per worker when it fits next to 4 A and 2 W stages (the read phase of
      // K > 1024 sites is bound by bytes in flight)
      if (x_tma && env_xdepth == 2 && pl.cached && a->A != nullptr && can_alias && rest >= 2 * x_bytes) {
        pl.x_depth = 2; pl.alias = 1; pl.out_bufs = 2; rest -= 2 * x_bytes;
      } else
      if (x_tma) {
        if (can_alias && rest >= x_bytes) { pl.alias = 1; pl.out_bufs = 2; rest -= x_bytes; }
        else if (rest >= x_bytes + out1) { pl.alias = 0; pl.out_bufs = 1; rest -= x_bytes + out1; }
        else return pl;                                     // does not fit
      } else {
        if (rest < out1) return pl;
        pl.out_bufs = 1; rest -= out1;
      }
      // the W stream needs ~2.5 stages of 32 KB in flight to cover the L2 latency: a third stage comes first
      if (rest >= w_stage) { ++pl.w_stages; rest -= w_stage; }
      if (!pl.alias && pl.out_bufs == 1 && env_ob != 1 && rest >= out1) { pl.out_bufs = 2; rest -= out1; }
      // beyond that: W to four stages, then (streamed A) the A ring to six, then whatever is left to W and A
      // (measured on 3072->768: 4 W + 6 A stages beat 6 + 4 and 3 + 8)
      while (pl.w_stages < 4 && rest >= w_stage) { ++pl.w_stages; rest -= w_stage; }
      if (!pl.resident)
        while (pl.a_stages < 6 && rest >= p.a_stage_bytes) { ++pl.a_stages; rest -= p.a_stage_bytes; }
      while (pl.w_stages < kMaxWStages && rest >= w_stage) { ++pl.w_stages; rest -= w_stage; }
      if (!pl.resident)
        while (pl.a_stages < kMaxAStages && rest >= p.a_stage_bytes) { ++pl.a_stages; rest -= p.a_stage_bytes; }
      pl.ok = 1;
      // preference: three W stages and double-buffered stores matter more than the chunk width
      pl.score = (pl.w_stages >= 3 ? 4 : 0) + (pl.out_bufs >= 2 ? 2 : 0) + (bn == 256 ? 1 : 0);
      return pl;
    };
      const bool x_ok = (env_xtma != 0 && p.K % 4 == 0) || a->A == nullptr;  // bins-in launches use the TMA variant's layout
    const int bn_cands[3] = {256, 192, 128};
    for (int xt = x_ok ? 1 : 0; xt >= 0 && !best.ok; --xt)
      for (int i = 0; i < 3; ++i) {
        int bn = bn_cands[i];
        if (env_bn > 0 && bn != env_bn) continue;
        if (env_bn <= 0 && narrow && bn != 128) continue;
        if (p.N < bn) { if (i == 0) bn = p.N; else continue; }   // narrow layers: one chunk of N columns
        else if (p.N % bn != 0 && i != 0) continue;               // 192 / 128 only when they tile N exactly
        if (bn % (16 * csz) != 0) continue;
        Plan pl = make_plan(bn, xt);
        if (pl.ok && (!best.ok || pl.score > best.score)) best = pl;
      }
  }
  if (!best.ok) { set_error("osq_fused_fq_linear: no shared-memory plan for K=%d N=%d", p.K, p.N); return OSQ_EINVAL; }
  p.BN = best.bn;
  p.NC = (p.N + p.BN - 1) / p.BN;
  p.acc_stages = kTmemCols / p.BN;
  if (p.acc_stages > kMaxAccStages) p.acc_stages = kMaxAccStages;
  p.w_stage_bytes = (p.BN / p.csz * kStageK + 1023) / 1024 * 1024;  // pair: each CTA stages half of the tile's rows
  p.resident = best.resident; p.cached = best.cached;
  p.codes_in = (a->A == nullptr) ? 1 : 0;
  p.a_stages = best.a_stages; p.w_stages = best.w_stages; p.out_bufs = best.out_bufs;
  p.x_tma = best.x_tma; p.alias_xo = best.alias; p.x_depth = best.x_tma ? best.x_depth : 1;
  // streamed A with a code cache: all accumulator stages are fed in one sweep over K (the later sweeps re-load the
  // bins by TMA).  Without a cache the workers re-convert per chunk and must keep one stage free for the epilogue
  // they run in between, so they stay at one chunk per sweep.
  p.cpp = p.cached ? p.acc_stages : 1;
  // column split: `cps` chunks per group of CTAs, a whole number of sweeps each
  p.cps = p.NC;
  if (nsplit_max > 1 && p.NC > 1) {
    int ns = nsplit_max < p.NC ? nsplit_max : p.NC;
    int cps = (p.NC + ns - 1) / ns;
    cps = (cps + p.cpp - 1) / p.cpp * p.cpp;
    p.cps = cps;
    p.nsplit = (p.NC + cps - 1) / cps;
    grid = p.mb_count * p.nsplit;
  }
  p.a_passes = p.resident ? 1 : (p.cps + p.cpp - 1) / p.cpp;
  const int const_bytes = 2 * ((p.BN + 31) & ~31) * (int)sizeof(float);
  const size_t smem_bytes = (size_t)p.a_stages * p.a_stage_bytes + (size_t)p.w_stages * p.w_stage_bytes +
                            (p.x_tma ? (size_t)x_bytes * (size_t)p.x_depth : 0) + (p.alias_xo ? 0 : (size_t)p.out_bufs * out1) +
                            sizeof(Smem) + (size_t)const_bytes;
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.gridDim = dim3((unsigned)grid);
  p.codes_box_bytes = (uint32_t)(p.M < p.rows_per_tile ? p.M : p.rows_per_tile) * kStageK;

  CUtensorMap map_w, map_y, map_y16, map_c, map_a;
  if (int rc = make_map_2d(&map_w, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, a->w_codes, (uint64_t)p.K, (uint64_t)p.N, kStageK,
                           (uint32_t)(p.BN / p.csz), CU_TENSOR_MAP_SWIZZLE_128B))
    return rc;
  static int env_s3d = -1;
  // measured (round 2): the pair-shared 256-byte-row stores are 2-3 % SLOWER in the kernel than per-warp 32 x 128 B tiles
  // (0.657 vs 0.670 of the HBM roofline over the four BERT-base sites) although they are 10 % faster in isolation
  // (profiles/r01_storebench_3d.txt): kept as a tested variant, off by default
  if (env_s3d < 0) { const char* e = getenv("OSQ_FUSED_STORE3D"); env_s3d = e ? atoi(e) : 0; }
  // pair stores need whole 32-column groups, chunks that split into 64-column pairs, and two 4 KB staging tiles per quarter
  if (a->out_scale != nullptr && !p.x_tma) {
    set_error("osq_fused_fq_linear: the output stage is built for the TMA landing-slot plans only (OSQ_FUSED_XTMA=0 is set?)");
    return OSQ_EINVAL;
  }
  p.store3d = (a->out_scale == nullptr && env_s3d != 0 && p.N % 32 == 0 && p.BN % 64 == 0 && p.M >= 32 && (p.alias_xo || p.out_bufs >= 1)) ? 1 : 0;
  static int env_drain = -1;
  if (env_drain < 0) { const char* e = getenv("OSQ_FUSED_DRAIN"); env_drain = e ? atoi(e) : 0; }
  static int env_pre = -1;
  if (env_pre < 0) { const char* e = getenv("OSQ_FUSED_PRE_L2"); env_pre = e ? atoi(e) : 0; }
  p.pre_l2 = env_pre;
  static int env_lsu = -1;
  if (env_lsu < 0) { const char* e = getenv("OSQ_FUSED_LSU_MOD"); env_lsu = e ? atoi(e) : 0; }
  p.lsu_mod = (p.trace == nullptr && env_lsu > 0) ? env_lsu : 0;
  static int env_e16 = -1;
  // measured: with plain Y stores the sixteen-warp epilogue changes nothing (the store path bounds the write phase: 54.1 vs 53.6 us on
  // 768->3072), without stores it is 14 % faster (36.9 vs 42.7 us).  Default: only where the epilogue computes (output stage).
  if (env_e16 < 0) { const char* e = getenv("OSQ_FUSED_EPI16"); env_e16 = e ? atoi(e) : 0; }
  // drain epilogue: needs the landing slots as staging (aliased plans), whole 64-column steps, no output stage
  p.drain = (env_drain != 0 && !p.store3d && a->out_scale == nullptr && p.x_tma && p.alias_xo && p.N % 64 == 0 && p.BN % 64 == 0 &&
             p.trace == nullptr) ? 1 : 0;
  // sixteen epilogue warps: needs one landing slot per worker as its staging tile (aliased plans)
  static int env_e16o = -1;
  if (env_e16o < 0) { const char* e = getenv("OSQ_FUSED_EPI16_OUT"); env_e16o = e ? atoi(e) : 1; }
  p.epi16 = ((a->out_scale == nullptr ? env_e16 != 0 : env_e16o != 0) && !p.drain && !p.store3d && p.x_tma && p.alias_xo) ? 1 : 0;
  if (p.store3d) {
    if (int rc = make_map_3d_y(&map_y, a->Y, (uint64_t)p.N, (uint64_t)p.M, 32)) return rc;
    if (int rc = make_map_3d_y(&map_y16, a->Y, (uint64_t)p.N, (uint64_t)p.M, 16)) return rc;
  } else {
  if (int rc = make_map_2d(&map_y, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a->Y, (uint64_t)p.N, (uint64_t)p.M, 32,
                           p.M < 32 ? (uint32_t)p.M : 32u, CU_TENSOR_MAP_SWIZZLE_128B))
    return rc;
  if (int rc = make_map_2d(&map_y16, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a->Y, (uint64_t)p.N, (uint64_t)p.M, 32,
                           p.M < 16 ? (uint32_t)p.M : 16u, CU_TENSOR_MAP_SWIZZLE_128B))
    return rc;
  }
  if (p.cached || p.codes_in) {
    if (int rc = make_map_2d(&map_c, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, a->a_codes, (uint64_t)p.K, (uint64_t)p.M, kStageK,
                             (uint32_t)(p.M < p.rows_per_tile ? p.M : p.rows_per_tile), CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  } else {
    map_c = map_w;  // unused by the kernel
  }
  // fp32 activation: [8 rows x 128 floats] boxes (one worker's rows of one k-block) for the landing slots; the
  // same map serves the optional L2 prefetcher (OSQ_FUSED_PREFETCH = distance in k-blocks; measured: 3 is neutral,
  // 6 and 12 are 3-8 % slower -> off by default)
  p.prefetch = (p.x_tma == 0 && !p.codes_in && p.K % 4 == 0) ? env_pf : 0;
  if ((p.x_tma && !p.codes_in) || p.prefetch > 0) {
    if (int rc = make_map_2d(&map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a->A, (uint64_t)p.K, (uint64_t)p.M, kStageK,
                             (uint32_t)(p.M < kRowsPerWorker ? p.M : kRowsPerWorker), CU_TENSOR_MAP_SWIZZLE_NONE))
      return rc;
  } else {
    map_a = map_w;  // unused by the kernel
  }
  static int env_verbose = -1;
  if (env_verbose < 0) { const char* e = getenv("OSQ_FUSED_VERBOSE"); env_verbose = e ? atoi(e) : 0; }
  if (env_verbose) {
    static int seen[16][3]; static int n_seen = 0;
    bool known = false;
    for (int i = 0; i < n_seen; ++i) known |= (seen[i][0] == p.M && seen[i][1] == p.K && seen[i][2] == p.N);
    if (!known && n_seen < 16) {
      seen[n_seen][0] = p.M; seen[n_seen][1] = p.K; seen[n_seen][2] = p.N; ++n_seen;
      fprintf(stderr, "[osq] fused M=%d K=%d N=%d: grid=%d (n-split %d) cluster=%d rows/tile=%d tiles/cta=%d BN=%d chunks=%d mode=%s a_stages=%d(%d B) "
                      "w_stages=%d(%d B) acc_stages=%d out_bufs=%d x_tma=%d(x%d) alias=%d sweeps=%d smem=%zu\n",
              p.M, p.K, p.N, grid, p.nsplit, p.csz, p.rows_per_tile, p.n_iters, p.BN, p.NC, p.codes_in ? (p.resident ? "bins-in resident" : "bins-in streamed") : p.resident ? "resident" : (p.cached ? "streamed+cache" : "streamed"),
              p.a_stages, p.a_stage_bytes, p.w_stages, p.w_stage_bytes, p.acc_stages, p.out_bufs, p.x_tma, p.x_depth, p.alias_xo, p.a_passes, smem_bytes);
    }
  }
  p.first_site = 1;
  p.more_sites = 0;
  out->p = p;
  out->map_w = map_w; out->map_y = map_y; out->map_y16 = map_y16; out->map_c = map_c; out->map_a = map_a;
  out->grid = grid;
  out->smem = smem_bytes;
  return OSQ_OK;
}

static int launch_fused(const FusedPlan& pl, void* stream) {
  const FusedParams& p = pl.p;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(kNumThreads);
  cfg.gridDim = dim3((unsigned)pl.grid);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)p.csz;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = p.pdl ? 2 : 1;
  const CUtensorMap &map_w = pl.map_w, &map_y = pl.map_y, &map_y16 = pl.map_y16, &map_c = pl.map_c, &map_a = pl.map_a;
  if (p.out_scale != nullptr && p.epi16) {
    if (p.csz == 2) OSQ_CUDA(cudaLaunchKernelEx(&cfg, fused_fq_linear_kernel_pair_epi_e16, map_w, map_y, map_y16, map_c, map_a, p));
    else OSQ_CUDA(cudaLaunchKernelEx(&cfg, fused_fq_linear_kernel_epi_e16, map_w, map_y, map_y16, map_c, map_a, p));
  } else if (p.out_scale != nullptr) {
    if (p.csz == 2) OSQ_CUDA(cudaLaunchKernelEx(&cfg, fused_fq_linear_kernel_pair_epi, map_w, map_y, map_y16, map_c, map_a, p));
    else OSQ_CUDA(cudaLaunchKernelEx(&cfg, fused_fq_linear_kernel_epi, map_w, map_y, map_y16, map_c, map_a, p));
  } else if (p.store3d) {
    if (p.csz == 2) OSQ_CUDA(cudaLaunchKernelEx(&cfg, fused_fq_linear_kernel_pair_s3, map_w, map_y, map_y16, map_c, map_a, p));
    else if (p.x_tma) OSQ_CUDA(cudaLaunchKernelEx(&cfg, fused_fq_linear_kernel_s3, map_w, map_y, map_y16, map_c, map_a, p));
    else OSQ_CUDA(cudaLaunchKernelEx(&cfg, fused_fq_linear_kernel_ldg_s3, map_w, map_y, map_y16, map_c, map_a, p));
  } else if (p.epi16) {
    if (p.csz == 2) OSQ_CUDA(cudaLaunchKernelEx(&cfg, fused_fq_linear_kernel_pair_e16, map_w, map_y, map_y16, map_c, map_a, p));
    else OSQ_CUDA(cudaLaunchKernelEx(&cfg, fused_fq_linear_kernel_e16, map_w, map_y, map_y16, map_c, map_a, p));
  } else if (p.drain) {
    if (p.csz == 2) OSQ_CUDA(cudaLaunchKernelEx(&cfg, fused_fq_linear_kernel_pair_dr, map_w, map_y, map_y16, map_c, map_a, p));
    else OSQ_CUDA(cudaLaunchKernelEx(&cfg, fused_fq_linear_kernel_dr, map_w, map_y, map_y16, map_c, map_a, p));
  } else {
    if (p.csz == 2) OSQ_CUDA(cudaLaunchKernelEx(&cfg, fused_fq_linear_kernel_pair, map_w, map_y, map_y16, map_c, map_a, p));
    else if (p.x_tma) OSQ_CUDA(cudaLaunchKernelEx(&cfg, fused_fq_linear_kernel, map_w, map_y, map_y16, map_c, map_a, p));
    else OSQ_CUDA(cudaLaunchKernelEx(&cfg, fused_fq_linear_kernel_ldg, map_w, map_y, map_y16, map_c, map_a, p));
  }
  OSQ_LAUNCH_CHECK();
  return OSQ_OK;
}

}  // namespace osq

extern "C" {

int osq_fused_fq_linear(const osq_fused_linear_t* a, void* stream) {
  using namespace osq;
  FusedPlan pl;
  if (int rc = plan_fused(a, &pl)) return rc;
  return launch_fused(pl, stream);
}

int osq_fused_fq_linear_multi(const osq_fused_linear_t* sites, int n_sites, void* stream) {
  using namespace osq;
  OSQ_CHECK_ARG(sites != nullptr && n_sites >= 1, "osq_fused_fq_linear_multi: no sites");
  static int env_multi = -1;
  // measured (round 2, BERT-base sites, M = 16384): the persistent multi-site grid is 10 % SLOWER than one launch per site
  // (2.50 vs 2.23 ms per 48-site step).  Its CTAs drift out of phase, so the chip reads and writes HBM at the same time
  // instead of in two bulk phases, and the mixed traffic costs more (DRAM bus turnarounds) than the saved launch gaps.
  // Off by default (the entry point then issues the sites one by one); OSQ_FUSED_MULTI=1 enables it.
  if (env_multi < 0) { const char* e = getenv("OSQ_FUSED_MULTI"); env_multi = e ? atoi(e) : 0; }
  // one persistent grid per run of compatible sites (same grid, same cluster size, TMA variant, plain tile stores);
  // anything else is launched on its own
  int i = 0;
  while (i < n_sites) {
    static thread_local FusedSites fs;
    FusedPlan first;
    if (int rc = plan_fused(&sites[i], &first)) return rc;
    int n = 0;
    size_t smem = first.smem;
    auto take = [&](const FusedPlan& pl) {
      FusedSite& d = fs.s[n++];
      d.w = pl.map_w; d.y = pl.map_y; d.y16 = pl.map_y16; d.codes = pl.map_c; d.a = pl.map_a; d.p = pl.p;
      if (pl.smem > smem) smem = pl.smem;
    };
    const bool ok0 = env_multi != 0 && first.p.x_tma && !first.p.store3d && first.p.trace == nullptr && first.p.out_scale == nullptr;
    if (!ok0) {
      if (int rc = launch_fused(first, stream)) return rc;
      ++i;
      continue;
    }
    take(first);
    int j = i + 1;
    for (; j < n_sites && n < kMaxSites; ++j) {
      FusedPlan pl;
      if (int rc = plan_fused(&sites[j], &pl)) return rc;
      if (!(pl.p.x_tma && !pl.p.store3d && pl.p.trace == nullptr && pl.p.out_scale == nullptr && pl.grid == first.grid &&
            pl.p.csz == first.p.csz)) break;
      take(pl);
    }
    if (n == 1) {
      if (int rc = launch_fused(first, stream)) return rc;
    } else {
      for (int k = 0; k < n; ++k) { fs.s[k].p.more_sites = (k + 1 < n) ? 1 : 0; fs.s[k].p.first_site = (k == 0) ? 1 : 0; }
      fs.n = n;
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.blockDim = dim3(kNumThreads);
      cfg.gridDim = dim3((unsigned)first.grid);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = (cudaStream_t)stream;
      cudaLaunchAttribute attr[2];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = (unsigned)first.p.csz;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[1].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = first.p.pdl ? 2 : 1;
      if (first.p.csz == 2) OSQ_CUDA(cudaLaunchKernelEx(&cfg, fused_fq_linear_multi_kernel_pair, fs));
      else OSQ_CUDA(cudaLaunchKernelEx(&cfg, fused_fq_linear_multi_kernel, fs));
      OSQ_LAUNCH_CHECK();
    }
    i = j;
  }
  return OSQ_OK;
}

}  // extern "C"
