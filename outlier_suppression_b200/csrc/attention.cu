// attention.cu -- K8 / K9: the two attention-side contractions with their activation quantizers as prologues
// (model/quant_bert.py:148-150 and :185-187; quant_roberta.py likewise).
//
//   K8  scores  = fq_q(Q) @ fq_k(K)^T  [* 1/sqrt(d)] [+ mask]          Q, K: [B, h, S, d] views of the [B, S, h*d] projections
//   K9  context = fq_p(P) @ fq_v(V)     [-> fq_o(context) + bins]      P: [B, h, Sq, Sk] probabilities, V like Q;
//                                                                       context leaves as [B, Sq, h*d] (permute + view of :189-191)
//
// Both operands of either product are fake-quantised ACTIVATIONS, so the product of the dequantised fp32 tensors is
//   s_a s_b * sum_k (a_k - Za)(b_k - Zb),   a_k, b_k = integer bins,
// an exact integer contraction (u8 x u8 -> s32) followed by one scale -- more accurate than the reference's fp32 GEMM over
// the dequantised values (whose partial sums round), and it never materialises the fake-quantised operands.
//
// These contractions are HBM-bound, not tensor-bound: per 128 x 128 output tile K8 performs 2 MFLOP for 64 KB written
// (d = 64), K9 reads 64 KB of probabilities per 1 MFLOP.  Warp-level mma.sync.m16n8k32 keeps the integer pipe far below its
// ceiling at those ratios; the design effort is in the memory path: quantise-on-load (one read of each fp32 operand, 128-bit
// streaming loads), bins staged in padded shared-memory tiles (conflict-free fragment loads), outputs staged per warp and
// written as whole 512-byte (K8) / 256-byte (K9) row segments.
//
// Algorithmic bytes:  K8: 4 (|Q| + |K|) + 4 B h Sq Sk      K9: 4 B h Sq Sk + 4 |V| + 4 |context| (+ |context| bins)
#include "common.cuh"

namespace osq {

constexpr int kAttThreads = 256;   // 8 warps x 16 query rows
constexpr int kAttRows = 128;      // query rows per CTA
constexpr int kPad = 16;           // bytes of padding per bin row: fragment loads of 8 rows x 4 words hit 32 distinct banks

struct AttnQ {
  const float* scale;
  const void* zp;
  int zp_is_int32;
  float g, qmin, qmax;
};

struct ScoresParams {
  const float* q;
  const float* k;
  int64_t qs[3], ks[3];   // batch, head, token strides in elements
  int heads, sq, sk;
  int kc;                 // key rows quantised into shared memory per chunk (multiple of 128, <= 512)
  AttnQ qq, kq;
  float out_mul;          // 1/sqrt(d) as ATen's scalar division computes it (multiplication by the fp32 reciprocal); 1 = off
  const float* mask;      // optional additive [batch, sk]
  float* out;             // [batch, heads, sq, sk]
};

struct ContextParams {
  const float* probs;     // [batch, heads, sq, sk]
  const float* v;
  int64_t vs[3];
  int heads, sq, sk;
  int kc;                 // keys quantised into shared memory per chunk (multiple of 64, <= 512)
  AttnQ pq, vq, oq;
  int has_oq;
  float* out;             // [batch, sq, heads * d]
  uint8_t* bins;          // optional, same shape
};

// Bins (q - qmin) of four adjacent elements: quant_bin4 (common.cuh) -- magic-number rounding of x * (1/s), the whole group redone
// with the IEEE division when any element is within 1e-4 of a rounding tie, i.e. bit-exact with util_quant.py:12-14.

// four 8 x 16-byte tiles of bins in one instruction: lane l supplies the address of row (l & 7) of tile (l >> 3); thread (g, t) of
// the warp receives bytes [4t, 4t + 4) of row g of each tile -- exactly the m16n8k32 fragment words
__device__ __forceinline__ void ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* row_ptr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"((uint32_t)__cvta_generic_to_shared(row_ptr)));
}

// sum of the four bytes of w, added to acc
__device__ __forceinline__ int bytesum(uint32_t w, int acc) { return (int)__dp4a(w, 0x01010101u, (unsigned int)acc); }

__device__ __forceinline__ void mma_u8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// ------------------------------------------------------------------------------------------------------------------
// K8: scores.  CTA = (128 query rows, head, batch).  The key bins of the head (up to 512 keys per chunk) are quantised ONCE into
// shared memory; after that every warp runs on its own -- 16 query rows x 128 keys per step: fragments from shared memory,
// integer zero-point corrections, staging and 256-byte row stores -- with no block-wide barrier in the loop.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kHalfStage = 72;   // floats per staged row of a 64-column half (64 + 8: float2 fragment stores and float4 row reads conflict-free)

template <int D>
__global__ void __launch_bounds__(kAttThreads, 2)
attn_scores_kernel(const ScoresParams p) {
  constexpr int kRow = D + kPad;           // bytes per bin row
  extern __shared__ __align__(16) uint8_t sm_raw[];
  const int KC = p.kc;                                   // key rows resident per chunk (multiple of 128)
  uint8_t* qc = sm_raw;                                  // [128][kRow]
  uint8_t* kc = qc + kAttRows * kRow;                    // [KC][kRow]
  int* qsum = reinterpret_cast<int*>(kc + (size_t)KC * kRow);  // [128]
  int* ksum = qsum + kAttRows;                           // [KC]  per key: kconst - Zq * (sum of the key's bins)
  float* kmask = reinterpret_cast<float*>(ksum + KC);    // [KC]  per key: additive mask (-0.0f without one: x + -0.0f == x bit for bit)
  float* stage = kmask + KC;                             // [8][16][kHalfStage]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int head = blockIdx.y, b = blockIdx.z, q0 = blockIdx.x * kAttRows;
  const bool wb = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 0;
  const QParam qq = load_qparam(p.qq.scale, p.qq.zp, p.qq.zp_is_int32, p.qq.g, p.qq.qmin, p.qq.qmax, wb);
  const QParam kq = load_qparam(p.kq.scale, p.kq.zp, p.kq.zp_is_int32, p.kq.g, p.kq.qmin, p.kq.qmax, wb);
  const ConvParam qcp = make_conv_param(qq.s, qq.z, p.qq.qmin, p.qq.qmax), kcp = make_conv_param(kq.s, kq.z, p.kq.qmin, p.kq.qmax);
  const int zcq = (int)qcp.zc, zck = (int)kcp.zc;
  const float sqk = __fmul_rn(qq.s, kq.s);
  const int kconst = D * zcq * zck;

  // ---- Q tile: quantise on load
  const float* qbase = p.q + (size_t)b * p.qs[0] + (size_t)head * p.qs[1];
  // (four independent 128-bit loads in flight per thread: the loop is latency-bound otherwise)
  for (int base = tid; base < kAttRows * (D / 4); base += 4 * kAttThreads) {
    float4 x[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = base + u * kAttThreads, r = idx / (D / 4), c4 = idx % (D / 4);
      x[u] = (idx < kAttRows * (D / 4) && q0 + r < p.sq) ? ldg_stream(reinterpret_cast<const float4*>(qbase + (size_t)(q0 + r) * p.qs[2]) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = base + u * kAttThreads, r = idx / (D / 4), c4 = idx % (D / 4);
      if (idx < kAttRows * (D / 4))
        *reinterpret_cast<uint32_t*>(qc + r * kRow + c4 * 4) = (q0 + r < p.sq) ? quant_bin4(x[u], qcp) : 0u;
    }
  }
  const float* kbase = p.k + (size_t)b * p.ks[0] + (size_t)head * p.ks[1];
  float* obase = p.out + (((size_t)b * p.heads + head) * (size_t)p.sq) * (size_t)p.sk;
  float* my_stage = stage + warp * 16 * kHalfStage;
  uint32_t a[D / 32][4];
  int qs_lo = 0, qs_hi = 0;
  for (int c0 = 0; c0 < p.sk; c0 += KC) {
    const int rows = min(KC, (p.sk - c0 + kAttRows - 1) / kAttRows * kAttRows);   // key rows of this chunk, whole 128-key steps
    if (c0 > 0) __syncthreads();   // every warp is through with the previous chunk's bins
    for (int base = tid; base < rows * (D / 4); base += 8 * kAttThreads) {
      float4 x[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int idx = base + u * kAttThreads, r = idx / (D / 4), c4 = idx % (D / 4);
        x[u] = (idx < rows * (D / 4) && c0 + r < p.sk) ? ldg_stream(reinterpret_cast<const float4*>(kbase + (size_t)(c0 + r) * p.ks[2]) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int idx = base + u * kAttThreads, r = idx / (D / 4), c4 = idx % (D / 4);
        if (idx < rows * (D / 4))
          *reinterpret_cast<uint32_t*>(kc + r * kRow + c4 * 4) = (c0 + r < p.sk) ? quant_bin4(x[u], kcp) : 0u;
      }
    }
    __syncthreads();
    for (int r = tid; r < rows; r += kAttThreads) {
      int s = 0;
#pragma unroll
      for (int c = 0; c < D / 4; ++c) s = bytesum(*reinterpret_cast<const uint32_t*>(kc + r * kRow + c * 4), s);
      ksum[r] = kconst - zcq * s;
      kmask[r] = (p.mask != nullptr && c0 + r < p.sk) ? __ldg(p.mask + (size_t)b * p.sk + c0 + r) : -0.0f;
    }
    if (c0 == 0) {
      if (tid < kAttRows) {
        int s = 0;
#pragma unroll
        for (int c = 0; c < D / 4; ++c) s = bytesum(*reinterpret_cast<const uint32_t*>(qc + tid * kRow + c * 4), s);
        qsum[tid] = s;
      }
      // this warp's A fragments stay in registers for every key step
#pragma unroll
      for (int ks = 0; ks < D / 32; ++ks)   // tiles: rows 0-7 / 8-15 at k 0-15, rows 0-7 / 8-15 at k 16-31
        ldsm_x4(a[ks][0], a[ks][1], a[ks][2], a[ks][3], qc + (warp * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * kRow + ks * 32 + (lane >> 4) * 16);
    }
    __syncthreads();
    if (c0 == 0) { qs_lo = -zck * qsum[warp * 16 + g]; qs_hi = -zck * qsum[warp * 16 + g + 8]; }   // per row: -Zk * (sum of the row's bins)
    if (q0 + warp * 16 >= p.sq) continue;   // this warp's rows do not exist (it still takes part in the chunk barriers)
    for (int kt = 0; kt < rows / kAttRows; ++kt) {
      const int kl = kt * kAttRows;          // first key of the step inside the chunk
      const int k0 = c0 + kl;
      int acc[16][4];
#pragma unroll
      for (int nt = 0; nt < 16; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0; }
#pragma unroll
      for (int ks = 0; ks < D / 32; ++ks) {
#pragma unroll
        for (int nt = 0; nt < 16; nt += 2) {   // tiles: keys of n-tile nt at k 0-15 / 16-31, keys of n-tile nt + 1 at k 0-15 / 16-31
          uint32_t b0, b1, b2, b3;
          ldsm_x4(b0, b1, b2, b3, kc + (kl + (nt + (lane >> 4)) * 8 + (lane & 7)) * kRow + ks * 32 + ((lane >> 3) & 1) * 16);
          mma_u8(acc[nt], a[ks][0], a[ks][1], a[ks][2], a[ks][3], b0, b1);
          mma_u8(acc[nt + 1], a[ks][0], a[ks][1], a[ks][2], a[ks][3], b2, b3);
        }
      }
      // ---- epilogue, one 64-column half at a time: exact integer zero-point corrections, one scale, optional 1/sqrt(d) and mask
#pragma unroll
      for (int half = 0; half < 2; ++half) {
#pragma unroll
        for (int n8 = 0; n8 < 8; ++n8) {
          const int nt = half * 8 + n8;
          const int c = nt * 8 + 2 * t;        // column inside the 128-key step
          const int2 kk = *reinterpret_cast<const int2*>(ksum + kl + c);
          const float2 mm = *reinterpret_cast<const float2*>(kmask + kl + c);
          // all terms are integers below 2^24: the sum is exact; then the reference's roundings: * (s_q s_k), * 1/sqrt(d), + mask
          auto fin = [&](int v, int rc, int cc, float m) -> float {
            return __fadd_rn(__fmul_rn(__fmul_rn((float)(v + rc + cc), sqk), p.out_mul), m);
          };
          const int cs = n8 * 8 + 2 * t;       // column inside the staged half
          *reinterpret_cast<float2*>(my_stage + g * kHalfStage + cs) = make_float2(fin(acc[nt][0], qs_lo, kk.x, mm.x), fin(acc[nt][1], qs_lo, kk.y, mm.y));
          *reinterpret_cast<float2*>(my_stage + (g + 8) * kHalfStage + cs) = make_float2(fin(acc[nt][2], qs_hi, kk.x, mm.x), fin(acc[nt][3], qs_hi, kk.y, mm.y));
        }
        __syncwarp();
        const int col = k0 + half * 64 + (lane & 15) * 4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = 2 * i + (lane >> 4);
          const int row = q0 + warp * 16 + r;
          if (row < p.sq && col < p.sk)   // sk % 4 == 0: a float4 is inside or outside as a whole
            __stcs(reinterpret_cast<float4*>(obase + (size_t)row * p.sk + col), *reinterpret_cast<const float4*>(my_stage + r * kHalfStage + (lane & 15) * 4));
        }
        __syncwarp();
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K9: context.  CTA = (128 query rows, head, batch).  The value bins of the head (up to 512 keys per chunk) are quantised ONCE into
// shared memory, transposed (the B operand wants the contraction index contiguous); after that every warp streams its own 16
// rows of probabilities, 64 keys per step, with no block-wide barrier in the loop.
// ------------------------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(kAttThreads, 2)
attn_context_kernel(const ContextParams p) {
  constexpr int kKT = 64;                  // keys per step
  constexpr int kRowP = kKT + kPad;        // bytes per probability bin row
  constexpr int kStage = D + 8;
  extern __shared__ __align__(16) uint8_t sm_raw[];
  const int KC = p.kc;                                   // keys resident per chunk (multiple of 64)
  const int kRowV = KC + kPad;                           // bytes per transposed value row
  uint8_t* pc = sm_raw;                                  // [128][kRowP]
  uint8_t* vT = pc + kAttRows * kRowP;                   // [D][kRowV]
  int* psum = reinterpret_cast<int*>(vT + (size_t)D * kRowV);  // [128]
  int* vsum = psum + kAttRows;                           // [D]
  float* stage = reinterpret_cast<float*>(vsum + D);     // [8][16][kStage]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int head = blockIdx.y, b = blockIdx.z, q0 = blockIdx.x * kAttRows;
  const bool wb = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 0;
  const QParam pq = load_qparam(p.pq.scale, p.pq.zp, p.pq.zp_is_int32, p.pq.g, p.pq.qmin, p.pq.qmax, wb);
  const QParam vq = load_qparam(p.vq.scale, p.vq.zp, p.vq.zp_is_int32, p.vq.g, p.vq.qmin, p.vq.qmax, wb);
  const ConvParam pcp = make_conv_param(pq.s, pq.z, p.pq.qmin, p.pq.qmax), vcp = make_conv_param(vq.s, vq.z, p.vq.qmin, p.vq.qmax);
  const int zcp = (int)pcp.zc, zcv = (int)vcp.zc;
  const float spv = __fmul_rn(pq.s, vq.s);

  const float* pbase = p.probs + (((size_t)b * p.heads + head) * (size_t)p.sq) * (size_t)p.sk;
  const float* vbase = p.v + (size_t)b * p.vs[0] + (size_t)head * p.vs[1];
  int acc[D / 8][4];
#pragma unroll
  for (int nt = 0; nt < D / 8; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0; }
  // row sums of the probability bins (zero-point correction) come out of the tensor core: one extra n-tile whose B fragment is all
  // ones accumulates sum_k p[row][k] in every column -- no cross-lane reduction in the conversion loop
  int acc1[4] = {0, 0, 0, 0};
  int vcol = 0;                              // threads < D: running bin sum of channel tid
  const bool warp_has_rows = q0 + warp * 16 < p.sq;
  for (int c0 = 0; c0 < p.sk; c0 += KC) {
    const int keys = min(KC, (p.sk - c0 + kKT - 1) / kKT * kKT);   // keys of this chunk, whole 64-key steps
    if (c0 > 0) __syncthreads();   // every warp is through with the previous chunk's value bins
    // values: four keys x four channels per thread, bins transposed in registers, one word (four keys of a channel) per store;
    // consecutive lanes take consecutive key quads (conflict-free stores; the 16-byte loads of a key's neighbours come from L1)
    for (int idx = tid; idx < (keys / 4) * (D / 4); idx += kAttThreads) {
      const int key4 = idx % (keys / 4), c4 = idx / (keys / 4);
      uint32_t w[4];
      float4 xv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int key = c0 + key4 * 4 + i;
        xv[i] = key < p.sk ? __ldg(reinterpret_cast<const float4*>(vbase + (size_t)key * p.vs[2]) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) w[i] = (c0 + key4 * 4 + i < p.sk) ? quant_bin4(xv[i], vcp) : 0u;
      const uint32_t t0 = __byte_perm(w[0], w[1], 0x5140), t1 = __byte_perm(w[2], w[3], 0x5140);
      const uint32_t t2 = __byte_perm(w[0], w[1], 0x7362), t3 = __byte_perm(w[2], w[3], 0x7362);
      uint8_t* dst = vT + (size_t)(c4 * 4) * kRowV + key4 * 4;
      *reinterpret_cast<uint32_t*>(dst) = __byte_perm(t0, t1, 0x5410);
      *reinterpret_cast<uint32_t*>(dst + kRowV) = __byte_perm(t0, t1, 0x7632);
      *reinterpret_cast<uint32_t*>(dst + 2 * kRowV) = __byte_perm(t2, t3, 0x5410);
      *reinterpret_cast<uint32_t*>(dst + 3 * kRowV) = __byte_perm(t2, t3, 0x7632);
    }
    __syncthreads();
    if (tid < D) {
      for (int c = 0; c < keys / 4; ++c) vcol = bytesum(*reinterpret_cast<const uint32_t*>(vT + (size_t)tid * kRowV + c * 4), vcol);
    }
    if (!warp_has_rows) continue;
    // probabilities: this warp's 16 rows x 64 keys per step, two rows per 128-bit load instruction.  The rows travel in two halves
    // of four loads each, and the next step's half is requested as soon as the current one has been converted: loads stay in
    // flight while the other half is converted and the fragments are multiplied (a warp alone would alternate between
    // waiting for DRAM and converting).
    const int n_steps = keys / kKT;
    const int prow_l = lane >> 4, pc4 = lane & 15;
    // one base pointer per lane (row 0 of its row pair, its 16-byte column of step 0); rows advance by 2 * sk floats, steps by 64
    const float* lane_base = pbase + (size_t)(q0 + warp * 16 + prow_l) * (size_t)p.sk + (size_t)(c0 + pc4 * 4);
    const uint32_t row_step = 2u * (uint32_t)p.sk;
    uint8_t* lane_pc = pc + (warp * 16 + prow_l) * kRowP + pc4 * 4;
    // whole tile inside the matrix (the usual case): no per-load predicates
    const bool rows_full = q0 + warp * 16 + 16 <= p.sq;
    auto load_half = [&](float4 (&x)[4], int kt, int half) {
      const float* src = lane_base + kt * kKT + (uint32_t)(half * 4) * row_step;
      if (rows_full && c0 + kt * kKT + kKT <= p.sk) {
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = ldg_stream(reinterpret_cast<const float4*>(src + (uint32_t)i * row_step));
      } else {
        const int key = c0 + kt * kKT + pc4 * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = q0 + warp * 16 + 2 * (half * 4 + i) + prow_l;
          x[i] = (row < p.sq && key < p.sk) ? ldg_stream(reinterpret_cast<const float4*>(src + (uint32_t)i * row_step)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    // Rows / keys outside the matrix were loaded as zeros and must contribute bin 0 (not the bin of 0.0): masked after the
    // conversion.  The conversion itself is branch-free; groups near a rounding tie (~0.1 %) are redone exactly afterwards.
    auto convert_half = [&](const float4 (&x)[4], int kt, int half) {
      const int key = c0 + kt * kKT + pc4 * 4;
      const bool key_ok = key < p.sk;
      uint32_t w[4];
      uint32_t risky_mask = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        bool risky;
        w[i] = quant_bin4_fast(x[i], pcp, risky);
        risky_mask |= (uint32_t)risky << i;
      }
      if (risky_mask != 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (risky_mask & (1u << i)) w[i] = quant_bin4_exact(x[i].x, x[i].y, x[i].z, x[i].w, pcp.s, pcp.zc, pcp.span);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r2 = half * 4 + i;
        if (!(rows_full && key_ok)) { if (!(key_ok && q0 + warp * 16 + 2 * r2 + prow_l < p.sq)) w[i] = 0; }
        *reinterpret_cast<uint32_t*>(lane_pc + (2 * r2) * kRowP) = w[i];
      }
    };
    float4 xa[4], xb[4];
    load_half(xa, 0, 0);
    load_half(xb, 0, 1);
    for (int kt = 0; kt < n_steps; ++kt) {
      const int kl = kt * kKT;
      convert_half(xa, kt, 0);
      if (kt + 1 < n_steps) load_half(xa, kt + 1, 0);
      convert_half(xb, kt, 1);
      if (kt + 1 < n_steps) load_half(xb, kt + 1, 1);
      __syncwarp();
#pragma unroll
      for (int ks = 0; ks < kKT / 32; ++ks) {
        uint32_t a0, a1, a2, a3;
        ldsm_x4(a0, a1, a2, a3, pc + (warp * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * kRowP + ks * 32 + (lane >> 4) * 16);
#pragma unroll
        for (int nt = 0; nt < D / 8; nt += 2) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4(b0, b1, b2, b3, vT + (size_t)((nt + (lane >> 4)) * 8 + (lane & 7)) * kRowV + kl + ks * 32 + ((lane >> 3) & 1) * 16);
          mma_u8(acc[nt], a0, a1, a2, a3, b0, b1);
          mma_u8(acc[nt + 1], a0, a1, a2, a3, b2, b3);
        }
        mma_u8(acc1, a0, a1, a2, a3, 0x01010101u, 0x01010101u);
      }
      __syncwarp();   // the fragments are consumed before the next step overwrites this warp's probability bins
    }
  }
  // ---- epilogue
  if (tid < D) vsum[tid] = vcol;
  __syncthreads();
  const int kconst = p.sk * zcp * zcv;
  float* my_stage = stage + warp * 16 * kStage;
  const int ps_lo = acc1[0], ps_hi = acc1[2];   // rows g and g + 8 of this warp (every column of the ones tile holds the row sum)
  (void)psum;
#pragma unroll
  for (int nt = 0; nt < D / 8; ++nt) {
    const int c = nt * 8 + 2 * t;
    const int vs0 = vsum[c], vs1 = vsum[c + 1];
    auto fin = [&](int v, int ps_, int vs_) -> float { return __fmul_rn((float)(v - zcv * ps_ - zcp * vs_ + kconst), spv); };
    *reinterpret_cast<float2*>(my_stage + g * kStage + c) = make_float2(fin(acc[nt][0], ps_lo, vs0), fin(acc[nt][1], ps_lo, vs1));
    *reinterpret_cast<float2*>(my_stage + (g + 8) * kStage + c) = make_float2(fin(acc[nt][2], ps_hi, vs0), fin(acc[nt][3], ps_hi, vs1));
  }
  __syncwarp();
  QParam oq = QParam{1.f, 0.f};
  float o_rinv = 1.f;
  if (p.has_oq) {
    oq = load_qparam(p.oq.scale, p.oq.zp, p.oq.zp_is_int32, p.oq.g, p.oq.qmin, p.oq.qmax, wb);
    o_rinv = __frcp_rn(oq.s);
  }
  constexpr int kLanesPerRow = D / 4;                 // 16 for d = 64
  constexpr int kRowsPerInstr = 32 / kLanesPerRow;    // 2
  static_assert(32 % kLanesPerRow == 0, "head size must be 32, 64 or 128");
  const int c4 = lane % kLanesPerRow;
#pragma unroll 4
  for (int i = 0; i < 16 / kRowsPerInstr; ++i) {
    const int r = i * kRowsPerInstr + lane / kLanesPerRow;
    const int row = q0 + warp * 16 + r;
    if (row < p.sq) {
      float4 o = *reinterpret_cast<const float4*>(my_stage + r * kStage + c4 * 4);
      const size_t off = (((size_t)b * p.sq + row) * (size_t)p.heads + head) * (size_t)D + (size_t)c4 * 4;
      if (p.has_oq) {
        float q0_, q1_, q2_, q3_;
        bool k0_, k1_, k2_, k3_;
        float4 y;
        y.x = fq_elem_fast(o.x, oq.s, o_rinv, oq.z, p.oq.qmin, p.oq.qmax, q0_, k0_);
        y.y = fq_elem_fast(o.y, oq.s, o_rinv, oq.z, p.oq.qmin, p.oq.qmax, q1_, k1_);
        y.z = fq_elem_fast(o.z, oq.s, o_rinv, oq.z, p.oq.qmin, p.oq.qmax, q2_, k2_);
        y.w = fq_elem_fast(o.w, oq.s, o_rinv, oq.z, p.oq.qmin, p.oq.qmax, q3_, k3_);
        if (k0_ | k1_ | k2_ | k3_) {
          y.x = fq_elem(o.x, oq.s, oq.z, p.oq.qmin, p.oq.qmax, q0_); y.y = fq_elem(o.y, oq.s, oq.z, p.oq.qmin, p.oq.qmax, q1_);
          y.z = fq_elem(o.z, oq.s, oq.z, p.oq.qmin, p.oq.qmax, q2_); y.w = fq_elem(o.w, oq.s, oq.z, p.oq.qmin, p.oq.qmax, q3_);
        }
        if (p.bins != nullptr)
          *reinterpret_cast<uint32_t*>(p.bins + off) = (uint32_t)__float2int_rn(q0_ - p.oq.qmin) | ((uint32_t)__float2int_rn(q1_ - p.oq.qmin) << 8) |
                                                       ((uint32_t)__float2int_rn(q2_ - p.oq.qmin) << 16) | ((uint32_t)__float2int_rn(q3_ - p.oq.qmin) << 24);
        o = y;
      }
      *reinterpret_cast<float4*>(p.out + off) = o;
    }
  }
}

static int check_quantizer(const osq_quantizer_t* q, const char* what) {
  OSQ_CHECK_ARG(q != nullptr && q->scale != nullptr && q->zero_point != nullptr, "%s: null quantizer", what);
  OSQ_CHECK_ARG(q->qmin < q->qmax && q->qmax - q->qmin <= 255, "%s: the quantizer must have at most 8 bits", what);
  OSQ_CHECK_ARG(!(q->lsq_grad_factor > 0.f && q->zp_is_int32), "%s: LSQ+ needs a float zero_point", what);
  return OSQ_OK;
}
static AttnQ to_attnq(const osq_quantizer_t* q) {
  AttnQ a;
  a.scale = q->scale; a.zp = q->zero_point; a.zp_is_int32 = q->zp_is_int32; a.g = q->lsq_grad_factor; a.qmin = (float)q->qmin; a.qmax = (float)q->qmax;
  return a;
}

}  // namespace osq

extern "C" {

int osq_attn_scores_fq_f32(const float* q, const float* k, int64_t batch, int64_t heads, int64_t sq, int64_t sk, int64_t d,
                           const int64_t* q_strides, const int64_t* k_strides, const osq_quantizer_t* qq, const osq_quantizer_t* kq,
                           float out_mul, const float* mask, float* scores, void* stream) {
  using namespace osq;
  OSQ_CHECK_ARG(batch >= 0 && heads > 0 && sq >= 0 && sk >= 0, "osq_attn_scores_fq_f32: bad shape");
  if (batch == 0 || sq == 0 || sk == 0) return OSQ_OK;
  OSQ_CHECK_ARG(q && k && scores && q_strides && k_strides, "osq_attn_scores_fq_f32: null pointer");
  OSQ_CHECK_ARG(d == 32 || d == 64 || d == 128, "osq_attn_scores_fq_f32: head size must be 32, 64 or 128");
  OSQ_CHECK_ARG(sk % 4 == 0, "osq_attn_scores_fq_f32: the key length must be a multiple of 4");
  OSQ_CHECK_ARG(heads <= 65535 && batch <= 65535, "osq_attn_scores_fq_f32: too many heads / batches for one grid");
  if (int rc = check_quantizer(qq, "osq_attn_scores_fq_f32 (q)")) return rc;
  if (int rc = check_quantizer(kq, "osq_attn_scores_fq_f32 (k)")) return rc;
  for (int i = 0; i < 3; ++i)
    OSQ_CHECK_ARG(q_strides[i] % 4 == 0 && k_strides[i] % 4 == 0, "osq_attn_scores_fq_f32: strides must be multiples of 4 elements");
  OSQ_CHECK_ARG((((uintptr_t)q | (uintptr_t)k | (uintptr_t)scores) & 15) == 0, "osq_attn_scores_fq_f32: pointers must be 16-byte aligned");
  ScoresParams p;
  p.q = q; p.k = k;
  for (int i = 0; i < 3; ++i) { p.qs[i] = q_strides[i]; p.ks[i] = k_strides[i]; }
  p.heads = (int)heads; p.sq = (int)sq; p.sk = (int)sk;
  p.qq = to_attnq(qq); p.kq = to_attnq(kq);
  p.out_mul = out_mul; p.mask = mask; p.out = scores;
  const dim3 grid((unsigned)((sq + kAttRows - 1) / kAttRows), (unsigned)heads, (unsigned)batch);
  p.kc = (int)((sk + kAttRows - 1) / kAttRows * kAttRows < 512 ? (sk + kAttRows - 1) / kAttRows * kAttRows : 512);
  const size_t smem = (size_t)(kAttRows + p.kc) * (size_t)(d + kPad) + (size_t)(kAttRows + 2 * p.kc) * sizeof(int) + 8 * 16 * kHalfStage * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_set[64] = {false};
  int dev = 0;
  OSQ_CUDA(cudaGetDevice(&dev));
  if (!attr_set[dev & 63]) {
    OSQ_CUDA(cudaFuncSetAttribute(attn_scores_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    OSQ_CUDA(cudaFuncSetAttribute(attn_scores_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    OSQ_CUDA(cudaFuncSetAttribute(attn_scores_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_set[dev & 63] = true;
  }
  if (d == 32) attn_scores_kernel<32><<<grid, kAttThreads, smem, st>>>(p);
  else if (d == 64) attn_scores_kernel<64><<<grid, kAttThreads, smem, st>>>(p);
  else attn_scores_kernel<128><<<grid, kAttThreads, smem, st>>>(p);
  OSQ_LAUNCH_CHECK();
  return OSQ_OK;
}

int osq_attn_context_fq_f32(const float* probs, const float* v, int64_t batch, int64_t heads, int64_t sq, int64_t sk, int64_t d,
                            const int64_t* v_strides, const osq_quantizer_t* pq, const osq_quantizer_t* vq, const osq_quantizer_t* oq,
                            float* context, uint8_t* bins, void* stream) {
  using namespace osq;
  OSQ_CHECK_ARG(batch >= 0 && heads > 0 && sq >= 0 && sk >= 0, "osq_attn_context_fq_f32: bad shape");
  if (batch == 0 || sq == 0) return OSQ_OK;
  OSQ_CHECK_ARG(probs && v && context && v_strides, "osq_attn_context_fq_f32: null pointer");
  OSQ_CHECK_ARG(d == 32 || d == 64 || d == 128, "osq_attn_context_fq_f32: head size must be 32, 64 or 128");
  OSQ_CHECK_ARG(sk % 4 == 0 && sk > 0, "osq_attn_context_fq_f32: the key length must be a positive multiple of 4");
  OSQ_CHECK_ARG(heads <= 65535 && batch <= 65535, "osq_attn_context_fq_f32: too many heads / batches for one grid");
  if (int rc = check_quantizer(pq, "osq_attn_context_fq_f32 (probs)")) return rc;
  if (int rc = check_quantizer(vq, "osq_attn_context_fq_f32 (v)")) return rc;
  if (oq != nullptr) { if (int rc = check_quantizer(oq, "osq_attn_context_fq_f32 (output)")) return rc; }
  OSQ_CHECK_ARG(oq != nullptr || bins == nullptr, "osq_attn_context_fq_f32: bins need the output quantizer");
  for (int i = 0; i < 3; ++i) OSQ_CHECK_ARG(v_strides[i] % 4 == 0, "osq_attn_context_fq_f32: strides must be multiples of 4 elements");
  OSQ_CHECK_ARG((((uintptr_t)probs | (uintptr_t)v | (uintptr_t)context) & 15) == 0 && (((uintptr_t)bins) & 3) == 0,
                "osq_attn_context_fq_f32: pointers must be 16-byte aligned (bins: 4)");
  ContextParams p;
  p.probs = probs; p.v = v;
  for (int i = 0; i < 3; ++i) p.vs[i] = v_strides[i];
  p.heads = (int)heads; p.sq = (int)sq; p.sk = (int)sk;
  p.pq = to_attnq(pq); p.vq = to_attnq(vq);
  p.has_oq = oq != nullptr ? 1 : 0;
  p.oq = oq != nullptr ? to_attnq(oq) : p.pq;
  p.out = context; p.bins = bins;
  const dim3 grid((unsigned)((sq + kAttRows - 1) / kAttRows), (unsigned)heads, (unsigned)batch);
  p.kc = (int)((sk + 63) / 64 * 64 < 512 ? (sk + 63) / 64 * 64 : 512);
  const size_t smem = (size_t)kAttRows * (64 + kPad) + (size_t)d * (size_t)(p.kc + kPad) + (kAttRows + d) * sizeof(int) + 8 * 16 * (size_t)(d + 8) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_set[64] = {false};
  int dev = 0;
  OSQ_CUDA(cudaGetDevice(&dev));
  if (!attr_set[dev & 63]) {
    OSQ_CUDA(cudaFuncSetAttribute(attn_context_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    OSQ_CUDA(cudaFuncSetAttribute(attn_context_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    OSQ_CUDA(cudaFuncSetAttribute(attn_context_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_set[dev & 63] = true;
  }
  if (d == 32) attn_context_kernel<32><<<grid, kAttThreads, smem, st>>>(p);
  else if (d == 64) attn_context_kernel<64><<<grid, kAttThreads, smem, st>>>(p);
  else attn_context_kernel<128><<<grid, kAttThreads, smem, st>>>(p);
  OSQ_LAUNCH_CHECK();
  return OSQ_OK;
}

}  // extern "C"
