// layernorm_fq.cu -- K7: GammaResidual + LayerNorm + the LayerNorm's output quantizer as ONE pass over the row.
//
//   u   = res * res_gamma + h                (model/util_layernorm.py:41-52, GammaResidual; res_gamma optional)
//   ln  = (u - mean(u)) * rsqrt(var(u) + eps) [* weight] [+ bias]   (util_layernorm.py:14-15 / :34-36, nn.LayerNorm)
//   y   = fq(ln)   + optional u8 bins         (util_layernorm.py:16-17 -> util_quant.py:11-15 / :48-55, exactly as K1)
//
// replaces three launches (residual add 12 B / element, LayerNorm 8, fake-quant 9) by one of 13 B / element
// (8 read, 4 + 1 written).  HBM-bound: one warp per row, the row lives in registers (hidden <= 1024) or is re-read from
// L2 (larger rows), two-pass mean / variance in fp32, 128-bit streaming loads and stores.
//
// Parity: the LayerNorm part is ordinary fp32 arithmetic (no implementation reproduces another's summation order bit for
// bit, torch's CPU and CUDA kernels differ from each other too): |ln - ln_ref| <= 2e-6 * max|ln|.  The quantizer part is
// bit-exact: (y, bins) == K1(ln) for the ln this kernel computes (`ln_out`, optional side output, is what the tests feed K1).
#include "common.cuh"

namespace osq {

constexpr int kLnThreads = 256;   // 8 rows per CTA
constexpr int kLnMaxVec = 8;      // float4 per lane held in registers: hidden <= 32 * 4 * 8 = 1024

struct LnFqParams {
  const float* h;
  const float* res;        // optional
  const float* res_gamma;  // optional [H]
  const float* weight;     // optional [H]
  const float* bias;       // optional [H]
  float* y;
  uint8_t* bins;           // optional
  float* ln_out;           // optional
  int64_t rows;
  int H;
  float eps;
  const float* scale;
  const void* zp;
  int zp_is_int32;
  float g, qmin, qmax;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// the quantizer on four adjacent elements: K1's division-free fast path, the whole group redone exactly near a tie
__device__ __forceinline__ float4 fq4(const float4 a, float s, float rinv, float z, float qmin, float qmax, uint32_t& word) {
  float4 o;
  float q0, q1, q2, q3;
  bool k0, k1, k2, k3;
  o.x = fq_elem_fast(a.x, s, rinv, z, qmin, qmax, q0, k0);
  o.y = fq_elem_fast(a.y, s, rinv, z, qmin, qmax, q1, k1);
  o.z = fq_elem_fast(a.z, s, rinv, z, qmin, qmax, q2, k2);
  o.w = fq_elem_fast(a.w, s, rinv, z, qmin, qmax, q3, k3);
  if (k0 | k1 | k2 | k3) {
    o.x = fq_elem(a.x, s, z, qmin, qmax, q0);
    o.y = fq_elem(a.y, s, z, qmin, qmax, q1);
    o.z = fq_elem(a.z, s, z, qmin, qmax, q2);
    o.w = fq_elem(a.w, s, z, qmin, qmax, q3);
  }
  word = (uint32_t)__float2int_rn(q0 - qmin) | ((uint32_t)__float2int_rn(q1 - qmin) << 8) | ((uint32_t)__float2int_rn(q2 - qmin) << 16) | ((uint32_t)__float2int_rn(q3 - qmin) << 24);
  return o;
}

// kVec = float4 per lane (H = 128 * kVec exactly); kVec = 0: generic H (multiple of 4), three sweeps over the row (the second and
// third hit L1 / L2)
template <int kVec>
__global__ void __launch_bounds__(kLnThreads, 4)   // <= 64 registers: four CTAs (32 warps) per SM hide the DRAM latency of the row loads
residual_layernorm_fq_kernel(const LnFqParams p) {
  const QParam qp = load_qparam(p.scale, p.zp, p.zp_is_int32, p.g, p.qmin, p.qmax, blockIdx.x == 0 && threadIdx.x == 0);
  const float s = qp.s, z = qp.z, rinv = __frcp_rn(qp.s);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wpc = kLnThreads / 32;
  const float inv_h = 1.f / (float)p.H;
  for (int64_t row = (int64_t)blockIdx.x * wpc + warp; row < p.rows; row += (int64_t)gridDim.x * wpc) {
    const size_t off = (size_t)row * (size_t)p.H;
    const float4* h4 = reinterpret_cast<const float4*>(p.h + off);
    const float4* r4 = p.res ? reinterpret_cast<const float4*>(p.res + off) : nullptr;
    auto load_u = [&](int c) -> float4 {   // c = float4 index inside the row
      float4 u = ldg_stream(h4 + c);
      if (r4 != nullptr) {
        float4 r = ldg_stream(r4 + c);
        if (p.res_gamma != nullptr) {
          const float4 gm = __ldg(reinterpret_cast<const float4*>(p.res_gamma) + c);
          r.x = __fmul_rn(r.x, gm.x); r.y = __fmul_rn(r.y, gm.y); r.z = __fmul_rn(r.z, gm.z); r.w = __fmul_rn(r.w, gm.w);
        }
        u.x = __fadd_rn(r.x, u.x); u.y = __fadd_rn(r.y, u.y); u.z = __fadd_rn(r.z, u.z); u.w = __fadd_rn(r.w, u.w);
      }
      return u;
    };
    auto finish = [&](int c, const float4 u, float mean, float rstd) {
      float4 n;
      n.x = __fmul_rn(__fsub_rn(u.x, mean), rstd); n.y = __fmul_rn(__fsub_rn(u.y, mean), rstd);
      n.z = __fmul_rn(__fsub_rn(u.z, mean), rstd); n.w = __fmul_rn(__fsub_rn(u.w, mean), rstd);
      if (p.weight != nullptr) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(p.weight) + c);
        n.x = __fmul_rn(n.x, w.x); n.y = __fmul_rn(n.y, w.y); n.z = __fmul_rn(n.z, w.z); n.w = __fmul_rn(n.w, w.w);
      }
      if (p.bias != nullptr) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias) + c);
        n.x = __fadd_rn(n.x, b.x); n.y = __fadd_rn(n.y, b.y); n.z = __fadd_rn(n.z, b.z); n.w = __fadd_rn(n.w, b.w);
      }
      if (p.ln_out != nullptr) __stcs(reinterpret_cast<float4*>(p.ln_out + off) + c, n);
      uint32_t word;
      const float4 o = fq4(n, s, rinv, z, p.qmin, p.qmax, word);
      __stcs(reinterpret_cast<float4*>(p.y + off) + c, o);
      if (p.bins != nullptr) *reinterpret_cast<uint32_t*>(p.bins + off + (size_t)c * 4) = word;
    };
    if constexpr (kVec > 0) {
      float4 u[kVec];
      float sum = 0.f;
#pragma unroll
      for (int v = 0; v < kVec; ++v) {
        u[v] = load_u(v * 32 + lane);
        sum += (u[v].x + u[v].y) + (u[v].z + u[v].w);
      }
      const float mean = warp_sum(sum) * inv_h;
      float sq = 0.f;
#pragma unroll
      for (int v = 0; v < kVec; ++v) {
        const float a = u[v].x - mean, b = u[v].y - mean, c = u[v].z - mean, d = u[v].w - mean;
        sq += (a * a + b * b) + (c * c + d * d);
      }
      const float rstd = rsqrtf(warp_sum(sq) * inv_h + p.eps);
#pragma unroll
      for (int v = 0; v < kVec; ++v) finish(v * 32 + lane, u[v], mean, rstd);
    } else {
      const int nvec = p.H >> 2;
      float sum = 0.f;
      for (int c = lane; c < nvec; c += 32) { const float4 u = load_u(c); sum += (u.x + u.y) + (u.z + u.w); }
      const float mean = warp_sum(sum) * inv_h;
      float sq = 0.f;
      for (int c = lane; c < nvec; c += 32) {
        const float4 u = load_u(c);
        const float a = u.x - mean, b = u.y - mean, cc = u.z - mean, d = u.w - mean;
        sq += (a * a + b * b) + (cc * cc + d * d);
      }
      const float rstd = rsqrtf(warp_sum(sq) * inv_h + p.eps);
      for (int c = lane; c < nvec; c += 32) finish(c, load_u(c), mean, rstd);
    }
  }
}

}  // namespace osq

extern "C" {

int osq_residual_layernorm_fq_f32(const float* h, const float* res, const float* res_gamma, const float* ln_weight,
                                  const float* ln_bias, float eps, int64_t rows, int64_t hidden, const float* scale,
                                  const void* zero_point, int zp_is_int32, float lsq_grad_factor, int qmin, int qmax, float* y,
                                  uint8_t* bins, float* ln_out, void* stream) {
  using namespace osq;
  OSQ_CHECK_ARG(rows >= 0 && hidden > 0, "osq_residual_layernorm_fq_f32: bad shape");
  if (rows == 0) return OSQ_OK;
  OSQ_CHECK_ARG(h && y && scale && zero_point, "osq_residual_layernorm_fq_f32: null pointer");
  OSQ_CHECK_ARG(hidden % 4 == 0 && hidden <= (1 << 20), "osq_residual_layernorm_fq_f32: hidden must be a multiple of 4");
  OSQ_CHECK_ARG(res != nullptr || res_gamma == nullptr, "osq_residual_layernorm_fq_f32: res_gamma without res");
  OSQ_CHECK_ARG(qmin < qmax, "osq_residual_layernorm_fq_f32: qmin >= qmax");
  OSQ_CHECK_ARG(bins == nullptr || qmax - qmin <= 255, "osq_residual_layernorm_fq_f32: uint8 bins need at most 8 bits");
  OSQ_CHECK_ARG(!(lsq_grad_factor > 0.f && zp_is_int32), "osq_residual_layernorm_fq_f32: LSQ+ needs a float zero_point");
  const uintptr_t al = (uintptr_t)h | (uintptr_t)y | (uintptr_t)res | (uintptr_t)res_gamma | (uintptr_t)ln_weight | (uintptr_t)ln_bias |
                       (uintptr_t)ln_out;
  OSQ_CHECK_ARG((al & 15) == 0 && (((uintptr_t)bins) & 3) == 0, "osq_residual_layernorm_fq_f32: pointers must be 16-byte aligned (bins: 4)");
  int sms = sm_count();
  if (sms <= 0) { set_error("no CUDA device"); return OSQ_ECUDA; }
  LnFqParams p;
  p.h = h; p.res = res; p.res_gamma = res_gamma; p.weight = ln_weight; p.bias = ln_bias; p.y = y; p.bins = bins; p.ln_out = ln_out;
  p.rows = rows; p.H = (int)hidden; p.eps = eps; p.scale = scale; p.zp = zero_point; p.zp_is_int32 = zp_is_int32;
  p.g = lsq_grad_factor; p.qmin = (float)qmin; p.qmax = (float)qmax;
  const int wpc = kLnThreads / 32;
  const int64_t want = (rows + wpc - 1) / wpc;
  const int grid = (int)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);   // (one CTA per 8-row block, up to 96 per SM, measured no faster)
  cudaStream_t st = (cudaStream_t)stream;
  const int vec = (hidden % 128 == 0 && hidden / 128 <= kLnMaxVec) ? (int)(hidden / 128) : 0;
  switch (vec) {
    case 1: residual_layernorm_fq_kernel<1><<<grid, kLnThreads, 0, st>>>(p); break;
    case 2: residual_layernorm_fq_kernel<2><<<grid, kLnThreads, 0, st>>>(p); break;
    case 4: residual_layernorm_fq_kernel<4><<<grid, kLnThreads, 0, st>>>(p); break;
    case 6: residual_layernorm_fq_kernel<6><<<grid, kLnThreads, 0, st>>>(p); break;   // 768: BERT / RoBERTa base
    case 8: residual_layernorm_fq_kernel<8><<<grid, kLnThreads, 0, st>>>(p); break;   // 1024: BART-large, BERT-large
    default: residual_layernorm_fq_kernel<0><<<grid, kLnThreads, 0, st>>>(p); break;
  }
  OSQ_LAUNCH_CHECK();
  return OSQ_OK;
}

}  // extern "C"
