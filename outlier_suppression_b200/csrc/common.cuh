// common.cuh -- shared helpers for libosq_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "osq.h"

namespace osq {

void set_error(const char* fmt, ...);
int sm_count();

#define OSQ_CHECK_ARG(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      ::osq::set_error(__VA_ARGS__);        \
      return OSQ_EINVAL;                    \
    }                                       \
  } while (0)

#define OSQ_CUDA(call)                                                            \
  do {                                                                            \
    cudaError_t e__ = (call);                                                     \
    if (e__ != cudaSuccess) {                                                     \
      ::osq::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),   \
                       __FILE__, __LINE__);                                       \
      return OSQ_ECUDA;                                                           \
    }                                                                             \
  } while (0)

#define OSQ_LAUNCH_CHECK() OSQ_CUDA(cudaPeekAtLastError())

constexpr int kMaxPartialBlocks = 2048;  // upper bound on the grid of any reduction kernel
// workspace layout: float partial_min[kMax], float partial_max[kMax], uint32 ticket, pad
constexpr int64_t kSelectWsOffset = (2 * kMaxPartialBlocks) * 4 + 64;  // multi-CTA radix select state (observer.cu)
constexpr int64_t kSelectHist0Offset = kSelectWsOffset + 2 * 256 * 4 + 64;  // first-digit table of the fused prune path: 2 x 2048 u32
constexpr int64_t kTraceOffset = kSelectHist0Offset + 2 * 2048 * 4;  // 32 x int64 globaltimer stamps of the last fused prune call (profiling aid)
constexpr int64_t kSelectHist1Offset = kTraceOffset + 32 * 8;   // second first-digit table (+ stamps): osq_prune_observe_many_f32 alternates tables
constexpr int64_t kWorkspaceBytes = kSelectHist1Offset + 2 * 2048 * 4 + 32 * 8;

// ---------------------------------------------------------------------------------------------
// Quantisation parameters as the kernels consume them.
// ---------------------------------------------------------------------------------------------
struct QParam {
  float s;  // effective scale
  float z;  // effective zero point (integer valued except for LSQ+'s 1-ulp drift)
};

// (t - t*g) + t*g with the reference's op order and NO fma contraction (util_quant.py:70-71)
__device__ __forceinline__ float grad_scale_value(float t, float g) {
  float tg = __fmul_rn(t, g);
  return __fadd_rn(__fsub_rn(t, tg), tg);
}

// Loads (scale, zero_point) device scalars and derives the effective parameters.
//   g == 0  : FixedFakeQuantize -- parameters used as stored (fake_quant.py:123-125)
//   g  > 0  : LSQPlusFakeQuantize -- in-place sanitise (fake_quant.py:188-191: |s| clamped to
//             eps, z clamped to [qmin,qmax]) then round_ste + grad_scale (util_quant.py:49-51).
// `writeback` (one thread of the grid) stores the sanitised raw parameters like the reference's
// in-place ops; sanitise is idempotent so concurrent readers are unaffected.
__device__ __forceinline__ QParam load_qparam(const float* scale, const void* zp, int zp_is_int32,
                                              float g, float qmin, float qmax, bool writeback) {
  float s = *scale;
  float z = zp_is_int32 ? (float)(*(const int32_t*)zp) : *(const float*)zp;
  if (g > 0.f) {
    float s_raw = s, z_raw = z;
    s = fmaxf(fabsf(s), 1.1920928955078125e-07f);
    z = fminf(fmaxf(z, qmin), qmax);
    if (writeback && !zp_is_int32) {
      if (s != s_raw) *const_cast<float*>(scale) = s;
      if (z != z_raw) *(float*)const_cast<void*>(zp) = z;
    }
    z = grad_scale_value(rintf(z), g);
    s = grad_scale_value(s, g);
  }
  return QParam{s, z};
}

// One element of util_quant.py:12-14.  Returns y; q receives the clamped bin (fp32).
// round_ste value (rint(t) - t) + t: rint(t) for finite t, NaN for +-inf, like the reference.
__device__ __forceinline__ float fq_elem(float x, float s, float z, float qmin, float qmax, float& q) {
  float t = __fdiv_rn(x, s);
  float r = rintf(t);
  r = __fadd_rn(__fsub_rn(r, t), t);
  float v = __fadd_rn(r, z);
  q = (v != v) ? v : fminf(fmaxf(v, qmin), qmax);  // torch.clamp propagates NaN
  return __fmul_rn(__fsub_rn(q, z), s);
}

// Division-free fast path of fq_elem.  t' = x * (1/s) differs from the true quotient by a few ulps; when t' is farther than
// 1e-4 from a rounding tie and small enough for that bound to hold (|t'| < 300), rint(x / s) == rint(t') and the rest of
// util_quant.py:12-14 is evaluated exactly as fq_elem does.  `risky` (near a tie, huge, inf, NaN: ~2e-4 of the elements)
// sends the caller to the exact division.  The IEEE division is 40 % of K1's instructions, and K1 is ALU-bound.
__device__ __forceinline__ float fq_elem_fast(float x, float s, float rinv, float z, float qmin, float qmax, float& q, bool& risky) {
  const float t = __fmul_rn(x, rinv);
  const float r = rintf(t);
  risky = !(fabsf(__fsub_rn(t, r)) < 0.4999f) || !(fabsf(t) < 300.f);
  const float v = __fadd_rn(r, z);
  q = fminf(fmaxf(v, qmin), qmax);
  return __fmul_rn(__fsub_rn(q, z), s);
}


constexpr float kMagic = 12582912.f;  // 1.5 * 2^23: fp32 ulp is 1 in [2^23, 2^24)

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


__device__ __forceinline__ uint32_t pack4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}

// four bins -> one packed word.  The fast path of the four elements is straight-line code (full ILP);
// ONE branch per float4 sends the whole group through the exact path when any element is near a tie.
static __device__ __noinline__ uint32_t quant_bin4_exact(float x0, float x1, float x2, float x3, float s, float zc, float span) {
  // exact path (true division), taken by ~0.1 % of the float4 groups: kept out of line so the unrolled
  // conversion loop stays a few KB of straight-line code
  float v0 = fminf(fmaxf(__fadd_rn(rintf(__fdiv_rn(x0, s)), zc), 0.f), span);
  float v1 = fminf(fmaxf(__fadd_rn(rintf(__fdiv_rn(x1, s)), zc), 0.f), span);
  float v2 = fminf(fmaxf(__fadd_rn(rintf(__fdiv_rn(x2, s)), zc), 0.f), span);
  float v3 = fminf(fmaxf(__fadd_rn(rintf(__fdiv_rn(x3, s)), zc), 0.f), span);
  return (uint32_t)v0 | ((uint32_t)v1 << 8) | ((uint32_t)v2 << 16) | ((uint32_t)v3 << 24);
}

__device__ __forceinline__ uint32_t quant_bin4(const float4 x, const ConvParam& c) {
  float u0 = fmaf(x.x, c.rinv, c.mz), u1 = fmaf(x.y, c.rinv, c.mz), u2 = fmaf(x.z, c.rinv, c.mz), u3 = fmaf(x.w, c.rinv, c.mz);
  const float e0 = fmaf(x.x, c.rinv, -__fsub_rn(u0, c.mz)), e1 = fmaf(x.y, c.rinv, -__fsub_rn(u1, c.mz));
  const float e2 = fmaf(x.z, c.rinv, -__fsub_rn(u2, c.mz)), e3 = fmaf(x.w, c.rinv, -__fsub_rn(u3, c.mz));
  const float worst = fmaxf(fmaxf(fabsf(e0), fabsf(e1)), fmaxf(fabsf(e2), fabsf(e3)));
  const bool nan_in = (e0 != e0) | (e1 != e1) | (e2 != e2) | (e3 != e3);
  if (!(worst <= 0.4999f) | nan_in) return quant_bin4_exact(x.x, x.y, x.z, x.w, c.s, c.zc, c.span);
  u0 = fminf(fmaxf(u0, c.lo), c.hi); u1 = fminf(fmaxf(u1, c.lo), c.hi);
  u2 = fminf(fmaxf(u2, c.lo), c.hi); u3 = fminf(fmaxf(u3, c.lo), c.hi);
  return pack4(__float_as_uint(u0), __float_as_uint(u1), __float_as_uint(u2), __float_as_uint(u3));
}

// branch-free fast path for one float4: returns the packed bins and sets `risky` when any element is within
// 1e-4 of a rounding tie (or huge / NaN): those groups (~0.1 %) are redone with the exact division AFTER the
// unrolled loop, so the hot loop is straight-line code that the scheduler can interleave across all rows
__device__ __forceinline__ uint32_t quant_bin4_fast(const float4 x, const ConvParam& c, bool& risky) {
  float u0 = fmaf(x.x, c.rinv, c.mz), u1 = fmaf(x.y, c.rinv, c.mz), u2 = fmaf(x.z, c.rinv, c.mz), u3 = fmaf(x.w, c.rinv, c.mz);
  const float e0 = fmaf(x.x, c.rinv, -__fsub_rn(u0, c.mz)), e1 = fmaf(x.y, c.rinv, -__fsub_rn(u1, c.mz));
  const float e2 = fmaf(x.z, c.rinv, -__fsub_rn(u2, c.mz)), e3 = fmaf(x.w, c.rinv, -__fsub_rn(u3, c.mz));
  const float worst = fmaxf(fmaxf(fabsf(e0), fabsf(e1)), fmaxf(fabsf(e2), fabsf(e3)));
  risky = !(worst <= 0.4999f);  // NaN residuals (non-finite A) are dropped by fmaxf: such inputs are outside the contract
  u0 = fminf(fmaxf(u0, c.lo), c.hi); u1 = fminf(fmaxf(u1, c.lo), c.hi);
  u2 = fminf(fmaxf(u2, c.lo), c.hi); u3 = fminf(fmaxf(u3, c.lo), c.hi);
  return pack4(__float_as_uint(u0), __float_as_uint(u1), __float_as_uint(u2), __float_as_uint(u3));
}

// conversion constants of one quantizer (effective scale s, effective zero point z)
__device__ __forceinline__ ConvParam make_conv_param(float s, float z, float qmin, float qmax) {
  ConvParam cp;
  cp.s = s;
  cp.rinv = __frcp_rn(s);
  cp.zc = rintf(z) - qmin;
  cp.span = qmax - qmin;
  cp.mz = kMagic + cp.zc;
  cp.lo = kMagic;
  cp.hi = kMagic + cp.span;
  return cp;
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// streaming 128-bit load that does not allocate in L1 (data is touched once)
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// observer.py:100-119 for one (min,max) pair.  Returns scale; zp by reference (fp32, integer valued).
__device__ __forceinline__ float calc_qparams(float mn, float mx, int qmin, int qmax, int symmetric, float& zp) {
  float lo = fminf(mn, 0.f), hi = fmaxf(mx, 0.f);
  float s;
  if (symmetric) {
    hi = fmaxf(-lo, hi);
    s = fmaxf(__fdiv_rn(hi, (float)(qmax - qmin) / 2.f), 1e-8f);
    zp = 0.f;
  } else {
    s = fmaxf(__fdiv_rn(__fsub_rn(hi, lo), (float)(qmax - qmin)), 1e-8f);
    float z = __fsub_rn((float)qmin, rintf(__fdiv_rn(lo, s)));
    zp = fminf(fmaxf(z, (float)qmin), (float)qmax);
  }
  return s;
}

// running-statistics epilogue executed by ONE thread (see osq_stat_epilogue_t in osq.h)
__device__ __forceinline__ void stat_epilogue(const osq_stat_epilogue_t& e, float cur_min, float cur_max) {
  if (e.mode == 0) return;
  float mn = *e.state_min, mx = *e.state_max;
  if (e.mode == 1) {
    if (isinf(mx)) {  // observer.py:194-196 (`first batch` branch: max_val still +-inf)
      mn = cur_min;
      mx = cur_max;
    } else {
      mn = __fadd_rn(__fmul_rn(mn, (float)e.cnt), cur_min);
      mx = __fadd_rn(__fmul_rn(mx, (float)e.cnt), cur_max);
    }
    mn = __fdiv_rn(mn, (float)(e.cnt + 1));
    mx = __fdiv_rn(mx, (float)(e.cnt + 1));
  } else {
    mn = fminf(mn, cur_min);
    mx = fmaxf(mx, cur_max);
  }
  *e.state_min = mn;
  *e.state_max = mx;
  if (e.scale_out != nullptr) {
    float zp;
    float s = calc_qparams(mn, mx, e.qmin, e.qmax, e.symmetric, zp);
    e.scale_out[0] = s;
    if (e.zp_out != nullptr) {
      if (e.zp_out_is_int32) *(int32_t*)e.zp_out = (int32_t)zp;
      else *(float*)e.zp_out = zp;
    }
  }
}

}  // namespace osq
