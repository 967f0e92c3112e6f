// fq_elementwise.cu -- K1 / K2: per-tensor and per-channel fake-quantize (one read, one write).
//
// HBM-bound: 8 algorithmic bytes per element (4 read + 4 written).  Grid = multiple of the SM
// count, 128-bit streaming loads/stores, device-resident quantisation parameters (no .item()).
#include <stdlib.h>

#include "common.cuh"

namespace osq {

constexpr int kFqThreads = 256;
constexpr int kFqUnroll = 4;  // float4s in flight per thread

// Programmatic dependent launch for the activation fake-quant kernels (they sit between fused Linears on the module path):
// launched with the programmatic-serialization attribute, each starts its CTAs while the previous kernel drains and parks them
// at griddepcontrol.wait; it releases its own dependents at once -- a dependent that reads this kernel's output orders itself
// with its own griddepcontrol.wait (the fused Linear does, for everything but its static weights).  Measured on the e2e module
// stack (96 kernels per step): 3.16 ms with it, 3.12 ms without -- CTAs parked on the SMs cost more than the launch gaps they
// save -- so it is OFF by default (OSQ_FQ_PDL=1 enables it); without the attribute the two instructions are no-ops.
__device__ __forceinline__ void fq_pdl_prologue() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <class... KArgs, class... Args>
static cudaError_t fq_launch(void (*kernel)(KArgs...), int grid, cudaStream_t st, Args... args) {
  static int env_pdl = -1;
  if (env_pdl < 0) { const char* e = getenv("OSQ_FQ_PDL"); env_pdl = e ? atoi(e) : 0; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3(kFqThreads, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = env_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// kCodes: 0 = no side output, 1 = int16 bins q (parity tests), 2 = uint8 bins q - qmin (the operand format of the
// fused Linear kernel: a downstream QLinear can consume them instead of re-reading and re-quantising the fp32 tensor)
template <int kCodes>
__device__ __forceinline__ void put_code(void* codes, int64_t i, float q, float qmin) {
  if (kCodes == 1) static_cast<int16_t*>(codes)[i] = (int16_t)rintf(q);
  if (kCodes == 2) static_cast<uint8_t*>(codes)[i] = (uint8_t)__float2int_rn(q - qmin);   // (nearest: LSQ+ zero points may sit 1 ulp off an integer)
}

// GELU(x) = (x * 0.5) * (1 + erf(x / sqrt(2))) with the operation order of ATen's CUDA kernel (approximate = 'none'):
// bit-identical to torch.nn.functional.gelu on the same device
__device__ __forceinline__ float gelu_erf_k1(float x) {
  return __fmul_rn(__fmul_rn(x, 0.5f), __fadd_rn(1.0f, erff(__fmul_rn(x, 0.70710678118654752440f))));
}
template <int kAct>
__device__ __forceinline__ float act_in(float x) { return kAct == 1 ? gelu_erf_k1(x) : x; }

// kAct: 0 = none, 1 = GELU applied to x before the fake-quant (quant_bert.py:278-280: intermediate_act_fn followed by its
// quantizer as ONE pass over the tensor)
template <int kCodes, int kAct = 0>
__global__ void __launch_bounds__(kFqThreads)
fq_per_tensor_kernel(const float* __restrict__ x, float* __restrict__ y, void* __restrict__ codes,
                     int64_t n, const float* __restrict__ scale, const void* __restrict__ zp,
                     int zp_is_int32, float g, float qmin, float qmax) {
  fq_pdl_prologue();
  const QParam p = load_qparam(scale, zp, zp_is_int32, g, qmin, qmax,
                               blockIdx.x == 0 && threadIdx.x == 0);
  const float s = p.s, z = p.z;
  const float rinv = __frcp_rn(s);
  // head: elements before the first 16-byte boundary, tail: after the last full float4
  int64_t head = (int64_t)((16 - ((uintptr_t)x & 15)) & 15) >> 2;
  if (head > n) head = n;
  const bool vec_ok = (((uintptr_t)x & 3) == 0) && ((((uintptr_t)x) & 15) == (((uintptr_t)y) & 15));
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  if (!vec_ok) {
    for (int64_t i = tid; i < n; i += nthreads) {
      float q;
      y[i] = fq_elem(act_in<kAct>(x[i]), s, z, qmin, qmax, q);
      put_code<kCodes>(codes, i, q, qmin);
    }
    return;
  }
  const int64_t nvec = (n - head) >> 2;
  const float4* xv = reinterpret_cast<const float4*>(x + head);
  float4* yv = reinterpret_cast<float4*>(y + head);
  const bool word_codes = kCodes == 2 && (((uintptr_t)codes + (uintptr_t)head) & 3) == 0;  // four bins per 32-bit store
  for (int64_t base = tid; base < nvec; base += nthreads * kFqUnroll) {
    float4 v[kFqUnroll];
#pragma unroll
    for (int u = 0; u < kFqUnroll; ++u) {
      int64_t i = base + (int64_t)u * nthreads;
      if (i < nvec) v[u] = ldg_stream(xv + i);
    }
#pragma unroll
    for (int u = 0; u < kFqUnroll; ++u) {
      int64_t i = base + (int64_t)u * nthreads;
      if (i < nvec) {
        float4 o;
        float q0, q1, q2, q3;
        bool k0, k1, k2, k3;
        const float a0 = act_in<kAct>(v[u].x), a1 = act_in<kAct>(v[u].y), a2 = act_in<kAct>(v[u].z), a3 = act_in<kAct>(v[u].w);
        o.x = fq_elem_fast(a0, s, rinv, z, qmin, qmax, q0, k0);
        o.y = fq_elem_fast(a1, s, rinv, z, qmin, qmax, q1, k1);
        o.z = fq_elem_fast(a2, s, rinv, z, qmin, qmax, q2, k2);
        o.w = fq_elem_fast(a3, s, rinv, z, qmin, qmax, q3, k3);
        if (k0 | k1 | k2 | k3) {  // rare: the whole group through the exact division
          o.x = fq_elem(a0, s, z, qmin, qmax, q0);
          o.y = fq_elem(a1, s, z, qmin, qmax, q1);
          o.z = fq_elem(a2, s, z, qmin, qmax, q2);
          o.w = fq_elem(a3, s, z, qmin, qmax, q3);
        }
        __stcs(yv + i, o);
        if (kCodes != 0) {
          const int64_t e = head + (i << 2);
          if (word_codes) {
            const uint32_t w = (uint32_t)__float2int_rn(q0 - qmin) | ((uint32_t)__float2int_rn(q1 - qmin) << 8) | ((uint32_t)__float2int_rn(q2 - qmin) << 16) |
                               ((uint32_t)__float2int_rn(q3 - qmin) << 24);
            *reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(codes) + e) = w;
          } else {
            put_code<kCodes>(codes, e, q0, qmin); put_code<kCodes>(codes, e + 1, q1, qmin);
            put_code<kCodes>(codes, e + 2, q2, qmin); put_code<kCodes>(codes, e + 3, q3, qmin);
          }
        }
      }
    }
  }
  // scalar head + tail
  const int64_t tail_start = head + (nvec << 2);
  for (int64_t i = tid; i < head + (n - tail_start); i += nthreads) {
    int64_t j = i < head ? i : tail_start + (i - head);
    float q;
    y[j] = fq_elem(act_in<kAct>(x[j]), s, z, qmin, qmax, q);
    put_code<kCodes>(codes, j, q, qmin);
  }
}

// one CTA per (row, column-slab): scale/zp are per row (ch_axis = 0)
template <bool kCodes>
__global__ void __launch_bounds__(kFqThreads)
fq_per_channel_kernel(const float* __restrict__ x, float* __restrict__ y, int16_t* __restrict__ codes,
                      int64_t rows, int64_t cols, const float* __restrict__ scale,
                      const int32_t* __restrict__ zp, float qmin, float qmax) {
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const float s = scale[r];
    const float z = (float)zp[r];
    const float* xr = x + r * cols;
    float* yr = y + r * cols;
    const bool vec_ok = ((cols & 3) == 0) && (((uintptr_t)xr & 15) == 0) && (((uintptr_t)yr & 15) == 0);
    if (vec_ok) {
      const float4* xv = reinterpret_cast<const float4*>(xr);
      float4* yv = reinterpret_cast<float4*>(yr);
      for (int64_t i = threadIdx.x; i < (cols >> 2); i += blockDim.x) {
        float4 v = ldg_stream(xv + i), o;
        float q0, q1, q2, q3;
        o.x = fq_elem(v.x, s, z, qmin, qmax, q0);
        o.y = fq_elem(v.y, s, z, qmin, qmax, q1);
        o.z = fq_elem(v.z, s, z, qmin, qmax, q2);
        o.w = fq_elem(v.w, s, z, qmin, qmax, q3);
        yv[i] = o;
        if (kCodes) {
          int16_t* c = codes + r * cols + (i << 2);
          c[0] = (int16_t)rintf(q0); c[1] = (int16_t)rintf(q1);
          c[2] = (int16_t)rintf(q2); c[3] = (int16_t)rintf(q3);
        }
      }
    } else {
      for (int64_t i = threadIdx.x; i < cols; i += blockDim.x) {
        float q;
        yr[i] = fq_elem(xr[i], s, z, qmin, qmax, q);
        if (kCodes) codes[r * cols + i] = (int16_t)rintf(q);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// LSQ+ backward (gradients of util_quant.py:48-55 as autograd derives them)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFqThreads)
lsqplus_backward_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, int64_t n,
                        const float* __restrict__ scale, const float* __restrict__ zp, float g, float qmin, float qmax,
                        double* __restrict__ grad_acc) {
  const QParam p = load_qparam(scale, zp, 0, g, qmin, qmax, false);
  const float s = p.s, z = p.z;
  double ds = 0.0, dz = 0.0;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  // one element of the backward pass (the gradients autograd derives from util_quant.py:48-55)
  auto elem = [&](float xv, float gy) -> float {
    const float t = __fdiv_rn(xv, s);
    float r = rintf(t);
    r = __fadd_rn(__fsub_rn(r, t), t);
    const float v = __fadd_rn(r, z);
    const bool inside = (v >= qmin) && (v <= qmax);
    const float gq = __fmul_rn(gy, s);  // d/d x_quant
    if (inside) {
      ds += (double)gy * ((double)__fsub_rn(v, z) - (double)t);
      return __fdiv_rn(gq, s);
    }
    const float q = fminf(fmaxf(v, qmin), qmax);
    ds += (double)gy * (double)__fsub_rn(q, z);
    dz -= (double)gq;
    return 0.f;
  };
  // 128-bit streaming loads / stores when the three arrays are 16-byte aligned (12 algorithmic bytes per element)
  const bool vec_ok = ((((uintptr_t)x) | ((uintptr_t)dy) | ((uintptr_t)dx)) & 15) == 0;
  const int64_t nvec = vec_ok ? (n >> 2) : 0;
  const float4* xv4 = reinterpret_cast<const float4*>(x);
  const float4* gy4 = reinterpret_cast<const float4*>(dy);
  float4* dx4 = reinterpret_cast<float4*>(dx);
  for (int64_t i = tid; i < nvec; i += 2 * nthreads) {
    const bool two = i + nthreads < nvec;
    const float4 a0 = ldg_stream(xv4 + i), g0 = ldg_stream(gy4 + i);
    const float4 a1 = two ? ldg_stream(xv4 + i + nthreads) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 g1 = two ? ldg_stream(gy4 + i + nthreads) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 o;
    o.x = elem(a0.x, g0.x); o.y = elem(a0.y, g0.y); o.z = elem(a0.z, g0.z); o.w = elem(a0.w, g0.w);
    __stcs(dx4 + i, o);
    if (two) {
      o.x = elem(a1.x, g1.x); o.y = elem(a1.y, g1.y); o.z = elem(a1.z, g1.z); o.w = elem(a1.w, g1.w);
      __stcs(dx4 + i + nthreads, o);
    }
  }
  for (int64_t i = (nvec << 2) + tid; i < n; i += nthreads) dx[i] = elem(x[i], dy[i]);
  __shared__ double sds[kFqThreads / 32], sdz[kFqThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ds += __shfl_xor_sync(0xffffffffu, ds, o);
    dz += __shfl_xor_sync(0xffffffffu, dz, o);
  }
  if ((threadIdx.x & 31) == 0) { sds[threadIdx.x >> 5] = ds; sdz[threadIdx.x >> 5] = dz; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int i = 0; i < kFqThreads / 32; ++i) { a += sds[i]; b += sdz[i]; }
    atomicAdd(grad_acc + 0, a * (double)g);
    atomicAdd(grad_acc + 1, b * (double)g);
  }
}

// K1d  bins only: the per-tensor fake-quantize of K1 / K1b without its fp32 output -- 5 bytes per element instead of 9.  For a
// quantizer whose output is consumed by fused QLinears alone (the dequantised tensor is then never read); `eff` receives the
// effective (scale, zero point) the launch used, so that osq_dequant_bins_f32 can rebuild the fp32 tensor later.
template <int kAct>
__global__ void __launch_bounds__(kFqThreads)
fq_bins_only_kernel(const float4* __restrict__ x, uint32_t* __restrict__ bins, int64_t nvec, const float* __restrict__ scale,
                    const void* __restrict__ zp, int zp_is_int32, float g, float qmin, float qmax, float* __restrict__ eff) {
  fq_pdl_prologue();
  const QParam p = load_qparam(scale, zp, zp_is_int32, g, qmin, qmax, blockIdx.x == 0 && threadIdx.x == 0);
  const float s = p.s, z = p.z, rinv = __frcp_rn(s);
  if (blockIdx.x == 0 && threadIdx.x == 0 && eff != nullptr) { eff[0] = s; eff[1] = z; }
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t base = tid; base < nvec; base += nthreads * kFqUnroll) {
    float4 v[kFqUnroll];
#pragma unroll
    for (int u = 0; u < kFqUnroll; ++u) {
      const int64_t i = base + (int64_t)u * nthreads;
      if (i < nvec) v[u] = ldg_stream(x + i);
    }
#pragma unroll
    for (int u = 0; u < kFqUnroll; ++u) {
      const int64_t i = base + (int64_t)u * nthreads;
      if (i < nvec) {
        float q0, q1, q2, q3;
        bool k0, k1, k2, k3;
        const float a0 = act_in<kAct>(v[u].x), a1 = act_in<kAct>(v[u].y), a2 = act_in<kAct>(v[u].z), a3 = act_in<kAct>(v[u].w);
        fq_elem_fast(a0, s, rinv, z, qmin, qmax, q0, k0);
        fq_elem_fast(a1, s, rinv, z, qmin, qmax, q1, k1);
        fq_elem_fast(a2, s, rinv, z, qmin, qmax, q2, k2);
        fq_elem_fast(a3, s, rinv, z, qmin, qmax, q3, k3);
        if (k0 | k1 | k2 | k3) {
          fq_elem(a0, s, z, qmin, qmax, q0); fq_elem(a1, s, z, qmin, qmax, q1);
          fq_elem(a2, s, z, qmin, qmax, q2); fq_elem(a3, s, z, qmin, qmax, q3);
        }
        bins[i] = (uint32_t)__float2int_rn(q0 - qmin) | ((uint32_t)__float2int_rn(q1 - qmin) << 8) | ((uint32_t)__float2int_rn(q2 - qmin) << 16) |
                  ((uint32_t)__float2int_rn(q3 - qmin) << 24);
      }
    }
  }
}

// (bins, eff) -> the fp32 tensor K1 would have written: y = (q - z) * s with q = bin + qmin.  Bit-identical to K1 whenever the
// effective zero point is integer valued (always for FixedFakeQuantize; LSQ+'s grad_scale leaves rint(z) untouched except for
// a rare 1-ulp drift, in which case q is rebuilt as clamp((q_int - rint(z)) + z) exactly as util_quant.py:13 forms it).
__global__ void __launch_bounds__(kFqThreads)
dequant_bins_kernel(const uint32_t* __restrict__ bins, float4* __restrict__ y, int64_t nvec, const float* __restrict__ eff, float qmin,
                    float qmax) {
  const float s = eff[0], z = eff[1], t = rintf(eff[1]);
  const bool drift = z != t;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  auto one = [&](uint32_t b) -> float {
    float q = __fadd_rn((float)b, qmin);
    if (drift && q > qmin && q < qmax) q = fminf(fmaxf(__fadd_rn(__fsub_rn(q, t), z), qmin), qmax);
    return __fmul_rn(__fsub_rn(q, z), s);
  };
  for (int64_t i = tid; i < nvec; i += nthreads) {
    const uint32_t w = __ldg(bins + i);
    __stcs(y + i, make_float4(one(w & 0xFFu), one((w >> 8) & 0xFFu), one((w >> 16) & 0xFFu), one(w >> 24)));
  }
}

__global__ void calc_qparams_kernel(const float* __restrict__ mn, const float* __restrict__ mx, int64_t n, int qmin,
                                    int qmax, int symmetric, float* __restrict__ scale, float* __restrict__ zp_f32,
                                    int32_t* __restrict__ zp_i32) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float zp;
  scale[i] = calc_qparams(mn[i], mx[i], qmin, qmax, symmetric, zp);
  if (zp_f32) zp_f32[i] = zp;
  if (zp_i32) zp_i32[i] = (int32_t)zp;
}

}  // namespace osq

extern "C" {

int osq_calc_qparams_f32(const float* min_val, const float* max_val, int64_t n, int qmin, int qmax, int symmetric,
                         float* scale, float* zp_f32, int32_t* zp_i32, void* stream) {
  using namespace osq;
  OSQ_CHECK_ARG(min_val && max_val && scale && n >= 0, "osq_calc_qparams_f32: bad argument");
  if (n == 0) return OSQ_OK;
  calc_qparams_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(min_val, max_val, n, qmin, qmax, symmetric,
                                                                                   scale, zp_f32, zp_i32);
  OSQ_LAUNCH_CHECK();
  return OSQ_OK;
}

int osq_lsqplus_backward_f32(const float* x, const float* dy, float* dx, int64_t n, const float* scale,
                             const float* zero_point, float lsq_grad_factor, int qmin, int qmax, double* grad_acc,
                             void* stream) {
  using namespace osq;
  OSQ_CHECK_ARG(x && dy && dx && scale && zero_point && grad_acc && n >= 0, "osq_lsqplus_backward_f32: bad argument");
  OSQ_CHECK_ARG(lsq_grad_factor > 0.f, "osq_lsqplus_backward_f32: grad factor must be > 0");
  if (n == 0) return OSQ_OK;
  int sms = sm_count();
  if (sms <= 0) { set_error("no CUDA device"); return OSQ_ECUDA; }
  int64_t want = (n + kFqThreads * 8 - 1) / (kFqThreads * 8);
  int grid = (int)(want < (int64_t)sms * 8 ? (want < 1 ? 1 : want) : (int64_t)sms * 8);
  lsqplus_backward_kernel<<<grid, kFqThreads, 0, (cudaStream_t)stream>>>(x, dy, dx, n, scale, zero_point, lsq_grad_factor,
                                                                        (float)qmin, (float)qmax, grad_acc);
  OSQ_LAUNCH_CHECK();
  return OSQ_OK;
}

static int launch_fq_per_tensor(const char* who, const float* x, float* y, void* codes, int code_kind, int64_t n, const float* scale,
                               const void* zero_point, int zp_is_int32, float lsq_grad_factor, int qmin, int qmax, void* stream) {
  using namespace osq;
  OSQ_CHECK_ARG(n >= 0, "%s: n < 0", who);
  if (n == 0) return OSQ_OK;
  OSQ_CHECK_ARG(x && y && scale && zero_point, "%s: null pointer", who);
  OSQ_CHECK_ARG(qmin < qmax, "%s: qmin >= qmax", who);
  OSQ_CHECK_ARG(!(lsq_grad_factor > 0.f && zp_is_int32), "%s: LSQ+ needs a float zero_point", who);
  OSQ_CHECK_ARG(code_kind != 2 || qmax - qmin <= 255, "%s: uint8 bins need at most 8 bits", who);
  int sms = sm_count();
  if (sms <= 0) { set_error("no CUDA device"); return OSQ_ECUDA; }
  int64_t per_block = (int64_t)kFqThreads * 4 * kFqUnroll;
  int64_t want = (n + per_block - 1) / per_block;
  int grid = (int)(want < (int64_t)sms * 8 ? (want < 1 ? 1 : want) : (int64_t)sms * 8);
  cudaStream_t st = (cudaStream_t)stream;
  const float fmin_ = (float)qmin, fmax_ = (float)qmax;
  if (codes == nullptr)
    OSQ_CUDA(fq_launch(fq_per_tensor_kernel<0>, grid, st, x, y, (void*)nullptr, n, scale, zero_point, zp_is_int32, lsq_grad_factor, fmin_, fmax_));
  else if (code_kind == 1)
    OSQ_CUDA(fq_launch(fq_per_tensor_kernel<1>, grid, st, x, y, codes, n, scale, zero_point, zp_is_int32, lsq_grad_factor, fmin_, fmax_));
  else
    OSQ_CUDA(fq_launch(fq_per_tensor_kernel<2>, grid, st, x, y, codes, n, scale, zero_point, zp_is_int32, lsq_grad_factor, fmin_, fmax_));
  OSQ_LAUNCH_CHECK();
  return OSQ_OK;
}

int osq_fq_per_tensor_f32(const float* x, float* y, int16_t* codes, int64_t n, const float* scale,
                          const void* zero_point, int zp_is_int32, float lsq_grad_factor, int qmin,
                          int qmax, void* stream) {
  return launch_fq_per_tensor("osq_fq_per_tensor_f32", x, y, codes, 1, n, scale, zero_point, zp_is_int32, lsq_grad_factor, qmin, qmax, stream);
}

int osq_fq_per_tensor_bins_f32(const float* x, float* y, uint8_t* bins, int64_t n, const float* scale,
                               const void* zero_point, int zp_is_int32, float lsq_grad_factor, int qmin,
                               int qmax, void* stream) {
  using namespace osq;
  OSQ_CHECK_ARG(bins != nullptr, "osq_fq_per_tensor_bins_f32: null bins");
  return launch_fq_per_tensor("osq_fq_per_tensor_bins_f32", x, y, bins, 2, n, scale, zero_point, zp_is_int32, lsq_grad_factor, qmin, qmax, stream);
}

int osq_act_fq_per_tensor_bins_f32(const float* x, float* y, uint8_t* bins, int64_t n, int act, const float* scale,
                                   const void* zero_point, int zp_is_int32, float lsq_grad_factor, int qmin,
                                   int qmax, void* stream) {
  using namespace osq;
  if (act == 0) return launch_fq_per_tensor("osq_act_fq_per_tensor_bins_f32", x, y, bins, 2, n, scale, zero_point, zp_is_int32,
                                            lsq_grad_factor, qmin, qmax, stream);
  OSQ_CHECK_ARG(act == 1, "osq_act_fq_per_tensor_bins_f32: act must be 0 (none) or 1 (GELU, erf form)");
  OSQ_CHECK_ARG(n >= 0, "osq_act_fq_per_tensor_bins_f32: n < 0");
  if (n == 0) return OSQ_OK;
  OSQ_CHECK_ARG(x && y && scale && zero_point, "osq_act_fq_per_tensor_bins_f32: null pointer");
  OSQ_CHECK_ARG(qmin < qmax, "osq_act_fq_per_tensor_bins_f32: qmin >= qmax");
  OSQ_CHECK_ARG(!(lsq_grad_factor > 0.f && zp_is_int32), "osq_act_fq_per_tensor_bins_f32: LSQ+ needs a float zero_point");
  OSQ_CHECK_ARG(bins == nullptr || qmax - qmin <= 255, "osq_act_fq_per_tensor_bins_f32: uint8 bins need at most 8 bits");
  int sms = sm_count();
  if (sms <= 0) { set_error("no CUDA device"); return OSQ_ECUDA; }
  int64_t per_block = (int64_t)kFqThreads * 4 * kFqUnroll;
  int64_t want = (n + per_block - 1) / per_block;
  int grid = (int)(want < (int64_t)sms * 8 ? (want < 1 ? 1 : want) : (int64_t)sms * 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (bins != nullptr)
    OSQ_CUDA(fq_launch(fq_per_tensor_kernel<2, 1>, grid, st, x, y, (void*)bins, n, scale, zero_point, zp_is_int32, lsq_grad_factor, (float)qmin, (float)qmax));
  else
    OSQ_CUDA(fq_launch(fq_per_tensor_kernel<0, 1>, grid, st, x, y, (void*)nullptr, n, scale, zero_point, zp_is_int32, lsq_grad_factor, (float)qmin, (float)qmax));
  OSQ_LAUNCH_CHECK();
  return OSQ_OK;
}

int osq_fq_per_tensor_bins_only_f32(const float* x, uint8_t* bins, int64_t n, int act, const float* scale, const void* zero_point,
                                    int zp_is_int32, float lsq_grad_factor, int qmin, int qmax, float* eff, void* stream) {
  using namespace osq;
  OSQ_CHECK_ARG(n >= 0, "osq_fq_per_tensor_bins_only_f32: n < 0");
  if (n == 0) return OSQ_OK;
  OSQ_CHECK_ARG(x && bins && scale && zero_point, "osq_fq_per_tensor_bins_only_f32: null pointer");
  OSQ_CHECK_ARG(act == 0 || act == 1, "osq_fq_per_tensor_bins_only_f32: act must be 0 (none) or 1 (GELU, erf form)");
  OSQ_CHECK_ARG(qmin < qmax && qmax - qmin <= 255, "osq_fq_per_tensor_bins_only_f32: uint8 bins need at most 8 bits");
  OSQ_CHECK_ARG(!(lsq_grad_factor > 0.f && zp_is_int32), "osq_fq_per_tensor_bins_only_f32: LSQ+ needs a float zero_point");
  OSQ_CHECK_ARG(n % 4 == 0 && (((uintptr_t)x) & 15) == 0 && (((uintptr_t)bins) & 3) == 0,
                "osq_fq_per_tensor_bins_only_f32: n must be a multiple of 4, x 16-byte and bins 4-byte aligned");
  int sms = sm_count();
  if (sms <= 0) { set_error("no CUDA device"); return OSQ_ECUDA; }
  const int64_t nvec = n >> 2;
  const int64_t per_block = (int64_t)kFqThreads * kFqUnroll;
  const int64_t want = (nvec + per_block - 1) / per_block;
  const int grid = (int)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);
  if (act == 1)
    OSQ_CUDA(fq_launch(fq_bins_only_kernel<1>, grid, (cudaStream_t)stream, reinterpret_cast<const float4*>(x), reinterpret_cast<uint32_t*>(bins), nvec,
                       scale, zero_point, zp_is_int32, lsq_grad_factor, (float)qmin, (float)qmax, eff));
  else
    OSQ_CUDA(fq_launch(fq_bins_only_kernel<0>, grid, (cudaStream_t)stream, reinterpret_cast<const float4*>(x), reinterpret_cast<uint32_t*>(bins), nvec,
                       scale, zero_point, zp_is_int32, lsq_grad_factor, (float)qmin, (float)qmax, eff));
  OSQ_LAUNCH_CHECK();
  return OSQ_OK;
}

int osq_dequant_bins_f32(const uint8_t* bins, const float* eff, int qmin, int qmax, float* y, int64_t n, void* stream) {
  using namespace osq;
  OSQ_CHECK_ARG(n >= 0, "osq_dequant_bins_f32: n < 0");
  if (n == 0) return OSQ_OK;
  OSQ_CHECK_ARG(bins && eff && y, "osq_dequant_bins_f32: null pointer");
  OSQ_CHECK_ARG(qmin < qmax && qmax - qmin <= 255, "osq_dequant_bins_f32: uint8 bins hold at most 8 bits");
  OSQ_CHECK_ARG(n % 4 == 0 && (((uintptr_t)y) & 15) == 0 && (((uintptr_t)bins) & 3) == 0,
                "osq_dequant_bins_f32: n must be a multiple of 4, y 16-byte and bins 4-byte aligned");
  int sms = sm_count();
  if (sms <= 0) { set_error("no CUDA device"); return OSQ_ECUDA; }
  const int64_t nvec = n >> 2;
  const int64_t want = (nvec + kFqThreads - 1) / kFqThreads;
  const int grid = (int)(want < (int64_t)sms * 16 ? want : (int64_t)sms * 16);
  dequant_bins_kernel<<<grid, kFqThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint32_t*>(bins), reinterpret_cast<float4*>(y), nvec, eff,
                                                                   (float)qmin, (float)qmax);
  OSQ_LAUNCH_CHECK();
  return OSQ_OK;
}

int osq_fq_per_channel_f32(const float* x, float* y, int16_t* codes, int64_t rows, int64_t cols,
                           const float* scale, const int32_t* zero_point, int qmin, int qmax,
                           void* stream) {
  using namespace osq;
  OSQ_CHECK_ARG(rows >= 0 && cols >= 0, "osq_fq_per_channel_f32: negative shape");
  if (rows == 0 || cols == 0) return OSQ_OK;
  OSQ_CHECK_ARG(x && y && scale && zero_point, "osq_fq_per_channel_f32: null pointer");
  OSQ_CHECK_ARG(qmin < qmax, "osq_fq_per_channel_f32: qmin >= qmax");
  int sms = sm_count();
  if (sms <= 0) { set_error("no CUDA device"); return OSQ_ECUDA; }
  int grid = (int)(rows < (int64_t)sms * 16 ? rows : (int64_t)sms * 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (codes)
    fq_per_channel_kernel<true><<<grid, kFqThreads, 0, st>>>(x, y, codes, rows, cols, scale, zero_point,
                                                             (float)qmin, (float)qmax);
  else
    fq_per_channel_kernel<false><<<grid, kFqThreads, 0, st>>>(x, y, nullptr, rows, cols, scale, zero_point,
                                                              (float)qmin, (float)qmax);
  OSQ_LAUNCH_CHECK();
  return OSQ_OK;
}

}  // extern "C"
