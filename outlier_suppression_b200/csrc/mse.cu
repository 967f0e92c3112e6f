// mse.cu -- K5: MSE-loss evaluation for the (Avg)MSEFast observers.
//
//   osq_mse_multi_f32      : sum of squared fake-quant error for up to 8 candidate qparams per pass
//                            over a (masked) activation -- 4 algorithmic bytes / element / pass.
//   osq_mse_brent_rows_f32 : MSEFastObserver(ch_axis=0): one CTA per weight row, the row resident in
//                            shared memory, SciPy's bounded Brent (fminbound) restated on-chip in fp64,
//                            so the ~15 loss evaluations per channel cost zero HBM traffic and zero
//                            host round trips (the reference does one .cpu() sync per evaluation).
#include "common.cuh"

namespace osq {

constexpr int kMseThreads = 512;
constexpr int kMseWarps = kMseThreads / 32;
constexpr int kMseMaxCand = 8;

struct Cands {
  float s[kMseMaxCand];
  float z[kMseMaxCand];
};

__device__ __forceinline__ float sq_err(float x, float s, float z, float qmin, float qmax) {
  float q;
  float y = fq_elem(x, s, z, qmin, qmax, q);
  float d = __fsub_rn(y, x);
  return __fmul_rn(d, d);
}

template <int C>
__global__ void __launch_bounds__(kMseThreads)
mse_multi_kernel(const float* __restrict__ x, osq_tokens_t tk, const int64_t* __restrict__ lens, int n_lens,
                 const float* __restrict__ cand_scale, const float* __restrict__ cand_zp, float qmin, float qmax,
                 double* __restrict__ loss_sum, int64_t* __restrict__ n_valid) {
  float s[C], z[C], acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) { s[c] = cand_scale[c]; z[c] = cand_zp[c]; acc[c] = 0.f; }
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t)blockIdx.x * kMseWarps + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * kMseWarps;
  const int64_t n_seg = tk.B * tk.S * tk.F1;
  if (blockIdx.x == 0 && threadIdx.x == 0 && n_valid != nullptr) {
    int64_t t = 0;
    if (lens == nullptr) t = tk.B * tk.S;
    else
      for (int64_t b = 0; b < tk.B && b < n_lens; ++b) {
        int64_t l = lens[b];
        t += l < 0 ? 0 : (l > tk.S ? tk.S : l);
      }
    *n_valid = t * tk.F1 * tk.F2;
  }
  double dacc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) dacc[c] = 0.0;
  for (int64_t seg = warp_global; seg < n_seg; seg += n_warps) {
    const int64_t f1 = seg % tk.F1;
    const int64_t bs = seg / tk.F1;
    const int64_t sidx = bs % tk.S, b = bs / tk.S;
    if (lens != nullptr && (b >= n_lens || sidx >= lens[b])) continue;
    const float* p = x + b * tk.sb + sidx * tk.ss + f1 * tk.sf1;
    if (tk.sf2 == 1 && (((uintptr_t)p) & 15) == 0 && (tk.F2 & 3) == 0) {
      const float4* v = reinterpret_cast<const float4*>(p);
      for (int64_t i = lane; i < (tk.F2 >> 2); i += 32) {
        float4 a = ldg_stream(v + i);
#pragma unroll
        for (int c = 0; c < C; ++c)
          acc[c] += sq_err(a.x, s[c], z[c], qmin, qmax) + sq_err(a.y, s[c], z[c], qmin, qmax) +
                    sq_err(a.z, s[c], z[c], qmin, qmax) + sq_err(a.w, s[c], z[c], qmin, qmax);
      }
    } else {
      for (int64_t i = lane; i < tk.F2; i += 32) {
        float a = __ldg(p + i * tk.sf2);
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] += sq_err(a, s[c], z[c], qmin, qmax);
      }
    }
    // fold the short fp32 partial into fp64 once per segment (keeps fp32 accumulation runs short)
#pragma unroll
    for (int c = 0; c < C; ++c) { dacc[c] += (double)acc[c]; acc[c] = 0.f; }
  }
  __shared__ double sacc[kMseWarps][kMseMaxCand];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    double v = dacc[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sacc[threadIdx.x >> 5][c] = v;
  }
  __syncthreads();
  if (threadIdx.x < C) {
    double t = 0;
    for (int w = 0; w < kMseWarps; ++w) t += sacc[w][threadIdx.x];
    atomicAdd(loss_sum + threadIdx.x, t);
  }
}

// ---------------------------------------------------------------------------------------------
// bounded Brent (Forsythe/Malcolm/Moler fminbound, as in scipy.optimize._minimize_scalar_bounded:
// xatol = 1e-5, maxiter = 500, golden mean 0.5*(3-sqrt(5)), sqrt_eps = sqrt(2.2e-16)); fp64 state.
// Every thread of the CTA runs the same state machine on the same (broadcast) loss values.
// ---------------------------------------------------------------------------------------------
struct RowLoss {
  const float* row;  // shared or global
  int64_t cols;
  float qmin, qmax;
  int symmetric, one_side;
  double* red;  // shared scratch [warps]
  float* bcast;

  __device__ float operator()(double r) const {
    // calculate_qparams on (new_min, new_max) = (-r or 0, r or 0) in fp64, observer.py:453-456,100-119
    const double span = (double)(qmax - qmin);
    double scale64;
    float zp = 0.f;
    if (symmetric) {
      scale64 = r / (span / 2.0);
    } else {
      scale64 = r / span;  // one-sided: (max_pos - min_neg) == r
    }
    const double eps = (double)1e-8f;
    if (!(scale64 > eps)) scale64 = eps;
    if (!symmetric) {
      const double mn = (one_side == 1) ? 0.0 : -r;
      double z = (double)qmin - rint(mn / scale64);
      z = z < (double)qmin ? (double)qmin : (z > (double)qmax ? (double)qmax : z);
      zp = (float)z;
    }
    const float s = (float)scale64;
    float acc = 0.f;
    for (int64_t i = threadIdx.x; i < cols; i += blockDim.x) acc += sq_err(row[i], s, zp, qmin, qmax);
    double v = (double)acc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
      *bcast = (float)(t / (double)cols);  // the reference's loss is an fp32 mean
    }
    __syncthreads();
    return *bcast;
  }
};

template <class F>
__device__ double fminbound(const F& func, double x1, double x2, int* nfev) {
  const double sqrt_eps = sqrt(2.2e-16);
  const double golden_mean = 0.5 * (3.0 - sqrt(5.0));
  const double xatol = 1e-5;
  const int maxfun = 500;
  double a = x1, b = x2;
  double fulc = a + golden_mean * (b - a);
  double nfc = fulc, xf = fulc;
  double rat = 0.0, e = 0.0;
  double x = xf;
  double fx = (double)func(x);
  int num = 1;
  double ffulc = fx, fnfc = fx;
  double xm = 0.5 * (a + b);
  double tol1 = sqrt_eps * fabs(xf) + xatol / 3.0;
  double tol2 = 2.0 * tol1;
  while (fabs(xf - xm) > (tol2 - 0.5 * (b - a))) {
    bool golden = true;
    if (fabs(e) > tol1) {  // try a parabolic step
      golden = false;
      double r = (xf - nfc) * (fx - ffulc);
      double q = (xf - fulc) * (fx - fnfc);
      double p = (xf - fulc) * q - (xf - nfc) * r;
      q = 2.0 * (q - r);
      if (q > 0.0) p = -p;
      q = fabs(q);
      r = e;
      e = rat;
      if ((fabs(p) < fabs(0.5 * q * r)) && (p > q * (a - xf)) && (p < q * (b - xf))) {
        rat = (p + 0.0) / q;
        x = xf + rat;
        if (((x - a) < tol2) || ((b - x) < tol2)) {
          double d = xm - xf;
          double si = (d > 0.0 ? 1.0 : (d < 0.0 ? -1.0 : 0.0)) + (d == 0.0 ? 1.0 : 0.0);
          rat = tol1 * si;
        }
      } else {
        golden = true;
      }
    }
    if (golden) {
      e = (xf >= xm) ? (a - xf) : (b - xf);
      rat = golden_mean * e;
    }
    double si = (rat > 0.0 ? 1.0 : (rat < 0.0 ? -1.0 : 0.0)) + (rat == 0.0 ? 1.0 : 0.0);
    x = xf + si * fmax(fabs(rat), tol1);
    double fu = (double)func(x);
    ++num;
    if (fu <= fx) {
      if (x >= xf) a = xf; else b = xf;
      fulc = nfc; ffulc = fnfc;
      nfc = xf; fnfc = fx;
      xf = x; fx = fu;
    } else {
      if (x < xf) a = x; else b = x;
      if ((fu <= fnfc) || (nfc == xf)) {
        fulc = nfc; ffulc = fnfc;
        nfc = x; fnfc = fu;
      } else if ((fu <= ffulc) || (fulc == xf) || (fulc == nfc)) {
        fulc = x; ffulc = fu;
      }
    }
    xm = 0.5 * (a + b);
    tol1 = sqrt_eps * fabs(xf) + xatol / 3.0;
    tol2 = 2.0 * tol1;
    if (num >= maxfun) break;
  }
  if (nfev) *nfev = num;
  return xf;
}

constexpr int kBrentThreads = 256;

__global__ void __launch_bounds__(kBrentThreads)
mse_brent_rows_kernel(const float* __restrict__ w, int64_t rows, int64_t cols, float qmin, float qmax, int symmetric,
                      int one_side, int row_in_smem, float* __restrict__ out_min, float* __restrict__ out_max,
                      int32_t* __restrict__ evals) {
  extern __shared__ __align__(16) float srow[];
  __shared__ double red[kBrentThreads / 32];
  __shared__ float bcast;
  __shared__ float smn[kBrentThreads / 32], smx[kBrentThreads / 32];
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const float* grow = w + r * cols;
    float mn = INFINITY, mx = -INFINITY;
    for (int64_t i = threadIdx.x; i < cols; i += blockDim.x) {
      float v = grow[i];
      if (row_in_smem) srow[i] = v;
      mn = fminf(mn, v);
      mx = fmaxf(mx, v);
    }
    mn = warp_min(mn);
    mx = warp_max(mx);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
    __syncthreads();
    mn = INFINITY; mx = -INFINITY;
    for (int i = 0; i < kBrentThreads / 32; ++i) { mn = fminf(mn, smn[i]); mx = fmaxf(mx, smx[i]); }
    const double xrange = (double)fmaxf(fabsf(mn), mx);  // observer.py:484
    RowLoss f{row_in_smem ? srow : grow, cols, qmin, qmax, symmetric, one_side, red, &bcast};
    int nfev = 0;
    const double lo = fmin(0.1, 0.01 * xrange);
    const double best = fminbound(f, lo, xrange, &nfev);
    if (threadIdx.x == 0) {
      out_min[r] = (one_side == 1) ? 0.f : (float)(-best);
      out_max[r] = (one_side == 2) ? 0.f : (float)best;
      if (evals) evals[r] = nfev;
    }
    __syncthreads();
  }
}

}  // namespace osq

extern "C" {

int osq_mse_multi_f32(const float* x, const osq_tokens_t* tok, const int64_t* lens, int n_lens,
                      const float* cand_scale, const float* cand_zp, int n_cand, int qmin, int qmax,
                      double* loss_sum, int64_t* n_valid, void* stream) {
  using namespace osq;
  OSQ_CHECK_ARG(x && tok && cand_scale && cand_zp && loss_sum, "osq_mse_multi_f32: null pointer");
  OSQ_CHECK_ARG(n_cand >= 1, "osq_mse_multi_f32: n_cand < 1");
  OSQ_CHECK_ARG(qmin < qmax, "osq_mse_multi_f32: qmin >= qmax");
  int sms = sm_count();
  if (sms <= 0) { set_error("no CUDA device"); return OSQ_ECUDA; }
  cudaStream_t st = (cudaStream_t)stream;
  OSQ_CUDA(cudaMemsetAsync(loss_sum, 0, sizeof(double) * n_cand, st));
  int64_t n_seg = tok->B * tok->S * tok->F1;
  int64_t g = (n_seg + kMseWarps - 1) / kMseWarps;
  if (g > (int64_t)sms * 4) g = (int64_t)sms * 4;
  if (g < 1) g = 1;
  for (int c0 = 0; c0 < n_cand;) {
    const int rem = n_cand - c0;
    const int width = rem >= 8 ? 8 : (rem >= 4 ? 4 : (rem >= 2 ? 2 : 1));
    int64_t* nv = (c0 == 0) ? n_valid : nullptr;
#define OSQ_MSE_LAUNCH(CC)                                                                                      \
  mse_multi_kernel<CC><<<(int)g, kMseThreads, 0, st>>>(x, *tok, lens, n_lens, cand_scale + c0, cand_zp + c0,   \
                                                       (float)qmin, (float)qmax, loss_sum + c0, nv)
    if (width == 8) OSQ_MSE_LAUNCH(8);
    else if (width == 4) OSQ_MSE_LAUNCH(4);
    else if (width == 2) OSQ_MSE_LAUNCH(2);
    else OSQ_MSE_LAUNCH(1);
#undef OSQ_MSE_LAUNCH
    OSQ_LAUNCH_CHECK();
    c0 += width;
  }
  return OSQ_OK;
}

int osq_mse_brent_rows_f32(const float* w, int64_t rows, int64_t cols, int qmin, int qmax, int one_side,
                           float* out_min, float* out_max, int32_t* evals, void* stream) {
  using namespace osq;
  OSQ_CHECK_ARG(w && out_min && out_max && rows > 0 && cols > 0, "osq_mse_brent_rows_f32: bad argument");
  OSQ_CHECK_ARG(one_side >= 0 && one_side <= 2, "osq_mse_brent_rows_f32: one_side must be 0 (no), 1 (pos) or 2 (neg)");
  const int symmetric = qmin < 0;
  OSQ_CHECK_ARG(symmetric || one_side != 0, "osq_mse_brent_rows_f32: asymmetric two-sided rows need the 2-D search (host driven)");
  int sms = sm_count();
  if (sms <= 0) { set_error("no CUDA device"); return OSQ_ECUDA; }
  const int row_in_smem = cols * 4 <= 160 * 1024;
  const size_t smem = row_in_smem ? (size_t)cols * 4 : 0;
  // the opt-in is per device, not per process (one process may drive several GPUs)
  static bool attr_set[64] = {false};
  int dev = 0;
  OSQ_CUDA(cudaGetDevice(&dev));
  if (!attr_set[dev & 63]) {
    OSQ_CUDA(cudaFuncSetAttribute(mse_brent_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_set[dev & 63] = true;
  }
  int64_t g = rows < (int64_t)sms * 8 ? rows : (int64_t)sms * 8;
  mse_brent_rows_kernel<<<(int)g, kBrentThreads, smem, (cudaStream_t)stream>>>(w, rows, cols, (float)qmin, (float)qmax, symmetric,
                                                                             one_side, row_in_smem, out_min, out_max, evals);
  OSQ_LAUNCH_CHECK();
  return OSQ_OK;
}

}  // extern "C"
