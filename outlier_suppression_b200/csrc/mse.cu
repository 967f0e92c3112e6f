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
__device__ double fminbound(const F& func, double x1, double x2, int* nfev, double* fval = nullptr) {
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
  if (fval) *fval = fx;
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


// ---------------------------------------------------------------------------------------------
// K5c: (Avg)MSEFastObserver per-TENSOR search (observer.py:434-494) as ONE cooperative launch.  Every CTA runs the same
// bounded-Brent state machine(s) in fp64 on identical loss values; each loss evaluation is a grid-wide masked reduction
// over the activation (which stays L2 resident: 12.6 MB at [32, 128, 768]) -- per-CTA fp64 partials, one grid barrier,
// every CTA folds the partials in the same order.  The asymmetric two-sided case nests the search exactly like the
// reference: an outer Brent over the range whose objective is the minimum of an inner Brent over the shift, then one
// more inner search at the best range (~640 loss evaluations); symmetric / one-sided tensors take the 1-D search.
// No host round trip at all: the reference (and round 1 of this backend) synchronise once or twice per evaluation.
// ---------------------------------------------------------------------------------------------
constexpr int kTensorThreads = 512;
constexpr int kTensorWarps = kTensorThreads / 32;
constexpr int kMaxGridCtas = 1024;

struct GridCtx {
  const float* x;
  osq_tokens_t tk;
  const int64_t* lens;
  int n_lens;
  float qmin, qmax;
  int symmetric;
  double* partials;       // [2][2][kMaxGridCtas] global: double-buffered per evaluation, two values per CTA
  unsigned int* barrier;  // global arrival counter (zeroed before the launch)
  unsigned int* epoch;    // this CTA's barrier count so far (register-held by the caller; pointer to a local)
  double* red;            // shared [2][kTensorWarps]
  double* bc;             // shared [2] broadcast
  double numel;
};

__device__ __forceinline__ void grid_barrier(const GridCtx& c) {
  __syncthreads();
  const unsigned int target = (++*c.epoch) * gridDim.x;   // every thread keeps its own (identical) count
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(c.barrier, 1u);
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(c.barrier) : "memory");
    } while (seen < target);
  }
  __syncthreads();
}

// grid-wide sum of two per-thread doubles (second may be unused); identical result in every CTA
__device__ __forceinline__ void grid_sum2(const GridCtx& c, double v0, double v1, double& s0, double& s1, bool take_minmax) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double a = __shfl_xor_sync(0xffffffffu, v0, o), b = __shfl_xor_sync(0xffffffffu, v1, o);
    if (take_minmax) { v0 = fmin(v0, a); v1 = fmax(v1, b); } else { v0 += a; v1 += b; }
  }
  if (lane == 0) { c.red[warp] = v0; c.red[kTensorWarps + warp] = v1; }
  __syncthreads();
  const unsigned int buf = (*c.epoch) & 1u;
  double* p0 = c.partials + (size_t)buf * 2 * kMaxGridCtas;
  double* p1 = p0 + kMaxGridCtas;
  if (threadIdx.x == 0) {
    double t0 = c.red[0], t1 = c.red[kTensorWarps];
    for (int w = 1; w < kTensorWarps; ++w) {
      if (take_minmax) { t0 = fmin(t0, c.red[w]); t1 = fmax(t1, c.red[kTensorWarps + w]); }
      else { t0 += c.red[w]; t1 += c.red[kTensorWarps + w]; }
    }
    p0[blockIdx.x] = t0;
    p1[blockIdx.x] = t1;
  }
  grid_barrier(c);
  if (warp == 0) {
    double t0 = take_minmax ? INFINITY : 0.0, t1 = take_minmax ? -INFINITY : 0.0;
    for (int i = lane; i < (int)gridDim.x; i += 32) {
      double a, b;
      asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(a) : "l"(p0 + i));
      asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(b) : "l"(p1 + i));
      if (take_minmax) { t0 = fmin(t0, a); t1 = fmax(t1, b); } else { t0 += a; t1 += b; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double a = __shfl_xor_sync(0xffffffffu, t0, o), b = __shfl_xor_sync(0xffffffffu, t1, o);
      if (take_minmax) { t0 = fmin(t0, a); t1 = fmax(t1, b); } else { t0 += a; t1 += b; }
    }
    if (lane == 0) { c.bc[0] = t0; c.bc[1] = t1; }
  }
  __syncthreads();
  s0 = c.bc[0];
  s1 = c.bc[1];
  __syncthreads();
}

// visits every valid element of this CTA's share of the masked activation
template <class F>
__device__ __forceinline__ void for_valid(const GridCtx& c, F&& f) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t)blockIdx.x * kTensorWarps + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * kTensorWarps;
  const osq_tokens_t& tk = c.tk;
  const int64_t n_seg = tk.B * tk.S * tk.F1;
  for (int64_t seg = warp_global; seg < n_seg; seg += n_warps) {
    const int64_t f1 = seg % tk.F1;
    const int64_t bs = seg / tk.F1;
    const int64_t sidx = bs % tk.S, b = bs / tk.S;
    if (c.lens != nullptr && (b >= c.n_lens || sidx >= c.lens[b])) continue;
    const float* p = c.x + b * tk.sb + sidx * tk.ss + f1 * tk.sf1;
    if (tk.sf2 == 1 && (((uintptr_t)p) & 15) == 0 && (tk.F2 & 3) == 0) {
      const float4* v = reinterpret_cast<const float4*>(p);
      for (int64_t i = lane; i < (tk.F2 >> 2); i += 32) {
        const float4 a = __ldg(v + i);   // re-read ~640 times: keep it cacheable
        f(a.x); f(a.y); f(a.z); f(a.w);
      }
    } else {
      for (int64_t i = lane; i < tk.F2; i += 32) f(__ldg(p + i * tk.sf2));
    }
  }
}

// loss_fx(new_min, new_max) of observer.py:420-432: qparams in fp64 (the candidates are np.float64), scale rounded to fp32
// by `.item()` -> fp32 division, loss = fp32 mean of the squared error (accumulated in fp64 here)
__device__ float grid_loss(const GridCtx& c, double new_min, double new_max) {
  const double span = (double)(c.qmax - c.qmin);
  const double lo = fmin(new_min, 0.0), hi = fmax(new_max, 0.0);
  double scale64;
  float zp = 0.f;
  const double eps = (double)1e-8f;
  if (c.symmetric) {
    scale64 = fmax(-lo, hi) / (span / 2.0);
    if (!(scale64 > eps)) scale64 = eps;
  } else {
    scale64 = (hi - lo) / span;
    if (!(scale64 > eps)) scale64 = eps;
    double z = (double)c.qmin - rint(lo / scale64);
    z = z < (double)c.qmin ? (double)c.qmin : (z > (double)c.qmax ? (double)c.qmax : z);
    zp = (float)z;
  }
  const float s = (float)scale64;
  float acc = 0.f;
  double dacc = 0.0;
  int run = 0;
  for_valid(c, [&](float v) {
    acc += sq_err(v, s, zp, c.qmin, c.qmax);
    if (++run == 64) { dacc += (double)acc; acc = 0.f; run = 0; }   // short fp32 runs
  });
  dacc += (double)acc;
  double tot, unused;
  grid_sum2(c, dacc, 0.0, tot, unused, false);
  return (float)(tot / c.numel);
}

struct Loss1D {
  const GridCtx* c;
  int one_side;  // 0 no, 1 pos, 2 neg
  __device__ float operator()(double r) const { return grid_loss(*c, one_side == 1 ? 0.0 : -r, one_side == 2 ? 0.0 : r); }
};
struct ShiftLoss {
  const GridCtx* c;
  double xrange, x_min, x_max;
  __device__ float operator()(double shift) const {
    return grid_loss(*c, fmax(0.0 - shift, x_min), fmin(xrange - shift, x_max));   // observer.py:436-440
  }
};
struct RangeLoss {
  const GridCtx* c;
  double x_min, x_max, span;
  float qmin, qmax;
  int* evals;
  __device__ float operator()(double xrange) const {
    const double d = xrange / span;
    ShiftLoss f{c, xrange, x_min, x_max};
    int n = 0;
    double fval = 0.0;
    fminbound(f, d * (double)qmin, d * (double)qmax, &n, &fval);
    *evals += n;
    return (float)fval;
  }
};

__global__ void __launch_bounds__(kTensorThreads, 1)
mse_brent_tensor_kernel(const float* __restrict__ x, osq_tokens_t tk, const int64_t* __restrict__ lens, int n_lens, float qmin,
                        float qmax, int symmetric, int* __restrict__ one_side_state, double* __restrict__ partials,
                        unsigned int* __restrict__ barrier, double* __restrict__ out /* best_min, best_max, x_min, x_max */,
                        int32_t* __restrict__ evals_out) {
  __shared__ double red[2 * kTensorWarps];
  __shared__ double bc[2];
  unsigned int epoch = 0;
  GridCtx c{x, tk, lens, n_lens, qmin, qmax, symmetric, partials, barrier, &epoch, red, bc, 0.0};
  // ---- masked min / max and element count (observer.py:500-505) ----
  float mn = INFINITY, mx = -INFINITY;
  double cnt = 0.0;
  for_valid(c, [&](float v) { mn = fminf(mn, v); mx = fmaxf(mx, v); cnt += 1.0; });
  double x_min, x_max, numel, unused;
  grid_sum2(c, (double)mn, (double)mx, x_min, x_max, true);
  grid_sum2(c, cnt, 0.0, numel, unused, false);
  c.numel = numel > 0.0 ? numel : 1.0;
  // one_side_dist is decided on the first batch and kept (observer.py:528-529): -1 = undecided
  int one_side = *one_side_state;
  if (one_side < 0) one_side = (x_min >= 0.0) ? 1 : ((x_max <= 0.0) ? 2 : 0);
  int evals = 0;
  double best_min, best_max;
  if (one_side != 0 || symmetric) {   // observer.py:483-494
    const double xrange = fmax(fabs(x_min), x_max);
    Loss1D f{&c, one_side};
    const double r = fminbound(f, fmin(0.1, 0.01 * xrange), xrange, &evals);
    best_min = one_side == 1 ? 0.0 : -r;
    best_max = one_side == 2 ? 0.0 : r;
  } else {                            // observer.py:458-481
    const double span = (double)(qmax - qmin);
    // the reference forms the range in fp32 (tensor arithmetic) before handing it to SciPy
    const double total = (double)((float)x_max - (float)x_min);
    RangeLoss outer{&c, x_min, x_max, span, qmin, qmax, &evals};
    const double final_range = fminbound(outer, fmin(0.1, 0.01 * total), total, nullptr);
    const double d = final_range / span;
    ShiftLoss inner{&c, final_range, x_min, x_max};
    int n = 0;
    const double final_shift = fminbound(inner, d * (double)qmin, d * (double)qmax, &n);
    evals += n;
    best_min = fmax(0.0 - final_shift, x_min);
    best_max = fmin(final_range - final_shift, x_max);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    out[0] = best_min; out[1] = best_max; out[2] = x_min; out[3] = x_max;
    *one_side_state = one_side;
    if (evals_out) *evals_out = evals;
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


int osq_mse_brent_tensor_f32(const float* x, const osq_tokens_t* tok, const int64_t* lens, int n_lens, int qmin, int qmax,
                             int symmetric, int32_t* one_side_state, double* out4, int32_t* evals, void* scratch, void* stream) {
  using namespace osq;
  OSQ_CHECK_ARG(x && tok && one_side_state && out4 && scratch, "osq_mse_brent_tensor_f32: null pointer");
  OSQ_CHECK_ARG(qmin < qmax, "osq_mse_brent_tensor_f32: qmin >= qmax");
  int sms = sm_count();
  if (sms <= 0) { set_error("no CUDA device"); return OSQ_ECUDA; }
  int dev = 0, coop = 0;
  OSQ_CUDA(cudaGetDevice(&dev));
  OSQ_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  OSQ_CHECK_ARG(coop != 0, "osq_mse_brent_tensor_f32: device does not support cooperative launches");
  int per_sm = 0;
  OSQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mse_brent_tensor_kernel, kTensorThreads, 0));
  if (per_sm < 1) { set_error("osq_mse_brent_tensor_f32: kernel does not fit on an SM"); return OSQ_ECUDA; }
  int64_t n_seg = tok->B * tok->S * tok->F1;
  int64_t g = (n_seg + kTensorWarps - 1) / kTensorWarps;
  if (g > sms) g = sms;                 // one CTA per SM: the barrier, not the bandwidth, paces the ~640 evaluations
  if (g < 1) g = 1;
  if (g > kMaxGridCtas) g = kMaxGridCtas;
  cudaStream_t st = (cudaStream_t)stream;
  // scratch: [2][2][kMaxGridCtas] doubles, then the barrier counter
  double* partials = static_cast<double*>(scratch);
  unsigned int* barrier = reinterpret_cast<unsigned int*>(partials + 4 * kMaxGridCtas);
  OSQ_CUDA(cudaMemsetAsync(barrier, 0, sizeof(unsigned int), st));
  osq_tokens_t tk = *tok;
  float fqmin = (float)qmin, fqmax = (float)qmax;
  void* args[] = {(void*)&x, (void*)&tk, (void*)&lens, (void*)&n_lens, (void*)&fqmin, (void*)&fqmax, (void*)&symmetric,
                  (void*)&one_side_state, (void*)&partials, (void*)&barrier, (void*)&out4, (void*)&evals};
  OSQ_CUDA(cudaLaunchCooperativeKernel((const void*)mse_brent_tensor_kernel, dim3((unsigned)g), dim3(kTensorThreads), args, 0, st));
  return OSQ_OK;
}

int64_t osq_mse_tensor_scratch_bytes(void) { return (int64_t)(4 * osq::kMaxGridCtas) * 8 + 64; }

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
