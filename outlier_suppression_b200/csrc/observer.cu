// observer.cu -- K3 / K4: calibration-observer reductions with the pad-token mask applied in-kernel.
//
// All of these read the activation exactly once (4 algorithmic bytes / element) and are HBM-bound.
// Work unit = one contiguous feature segment (F2 floats) of one token; a warp owns a segment,
// lanes issue 128-bit streaming loads, results fold through warp shuffles, then per-CTA partials,
// then the last CTA to finish (ticket counter) folds the partials and runs the running-statistics
// epilogue (running average / extrema + calculate_qparams) so a whole observer.forward() is ONE
// launch with no host synchronisation.
#include <stdlib.h>

#include "common.cuh"

namespace osq {

__device__ __forceinline__ long long obs_gtimer() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define OBS_STAMP(slot) do { if (threadIdx.x == 0 && blockIdx.x == 0 && trace != nullptr) trace[(slot)] = obs_gtimer(); } while (0)

constexpr int kObsThreads = 512;
constexpr int kObsWarps = kObsThreads / 32;

struct Partials {
  float* pmin;
  float* pmax;
  unsigned int* ticket;
};
__host__ __device__ inline Partials carve(void* ws) {
  Partials p;
  p.pmin = (float*)ws;
  p.pmax = p.pmin + kMaxPartialBlocks;
  p.ticket = (unsigned int*)(p.pmax + kMaxPartialBlocks);
  return p;
}

__device__ __forceinline__ bool token_valid(const int64_t* lens, int n_lens, int64_t b, int64_t s) {
  if (lens == nullptr) return true;
  if (b >= n_lens) return false;  // zip(observation_mask, x) stops at the shorter one (observer.py:82)
  return s < lens[b];
}

// min/max over one contiguous run of `len` floats (stride 1) handled by a full warp
__device__ __forceinline__ void warp_scan_segment(const float* __restrict__ p, int64_t len, int lane,
                                                  float& mn, float& mx) {
  if ((((uintptr_t)p) & 15) == 0 && len >= 128) {
    const float4* v = reinterpret_cast<const float4*>(p);
    const int64_t nv = len >> 2;
    int64_t i = lane;
    for (; i + 96 < nv; i += 128) {  // 4 independent 128-bit loads in flight per lane
      float4 a = ldg_stream(v + i), b = ldg_stream(v + i + 32), c = ldg_stream(v + i + 64), d = ldg_stream(v + i + 96);
      mn = fminf(mn, fminf(fminf(fminf(a.x, a.y), fminf(a.z, a.w)), fminf(fminf(b.x, b.y), fminf(b.z, b.w))));
      mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)), fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w))));
      mn = fminf(mn, fminf(fminf(fminf(c.x, c.y), fminf(c.z, c.w)), fminf(fminf(d.x, d.y), fminf(d.z, d.w))));
      mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(c.x, c.y), fmaxf(c.z, c.w)), fmaxf(fmaxf(d.x, d.y), fmaxf(d.z, d.w))));
    }
    for (; i < nv; i += 32) {
      float4 a = ldg_stream(v + i);
      mn = fminf(mn, fminf(fminf(a.x, a.y), fminf(a.z, a.w)));
      mx = fmaxf(mx, fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)));
    }
    for (int64_t j = (nv << 2) + lane; j < len; j += 32) {
      float a = p[j];
      mn = fminf(mn, a);
      mx = fmaxf(mx, a);
    }
  } else {
    for (int64_t j = lane; j < len; j += 32) {
      float a = __ldg(p + j);
      mn = fminf(mn, a);
      mx = fmaxf(mx, a);
    }
  }
}

// generic strided segment (sf2 != 1): lanes stride over f2
__device__ __forceinline__ void warp_scan_strided(const float* __restrict__ p, int64_t len, int64_t stride,
                                                  int lane, float& mn, float& mx) {
  for (int64_t j = lane; j < len; j += 32) {
    float a = __ldg(p + j * stride);
    mn = fminf(mn, a);
    mx = fmaxf(mx, a);
  }
}

// Folds per-CTA (mn, mx) into the workspace; the last CTA reduces all partials.
// Returns true in thread 0 of the last CTA with the final values in (mn, mx).
__device__ __forceinline__ bool grid_fold(float& mn, float& mx, Partials ws) {
  __shared__ float smn[kObsWarps], smx[kObsWarps];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  mn = warp_min(mn);
  mx = warp_max(mx);
  if (lane == 0) { smn[warp] = mn; smx[warp] = mx; }
  __syncthreads();
  if (warp == 0) {
    mn = lane < (blockDim.x >> 5) ? smn[lane] : INFINITY;
    mx = lane < (blockDim.x >> 5) ? smx[lane] : -INFINITY;
    mn = warp_min(mn);
    mx = warp_max(mx);
    if (lane == 0) {
      ws.pmin[blockIdx.x] = mn;
      ws.pmax[blockIdx.x] = mx;
      __threadfence();
      unsigned int t = atomicAdd(ws.ticket, 1u);
      is_last = (t == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (!is_last) return false;
  __threadfence();
  mn = INFINITY;
  mx = -INFINITY;
  for (int i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
    mn = fminf(mn, __ldcg(ws.pmin + i));
    mx = fmaxf(mx, __ldcg(ws.pmax + i));
  }
  mn = warp_min(mn);
  mx = warp_max(mx);
  __syncthreads();
  if (lane == 0) { smn[warp] = mn; smx[warp] = mx; }
  __syncthreads();
  if (warp == 0) {
    mn = lane < (blockDim.x >> 5) ? smn[lane] : INFINITY;
    mx = lane < (blockDim.x >> 5) ? smx[lane] : -INFINITY;
    mn = warp_min(mn);
    mx = warp_max(mx);
    if (lane == 0) {
      *ws.ticket = 0;  // re-arm the workspace for the next launch on this stream
      return true;
    }
  }
  return false;
}

// ---------------------------------------------------------------------------------------------
// K3: masked global min/max
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kObsThreads)
minmax_masked_kernel(const float* __restrict__ x, osq_tokens_t tk, const int64_t* __restrict__ lens,
                     int n_lens, float* __restrict__ cur, osq_stat_epilogue_t epi, void* wsp) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t)blockIdx.x * kObsWarps + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * kObsWarps;
  const int64_t n_seg = tk.B * tk.S * tk.F1;
  float mn = INFINITY, mx = -INFINITY;
  for (int64_t seg = warp_global; seg < n_seg; seg += n_warps) {
    const int64_t f1 = seg % tk.F1;
    const int64_t bs = seg / tk.F1;
    const int64_t s = bs % tk.S, b = bs / tk.S;
    if (!token_valid(lens, n_lens, b, s)) continue;
    const float* p = x + b * tk.sb + s * tk.ss + f1 * tk.sf1;
    if (tk.sf2 == 1) warp_scan_segment(p, tk.F2, lane, mn, mx);
    else warp_scan_strided(p, tk.F2, tk.sf2, lane, mn, mx);
  }
  if (grid_fold(mn, mx, carve(wsp))) {
    cur[0] = mn;
    cur[1] = mx;
    stat_epilogue(epi, mn, mx);
  }
}

__global__ void __launch_bounds__(kObsThreads)
minmax_flat_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ cur,
                   osq_stat_epilogue_t epi, void* wsp) {
  // contiguous: every CTA takes an equal slab, warps split the slab
  const int lane = threadIdx.x & 31;
  const int64_t chunk = 4096;  // floats per warp step (16 KB)
  const int64_t warp_global = (int64_t)blockIdx.x * kObsWarps + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * kObsWarps;
  float mn = INFINITY, mx = -INFINITY;
  for (int64_t off = warp_global * chunk; off < n; off += n_warps * chunk) {
    int64_t len = n - off < chunk ? n - off : chunk;
    warp_scan_segment(x + off, len, lane, mn, mx);
  }
  if (grid_fold(mn, mx, carve(wsp))) {
    cur[0] = mn;
    cur[1] = mx;
    stat_epilogue(epi, mn, mx);
  }
}

// ---------------------------------------------------------------------------------------------
// K4a: per-token min/max (one warp per token)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kObsThreads)
token_minmax_kernel(const float* __restrict__ x, osq_tokens_t tk, const int64_t* __restrict__ lens,
                    int n_lens, float* __restrict__ tmin, float* __restrict__ tmax,
                    int32_t* __restrict__ n_valid, unsigned int* __restrict__ hist0, int end_wait) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t)blockIdx.x * kObsWarps + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * kObsWarps;
  const int64_t n_tok = tk.B * tk.S;
  // a dependent launch (the cluster select of osq_prune_observe_f32) may be scheduled as soon as SMs free up; it parks
  // on griddepcontrol.wait until this grid has completed and flushed
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // first radix digit of the select that follows: counted per CTA in shared memory (one atomic per token and side, by
  // lane 0 of the token's warp), flushed as a handful of L2 reductions at the end -- same-address L2 atomics serialise,
  // and the leading digits of per-token extrema are nearly constant
  __shared__ unsigned int sh0[2 * 2048];
  long long* trace = reinterpret_cast<long long*>(hist0 + 2 * 2048);   // workspace: the stamps follow the table
  if (hist0 != nullptr) {
    OBS_STAMP(0);
    for (int i = threadIdx.x; i < 2 * 2048; i += blockDim.x) sh0[i] = 0;
    __syncthreads();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int64_t t = 0;
    if (lens == nullptr) t = n_tok;
    else
      for (int64_t b = 0; b < tk.B && b < n_lens; ++b) {
        int64_t l = lens[b];
        t += l < 0 ? 0 : (l > tk.S ? tk.S : l);
      }
    *n_valid = (int32_t)t;
  }
  for (int64_t t = warp_global; t < n_tok; t += n_warps) {
    const int64_t s = t % tk.S, b = t / tk.S;
    float mn = INFINITY, mx = -INFINITY;
    if (token_valid(lens, n_lens, b, s)) {
      const float* p = x + b * tk.sb + s * tk.ss;
      for (int64_t f1 = 0; f1 < tk.F1; ++f1) {
        if (tk.sf2 == 1) warp_scan_segment(p + f1 * tk.sf1, tk.F2, lane, mn, mx);
        else warp_scan_strided(p + f1 * tk.sf1, tk.F2, tk.sf2, lane, mn, mx);
      }
      mn = warp_min(mn);
      mx = warp_max(mx);
    }
    if (lane == 0) {
      tmin[t] = mn;
      tmax[t] = mx;
      if (hist0 != nullptr && mn <= mx) {
        atomicAdd(&sh0[__float_as_uint(fabsf(mx)) >> 20], 1u);
        atomicAdd(&sh0[2048 + (__float_as_uint(fabsf(mn)) >> 20)], 1u);
      }
    }
  }
  if (hist0 != nullptr) {
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * 2048; i += blockDim.x) {
      const unsigned int c = sh0[i];
      if (c) atomicAdd(hist0 + i, c);   // fire-and-forget reduction in L2
    }
    OBS_STAMP(1);
  }
  // osq_prune_observe_many_f32 launches this grid as a programmatic dependent of the PREVIOUS batch's one-CTA select tail, so it runs
  // next to that tail instead of behind it (it touches none of the tail's buffers: vectors and tables alternate).  The tails
  // themselves must stay ordered (both update the running statistics): one thread holds this grid open until every prerequisite
  // grid has completed.
  if (end_wait && blockIdx.x == 0 && threadIdx.x == 0) asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// K4b: prune selection on the [T] vectors (single CTA; the vectors are tiny next to the activation)
// ---------------------------------------------------------------------------------------------
// torch.quantile(v, p) on an ascending-sorted fp32 vector of n entries ('linear' interpolation):
// rank = p*(n-1) in fp32, lerp with the fused multiply-add form ATen's CPU kernel uses.
__device__ __forceinline__ float quantile_sorted(const float* __restrict__ sorted, int n, float p) {
  float rank = __fmul_rn(p, (float)(n - 1));
  int lo = (int)rank;  // trunc, like .toType(kLong)
  int hi = (int)ceilf(rank);
  float w = __fsub_rn(rank, (float)lo);
  float a = sorted[lo], b = sorted[hi];
  float d = __fsub_rn(b, a);
  return (w < 0.5f) ? fmaf(w, d, a) : fmaf(-d, __fsub_rn(1.f, w), b);
}

__global__ void __launch_bounds__(1024)
prune_select_kernel(const float* __restrict__ tmin, const float* __restrict__ tmax,
                    const float* __restrict__ abs_tmin_sorted, const float* __restrict__ abs_tmax_sorted,
                    int64_t n_slots, const int32_t* __restrict__ n_valid, float percentile,
                    float* __restrict__ cur, osq_stat_epilogue_t epi) {
  __shared__ float smn[32], smx[32];
  const int T = *n_valid;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float lower = INFINITY, upper = -INFINITY;
  if (T > 0) {
    const float up_thr = quantile_sorted(abs_tmax_sorted, T, percentile);
    const float lo_thr = -quantile_sorted(abs_tmin_sorted, T, percentile);
    for (int64_t i = threadIdx.x; i < n_slots; i += blockDim.x) {
      float a = tmin[i], b = tmax[i];
      if (a <= b) {  // valid token (invalid ones hold +inf / -inf)
        if (b <= up_thr) upper = fmaxf(upper, b);
        if (a >= lo_thr) lower = fminf(lower, a);
      }
    }
  }
  lower = warp_min(lower);
  upper = warp_max(upper);
  if (lane == 0) { smn[warp] = lower; smx[warp] = upper; }
  __syncthreads();
  if (warp == 0) {
    lower = lane < (blockDim.x >> 5) ? smn[lane] : INFINITY;
    upper = lane < (blockDim.x >> 5) ? smx[lane] : -INFINITY;
    lower = warp_min(lower);
    upper = warp_max(upper);
    if (lane == 0) {
      cur[0] = lower;
      cur[1] = upper;
      stat_epilogue(epi, lower, upper);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K4b': the same selection WITHOUT a sort: exact order statistics of |tmax| and |tmin| by an 8-bit-per-pass
// radix select on the fp32 bit patterns (monotone for non-negative floats), one CTA, one launch.
// Replaces two torch.abs + two torch.sort (>= 10 launches) of the first version.
// ---------------------------------------------------------------------------------------------
// single-CTA sweep over a small vector with 8 independent loads in flight per thread (the [T] vectors live in L2;
// with one dependent load per iteration every pass would cost 64 L2 round trips)
template <class F>
__device__ __forceinline__ void sweep8(const float* __restrict__ v, int64_t n, F&& f, int64_t first = -1, int64_t stride = 0) {
  if (first < 0) { first = threadIdx.x; stride = blockDim.x; }
  for (int64_t base = first; base < n; base += stride * 8) {
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = (base + j * stride < n) ? __ldg(v + base + j * stride) : __int_as_float(0x7fc00000);
#pragma unroll
    for (int j = 0; j < 8; ++j) f(x[j], base + j * stride < n);
  }
}

struct SelectScratch {
  unsigned int hist[256];
  unsigned int prefix, k_rem;
  float fmin_above;
  unsigned int cnt_le;
};

// warp-aggregated histogram update: lanes with the same digit elect one leader (the leading digits of fp32
// magnitudes are nearly constant, a plain atomicAdd would serialise the whole warp on one bin)
__device__ __forceinline__ void hist_add(unsigned int* hist, float xv, bool ok, unsigned int mask, unsigned int prefix, int pass) {
  const unsigned int u = __float_as_uint(fabsf(xv));
  const bool in = ok && (u & mask) == prefix;
  const unsigned int digit = in ? ((u >> (8 * pass)) & 255u) : 256u;
  const unsigned int peers = __match_any_sync(__activemask(), digit);
  if (in && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&hist[digit], (unsigned int)__popc(peers));
}

// one full warp: first bin with (count of all lower bins) + hist[bin] > krem, and that count of lower bins
__device__ __forceinline__ void warp_pick_bin(const unsigned int* hist, unsigned int krem, unsigned int& bin, unsigned int& below) {
  const int lane = threadIdx.x & 31;
  unsigned int h[8], sum = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { h[j] = hist[lane * 8 + j]; sum += h[j]; }
  unsigned int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const unsigned int vote = __ballot_sync(0xffffffffu, incl > krem);
  const int src = vote ? __ffs(vote) - 1 : 31;
  unsigned int acc = incl - sum, pick = lane * 8 + 7;
#pragma unroll
  for (int j = 7; j >= 0; --j) {  // descending so the smallest qualifying j wins
    unsigned int lower = incl - sum;
#pragma unroll
    for (int i = 0; i < j; ++i) lower += h[i];
    if (lower + h[j] > krem) { pick = lane * 8 + j; acc = lower; }
  }
  bin = __shfl_sync(0xffffffffu, pick, src);
  below = __shfl_sync(0xffffffffu, acc, src);
}

// value of rank k (0-based, ascending) of {|v[i]|}; also the value of rank k+1 (needed by the lerp)
template <bool kIsMax>
__device__ void select_two(const float* __restrict__ v, int64_t n, int k, int n_valid, SelectScratch& sc, float& a, float& b) {
  unsigned int prefix = 0, mask = 0, krem = (unsigned int)k;
  for (int pass = 3; pass >= 0; --pass) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sc.hist[i] = 0;
    __syncthreads();
    sweep8(v, n, [&](float xv, bool ok) { hist_add(sc.hist, xv, ok, mask, prefix, pass); });
    __syncthreads();
    if (threadIdx.x < 32) {
      unsigned int bin, below;
      warp_pick_bin(sc.hist, krem, bin, below);
      if (threadIdx.x == 0) {
        sc.prefix = prefix | (bin << (8 * pass));
        sc.k_rem = krem - below;
      }
    }
    __syncthreads();
    prefix = sc.prefix;
    krem = sc.k_rem;
    mask |= 0xFFu << (8 * pass);
    __syncthreads();
  }
  a = __uint_as_float(prefix);
  // rank k+1: a again if it has duplicates beyond rank k, else the smallest value above a
  if (threadIdx.x == 0) { sc.cnt_le = 0; sc.fmin_above = INFINITY; }
  __syncthreads();
  unsigned int cnt = 0;
  float above = INFINITY;
  sweep8(v, n, [&](float xv, bool ok) {
    const float x = fabsf(xv);
    if (ok) { if (x <= a) ++cnt; else above = fminf(above, x); }
  });
  above = warp_min(above);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&sc.cnt_le, cnt);
    atomicMin(reinterpret_cast<unsigned int*>(&sc.fmin_above), __float_as_uint(above));  // non-negative floats order like uints
  }
  __syncthreads();
  b = (sc.cnt_le >= (unsigned int)k + 2u || k + 1 >= n_valid) ? a : sc.fmin_above;
  __syncthreads();
}

__device__ __forceinline__ float quantile_from_pair(float a, float b, float rank, int lo) {
  const float w = __fsub_rn(rank, (float)lo);
  const float d = __fsub_rn(b, a);
  return (w < 0.5f) ? fmaf(w, d, a) : fmaf(-d, __fsub_rn(1.f, w), b);
}

__global__ void __launch_bounds__(1024)
prune_select_unsorted_kernel(const float* __restrict__ tmin, const float* __restrict__ tmax, int64_t n_slots,
                             const int32_t* __restrict__ n_valid, float percentile, float* __restrict__ cur,
                             osq_stat_epilogue_t epi) {
  __shared__ SelectScratch sc;
  __shared__ float smn[32], smx[32];
  const int T = *n_valid;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float lower = INFINITY, upper = -INFINITY;
  if (T > 0) {
    // torch.quantile: rank = p * (T - 1) in fp32, below = trunc(rank), above = ceil(rank), lerp
    const float rank = __fmul_rn(percentile, (float)(T - 1));
    const int lo = (int)rank;
    const bool need_pair = (int)ceilf(rank) != lo;
    float a, b;
    select_two<true>(tmax, n_slots, lo, T, sc, a, b);
    const float up_thr = need_pair ? quantile_from_pair(a, b, rank, lo) : quantile_from_pair(a, a, rank, lo);
    select_two<false>(tmin, n_slots, lo, T, sc, a, b);
    const float lo_thr = -(need_pair ? quantile_from_pair(a, b, rank, lo) : quantile_from_pair(a, a, rank, lo));
    const int64_t stride = blockDim.x;
    for (int64_t base = threadIdx.x; base < n_slots; base += stride * 4) {  // 8 independent loads in flight
      float mn[4], mx[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool ok = base + j * stride < n_slots;
        mn[j] = ok ? __ldg(tmin + base + j * stride) : INFINITY;
        mx[j] = ok ? __ldg(tmax + base + j * stride) : -INFINITY;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (mn[j] <= mx[j]) {  // valid token (invalid ones hold +inf / -inf)
          if (mx[j] <= up_thr) upper = fmaxf(upper, mx[j]);
          if (mn[j] >= lo_thr) lower = fminf(lower, mn[j]);
        }
    }
  }
  lower = warp_min(lower);
  upper = warp_max(upper);
  if (lane == 0) { smn[warp] = lower; smx[warp] = upper; }
  __syncthreads();
  if (warp == 0) {
    lower = lane < (blockDim.x >> 5) ? smn[lane] : INFINITY;
    upper = lane < (blockDim.x >> 5) ? smx[lane] : -INFINITY;
    lower = warp_min(lower);
    upper = warp_max(upper);
    if (lane == 0) {
      cur[0] = lower;
      cur[1] = upper;
      stat_epilogue(epi, lower, upper);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K4b'': the same radix select spread over many CTAs for long token vectors (T >> 64K: a calibration batch
// of 1024 x 512 tokens has 4 MB of per-token extrema, one CTA would crawl through it ten times).
// One launch per 8-bit digit (both sides at once), one for the rank-(k+1) value, one for clip + aminmax +
// running statistics.  State lives in the caller's workspace; the last CTA of every launch (ticket) folds the
// global histogram and re-arms it, so the whole sequence needs no host synchronisation and no memset.
// ---------------------------------------------------------------------------------------------
constexpr int64_t kSelectSingleCtaMax = 32768;

struct SelectWs {
  unsigned int hist[2][256];
  unsigned int prefix[2], krem[2];
  unsigned int cnt_le[2], above_s[2];  // above_s = 0x7f800000 - bits(min value above a): atomicMax, zero = +inf
  float thr[2];                        // up_thr, lo_thr
  unsigned int ticket;
};
static_assert(sizeof(SelectWs) <= 2 * 256 * 4 + 64, "workspace too small");

__device__ __forceinline__ bool last_cta(unsigned int* ticket) {
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
}

__global__ void __launch_bounds__(1024)
select_pass_kernel(const float* __restrict__ tmin, const float* __restrict__ tmax, int64_t n,
                   const int32_t* __restrict__ n_valid, float percentile, int pass, SelectWs* __restrict__ ws) {
  __shared__ unsigned int hist[2][256];
  const int T = *n_valid;
  if (T <= 0) return;
  unsigned int prefix[2], krem[2];
  const unsigned int mask = pass == 3 ? 0u : (0xFFFFFFFFu << (8 * (pass + 1)));
  if (pass == 3) {
    prefix[0] = prefix[1] = 0;
    krem[0] = krem[1] = (unsigned int)(int)__fmul_rn(percentile, (float)(T - 1));
  } else {
    prefix[0] = __ldcg(&ws->prefix[0]); prefix[1] = __ldcg(&ws->prefix[1]);
    krem[0] = __ldcg(&ws->krem[0]); krem[1] = __ldcg(&ws->krem[1]);
  }
  for (int i = threadIdx.x; i < 512; i += blockDim.x) (&hist[0][0])[i] = 0;
  __syncthreads();
  const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  sweep8(tmax, n, [&](float xv, bool ok) { hist_add(hist[0], xv, ok, mask, prefix[0], pass); }, first, stride);
  sweep8(tmin, n, [&](float xv, bool ok) { hist_add(hist[1], xv, ok, mask, prefix[1], pass); }, first, stride);
  __syncthreads();
  for (int i = threadIdx.x; i < 512; i += blockDim.x) {
    const unsigned int c = (&hist[0][0])[i];
    if (c) atomicAdd(&ws->hist[0][0] + i, c);
  }
  if (!last_cta(&ws->ticket)) return;
  if (threadIdx.x < 64) {
    const int side = threadIdx.x >> 5;
    for (int i = threadIdx.x & 31; i < 256; i += 32) hist[side][i] = __ldcg(&ws->hist[side][i]);
    __syncwarp();
    unsigned int bin, below;
    warp_pick_bin(hist[side], krem[side], bin, below);
    if ((threadIdx.x & 31) == 0) {
      ws->prefix[side] = prefix[side] | (bin << (8 * pass));
      ws->krem[side] = krem[side] - below;
    }
  }
  __syncthreads();  // the two scanning warps have read the global histogram; now re-arm it
  for (int i = threadIdx.x; i < 512; i += blockDim.x) (&ws->hist[0][0])[i] = 0;
  if (threadIdx.x == 0) ws->ticket = 0;
}

__global__ void __launch_bounds__(1024)
select_final_kernel(const float* __restrict__ tmin, const float* __restrict__ tmax, int64_t n,
                    const int32_t* __restrict__ n_valid, float percentile, SelectWs* __restrict__ ws) {
  const int T = *n_valid;
  if (T <= 0) return;
  const float a0 = __uint_as_float(__ldcg(&ws->prefix[0])), a1 = __uint_as_float(__ldcg(&ws->prefix[1]));
  unsigned int cnt[2] = {0, 0};
  float above[2] = {INFINITY, INFINITY};
  const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  sweep8(tmax, n, [&](float xv, bool ok) { const float x = fabsf(xv); if (ok) { if (x <= a0) ++cnt[0]; else above[0] = fminf(above[0], x); } }, first, stride);
  sweep8(tmin, n, [&](float xv, bool ok) { const float x = fabsf(xv); if (ok) { if (x <= a1) ++cnt[1]; else above[1] = fminf(above[1], x); } }, first, stride);
#pragma unroll
  for (int sd = 0; sd < 2; ++sd) {
    above[sd] = warp_min(above[sd]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt[sd] += __shfl_xor_sync(0xffffffffu, cnt[sd], o);
    if ((threadIdx.x & 31) == 0) {
      if (cnt[sd]) atomicAdd(&ws->cnt_le[sd], cnt[sd]);
      if (above[sd] < INFINITY) atomicMax(&ws->above_s[sd], 0x7f800000u - __float_as_uint(above[sd]));
    }
  }
  if (!last_cta(&ws->ticket)) return;
  if (threadIdx.x == 0) {
    const float rank = __fmul_rn(percentile, (float)(T - 1));
    const int lo = (int)rank;
    const bool need_pair = (int)ceilf(rank) != lo;
    float thr[2];
#pragma unroll
    for (int sd = 0; sd < 2; ++sd) {
      const float a = sd ? a1 : a0;
      const unsigned int c = __ldcg(&ws->cnt_le[sd]);
      const float nxt = __uint_as_float(0x7f800000u - __ldcg(&ws->above_s[sd]));
      const float b = (c >= (unsigned int)lo + 2u || lo + 1 >= T) ? a : nxt;
      thr[sd] = quantile_from_pair(a, need_pair ? b : a, rank, lo);
      ws->cnt_le[sd] = 0;
      ws->above_s[sd] = 0;
    }
    ws->thr[0] = thr[0];
    ws->thr[1] = -thr[1];
    ws->ticket = 0;
  }
}

__global__ void __launch_bounds__(kObsThreads)
prune_apply_kernel(const float* __restrict__ tmin, const float* __restrict__ tmax, int64_t n,
                   const int32_t* __restrict__ n_valid, const SelectWs* __restrict__ sw, float* __restrict__ cur,
                   osq_stat_epilogue_t epi, void* wsp) {
  float lower = INFINITY, upper = -INFINITY;
  if (*n_valid > 0) {
    const float up_thr = __ldcg(&sw->thr[0]), lo_thr = __ldcg(&sw->thr[1]);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; base < n; base += stride * 4) {
      float mn[4], mx[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool ok = base + j * stride < n;
        mn[j] = ok ? __ldg(tmin + base + j * stride) : INFINITY;
        mx[j] = ok ? __ldg(tmax + base + j * stride) : -INFINITY;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (mn[j] <= mx[j]) {
          if (mx[j] <= up_thr) upper = fmaxf(upper, mx[j]);
          if (mn[j] >= lo_thr) lower = fminf(lower, mn[j]);
        }
    }
  }
  if (grid_fold(lower, upper, carve(wsp))) {
    cur[0] = lower;
    cur[1] = upper;
    stat_epilogue(epi, lower, upper);
  }
}


// ---------------------------------------------------------------------------------------------
// K4c: the token-pruning tail of AvgPruneMinMaxObserver (observer.py:50-70; run ~94 k times by token-wise clipping) as ONE
// small CTA parked behind the per-token pass by programmatic dependent launch -- with no shared-memory atomics on the
// hot path (they cost ~2 cycles per lane: a histogram over all T tokens on one SM is what made the earlier single-CTA
// selects take 25-50 us at T = 16384).
//
//  (0) the FIRST radix digit (top 11 bits of the fp32 pattern of |tmax| / |tmin|) is histogrammed by the per-token pass
//      itself, per CTA in shared memory and flushed as a few L2 reductions -- spread over all SMs, hidden under the HBM stream;
//  (a) the tail reads that 2 x 2048-bin table, re-arms it, and picks the first-digit bin of rank lo on both sides;
//  (b) ONE sweep of the [T] vectors (L2, 128-bit loads) compacts the members of the two bins into shared-memory lists
//      (warp-aggregated appends: one atomic per warp) and records the smallest magnitude above each bin;
//  (c) the remaining 20 bits are resolved 4 at a time over the lists only: every thread counts its members' digits in
//      packed 16-bit register fields, warps combine them with shuffles, one warp per side picks the digit -- 5 passes of
//      two barriers each, both sides at once;
//  (c') one more pass over each list yields rank lo + 1 (the second value torch.quantile's lerp interpolates): the
//      selected value again if it has a duplicate, else the smallest member above it, else the smallest magnitude above
//      the bin;
//  (d) a second sweep of the vectors does the clip + aminmax; thread 0 runs the running-statistics epilogue.
// A bin with more members than a list holds is processed straight from L2 with a membership test (same code, different
// source), so any T < 2^21 and any value distribution works.
// ---------------------------------------------------------------------------------------------
constexpr int kSelThreads = 1024;
constexpr int kSelBins0 = 2048;        // first digit: bits [30:20]
constexpr int kSelCap = 8192;          // entries per candidate list (one per side)
constexpr int64_t kSelMaxTokens = (int64_t)1 << 21;

// first bin b of hist[0..BINS) with (sum of bins below b) + hist[b] > krem; executed by one full warp.
// Lane l owns bins [P l, P l + P), P = BINS / 32: it sums them in a rotated order (conflict free), a warp scan finds the
// owning lane, then the whole warp scans that lane's P bins, P / 32 per lane.
template <int BINS>
__device__ __forceinline__ void warp_pick_bin_wide(const unsigned int* hist, unsigned int krem, unsigned int& bin, unsigned int& below) {
  const int lane = threadIdx.x & 31;
  constexpr int kPer = BINS / 32, kPer2 = kPer / 32;
  static_assert(kPer % 32 == 0 && kPer2 >= 1 && kPer2 <= 2, "BINS must be 1024 or 2048");
  unsigned int sum = 0;
#pragma unroll 8
  for (int j = 0; j < kPer; ++j) sum += hist[lane * kPer + ((j + lane) & (kPer - 1))];
  unsigned int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const unsigned int vote = __ballot_sync(0xffffffffu, incl > krem);
  const int src = vote ? __ffs(vote) - 1 : 31;
  const unsigned int base = __shfl_sync(0xffffffffu, incl - sum, src);   // count of all bins below lane src's range
  const unsigned int h0 = hist[src * kPer + kPer2 * lane];
  const unsigned int h1 = kPer2 == 2 ? hist[src * kPer + kPer2 * lane + 1] : 0u;
  unsigned int inc2 = h0 + h1;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int t = __shfl_up_sync(0xffffffffu, inc2, o);
    if (lane >= o) inc2 += t;
  }
  const unsigned int v2 = __ballot_sync(0xffffffffu, base + inc2 > krem);
  const int l2 = v2 ? __ffs(v2) - 1 : 31;
  const unsigned int before = base + inc2 - (h0 + h1);                   // count below this lane's bins
  const bool first = (kPer2 == 1) || (before + h0 > krem);
  const unsigned int my_bin = (unsigned int)(src * kPer + kPer2 * lane + (first ? 0 : 1));
  const unsigned int my_below = first ? before : before + h0;
  bin = __shfl_sync(0xffffffffu, my_bin, l2);
  below = __shfl_sync(0xffffffffu, my_below, l2);
}

// calls f(tmin[i], tmax[i]) for every slot, 8 slots per thread in flight (two 128-bit loads per vector) when aligned.
// Every lane of a warp calls f the same number of times (f may use warp votes); slots past the end are presented as
// (+inf, -inf), i.e. as invalid tokens, which every f ignores.
template <class F>
__device__ __forceinline__ void sweep_pairs(const float* __restrict__ tmin, const float* __restrict__ tmax, int64_t n, F&& f) {
  const int tid = threadIdx.x, lane = tid & 31;
  const float4 inv_a = make_float4(INFINITY, INFINITY, INFINITY, INFINITY), inv_b = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  if (((((uintptr_t)tmin) | ((uintptr_t)tmax)) & 15) == 0) {
    const float4* a4 = reinterpret_cast<const float4*>(tmin);
    const float4* b4 = reinterpret_cast<const float4*>(tmax);
    const int64_t nv = n >> 2;
    for (int64_t base = tid - lane; base < nv; base += 2 * kSelThreads) {   // warp-uniform trip count
      const int64_t i = base + lane;
      const bool one = i < nv, two = i + kSelThreads < nv;
      const float4 a0 = one ? __ldcg(a4 + i) : inv_a, b0 = one ? __ldcg(b4 + i) : inv_b;
      const float4 a1 = two ? __ldcg(a4 + i + kSelThreads) : inv_a, b1 = two ? __ldcg(b4 + i + kSelThreads) : inv_b;
      f(a0.x, b0.x); f(a0.y, b0.y); f(a0.z, b0.z); f(a0.w, b0.w);
      f(a1.x, b1.x); f(a1.y, b1.y); f(a1.z, b1.z); f(a1.w, b1.w);
    }
    if (tid < 32 && (n & 3)) {                                              // up to three trailing slots: warp 0
      const int64_t i = (nv << 2) + lane;
      f(i < n ? __ldcg(tmin + i) : INFINITY, i < n ? __ldcg(tmax + i) : -INFINITY);
    }
  } else {
    for (int64_t base = tid - lane; base < n; base += kSelThreads) {
      const int64_t i = base + lane;
      f(i < n ? __ldcg(tmin + i) : INFINITY, i < n ? __ldcg(tmax + i) : -INFINITY);
    }
  }
}

// The same sweep handing out whole quads (four adjacent slots of both vectors): callers that only rarely act on an element test
// the quad first -- one warp vote per quad instead of one or two per element (cross-lane instructions are what a single SM
// runs out of in this kernel).  `kBatch` quads per thread are requested before the first one is used.
template <int kBatch, class F>
__device__ __forceinline__ void sweep_quads(const float* __restrict__ tmin, const float* __restrict__ tmax, int64_t n, F&& f) {
  const int tid = threadIdx.x, lane = tid & 31;
  const float4 inv_a = make_float4(INFINITY, INFINITY, INFINITY, INFINITY), inv_b = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  if (((((uintptr_t)tmin) | ((uintptr_t)tmax)) & 15) == 0) {
    const float4* a4 = reinterpret_cast<const float4*>(tmin);
    const float4* b4 = reinterpret_cast<const float4*>(tmax);
    const int64_t nv = n >> 2;
    for (int64_t base = tid - lane; base < nv; base += (int64_t)kBatch * kSelThreads) {   // warp-uniform trip count
      float4 a[kBatch], b[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int64_t i = base + lane + (int64_t)u * kSelThreads;
        a[u] = i < nv ? __ldcg(a4 + i) : inv_a;
        b[u] = i < nv ? __ldcg(b4 + i) : inv_b;
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) f(a[u], b[u]);
    }
    if (tid < 32 && (n & 3)) {                                              // up to three trailing slots: warp 0
      const int64_t i = (nv << 2) + lane;
      f(make_float4(i < n ? __ldcg(tmin + i) : INFINITY, INFINITY, INFINITY, INFINITY),
        make_float4(i < n ? __ldcg(tmax + i) : -INFINITY, -INFINITY, -INFINITY, -INFINITY));
    }
  } else {
    for (int64_t base = tid - lane; base < n; base += kSelThreads) {
      const int64_t i = base + lane;
      f(make_float4(i < n ? __ldcg(tmin + i) : INFINITY, INFINITY, INFINITY, INFINITY),
        make_float4(i < n ? __ldcg(tmax + i) : -INFINITY, -INFINITY, -INFINITY, -INFINITY));
    }
  }
}
// total order of finite floats and infinities as unsigned integers (REDUX has integer min / max only)
__device__ __forceinline__ unsigned int f2ord(float f) { const unsigned int u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(unsigned int o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o); }

// sixteen 16-bit counters in eight 32-bit registers.  Combined across the warp with the REDUX unit (__reduce_add_sync:
// one instruction per word) -- a shuffle tree would cost 80 SHFL per warp, side and pass, and SHFL issues at one warp
// instruction per clock per SM: with 32 warps that alone was 13 us.
struct Packed16 {
  unsigned int q[8];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = 0u;
  }
  __device__ __forceinline__ void add(unsigned int digit, bool pred) {  // no dynamic register indexing
    const unsigned int inc = pred ? (1u << ((digit & 1u) << 4)) : 0u;
    const unsigned int w = digit >> 1;
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] += (w == (unsigned int)i) ? inc : 0u;
  }
  __device__ __forceinline__ void warp_sum() {   // fields never overflow (<= 65535 members per warp and pass)
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = __reduce_add_sync(0xffffffffu, q[i]);
  }
  __device__ __forceinline__ unsigned int get(int j) const { return (q[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu; }
};

// `keep_table`: the first-digit table belongs to a cache (osq_prune_select_cached_f32) and is left as it is; otherwise it is the
// workspace table of the launch pair and is re-armed (zeroed) for the next call on this stream.  `trace`: optional stamps.
static __device__ __forceinline__ void
prune_select_tail_body(const float* __restrict__ tmin, const float* __restrict__ tmax, int64_t n_slots,
                       const int32_t* __restrict__ n_valid, float percentile, const float* __restrict__ percentile_dev,
                       unsigned int* __restrict__ ghist0, float* __restrict__ cur, const osq_stat_epilogue_t& epi, const bool keep_table,
                       long long* trace, const bool release_early = false) {
  extern __shared__ __align__(16) unsigned int sel_smem[];
  unsigned int* h0 = sel_smem;                                                              // [2][kSelBins0]
  unsigned int (*list)[kSelCap] = reinterpret_cast<unsigned int (*)[kSelCap]>(h0 + 2 * kSelBins0);  // [2][kSelCap]
  __shared__ unsigned int part[2][32][16];     // per-warp digit counts of the current pass
  __shared__ unsigned int s_prefix[2], s_krem[2], s_nbin[2], s_nlist[2], s_minabove[2], s_cntle[2], s_next[2];
  __shared__ float red[2][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  OBS_STAMP(2);
  asm volatile("griddepcontrol.wait;" ::: "memory");  // token_minmax_kernel's vectors, n_valid and first-digit table are complete
  if (release_early) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the next batch's per-token pass starts next to this CTA
  OBS_STAMP(3);
  const int T = *n_valid;
  if (percentile_dev != nullptr) percentile = fminf(fmaxf(*percentile_dev, 0.f), 1.f);   // CUDA-graph replays: the ratio is data
  const float frank = __fmul_rn(percentile, (float)(T > 0 ? T - 1 : 0));   // torch.quantile: rank = p * (T - 1) in fp32
  const int lo = (int)frank;
  const bool need_pair = (int)ceilf(frank) != lo;
  float thr_up = INFINITY, thr_lo = -INFINITY;

  // ---- (a) first digit: read the table (and re-arm it for the next call on this stream), pick the two bins ----
  for (int i = tid; i < 2 * kSelBins0; i += kSelThreads) {
    const unsigned int c = __ldcg(ghist0 + i);
    h0[i] = c;
    if (c && !keep_table) ghist0[i] = 0;
  }
  if (tid < 2) { s_nlist[tid] = 0; s_minabove[tid] = 0xFFFFFFFFu; s_cntle[tid] = 0; s_next[tid] = 0xFFFFFFFFu; }
  __syncthreads();
  if (T > 0) {
    if (warp < 2) {   // side 0 = |tmax|, side 1 = |tmin|
      unsigned int bin, below;
      warp_pick_bin_wide<kSelBins0>(h0 + warp * kSelBins0, (unsigned int)lo, bin, below);
      if (lane == 0) { s_prefix[warp] = bin << 20; s_krem[warp] = (unsigned int)lo - below; s_nbin[warp] = h0[warp * kSelBins0 + bin]; }
    }
    __syncthreads();
    OBS_STAMP(4);
    // ---- (b) one sweep over the vectors: compact the members of the chosen bins, smallest magnitude above each bin ----
    const unsigned int bin_mx = s_prefix[0] >> 20, bin_mn = s_prefix[1] >> 20;
    const bool listed_mx = s_nbin[0] <= kSelCap, listed_mn = s_nbin[1] <= kSelCap;
    unsigned int above_mx = 0xFFFFFFFFu, above_mn = 0xFFFFFFFFu;
    auto append = [&](int side, bool pred, unsigned int u) {   // warp-aggregated: one shared atomic per warp and call
      const unsigned int m = __ballot_sync(0xffffffffu, pred);
      if (m == 0) return;
      const int leader = __ffs(m) - 1;
      unsigned int base = 0;
      if (lane == leader) base = atomicAdd(&s_nlist[side], (unsigned int)__popc(m));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (pred) list[side][base + __popc(m & ((1u << lane) - 1u))] = u;
    };
    sweep_quads<4>(tmin, tmax, n_slots, [&](const float4 a4, const float4 b4) {
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
      unsigned int ub[4], ua[4];
      bool in_mx[4], in_mn[4], any = false;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool ok = av[e] <= bv[e];
        ub[e] = __float_as_uint(fabsf(bv[e])); ua[e] = __float_as_uint(fabsf(av[e]));
        const unsigned int db = ub[e] >> 20, da = ua[e] >> 20;
        in_mx[e] = listed_mx && ok && db == bin_mx;
        in_mn[e] = listed_mn && ok && da == bin_mn;
        any |= in_mx[e] | in_mn[e];
        if (ok && db > bin_mx) above_mx = min(above_mx, ub[e]);
        if (ok && da > bin_mn) above_mn = min(above_mn, ua[e]);
      }
      // members of the two chosen bins are rare (a 2048-bin first digit): one vote per quad decides whether anybody appends
      if (__any_sync(0xffffffffu, any)) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (listed_mx) append(0, in_mx[e], ub[e]);
          if (listed_mn) append(1, in_mn[e], ua[e]);
        }
      }
    });
    above_mx = __reduce_min_sync(0xffffffffu, above_mx);
    above_mn = __reduce_min_sync(0xffffffffu, above_mn);
    if (lane == 0) {
      if (above_mx != 0xFFFFFFFFu) atomicMin(&s_minabove[0], above_mx);
      if (above_mn != 0xFFFFFFFFu) atomicMin(&s_minabove[1], above_mn);
    }
    __syncthreads();
    // members of side sd: from its list, or (crowded bin) straight from L2 with the membership test
    auto for_members = [&](int sd, auto&& g) {
      if (sd == 0 ? listed_mx : listed_mn) {
        const unsigned int n = s_nlist[sd];
        for (unsigned int i = tid; i < n; i += kSelThreads) g(list[sd][i]);
      } else {
        const unsigned int bin = sd == 0 ? bin_mx : bin_mn;
        sweep_pairs(tmin, tmax, n_slots, [&](float a, float b) {
          const unsigned int u = __float_as_uint(fabsf(sd == 0 ? b : a));
          if (a <= b && (u >> 20) == bin) g(u);
        });
      }
    };
    OBS_STAMP(5);
    // ---- (c) the remaining 20 bits, 4 per pass, counted in registers ----
    unsigned int mask = 0xFFFu << 20;
    auto pick_digit = [&](int sh) {   // after the per-warp digit counts are in part[][][]: one warp per side picks
      __syncthreads();
      if (warp < 2) {  // lanes 0..15 own one digit each
        unsigned int cnt = 0;
        if (lane < 16) {
#pragma unroll 8
          for (int w = 0; w < 32; ++w) cnt += part[warp][(w + lane) & 31][lane];
        }
        unsigned int incl = cnt;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
          const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        const unsigned int krem = s_krem[warp];
        const unsigned int vote = __ballot_sync(0xffffffffu, lane < 16 && incl > krem);
        const int d = vote ? __ffs(vote) - 1 : 15;
        const unsigned int below = __shfl_sync(0xffffffffu, incl - cnt, d);
        __syncwarp();
        if (lane == 0) { s_prefix[warp] |= (unsigned int)d << sh; s_krem[warp] = krem - below; }
      }
      __syncthreads();
    };
    // Listed sides (the normal case): ONE warp per side resolves the remaining 20 bits and rank lo + 1 on its own -- list in
    // shared memory, digit counts in packed 16-bit register fields, warp-wide REDUX, the pick in registers -- with no block
    // barrier at all.  Thirty-two warps taking turns at two barriers, a cross-lane reduction and a one-warp pick per pass
    // cost 2.2 us per pass whatever the list length; the cross-lane units of one SM are the bottleneck, not the ALUs.
    const bool reg_path = listed_mx && listed_mn;
    if (reg_path) {
      if (warp < 2) {
        const int sd = warp;
        const unsigned int n = s_nlist[sd];
        const unsigned int* lst = list[sd];
        unsigned int prefix = s_prefix[sd], krem = s_krem[sd];
#pragma unroll 1
        for (int sh = 16; sh >= 0; sh -= 4) {
          unsigned int q[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
          for (unsigned int i = lane; i < n; i += 32) {
            const unsigned int u = lst[i];
            const unsigned int d = (u >> sh) & 15u;
            const unsigned int inc = ((u & mask) == prefix) ? (1u << ((d & 1u) << 4)) : 0u;
#pragma unroll
            for (int w = 0; w < 8; ++w) q[w] += ((d >> 1) == (unsigned int)w) ? inc : 0u;
          }
#pragma unroll
          for (int w = 0; w < 8; ++w) q[w] = __reduce_add_sync(0xffffffffu, q[w]);   // <= 8192 members: no field overflows
          unsigned int below = 0, dsel = 15, done = 0;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const unsigned int c = (q[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu;
            const unsigned int hit = (!done && below + c > krem) ? 1u : 0u;
            dsel = hit ? (unsigned int)j : dsel;
            done |= hit;
            below += done ? 0u : c;
          }
          prefix |= dsel << sh;
          krem -= below;
          mask |= 0xFu << sh;
        }
        // rank lo + 1 inside the list: members <= a, smallest member above a
        unsigned int cle = 0, nxt = 0xFFFFFFFFu;
        for (unsigned int i = lane; i < n; i += 32) {
          const unsigned int u = lst[i];
          if (u <= prefix) ++cle; else nxt = min(nxt, u);
        }
        cle = __reduce_add_sync(0xffffffffu, cle);
        nxt = __reduce_min_sync(0xffffffffu, nxt);
        __syncwarp();   // every lane has read s_prefix / s_krem (above) before lane 0 overwrites them
        if (lane == 0) { s_prefix[sd] = prefix; s_krem[sd] = krem; s_cntle[sd] = cle; s_next[sd] = nxt; }
      }
    } else {
#pragma unroll 1
    for (int sh = 16; sh >= 0; sh -= 4) {
#pragma unroll
      for (int sd = 0; sd < 2; ++sd) {
        Packed16 cc;
        cc.clear();
        const unsigned int prefix = s_prefix[sd];
        for_members(sd, [&](unsigned int u) { cc.add((u >> sh) & 15u, (u & mask) == prefix); });
        cc.warp_sum();
        unsigned int mine = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) mine = (lane == j) ? cc.get(j) : mine;
        if (lane < 16) part[sd][warp][lane] = mine;
      }
      pick_digit(sh);
      mask |= 0xFu << sh;
    }
    }
    OBS_STAMP(6);
    // ---- (c') rank lo + 1 ----
    if (reg_path) {
      // done by the side's warp above
    } else {
#pragma unroll
    for (int sd = 0; sd < 2; ++sd) {
      const unsigned int a = s_prefix[sd];
      unsigned int cle = 0, nxt = 0xFFFFFFFFu;
      for_members(sd, [&](unsigned int u) { if (u <= a) ++cle; else nxt = min(nxt, u); });
      cle = __reduce_add_sync(0xffffffffu, cle);
      nxt = __reduce_min_sync(0xffffffffu, nxt);
      if (lane == 0) {
        if (cle) atomicAdd(&s_cntle[sd], cle);
        if (nxt != 0xFFFFFFFFu) atomicMin(&s_next[sd], nxt);
      }
    }
    }
    __syncthreads();
    // the number of ALL valid tokens <= a is (tokens in lower first-digit bins) + (members <= a); rank lo + 1 equals a
    // iff that count reaches lo + 2, else it is the next larger magnitude
    if (warp < 2) {
      unsigned int below = 0;
      const unsigned int bin = s_prefix[warp] >> 20;
      for (unsigned int i = lane; i < bin; i += 32) below += h0[warp * kSelBins0 + i];
      below = __reduce_add_sync(0xffffffffu, below);
      if (lane == 0) {
        const unsigned int a = s_prefix[warp];
        unsigned int b = a;
        if ((unsigned int)lo + 1u < (unsigned int)T && below + s_cntle[warp] < (unsigned int)lo + 2u)
          b = (s_next[warp] != 0xFFFFFFFFu) ? s_next[warp] : ((s_minabove[warp] != 0xFFFFFFFFu) ? s_minabove[warp] : a);
        s_next[warp] = b;   // reuse: the value of rank lo + 1
      }
    }
    __syncthreads();
    const float a_mx = __uint_as_float(s_prefix[0]), b_mx = __uint_as_float(s_next[0]);
    const float a_mn = __uint_as_float(s_prefix[1]), b_mn = __uint_as_float(s_next[1]);
    thr_up = quantile_from_pair(a_mx, need_pair ? b_mx : a_mx, frank, lo);
    thr_lo = -quantile_from_pair(a_mn, need_pair ? b_mn : a_mn, frank, lo);
  }
  // ---- (d) clip + aminmax over the kept tokens (observer.py:66-69, 227) ----
  OBS_STAMP(7);
  float lower = INFINITY, upper = -INFINITY;
  if (T > 0)
    sweep_quads<4>(tmin, tmax, n_slots, [&](const float4 a4, const float4 b4) {
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (av[e] <= bv[e]) {
          if (bv[e] <= thr_up) upper = fmaxf(upper, bv[e]);
          if (av[e] >= thr_lo) lower = fminf(lower, av[e]);
        }
      }
    });
  // (one REDUX per value instead of five shuffles: no NaNs here, the vectors come out of fminf / fmaxf)
  lower = ord2f(__reduce_min_sync(0xffffffffu, f2ord(lower)));
  upper = ord2f(__reduce_max_sync(0xffffffffu, f2ord(upper)));
  if (lane == 0) { red[0][warp] = lower; red[1][warp] = upper; }
  __syncthreads();
  if (warp == 0) {
    lower = warp_min(red[0][lane]);
    upper = warp_max(red[1][lane]);
    if (lane == 0) {
      cur[0] = lower;
      cur[1] = upper;
      stat_epilogue(epi, lower, upper);
      OBS_STAMP(8);
    }
  }
}

__global__ void __launch_bounds__(kSelThreads, 1)
prune_select_tail_kernel(const float* __restrict__ tmin, const float* __restrict__ tmax, int64_t n_slots,
                         const int32_t* __restrict__ n_valid, float percentile, const float* __restrict__ percentile_dev,
                         unsigned int* __restrict__ ghist0, float* __restrict__ cur, osq_stat_epilogue_t epi, int release_early) {
  prune_select_tail_body(tmin, tmax, n_slots, n_valid, percentile, percentile_dev, ghist0, cur, epi, false,
                         reinterpret_cast<long long*>(ghist0 + 2 * kSelBins0), release_early != 0);
}

// The select on CACHED per-token vectors, many problems per launch (one CTA each): token-wise clipping re-calibrates every
// observer for every candidate ratio, but its calibration forwards run with activation fake-quant off (token_wise_clipping.py:12-19),
// so the per-token extrema of every (observer, batch) are the same for all ratios -- only the quantile moves.  The vectors and
// their first-digit tables are recorded once (osq_token_minmax_hist_f32); each ratio then costs one launch of this kernel plus one
// replay launch instead of a model forward per batch.
__global__ void __launch_bounds__(kSelThreads, 1)
prune_select_cached_kernel(const osq_select_problem_t* __restrict__ problems, float percentile, const float* __restrict__ percentile_dev) {
  const osq_select_problem_t pr = problems[blockIdx.x];
  osq_stat_epilogue_t none;
  none.mode = 0; none.cnt = 0; none.state_min = nullptr; none.state_max = nullptr; none.scale_out = nullptr; none.zp_out = nullptr;
  none.zp_out_is_int32 = 0; none.qmin = 0; none.qmax = 1; none.symmetric = 0;
  prune_select_tail_body(pr.tmin, pr.tmax, pr.n_slots, pr.n_valid, percentile, percentile_dev, const_cast<unsigned int*>(pr.hist0), pr.cur, none,
                         true, nullptr);
}

// ---------------------------------------------------------------------------------------------
// AvgQuantileObserver (observer.py:253-282): histogram of |x| over the valid tokens in `bins` equal bins of [0, R],
// R = max(-min, max); first bin whose cumulative count reaches threshold * numel; clip; running average.
// Second launch of the observer (the first is K3, which leaves (min, max) in cur[]).  Binning follows ATen's CPU histc
// (probed on torch 2.11: pos = int((|x| * bins) / R) in fp32, the right edge falls into the last bin); the cumulative
// scan and its comparison run in fp32 like the reference's Python loop over an fp32 histogram.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxQuantileBins = 8192;

__device__ __forceinline__ void hist_abs_run(const float* __restrict__ p, int64_t len, int64_t stride, int lane, float fb, float R,
                                             int bins, unsigned int* hist) {
  for (int64_t j = lane; j < len; j += 32) {
    const float a = fabsf(__ldg(p + j * stride));
    if (!(a <= R)) continue;  // NaN (histc skips it)
    int pos = (int)__fdiv_rn(__fmul_rn(a, fb), R);
    if (pos >= bins) pos = bins - 1;
    atomicAdd(&hist[pos], 1u);
  }
}

__global__ void __launch_bounds__(kObsThreads)
abs_hist_kernel(const float* __restrict__ x, osq_tokens_t tk, const int64_t* __restrict__ lens, int n_lens, int bins,
                double threshold, unsigned int* __restrict__ ghist, float* __restrict__ cur, osq_stat_epilogue_t epi,
                void* wsp) {
  extern __shared__ unsigned int shist[];
  const int lane = threadIdx.x & 31;
  const float mn0 = cur[0], mx0 = cur[1];
  const float R = fmaxf(-mn0, mx0);
  const float fb = (float)bins;
  for (int i = threadIdx.x; i < bins; i += blockDim.x) shist[i] = 0;
  __syncthreads();
  const int64_t warp_global = (int64_t)blockIdx.x * kObsWarps + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * kObsWarps;
  const int64_t n_seg = tk.B * tk.S * tk.F1;
  for (int64_t seg = warp_global; seg < n_seg; seg += n_warps) {
    const int64_t f1 = seg % tk.F1;
    const int64_t bs = seg / tk.F1;
    const int64_t s_ = bs % tk.S, b = bs / tk.S;
    if (!token_valid(lens, n_lens, b, s_)) continue;
    hist_abs_run(x + b * tk.sb + s_ * tk.ss + f1 * tk.sf1, tk.F2, tk.sf2, lane, fb, R, bins, shist);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bins; i += blockDim.x) {
    const unsigned int c = shist[i];
    if (c) atomicAdd(&ghist[i], c);
  }
  Partials ws = carve(wsp);
  if (!last_cta(ws.ticket)) return;
  // ---- last CTA: the reference's sequential scan, then clip + running average + qparams ----
  for (int i = threadIdx.x; i < bins; i += blockDim.x) { shist[i] = __ldcg(&ghist[i]); ghist[i] = 0; }  // read + re-arm
  __syncthreads();
  if (threadIdx.x == 0) {
    *ws.ticket = 0;
    int64_t tokens = 0;
    if (lens == nullptr) tokens = tk.B * tk.S;
    else
      for (int64_t b = 0; b < tk.B && b < n_lens; ++b) {
        const int64_t l = lens[b];
        tokens += l < 0 ? 0 : (l > tk.S ? tk.S : l);
      }
    const double numel = (double)(tokens * tk.F1 * tk.F2);
    const float need = (float)(threshold * numel);   // a fp32 tensor compared with a Python float compares in fp32
    float total = 0.f, clip = R;
    for (int i = 0; i < bins; ++i) {
      const float c = (float)shist[i];
      if (__fadd_rn(total, c) >= need) {
        clip = __fmul_rn((float)i + 0.5f, __fdiv_rn(R, fb));
        break;
      }
      total = __fadd_rn(total, c);
    }
    const float lo = (-clip > mn0) ? -clip : mn0;   // Python max(min_val_cur, -clip_value)
    const float hi = (clip < mx0) ? clip : mx0;     // Python min(max_val_cur, clip_value)
    cur[0] = lo;
    cur[1] = hi;
    stat_epilogue(epi, lo, hi);
  }
}

// ---------------------------------------------------------------------------------------------
// Rank-sharded calibration (dist.py): after the one all-reduce of the slot table [n_obs, n_batches, 2] every rank
// replays observer.py:194-202  m <- (m*cnt + cur)/(cnt+1)  in batch order for ALL observers in one launch (one thread per
// observer) and refreshes every owning quantizer's (scale, zero_point) through a pointer table -- no host round trip.
// ---------------------------------------------------------------------------------------------
// `peers` != nullptr: slot (observer, batch b) is read from the table of the rank that processed batch b (b mod world) through
// its peer-mapped pointer -- plain loads over NVLink, no collective library in the path (dist.py: symmetric-memory slot table)
__global__ void __launch_bounds__(128)
replay_average_kernel(const float* __restrict__ table, const float* const* __restrict__ peers, int world, int n_obs, int n_batches,
                      int cnt0, const osq_replay_target_t* __restrict__ tgt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_obs) return;
  osq_replay_target_t t = tgt[i];
  osq_stat_epilogue_t e;
  e.mode = 1;
  e.state_min = t.state_min;
  e.state_max = t.state_max;
  e.scale_out = nullptr;   // qparams once, after the last batch
  e.zp_out = nullptr;
  e.zp_out_is_int32 = t.zp_out_is_int32;
  e.qmin = t.qmin; e.qmax = t.qmax; e.symmetric = t.symmetric;
  for (int b = 0; b < n_batches; ++b) {
    e.cnt = cnt0 + b;
    if (b == n_batches - 1) { e.scale_out = t.scale_out; e.zp_out = t.zp_out; }
    const float* src = peers != nullptr ? peers[b % world] : table;
    const float2 v = __ldcv(reinterpret_cast<const float2*>(src + ((int64_t)i * n_batches + b) * 2));   // never a cached copy
    stat_epilogue(e, v.x, v.y);
  }
}

// ---------------------------------------------------------------------------------------------
// Exchange + replay as ONE launch over NVLink peer memory (no collective library, no host-side barrier).
//
// Every rank owns a symmetric-memory region, mapped by all ranks:  float slots[2][n_obs * n_batches * 2]  (two generations,
// selected by the parity of the pass number), then  uint32 flags[world].  The launch
//   1. publishes this rank's slots (batches b with b mod world == rank) from its private slot table into its own region,
//      generation `pass & 1`                                                     (plain stores + __threadfence_system)
//   2. signals every peer:  peer.flags[rank] = pass                               (one remote release-store per peer)
//   3. waits until flags[r] >= pass for every r                                   (acquire loads of LOCAL memory)
//   4. replays observer.py:194-202 in batch order for every observer, loading slot (i, b) from rank b mod world's region.
// Two generations make a trailing barrier unnecessary: generation g is rewritten in pass p + 2, which a rank can only reach
// after every peer has signalled pass p + 1, i.e. after every peer has finished reading pass p.  `pass` lives in device memory
// (pass_counter) and is advanced by the kernel, so a CUDA graph can replay the launch.  A peer that never signals (a rank
// died) is given `timeout_ns`, then the launch records the failure in err_flag and returns without touching the targets.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) { uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

__global__ void __launch_bounds__(256)
replay_exchange_kernel(const float* __restrict__ local_table, float* const* __restrict__ regions, int rank, int world, int n_obs, int n_batches,
                       int cnt0, const osq_replay_target_t* __restrict__ tgt, uint32_t* __restrict__ pass_counter, uint32_t* __restrict__ err_flag,
                       long long timeout_ns) {
  __shared__ int s_fail;
  const int tid = threadIdx.x;
  const uint32_t pass = *pass_counter + 1u;
  const int n_f = n_obs * n_batches * 2;
  float* mine = regions[rank] + (size_t)(pass & 1u) * n_f;
  for (int j = tid; j < n_obs * n_batches; j += blockDim.x)
    if ((j % n_batches) % world == rank) {
      mine[2 * j] = local_table[2 * j];
      mine[2 * j + 1] = local_table[2 * j + 1];
    }
  if (tid == 0) s_fail = 0;
  __threadfence_system();
  __syncthreads();
  if (tid < world) {
    st_release_sys_u32(reinterpret_cast<uint32_t*>(regions[tid] + 2 * (size_t)n_f) + rank, pass);
    const uint32_t* my_flag = reinterpret_cast<const uint32_t*>(regions[rank] + 2 * (size_t)n_f) + tid;
    const long long t0 = obs_gtimer();
    while ((int32_t)(ld_acquire_sys_u32(my_flag) - pass) < 0) {
      __nanosleep(100);
      if (obs_gtimer() - t0 > timeout_ns) { s_fail = 1; break; }
    }
  }
  __syncthreads();
  if (s_fail) {
    if (tid == 0) { *err_flag = pass; *pass_counter = pass; }
    return;
  }
  for (int i = tid; i < n_obs; i += blockDim.x) {
    osq_replay_target_t t = tgt[i];
    osq_stat_epilogue_t e;
    e.mode = 1;
    e.state_min = t.state_min;
    e.state_max = t.state_max;
    e.scale_out = nullptr;   // qparams once, after the last batch
    e.zp_out = nullptr;
    e.zp_out_is_int32 = t.zp_out_is_int32;
    e.qmin = t.qmin; e.qmax = t.qmax; e.symmetric = t.symmetric;
    for (int b = 0; b < n_batches; ++b) {
      e.cnt = cnt0 + b;
      if (b == n_batches - 1) { e.scale_out = t.scale_out; e.zp_out = t.zp_out; }
      const float* src = regions[b % world] + (size_t)(pass & 1u) * n_f;
      const float2 v = __ldcv(reinterpret_cast<const float2*>(src + ((int64_t)i * n_batches + b) * 2));   // never a cached copy
      stat_epilogue(e, v.x, v.y);
    }
  }
  if (tid == 0) *pass_counter = pass;
}

// ---------------------------------------------------------------------------------------------
// per-row min/max + running extrema + per-row qparams (weights, MinMaxObserver ch_axis=0)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rowwise_minmax_qparams_kernel(const float* __restrict__ w, int64_t rows, int64_t cols, int first,
                              float* __restrict__ state_min, float* __restrict__ state_max,
                              float* __restrict__ scale_out, int32_t* __restrict__ zp_out, int qmin,
                              int qmax, int symmetric) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = warp_global; r < rows; r += n_warps) {
    float mn = INFINITY, mx = -INFINITY;
    warp_scan_segment(w + r * cols, cols, lane, mn, mx);
    mn = warp_min(mn);
    mx = warp_max(mx);
    if (lane == 0) {
      if (!first) {
        mn = fminf(mn, state_min[r]);
        mx = fmaxf(mx, state_max[r]);
      }
      state_min[r] = mn;
      state_max[r] = mx;
      if (scale_out != nullptr) {
        float zp;
        scale_out[r] = calc_qparams(mn, mx, qmin, qmax, symmetric, zp);
        if (zp_out != nullptr) zp_out[r] = (int32_t)zp;
      }
    }
  }
}

static int reduction_grid(int64_t units_of_work) {
  int sms = sm_count();
  if (sms <= 0) return -1;
  int64_t g = (units_of_work + kObsWarps - 1) / kObsWarps;
  static int per_sm = -1;          // experiment knob; default 4 CTAs of 512 threads = 64 warps / SM
  if (per_sm < 0) { const char* e = getenv("OSQ_OBS_CTAS_PER_SM"); per_sm = e ? atoi(e) : 4; if (per_sm < 1 || per_sm > 4) per_sm = 4; }
  int64_t cap = (int64_t)sms * per_sm;
  if (cap > kMaxPartialBlocks) cap = kMaxPartialBlocks;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

static int check_tokens(const osq_tokens_t* t, const char* who) {
  OSQ_CHECK_ARG(t != nullptr, "%s: null token geometry", who);
  OSQ_CHECK_ARG(t->B >= 0 && t->S >= 0 && t->F1 >= 0 && t->F2 >= 0, "%s: negative size", who);
  return OSQ_OK;
}

}  // namespace osq

extern "C" {

int osq_minmax_masked_f32(const float* x, const osq_tokens_t* tok, const int64_t* lens, int n_lens,
                          float* cur_minmax, const osq_stat_epilogue_t* epi, void* workspace,
                          void* stream) {
  using namespace osq;
  if (int rc = check_tokens(tok, "osq_minmax_masked_f32")) return rc;
  OSQ_CHECK_ARG(x && cur_minmax && epi && workspace, "osq_minmax_masked_f32: null pointer");
  int grid = reduction_grid(tok->B * tok->S * tok->F1);
  if (grid < 0) { set_error("no CUDA device"); return OSQ_ECUDA; }
  minmax_masked_kernel<<<grid, kObsThreads, 0, (cudaStream_t)stream>>>(x, *tok, lens, n_lens, cur_minmax, *epi, workspace);
  OSQ_LAUNCH_CHECK();
  return OSQ_OK;
}

int osq_minmax_flat_f32(const float* x, int64_t n, float* cur_minmax, const osq_stat_epilogue_t* epi,
                        void* workspace, void* stream) {
  using namespace osq;
  OSQ_CHECK_ARG(x && cur_minmax && epi && workspace && n > 0, "osq_minmax_flat_f32: bad argument");
  int grid = reduction_grid((n + 4095) / 4096);
  if (grid < 0) { set_error("no CUDA device"); return OSQ_ECUDA; }
  minmax_flat_kernel<<<grid, kObsThreads, 0, (cudaStream_t)stream>>>(x, n, cur_minmax, *epi, workspace);
  OSQ_LAUNCH_CHECK();
  return OSQ_OK;
}

int osq_token_minmax_f32(const float* x, const osq_tokens_t* tok, const int64_t* lens, int n_lens,
                         float* tmin, float* tmax, int32_t* n_valid, void* stream) {
  using namespace osq;
  if (int rc = check_tokens(tok, "osq_token_minmax_f32")) return rc;
  OSQ_CHECK_ARG(x && tmin && tmax && n_valid, "osq_token_minmax_f32: null pointer");
  OSQ_CHECK_ARG(tok->B * tok->S < (int64_t)INT32_MAX, "osq_token_minmax_f32: too many tokens");
  int grid = reduction_grid(tok->B * tok->S);
  if (grid < 0) { set_error("no CUDA device"); return OSQ_ECUDA; }
  token_minmax_kernel<<<grid, kObsThreads, 0, (cudaStream_t)stream>>>(x, *tok, lens, n_lens, tmin, tmax, n_valid, nullptr, 0);
  OSQ_LAUNCH_CHECK();
  return OSQ_OK;
}

int osq_token_minmax_hist_f32(const float* x, const osq_tokens_t* tok, const int64_t* lens, int n_lens, float* tmin, float* tmax,
                              int32_t* n_valid, uint32_t* hist0, void* stream) {
  using namespace osq;
  if (int rc = check_tokens(tok, "osq_token_minmax_hist_f32")) return rc;
  OSQ_CHECK_ARG(x && tmin && tmax && n_valid && hist0, "osq_token_minmax_hist_f32: null pointer");
  OSQ_CHECK_ARG(tok->B * tok->S < kSelMaxTokens, "osq_token_minmax_hist_f32: too many tokens for the cached select");
  int grid = reduction_grid(tok->B * tok->S);
  if (grid < 0) { set_error("no CUDA device"); return OSQ_ECUDA; }
  token_minmax_kernel<<<grid, kObsThreads, 0, (cudaStream_t)stream>>>(x, *tok, lens, n_lens, tmin, tmax, n_valid, hist0, 0);
  OSQ_LAUNCH_CHECK();
  return OSQ_OK;
}

int osq_prune_select_cached_f32(const osq_select_problem_t* problems, int n_problems, float percentile, const float* percentile_dev,
                                void* stream) {
  using namespace osq;
  OSQ_CHECK_ARG(problems != nullptr && n_problems >= 0, "osq_prune_select_cached_f32: bad argument");
  OSQ_CHECK_ARG(percentile >= 0.f && percentile <= 1.f, "osq_prune_select_cached_f32: percentile outside [0,1]");
  if (n_problems == 0) return OSQ_OK;
  constexpr size_t kSelSmem = (size_t)(2 * kSelBins0 + 2 * kSelCap) * 4;
  static bool attr_set[64] = {false};
  int dev = 0;
  OSQ_CUDA(cudaGetDevice(&dev));
  if (!attr_set[dev & 63]) {
    OSQ_CUDA(cudaFuncSetAttribute(prune_select_cached_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSelSmem));
    attr_set[dev & 63] = true;
  }
  prune_select_cached_kernel<<<n_problems, kSelThreads, kSelSmem, (cudaStream_t)stream>>>(problems, percentile, percentile_dev);
  OSQ_LAUNCH_CHECK();
  return OSQ_OK;
}

int osq_prune_select_f32(const float* tmin, const float* tmax, const float* abs_tmin_sorted,
                         const float* abs_tmax_sorted, int64_t n_slots, const int32_t* n_valid,
                         float percentile, float* cur_minmax, const osq_stat_epilogue_t* epi,
                         void* workspace, void* stream) {
  using namespace osq;
  (void)workspace;
  OSQ_CHECK_ARG(tmin && tmax && abs_tmin_sorted && abs_tmax_sorted && n_valid && cur_minmax && epi,
                "osq_prune_select_f32: null pointer");
  OSQ_CHECK_ARG(n_slots > 0, "osq_prune_select_f32: n_slots <= 0");
  OSQ_CHECK_ARG(percentile >= 0.f && percentile <= 1.f, "osq_prune_select_f32: percentile outside [0,1]");
  prune_select_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(tmin, tmax, abs_tmin_sorted, abs_tmax_sorted, n_slots,
                                                           n_valid, percentile, cur_minmax, *epi);
  OSQ_LAUNCH_CHECK();
  return OSQ_OK;
}

int osq_prune_select_unsorted_f32(const float* tmin, const float* tmax, int64_t n_slots, const int32_t* n_valid,
                                  float percentile, float* cur_minmax, const osq_stat_epilogue_t* epi,
                                  void* workspace, void* stream) {
  using namespace osq;
  OSQ_CHECK_ARG(tmin && tmax && n_valid && cur_minmax && epi && workspace, "osq_prune_select_unsorted_f32: null pointer");
  OSQ_CHECK_ARG(n_slots > 0, "osq_prune_select_unsorted_f32: n_slots <= 0");
  OSQ_CHECK_ARG(percentile >= 0.f && percentile <= 1.f, "osq_prune_select_unsorted_f32: percentile outside [0,1]");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_slots <= kSelectSingleCtaMax) {
    prune_select_unsorted_kernel<<<1, 1024, 0, st>>>(tmin, tmax, n_slots, n_valid, percentile, cur_minmax, *epi);
    OSQ_LAUNCH_CHECK();
    return OSQ_OK;
  }
  int sms = sm_count();
  if (sms <= 0) { set_error("no CUDA device"); return OSQ_ECUDA; }
  int64_t g = (n_slots + 8191) / 8192;  // 8 slots per thread and side
  if (g > sms) g = sms;
  SelectWs* sw = reinterpret_cast<SelectWs*>(static_cast<char*>(workspace) + kSelectWsOffset);
  for (int pass = 3; pass >= 0; --pass)
    select_pass_kernel<<<(int)g, 1024, 0, st>>>(tmin, tmax, n_slots, n_valid, percentile, pass, sw);
  select_final_kernel<<<(int)g, 1024, 0, st>>>(tmin, tmax, n_slots, n_valid, percentile, sw);
  int64_t g2 = (n_slots + kObsThreads * 4 - 1) / (kObsThreads * 4);
  if (g2 > sms) g2 = sms;
  prune_apply_kernel<<<(int)g2, kObsThreads, 0, st>>>(tmin, tmax, n_slots, n_valid, sw, cur_minmax, *epi, workspace);
  OSQ_LAUNCH_CHECK();
  return OSQ_OK;
}


int osq_prune_observe_f32(const float* x, const osq_tokens_t* tok, const int64_t* lens, int n_lens, float percentile,
                          const float* percentile_dev, float* tmin, float* tmax, int32_t* n_valid, float* cur_minmax,
                          const osq_stat_epilogue_t* epi, void* workspace, void* stream) {
  using namespace osq;
  if (int rc = check_tokens(tok, "osq_prune_observe_f32")) return rc;
  OSQ_CHECK_ARG(x && tmin && tmax && n_valid && cur_minmax && epi && workspace, "osq_prune_observe_f32: null pointer");
  OSQ_CHECK_ARG(percentile >= 0.f && percentile <= 1.f, "osq_prune_observe_f32: percentile outside [0,1]");
  const int64_t n_slots = tok->B * tok->S;
  OSQ_CHECK_ARG(n_slots > 0 && n_slots < (int64_t)INT32_MAX, "osq_prune_observe_f32: token count out of range");
  if (n_slots >= kSelMaxTokens) {  // beyond the packed 16-bit counters of the tail: the multi-launch select over L2
    OSQ_CHECK_ARG(percentile_dev == nullptr, "osq_prune_observe_f32: a device-resident percentile needs fewer than 2^21 tokens");
    if (int rc = osq_token_minmax_f32(x, tok, lens, n_lens, tmin, tmax, n_valid, stream)) return rc;
    return osq_prune_select_unsorted_f32(tmin, tmax, n_slots, n_valid, percentile, cur_minmax, epi, workspace, stream);
  }
  int grid = reduction_grid(n_slots);
  if (grid < 0) { set_error("no CUDA device"); return OSQ_ECUDA; }
  cudaStream_t st = (cudaStream_t)stream;
  constexpr size_t kSelSmem = (size_t)(2 * kSelBins0 + 2 * kSelCap) * 4;
  static bool attr_set[64] = {false};
  int dev = 0;
  OSQ_CUDA(cudaGetDevice(&dev));
  if (!attr_set[dev & 63]) {
    OSQ_CUDA(cudaFuncSetAttribute(prune_select_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSelSmem));
    // both kernels of the pair ask for the same shared-memory carve-out: an SM whose L1 / shared split has to change
    // between two kernels drains first, which costs more than the tail kernel itself
    static int carve = -2;
    if (carve == -2) { const char* e = getenv("OSQ_OBS_CARVEOUT"); carve = e ? atoi(e) : 50; }
    if (carve >= 0) {
      OSQ_CUDA(cudaFuncSetAttribute(prune_select_tail_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
      OSQ_CUDA(cudaFuncSetAttribute(token_minmax_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
    }
    attr_set[dev & 63] = true;
  }
  unsigned int* hist0 = reinterpret_cast<unsigned int*>(static_cast<char*>(workspace) + kSelectHist0Offset);
  token_minmax_kernel<<<grid, kObsThreads, 0, st>>>(x, *tok, lens, n_lens, tmin, tmax, n_valid, hist0, 0);
  OSQ_LAUNCH_CHECK();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(1, 1, 1);
  cfg.blockDim = dim3(kSelThreads, 1, 1);
  cfg.dynamicSmemBytes = kSelSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // park behind the per-token pass (griddepcontrol.wait)
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  OSQ_CUDA(cudaLaunchKernelEx(&cfg, prune_select_tail_kernel, (const float*)tmin, (const float*)tmax, n_slots,
                              (const int32_t*)n_valid, percentile, percentile_dev, hist0, cur_minmax, *epi, 0));
  return OSQ_OK;
}

int osq_prune_observe_many_f32(const float* const* xs, int n, const osq_tokens_t* tok, const int64_t* lens, int n_lens, float percentile,
                               const float* percentile_dev, float* tmin2, float* tmax2, int32_t* n_valid2, float* const* curs,
                               const osq_stat_epilogue_t* epis, void* workspace, void* stream) {
  using namespace osq;
  if (int rc = check_tokens(tok, "osq_prune_observe_many_f32")) return rc;
  OSQ_CHECK_ARG(xs && n >= 1 && tmin2 && tmax2 && n_valid2 && curs && epis && workspace, "osq_prune_observe_many_f32: bad argument");
  OSQ_CHECK_ARG(percentile >= 0.f && percentile <= 1.f, "osq_prune_observe_many_f32: percentile outside [0,1]");
  const int64_t n_slots = tok->B * tok->S;
  OSQ_CHECK_ARG(n_slots > 0 && n_slots < kSelMaxTokens, "osq_prune_observe_many_f32: token count out of range (use osq_prune_observe_f32)");
  for (int i = 0; i < n; ++i) OSQ_CHECK_ARG(xs[i] && curs[i], "osq_prune_observe_many_f32: null activation / output pointer");
  // make sure the shared-memory attributes of the pair are set (first use of the device)
  if (n == 1) return osq_prune_observe_f32(xs[0], tok, lens, n_lens, percentile, percentile_dev, tmin2, tmax2, n_valid2, curs[0], &epis[0], workspace, stream);
  {
    static bool warmed[64] = {false};
    int dev = 0;
    OSQ_CUDA(cudaGetDevice(&dev));
    if (!warmed[dev & 63]) {   // the single-call path sets the function attributes; run it for batch 0 and continue with the rest
      if (int rc = osq_prune_observe_f32(xs[0], tok, lens, n_lens, percentile, percentile_dev, tmin2, tmax2, n_valid2, curs[0], &epis[0], workspace, stream)) return rc;
      warmed[dev & 63] = true;
      return osq_prune_observe_many_f32(xs + 1, n - 1, tok, lens, n_lens, percentile, percentile_dev, tmin2, tmax2, n_valid2, curs + 1, epis + 1, workspace, stream);
    }
  }
  int grid = reduction_grid(n_slots);
  if (grid < 0) { set_error("no CUDA device"); return OSQ_ECUDA; }
  cudaStream_t st = (cudaStream_t)stream;
  constexpr size_t kSelSmem = (size_t)(2 * kSelBins0 + 2 * kSelCap) * 4;
  const int64_t n4 = (n_slots + 3) & ~(int64_t)3;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  for (int i = 0; i < n; ++i) {
    const int par = i & 1;   // vectors and first-digit tables alternate: batch i + 1's per-token pass never touches what tail i reads
    float* tmin = tmin2 + par * n4;
    float* tmax = tmax2 + par * n4;
    int32_t* n_valid = n_valid2 + par;
    unsigned int* hist0 = reinterpret_cast<unsigned int*>(static_cast<char*>(workspace) + (par ? kSelectHist1Offset : kSelectHist0Offset));
    cudaLaunchConfig_t tcfg = {};
    tcfg.gridDim = dim3((unsigned)grid, 1, 1);
    tcfg.blockDim = dim3(kObsThreads, 1, 1);
    tcfg.stream = st;
    tcfg.attrs = attr;
    tcfg.numAttrs = i > 0 ? 1 : 0;   // batch 0 is an ordinary launch: whatever produced the activations has completed
    OSQ_CUDA(cudaLaunchKernelEx(&tcfg, token_minmax_kernel, xs[i], *tok, lens, n_lens, tmin, tmax, n_valid, hist0, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1, 1, 1);
    cfg.blockDim = dim3(kSelThreads, 1, 1);
    cfg.dynamicSmemBytes = kSelSmem;
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    OSQ_CUDA(cudaLaunchKernelEx(&cfg, prune_select_tail_kernel, (const float*)tmin, (const float*)tmax, n_slots, (const int32_t*)n_valid,
                                percentile, percentile_dev, hist0, curs[i], epis[i], i + 1 < n ? 1 : 0));
  }
  return OSQ_OK;
}

int osq_quantile_observe_f32(const float* x, const osq_tokens_t* tok, const int64_t* lens, int n_lens, int bins,
                             double threshold, uint32_t* hist, float* cur_minmax, const osq_stat_epilogue_t* epi,
                             void* workspace, void* stream) {
  using namespace osq;
  if (int rc = check_tokens(tok, "osq_quantile_observe_f32")) return rc;
  OSQ_CHECK_ARG(x && hist && cur_minmax && epi && workspace, "osq_quantile_observe_f32: null pointer");
  OSQ_CHECK_ARG(bins >= 1 && bins <= kMaxQuantileBins, "osq_quantile_observe_f32: bins must be in [1, 8192]");
  OSQ_CHECK_ARG(threshold >= 0.0 && threshold <= 1.0, "osq_quantile_observe_f32: threshold outside [0,1]");
  int grid = reduction_grid(tok->B * tok->S * tok->F1);
  if (grid < 0) { set_error("no CUDA device"); return OSQ_ECUDA; }
  cudaStream_t st = (cudaStream_t)stream;
  osq_stat_epilogue_t none = *epi;
  none.mode = 0;
  minmax_masked_kernel<<<grid, kObsThreads, 0, st>>>(x, *tok, lens, n_lens, cur_minmax, none, workspace);
  abs_hist_kernel<<<grid, kObsThreads, (size_t)bins * 4, st>>>(x, *tok, lens, n_lens, bins, threshold, hist, cur_minmax, *epi,
                                                              workspace);
  OSQ_LAUNCH_CHECK();
  return OSQ_OK;
}

int osq_replay_average_f32(const float* table, int n_obs, int n_batches, int cnt0, const osq_replay_target_t* targets,
                           void* stream) {
  using namespace osq;
  OSQ_CHECK_ARG(table && targets && n_obs > 0 && n_batches > 0 && cnt0 >= 0, "osq_replay_average_f32: bad argument");
  replay_average_kernel<<<(n_obs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(table, nullptr, 1, n_obs, n_batches, cnt0, targets);
  OSQ_LAUNCH_CHECK();
  return OSQ_OK;
}

int osq_replay_average_peer_f32(const float* const* peer_tables, int world, int n_obs, int n_batches, int cnt0,
                                const osq_replay_target_t* targets, void* stream) {
  using namespace osq;
  OSQ_CHECK_ARG(peer_tables && targets && world >= 1 && n_obs > 0 && n_batches > 0 && cnt0 >= 0, "osq_replay_average_peer_f32: bad argument");
  replay_average_kernel<<<(n_obs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(nullptr, peer_tables, world, n_obs, n_batches, cnt0, targets);
  OSQ_LAUNCH_CHECK();
  return OSQ_OK;
}

int osq_replay_exchange_f32(const float* local_table, float* const* regions, int rank, int world, int n_obs, int n_batches, int cnt0,
                            const osq_replay_target_t* targets, uint32_t* pass_counter, uint32_t* err_flag, void* stream) {
  using namespace osq;
  OSQ_CHECK_ARG(local_table && regions && targets && pass_counter && err_flag, "osq_replay_exchange_f32: null pointer");
  OSQ_CHECK_ARG(world >= 1 && world <= 256 && rank >= 0 && rank < world && n_obs > 0 && n_batches > 0 && cnt0 >= 0,
                "osq_replay_exchange_f32: bad argument");
  static long long timeout_ns = -1;
  if (timeout_ns < 0) { const char* e = getenv("OSQ_EXCHANGE_TIMEOUT_MS"); timeout_ns = (long long)(e ? atoi(e) : 5000) * 1000000ll; }
  replay_exchange_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(local_table, regions, rank, world, n_obs, n_batches, cnt0, targets, pass_counter,
                                                            err_flag, timeout_ns);
  OSQ_LAUNCH_CHECK();
  return OSQ_OK;
}

int osq_rowwise_minmax_qparams_f32(const float* w, int64_t rows, int64_t cols, int first,
                                   float* state_min, float* state_max, float* scale_out,
                                   int32_t* zp_out, int qmin, int qmax, int symmetric, void* stream) {
  using namespace osq;
  OSQ_CHECK_ARG(w && state_min && state_max && rows > 0 && cols > 0, "osq_rowwise_minmax_qparams_f32: bad argument");
  int sms = sm_count();
  if (sms <= 0) { set_error("no CUDA device"); return OSQ_ECUDA; }
  int64_t g = (rows + 7) / 8;
  if (g > (int64_t)sms * 8) g = (int64_t)sms * 8;
  rowwise_minmax_qparams_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(w, rows, cols, first, state_min, state_max,
                                                                       scale_out, zp_out, qmin, qmax, symmetric);
  OSQ_LAUNCH_CHECK();
  return OSQ_OK;
}

}  // extern "C"
