/* TEST INFRASTRUCTURE ONLY -- plain C restatement of the integer/float arithmetic of the hot path.
 * A second, independent pin next to oracle/osq_oracle.py (torch): C `float` division and rintf() are the
 * IEEE operations the reference's CPU path performs.  Checked against tests/golden/*.npz (vectors produced
 * by the unmodified reference) in tests/test_oracle_c.py.  Never linked into the product.
 *
 * Build: oracle/build_c.py  ->  oracle/_build/libosq_oracle.so   (gcc -O2 -ffp-contract=off)
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

/* util_quant.py:4-8 -- forward value of round_ste: (round(t) - t) + t (NaN for +-inf) */
static inline float round_ste_value(float t) {
  volatile float r = rintf(t); /* round-half-to-even under the default rounding mode */
  volatile float d = r - t;
  return d + t;
}

/* util_quant.py:11-15.  y and q (the clamped bin, as float) may be NULL. */
void osqo_fq_per_tensor(const float* x, int64_t n, float scale, float zero_point, float qmin, float qmax,
                        float* y, float* q) {
  for (int64_t i = 0; i < n; ++i) {
    float t = x[i] / scale;
    float v = round_ste_value(t) + zero_point;
    float c = (v != v) ? v : (v < qmin ? qmin : (v > qmax ? qmax : v)); /* torch.clamp keeps NaN */
    if (q) q[i] = c;
    if (y) y[i] = (c - zero_point) * scale;
  }
}

/* util_quant.py:18-26, ch_axis = 0 on a [rows, cols] matrix */
void osqo_fq_per_channel(const float* x, int64_t rows, int64_t cols, const float* scale, const int32_t* zp,
                         float qmin, float qmax, float* y, float* q) {
  for (int64_t r = 0; r < rows; ++r)
    osqo_fq_per_tensor(x + r * cols, cols, scale[r], (float)zp[r], qmin, qmax, y ? y + r * cols : NULL,
                       q ? q + r * cols : NULL);
}

/* util_quant.py:70-71 */
static inline float grad_scale_value(float t, float g) {
  volatile float tg = t * g;
  volatile float d = t - tg;
  return d + tg;
}

/* effective LSQ+ parameters, util_quant.py:49-51 */
void osqo_lsqplus_effective(float scale, float zero_point, float g, float* s_eff, float* z_eff) {
  *s_eff = grad_scale_value(scale, g);
  *z_eff = grad_scale_value(rintf(zero_point), g);
}

/* observer.py:100-119 for one (min, max) pair */
void osqo_qparams(float mn, float mx, int qmin, int qmax, int symmetric, float* scale, float* zp) {
  float lo = mn < 0.f ? mn : 0.f, hi = mx > 0.f ? mx : 0.f;
  float s;
  if (symmetric) {
    if (-lo > hi) hi = -lo;
    s = hi / ((float)(qmax - qmin) / 2.f);
    if (!(s > 1e-8f)) s = 1e-8f;
    *zp = 0.f;
  } else {
    s = (hi - lo) / (float)(qmax - qmin);
    if (!(s > 1e-8f)) s = 1e-8f;
    float z = (float)qmin - rintf(lo / s);
    *zp = z < (float)qmin ? (float)qmin : (z > (float)qmax ? (float)qmax : z);
  }
  *scale = s;
}

/* observer.py:72-98 + :64-65: per-token min/max of a [B, S, F1, F2] strided view, pad tokens skipped.
 * tmin/tmax are compacted (batch-major); returns the number of valid tokens. */
int64_t osqo_token_minmax(const float* x, int64_t B, int64_t S, int64_t F1, int64_t F2, int64_t sb, int64_t ss,
                          int64_t sf1, int64_t sf2, const int64_t* lens, int64_t n_lens, float* tmin, float* tmax) {
  int64_t t = 0;
  for (int64_t b = 0; b < B; ++b) {
    if (lens && b >= n_lens) break;
    int64_t L = lens ? (lens[b] < S ? lens[b] : S) : S;
    for (int64_t s = 0; s < L; ++s, ++t) {
      float mn = INFINITY, mx = -INFINITY;
      for (int64_t f1 = 0; f1 < F1; ++f1)
        for (int64_t f2 = 0; f2 < F2; ++f2) {
          float v = x[b * sb + s * ss + f1 * sf1 + f2 * sf2];
          if (v < mn) mn = v;
          if (v > mx) mx = v;
        }
      tmin[t] = mn;
      tmax[t] = mx;
    }
  }
  return t;
}

/* exact integer contraction of the bins: acc[m,n] = sum_k (qa[m,k] - za) * qw[n,k]   (int64) */
void osqo_code_gemm(const int16_t* qa, int za, const int8_t* qw, int64_t M, int64_t K, int64_t N, int64_t* acc) {
  for (int64_t m = 0; m < M; ++m)
    for (int64_t n = 0; n < N; ++n) {
      int64_t a = 0;
      for (int64_t k = 0; k < K; ++k) a += (int64_t)(qa[m * K + k] - za) * (int64_t)qw[n * K + k];
      acc[m * N + n] = a;
    }
}
