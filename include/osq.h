/*
 * osq.h -- C ABI of libosq_b200.so: the B200 (sm_100a) fake-quantize / observer / fused
 * fake-quant+Linear hot path of wimh966/outlier_suppression.
 *
 * The reference is pure Python (no FFI); each entry point below replaces the launch sequence of
 * the reference function it cites (paths relative to /root/reference/quant_transformer).  A
 * maintainer binds them with ctypes exactly as outlier_suppression_b200/_lib.py does
 * (INTEGRATION.md shows the stub to drop into quantization/).
 *
 * Conventions
 *   - every pointer is a raw DEVICE pointer unless its name ends in _host; the caller owns all
 *     memory (outputs and workspaces are pre-allocated by the caller);
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy
 *     default stream) and never synchronises the device;
 *   - return value: 0 on success, a negative OSQ_E* code on failure; the message for the calling
 *     thread is available through osq_last_error();
 *   - all floating point data is fp32; quantisation parameters live in device memory so that no
 *     host<->device sync (the reference's `.item()` calls, fake_quant.py:124) is ever needed.
 */
#ifndef OSQ_H_
#define OSQ_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OSQ_OK 0
#define OSQ_EINVAL (-1)   /* bad argument / unsupported shape */
#define OSQ_ECUDA (-2)    /* CUDA runtime / driver error      */
#define OSQ_EARCH (-3)    /* device is not sm_100             */

#define OSQ_VERSION 100

int osq_version(void);
const char* osq_last_error(void);
/* number of SMs of the current device (used to size caller-owned workspaces); <0 on error */
int osq_sm_count(void);
/* bytes of caller-owned scratch every reduction entry point needs (partials + ticket counter) */
int64_t osq_workspace_bytes(void);

/* ---------------------------------------------------------------------------------------------
 * K1  per-tensor fake-quantize.     replaces util_quant.py:11-15 (fake_quantize_per_tensor_affine)
 *     and util_quant.py:48-55 (learnable-plus variant) as called from fake_quant.py:107-126,178-209.
 *
 *   q = clamp((rint(x/s) - x/s + x/s) + z, qmin, qmax);   y = (q - z) * s        (all fp32, true division)
 *
 *   scale/zero_point are device scalars.  lsq_grad_factor > 0 selects the LSQ+ forward: the
 *   effective parameters  z' = gs(rint(z), g), s' = gs(s, g), gs(t,g) = (t - t*g) + t*g  are
 *   derived in-kernel with the reference's exact fp32 op order (util_quant.py:49-51,70-71).
 *   zp_is_int32 != 0: zero_point points at an int32 (FixedFakeQuantize buffer), else at a float.
 *   codes (optional, may be NULL): int16 bin per element, rint(q)  (debug / parity side output).
 *   x and y may alias.  n elements, any alignment.
 * ------------------------------------------------------------------------------------------- */
int osq_fq_per_tensor_f32(const float* x, float* y, int16_t* codes, int64_t n,
                          const float* scale, const void* zero_point, int zp_is_int32,
                          float lsq_grad_factor, int qmin, int qmax, void* stream);

/* K7  GammaResidual + LayerNorm + the LayerNorm's output quantizer as ONE pass (13 B / element instead of 12 + 8 + 9):
 *       u = res * res_gamma + h            (res, res_gamma optional)     model/util_layernorm.py:41-52 (GammaResidual.forward)
 *       ln = (u - mean) * rsqrt(var + eps) [* ln_weight] [+ ln_bias]     model/util_layernorm.py:14-15 (QuantizedLayerNorm) and
 *                                                                        :34-36 (QuantizedSplitLayerNorm: no weight, bias = beta / gamma)
 *       y = fq(ln), bins = q - qmin (optional)                           util_layernorm.py:16-17 -> util_quant.py:11-15 / :48-55
 *     replaces quant_bert.py:211-217 / :296-303 (dense -> dropout -> before_LayerNorm_residual -> LayerNorm) behind the dense.
 *     `ln_out` (optional) receives the un-quantised LayerNorm output.  The quantizer part is bit-identical to K1 applied to
 *     ln_out; the LayerNorm part is plain fp32 (two-pass mean / variance), within 2e-6 of torch's kernels.
 *     hidden % 4 == 0; all fp32 pointers 16-byte aligned. */
int osq_residual_layernorm_fq_f32(const float* h, const float* res, const float* res_gamma, const float* ln_weight,
                                  const float* ln_bias, float eps, int64_t rows, int64_t hidden, const float* scale,
                                  const void* zero_point, int zp_is_int32, float lsq_grad_factor, int qmin, int qmax, float* y,
                                  uint8_t* bins, float* ln_out, void* stream);

/* K1d the per-tensor fake-quantize WITHOUT its fp32 output: only the uint8 bins (bin - qmin) leave, 5 bytes per element instead of 9.
 *     For an activation quantizer (fake_quant.py:107-126 / :170-209) whose output is consumed by fused QLinears alone: those read
 *     the bins (A = NULL, a_codes = bins) and the dequantised tensor of util_quant.py:14 is never needed.  `eff` (device float[2],
 *     optional) receives the effective (scale, zero_point) the launch used -- after LSQ+'s sanitise / grad_scale -- for
 *     osq_dequant_bins_f32.  act = 1 applies GELU (erf form, as K1c) in front: quant_bert.py:278-280 feeding output.dense alone.
 *     n % 4 == 0, x 16-byte aligned. */
int osq_fq_per_tensor_bins_only_f32(const float* x, uint8_t* bins, int64_t n, int act, const float* scale, const void* zero_point,
                                    int zp_is_int32, float lsq_grad_factor, int qmin, int qmax, float* eff, void* stream);

/*     the fp32 tensor K1 would have written for those bins: y = (bin + qmin - z) * s with (s, z) = eff (util_quant.py:14).
 *     Bit-identical to K1's output whenever z is integer valued (FixedFakeQuantize always; LSQ+ except for its rare 1-ulp
 *     zero-point drift, where the un-clamped bins are rebuilt as clamp((q - rint(z)) + z) like util_quant.py:13). */
int osq_dequant_bins_f32(const uint8_t* bins, const float* eff, int qmin, int qmax, float* y, int64_t n, void* stream);

/* K2  per-channel (ch_axis = 0) fake-quantize of a [rows, cols] matrix.
 *     replaces util_quant.py:18-26 (fake_quantize_per_channel_affine), fake_quant.py:119-122. */
/* K1b the same per-tensor fake-quantize with the bins as a uint8 side output in the operand format of
 *     osq_fused_fq_linear (bin - qmin, needs qmax - qmin <= 255): a QLinear that consumes y (quantized_module.py:71-72)
 *     can be fed `bins` instead of re-reading and re-quantising the fp32 tensor (A = NULL, a_codes = bins). */
int osq_fq_per_tensor_bins_f32(const float* x, float* y, uint8_t* bins, int64_t n, const float* scale,
                               const void* zero_point, int zp_is_int32, float lsq_grad_factor, int qmin,
                               int qmax, void* stream);

/* K1c an activation and the quantizer behind it as ONE pass: y = fq(act(x)), bins optional (NULL: none).  replaces
 *     model/quant_bert.py:278-280 (intermediate_act_fn followed by intermediate_act_fn_post_act_fake_quantize): act = 1 is GELU
 *     in its erf form with the operation order of ATen's CUDA kernel (bit-identical to torch.nn.functional.gelu on the
 *     same device followed by K1 / K1b); act = 0 is K1b.  9 bytes per element (8 without bins) instead of 8 + 9. */
int osq_act_fq_per_tensor_bins_f32(const float* x, float* y, uint8_t* bins, int64_t n, int act, const float* scale,
                                   const void* zero_point, int zp_is_int32, float lsq_grad_factor, int qmin,
                                   int qmax, void* stream);

int osq_fq_per_channel_f32(const float* x, float* y, int16_t* codes, int64_t rows, int64_t cols,
                           const float* scale, const int32_t* zero_point, int qmin, int qmax,
                           void* stream);

/* ---------------------------------------------------------------------------------------------
 * Token geometry shared by the observer kernels.  replaces observer.py:72-98 (remove_padding /
 * reshape_batch_embedding): the activation is viewed as [B, S, F1, F2] with element strides
 * (sb, ss, sf1, sf2); token (b, s) owns the F1*F2 features; it is valid iff s < lens[b]
 * (lens == NULL: every token valid; b >= n_lens: batch entry skipped -- the reference's zip()).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int64_t B, S, F1, F2;
  int64_t sb, ss, sf1, sf2;
} osq_tokens_t;

/* Running-statistics epilogue shared by the observer entry points.
 *   mode 0: none (only cur_minmax[2] is written)
 *   mode 1: running average  m <- (m*cnt + cur)/(cnt+1)   observer.py:194-202 (Avg* observers);
 *           cnt == 0 takes the `first batch` branch (m <- cur).
 *   mode 2: running extrema  m <- min/max(m, cur)          observer.py:143-144
 * If scale_out != NULL the quantisation parameters of observer.py:100-119 (calculate_qparams) are
 * recomputed from the updated state: scale_out[0] (fp32) and zp_out (int32 when zp_out_is_int32,
 * else fp32 -- the LSQ+ Parameter). */
typedef struct {
  int mode;
  int cnt;
  float* state_min;    /* [1] device: the observer's min_val buffer (updated in place) */
  float* state_max;    /* [1] device: the observer's max_val buffer */
  float* scale_out;    /* [1] device or NULL */
  void* zp_out;        /* [1] device or NULL */
  int zp_out_is_int32;
  int qmin, qmax, symmetric;
} osq_stat_epilogue_t;

/* K3  masked global min/max.  replaces observer.py:188-193 (+ _aminmax) for AvgMinMaxObserver /
 *     MinMaxObserver(ch_axis=-1).  cur_minmax[2] device output.  workspace: osq_workspace_bytes(). */
int osq_minmax_masked_f32(const float* x, const osq_tokens_t* tok, const int64_t* lens, int n_lens,
                          float* cur_minmax, const osq_stat_epilogue_t* epi, void* workspace,
                          void* stream);

/* flat (no token geometry) global min/max of n elements: observer.py:227 with seq_pos == -1 */
int osq_minmax_flat_f32(const float* x, int64_t n, float* cur_minmax,
                        const osq_stat_epilogue_t* epi, void* workspace, void* stream);

/* K4a per-token min/max over the feature axis.  replaces observer.py:64-65 (value.max(1)/min(1))
 *     fused with the pad removal.  tmin/tmax have B*S entries indexed b*S+s; invalid tokens get
 *     (+inf, -inf).  n_valid (device int32[1]) receives the number of valid tokens T. */
int osq_token_minmax_f32(const float* x, const osq_tokens_t* tok, const int64_t* lens, int n_lens,
                         float* tmin, float* tmax, int32_t* n_valid, void* stream);

/* K4b token pruning on the per-token vectors.  replaces observer.py:50-70 (quantile_range /
 *     cac_thres / prune_token) + the clip + _aminmax of :69,:227, which equal (lower, upper).
 *     abs_tmax_sorted / abs_tmin_sorted: ascending sort of |tmax|, |tmin| over the n_slots entries
 *     with invalid tokens mapped to +inf (so the first T entries are the valid ones).
 *     Quantile rank/lerp follow torch.quantile(..., interpolation='linear') in fp32. */
int osq_prune_select_f32(const float* tmin, const float* tmax, const float* abs_tmin_sorted,
                         const float* abs_tmax_sorted, int64_t n_slots, const int32_t* n_valid,
                         float percentile, float* cur_minmax, const osq_stat_epilogue_t* epi,
                         void* workspace, void* stream);

/* K4b' the same selection without any sort: the two order statistics torch.quantile interpolates between are
 *      found by an exact radix select on the fp32 bit patterns of |tmax| / |tmin|; bit-identical to
 *      osq_prune_select_f32 on sorted inputs.  Up to 32768 slots: one CTA, one launch.  Above: six launches
 *      spread over the SMs (four digit passes, the rank+1 value, clip + aminmax + running statistics) with
 *      their state in `workspace` (osq_workspace_bytes(), zero-initialised once; re-armed by the kernels). */
int osq_prune_select_unsorted_f32(const float* tmin, const float* tmax, int64_t n_slots,
                                  const int32_t* n_valid, float percentile, float* cur_minmax,
                                  const osq_stat_epilogue_t* epi, void* workspace, void* stream);


/* K4c the whole AvgPruneMinMaxObserver step (observer.py:50-70,214-237) as two back-to-back launches: the per-token
 *     pass of K4a -- which also histograms the first radix digit of |tmax| / |tmin| into the workspace with fire-and-forget
 *     L2 reductions -- then, by programmatic dependent launch, ONE small CTA that picks the first-digit bins of rank lo and
 *     lo+1 on both sides, compacts their members into shared memory in one sweep of the [T] vectors, finishes the exact
 *     select there, and runs the clip / aminmax sweep and the running-statistics epilogue.  Bit-identical to K4a + K4b'.
 *     tmin / tmax ([B*S] fp32) and n_valid (int32[1]) are caller-owned scratch; workspace: osq_workspace_bytes(), zeroed once.
 *     percentile_dev (optional device float[1]): when given it overrides `percentile` at run time, so a calibration forward
 *     captured once in a CUDA graph can be replayed for every ratio of token_wise_clipping.find_ratio (:50-66). */
int osq_prune_observe_f32(const float* x, const osq_tokens_t* tok, const int64_t* lens, int n_lens, float percentile,
                          const float* percentile_dev, float* tmin, float* tmax, int32_t* n_valid, float* cur_minmax,
                          const osq_stat_epilogue_t* epi, void* workspace, void* stream);

/*     The same step for `n` calibration batches of one geometry in ONE call (observer.py:214-237 called once per batch by
 *     ptq_glue_quant.calibrate / token_wise_clipping.calibrate): results and running statistics are what n consecutive calls of
 *     osq_prune_observe_f32 produce (epis[i] carries batch i's cnt), but batch i + 1's per-token pass is launched as a programmatic
 *     dependent of batch i's one-CTA select tail and runs next to it instead of behind it -- only this library's own launches sit in
 *     between, which is what makes skipping the dependency safe.  xs / curs / epis are HOST arrays of n entries; tmin2 / tmax2
 *     ([2 * ceil4(B*S)] fp32) and n_valid2 (int32[2]) are two alternating scratch sets; the workspace holds two first-digit tables. */
int osq_prune_observe_many_f32(const float* const* xs, int n, const osq_tokens_t* tok, const int64_t* lens, int n_lens, float percentile,
                               const float* percentile_dev, float* tmin2, float* tmax2, int32_t* n_valid2, float* const* curs,
                               const osq_stat_epilogue_t* epis, void* workspace, void* stream);

/* Token-wise clipping (solver/token_wise_clipping.py:50-66) re-calibrates every AvgPruneMinMaxObserver for every candidate ratio, with
 * activation fake-quant switched off (set_ratio, :12-19): the per-token extrema of an (observer, batch) pair do not depend on the
 * ratio.  osq_token_minmax_hist_f32 records them once -- tmin / tmax [B*S], n_valid and the first-digit table hist0 (uint32
 * [2 * 2048 + 64], zeroed by the caller) -- and osq_prune_select_cached_f32 evaluates observer.py:50-70 on any number of recorded
 * problems in ONE launch (one CTA each): cur[0..1] = (min, max) after pruning at `percentile` (or *percentile_dev), bit-identical
 * to osq_prune_observe_f32 on the original activation.  The running average / qparams follow with osq_replay_average_f32. */
typedef struct {
  const float* tmin;
  const float* tmax;
  int64_t n_slots;         /* B * S */
  const int32_t* n_valid;
  const uint32_t* hist0;
  float* cur;              /* device float[2] */
} osq_select_problem_t;

int osq_token_minmax_hist_f32(const float* x, const osq_tokens_t* tok, const int64_t* lens, int n_lens, float* tmin, float* tmax,
                              int32_t* n_valid, uint32_t* hist0, void* stream);
int osq_prune_select_cached_f32(const osq_select_problem_t* problems, int n_problems, float percentile, const float* percentile_dev,
                                void* stream);

/* AvgQuantileObserver.forward (observer.py:253-282) in two launches: K3 (masked min/max -> cur_minmax), then a histogram
 * of |x| over the valid tokens in `bins` equal bins of [0, R], R = max(-min, max) (ATen CPU histc binning, fp32), the
 * reference's sequential cumulative scan (`cur_total + cnt >= threshold * numel`, compared in fp32), the clip of
 * (min, max) to +-(i + 0.5) * R / bins and the running average / qparams of `epi`.  hist: caller-owned uint32[bins],
 * zero-initialised once (the kernel re-arms it).  bins <= 8192. */
int osq_quantile_observe_f32(const float* x, const osq_tokens_t* tok, const int64_t* lens, int n_lens, int bins,
                             double threshold, uint32_t* hist, float* cur_minmax, const osq_stat_epilogue_t* epi,
                             void* workspace, void* stream);

/* Rank-sharded calibration (no counterpart in the reference, which is single-GPU): replays the running average of
 * observer.py:194-202 over a slot table [n_obs, n_batches, 2] (per-batch (min, max) of every observer, summed across ranks
 * by ONE all-reduce) in batch order, one thread per observer, and rewrites every observer's state and its quantizer's
 * (scale, zero_point) (observer.py:100-119) through `targets` (device array) -- no host round trip. */
typedef struct {
  float* state_min;   /* [1] device: observer.min_val (updated in place; +-inf = no batch seen yet) */
  float* state_max;   /* [1] device: observer.max_val */
  float* scale_out;   /* [1] device or NULL */
  void* zp_out;       /* [1] device or NULL */
  int zp_out_is_int32;
  int qmin, qmax, symmetric;
} osq_replay_target_t;

int osq_replay_average_f32(const float* table, int n_obs, int n_batches, int cnt0, const osq_replay_target_t* targets,
                           void* stream);
/* The same replay with NO collective: peer_tables (device array of `world` device pointers) are the slot tables of all ranks,
 * mapped into this process (NVLink peer memory / CUDA symmetric memory); slot (observer, batch b) is loaded from the table of
 * rank b mod world, the rank that processed batch b.  The caller orders the launch after every rank's writes (a cross-rank
 * barrier on the stream) and keeps the tables untouched until every rank has replayed. */
int osq_replay_average_peer_f32(const float* const* peer_tables, int world, int n_obs, int n_batches, int cnt0,
                                const osq_replay_target_t* targets, void* stream);

/* Exchange AND replay as one launch over peer memory (NVLink / NVSwitch), no collective library and no host-side barrier:
 * `regions[r]` (device array of `world` pointers) is rank r's symmetric-memory region, mapped into this process, laid out as
 * float slots[2][n_obs * n_batches * 2] followed by uint32 flags[world].  The kernel publishes this rank's slots (batches b with
 * b mod world == rank) from `local_table` into generation (pass & 1) of its own region, signals every peer's flags[rank] = pass
 * with a remote release-store, waits (acquire loads of its own flags) until every rank has signalled, then runs the replay of
 * osq_replay_average_f32 with slot (i, b) loaded from rank b mod world's region.  pass = *pass_counter + 1, written back by the
 * kernel (graph-replayable); the two generations remove the need for a trailing barrier.  A peer that does not signal within
 * OSQ_EXCHANGE_TIMEOUT_MS (default 5000) makes the launch store `pass` to *err_flag and return without touching the targets. */
int osq_replay_exchange_f32(const float* local_table, float* const* regions, int rank, int world, int n_obs, int n_batches, int cnt0,
                            const osq_replay_target_t* targets, uint32_t* pass_counter, uint32_t* err_flag, void* stream);

/* per-row min/max of a [rows, cols] matrix with the running-extrema update of
 * MinMaxObserver(ch_axis=0) (observer.py:141-144) and per-row calculate_qparams.
 * state_min/state_max [rows] are updated in place (first = 1 overwrites them). */
int osq_rowwise_minmax_qparams_f32(const float* w, int64_t rows, int64_t cols, int first,
                                   float* state_min, float* state_max, float* scale_out,
                                   int32_t* zp_out, int qmin, int qmax, int symmetric, void* stream);

/* observer.py:100-119 (calculate_qparams) for n independent (min, max) pairs; true fp32 division.
 * zp_f32 / zp_i32: either may be NULL. */
int osq_calc_qparams_f32(const float* min_val, const float* max_val, int64_t n, int qmin, int qmax,
                         int symmetric, float* scale, float* zp_f32, int32_t* zp_i32, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K5  MSE-grid evaluation.  replaces observer.py:420-432 (MSEFastObserver.loss_fx) evaluated for C
 *     candidate (min, max) pairs in ONE pass over x:  loss[c] = mean((fq(x; qparams(c)) - x)^2).
 *     Sums are accumulated in fp64 (the reference's fp32 `.mean()` is order-dependent; tolerance
 *     stated in tests).  cand_scale[C] (fp32), cand_zp[C] (fp32, integer valued) are the candidates'
 *     quantisation parameters.  loss_sum[C] (fp64) receives the SUM of squared errors (caller
 *     divides by the number of valid elements, returned in n_valid (int64 device)).
 * ------------------------------------------------------------------------------------------- */
int osq_mse_multi_f32(const float* x, const osq_tokens_t* tok, const int64_t* lens, int n_lens,
                      const float* cand_scale, const float* cand_zp, int n_cand, int qmin, int qmax,
                      double* loss_sum, int64_t* n_valid, void* stream);

/* K5c the whole per-TENSOR search of MSEFastObserver / AvgMSEFastObserver (observer.py:434-494,496-533) as ONE cooperative
 *     launch: masked min / max, the one_side_dist decision (*one_side_state: -1 undecided, 0 no, 1 pos, 2 neg; decided on
 *     the first call and kept, observer.py:528-529), then the 1-D bounded Brent (symmetric or one-sided) or the nested
 *     2-D search (outer Brent over the range, inner Brent over the shift per evaluation, a final inner search) with SciPy's
 *     _minimize_scalar_bounded restated in fp64; every loss evaluation is a grid-wide reduction behind one grid barrier.
 *     out4 (device double[4]): best_min, best_max, x_min, x_max.  evals (device int32, optional): loss evaluations.
 *     scratch: osq_mse_tensor_scratch_bytes() of device memory.  Same tolerance as K5b (tests). */
int osq_mse_brent_tensor_f32(const float* x, const osq_tokens_t* tok, const int64_t* lens, int n_lens, int qmin, int qmax,
                             int symmetric, int32_t* one_side_state, double* out4, int32_t* evals, void* scratch, void* stream);
int64_t osq_mse_tensor_scratch_bytes(void);

/* K5b per-row bounded-Brent search of MSEFastObserver(ch_axis=0, symmetric) -- observer.py:483-517
 *     with scipy.optimize.minimize_scalar(method='Bounded') restated on-chip (one CTA per row, row
 *     resident in shared memory).  out_min/out_max [rows]; evals [rows] (int32) optional. */
int osq_mse_brent_rows_f32(const float* w, int64_t rows, int64_t cols, int qmin, int qmax,
                           int one_side /*0 no,1 pos,2 neg*/, float* out_min, float* out_max,
                           int32_t* evals, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Weight preparation for the fused kernel (once per weight version; replaces the per-forward
 * weight fake-quant of quantized_module.py:72).  W [N, K] fp32 (gamma already folded by
 * gamma_migration.py:46-76), scale[N]/zp[N] from the weight quantizer.
 *   codes  [N, K] int8  : q - zp  (the dequantised weight is codes * scale[n])
 *   rowsum [N]    int32 : sum_k codes[n, k]   (zero-point correction of the u8 x s8 contraction)
 * Requires qmin - zp >= -128 and qmax - zp <= 127 for every row: any zp in [qmin, qmax] qualifies when
 * qmax - qmin <= 127 (<= 7 bits); an 8-bit range must be the symmetric one, [-128, 127] with zp == 0.
 * Asymmetric 8-bit weights (range [0, 255]) return OSQ_EINVAL; out-of-range bins saturate (never wrap).
 * ------------------------------------------------------------------------------------------- */
int osq_pack_weight_s8(const float* w, int64_t N, int64_t K, const float* scale, const int32_t* zp,
                       int qmin, int qmax, int8_t* codes, int32_t* rowsum, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K6  fused activation fake-quant + per-channel weight fake-quant + Linear.
 *     replaces  fake_quant.py:107-126 / :178-209 (activation quantizer)  ->
 *               quantized_module.py:71-72 (QLinear.forward: weight fq + F.linear)
 *
 *   Y[m, n] = s_a * w_scale[n] * sum_k (qa[m,k] - Z) * wcodes[n,k] + bias[n]
 *   qa = clamp(rint(A/s_a) + Z, a_qmin, a_qmax)  with  Z = rint(z_a)  (LSQ+ effective s', z' when
 *   lsq_grad_factor > 0, as in K1).  A [M, K] fp32 row-major (raw, un-quantised, or already
 *   fake-quantised -- fq is idempotent), Y [M, N] fp32 row-major.
 *
 *   One persistent, warp-specialised tcgen05 kernel: fp32 A arrives by TMA in per-warp landing slots, converter
 *   warps turn it into integer bins in the UMMA shared-memory layout, the packed weight bins arrive by TMA,
 *   `tcgen05.mma kind::i8` (u8 x s8 -> s32, exact for every bit-width <= 8) accumulates in TMEM, the epilogue
 *   applies the zero-point correction, scales and bias and writes fp32 Y with TMA stores.
 *   mma_kind: 0 = auto, 1 = kind::i8 (the only kind built; anything else is rejected).
 *
 *   Shape contract: K % 128 == 0, K <= 32768 (int32 accumulators), N % 16 == 0, M >= 1; A, Y, w_codes 16-byte
 *   aligned.  a_codes (optional uint8 [M, K], 16-byte aligned): receives the activation bins (bin - a_qmin) the
 *   kernel fed to the tensor core (parity side output).  When K > 1024 the converted block no longer fits in
 *   shared memory; if a_codes is given it doubles as an L2-resident code cache so fp32 A is read and quantised
 *   once (otherwise A is re-read and re-quantised for every chunk of N).
 *   Bins-in: A == NULL and a_codes != NULL -> a_codes already holds the activation bins (osq_fq_per_tensor_bins_f32 of
 *   the upstream activation quantizer); the kernel TMA-loads them straight into the UMMA layout and converts nothing.
 *   The result is bit-identical to the fp32-in launch.
 *   Several Linears that consume the same A (BERT query | key | value) may be served by one call over their
 *   row-concatenated w_codes / w_scale / w_rowsum / bias: every output column is computed independently.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const float* A;        /* [M, K] fp32, or NULL for a bins-in launch (then a_codes is the input) */
  int64_t M, K;
  const float* a_scale;  /* device [1] */
  const void* a_zp;      /* device [1] */
  int a_zp_is_int32;
  float lsq_grad_factor; /* 0 for FixedFakeQuantize */
  int a_qmin, a_qmax;
  const int8_t* w_codes; /* [N, K] from osq_pack_weight_s8 */
  const float* w_scale;  /* [N] */
  const int32_t* w_rowsum; /* [N] */
  const float* bias;     /* [N] or NULL */
  float* Y;
  int64_t N;
  int mma_kind;
  uint8_t* a_codes;      /* optional [M, K] u8: bins side output AND the kernel's code cache (see above) */
  void* debug_trace;     /* optional device int64[2048]: clock64 timeline of CTA 0 + per-CTA globaltimer start/end (profiling aid), else NULL */
  /* ---- optional output stage: the NEXT activation quantizer fused into the epilogue (quant_bert.py:277-280: dense ->
   * intermediate_act_fn -> intermediate_act_fn_post_act_fake_quantize).  With out_scale != NULL the kernel writes
   * Y = fq(act(Linear)) instead of the Linear's output -- bit-identical to K1 applied to act(Linear) -- and, if out_bins is
   * given, that quantizer's bins (bin - out_qmin, u8 [M, N], 16-byte aligned) in the operand format of a bins-in launch,
   * so Linear -> activation -> quantizer -> Linear is two launches with a 1 B / element hand-off. */
  int out_act;             /* 0 = none, 1 = GELU (erf form; op order of ATen's CUDA kernel) */
  const float* out_scale;  /* device [1] or NULL (no output stage) */
  const void* out_zp;      /* device [1] */
  int out_zp_is_int32;
  float out_lsq_grad_factor; /* 0 for FixedFakeQuantize; LSQ+: 1 / sqrt(M * N * out_qmax) */
  int out_qmin, out_qmax;
  uint8_t* out_bins;       /* optional */
} osq_fused_linear_t;

int osq_fused_fq_linear(const osq_fused_linear_t* args, void* stream);

/* K6 for a list of INDEPENDENT sites (no site reads what another site of the list writes): runs of compatible sites --
 * up to 4, same grid and launch plan family -- execute as ONE persistent launch in which every CTA walks the list, so the
 * launch gap, the wait for the slowest CTA and the first-load latency are paid once per run instead of once per site;
 * incompatible sites fall back to individual launches, in list order.  Results are bit-identical to n_sites calls of
 * osq_fused_fq_linear.  Typical use: the Linears of several layers fed from independent buffers (BASELINE config 2's
 * synthetic site stack), or sibling projections that cannot share one weight concatenation. */
int osq_fused_fq_linear_multi(const osq_fused_linear_t* sites, int n_sites, void* stream);

/* A per-tensor activation quantizer as the attention kernels consume it: device-resident parameters (no .item() sync),
 * FixedFakeQuantize (lsq_grad_factor = 0, fake_quant.py:123-125) or LSQPlusFakeQuantize (lsq_grad_factor = 1 / sqrt(numel * qmax),
 * float zero_point, fake_quant.py:188-195); at most 8 bits. */
typedef struct {
  const float* scale;      /* device [1] */
  const void* zero_point;  /* device [1], float or int32 */
  int zp_is_int32;
  float lsq_grad_factor;
  int qmin, qmax;
} osq_quantizer_t;

/* K8  attention scores with the query / key quantizers as prologue, one launch:
 *       scores = fq_q(Q) @ fq_k(K)^T  [* out_mul]  [+ mask]
 *     replaces model/quant_bert.py:148-150 (query_permute_post_act_fake_quantize, key_transpose_post_act_fake_quantize,
 *     torch.matmul) and, with out_mul / mask, :169-172 (/ sqrt(d), + attention_mask).  Q and K are the [batch, tokens, heads * d]
 *     projections viewed as [batch, heads, tokens, d] (transpose_for_scores): strides {batch, head, token} in elements, the
 *     channel stride is 1.  Both operands are fake-quantised activations, so the product of the dequantised tensors equals
 *     s_q s_k * sum (q_bin - Zq)(k_bin - Zk): an exact integer contraction (u8 x u8 -> s32) and one scale.  out_mul = the fp32
 *     reciprocal ATen multiplies by for `scores / math.sqrt(d)` (1 = none); mask = additive [batch, sk] or NULL.
 *     d in {32, 64, 128}; sk % 4 == 0; scores [batch, heads, sq, sk] fp32. */
int osq_attn_scores_fq_f32(const float* q, const float* k, int64_t batch, int64_t heads, int64_t sq, int64_t sk, int64_t d,
                           const int64_t* q_strides, const int64_t* k_strides, const osq_quantizer_t* qq, const osq_quantizer_t* kq,
                           float out_mul, const float* mask, float* scores, void* stream);

/* K9  attention context with the probability / value quantizers as prologue and the context quantizer as epilogue:
 *       context = fq_p(P) @ fq_v(V)            written as [batch, sq, heads * d] (the permute + view of quant_bert.py:189-191)
 *       oq != NULL: context <- fq_o(context), bins (optional) = its uint8 bins for the bins-in Linear behind it
 *     replaces quant_bert.py:185-187 (attention_probs_post_act_fake_quantize, value_permute_post_act_fake_quantize, torch.matmul)
 *     and :192-193 (context_view_post_act_fake_quantize).  P [batch, heads, sq, sk] fp32 is read once (it is 8x larger than
 *     everything else this kernel touches); V like K8's operands. */
int osq_attn_context_fq_f32(const float* probs, const float* v, int64_t batch, int64_t heads, int64_t sq, int64_t sk, int64_t d,
                            const int64_t* v_strides, const osq_quantizer_t* pq, const osq_quantizer_t* vq, const osq_quantizer_t* oq,
                            float* context, uint8_t* bins, void* stream);

/* LSQ+ backward (fine stage `learn_scale`, token_wise_clipping.py:72-108; gradients of
 * util_quant.py:48-55):  dx = dy * 1[qmin <= q <= qmax];
 *   dscale += g * sum dy * (inside ? rint(x/s) - x/s : q_clamped - z);  dzp += g * sum dy * (inside ? 0 : -s)
 * grad_acc (device double[2]) must be zeroed by the caller. */
int osq_lsqplus_backward_f32(const float* x, const float* dy, float* dx, int64_t n,
                             const float* scale, const float* zero_point, float lsq_grad_factor,
                             int qmin, int qmax, double* grad_acc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OSQ_H_ */
