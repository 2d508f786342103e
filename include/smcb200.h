/* smcb200 - B200-native Sequential-Monte-Carlo inner loop, C ABI.
 *
 * Drop-in boundary for the hot path of tingiskhan/pyfilter (reference v0.29.0, pure Python on PyTorch).  The reference has no
 * FFI; its boundary is the Python plug-in surface.  Every entry point below names the reference
 * interface it stands behind (paths relative to /root/reference/pyfilter/).  pyfilter_b200/ binds these with ctypes and
 * re-creates the reference's classes on top (INTEGRATION.md shows the binding a pyfilter maintainer would add).
 *
 * Conventions
 *   - plain C types only; every pointer named *_dev is a CUDA device pointer, *_host a host pointer;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); work is enqueued on it, nothing synchronises
 *     unless stated;
 *   - functions return 0 on success, a negative SMCB_E* code otherwise; smcb_last_error() gives the message (thread local);
 *   - a handle is not thread-safe; distinct handles are independent;
 *   - device layout: one column per independent filter (the reference's batch dimension, filters/base.py:93-119); a column is
 *     a contiguous row of `ld` elements (ld = particles rounded up to 4096); states are SoA x[dim][column][particle];
 *   - kernels of one move are chained with programmatic dependent launch (SMCB_NO_PDL=1 in the environment disables it);
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns SMCB_ENODEVICE.
 */
#ifndef SMCB200_H
#define SMCB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMCB_VERSION 200

enum { SMCB_OK = 0, SMCB_EINVAL = -1, SMCB_ECUDA = -2, SMCB_ENODEVICE = -3, SMCB_EUNSUPPORTED = -4, SMCB_ESTATE = -5 };

/* models: the user-supplied stochproc callables of the reference become a compiled zoo */
enum { SMCB_LG_AR1 = 0, SMCB_SINE_EM = 1, SMCB_SV_AR1 = 2, SMCB_LORENZ63_EM = 3,
       SMCB_USER_MODEL = 4 /* a build of the library compiled with -DSMCB_USER_MODEL_HEADER=<file>: the user's mean_scale / observation
                              density as device functions (csrc/models.h); raw parameters: UserModel::NRAW values per column */ };
/* proposals: filters/particle/proposals/bootstrap.py:4-17, proposals/linear.py:13-89 */
/*            proposals/linearized.py:9-73 with proposals/utils.py:30-146 (ModeFinder; closed-form derivatives of the zoo's models) */
/*            proposals/nested.py:8-50 (num_samples inner draws per particle, one of them proposed) */
enum { SMCB_BOOTSTRAP = 0, SMCB_LINEAR_GAUSSIAN_OBSERVATIONS = 1, SMCB_LINEARIZED = 2, SMCB_NESTED = 3 };
/* filters: filters/particle/sisr.py:7-56, filters/particle/apf.py:9-46 */
enum { SMCB_SISR = 0, SMCB_APF = 1, SMCB_GPF = 2 };   /* GPF: filters/particle/gpf.py with its default GaussianProposal */
/* resamplers: resampling.py:24-52 (systematic), :55-65 (multinomial) */
enum { SMCB_SYSTEMATIC = 0, SMCB_MULTINOMIAL = 1 };

#define SMCB_MAX_RAW_PARAMS 8

/* Raw parameter order per model (float32, one value per column or one shared value):
 *   SMCB_LG_AR1       alpha, beta, sigma, a, b, s             x' = alpha + beta x + sigma e ;       y = b + a x + s v
 *   SMCB_SINE_EM      gamma, sigma, dt, a, b, s               x' = x + sin(x - gamma) dt + sigma sqrt(dt) e ; y = b + a x + s v
 *   SMCB_SV_AR1       mu, phi, sigma_v                        x' = mu + phi (x - mu) + sigma_v e ;  y ~ N(0, exp(x/2))
 *   SMCB_LORENZ63_EM  s, r, b, sigma, dt, obs_a, obs_s        Euler-Maruyama Lorenz-63 ;            y = obs_a (x1, x3) + obs_s v
 */
typedef struct smcb_config {
  int32_t model;          /* SMCB_LG_AR1 ...                                              (the `model` argument, filters/base.py:22) */
  int32_t proposal;       /* SMCB_BOOTSTRAP ...                                           (`proposal`, filters/particle/base.py:24) */
  int32_t algorithm;      /* SMCB_SISR / SMCB_APF / SMCB_GPF                                          (the filter class) */
  int32_t resampler;      /* SMCB_SYSTEMATIC / SMCB_MULTINOMIAL                            (`resampling`, filters/particle/base.py:23) */
  int64_t particles;      /* N                                                            (`particles`, filters/particle/base.py:22) */
  int32_t batch;          /* B >= 1 independent filters                                    (set_batch_shape, filters/base.py:93-119) */
  int32_t n_raw_params;   /* number of rows in `params_host` */
  const float* params_host; /* (n_raw_params, param_cols) row-major, host memory */
  int32_t param_cols;     /* 1 (shared) or `batch` (one value per column) */
  float ess_threshold;    /* relative ESS threshold of SISR                               (`ess_threshold`, filters/particle/base.py:25,42) */
  uint64_t seed;          /* Philox key */
  int32_t history_rows;   /* rows kept for filter means / variances / log-likelihood increments (T + 1 for batch_filter) */
  int32_t fold_lookahead; /* APF: fold log p(y_{t+1}|.) into the stored weights when y_{t+1} is known (saves one pass) */
  int32_t exact_weights;  /* 0 (default): the resampling weights the handle derives from its log-weights are rounded to multiples
                             of 2^-52 (|dW| <= 2^-53 ~ 1.1e-16 absolute), so the sequential fp64 prefix sum of resampling.py:47
                             never rounds and every column takes the chain-free path; 1: weights are exactly
                             fl32(exp(lw - max) * fl32(1/sum)) and columns with tiny weights take the transducer scan.  Either way
                             the ancestors are bit-exact for the weights used (smcb_filter_dump_noise returns them). */
  int32_t column_offset;  /* global index of this handle's column 0.  The Philox counters are (particle group, column_offset + column,
                             move, purpose): shards of ONE batch of filters on different ranks (or handles) pass the offset of their
                             first column so that no two columns of the batch share a random stream, and the result does not depend
                             on how the batch is split */
  int32_t lin_steps;      /* SMCB_LINEARIZED: `n_steps`, `alpha`, `use_second_order` of proposals/linearized.py:22 (the default functorch  */
  float lin_alpha;        /* path of ModeFinder.find_mode, proposals/utils.py:96-146)                                                   */
  int32_t lin_second_order;
  int32_t nested_samples; /* SMCB_NESTED: `num_samples` of proposals/nested.py:17 (1 .. 256) */
} smcb_config;

typedef struct smcb_filter smcb_filter;

typedef struct smcb_info {
  int64_t particles, ld;
  int32_t batch, state_dim, obs_dim;
  int32_t t;              /* filter moves completed */
  int32_t history_rows;
  int32_t slow_tiles;     /* exact-scan tiles that needed the sequential fallback so far (diagnostic; synchronises) */
  int64_t kernel_launches;/* kernels launched by this handle so far */
  int64_t lb_windows;     /* reserved (0) */
  int32_t lb_fail;        /* reserved (0) */
  int32_t reserved;
} smcb_info;

/* library */
int smcb_version(void);
/* (sizeof(smcb_config) << 16) | sizeof(smcb_info) of the BUILT library: a binding compares it with its own view of the structs before
 * it passes one (a stale binary next to newer sources otherwise reads them with the wrong layout) */
int smcb_abi_signature(void);
const char* smcb_last_error(void);
int smcb_device_count(void);

/* lifetime ............................................................ ParticleFilter.__init__ (filters/particle/base.py:19-48) */
int smcb_filter_create(const smcb_config* cfg, smcb_filter** out);
int smcb_filter_destroy(smcb_filter* f);
/* new parameter values for an existing handle (SMC2 / PMMH rebuild the model per theta: filters/base.py:75-83) */
int smcb_filter_set_params(smcb_filter* f, const float* params_host, int32_t n_raw_params, int32_t param_cols, void* stream);
int smcb_filter_info(smcb_filter* f, smcb_info* out);
/* a new Philox key for the coming moves (the proposal filter of a PMMH sweep re-filters the same data again and again:
 * inference/batch/mcmc/utils.py:52-55 - every run needs its own random stream) */
int smcb_filter_set_seed(smcb_filter* f, uint64_t seed);

/* ParticleFilter.initialize (filters/particle/base.py:87-103): x_0 ~ p_0, log w = 0, ll = 0, prev_inds = arange; history row 0 */
int smcb_filter_initialize(smcb_filter* f, void* stream);
/* init_state= of batch_filter (filters/base.py:141,152): the caller has written x / log w through the borrowed pointers below
 * (smcb_filter_ptr); recompute normalisers, ESS and moments for that state and set the move counter to `t`. */
int smcb_filter_refresh_state(smcb_filter* f, int32_t t, void* stream);

/* observations for the coming moves: y_dev is (count, obs_dim) float32 on the device, y_dev[0] belongs to move `base_t` */
int smcb_filter_set_observations(smcb_filter* f, const float* y_dev, int32_t count, int32_t base_t, void* stream);

/* BaseFilter.filter (filters/base.py:188-221) = predict (sisr.py:14-48 / apf.py:16-23) + correct (sisr.py:50-56 / apf.py:25-46)
 * for the next `steps` observations set above; NaN observations propagate only (filters/base.py:213-214). */
int smcb_filter_run(smcb_filter* f, int32_t steps, void* stream);

/* the online pattern of SMC2 / NESS (inference/sequential/smc2.py:53-65, ness.py:56): `steps` launches of ONE move each; when an exchange
 * is attached (below) every move also gathers that exchange - inside the move's own launch for the resident column kernel (the block
 * that finishes a rank's move last waits for the other ranks' values), otherwise with the reader kernel behind it.  COLLECTIVE: every
 * rank of the batch must run this call concurrently (one process per GPU); handles of several "ranks" driven from ONE stream must use
 * smcb_filter_run + smcb_filter_exchange_wait instead. */
int smcb_filter_run_stepwise(smcb_filter* f, int32_t steps, void* stream);

/* measurement aid: runs `steps` moves like smcb_filter_run with CUDA events around every kernel group and returns the summed
 * device time in milliseconds: out_ms_host[0..4] = {APF pre-weight (+finalize), normalize, describe (+chain), expand, fused
 * step}.  Synchronises the stream. */
int smcb_filter_profile(smcb_filter* f, int32_t steps, float* out_ms_host, void* stream);

/* BaseFilter.batch_filter (filters/base.py:140-158) end to end on HOST buffers: initialises, copies y (T, obs_dim) to the
 * device, runs T moves, copies back filter means / variances ((T+1, B, D) each), per-move log-likelihood increments (T+1, B)
 * (row 0 is zero) and the total log-likelihood (B).  Any output pointer may be NULL.  Synchronises the stream. */
int smcb_filter_batch_filter_host(smcb_filter* f, const float* y_host, int32_t T, float* means_host, float* vars_host,
                                  float* ll_steps_host, float* ll_total_host, void* stream);

/* Shards of ONE batch of filters on several GPUs (the theta-particles of SMC2 / NESS split by columns): the per-column
 * log-likelihood values every rank needs for the theta-level ESS test (inference/sequential/state.py:35-44, smc2.py:59-63) are stored
 * by the finalising kernel straight into every rank's buffer over NVLink peer memory - no collective launch.  `peer_ptrs_host[r]` is
 * the device address of rank r's buffer as mapped into THIS process (torch symmetric memory / cudaIpc: plumbing), each at least
 * 2 * 2 * total_columns * 8 bytes and zeroed; `rank` is this handle's position, `first_column` the global index of its column 0.
 * From then on the move that ends every smcb_filter_run publishes (increment of the last move, running total) of its columns.
 * NULL detaches. */
int smcb_filter_attach_exchange(smcb_filter* f, const uint64_t* peer_ptrs_host, int32_t world, int32_t rank, int32_t total_columns,
                                int32_t first_column);
/* enqueues the reader of the latest exchange: waits (on the device, polling this rank's buffer) until the values of every column of
 * the batch have arrived and returns a dense device array (2, total_columns): row 0 the increments, row 1 the totals */
int smcb_filter_exchange_wait(smcb_filter* f, float** out_dev, void* stream);

/* parity hooks: inject the transition noise (D, B, ld) / systematic offsets (B) / multinomial uniforms
 * (B, ld) float64, or dump the ones the kernels generated and the normalised weights (B, ld) the resampler consumed.  NULL
 * switches a hook off. */
int smcb_filter_set_noise(smcb_filter* f, const float* eps_dev, const float* u_dev, const double* U_dev);
int smcb_filter_dump_noise(smcb_filter* f, float* eps_dev, float* u_dev, float* w_dev);
/* NestedProposal: inject the N(0,1) values of the inner samples (num_samples, D, B, ld) and the Exp(1) values (num_samples, B, ld) of
 * the categorical draw - hidden_density.sample(num_samples) and Categorical.sample in proposals/nested.py:29,40; torch.multinomial
 * picks its single sample per row as argmax(probs / Exp(1)).  NULL: Philox (one uniform per particle inverts the prefix sums). */
int smcb_filter_set_nested_noise(smcb_filter* f, const float* z_dev, const float* e_dev);

/* borrowed device pointers into the handle's state (valid until destroy) ........ ParticleFilterCorrection (particle/state.py:72-211) */
enum {
  SMCB_PTR_X = 0,          /* float (D, B, ld)  current particles           `_x`                                            */
  SMCB_PTR_LOGW = 1,       /* float (B, ld)     log-weights                 `_w`                                            */
  SMCB_PTR_PREV_INDS = 2,  /* int32 (B, ld)     ancestors of the last move  `_prev_inds` (widen to int64 on the host side)  */
  SMCB_PTR_MEAN = 3,       /* float (B, D)      latest filter mean          `_mean`                                         */
  SMCB_PTR_VAR = 4,        /* float (B, D)      latest filter variance      `_var`                                          */
  SMCB_PTR_LL = 5,         /* float (B)         latest log p(y_t|y_{1:t-1}) `_ll`                                           */
  SMCB_PTR_LL_TOTAL = 6,   /* float (B)         running log-likelihood      FilterResult.loglikelihood (filters/result.py:42-48) */
  SMCB_PTR_HIST_MEAN = 7,  /* float (rows,B,D)  FilterResult.filter_means   (filters/result.py:50-57)                        */
  SMCB_PTR_HIST_VAR = 8,   /* float (rows,B,D)  FilterResult.filter_variance                                                */
  SMCB_PTR_HIST_LL = 9,    /* float (rows,B)    per-move increments                                                         */
  SMCB_PTR_ESS = 10,       /* float (B) packed copy refreshed by smcb_filter_sync_stats                                     */
  SMCB_PTR_X_OTHER = 11,   /* float (D, B, ld)  the ping-pong twin of SMCB_PTR_X (previous particles)                       */
  SMCB_PTR_RESAMPLE_LOGW = 12 /* float (B, ld)  APF: the resampling log-weights g(y_{t+1}|x_t) + log w_t of the COMING move (apf.py:29),
                                 valid after a move that folded the look-ahead or after the pre-weight pass                         */
};
int smcb_filter_ptr(smcb_filter* f, int32_t what, void** ptr_dev);
/* gathers ESS / resample flags of every column into the packed SMCB_PTR_ESS buffer (get_ess, utils.py:8-20) */
int smcb_filter_sync_stats(smcb_filter* f, void* stream);

/* the proposal plug-in surface (filters/particle/proposals/base.py:52-85) as stand-alone passes of the handle's proposal over a caller's
 * particles `x_dev` (D, B, ld) in the handle's layout (NULL: the handle's current particles); y_dev is ONE observation (obs_dim) on the
 * device.  pre_weight -> log p(y | .) as the APF uses it (apf.py:27; proposals/base.py:69-85, proposals/linear.py:57-86) into out_dev
 * (B, ld).  sample_and_weight -> proposed particles (D, B, ld) and weight increments (B, ld) (proposals/bootstrap.py:10-14,
 * proposals/linear.py:38-55); `eps_dev` (D, B, ld) injects the N(0,1) draws, NULL draws them from Philox at move index `t`. */
int smcb_filter_pre_weight(smcb_filter* f, const float* y_dev, const float* x_dev, float* out_dev, void* stream);
int smcb_filter_sample_and_weight(smcb_filter* f, const float* y_dev, const float* x_dev, const float* eps_dev, int32_t t, float* x_out_dev,
                                  float* w_out_dev, void* stream);
/* relative ESS threshold of the coming SISR moves (filters/particle/base.py:42); a negative value switches resampling off, a value above 1
 * forces it - SISR.correct (sisr.py:50-56) propagates a prediction that SISR.predict has already resampled */
int smcb_filter_set_ess_threshold(smcb_filter* f, float relative_threshold);
/* ParticleFilterCorrection.predict_path (particle/state.py:173-174 -> model.sample_states(num_steps, x_0)): every particle of `x_dev`
 * (NULL: the current ones) simulated `steps` transitions ahead with its observations: x_out_dev (steps, D, B, ld), y_out_dev (steps,
 * obs_dim, B, ld).  A Philox stream of its own (the filter's noise is not replayed). */
int smcb_filter_predict_path(smcb_filter* f, int32_t steps, const float* x_dev, float* x_out_dev, float* y_out_dev, void* stream);

/* one backward step of forward-filtering backward-sampling as the reference writes it (filters/particle/base.py:105-128) for a
 * non-batched filter: for every smoothed particle i (its value at the later time: xnext_dev[i]) an index is drawn from
 * Categorical(logits = lw + log p(xnext_i | x)) over the N particles x_dev (N, D) of the earlier state with log-weights lw_dev (N) -
 * by inversion of the cumulative unnormalised probabilities with the uniform U_dev[i] (float64, NULL: Philox(seed, t)); idx_out_dev (N)
 * int64 and x_out_dev (N, D) = x[idx].  Reference layout, contiguous.  O(N^2), like the reference. */
int smcb_filter_ffbs_step(smcb_filter* f, const float* x_dev, const float* lw_dev, const float* xnext_dev, const double* U_dev, uint64_t seed,
                          int32_t t, int64_t* idx_out_dev, float* x_out_dev, void* stream);

/* theta-level operations of SMC2 / PMMH on the RESIDENT state of a handle (columns = theta-particles): FilterResult.resample
 * (filters/result.py:76-95, particle/state.py:150-158): column b <- column idx_dev[b] (int64 (B)) for particles, log-weights, APF
 * resampling weights, ancestors, statistics, running log-likelihood, latest moments and - with entire_history - the moment / likelihood
 * history rows; FilterResult.exchange (filters/result.py:97-117, particle/state.py:160-168): column b of `dst` <- column b of `src` where
 * mask_dev[b] != 0 (uint8 (B)); both handles must hold the same shapes and stand at the same move index. */
int smcb_filter_resample_columns(smcb_filter* f, const int64_t* idx_dev, int32_t entire_history, void* stream);
int smcb_filter_exchange_columns(smcb_filter* dst, smcb_filter* src, const uint8_t* mask_dev, void* stream);

/* the columns of a handle as self-contained records, for theta-resampling ACROSS handles / ranks (a sharded SMC2 run: the ancestor of a
 * column may live on another GPU, particle/state.py:150-158 on a distributed batch).  export: record[b] = everything the handle holds
 * for column b (particles, both weight rows, ancestors, statistics, running log-likelihood, latest moments, the moment / likelihood
 * history), smcb_filter_column_record_elems() 4-byte elements each, into buf_dev (B, record).  import: column b of the handle <- record
 * idx_dev[b] of a buffer holding n_records records (the all-gather of the ranks' exports; idx NULL: identity); both handles must have
 * been created with the same shapes and stand at the same move index.  `folded` (>= 0) tells whether the imported APF resampling weights
 * are valid for the coming move.  import synchronises (range check). */
int64_t smcb_filter_column_record_elems(smcb_filter* f);
int smcb_filter_export_columns(smcb_filter* f, void* buf_dev, void* stream);
int smcb_filter_import_columns(smcb_filter* f, const void* buf_dev, int32_t n_records, const int64_t* idx_dev, int32_t folded, void* stream);

/* stand-alone operators ......................................................................................................
 * All take a weight matrix with element strides (stride_n, stride_b): the reference's particle-major (N,B) tensor has
 * (B, 1); a single column has (1, 0).  `out_dev` strides are given the same way. */
/* pyfilter.utils.normalize (utils.py:49-64): NaN-safe soft-max over the particle axis; does NOT mutate the input */
int smcb_normalize(const float* logw_dev, int64_t n, int32_t B, int64_t stride_n, int64_t stride_b, float* out_dev,
                   int64_t out_stride_n, int64_t out_stride_b, float* ess_out_dev, void* stream);
/* pyfilter.utils.get_ess (utils.py:8-20): ESS = 1 / sum_i W_i^2 per column, from log-weights (normalized = 0) or from normalised
 * weights (normalized = 1) */
int smcb_get_ess(const float* w_dev, int64_t n, int32_t B, int64_t stride_n, int64_t stride_b, int32_t normalized,
                 float* ess_out_dev, void* stream);
/* pyfilter.resampling.systematic (resampling.py:24-52): int64 ancestors; `u_dev` (B) overrides the sampled offsets (the
 * reference's testing hook `u=`); `normalized` mirrors the keyword of the same name */
int smcb_systematic(const float* w_dev, int64_t n, int32_t B, int64_t stride_n, int64_t stride_b, int32_t normalized,
                    const float* u_dev, uint64_t seed, int64_t* out_dev, int64_t out_stride_n, int64_t out_stride_b,
                    void* stream);
/* pyfilter.resampling.multinomial (resampling.py:55-65): `U_dev` (B, n) float64 uniforms in draw order override Philox */
int smcb_multinomial(const float* w_dev, int64_t n, int32_t B, int64_t stride_n, int64_t stride_b, int32_t normalized,
                     const double* U_dev, uint64_t seed, int64_t* out_dev, int64_t out_stride_n, int64_t out_stride_b,
                     void* stream);

/* pyfilter.resampling.residual (resampling.py:68-105) on NORMALISED weights: floor(n w) deterministic copies in particle order, then
 * multinomial draws (torch.multinomial CPU semantics, as smcb_multinomial) on the fractional parts; `U_dev` (B, n) float64 uniforms in
 * draw order override Philox.  (The reference implements one column only; columns are independent here.) */
int smcb_residual(const float* w_dev, int64_t n, int32_t B, int64_t stride_n, int64_t stride_b, const double* U_dev, uint64_t seed,
                  int64_t* out_dev, int64_t out_stride_n, int64_t out_stride_b, void* stream);
/* filters.utils.batched_gather (filters/utils.py:4-21) on the reference's layout, contiguous: out[i, b, :] = x[idx[i, b], b, :] with
 * x (n, B, D) float32 and idx (n, B) int64.  With `prev_dev` (n, B) int64 the index is first pushed one generation back and written
 * back, idx[i, b] <- prev[idx[i, b], b]: one backward step of fixed-lag smoothing (ancestral tracing, filters/particle/base.py:130-146).
 * Synchronises the stream (an index out of range is an error, as in torch.gather). */
int smcb_batched_gather(const float* x_dev, int64_t n, int32_t B, int32_t D, int64_t* idx_dev, const int64_t* prev_dev, float* out_dev,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SMCB200_H */
