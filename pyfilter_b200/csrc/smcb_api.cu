// C ABI of libsmcb200 (include/smcb200.h): handle management, kernel dispatch over the compiled model zoo, the time loop.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <string>
#include <vector>
#include <algorithm>

#include "../../include/smcb200.h"
#include "resample.cuh"
#include "step.cuh"
#include "column.cuh"
#include "move.cuh"
#include "operators.cuh"
#include "plugin.cuh"
#include "gpf.cuh"

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CU(call)                                                                                            \
  do {                                                                                                      \
    cudaError_t e_ = (call);                                                                                \
    if (e_ != cudaSuccess)                                                                                  \
      return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? SMCB_ENODEVICE : SMCB_ECUDA, \
                  std::string(#call) + ": " + cudaGetErrorString(e_));                                      \
  } while (0)

// Launch with programmatic stream serialisation (common.cuh: pdl_wait / pdl_trigger); SMCB_NO_PDL=1 falls back to plain launches.
static bool use_pdl() {
  static int v = -1;
  if (v < 0) v = getenv("SMCB_NO_PDL") ? 0 : 1;
  return v == 1;
}
template <typename K, typename A>
static void launch_pdl(K kernel, dim3 grid, dim3 block, cudaStream_t s, const A& args) {
  static bool pdl_ok = true;  // cleared when the driver refuses the attribute: plain launches from then on
  if (use_pdl() && pdl_ok) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, kernel, args) == cudaSuccess) return;
    cudaGetLastError();
    pdl_ok = false;
  }
  kernel<<<grid, block, 0, s>>>(args);
}

extern "C" int smcb_version(void) { return SMCB_VERSION; }
extern "C" int smcb_abi_signature(void) { return (int)((sizeof(smcb_config) << 16) | sizeof(smcb_info)); }
extern "C" const char* smcb_last_error(void) { return g_err.c_str(); }
extern "C" int smcb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

#ifdef SMCB_USER_MODEL_HEADER
static const int kDims[SMCB_NUM_MODELS][2] = {{1, 1}, {1, 1}, {1, 1}, {3, 2}, {UserModel::D, UserModel::OD}};
static const int kRaw[SMCB_NUM_MODELS] = {6, 6, 3, 7, UserModel::NRAW};
#define SMCB_USER_CASE(BODY) case 4: { BODY; } break;
#else
static const int kDims[SMCB_NUM_MODELS][2] = {{1, 1}, {1, 1}, {1, 1}, {3, 2}, {1, 1}};
static const int kRaw[SMCB_NUM_MODELS] = {6, 6, 3, 7, 0};
#define SMCB_USER_CASE(BODY)
#endif

// ---- host-side parameter row (layout in models.h) -------------------------------------------------------------------------
static void derive_params(int model, const double* r, float* P, const smcb_config& cfg) {
  const double c = 0.91893853320467274178;  // log sqrt(2 pi)
  for (int i = 0; i < SMCB_NPARAM; ++i) P[i] = 0.f;
  double inc = 1.0, sigma = 1.0, a = 0, s = 1;
  bool linear = false, lorenz_lgo = false;
  switch (model) {
    case SMCB_MODEL_LG_AR1:
      for (int i = 0; i < 6; ++i) P[i] = (float)r[i];
      sigma = r[2]; a = r[3]; s = r[5]; linear = true;
      P[P_X0_LOC] = (float)r[0];
      P[P_X0_SCALE] = (float)(r[2] / sqrt(1.0 - r[1] * r[1]));
      break;
    case SMCB_MODEL_SINE_EM:
      for (int i = 0; i < 6; ++i) P[i] = (float)r[i];
      sigma = r[1]; a = r[3]; s = r[5]; linear = true; inc = sqrt(r[2]);
      P[P_X0_LOC] = 0.f; P[P_X0_SCALE] = 1.f;
      break;
    case SMCB_MODEL_SV_AR1:
      sigma = r[2];
      for (int i = 0; i < 3; ++i) P[i] = (float)r[i];
      P[P_X0_LOC] = (float)r[0];
      P[P_X0_SCALE] = (float)(r[2] / sqrt(1.0 - r[1] * r[1]));
      break;
    case SMCB_MODEL_LORENZ63_EM: {
      for (int i = 0; i < 6; ++i) P[i] = (float)r[i];
      const double os = r[6];
      P[6] = (float)(1.0 / (2.0 * os * os));
      P[7] = (float)(log(os) + c);
      inc = sqrt(r[4]);
      const float m0[3] = {-5.91652f, -5.52332f, 24.5723f};
      for (int d = 0; d < 3; ++d) { P[P_X0_LOC + d] = m0[d]; P[P_X0_SCALE + d] = (float)sqrt(10.0); }
      P[P_LGO_OBS_S] = (float)os;
      // LinearGaussianObservations on y = a (x1, x3) + s nu: diagonal matrices (models.h), the unobserved coordinate keeps P = 1 / sigma^-2
      sigma = r[3]; a = r[5]; s = os; lorenz_lgo = true;
      break;
    }
#ifdef SMCB_USER_MODEL_HEADER
    case SMCB_MODEL_USER:
      P[P_INC_SCALE] = 1.f;
      UserModel::derive(r, P);
      return;
#endif
  }
  P[P_INC_SCALE] = (float)inc;
  {  // Linearized proposal: its settings and the constants of the transition density (models.h)
    const double tv = sigma * inc;
    P[P_LIN_STEPS] = (float)cfg.lin_steps; P[P_LIN_ALPHA] = cfg.lin_alpha; P[P_LIN_SECOND] = cfg.lin_second_order ? 1.f : 0.f;
    P[P_LIN_T_INVVAR] = (float)(1.0 / (tv * tv));
    P[P_LIN_T_INV2VAR] = (float)(1.0 / (2.0 * tv * tv));
    P[P_LIN_T_LOGNORM] = (float)(log(fabs(tv)) + c);
    P[P_NESTED_M] = (float)cfg.nested_samples;
  }
  if (lorenz_lgo) {
    const double hvi = 1.0 / (sigma * sigma), ovi = 1.0 / (s * s);
    const double cov = 1.0 / (hvi + a * a * ovi), cov1 = 1.0 / hvi;
    const double pre = s * s + a * a * sigma * sigma;
    P[P_LGO_HVI] = (float)hvi; P[P_LGO_OVI] = (float)ovi;
    P[P_LGO_COV] = (float)cov; P[P_LGO_KSTD] = (float)sqrt(cov);
    P[P_LGO_K_INV2VAR] = (float)(1.0 / (2.0 * cov)); P[P_LGO_K_LOGNORM] = (float)(0.5 * log(cov) + c);
    P[P_LGO_COV1] = (float)cov1; P[P_LGO_KSTD1] = (float)sqrt(cov1);
    P[P_LGO_K1_INV2VAR] = (float)(1.0 / (2.0 * cov1)); P[P_LGO_K1_LOGNORM] = (float)(0.5 * log(cov1) + c);
    P[P_LGO_PRE_INV2VAR] = (float)(1.0 / (2.0 * pre)); P[P_LGO_PRE_LOGNORM] = (float)(0.5 * log(pre) + c);
    P[P_LGO_INC_INV2VAR] = (float)(1.0 / (2.0 * inc * inc));
    P[P_LGO_INC_LOGNORM] = (float)(log(inc) + c + log(fabs(sigma)));
    P[P_LGO_INV_SIGMA] = (float)(1.0 / sigma);
  }
  if (linear) {
    P[P_OBS_INV2VAR] = (float)(1.0 / (2.0 * s * s));
    P[P_OBS_LOGNORM] = (float)(log(s) + c);
    const double hvi = 1.0 / (sigma * sigma), ovi = 1.0 / (s * s);
    const double cov = 1.0 / (hvi + a * a * ovi);
    const double pre = s * s + a * a * sigma * sigma;
    P[P_LGO_HVI] = (float)hvi;
    P[P_LGO_OVI] = (float)ovi;
    P[P_LGO_COV] = (float)cov;
    P[P_LGO_KSTD] = (float)sqrt(cov);
    P[P_LGO_K_INV2VAR] = (float)(1.0 / (2.0 * cov));
    P[P_LGO_K_LOGNORM] = (float)(0.5 * log(cov) + c);
    P[P_LGO_PRE_INV2VAR] = (float)(1.0 / (2.0 * pre));
    P[P_LGO_PRE_LOGNORM] = (float)(0.5 * log(pre) + c);
    P[P_LGO_INC_INV2VAR] = (float)(1.0 / (2.0 * inc * inc));
    P[P_LGO_INC_LOGNORM] = (float)(log(inc) + c + log(fabs(sigma)));
    P[P_LGO_INV_SIGMA] = (float)(1.0 / sigma);
  }
}

struct smcb_filter {
  smcb_config cfg;
  int D, OD, B, tiles_per_col, blocks_per_col, iters;
  int64_t n, ld;
  float* P_dev = nullptr;
  float* xbuf[2] = {nullptr, nullptr};
  float* lwbuf[2] = {nullptr, nullptr};  // log-weights and APF resampling log-weights: like the state they ping-pong with the move index
  float* rwbuf[2] = {nullptr, nullptr};  // (the state at move index t lives in xbuf / lwbuf / rwbuf [t & 1])
  int32_t *anc = nullptr, *prev_inds = nullptr;
  ColStats* stats = nullptr;
  Partial* partials = nullptr;
  Ctrl* ctrl = nullptr;
  int32_t* col_ticket = nullptr;
  double *tilesum = nullptr, *prefix = nullptr, *sin = nullptr;
  int32_t* tileflag = nullptr;
  XsDesc *desc = nullptr, *desc2 = nullptr;
  SegTable* tables = nullptr;
  uint32_t* tilemin = nullptr;
  int32_t *ncounter = nullptr, *dcounter = nullptr, *verdict = nullptr;
  FusedSlot* fslots = nullptr;
  float* u_col = nullptr;
  float *hist_mean = nullptr, *hist_var = nullptr, *hist_ll = nullptr;
  float *latest_mean = nullptr, *latest_var = nullptr, *latest_ll = nullptr, *ll_total = nullptr, *ess_packed = nullptr;
  float* y_own = nullptr;
  int y_own_cap = 0;
  float* cbuf = nullptr;  // multinomial: sequential float32 prefix sums (B, ld)
  float* midbuf = nullptr;  // multinomial: packed middle level of the draw's search (B, ceil(n / 64))
  float* wn = nullptr;    // normalised weights of the current resampling pass (B, ld)
  long long* dbg = nullptr;  // SMCB_DEBUG_TIMELINE=1: per-tile timeline of the scan kernel
  float* gpf_partial = nullptr;    // GPF: block sums of the predictive moments, (B, blocks_per_col, 10)
  float* gpf_dist = nullptr;       //      mean and Cholesky factor of the Gaussian approximation, (B, 12)
  const float* nest_z = nullptr;   // NestedProposal: injected inner draws (smcb_filter_set_nested_noise)
  const float* nest_e = nullptr;
  const float *eps_in = nullptr, *u_in = nullptr;
  const double* U_in = nullptr;
  float *eps_out = nullptr, *u_out = nullptr, *w_out = nullptr;
  int t_host = 0, y_base = 0, y_count = 0;
  const float* y_dev = nullptr;         // observations on the device (set_observations)
  unsigned long long fused_epoch = 0;   // tags the tile-sum slots of resample_fused_kernel, bumped per launch
  bool folded_for_next = false;
  int64_t launches = 0;
  // move_kernel (move.cuh): per-tile partial records, the tile-ticket counter and the host's copy of its value, the persistent grid
  Partial* tile_partials = nullptr;
  uint32_t* tile_counter = nullptr;
  long long* wd = nullptr;
  uint32_t ticket_next = 0;
  int mv_grid = 0;            // blocks of a move_kernel launch = B * mv_T (0: geometry not chosen yet)
  int mv_t1 = 0, mv_items1 = 16, mv_items2 = 16, mv_T = 0;   // tile classes of a column (move.cuh, MoveArgs)
  int mv_tiles_cap = 0;       // per-column capacity of the per-tile buffers (tiles of the smallest class)
  unsigned long long *mslots = nullptr, *gwords = nullptr;
  unsigned mv_epoch = 0;
  uint32_t mv_launches = 0;   // move_kernel launches since the group counters were cleared
  // peer-memory exchange of the per-column log-likelihoods (smcb_filter_attach_exchange)
  ExchangeArgs xch = {};
  uint32_t xch_seq = 0;       // exchanges published so far
  int xch_rank = 0;
  float* xch_out = nullptr;   // (2, total) dense copy made by smcb_filter_exchange_wait
  int32_t* xch_ticket = nullptr;   // column kernel: elects the block that gathers the exchange inside the move's own launch
  uint32_t xch_inline_seq = 0;     // the exchange that launch has already gathered into xch_out
  bool want_inline = false;        // set by smcb_filter_run_stepwise: the ranks run concurrently, a block may wait for the peers' values
};

template <typename T>
static cudaError_t dalloc(T** p, size_t count) {
  cudaError_t e = cudaMalloc((void**)p, count * sizeof(T));
  if (e == cudaSuccess) e = cudaMemset(*p, 0, count * sizeof(T));
  return e;
}

static int upload_params(smcb_filter* f, const float* params_host, int n_raw, int cols, cudaStream_t s) {
  if (n_raw != kRaw[f->cfg.model]) return fail(SMCB_EINVAL, "wrong number of raw parameters for this model");
  if (cols != 1 && cols != f->B) return fail(SMCB_EINVAL, "param_cols must be 1 or the batch size");
  std::vector<float> P((size_t)f->B * SMCB_NPARAM);
  for (int b = 0; b < f->B; ++b) {
    double r[SMCB_MAX_RAW_PARAMS];
    for (int k = 0; k < n_raw; ++k) r[k] = params_host[(size_t)k * cols + (cols == 1 ? 0 : b)];
    derive_params(f->cfg.model, r, &P[(size_t)b * SMCB_NPARAM], f->cfg);
  }
  CU(cudaMemcpyAsync(f->P_dev, P.data(), P.size() * sizeof(float), cudaMemcpyHostToDevice, s));
  CU(cudaStreamSynchronize(s));  // P is a stack/heap temporary
  return SMCB_OK;
}

extern "C" int smcb_filter_destroy(smcb_filter* f) {
  if (!f) return SMCB_OK;
  void* ptrs[] = {f->P_dev, f->xbuf[0], f->xbuf[1], f->lwbuf[0], f->lwbuf[1], f->rwbuf[0], f->rwbuf[1], f->anc, f->prev_inds, f->stats, f->partials, f->ctrl,
                  f->tilesum, f->prefix, f->sin, f->tileflag, f->desc, f->desc2, f->tables, f->dcounter, f->hist_mean, f->hist_var, f->hist_ll, f->latest_mean, f->latest_var, f->latest_ll,
                  f->ll_total, f->ess_packed, f->y_own, f->cbuf, f->col_ticket, f->wn, f->dbg, f->tilemin, f->ncounter, f->verdict, f->fslots,
                  f->u_col, f->tile_partials, f->tile_counter, f->wd, f->mslots, f->gwords, f->xch_out, f->xch_ticket, f->midbuf, f->gpf_partial, f->gpf_dist};
  for (void* p : ptrs) if (p) cudaFree(p);
  delete f;
  return SMCB_OK;
}

extern "C" int smcb_filter_create(const smcb_config* cfg, smcb_filter** out) {
  if (!cfg || !out) return fail(SMCB_EINVAL, "null argument");
  if (cfg->model < 0 || cfg->model >= SMCB_NUM_MODELS) return fail(SMCB_EUNSUPPORTED, "unknown model id (only the compiled zoo is supported; there is no CPU fallback)");
#ifndef SMCB_USER_MODEL_HEADER
  if (cfg->model == SMCB_MODEL_USER) return fail(SMCB_EUNSUPPORTED, "this build of libsmcb200 carries no user model (pyfilter_b200.timeseries.compile_user_model builds one)");
#endif
  if (cfg->model == SMCB_MODEL_USER && cfg->proposal != SMCB_BOOTSTRAP) return fail(SMCB_EUNSUPPORTED, "Model combination not supported!");
  if (cfg->proposal != SMCB_BOOTSTRAP && cfg->proposal != SMCB_LINEAR_GAUSSIAN_OBSERVATIONS && cfg->proposal != SMCB_LINEARIZED &&
      cfg->proposal != SMCB_NESTED)
    return fail(SMCB_EUNSUPPORTED, "unknown proposal");
  if (cfg->proposal == SMCB_NESTED && (cfg->nested_samples < 1 || cfg->nested_samples > SMCB_NESTED_MAX))
    return fail(SMCB_EINVAL, "NestedProposal: num_samples must be in 1 .. 256");
  if (cfg->proposal == SMCB_LINEARIZED && cfg->lin_steps < 1) return fail(SMCB_EINVAL, "``n_steps`` must be >= 1");   // proposals/linearized.py:39
  if (cfg->proposal == SMCB_LINEAR_GAUSSIAN_OBSERVATIONS && !(cfg->model == SMCB_LG_AR1 || cfg->model == SMCB_SINE_EM || cfg->model == SMCB_LORENZ63_EM))
    return fail(SMCB_EUNSUPPORTED, "Model combination not supported!");  // same condition as proposals/linear.py:32-36
  if (cfg->algorithm != SMCB_SISR && cfg->algorithm != SMCB_APF && cfg->algorithm != SMCB_GPF) return fail(SMCB_EUNSUPPORTED, "unknown filter algorithm");
  if (cfg->algorithm == SMCB_GPF && cfg->proposal != SMCB_BOOTSTRAP)   // the GaussianProposal of gpf.py:24 is the one compiled
    return fail(SMCB_EUNSUPPORTED, "GPF runs with its default GaussianProposal");
  if (cfg->resampler != SMCB_SYSTEMATIC && cfg->resampler != SMCB_MULTINOMIAL) return fail(SMCB_EUNSUPPORTED, "unknown resampler");
  if (cfg->particles < 1 || cfg->particles >= ((int64_t)1 << 31) - (1 << 21)) return fail(SMCB_EINVAL, "particles out of range");
  if (cfg->resampler == SMCB_MULTINOMIAL && cfg->particles > (1 << 24)) return fail(SMCB_EINVAL, "number of categories cannot exceed 2^24");  // torch.multinomial's limit
  if (cfg->batch < 1) return fail(SMCB_EINVAL, "batch must be >= 1");
  if (smcb_device_count() < 1) return fail(SMCB_ENODEVICE, "no CUDA device: libsmcb200 has no CPU fallback");
  smcb_filter* f = new smcb_filter();
  f->cfg = *cfg;
  f->cfg.params_host = nullptr;
  f->D = kDims[cfg->model][0];
  f->OD = kDims[cfg->model][1];
  f->B = cfg->batch;
  f->n = cfg->particles;
  f->tiles_per_col = (int)((f->n + RS_TILE - 1) / RS_TILE);
  f->ld = (int64_t)f->tiles_per_col * RS_TILE;
  {
    const int64_t chunk = ST_NT * ST_VEC;
    const int64_t nchunks = (f->n + chunk - 1) / chunk;
    // one resident wave of step-kernel blocks PER COLUMN: how a column is cut into blocks (hence the order in which its partial sums are
    // folded) must not depend on how many columns the handle holds - shards of one batch reproduce the unsharded run bit for bit
    int64_t cap = (int64_t)smcb_sm_count() * SMCB_ST_MINB;
    if (cap < 1) cap = 1;
    f->iters = (int)((nchunks + cap - 1) / cap);
    f->blocks_per_col = (int)((nchunks + f->iters - 1) / f->iters);
  }
  const size_t cells = (size_t)f->B * f->ld;
  const size_t slack = RS_TILE;   // the last tile of a column may read past its row (move.cuh): one tile of slack behind the last row
  f->mv_tiles_cap = (int)((f->n + MV_NT * 4 - 1) / (MV_NT * 4)) + 1;
  const int rows = cfg->history_rows > 0 ? cfg->history_rows : 1;
  f->cfg.history_rows = rows;
  cudaError_t e = cudaSuccess;
#define A_(call) if (e == cudaSuccess) e = (call)
  A_(dalloc(&f->P_dev, (size_t)f->B * SMCB_NPARAM));
  A_(dalloc(&f->xbuf[0], cells * f->D + slack));
  A_(dalloc(&f->xbuf[1], cells * f->D + slack));
  A_(dalloc(&f->lwbuf[0], cells + slack));
  A_(dalloc(&f->lwbuf[1], cells + slack));
  A_(dalloc(&f->rwbuf[0], cells + slack));
  A_(dalloc(&f->rwbuf[1], cells + slack));
  A_(dalloc(&f->wn, cells));
  A_(dalloc(&f->anc, cells));
  A_(dalloc(&f->prev_inds, cells));
  A_(dalloc(&f->stats, (size_t)f->B));
  A_(dalloc(&f->partials, (size_t)f->B * f->blocks_per_col));
  A_(dalloc(&f->ctrl, (size_t)1));
  A_(dalloc(&f->col_ticket, (size_t)f->B));
  A_(dalloc(&f->tilesum, (size_t)f->B * f->tiles_per_col));
  A_(dalloc(&f->prefix, (size_t)f->B * f->tiles_per_col));
  A_(dalloc(&f->sin, (size_t)f->B * f->tiles_per_col));
  A_(dalloc(&f->tileflag, (size_t)f->B * f->tiles_per_col));
  A_(dalloc(&f->desc, (size_t)f->B * f->tiles_per_col));
  A_(dalloc(&f->desc2, (size_t)f->B * f->tiles_per_col));
  A_(dalloc(&f->tables, (size_t)f->B * f->tiles_per_col));
  A_(dalloc(&f->dcounter, (size_t)f->B));
  A_(dalloc(&f->fslots, (size_t)f->B * f->tiles_per_col));
  A_(dalloc(&f->tilemin, (size_t)f->B * f->tiles_per_col));
  A_(dalloc(&f->ncounter, (size_t)f->B));
  A_(dalloc(&f->verdict, (size_t)f->B));
  A_(dalloc(&f->u_col, (size_t)f->B));
  A_(dalloc(&f->tile_partials, (size_t)f->B * f->mv_tiles_cap));
  A_(dalloc(&f->tile_counter, (size_t)1));
  A_(dalloc(&f->mslots, (size_t)f->B * f->mv_tiles_cap));
  A_(dalloc(&f->gwords, (size_t)2 * f->B * ((f->mv_tiles_cap + MV_GROUP - 1) / MV_GROUP) * MV_GPAD));
  A_(dalloc(&f->wd, (size_t)4));
  A_(dalloc(&f->hist_mean, (size_t)rows * f->B * f->D));
  A_(dalloc(&f->hist_var, (size_t)rows * f->B * f->D));
  A_(dalloc(&f->hist_ll, (size_t)rows * f->B));
  A_(dalloc(&f->latest_mean, (size_t)f->B * f->D));
  A_(dalloc(&f->latest_var, (size_t)f->B * f->D));
  A_(dalloc(&f->latest_ll, (size_t)f->B));
  A_(dalloc(&f->ll_total, (size_t)f->B));
  A_(dalloc(&f->ess_packed, (size_t)f->B * 2));
  if (cfg->algorithm == SMCB_GPF) {
    A_(dalloc(&f->gpf_partial, (size_t)f->B * f->blocks_per_col * 10));
    A_(dalloc(&f->gpf_dist, (size_t)f->B * 12));
    f->cfg.ess_threshold = 0.f;   // nothing resamples (gpf.py:27-30): the propagate-only moves never see a resampling flag
  }
  if (cfg->resampler == SMCB_MULTINOMIAL) A_(dalloc(&f->cbuf, cells));
  if (cfg->resampler == SMCB_MULTINOMIAL) A_(dalloc(&f->midbuf, (size_t)f->B * ((f->n + MN_MID - 1) / MN_MID)));
  if (getenv("SMCB_DEBUG_TIMELINE")) A_(dalloc(&f->dbg, (size_t)32 + (size_t)16 * f->B * (f->mv_tiles_cap > f->tiles_per_col ? f->mv_tiles_cap : f->tiles_per_col)));
#undef A_
  if (e != cudaSuccess) {
    smcb_filter_destroy(f);
    return fail(SMCB_ECUDA, std::string("device allocation failed: ") + cudaGetErrorString(e));
  }
  int rc = upload_params(f, cfg->params_host, cfg->n_raw_params, cfg->param_cols, 0);
  if (rc != SMCB_OK) { smcb_filter_destroy(f); return rc; }
  *out = f;
  return SMCB_OK;
}

extern "C" int smcb_filter_set_params(smcb_filter* f, const float* params_host, int32_t n_raw, int32_t cols, void* stream) {
  if (!f || !params_host) return fail(SMCB_EINVAL, "null argument");
  return upload_params(f, params_host, n_raw, cols, (cudaStream_t)stream);
}

extern "C" int smcb_filter_set_seed(smcb_filter* f, uint64_t seed) {
  if (!f) return fail(SMCB_EINVAL, "null handle");
  f->cfg.seed = seed;
  return SMCB_OK;
}

extern "C" int smcb_filter_info(smcb_filter* f, smcb_info* o) {
  if (!f || !o) return fail(SMCB_EINVAL, "null argument");
  o->particles = f->n; o->ld = f->ld; o->batch = f->B; o->state_dim = f->D; o->obs_dim = f->OD;
  o->t = f->t_host; o->history_rows = f->cfg.history_rows; o->kernel_launches = f->launches;
  Ctrl c;
  CU(cudaMemcpy(&c, f->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost));
  o->slow_tiles = c.slow_tiles; o->lb_fail = 0; o->lb_windows = 0; o->reserved = 0;
  return SMCB_OK;
}

// ---- kernel dispatch ------------------------------------------------------------------------------------------------------------
static StepArgs make_args(smcb_filter* f) {
  StepArgs a;
  memset(&a, 0, sizeof(a));
  a.n = f->n; a.ld = f->ld; a.B = f->B; a.blocks_per_col = f->blocks_per_col; a.iters = f->iters;
  a.P = f->P_dev; a.xbuf[0] = f->xbuf[0]; a.xbuf[1] = f->xbuf[1];
  a.lw = f->lwbuf[f->t_host & 1]; a.rw = f->rwbuf[f->t_host & 1];                  // current state
  a.lw_out = f->lwbuf[(f->t_host + 1) & 1]; a.rw_out = f->rwbuf[(f->t_host + 1) & 1];  // state after ONE move (run_column overrides)
  a.anc = f->anc; a.prev_inds = f->prev_inds; a.stats = f->stats; a.partials = f->partials; a.ctrl = f->ctrl; a.col_ticket = f->col_ticket;
  a.eps_in = f->eps_in; a.eps_out = f->eps_out; a.seed = f->cfg.seed;
  philox_round_keys((uint32_t)f->cfg.seed, (uint32_t)(f->cfg.seed >> 32), a.pkeys);
  a.fold = f->cfg.fold_lookahead; a.store_lw = 1; a.sample_x0 = 0; a.ess_threshold = f->cfg.ess_threshold;
  a.col0 = f->cfg.column_offset;
  a.nest_z = f->nest_z; a.nest_e = f->nest_e;
  a.xch = f->xch; a.xch.seq = 0;   // set by the launch that finalises the last move of a run
  a.hist_mean = f->hist_mean; a.hist_var = f->hist_var; a.hist_ll = f->hist_ll; a.hist_rows = f->cfg.history_rows;
  a.dbg = f->dbg;
  a.latest_mean = f->latest_mean; a.latest_var = f->latest_var; a.latest_ll = f->latest_ll; a.ll_total = f->ll_total;
  return a;
}

#define FOR_MODEL(M, BODY)                                                  \
  switch (M) {                                                              \
    case 0: { constexpr int MODEL = 0; BODY; } break;                        \
    case 1: { constexpr int MODEL = 1; BODY; } break;                        \
    case 2: { constexpr int MODEL = 2; BODY; } break;                        \
    case 3: { constexpr int MODEL = 3; BODY; } break;                        \
    SMCB_USER_CASE(constexpr int MODEL = 4; BODY)                            \
  }

// every compiled (model, proposal) pair: LinearGaussianObservations needs linear-Gaussian observations, a user model takes Bootstrap
#ifdef SMCB_USER_MODEL_HEADER
#define SMCB_USER_PAIR(BODY) case 16: { constexpr int MODEL = 4, PROP = 0; BODY; } break;
#else
#define SMCB_USER_PAIR(BODY)
#endif
#define FOR_MODEL_PROP(M, PR, BODY)                                                   \
  switch ((M) * 4 + (PR)) {                                                           \
    case 0: { constexpr int MODEL = 0, PROP = 0; BODY; } break;                        \
    case 1: { constexpr int MODEL = 0, PROP = 1; BODY; } break;                        \
    case 2: { constexpr int MODEL = 0, PROP = 2; BODY; } break;                        \
    case 3: { constexpr int MODEL = 0, PROP = 3; BODY; } break;                        \
    case 4: { constexpr int MODEL = 1, PROP = 0; BODY; } break;                        \
    case 5: { constexpr int MODEL = 1, PROP = 1; BODY; } break;                        \
    case 6: { constexpr int MODEL = 1, PROP = 2; BODY; } break;                        \
    case 7: { constexpr int MODEL = 1, PROP = 3; BODY; } break;                        \
    case 8: { constexpr int MODEL = 2, PROP = 0; BODY; } break;                        \
    case 10: { constexpr int MODEL = 2, PROP = 2; BODY; } break;                       \
    case 11: { constexpr int MODEL = 2, PROP = 3; BODY; } break;                       \
    case 12: { constexpr int MODEL = 3, PROP = 0; BODY; } break;                       \
    case 13: { constexpr int MODEL = 3, PROP = 1; BODY; } break;                       \
    case 14: { constexpr int MODEL = 3, PROP = 2; BODY; } break;                       \
    case 15: { constexpr int MODEL = 3, PROP = 3; BODY; } break;                       \
    SMCB_USER_PAIR(BODY)                                                              \
  }

template <int MODEL, int PROP>
static void launch_step_alg(int alg, dim3 g, cudaStream_t s, const StepArgs& a) {
  if (alg == SMCB_SISR) launch_pdl(step_kernel<MODEL, PROP, SMCB_ALG_SISR>, g, dim3(ST_NT), s, a);
  else launch_pdl(step_kernel<MODEL, PROP, SMCB_ALG_APF>, g, dim3(ST_NT), s, a);
}

static void launch_step(smcb_filter* f, const StepArgs& a, cudaStream_t s) {
  dim3 g(f->blocks_per_col, f->B);
  const int prop = f->cfg.proposal, alg = f->cfg.algorithm;
  FOR_MODEL_PROP(f->cfg.model, prop, (launch_step_alg<MODEL, PROP>(alg, g, s, a)));
  f->launches++;
}

// ---- move_kernel (move.cuh): one kernel per move, one block per tile, tiles drawn from a ticket counter ---------------------------
// Tile classes of a column (MoveArgs.t1 / items1 / items2).  `slots` = blocks resident at a time.  Tiles of 4096 particles (16 per
// thread) unless the last wave of such tiles would use less than a quarter of the slots: then the columns' remainders behind the full
// waves are cut into 1024-particle tiles (measurements: move.cuh, profiles/README.md).  SMCB_MV_GEOM="items1,items2,t1" overrides
// (diagnostics, tools/geom_sweep.py).
static void mv_choose_geometry(smcb_filter* f, int slots) {
  const int64_t n = f->n;
  const int B = f->B;
  const int64_t L1 = (int64_t)MV_NT * 16;
  const int64_t Tu = (n + L1 - 1) / L1;
  int bt1 = (int)Tu, bi1 = 16, bi2 = 16, bT = (int)Tu;
  const int64_t total = Tu * B;
  // (a single filter only: the cut of a column must not depend on how many columns the handle holds - shards of one batch on
  // different ranks reproduce the unsharded run bit for bit, and the fold of a column's per-tile records follows its tiles)
  if (B == 1 && total > slots && total % slots != 0 && total % slots < slots / 4) {
    const int64_t t1 = ((total / slots) * slots) / B;   // whole waves of large tiles
    if (t1 >= 1 && t1 < Tu) {
      const int64_t L2 = (int64_t)MV_NT * 4;
      const int64_t t2 = (n - t1 * L1 + L2 - 1) / L2;
      if (t2 * B <= slots && t1 + t2 <= f->mv_tiles_cap) { bt1 = (int)t1; bi2 = 4; bT = (int)(t1 + t2); }
    }
  }
  if (const char* g = getenv("SMCB_MV_GEOM")) {
    int i1 = 16, i2 = 16, t1 = 0;
    if (sscanf(g, "%d,%d,%d", &i1, &i2, &t1) == 3 && (i1 == 16 || i1 == 4) && (i2 == 16 || i2 == 4)) {
      const int64_t L1 = (int64_t)MV_NT * i1, L2 = (int64_t)MV_NT * i2;
      int64_t T1 = t1 < 0 ? 0 : t1;
      if (T1 * L1 >= n) T1 = (n + L1 - 1) / L1;
      const int64_t rest = n - T1 * L1;
      const int64_t T2 = rest > 0 ? (rest + L2 - 1) / L2 : 0;
      if (T1 + T2 <= f->mv_tiles_cap) { bt1 = (int)T1; bi1 = i1; bi2 = i2; bT = (int)(T1 + T2); }
    }
  }
  f->mv_t1 = bt1; f->mv_items1 = bi1; f->mv_items2 = bi2; f->mv_T = bT;
}

template <int MODEL, int PROP, int ALG>
static cudaError_t launch_move_t(smcb_filter* f, MoveArgs& m, cudaStream_t s) {
  constexpr int D = Model<MODEL>::D;
  const size_t dyn = sizeof(float) * (size_t)D * MV_TILE;
  auto kernel = move_kernel<MODEL, PROP, ALG>;
  if (f->mv_grid == 0) {  // once per handle: dynamic shared memory limit, resident blocks, tile classes; one block per tile
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, MV_NT, dyn);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    mv_choose_geometry(f, per_sm * smcb_sm_count());
    f->mv_grid = f->mv_T * f->B;
  }
  m.t1 = f->mv_t1; m.items1 = f->mv_items1; m.items2 = f->mv_items2;
  m.tiles_per_col = f->mv_T; m.total_tiles = f->mv_grid;
  m.s.blocks_per_col = f->mv_T;
  m.ticket_base = f->ticket_next;
  f->ticket_next += (uint32_t)m.total_tiles;  // every block draws exactly one ticket
  static bool pdl_ok = true;
  if (use_pdl() && pdl_ok) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(f->mv_grid); cfg.blockDim = dim3(MV_NT); cfg.dynamicSmemBytes = dyn; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, kernel, m) == cudaSuccess) return cudaSuccess;
    cudaGetLastError();
    pdl_ok = false;
  }
  kernel<<<f->mv_grid, MV_NT, dyn, s>>>(m);
  return cudaGetLastError();
}
template <int MODEL, int PROP>
static cudaError_t launch_move_alg(smcb_filter* f, MoveArgs& m, cudaStream_t s) {
  if (f->cfg.algorithm == SMCB_SISR) return launch_move_t<MODEL, PROP, SMCB_ALG_SISR>(f, m, s);
  return launch_move_t<MODEL, PROP, SMCB_ALG_APF>(f, m, s);
}
static int launch_move(smcb_filter* f, const StepArgs& a, cudaStream_t s) {
  MoveArgs m;
  memset(&m, 0, sizeof(m));
  m.s = a;
  m.s.partials = f->tile_partials;   // (tile geometry: launch_move_t)
  f->mv_epoch = f->mv_epoch % 1023u + 1u;   // 1 .. 1023, never the previous launch's
  m.tile_counter = f->tile_counter; m.mslots = f->mslots; m.gwords = f->gwords; m.epoch = f->mv_epoch;
  m.launch_index = f->mv_launches++;
  m.u_in = f->u_in; m.u_out = f->u_out; m.w_out = f->w_out; m.wd = f->wd;
  m.tl = f->dbg ? f->dbg + 32 : nullptr;
  if (!f->u_in && f->B <= MV_U_HOST) {  // few columns: the systematic offsets (one Philox block per column and move) come as arguments
    m.n_u_host = f->B;
    for (int b = 0; b < f->B; ++b) {
      const Philox4 r4 = philox4x32_10((uint32_t)(b + a.col0), 0u, (uint32_t)a.t_host, SMCB_RNG_SYSTEMATIC, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
      m.u_host[b] = smcb_u01(r4.x);
    }
  }
  const int prop = f->cfg.proposal;
  cudaError_t e = cudaSuccess;
  FOR_MODEL_PROP(f->cfg.model, prop, (e = launch_move_alg<MODEL, PROP>(f, m, s)));
  if (e != cudaSuccess) return fail(SMCB_ECUDA, std::string("move_kernel: ") + cudaGetErrorString(e));
  f->launches++;
  return SMCB_OK;
}
// the move kernel serves systematic resampling with rounding-free weights of at most 2^23 particles (the lean probe count)
static bool move_path_ok(const smcb_filter* f) {
  const bool off = getenv("SMCB_NO_MOVE") != nullptr;  // diagnostics / tests: force the two-kernel pipeline
  return !off && f->cfg.algorithm != SMCB_GPF && f->cfg.resampler == SMCB_SYSTEMATIC && !f->cfg.exact_weights && f->n <= (1 << 23) &&
         (int64_t)f->mv_tiles_cap * f->B < (1 << 30);
}

static void launch_preweight(smcb_filter* f, const StepArgs& a, cudaStream_t s) {
  dim3 g(f->blocks_per_col, f->B);
  const int prop = f->cfg.proposal;
  FOR_MODEL_PROP(f->cfg.model, prop, (preweight_kernel<MODEL, PROP><<<g, ST_NT, 0, s>>>(a)));
  f->launches++;
}

static void launch_state(smcb_filter* f, const StepArgs& a, cudaStream_t s) {
  dim3 g(f->blocks_per_col, f->B);
  FOR_MODEL(f->cfg.model, (state_kernel<MODEL><<<g, ST_NT, 0, s>>>(a)));
  f->launches++;
}

static void launch_finalize(smcb_filter* f, StepArgs a, int mode, cudaStream_t s, bool pdl = false) {
  a.fin_mode = mode;
  const bool apf = f->cfg.algorithm == SMCB_APF;
  const dim3 g(f->B), b(ST_NT);
  if (pdl) {  // behind move_kernel: overlaps the launch with the last tiles of the move
    if (f->D == 1) {
      if (apf) launch_pdl(finalize_kernel<1, 1, SMCB_ALG_APF>, g, b, s, a);
      else launch_pdl(finalize_kernel<1, 1, SMCB_ALG_SISR>, g, b, s, a);
    } else {
      if (apf) launch_pdl(finalize_kernel<3, 2, SMCB_ALG_APF>, g, b, s, a);
      else launch_pdl(finalize_kernel<3, 2, SMCB_ALG_SISR>, g, b, s, a);
    }
  } else if (f->D == 1) {
    if (apf) finalize_kernel<1, 1, SMCB_ALG_APF><<<g, b, 0, s>>>(a);
    else finalize_kernel<1, 1, SMCB_ALG_SISR><<<g, b, 0, s>>>(a);
  } else {
    if (apf) finalize_kernel<3, 2, SMCB_ALG_APF><<<g, b, 0, s>>>(a);
    else finalize_kernel<3, 2, SMCB_ALG_SISR><<<g, b, 0, s>>>(a);
  }
  f->launches++;
}

static int push_ctrl(smcb_filter* f, const float* y_dev, cudaStream_t s) {
  // only the fields the host owns: t, y_base, y_count, y (the device owns ticket/epoch/tile_counter/slow_tiles)
  struct { int32_t t, y_base, y_count; } head = {f->t_host, f->y_base, f->y_count};
  CU(cudaMemcpyAsync(f->ctrl, &head, sizeof(head), cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync((char*)f->ctrl + offsetof(Ctrl, y), &y_dev, sizeof(y_dev), cudaMemcpyHostToDevice, s));
  f->y_dev = y_dev;
  return SMCB_OK;
}

extern "C" int smcb_filter_initialize(smcb_filter* f, void* stream) {
  if (!f) return fail(SMCB_EINVAL, "null handle");
  cudaStream_t s = (cudaStream_t)stream;
  f->t_host = 0; f->folded_for_next = false;
  const float* ynull = nullptr;
  f->y_count = 0; f->y_base = 0;
  int rc = push_ctrl(f, ynull, s);
  if (rc) return rc;
  CU(cudaMemsetAsync(f->ll_total, 0, sizeof(float) * f->B, s));
  CU(cudaMemsetAsync(f->stats, 0, sizeof(ColStats) * f->B, s));
  CU(cudaMemsetAsync(f->gwords, 0, sizeof(unsigned long long) * 2 * f->B * ((f->mv_tiles_cap + MV_GROUP - 1) / MV_GROUP) * MV_GPAD, s));
  f->mv_launches = 0;
  StepArgs a = make_args(f);
  a.sample_x0 = 1;
  launch_state(f, a, s);
  launch_finalize(f, a, FIN_STATE, s);
  CU(cudaGetLastError());
  return SMCB_OK;
}

extern "C" int smcb_filter_refresh_state(smcb_filter* f, int32_t t, void* stream) {
  if (!f) return fail(SMCB_EINVAL, "null handle");
  cudaStream_t s = (cudaStream_t)stream;
  f->t_host = t; f->folded_for_next = false;
  struct { int32_t t; } head = {t};
  CU(cudaMemcpyAsync(f->ctrl, &head, sizeof(head), cudaMemcpyHostToDevice, s));
  StepArgs a = make_args(f);
  a.sample_x0 = 0;
  launch_state(f, a, s);
  launch_finalize(f, a, FIN_STATE, s);
  CU(cudaGetLastError());
  return SMCB_OK;
}

extern "C" int smcb_filter_set_observations(smcb_filter* f, const float* y_dev, int32_t count, int32_t base_t, void* stream) {
  if (!f) return fail(SMCB_EINVAL, "null handle");
  if (count < 0) return fail(SMCB_EINVAL, "negative count");
  f->y_base = base_t; f->y_count = count;
  return push_ctrl(f, y_dev, (cudaStream_t)stream);
}

// one filter move; `ev` (optional) receives 6 events bracketing the 5 kernel groups
// `last`: the caller reads the state after this move (API-visible log-weights and ancestors are stored)
static int run_one(smcb_filter* f, cudaStream_t s, cudaEvent_t* ev, bool last = true) {
  const bool apf = f->cfg.algorithm == SMCB_APF;
  const int t = f->t_host;
  if (t - f->y_base < 0 || t - f->y_base >= f->y_count) return fail(SMCB_ESTATE, "no observation set for this move");
  StepArgs a = make_args(f);
  a.t_host = t;
  a.y_t = f->y_dev ? f->y_dev + (int64_t)(t - f->y_base) * f->OD : nullptr;
  a.y_next = (f->y_dev && t + 1 - f->y_base < f->y_count) ? f->y_dev + (int64_t)(t + 1 - f->y_base) * f->OD : nullptr;
  if (f->cfg.algorithm == SMCB_GPF) {   // gpf.cuh: propagate + predictive moments, Gaussian fit, draw + weight, fold
    GpfArgs g;
    g.s = a; g.partial = f->gpf_partial; g.dist = f->gpf_dist;
    g.s.fin_host = 1;
    const dim3 grid(f->blocks_per_col, f->B);
    FOR_MODEL(f->cfg.model, (gpf_predict_kernel<MODEL><<<grid, ST_NT, 0, s>>>(g)));
    if (f->D == 1) gpf_moments_kernel<1><<<f->B, ST_NT, 0, s>>>(g);
    else if (f->D == 2) gpf_moments_kernel<2><<<f->B, ST_NT, 0, s>>>(g);
    else gpf_moments_kernel<3><<<f->B, ST_NT, 0, s>>>(g);
    FOR_MODEL(f->cfg.model, (gpf_correct_kernel<MODEL><<<grid, ST_NT, 0, s>>>(g)));
    f->launches += 3;
    launch_finalize(f, g.s, FIN_STEP, s);
    if (ev) for (int k = 0; k <= 5; ++k) cudaEventRecord(ev[k], s);
    f->t_host = t + 1;
    return SMCB_OK;
  }
  if (ev) cudaEventRecord(ev[0], s);
  if (apf && !f->folded_for_next) {  // apf.py:27-29 evaluated now because the previous move could not fold it
    launch_preweight(f, a, s);
    launch_finalize(f, a, FIN_PREWEIGHT, s);
  }
  if (ev) cudaEventRecord(ev[1], s);
  a.store_lw = last ? 1 : 0;
  if (last && f->xch.peer[0]) a.xch.seq = ++f->xch_seq;   // the move that ends a run publishes its columns' log-likelihoods to every rank
  if (move_path_ok(f)) {
    int rc = launch_move(f, a, s);
    if (rc) return rc;
    {  // the per-tile records of the move are folded by a one-block-per-column kernel chained behind it
      StepArgs fa = a;
      fa.partials = f->tile_partials; fa.blocks_per_col = f->mv_T; fa.fin_host = 1;
      launch_finalize(f, fa, FIN_STEP, s, true);
    }
    if (ev) for (int g = 2; g <= 5; ++g) cudaEventRecord(ev[g], s);
    f->folded_for_next = apf && f->cfg.fold_lookahead && (t + 1 - f->y_base) < f->y_count;
    f->t_host = t + 1;
    return SMCB_OK;
  }
  ResampleArgs r;
  memset(&r, 0, sizeof(r));
  r.w = apf ? a.rw : a.lw;
  r.wn = f->wn;
  r.n = f->n; r.ld = f->ld; r.B = f->B; r.tiles_per_col = f->tiles_per_col;
  r.input_is_w = 0; r.use_rw = apf ? 1 : 0; r.stats = f->stats;
  r.u_in = f->u_in; r.u_out = f->u_out; r.seed = f->cfg.seed; r.col0 = f->cfg.column_offset;
  r.tilesum = f->tilesum; r.prefix = f->prefix; r.sin = f->sin; r.tileflag = f->tileflag; r.desc = f->desc; r.desc2 = f->desc2; r.tables = f->tables;
  r.anc = f->anc; r.w_out = f->w_out; r.ctrl = f->ctrl;
  r.tilemin = f->tilemin; r.ncounter = f->ncounter; r.dcounter = f->dcounter; r.verdict = f->verdict; r.u_col = f->u_col;
  r.dbg = f->dbg;
  r.quantize = f->cfg.exact_weights ? 0 : 1;
  r.presanitized = 1;  // lw / rw were stored by the step, state and pre-weight kernels
  const dim3 rgrid(r.tiles_per_col, r.B);
  // rounding-free weights of at most 2^23 particles with Philox offsets and no dump of the weights' prefix: one fused kernel
  static const bool no_fused = getenv("SMCB_NO_FUSED") != nullptr;
  if (f->cfg.resampler == SMCB_SYSTEMATIC && r.quantize && f->n <= (1 << 23) && !f->u_in && !no_fused) {
    r.fslots = f->fslots;
    r.t_host = t; r.epoch_host = ++f->fused_epoch;
    launch_pdl(resample_fused_kernel, dim3(r.tiles_per_col * r.B), dim3(RS_NT), s, r);
    f->launches++;
    if (ev) { cudaEventRecord(ev[2], s); cudaEventRecord(ev[3], s); cudaEventRecord(ev[4], s); }
    launch_step(f, a, s);
    if (ev) cudaEventRecord(ev[5], s);
    f->folded_for_next = apf && f->cfg.fold_lookahead && (t + 1 - f->y_base) < f->y_count;
    f->t_host = t + 1;
    return SMCB_OK;
  }
  launch_pdl(normalize_kernel, rgrid, dim3(RS_NT), s, r);
  f->launches++;
  if (ev) cudaEventRecord(ev[2], s);
  if (f->cfg.resampler == SMCB_SYSTEMATIC) {
    // quantised weights of at most 2^23 particles with Philox offsets are benign by construction: nothing to describe or chain
    r.force_benign = (r.quantize && f->n <= (1 << 23) && !f->u_in) ? 1 : 0;
    if (!r.force_benign) {
      launch_pdl(describe_kernel<53>, rgrid, dim3(RS_NT), s, r);  // returns at once for benign columns (verdict decided on the device)
      f->launches++;
    }
    if (ev) cudaEventRecord(ev[3], s);
    launch_pdl(expand_kernel<53, RS_OUT_ANCESTORS>, rgrid, dim3(RS_NT), s, r);
    f->launches++;
  } else {
    r.c_out = f->cbuf; r.mid_out = f->midbuf;
    if (ev) cudaEventRecord(ev[3], s);
    op_launch_multinomial_after_normalize(r, f->U_in, f->ld, s);
    f->launches += 4;
  }
  if (ev) cudaEventRecord(ev[4], s);
  launch_step(f, a, s);  // the block that completes a column also folds its partials (finalize_column, FIN_STEP)
  if (ev) cudaEventRecord(ev[5], s);
  f->folded_for_next = apf && f->cfg.fold_lookahead && (t + 1 - f->y_base) < f->y_count;
  f->t_host = t + 1;
  return SMCB_OK;
}

// ---- resident-column path: one block per column runs all `steps` moves in one launch (column.cuh) -----------------------------------
static bool column_path_ok(const smcb_filter* f) {
  if (getenv("SMCB_NO_COLUMN")) return false;   // diagnostics / tests: force the multi-kernel pipeline
  if (f->cfg.algorithm == SMCB_GPF) return false;
  if (f->n > RS_TILE || f->cfg.resampler != SMCB_SYSTEMATIC || f->cfg.exact_weights) return false;
  if (f->cfg.algorithm == SMCB_APF && !f->cfg.fold_lookahead) return false;
  return true;
}

template <int MODEL, int PROP, int ALG, int NT, int MINB>
static cudaError_t launch_column_nt(int B, size_t dyn, cudaStream_t s, const ColumnArgs& c) {
  static size_t set_for = 0;   // the attribute is set once per instantiation (a driver call per launch costs microseconds on the online path)
  cudaError_t e = cudaSuccess;
  if (set_for < dyn) {
    e = cudaFuncSetAttribute(column_kernel<MODEL, PROP, ALG, NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e == cudaSuccess) set_for = dyn;
  }
  if (e != cudaSuccess) return e;
  static bool pdl_ok = true;
  if (use_pdl() && pdl_ok) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(B); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = dyn; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, column_kernel<MODEL, PROP, ALG, NT, MINB>, c) == cudaSuccess) return cudaSuccess;
    cudaGetLastError();
    pdl_ok = false;
  }
  column_kernel<MODEL, PROP, ALG, NT, MINB><<<B, NT, dyn, s>>>(c);
  return cudaSuccess;
}
template <int MODEL, int PROP>
static cudaError_t launch_column_alg(int alg, int minb, int B, size_t dyn, cudaStream_t s, const ColumnArgs& c) {
  if (alg == SMCB_SISR) {
    if (minb == 0) return launch_column_nt<MODEL, PROP, SMCB_ALG_SISR, 1024, 1>(B, dyn, s, c);
    if (minb == 1) return launch_column_nt<MODEL, PROP, SMCB_ALG_SISR, 512, 1>(B, dyn, s, c);
    return launch_column_nt<MODEL, PROP, SMCB_ALG_SISR, 512, 2>(B, dyn, s, c);
  }
  if (minb == 0) return launch_column_nt<MODEL, PROP, SMCB_ALG_APF, 1024, 1>(B, dyn, s, c);
  if (minb == 1) return launch_column_nt<MODEL, PROP, SMCB_ALG_APF, 512, 1>(B, dyn, s, c);
  return launch_column_nt<MODEL, PROP, SMCB_ALG_APF, 512, 2>(B, dyn, s, c);
}

static int run_column(smcb_filter* f, int steps, cudaStream_t s) {
  const bool apf = f->cfg.algorithm == SMCB_APF;
  const int t = f->t_host;
  if (steps < 1) return SMCB_OK;
  if (t - f->y_base < 0 || t + steps - 1 - f->y_base >= f->y_count || !f->y_dev) return fail(SMCB_ESTATE, "no observation set for this move");
  StepArgs a = make_args(f);
  a.t_host = t;
  if (f->xch.peer[0]) a.xch.seq = ++f->xch_seq;
  ColumnArgs c;
  memset(&c, 0, sizeof(c));
  if (a.xch.seq && f->want_inline && f->xch_out && f->xch_ticket && !getenv("SMCB_NO_INLINE_EXCHANGE")) {
    c.xch_out = f->xch_out; c.xch_ticket = f->xch_ticket; c.xch_rank = f->xch_rank;
    f->xch_inline_seq = a.xch.seq;
  }
  c.preweight_first = (apf && !f->folded_for_next) ? 1 : 0;  // apf.py:27-29 evaluated by the block itself: the previous launch could not fold it
  c.s = a;
  c.s.lw_out = f->lwbuf[(t + steps) & 1]; c.s.rw_out = f->rwbuf[(t + steps) & 1];
  c.steps = steps;
  c.y = f->y_dev + (int64_t)(t - f->y_base) * f->OD;
  c.y_avail = f->y_count - (t - f->y_base);
  c.u_in = f->u_in; c.u_out = f->u_out; c.w_out = f->w_out; c.quantize = 1;
  const size_t dyn = sizeof(float) * (size_t)f->D * RS_TILE;
  const int prop = f->cfg.proposal, alg = f->cfg.algorithm;
  // 512 threads x 8 particles; fewer columns than SMs: one block per SM with twice the registers, otherwise two blocks per SM
  int nt = (f->B <= smcb_sm_count()) ? 1 : 2;
  if (const char* v = getenv("SMCB_COLUMN_MINB")) nt = atoi(v);   // diagnostics: 0 = 1024 threads x 4, 1 = 512 x 8 one block per SM, 2 = 512 x 8 two per SM
  cudaError_t e = cudaSuccess;
  FOR_MODEL_PROP(f->cfg.model, prop, (e = launch_column_alg<MODEL, PROP>(alg, nt, f->B, dyn, s, c)));
  if (e != cudaSuccess) return fail(SMCB_ECUDA, cudaGetErrorString(e));
  f->launches++;
  const int t1 = t + steps;
  f->folded_for_next = apf && f->cfg.fold_lookahead && (t1 - f->y_base) < f->y_count;
  f->t_host = t1;
  CU(cudaGetLastError());
  return SMCB_OK;
}

extern "C" int smcb_filter_run(smcb_filter* f, int32_t steps, void* stream) {
  if (!f) return fail(SMCB_EINVAL, "null handle");
  cudaStream_t s = (cudaStream_t)stream;
  if (column_path_ok(f)) return run_column(f, steps, s);
  for (int k = 0; k < steps; ++k) {
    int rc = run_one(f, s, nullptr, k == steps - 1);
    if (rc) return rc;
  }
  CU(cudaGetLastError());
  return SMCB_OK;
}

extern "C" int smcb_filter_exchange_wait(smcb_filter* f, float** out_dev, void* stream);
// SMC2 / NESS drive the filter one observation at a time and look at the likelihood increments after every move (smc2.py:53-65,
// ness.py:56): `steps` launches of one move each; with an exchange attached every move is followed by the reader of that exchange.
extern "C" int smcb_filter_run_stepwise(smcb_filter* f, int32_t steps, void* stream) {
  if (!f) return fail(SMCB_EINVAL, "null handle");
  for (int k = 0; k < steps; ++k) {
    f->want_inline = true;   // every rank of the batch runs this loop at the same time (one process per GPU): the block that finishes a
    int rc = smcb_filter_run(f, 1, stream);   // rank's move last may wait, inside the launch, for the values of the other ranks
    f->want_inline = false;
    if (rc) return rc;
    if (f->xch.peer[0]) {
      float* out = nullptr;
      rc = smcb_filter_exchange_wait(f, &out, stream);   // (no launch when the move has gathered the exchange itself)
      if (rc) return rc;
    }
  }
  return SMCB_OK;
}

extern "C" int smcb_filter_profile(smcb_filter* f, int32_t steps, float* out_ms_host, void* stream) {
  if (!f || !out_ms_host || steps < 1) return fail(SMCB_EINVAL, "bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  std::vector<cudaEvent_t> ev((size_t)steps * 6);
  for (auto& e : ev) CU(cudaEventCreate(&e));
  int rc = SMCB_OK;
  for (int k = 0; k < steps && rc == SMCB_OK; ++k) rc = run_one(f, s, &ev[(size_t)k * 6], k == steps - 1);
  cudaError_t e2 = cudaStreamSynchronize(s);
  for (int g = 0; g < 5; ++g) out_ms_host[g] = 0.f;
  if (rc == SMCB_OK && e2 == cudaSuccess)
    for (int k = 0; k < steps; ++k)
      for (int g = 0; g < 5; ++g) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev[(size_t)k * 6 + g], ev[(size_t)k * 6 + g + 1]);
        out_ms_host[g] += ms;
      }
  for (auto& e : ev) cudaEventDestroy(e);
  if (e2 != cudaSuccess) return fail(SMCB_ECUDA, cudaGetErrorString(e2));
  return rc;
}

extern "C" int smcb_filter_batch_filter_host(smcb_filter* f, const float* y_host, int32_t T, float* means_host, float* vars_host,
                                             float* ll_steps_host, float* ll_total_host, void* stream) {
  if (!f || !y_host) return fail(SMCB_EINVAL, "null argument");
  if (T + 1 > f->cfg.history_rows && (means_host || vars_host || ll_steps_host))
    return fail(SMCB_EINVAL, "history_rows of the handle is smaller than T + 1");
  cudaStream_t s = (cudaStream_t)stream;
  if (f->y_own_cap < T * f->OD) {
    if (f->y_own) cudaFree(f->y_own);
    f->y_own = nullptr; f->y_own_cap = 0;
    CU(cudaMalloc((void**)&f->y_own, sizeof(float) * (size_t)T * f->OD));
    f->y_own_cap = T * f->OD;
  }
  CU(cudaMemcpyAsync(f->y_own, y_host, sizeof(float) * (size_t)T * f->OD, cudaMemcpyHostToDevice, s));
  int rc = smcb_filter_initialize(f, stream);
  if (rc) return rc;
  rc = smcb_filter_set_observations(f, f->y_own, T, 0, stream);
  if (rc) return rc;
  rc = smcb_filter_run(f, T, stream);
  if (rc) return rc;
  const size_t rows = (size_t)T + 1;
  if (means_host) CU(cudaMemcpyAsync(means_host, f->hist_mean, sizeof(float) * rows * f->B * f->D, cudaMemcpyDeviceToHost, s));
  if (vars_host) CU(cudaMemcpyAsync(vars_host, f->hist_var, sizeof(float) * rows * f->B * f->D, cudaMemcpyDeviceToHost, s));
  if (ll_steps_host) CU(cudaMemcpyAsync(ll_steps_host, f->hist_ll, sizeof(float) * rows * f->B, cudaMemcpyDeviceToHost, s));
  if (ll_total_host) CU(cudaMemcpyAsync(ll_total_host, f->ll_total, sizeof(float) * f->B, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return SMCB_OK;
}

extern "C" int smcb_filter_attach_exchange(smcb_filter* f, const uint64_t* peer_ptrs_host, int32_t world, int32_t rank, int32_t total_columns,
                                           int32_t first_column) {
  if (!f) return fail(SMCB_EINVAL, "null handle");
  if (!peer_ptrs_host) {  // detach
    f->xch = ExchangeArgs{}; f->xch_seq = 0;
    return SMCB_OK;
  }
  if (world < 1 || world > SMCB_MAX_PEERS || rank < 0 || rank >= world) return fail(SMCB_EINVAL, "world size must be 1 .. 8, rank inside it");
  if (total_columns < f->B || first_column < 0 || first_column + f->B > total_columns) return fail(SMCB_EINVAL, "this rank's columns do not fit the batch");
  ExchangeArgs x = {};
  for (int r = 0; r < world; ++r) {
    if (!peer_ptrs_host[r]) return fail(SMCB_EINVAL, "null peer buffer");
    x.peer[r] = (unsigned long long*)(uintptr_t)peer_ptrs_host[r];
  }
  x.world = world; x.total = total_columns; x.lo = first_column; x.seq = 0;
  if (f->xch_out) { cudaFree(f->xch_out); f->xch_out = nullptr; }
  CU(cudaMalloc((void**)&f->xch_out, sizeof(float) * 2 * (size_t)total_columns));
  if (!f->xch_ticket) CU(dalloc(&f->xch_ticket, (size_t)1));
  f->xch = x; f->xch_seq = 0; f->xch_rank = rank;
  return SMCB_OK;
}

extern "C" int smcb_filter_exchange_wait(smcb_filter* f, float** out_dev, void* stream) {
  if (!f || !out_dev) return fail(SMCB_EINVAL, "null argument");
  if (!f->xch.peer[0] || !f->xch_seq) return fail(SMCB_ESTATE, "no exchange attached or nothing published yet");
  if (f->xch_inline_seq == f->xch_seq) {   // the column kernel's last block has gathered this exchange inside the move's launch
    *out_dev = f->xch_out;
    return SMCB_OK;
  }
  unsigned long long* mine = f->xch.peer[f->xch_rank];   // the buffer the peers (and this rank) stored into
  exchange_wait_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(mine, f->xch.total, f->xch_seq, f->xch_out, f->wd ? f->wd + 3 : nullptr);
  f->launches++;
  CU(cudaGetLastError());
  *out_dev = f->xch_out;
  return SMCB_OK;
}

extern "C" int smcb_filter_set_noise(smcb_filter* f, const float* eps_dev, const float* u_dev, const double* U_dev) {
  if (!f) return fail(SMCB_EINVAL, "null handle");
  f->eps_in = eps_dev; f->u_in = u_dev; f->U_in = U_dev;
  return SMCB_OK;
}
extern "C" int smcb_filter_set_nested_noise(smcb_filter* f, const float* z_dev, const float* e_dev) {
  if (!f) return fail(SMCB_EINVAL, "null handle");
  f->nest_z = z_dev; f->nest_e = e_dev;
  return SMCB_OK;
}
extern "C" int smcb_filter_dump_noise(smcb_filter* f, float* eps_dev, float* u_dev, float* w_dev) {
  if (!f) return fail(SMCB_EINVAL, "null handle");
  f->eps_out = eps_dev; f->u_out = u_dev; f->w_out = w_dev;
  return SMCB_OK;
}

extern "C" int smcb_filter_ptr(smcb_filter* f, int32_t what, void** p) {
  if (!f || !p) return fail(SMCB_EINVAL, "null argument");
  switch (what) {
    case SMCB_PTR_X: *p = f->xbuf[f->t_host & 1]; break;
    case SMCB_PTR_X_OTHER: *p = f->xbuf[(f->t_host + 1) & 1]; break;
    case SMCB_PTR_LOGW: *p = f->lwbuf[f->t_host & 1]; break;
    case SMCB_PTR_PREV_INDS: *p = f->prev_inds; break;
    case SMCB_PTR_MEAN: *p = f->latest_mean; break;
    case SMCB_PTR_VAR: *p = f->latest_var; break;
    case SMCB_PTR_LL: *p = f->latest_ll; break;
    case SMCB_PTR_LL_TOTAL: *p = f->ll_total; break;
    case SMCB_PTR_HIST_MEAN: *p = f->hist_mean; break;
    case SMCB_PTR_HIST_VAR: *p = f->hist_var; break;
    case SMCB_PTR_HIST_LL: *p = f->hist_ll; break;
    case SMCB_PTR_ESS: *p = f->ess_packed; break;
    case SMCB_PTR_RESAMPLE_LOGW: *p = f->rwbuf[f->t_host & 1]; break;
    case 20: *p = f->dbg; break;  /* diagnostics */
    case 21: *p = f->verdict; break;
    case 22: *p = f->wd; break;
    default: return fail(SMCB_EINVAL, "unknown pointer id");
  }
  return SMCB_OK;
}

__global__ void pack_stats_kernel(const ColStats* st, float* out, int B) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) { out[b] = st[b].ess; out[B + b] = (float)st[b].resample; }
}
extern "C" int smcb_filter_sync_stats(smcb_filter* f, void* stream) {
  if (!f) return fail(SMCB_EINVAL, "null handle");
  pack_stats_kernel<<<(f->B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(f->stats, f->ess_packed, f->B);
  f->launches++;
  CU(cudaGetLastError());
  return SMCB_OK;
}

// ---- stand-alone operators ---------------------------------------------------------------------------------------------------------
struct OpWorkspace {
  float* w = nullptr; float* wn = nullptr; int32_t* anc = nullptr; double* tilesum = nullptr; Ctrl* ctrl = nullptr;
  double *prefix = nullptr, *sin = nullptr; int32_t *tileflag = nullptr, *dcounter = nullptr; XsDesc *desc = nullptr, *desc2 = nullptr; SegTable* tables = nullptr;
  ColStats* stats = nullptr; NormPartial* parts = nullptr; float* cbuf = nullptr; float* mid = nullptr;
  uint32_t* tilemin = nullptr; int32_t* ncounter = nullptr; int32_t* verdict = nullptr; float* u_col = nullptr;
  int64_t ld = 0; int tiles = 0, nblk = 0;
};

// One slab per calling thread, kept between calls: the operators are called once per filter step by code written against the reference
// (`resampling=` callables, normalize / get_ess on the theta level), and nineteen stream-ordered allocations plus six memsets per call
// cost more host time than the kernels take.  The slab is reused while it is large enough and the stream is the same (work on one
// stream is ordered, so the previous call has finished with it by the time the next one touches it); otherwise it is released on its
// own stream and allocated anew.  The buffers that must start zeroed sit together at the front: one memset.
struct OpSlab { char* base = nullptr; size_t bytes = 0; cudaStream_t stream = nullptr; int device = -1; };
static thread_local OpSlab g_slab;

static int op_alloc(OpWorkspace& ws, int64_t n, int B, cudaStream_t s, bool want_cbuf = false) {
  ws.tiles = (int)((n + RS_TILE - 1) / RS_TILE);
  ws.ld = (int64_t)ws.tiles * RS_TILE;
  ws.nblk = ws.tiles;
  const size_t cells = (size_t)B * ws.ld;
  const size_t nt = (size_t)B * ws.tiles;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  // zeroed region
  const size_t o_dcounter = take((size_t)B * sizeof(int32_t)), o_ncounter = take((size_t)B * sizeof(int32_t)), o_verdict = take((size_t)B * sizeof(int32_t));
  const size_t o_ctrl = take(sizeof(Ctrl)), o_stats = take((size_t)B * sizeof(ColStats)), o_w = take(cells * sizeof(float));
  const size_t zero_bytes = off;
  const size_t o_wn = take(cells * sizeof(float)), o_anc = take(cells * sizeof(int32_t));
  const size_t o_tilesum = take(nt * sizeof(double)), o_prefix = take(nt * sizeof(double)), o_sin = take(nt * sizeof(double));
  const size_t o_tileflag = take(nt * sizeof(int32_t)), o_desc = take(nt * sizeof(XsDesc)), o_desc2 = take(nt * sizeof(XsDesc));
  const size_t o_tables = take(nt * sizeof(SegTable)), o_parts = take((size_t)B * ws.nblk * sizeof(NormPartial));
  const size_t o_tilemin = take(nt * sizeof(uint32_t)), o_ucol = take((size_t)B * sizeof(float));
  const size_t o_cbuf = want_cbuf ? take(cells * sizeof(float)) : 0;
  const size_t o_mid = want_cbuf ? take((size_t)B * ((n + MN_MID - 1) / MN_MID) * sizeof(float)) : 0;
  int dev = 0;
  CU(cudaGetDevice(&dev));
  OpSlab& sl = g_slab;
  if (!sl.base || sl.bytes < off || sl.stream != s || sl.device != dev) {
    if (sl.base) { cudaFreeAsync(sl.base, sl.stream); sl = OpSlab(); }
    CU(cudaMallocAsync((void**)&sl.base, off, s));
    sl.bytes = off; sl.stream = s; sl.device = dev;
  }
  char* p = sl.base;
  CU(cudaMemsetAsync(p, 0, zero_bytes, s));
  ws.dcounter = (int32_t*)(p + o_dcounter); ws.ncounter = (int32_t*)(p + o_ncounter); ws.verdict = (int32_t*)(p + o_verdict);
  ws.ctrl = (Ctrl*)(p + o_ctrl); ws.stats = (ColStats*)(p + o_stats); ws.w = (float*)(p + o_w);
  ws.wn = (float*)(p + o_wn); ws.anc = (int32_t*)(p + o_anc);
  ws.tilesum = (double*)(p + o_tilesum); ws.prefix = (double*)(p + o_prefix); ws.sin = (double*)(p + o_sin);
  ws.tileflag = (int32_t*)(p + o_tileflag); ws.desc = (XsDesc*)(p + o_desc); ws.desc2 = (XsDesc*)(p + o_desc2);
  ws.tables = (SegTable*)(p + o_tables); ws.parts = (NormPartial*)(p + o_parts);
  ws.tilemin = (uint32_t*)(p + o_tilemin); ws.u_col = (float*)(p + o_ucol);
  ws.cbuf = want_cbuf ? (float*)(p + o_cbuf) : nullptr;
  ws.mid = want_cbuf ? (float*)(p + o_mid) : nullptr;
  return SMCB_OK;
}
static void op_free(OpWorkspace& ws, cudaStream_t) { ws = OpWorkspace(); }   // the slab stays with the thread for the next call

static int op_prepare(OpWorkspace& ws, const float* w_dev, int64_t n, int B, int64_t sn, int64_t sb, bool need_stats, cudaStream_t s, bool want_cbuf = false) {
  int rc = op_alloc(ws, n, B, s, want_cbuf);
  if (rc) return rc;
  op_launch_gather_rows(w_dev, n, B, sn, sb, ws.w, ws.ld, s);
  if (need_stats) op_launch_colstats(ws.w, n, B, ws.ld, ws.nblk, ws.parts, ws.stats, s);
  CU(cudaGetLastError());
  return SMCB_OK;
}

extern "C" int smcb_normalize(const float* logw_dev, int64_t n, int32_t B, int64_t sn, int64_t sb, float* out_dev, int64_t osn,
                              int64_t osb, float* ess_out_dev, void* stream) {
  if (!logw_dev || n < 1 || B < 1) return fail(SMCB_EINVAL, "bad argument");
  if (smcb_device_count() < 1) return fail(SMCB_ENODEVICE, "no CUDA device: libsmcb200 has no CPU fallback");
  cudaStream_t s = (cudaStream_t)stream;
  OpWorkspace ws;
  int rc = op_prepare(ws, logw_dev, n, B, sn, sb, true, s);
  if (rc == SMCB_OK) {
    if (out_dev) op_launch_apply_weights(ws.w, ws.stats, n, B, ws.ld, out_dev, osn, osb, s);
    if (ess_out_dev) op_launch_copy_ess(ws.stats, ess_out_dev, B, s);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = fail(SMCB_ECUDA, cudaGetErrorString(e));
  }
  op_free(ws, s);
  return rc;
}

extern "C" int smcb_get_ess(const float* w_dev, int64_t n, int32_t B, int64_t sn, int64_t sb, int32_t normalized, float* ess_out_dev,
                            void* stream) {
  if (!w_dev || !ess_out_dev || n < 1 || B < 1) return fail(SMCB_EINVAL, "bad argument");
  if (!normalized) return smcb_normalize(w_dev, n, B, sn, sb, nullptr, 0, 0, ess_out_dev, stream);
  if (smcb_device_count() < 1) return fail(SMCB_ENODEVICE, "no CUDA device: libsmcb200 has no CPU fallback");
  ess_normalized_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(w_dev, n, sn, sb, ess_out_dev);
  CU(cudaGetLastError());
  return SMCB_OK;
}

static int op_resample(const float* w_dev, int64_t n, int32_t B, int64_t sn, int64_t sb, int32_t normalized, const float* u_dev,
                       const double* U_dev, uint64_t seed, int64_t* out_dev, int64_t osn, int64_t osb, int kind, cudaStream_t s) {
  if (!w_dev || !out_dev || n < 1 || B < 1) return fail(SMCB_EINVAL, "bad argument");
  if (n >= ((int64_t)1 << 31) - RS_TILE) return fail(SMCB_EINVAL, "too many particles");
  if (kind == SMCB_MULTINOMIAL && n > (1 << 24)) return fail(SMCB_EINVAL, "number of categories cannot exceed 2^24");
  if (smcb_device_count() < 1) return fail(SMCB_ENODEVICE, "no CUDA device: libsmcb200 has no CPU fallback");
  OpWorkspace ws;
  int rc = op_prepare(ws, w_dev, n, B, sn, sb, !normalized, s, kind == SMCB_MULTINOMIAL);
  if (rc == SMCB_OK) {
    ResampleArgs r;
    memset(&r, 0, sizeof(r));
    r.w = ws.w; r.wn = normalized ? ws.w : ws.wn; r.n = n; r.ld = ws.ld; r.B = B; r.tiles_per_col = ws.tiles;
    r.input_is_w = normalized ? 1 : 0; r.use_rw = 0; r.stats = normalized ? nullptr : ws.stats;
    r.u_in = u_dev; r.seed = seed; r.tilesum = ws.tilesum; r.anc = ws.anc; r.ctrl = ws.ctrl;
    r.prefix = ws.prefix; r.sin = ws.sin; r.tileflag = ws.tileflag; r.desc = ws.desc; r.desc2 = ws.desc2; r.tables = ws.tables; r.dcounter = ws.dcounter;
    r.tilemin = ws.tilemin; r.ncounter = ws.ncounter; r.verdict = ws.verdict; r.u_col = ws.u_col;
    if (kind == SMCB_SYSTEMATIC) op_launch_systematic(r, s);
    else {
      r.c_out = ws.cbuf; r.mid_out = ws.mid;
      if (rc == SMCB_OK) rc = op_launch_multinomial(r, U_dev, n, s);
    }
    if (rc == SMCB_OK) {
      op_launch_scatter_i64(ws.anc, n, B, ws.ld, out_dev, osn, osb, s);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) rc = fail(SMCB_ECUDA, cudaGetErrorString(e));
    }
  }
  op_free(ws, s);
  return rc;
}

extern "C" int smcb_systematic(const float* w_dev, int64_t n, int32_t B, int64_t sn, int64_t sb, int32_t normalized,
                               const float* u_dev, uint64_t seed, int64_t* out_dev, int64_t osn, int64_t osb, void* stream) {
  return op_resample(w_dev, n, B, sn, sb, normalized, u_dev, nullptr, seed, out_dev, osn, osb, SMCB_SYSTEMATIC, (cudaStream_t)stream);
}
extern "C" int smcb_multinomial(const float* w_dev, int64_t n, int32_t B, int64_t sn, int64_t sb, int32_t normalized,
                                const double* U_dev, uint64_t seed, int64_t* out_dev, int64_t osn, int64_t osb, void* stream) {
  return op_resample(w_dev, n, B, sn, sb, normalized, nullptr, U_dev, seed, out_dev, osn, osb, SMCB_MULTINOMIAL, (cudaStream_t)stream);
}

// ---- the callers and plug-ins either side of the fused move (plugin.cuh) ----------------------------------------------------------
extern "C" int smcb_filter_set_ess_threshold(smcb_filter* f, float relative_threshold) {
  if (!f) return fail(SMCB_EINVAL, "null handle");
  f->cfg.ess_threshold = relative_threshold;
  return SMCB_OK;
}

static int proposal_op(smcb_filter* f, int mode, const float* y_dev, const float* x_dev, const float* eps_dev, int32_t t, float* x_out, float* w_out,
                       cudaStream_t s) {
  if (!f || !y_dev || !w_out || (mode == 1 && !x_out)) return fail(SMCB_EINVAL, "null argument");
  ProposalOpArgs c;
  memset(&c, 0, sizeof(c));
  c.s = make_args(f);
  c.s.eps_in = eps_dev; c.s.eps_out = nullptr;
  c.x_in = x_dev ? x_dev : f->xbuf[f->t_host & 1];
  c.y = y_dev; c.x_out = x_out; c.w_out = w_out; c.mode = mode; c.t = t;
  const int64_t chunk = ST_NT * ST_VEC;
  dim3 g((unsigned)((f->n + chunk - 1) / chunk), f->B);
  const int prop = f->cfg.proposal;
  FOR_MODEL_PROP(f->cfg.model, prop, (proposal_op_kernel<MODEL, PROP><<<g, ST_NT, 0, s>>>(c)));
  f->launches++;
  CU(cudaGetLastError());
  return SMCB_OK;
}
extern "C" int smcb_filter_pre_weight(smcb_filter* f, const float* y_dev, const float* x_dev, float* out_dev, void* stream) {
  return proposal_op(f, 0, y_dev, x_dev, nullptr, 0, nullptr, out_dev, (cudaStream_t)stream);
}
extern "C" int smcb_filter_sample_and_weight(smcb_filter* f, const float* y_dev, const float* x_dev, const float* eps_dev, int32_t t,
                                             float* x_out_dev, float* w_out_dev, void* stream) {
  return proposal_op(f, 1, y_dev, x_dev, eps_dev, t, x_out_dev, w_out_dev, (cudaStream_t)stream);
}

extern "C" int smcb_filter_predict_path(smcb_filter* f, int32_t steps, const float* x_dev, float* x_out_dev, float* y_out_dev, void* stream) {
  if (!f || !x_out_dev || !y_out_dev || steps < 1) return fail(SMCB_EINVAL, "bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  PathArgs c;
  memset(&c, 0, sizeof(c));
  c.s = make_args(f);
  c.x_in = x_dev ? x_dev : f->xbuf[f->t_host & 1];
  c.x_out = x_out_dev; c.y_out = y_out_dev; c.steps = steps; c.t0 = f->t_host;
  const int64_t chunk = ST_NT * ST_VEC;
  dim3 g((unsigned)((f->n + chunk - 1) / chunk), f->B);
  FOR_MODEL(f->cfg.model, (predict_path_kernel<MODEL><<<g, ST_NT, 0, s>>>(c)));
  f->launches++;
  CU(cudaGetLastError());
  return SMCB_OK;
}

extern "C" int smcb_batched_gather(const float* x_dev, int64_t n, int32_t B, int32_t D, int64_t* idx_dev, const int64_t* prev_dev, float* out_dev,
                                   void* stream) {
  if (!x_dev || !idx_dev || !out_dev || n < 1 || B < 1 || D < 1) return fail(SMCB_EINVAL, "bad argument");
  if (smcb_device_count() < 1) return fail(SMCB_ENODEVICE, "no CUDA device: libsmcb200 has no CPU fallback");
  cudaStream_t s = (cudaStream_t)stream;
  int* bad = nullptr;
  CU(cudaMallocAsync((void**)&bad, sizeof(int), s));
  CU(cudaMemsetAsync(bad, 0, sizeof(int), s));
  const int64_t cells = n * B;
  gather_lineage_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, s>>>(x_dev, n, B, D, idx_dev, prev_dev, out_dev, bad);
  int h = 0;
  cudaError_t e = cudaMemcpyAsync(&h, bad, sizeof(int), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFreeAsync(bad, s);
  if (e != cudaSuccess) return fail(SMCB_ECUDA, cudaGetErrorString(e));
  if (h) return fail(SMCB_EINVAL, "index out of range");   // torch.gather raises the same way
  return SMCB_OK;
}

// the per-column buffers a theta-level permutation / exchange touches
struct ColumnPlanes { float* p[8]; int planes[8]; int count; };
static ColumnPlanes column_planes(smcb_filter* f) {
  ColumnPlanes c;
  const int cur = f->t_host & 1;
  c.count = 0;
  c.p[c.count] = f->xbuf[cur]; c.planes[c.count++] = f->D;
  c.p[c.count] = f->lwbuf[cur]; c.planes[c.count++] = 1;
  c.p[c.count] = f->rwbuf[cur]; c.planes[c.count++] = 1;
  c.p[c.count] = (float*)f->prev_inds; c.planes[c.count++] = 1;
  return c;
}
struct ColumnSmall { float* p; int rec, rows; };
static int column_small(smcb_filter* f, bool history, ColumnSmall* out) {
  int k = 0;
  out[k++] = {(float*)f->stats, (int)(sizeof(ColStats) / 4), 1};
  out[k++] = {f->ll_total, 1, 1};
  out[k++] = {f->latest_mean, f->D, 1};
  out[k++] = {f->latest_var, f->D, 1};
  out[k++] = {f->latest_ll, 1, 1};
  if (history) {
    const int rows = std::min(f->cfg.history_rows, f->t_host + 1);
    out[k++] = {f->hist_mean, f->D, rows};
    out[k++] = {f->hist_var, f->D, rows};
    out[k++] = {f->hist_ll, 1, rows};
  }
  return k;
}

extern "C" int smcb_filter_resample_columns(smcb_filter* f, const int64_t* idx_dev, int32_t entire_history, void* stream) {
  if (!f || !idx_dev) return fail(SMCB_EINVAL, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  int* bad = nullptr;
  CU(cudaMallocAsync((void**)&bad, sizeof(int), s));
  CU(cudaMemsetAsync(bad, 0, sizeof(int), s));
  const ColumnPlanes cp = column_planes(f);
  const int64_t ld4 = f->ld / 4;
  for (int k = 0; k < cp.count; ++k) {   // out of place through a scratch copy of the plane group, then back
    const size_t bytes = sizeof(float) * (size_t)cp.planes[k] * f->B * f->ld;
    float* tmp = nullptr;
    CU(cudaMallocAsync((void**)&tmp, bytes, s));
    dim3 g((unsigned)std::min<int64_t>((ld4 + 255) / 256, 64), f->B, cp.planes[k]);
    column_gather_kernel<<<g, 256, 0, s>>>((const float4*)cp.p[k], (float4*)tmp, f->B, ld4, idx_dev, nullptr, bad);
    CU(cudaMemcpyAsync(cp.p[k], tmp, bytes, cudaMemcpyDeviceToDevice, s));
    CU(cudaFreeAsync(tmp, s));
    f->launches++;
  }
  ColumnSmall sm[8];
  const int ns = column_small(f, entire_history != 0, sm);
  for (int k = 0; k < ns; ++k) {
    const size_t count = (size_t)sm[k].rows * f->B * sm[k].rec;
    float* tmp = nullptr;
    CU(cudaMallocAsync((void**)&tmp, count * sizeof(float), s));
    column_gather_small_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(sm[k].p, tmp, f->B, sm[k].rec, sm[k].rows, idx_dev, nullptr);
    CU(cudaMemcpyAsync(sm[k].p, tmp, count * sizeof(float), cudaMemcpyDeviceToDevice, s));
    CU(cudaFreeAsync(tmp, s));
    f->launches++;
  }
  int h = 0;
  cudaError_t e = cudaMemcpyAsync(&h, bad, sizeof(int), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFreeAsync(bad, s);
  if (e != cudaSuccess) return fail(SMCB_ECUDA, cudaGetErrorString(e));
  if (h) return fail(SMCB_EINVAL, "column index out of range");
  CU(cudaGetLastError());
  return SMCB_OK;
}

extern "C" int smcb_filter_exchange_columns(smcb_filter* dst, smcb_filter* src, const uint8_t* mask_dev, void* stream) {
  if (!dst || !src || !mask_dev) return fail(SMCB_EINVAL, "null argument");
  if (dst->B != src->B || dst->n != src->n || dst->D != src->D || dst->ld != src->ld) return fail(SMCB_EINVAL, "the two handles hold different shapes");
  if (dst->t_host != src->t_host) return fail(SMCB_ESTATE, "the two handles are at different move indices");
  cudaStream_t s = (cudaStream_t)stream;
  const ColumnPlanes a = column_planes(dst), b = column_planes(src);
  const int64_t ld4 = dst->ld / 4;
  for (int k = 0; k < a.count; ++k) {
    dim3 g((unsigned)std::min<int64_t>((ld4 + 255) / 256, 64), dst->B, a.planes[k]);
    column_gather_kernel<<<g, 256, 0, s>>>((const float4*)b.p[k], (float4*)a.p[k], dst->B, ld4, nullptr, mask_dev, nullptr);
    dst->launches++;
  }
  ColumnSmall sa[8], sb[8];
  const int ns = column_small(dst, true, sa);
  column_small(src, true, sb);
  for (int k = 0; k < ns; ++k) {
    const int rows = std::min(sa[k].rows, sb[k].rows);
    const size_t count = (size_t)rows * dst->B * sa[k].rec;
    column_gather_small_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(sb[k].p, sa[k].p, dst->B, sa[k].rec, rows, nullptr, mask_dev);
    dst->launches++;
  }
  dst->folded_for_next = dst->folded_for_next && src->folded_for_next;   // rw is valid only when both handles folded the look-ahead
  CU(cudaGetLastError());
  return SMCB_OK;
}

// pyfilter.resampling.residual (resampling.py:68-105) on NORMALISED weights: deterministic copies + multinomial draws on the fractions
extern "C" int smcb_residual(const float* w_dev, int64_t n, int32_t B, int64_t sn, int64_t sb, const double* U_dev, uint64_t seed, int64_t* out_dev,
                             int64_t osn, int64_t osb, void* stream) {
  if (!w_dev || !out_dev || n < 1 || B < 1) return fail(SMCB_EINVAL, "bad argument");
  if (n > (1 << 24)) return fail(SMCB_EINVAL, "number of categories cannot exceed 2^24");
  if (smcb_device_count() < 1) return fail(SMCB_ENODEVICE, "no CUDA device: libsmcb200 has no CPU fallback");
  cudaStream_t s = (cudaStream_t)stream;
  OpWorkspace ws;
  int32_t *counts = nullptr, *ksum = nullptr;
  int rc = op_prepare(ws, w_dev, n, B, sn, sb, false, s, true);
  if (rc == SMCB_OK) {
    const size_t cells = (size_t)B * ws.ld;
    cudaError_t e = cudaMallocAsync((void**)&counts, cells * sizeof(int32_t), s);
    if (e == cudaSuccess) e = cudaMallocAsync((void**)&ksum, (size_t)B * sizeof(int32_t), s);
    if (e == cudaSuccess) e = cudaMemsetAsync(ws.wn, 0, cells * sizeof(float), s);
    if (e != cudaSuccess) rc = fail(SMCB_ECUDA, cudaGetErrorString(e));
  }
  if (rc == SMCB_OK) {
    int32_t* tile_sum = (int32_t*)ws.tileflag;   // (B, tiles) int32 scratch of the slab, unused until the multinomial pipeline below
    const dim3 rg(ws.tiles, B);
    residual_counts_kernel<<<rg, RES_NT, 0, s>>>(ws.w, n, ws.ld, ws.tiles, counts, ws.wn, tile_sum);
    residual_scan_kernel<<<B, 1024, 0, s>>>(tile_sum, ws.tiles, ksum);
    residual_expand_kernel<<<rg, RES_NT, 0, s>>>(counts, ws.wn, n, ws.ld, ws.tiles, tile_sum, ksum, ws.anc);
    ResampleArgs r;
    memset(&r, 0, sizeof(r));
    r.w = ws.wn; r.wn = ws.wn; r.n = n; r.ld = ws.ld; r.B = B; r.tiles_per_col = ws.tiles;
    r.input_is_w = 1; r.use_rw = 0; r.stats = nullptr;
    r.seed = seed; r.tilesum = ws.tilesum; r.anc = ws.anc; r.ctrl = ws.ctrl;
    r.prefix = ws.prefix; r.sin = ws.sin; r.tileflag = ws.tileflag; r.desc = ws.desc; r.desc2 = ws.desc2; r.tables = ws.tables; r.dcounter = ws.dcounter;
    r.tilemin = ws.tilemin; r.ncounter = ws.ncounter; r.verdict = ws.verdict; r.u_col = ws.u_col;
    r.c_out = ws.cbuf; r.mid_out = ws.mid; r.draw_offset = ksum;
    rc = op_launch_multinomial(r, U_dev, n, s);
    if (rc == SMCB_OK) {
      op_launch_scatter_i64(ws.anc, n, B, ws.ld, out_dev, osn, osb, s);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) rc = fail(SMCB_ECUDA, cudaGetErrorString(e));
    }
  }
  if (counts) cudaFreeAsync(counts, s);
  if (ksum) cudaFreeAsync(ksum, s);
  op_free(ws, s);
  return rc;
}

// one backward step of FFBS (filters/particle/base.py:112-126) for a non-batched filter
extern "C" int smcb_filter_ffbs_step(smcb_filter* f, const float* x_dev, const float* lw_dev, const float* xnext_dev, const double* U_dev, uint64_t seed,
                                     int32_t t, int64_t* idx_out_dev, float* x_out_dev, void* stream) {
  if (!f || !x_dev || !lw_dev || !xnext_dev || !idx_out_dev || !x_out_dev) return fail(SMCB_EINVAL, "null argument");
  if (f->B != 1) return fail(SMCB_EUNSUPPORTED, "FFBS is implemented for non-batched filters (like the reference's working branch)");
  cudaStream_t s = (cudaStream_t)stream;
  FfbsArgs a;
  memset(&a, 0, sizeof(a));
  a.P = f->P_dev; a.x = x_dev; a.lw = lw_dev; a.xnext = xnext_dev; a.U = U_dev; a.idx = idx_out_dev; a.xout = x_out_dev;
  a.n = f->n; a.seed = seed; a.t = t;
  FOR_MODEL(f->cfg.model, (ffbs_step_kernel<MODEL><<<(unsigned)f->n, 256, 0, s>>>(a)));
  f->launches++;
  CU(cudaGetLastError());
  return SMCB_OK;
}

// ---- columns as records: export / import (cross-rank theta-resampling of a sharded batch) -------------------------------------------
static int64_t column_record_layout(smcb_filter* f, ColPackArgs& a) {
  memset(&a, 0, sizeof(a));
  const int cur = f->t_host & 1;
  int k = 0;
  int64_t off = 0;
  auto add = [&](void* ptr, int rows, int64_t inner, int64_t row_stride, int64_t col_stride) {
    a.d[k].ptr = (uint32_t*)ptr; a.d[k].rows = rows; a.d[k].inner = (int32_t)inner; a.d[k].row_stride = row_stride; a.d[k].col_stride = col_stride;
    a.d[k].offset = off;
    off += (int64_t)rows * inner;
    ++k;
  };
  const int64_t plane = (int64_t)f->B * f->ld;
  add(f->xbuf[cur], f->D, f->ld, plane, f->ld);              // particles (D, B, ld)
  add(f->lwbuf[cur], 1, f->ld, 0, f->ld);
  add(f->rwbuf[cur], 1, f->ld, 0, f->ld);
  add(f->prev_inds, 1, f->ld, 0, f->ld);
  add(f->stats, 1, sizeof(ColStats) / 4, 0, sizeof(ColStats) / 4);
  add(f->ll_total, 1, 1, 0, 1);
  add(f->latest_mean, 1, f->D, 0, f->D);
  add(f->latest_var, 1, f->D, 0, f->D);
  add(f->latest_ll, 1, 1, 0, 1);
  const int rows = f->cfg.history_rows;
  add(f->hist_mean, rows, f->D, (int64_t)f->B * f->D, f->D);   // (rows, B, D)
  add(f->hist_var, rows, f->D, (int64_t)f->B * f->D, f->D);
  add(f->hist_ll, rows, 1, f->B, 1);
  a.ndesc = k; a.B = f->B; a.record = off;
  return off;
}
extern "C" int64_t smcb_filter_column_record_elems(smcb_filter* f) {
  if (!f) return 0;
  ColPackArgs a;
  return column_record_layout(f, a);
}
extern "C" int smcb_filter_export_columns(smcb_filter* f, void* buf_dev, void* stream) {
  if (!f || !buf_dev) return fail(SMCB_EINVAL, "null argument");
  ColPackArgs a;
  column_record_layout(f, a);
  a.buf = (uint32_t*)buf_dev;
  column_pack_kernel<true><<<dim3(16, f->B), 256, 0, (cudaStream_t)stream>>>(a);
  f->launches++;
  CU(cudaGetLastError());
  return SMCB_OK;
}
extern "C" int smcb_filter_import_columns(smcb_filter* f, const void* buf_dev, int32_t n_records, const int64_t* idx_dev, int32_t folded, void* stream) {
  if (!f || !buf_dev || n_records < 1) return fail(SMCB_EINVAL, "bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  ColPackArgs a;
  column_record_layout(f, a);
  a.buf = (uint32_t*)buf_dev; a.idx = idx_dev; a.n_records = n_records;
  int* bad = nullptr;
  CU(cudaMallocAsync((void**)&bad, sizeof(int), s));
  CU(cudaMemsetAsync(bad, 0, sizeof(int), s));
  a.bad = bad;
  column_pack_kernel<false><<<dim3(16, f->B), 256, 0, s>>>(a);
  f->launches++;
  int h = 0;
  cudaError_t e = cudaMemcpyAsync(&h, bad, sizeof(int), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFreeAsync(bad, s);
  if (e != cudaSuccess) return fail(SMCB_ECUDA, cudaGetErrorString(e));
  if (h) return fail(SMCB_EINVAL, "record index out of range");
  if (folded >= 0) f->folded_for_next = folded != 0;
  return SMCB_OK;
}
