// The fused per-time-step kernels of the particle filter (reference filters/particle/{sisr,apf}.py, proposals/{bootstrap,linear}.py).
//
//   state_kernel      draws x_0 (or takes the caller's state) and produces the soft-max partials of the log-weights
//   preweight_kernel  APF look-ahead weights  rw = lw + log p(y_t | .)  when they were not folded into the previous step
//   step_kernel       ancestor gather -> transition / proposal sample -> observation log-density -> APF correction ->
//                     128-bit stores of x_t and the log-weight, plus per-block partials for every reduction the step needs
//                     (normalisers, ESS, log-likelihood increment, filter mean and variance)
//   finalize_kernel   one block per column folds the partials into ColStats + the moment / likelihood history
// Layout: state is SoA  x[dim][column][particle]  (pitch ld), weights  lw[column][particle]; one thread owns 4 consecutive
// particles of one column so every global access except the ancestor gather is a coalesced 128-bit transaction.
#pragma once
#include <type_traits>
#include "common.cuh"
#include "philox.h"

#define ST_NT 256
#define ST_VEC 4
#ifndef SMCB_ST_MINB
#define SMCB_ST_MINB 4   // resident step-kernel blocks per SM the register budget is sized for (one wave = 148 * SMCB_ST_MINB blocks)
#endif

struct StepArgs {
  int64_t n, ld;
  int32_t B, blocks_per_col, iters;  // every block runs `iters` strided chunks of ST_NT*ST_VEC particles
  const float* P;                    // (B, SMCB_NPARAM)
  float* xbuf[2];                    // ping-pong state buffers, each (D, B, ld); the live one is xbuf[ctrl->t & 1]
  float* lw;                         // (B, ld) log-weights of the CURRENT state (input of a move)
  float* rw;                         // (B, ld) APF resampling log-weights g + lw of the current state
  float* lw_out;                     // (B, ld) the rows a MOVE writes: the weight rows ping-pong with the move index like the state buffers,
  float* rw_out;                     //         because move_kernel reads a tile's weights while other tiles already store new ones
  const int32_t* anc;                // (B, ld)
  int32_t* prev_inds;                // (B, ld) ancestors of the latest move as the API reports them (sisr.py:32 / apf.py:46)
  ColStats* stats;                   // (B)
  Partial* partials;                 // (B, blocks_per_col)
  Ctrl* ctrl;
  int32_t* col_ticket;               // (B) last-block-done counters of the fused finalize
  const float* eps_in;               // optional injected N(0,1) draws (D, B, ld)
  float* eps_out;                    // optional dump of the draws used
  uint64_t seed;
  int32_t fold;                      // APF: fold the next look-ahead weight into rw when y_{t+1} is known
  int32_t store_lw;                  // also store lw_t when folding (API-visible weights)
  int32_t sample_x0;                 // state_kernel: draw x_0 from the initial distribution
  float ess_threshold;               // relative threshold (filters/particle/base.py:42)
  // history (finalize)
  float* hist_mean; float* hist_var; float* hist_ll;  // (rows, B, D), (rows, B, D), (rows, B)
  int32_t hist_rows;
  float* latest_mean; float* latest_var; float* latest_ll; float* ll_total;  // (B, D), (B, D), (B), (B)
  int32_t fin_mode;
  int32_t fin_host;                  // finalize_kernel: the move index and the observation pointers are the ones below (t_host, y_t, y_next) - no
                                     // chain of dependent loads (ctrl->t -> ctrl->y -> y[t]) in front of the fold
  long long* dbg;                    // optional diagnostics (SMCB_DEBUG_TIMELINE)
  // step_kernel only: what the host knows at launch time, so that the prologue has no chain of dependent loads (ctrl->t -> ctrl->y -> y[t])
  int32_t t_host;                    // == ctrl->t when the kernel runs
  const float* y_t;                  // observation of this move (NULL: none)
  const float* y_next;               // observation of the next move (NULL: unknown -> no folded look-ahead)
  uint32_t pkeys[20];                // Philox round keys of `seed` (philox_round_keys)
  ExchangeArgs xch;                  // peer-memory exchange of the per-column log-likelihoods (common.cuh); xch.seq == 0: none
  int32_t col0;                      // global index of column 0 (smcb_config.column_offset): the Philox counters use col0 + column
  // NestedProposal (proposals/nested.py): optional injected draws - the inner samples' N(0,1) values (M, D, B, ld) and the float64
  // uniform of every particle's categorical draw (B, ld); NULL: Philox
  const float* nest_z;
  const float* nest_e;
};
__device__ __forceinline__ long long st_now() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
enum { FIN_STATE = 0, FIN_PREWEIGHT = 1, FIN_STEP = 2 };

// ---- soft-max accumulators ---------------------------------------------------------------------------------------------
template <int K>
struct SoftAcc {  // running max m and K sums of exp(v - m) * payload
  float m;
  float s[K];
  __device__ __forceinline__ void init() {
    m = -INFINITY;
#pragma unroll
    for (int k = 0; k < K; ++k) s[k] = 0.f;
  }
  __device__ __forceinline__ float raise(float nm) {  // make `nm` the reference point; returns exp(old - new)
    float sc = 1.f;
    if (nm > m) {
      sc = (m == -INFINITY) ? 0.f : __expf(m - nm);
#pragma unroll
      for (int k = 0; k < K; ++k) s[k] *= sc;
      m = nm;
    }
    return sc;
  }
  __device__ __forceinline__ void merge(const SoftAcc& o) {
    float nm = fmaxf(m, o.m);
    float sa = (m == -INFINITY) ? 0.f : __expf(m - nm);
    float sb = (o.m == -INFINITY) ? 0.f : __expf(o.m - nm);
#pragma unroll
    for (int k = 0; k < K; ++k) s[k] = s[k] * sa + o.s[k] * sb;
    m = nm;
  }
};

template <int K>
__device__ __forceinline__ SoftAcc<K> softacc_shfl_xor(const SoftAcc<K>& a, int o) {
  SoftAcc<K> r;
  r.m = __shfl_xor_sync(0xffffffffu, a.m, o);
#pragma unroll
  for (int k = 0; k < K; ++k) r.s[k] = __shfl_xor_sync(0xffffffffu, a.s[k], o);
  return r;
}

// block-wide merge; the result is valid in thread 0.  `scratch` holds (NT/32) records.
// Two phases so that only ONE exp per thread is spent: block max of the reference points, rescale, plain sums.
template <int K>
__device__ __forceinline__ void softacc_block_reduce(SoftAcc<K>& a, SoftAcc<K>* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float mb = a.m;
#pragma unroll
  for (int o = 16; o; o >>= 1) mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, o));
  __syncthreads();
  if (lane == 0) scratch[wid].m = mb;
  __syncthreads();
#pragma unroll
  for (int w = 0; w < ST_NT / 32; ++w) mb = fmaxf(mb, scratch[w].m);
  const float sc = (a.m == -INFINITY) ? 0.f : __expf(a.m - mb);
#pragma unroll
  for (int k = 0; k < K; ++k) {
    float v = a.s[k] * sc;
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    a.s[k] = v;
  }
  a.m = mb;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) scratch[wid].s[k] = a.s[k];
  }
  __syncthreads();
  if (wid == 0) {  // same fold as softacc4_block_reduce (one lane per warp, butterfly): the pre-weight pass and the folded look-ahead
                   // of the step kernel must produce the same normaliser bit for bit
#pragma unroll
    for (int k = 0; k < K; ++k) a.s[k] = warp_sum((lane < ST_NT / 32) ? scratch[lane].s[k] : 0.f);
  }
}

// the four accumulators of the step in one pass (three barriers instead of twelve); results valid in thread 0
template <int K, int NT = ST_NT>
struct Fin4Scratch {
  float m[4][NT / 32];
  float a[K][NT / 32];
  float q[NT / 32], r2[NT / 32], r3[NT / 32];
};
template <int K, int NT = ST_NT, bool PROTECT = true>  // PROTECT = false: the scratch area is known to be idle (saves a barrier)
__device__ __forceinline__ void softacc4_block_reduce(SoftAcc<K>& A, SoftAcc<1>& Q, SoftAcc<1>& R2, SoftAcc<1>& R3, Fin4Scratch<K, NT>& sc) {
  // Cost per thread is independent of the block size: the NT/32 per-warp values are folded by every warp with one lane per value
  // (redux for the maxima) instead of a loop over the warps in every thread.
  static_assert(NT <= 1024, "one lane per warp");
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float mb[4] = {A.m, Q.m, R2.m, R3.m};
#pragma unroll
  for (int i = 0; i < 4; ++i) mb[i] = warp_redux_max(mb[i]);
  if (PROTECT) __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) sc.m[i][wid] = mb[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) mb[i] = warp_redux_max((lane < NT / 32) ? sc.m[i][lane] : -INFINITY);
  const float sA = (A.m == -INFINITY) ? 0.f : __expf(A.m - mb[0]);
  const float sQ = (Q.m == -INFINITY) ? 0.f : __expf(Q.m - mb[1]);
  const float s2 = (R2.m == -INFINITY) ? 0.f : __expf(R2.m - mb[2]);
  const float s3 = (R3.m == -INFINITY) ? 0.f : __expf(R3.m - mb[3]);
  float va[K];
#pragma unroll
  for (int k = 0; k < K; ++k) va[k] = warp_sum(A.s[k] * sA);
  const float vq = warp_sum(Q.s[0] * sQ), v2 = warp_sum(R2.s[0] * s2), v3 = warp_sum(R3.s[0] * s3);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) sc.a[k][wid] = va[k];
    sc.q[wid] = vq; sc.r2[wid] = v2; sc.r3[wid] = v3;
  }
  __syncthreads();
  A.m = mb[0]; Q.m = mb[1]; R2.m = mb[2]; R3.m = mb[3];
  if (wid == 0) {  // the warp of thread 0 folds the per-warp sums, one lane per warp
    const bool have = lane < NT / 32;
#pragma unroll
    for (int k = 0; k < K; ++k) A.s[k] = warp_sum(have ? sc.a[k][lane] : 0.f);
    Q.s[0] = warp_sum(have ? sc.q[lane] : 0.f);
    R2.s[0] = warp_sum(have ? sc.r2[lane] : 0.f);
    R3.s[0] = warp_sum(have ? sc.r3[lane] : 0.f);
  }
}

// set 1 over lw: [0] sum e, [1..D] sum e (x - shift), [D+1..2D] sum e (x - shift)^2 ; sum e^2 kept apart (scales with sc^2)
template <int D>
struct Moments {
  SoftAcc<1 + 2 * D> a;
  SoftAcc<1> q;  // reference point 2*m: sum e^2 = sum exp(2 lw - 2 m)
  __device__ __forceinline__ void init() { a.init(); q.init(); }
  __device__ __forceinline__ void add4(const float (&lw)[4], const float (&x)[D][4], const float* shift, const bool (&valid)[4]) {
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 4; ++k) if (valid[k]) mx = fmaxf(mx, lw[k]);
    if (mx == -INFINITY) return;
    a.raise(mx);
    q.raise(2.f * mx);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!valid[k]) continue;
      float e = __expf(lw[k] - a.m);
      a.s[0] += e;
      q.s[0] += e * e;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        float c = x[d][k] - shift[d];
        a.s[1 + d] += e * c;
        a.s[1 + D + d] += e * c * c;
      }
    }
  }
};

// ---- proposals -------------------------------------------------------------------------------------------------------------
// Everything the step needs from one particle: new state, weight increment log p(y|x') (+ proposal correction), and the
// look-ahead weight of the ancestor that the APF subtracts (apf.py:43) - recomputed from the gathered ancestor state.
template <int MODEL, int PROP>
struct Proposal;

template <int MODEL>
struct Proposal<MODEL, SMCB_PROPOSAL_BOOTSTRAP> {
  typedef Model<MODEL> M;
  // proposals/base.py:69-85 + pre_weight_funcs.py:9-11:  log p(y | loc(x))
  __device__ static __forceinline__ float pre_weight(const float* y, const float* x, const float* P) {
    float loc[M::D], sc;
    M::loc_scale(x, P, loc, sc);
    return M::obs_lp(y, loc, P);
  }
  // proposals/bootstrap.py:10-14
  __device__ static __forceinline__ void sample_and_weight(const float* y, const float* xa, const float* z, const float* P,
                                                           bool observed, float* xn, float& inc, float& g_anc) {
    float loc[M::D], sc;
    M::loc_scale(xa, P, loc, sc);
#pragma unroll
    for (int d = 0; d < M::D; ++d) xn[d] = __fadd_rn(loc[d], __fmul_rn(sc, __fmul_rn(z[d], P[P_INC_SCALE])));
    inc = 0.f; g_anc = 0.f;
    if (observed) {
      inc = M::obs_lp(y, xn, P);
      g_anc = M::obs_lp(y, loc, P);
    }
  }
};

template <int MODEL, int PROP, bool VECTOR = (Model<MODEL>::D > 1)>
struct ProposalLGO;
template <int MODEL>
struct Proposal<MODEL, SMCB_PROPOSAL_LINEAR_GAUSS> : ProposalLGO<MODEL, SMCB_PROPOSAL_LINEAR_GAUSS> {};

// multi-dimensional LinearGaussianObservations for the Lorenz model (proposals/linear.py:38-86, proposals/utils.py:219-267 with
// A = a [[1,0,0],[0,0,1]], examples/lorenz.ipynb:214): every matrix of find_optimal_density is diagonal, so the 3x3 inverse and the
// Cholesky factor are per-column constants (models.h: P_LGO_*), the kernel N(k, P) factorises over the coordinates and
// MultivariateNormal.log_prob is the sum of three scalar normal log-densities.
template <int MODEL>
struct ProposalLGO<MODEL, SMCB_PROPOSAL_LINEAR_GAUSS, true> {
  typedef Model<MODEL> M;
  static_assert(M::D == 3 && M::OD == 2, "vector LinearGaussianObservations is compiled for the Lorenz-63 model");
  // log N(y; A x_{t-1}, diag(s^2 + a^2 sigma^2))   (proposals/linear.py:57-86; centred on the previous state)
  __device__ static __forceinline__ float pre_weight(const float* y, const float* x, const float* P) {
    const float l0 = smcb_normal_lp(y[0], __fmul_rn(P[5], x[0]), P[P_LGO_PRE_INV2VAR], P[P_LGO_PRE_LOGNORM]);
    const float l1 = smcb_normal_lp(y[1], __fmul_rn(P[5], x[2]), P[P_LGO_PRE_INV2VAR], P[P_LGO_PRE_LOGNORM]);
    return __fadd_rn(l0, l1);
  }
  __device__ static __forceinline__ void sample_and_weight(const float* y, const float* xa, const float* z, const float* P,
                                                           bool observed, float* xn, float& inc, float& g_anc) {
    float m[3], sc;
    M::loc_scale(xa, P, m, sc);
    inc = 0.f; g_anc = 0.f;
    if (!observed) {
#pragma unroll
      for (int d = 0; d < 3; ++d) xn[d] = __fadd_rn(m[d], __fmul_rn(sc, __fmul_rn(z[d], P[P_INC_SCALE])));
      return;
    }
    // k = P (sigma^-2 m + A^T s^-2 y)
    float k[3];
    const float yo[3] = {y[0], 0.f, y[1]};
    float x_lp = 0.f, k_lp = 0.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float t1 = __fmul_rn(P[P_LGO_HVI], m[d]);
      if (d == 1) {
        k[d] = __fmul_rn(P[P_LGO_COV1], t1);
        xn[d] = __fadd_rn(k[d], __fmul_rn(P[P_LGO_KSTD1], z[d]));
        k_lp = __fadd_rn(k_lp, smcb_normal_lp(xn[d], k[d], P[P_LGO_K1_INV2VAR], P[P_LGO_K1_LOGNORM]));
      } else {
        const float t3 = __fmul_rn(P[5], __fmul_rn(P[P_LGO_OVI], yo[d]));
        k[d] = __fmul_rn(P[P_LGO_COV], __fadd_rn(t1, t3));
        xn[d] = __fadd_rn(k[d], __fmul_rn(P[P_LGO_KSTD], z[d]));
        k_lp = __fadd_rn(k_lp, smcb_normal_lp(xn[d], k[d], P[P_LGO_K_INV2VAR], P[P_LGO_K_LOGNORM]));
      }
      const float e = __fmul_rn(__fsub_rn(xn[d], m[d]), P[P_LGO_INV_SIGMA]);
      x_lp = __fadd_rn(x_lp, smcb_normal_lp(e, 0.f, P[P_LGO_INC_INV2VAR], P[P_LGO_INC_LOGNORM]));
    }
    const float y_lp = M::obs_lp(y, xn, P);
    inc = __fsub_rn(__fadd_rn(y_lp, x_lp), k_lp);
    g_anc = pre_weight(y, xa, P);
  }
};

template <int MODEL>
struct ProposalLGO<MODEL, SMCB_PROPOSAL_LINEAR_GAUSS, false> {
  typedef Model<MODEL> M;
  static_assert(M::LINEAR_OBS && M::D == 1, "LinearGaussianObservations needs y = b + a x + s nu with a scalar state");
  // proposals/linear.py:57-86: log N(y; b + a x_{t-1}, sqrt(s^2 + a^2 sigma^2))   (centred on the previous state)
  __device__ static __forceinline__ float pre_weight(const float* y, const float* x, const float* P) {
    return smcb_normal_lp(y[0], __fadd_rn(P[P_OBS_B], __fmul_rn(P[P_OBS_A], x[0])), P[P_LGO_PRE_INV2VAR], P[P_LGO_PRE_LOGNORM]);
  }
  // proposals/linear.py:38-55, proposals/utils.py:219-267, proposals/base.py:45-50
  __device__ static __forceinline__ void sample_and_weight(const float* y, const float* xa, const float* z, const float* P,
                                                           bool observed, float* xn, float& inc, float& g_anc) {
    float m[1], sc;
    M::loc_scale(xa, P, m, sc);
    inc = 0.f; g_anc = 0.f;
    if (!observed) {  // particle/state.py:38-42: missing observations propagate through the dynamics
      xn[0] = __fadd_rn(m[0], __fmul_rn(sc, __fmul_rn(z[0], P[P_INC_SCALE])));
      return;
    }
    float t1 = __fmul_rn(P[P_LGO_HVI], m[0]);
    float t3 = __fmul_rn(P[P_OBS_A], __fmul_rn(P[P_LGO_OVI], __fsub_rn(y[0], P[P_OBS_B])));
    float k = __fmul_rn(P[P_LGO_COV], __fadd_rn(t1, t3));
    xn[0] = __fadd_rn(__fmul_rn(z[0], P[P_LGO_KSTD]), k);
    float y_lp = M::obs_lp(y, xn, P);
    float e = __fmul_rn(__fsub_rn(xn[0], m[0]), P[P_LGO_INV_SIGMA]);
    float x_lp = smcb_normal_lp(e, 0.f, P[P_LGO_INC_INV2VAR], P[P_LGO_INC_LOGNORM]);
    float k_lp = smcb_normal_lp(xn[0], k, P[P_LGO_K_INV2VAR], P[P_LGO_K_LOGNORM]);
    inc = __fsub_rn(__fadd_rn(y_lp, x_lp), k_lp);
    g_anc = pre_weight(y, xa, P);
  }
};

// Linearized (proposals/linearized.py:53-70 with ModeFinder.find_mode, proposals/utils.py:96-146, the default functorch path, as
// written): from x = mean, n_steps times x += step, step = the CONSTANT alpha for the first-order variant (the gradient is evaluated
// there but not used: utils.py:119,137) or cov * gradient with cov = -(H - clip(2 H, 0))^-1 for use_second_order (vector state:
// -pinv(H - clip(2 lambda_min, 0) I); the Hessians of the zoo's models are diagonal, so eigenvalues, pseudo-inverse and Cholesky factor
// are per coordinate); gradient and H are those of log p(y | x) + log p(x | x_prev) - closed forms here (Model::obs_grad_hess),
// automatic differentiation there.  Kernel N(x, std), std = the hidden scale (first order) / sqrt(cov) of the last step.
template <int MODEL>
struct Proposal<MODEL, SMCB_PROPOSAL_LINEARIZED> {
  typedef Model<MODEL> M;
  __device__ static __forceinline__ float pre_weight(const float* y, const float* x, const float* P) {   // proposals/base.py:69-85
    float loc[M::D], sc;
    M::loc_scale(x, P, loc, sc);
    return M::obs_lp(y, loc, P);
  }
  __device__ static __forceinline__ void sample_and_weight(const float* y, const float* xa, const float* z, const float* P,
                                                           bool observed, float* xn, float& inc, float& g_anc) {
    constexpr int D = M::D;
    float m[D], sc;
    M::loc_scale(xa, P, m, sc);
    inc = 0.f; g_anc = 0.f;
    if (!observed) {
#pragma unroll
      for (int d = 0; d < D; ++d) xn[d] = __fadd_rn(m[d], __fmul_rn(sc, __fmul_rn(z[d], P[P_INC_SCALE])));
      return;
    }
    float x[D], sd[D];
#pragma unroll
    for (int d = 0; d < D; ++d) { x[d] = m[d]; sd[d] = sc; }
    const int steps = (int)P[P_LIN_STEPS];
    const bool second = P[P_LIN_SECOND] != 0.f;
    const float tiv = P[P_LIN_T_INVVAR];
    for (int it = 0; it < steps; ++it) {
      if (second) {
        float g[D], h[D];
        M::obs_grad_hess(y, x, P, g, h);
        float lam = INFINITY;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          g[d] = __fsub_rn(g[d], __fmul_rn(__fsub_rn(x[d], m[d]), tiv));
          h[d] = __fsub_rn(h[d], tiv);
          lam = fminf(lam, h[d]);
        }
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const float dh = fmaxf(__fmul_rn(2.0f, (D == 1) ? h[d] : lam), 0.f);
          const float den = __fsub_rn(h[d], dh);
          const float cov = (den == 0.f && D > 1) ? 0.f : -__frcp_rn(den);   // pinv of a diagonal matrix; the scalar case divides as written
          x[d] = __fadd_rn(x[d], __fmul_rn(cov, g[d]));
          sd[d] = __fsqrt_rn(cov);
        }
      } else {
#pragma unroll
        for (int d = 0; d < D; ++d) x[d] = __fadd_rn(x[d], P[P_LIN_ALPHA]);
      }
    }
    float x_lp = 0.f, k_lp = 0.f;
    const float rsc = __frcp_rn(sc);
#pragma unroll
    for (int d = 0; d < D; ++d) {
      xn[d] = __fadd_rn(x[d], __fmul_rn(sd[d], z[d]));
      const float e = __fsub_rn(xn[d], m[d]);
      x_lp = __fadd_rn(x_lp, __fsub_rn(__fmul_rn(-__fmul_rn(e, e), P[P_LIN_T_INV2VAR]), P[P_LIN_T_LOGNORM]));
      k_lp = __fadd_rn(k_lp, smcb_normal_lp_scale(xn[d], x[d], sd[d]));
    }
    (void)rsc;
    inc = __fsub_rn(__fadd_rn(M::obs_lp(y, xn, P), x_lp), k_lp);
    g_anc = pre_weight(y, xa, P);
  }
};

// NestedProposal (proposals/nested.py:27-47, Naesseth et al.): M inner samples from the transition density per particle; the proposed
// particle is one of them, drawn from Categorical(softmax of their observation log-densities), the weight is log mean exp of those
// log-densities.  torch semantics reproduced: nan_to_num(nan = -inf, posinf = -inf) on the log-densities, a float32 soft-max accumulated
// in sample order, NaN probabilities -> 1 / M, and torch.multinomial's single-sample rule for the draw (argmax of probs / Exp(1)) when the
// exponentials are injected; the library's own draw inverts the prefix sums at one uniform.  The inner samples are never stored: their
// log-densities sit in a local array, the chosen sample is regenerated from its Philox counter (particle, column, move, group of four).
#define SMCB_RNG_NESTED 0x100u
#define SMCB_RNG_NESTED_U 0x80u
template <int MODEL>
struct Proposal<MODEL, SMCB_PROPOSAL_NESTED> {
  typedef Model<MODEL> M;
  __device__ static __forceinline__ float pre_weight(const float* y, const float* x, const float* P) {   // proposals/base.py:69-85
    float loc[M::D], sc;
    M::loc_scale(x, P, loc, sc);
    return M::obs_lp(y, loc, P);
  }
  // standard normals of the four inner samples 4 g .. 4 g + 3 of particle i: one Philox block per state dimension
  __device__ static __forceinline__ void inner_normals4(const StepArgs& a, int col, int64_t i, int t, int g, int Ms, float (&z)[M::D][4]) {
#pragma unroll
    for (int d = 0; d < M::D; ++d) {
      if (a.nest_z) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int s = min(4 * g + q, Ms - 1);
          z[d][q] = a.nest_z[(((int64_t)s * M::D + d) * a.B + col) * a.ld + i];
        }
      } else {
        const Philox4 r = philox4x32_10_keys((uint32_t)i, (uint32_t)(col + a.col0), (uint32_t)t, SMCB_RNG_NESTED + (uint32_t)g * 4u + d, a.pkeys);
        smcb_normal4(r, z[d]);
      }
    }
  }
  __device__ static __noinline__ void sample_and_weight(const StepArgs& a, int col, int64_t i, int t, const float* y, const float* xa,
                                                        const float* zk, const float* P, bool observed, float* xn, float& inc, float& g_anc) {
    constexpr int D = M::D;
    float m[D], sc;
    M::loc_scale(xa, P, m, sc);
    inc = 0.f; g_anc = 0.f;
    if (!observed) {
#pragma unroll
      for (int d = 0; d < D; ++d) xn[d] = __fadd_rn(m[d], __fmul_rn(sc, __fmul_rn(zk[d], P[P_INC_SCALE])));
      return;
    }
    const int Ms = min((int)P[P_NESTED_M], SMCB_NESTED_MAX);
    float lp[SMCB_NESTED_MAX];
    float mx = -INFINITY;
    for (int g = 0; 4 * g < Ms; ++g) {
      float z[D][4];
      inner_normals4(a, col, i, t, g, Ms, z);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float xs[D];
#pragma unroll
        for (int d = 0; d < D; ++d) xs[d] = __fadd_rn(m[d], __fmul_rn(sc, __fmul_rn(z[d][q], P[P_INC_SCALE])));
        float l = M::obs_lp(y, xs, P);
        if (l != l || l == INFINITY) l = -INFINITY;          // nan_to_num(-inf, -inf): -inf itself becomes the lowest finite float
        else if (l == -INFINITY) l = -SMCB_FLT_MAX;
        if (4 * g + q < Ms) {
          lp[4 * g + q] = l;
          mx = fmaxf(mx, l);
        }
      }
    }
    // soft-max over the samples (float32, in sample order) and the mean of exp(log-density)
    float se = 0.f, sw = 0.f;
    for (int s = 0; s < Ms; ++s) {
      se += expf(lp[s] - mx);
      sw += expf(lp[s]);
    }
    inc = logf(sw / (float)Ms);                              // log_prob.exp().mean(dim=0).log()
    // Categorical(probs).sample(): torch.multinomial draws ONE sample per row as argmax(probs / Exp(1)) (its n_sample == 1 path); with
    // injected exponentials that rule is followed, the library's own draw inverts the prefix sums at one Philox uniform
    const float fill = 1.0f / (float)Ms;
    int best = Ms - 1;
    if (a.nest_e) {
      float top = -INFINITY;
      best = 0;
      for (int s = 0; s < Ms; ++s) {
        float pr = __fdiv_rn(expf(lp[s] - mx), se);
        if (pr != pr) pr = fill;
        const float q = __fdiv_rn(pr, a.nest_e[((int64_t)s * a.B + col) * a.ld + i]);
        if (q > top) { top = q; best = s; }
      }
    } else {
      const Philox4 r = philox4x32_10_keys((uint32_t)i, (uint32_t)(col + a.col0), (uint32_t)t, SMCB_RNG_NESTED_U, a.pkeys);
      const double U = smcb_u01_double(r.x, r.y);
      float total = 0.f;
      for (int s = 0; s < Ms; ++s) {
        float pr = expf(lp[s] - mx) / se;
        if (pr != pr) pr = fill;
        total = __fadd_rn(total, pr);
      }
      float c = 0.f;
      for (int s = 0; s < Ms; ++s) {
        float pr = expf(lp[s] - mx) / se;
        if (pr != pr) pr = fill;
        c = __fadd_rn(c, pr);
        if ((double)__fdiv_rn(c, total) >= U) { best = s; break; }
      }
    }
    float z[D][4];
    inner_normals4(a, col, i, t, best >> 2, Ms, z);
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const int q = best & 3;
      const float zb = q == 0 ? z[d][0] : q == 1 ? z[d][1] : q == 2 ? z[d][2] : z[d][3];
      xn[d] = __fadd_rn(m[d], __fmul_rn(sc, __fmul_rn(zb, P[P_INC_SCALE])));
    }
    g_anc = pre_weight(y, xa, P);
  }
};

// one entry point for every proposal: the nested one draws its own inner samples and needs to know which particle it is working on
template <int MODEL, int PROP>
__device__ __forceinline__ void prop_sample_and_weight(const StepArgs& a, int col, int64_t i, int t, const float* y, const float* xa,
                                                       const float* zk, const float* P, bool observed, float* xn, float& inc, float& g_anc) {
  if constexpr (PROP == SMCB_PROPOSAL_NESTED) Proposal<MODEL, PROP>::sample_and_weight(a, col, i, t, y, xa, zk, P, observed, xn, inc, g_anc);
  else Proposal<MODEL, PROP>::sample_and_weight(y, xa, zk, P, observed, xn, inc, g_anc);
}

// ---- helpers -----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ const float* st_obs(const Ctrl* c, int t, int od) {
  int k = t - c->y_base;
  return (c->y && k >= 0 && k < c->y_count) ? c->y + (int64_t)k * od : nullptr;
}
template <int OD>
__device__ __forceinline__ bool st_load_obs(const float* p, float* y) {  // false when missing or all-NaN (filters/base.py:213)
  if (!p) return false;
  bool all_nan = true;
#pragma unroll
  for (int d = 0; d < OD; ++d) { y[d] = p[d]; all_nan = all_nan && (y[d] != y[d]); }
  return !all_nan;
}

template <int D, bool PLAIN = false>  // PLAIN: the caller knows that no draws are injected or dumped
__device__ __forceinline__ void st_noise4(const StepArgs& a, int col, int64_t i0, int t, uint32_t purpose, float (&z)[D][4]) {
#pragma unroll
  for (int d = 0; d < D; ++d) {
    if (!PLAIN && a.eps_in) {
      const float4 q = *reinterpret_cast<const float4*>(a.eps_in + ((int64_t)d * a.B + col) * a.ld + i0);
      z[d][0] = q.x; z[d][1] = q.y; z[d][2] = q.z; z[d][3] = q.w;
    } else {
      Philox4 r = philox4x32_10_keys((uint32_t)(i0 >> 2), (uint32_t)(col + a.col0), (uint32_t)t, purpose + d, a.pkeys);
      smcb_normal4(r, z[d]);
    }
    if (!PLAIN && a.eps_out)
      *reinterpret_cast<float4*>(a.eps_out + ((int64_t)d * a.B + col) * a.ld + i0) = make_float4(z[d][0], z[d][1], z[d][2], z[d][3]);
  }
}

__device__ __forceinline__ void st_write_partial1(Partial& p, const SoftAcc<3>& a, const SoftAcc<1>& q) {
  p.m1 = a.m; p.z1 = a.s[0]; p.sx[0] = a.s[1]; p.sxx[0] = a.s[2]; p.sx[1] = p.sx[2] = p.sxx[1] = p.sxx[2] = 0.f;
  // q holds sum exp(2 lw - q.m) with q.m == 2 a.m up to rounding; express it relative to 2*m1
  p.zz1 = (q.m == -INFINITY) ? 0.f : q.s[0] * __expf(q.m - 2.f * a.m);
}
__device__ __forceinline__ void st_write_partial1(Partial& p, const SoftAcc<7>& a, const SoftAcc<1>& q) {
  p.m1 = a.m; p.z1 = a.s[0];
#pragma unroll
  for (int d = 0; d < 3; ++d) { p.sx[d] = a.s[1 + d]; p.sxx[d] = a.s[4 + d]; }
  p.zz1 = (q.m == -INFINITY) ? 0.f : q.s[0] * __expf(q.m - 2.f * a.m);
}

// ---- state kernel: x_0 ~ p_0 (filters/particle/base.py:87-103) or a caller-supplied state; partials of lw --------------------
template <int MODEL>
__global__ void __launch_bounds__(ST_NT) state_kernel(StepArgs a) {
  typedef Model<MODEL> M;
  constexpr int D = M::D;
  __shared__ float Ps[SMCB_NPARAM];
  __shared__ SoftAcc<1 + 2 * D> sA[ST_NT / 32];
  __shared__ SoftAcc<1> sQ[ST_NT / 32];
  const int col = blockIdx.y, tid = threadIdx.x;
  if (tid < SMCB_NPARAM) Ps[tid] = a.P[(int64_t)col * SMCB_NPARAM + tid];
  __syncthreads();
  const int t = a.ctrl->t;
  float* xcur = a.xbuf[t & 1];
  float shift[D];
#pragma unroll
  for (int d = 0; d < D; ++d) shift[d] = a.sample_x0 ? Ps[P_X0_LOC + d] : a.stats[col].shift[d];
  if (a.sample_x0 && blockIdx.x == 0 && tid < D) a.stats[col].shift[tid] = Ps[P_X0_LOC + tid];  // the finalize kernel adds the shift back
  Moments<D> mom; mom.init();
  for (int it = 0; it < a.iters; ++it) {
    const int64_t i0 = ((int64_t)(it * a.blocks_per_col + blockIdx.x) * ST_NT + tid) * ST_VEC;
    if (i0 >= a.n) continue;
    float x[D][4], lw[4];
    bool valid[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) valid[k] = i0 + k < a.n;
    if (a.sample_x0) {
      float z[D][4];
      st_noise4<D>(a, col, i0, 0, SMCB_RNG_INIT, z);
#pragma unroll
      for (int d = 0; d < D; ++d) {
#pragma unroll
        for (int k = 0; k < 4; ++k) x[d][k] = __fadd_rn(Ps[P_X0_LOC + d], __fmul_rn(Ps[P_X0_SCALE + d], z[d][k]));
        *reinterpret_cast<float4*>(xcur + ((int64_t)d * a.B + col) * a.ld + i0) = make_float4(x[d][0], x[d][1], x[d][2], x[d][3]);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) lw[k] = 0.f;
      *reinterpret_cast<float4*>(a.lw + (int64_t)col * a.ld + i0) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<int4*>(a.prev_inds + (int64_t)col * a.ld + i0) = make_int4((int)i0, (int)i0 + 1, (int)i0 + 2, (int)i0 + 3);
    } else {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const float4 q = *reinterpret_cast<const float4*>(xcur + ((int64_t)d * a.B + col) * a.ld + i0);
        x[d][0] = q.x; x[d][1] = q.y; x[d][2] = q.z; x[d][3] = q.w;
      }
      float4 q = *reinterpret_cast<const float4*>(a.lw + (int64_t)col * a.ld + i0);
      lw[0] = smcb_sanitize(q.x); lw[1] = smcb_sanitize(q.y); lw[2] = smcb_sanitize(q.z); lw[3] = smcb_sanitize(q.w);
      *reinterpret_cast<float4*>(a.lw + (int64_t)col * a.ld + i0) = make_float4(lw[0], lw[1], lw[2], lw[3]);  // utils.py:57 mutates
    }
    mom.add4(lw, x, shift, valid);
  }
  softacc_block_reduce(mom.a, sA);
  softacc_block_reduce(mom.q, sQ);
  if (tid == 0) {
    Partial& p = a.partials[(int64_t)col * a.blocks_per_col + blockIdx.x];
    st_write_partial1(p, mom.a, mom.q);
    p.m2 = p.m3 = -INFINITY; p.z2 = p.z3 = 0.f;
  }
}

// ---- APF look-ahead weights when not folded: rw = lw + log p(y_t | .)  (apf.py:27-29) ----------------------------------------
template <int MODEL, int PROP>
__global__ void __launch_bounds__(ST_NT) preweight_kernel(StepArgs a) {
  typedef Model<MODEL> M;
  constexpr int D = M::D, OD = M::OD;
  __shared__ float Ps[SMCB_NPARAM];
  __shared__ SoftAcc<1> sR[ST_NT / 32];
  const int col = blockIdx.y, tid = threadIdx.x;
  if (tid < SMCB_NPARAM) Ps[tid] = a.P[(int64_t)col * SMCB_NPARAM + tid];
  __syncthreads();
  const int t = a.ctrl->t;
  float y[OD];
  const bool observed = st_load_obs<OD>(st_obs(a.ctrl, t, OD), y);
  const float* xcur = a.xbuf[t & 1];
  SoftAcc<1> r2; r2.init();
  if (observed) {
    for (int it = 0; it < a.iters; ++it) {
      const int64_t i0 = ((int64_t)(it * a.blocks_per_col + blockIdx.x) * ST_NT + tid) * ST_VEC;
      if (i0 >= a.n) continue;
      float x[D][4];
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const float4 q = *reinterpret_cast<const float4*>(xcur + ((int64_t)d * a.B + col) * a.ld + i0);
        x[d][0] = q.x; x[d][1] = q.y; x[d][2] = q.z; x[d][3] = q.w;
      }
      const float4 l = *reinterpret_cast<const float4*>(a.lw + (int64_t)col * a.ld + i0);
      float lw[4] = {l.x, l.y, l.z, l.w}, rw[4];
      float mx = -INFINITY;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float xs[D];
#pragma unroll
        for (int d = 0; d < D; ++d) xs[d] = x[d][k];
        rw[k] = smcb_sanitize(__fadd_rn(Proposal<MODEL, PROP>::pre_weight(y, xs, Ps), lw[k]));
        if (i0 + k < a.n) mx = fmaxf(mx, rw[k]);
      }
      *reinterpret_cast<float4*>(a.rw + (int64_t)col * a.ld + i0) = make_float4(rw[0], rw[1], rw[2], rw[3]);
      if (mx > -INFINITY) {
        r2.raise(mx);
#pragma unroll
        for (int k = 0; k < 4; ++k) if (i0 + k < a.n) r2.s[0] += __expf(rw[k] - r2.m);
      }
    }
  }
  softacc_block_reduce(r2, sR);
  if (tid == 0) {
    Partial& p = a.partials[(int64_t)col * a.blocks_per_col + blockIdx.x];
    p.m2 = r2.m; p.z2 = r2.s[0];
  }
}

// ---- finalize: folds the per-block partials of one column ------------------------------------------------------------------------
// FIN_STATE      after state_kernel:      normalisers/ESS/moments of the current weights; history row ctrl->t; no likelihood
// FIN_PREWEIGHT  after preweight_kernel:  normalisers of rw and the look-ahead term of apf.py:44
// FIN_STEP       after step_kernel:       everything, history row t+1, running log-likelihood, then t <- t+1
// Runs either as its own kernel (one block per column) or as the tail of the step kernel in the block that finishes a column last.
template <int D>
struct FinSmem {
  SoftAcc<1 + 2 * D> A[ST_NT / 32];
  SoftAcc<1> Q[ST_NT / 32];
  Fin4Scratch<1 + 2 * D> f4;
};

// what the finalizing thread needs besides the partials, loaded BEFORE the serial tail (every dependent global round trip
// of the tail costs about a microsecond)
struct FinPre {
  ColStats st;
  bool observed;   // y_t is present
  bool fold;       // y_{t+1} is present and the look-ahead is folded
  float ll_total;  // running log-likelihood before this move
};
template <int OD>
__device__ __forceinline__ FinPre fin_preload(const StepArgs& a, int col, int mode, int t) {
  FinPre p;
  p.st = a.stats[col];
  float y[OD];
  p.observed = (mode != FIN_STATE) && st_load_obs<OD>(st_obs(a.ctrl, t, OD), y);
  p.fold = (mode == FIN_STEP) && a.fold && st_load_obs<OD>(st_obs(a.ctrl, t + 1, OD), y);
  p.ll_total = a.ll_total[col];
  return p;
}

// The single-thread part of FIN_STATE / FIN_STEP: the folded sums of a column -> ColStats, moments, likelihood increment, history rows.
// Shared by finalize_column (multi-kernel pipeline) and column_kernel (one block owns the column).  Returns the new statistics in `st`
// (also stored to a.stats[col]) and the likelihood increment of the move.
template <int D, int OD, int ALG>
__device__ __forceinline__ float fin_apply(const StepArgs& a, int col, int mode, int t, const SoftAcc<1 + 2 * D>& A, const SoftAcc<1>& Q,
                                           const SoftAcc<1>& R2, const SoftAcc<1>& R3, const FinPre& pre, ColStats& st) {
  const float nf = (float)a.n;
  const bool observed = (mode == FIN_STEP) && pre.observed;
  const float ll_aux_prev = st.ll_aux;
  st.m_lw = A.m; st.z_lw = A.s[0]; st.inv_z_lw = 1.0f / A.s[0];
  // Q.m may differ from 2*A.m by rounding of the merges: bring sum e^2 to the reference point 2*A.m
  const float zz = (Q.m == -INFINITY) ? 0.f : Q.s[0] * __expf(Q.m - 2.f * A.m);
  st.ess = (A.s[0] * A.s[0]) / zz;
  float mean[D], var[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const float dm = A.s[1 + d] * st.inv_z_lw;            // E[x - shift]
    mean[d] = st.shift[d] + dm;
    var[d] = fmaxf(A.s[1 + D + d] * st.inv_z_lw - dm * dm, 0.f);
  }
  float ll = 0.f;
  if (mode == FIN_STEP && observed) {
    if (ALG == SMCB_ALG_SISR) ll = R3.m + logf(R3.s[0]);                       // filters/particle/utils.py:16-22
    else ll = (A.m + logf(A.s[0]) - logf(nf)) + ll_aux_prev;                   // apf.py:44
  }
  // next step's resampling decision and (APF) folded normalisers
  st.fold_valid = 0;
  if (ALG == SMCB_ALG_SISR) st.resample = (st.ess < a.ess_threshold * nf) ? 1 : 0;   // sisr.py:18-19
  else {
    st.resample = 0;  // set by the pre-weight pass unless the look-ahead was folded below
    const bool fold = pre.fold;
    if (fold) {
      st.m_rw = R2.m; st.z_rw = R2.s[0]; st.inv_z_rw = 1.0f / R2.s[0];
      st.ll_aux = logf(R2.s[0]) + (R2.m - A.m) - logf(A.s[0]);
      st.fold_valid = 1;
      st.resample = 1;
    }
  }
#pragma unroll
  for (int d = 0; d < D; ++d) st.shift[d] = mean[d];
  a.stats[col] = st;
  const int rowi = (mode == FIN_STEP) ? t + 1 : t;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    a.latest_mean[col * D + d] = mean[d];
    a.latest_var[col * D + d] = var[d];
    if (a.hist_mean && rowi < a.hist_rows) {
      a.hist_mean[((int64_t)rowi * a.B + col) * D + d] = mean[d];
      a.hist_var[((int64_t)rowi * a.B + col) * D + d] = var[d];
    }
  }
  a.latest_ll[col] = ll;
  if (mode == FIN_STEP) a.ll_total[col] = pre.ll_total + ll;
  if (a.hist_ll && rowi < a.hist_rows) a.hist_ll[(int64_t)rowi * a.B + col] = ll;
  return ll;
}

template <int D, int OD, int ALG>
__device__ __forceinline__ void finalize_column(const StepArgs& a, int col, int mode, int t, FinSmem<D>& fs, const FinPre& pre) {
  const int tid = threadIdx.x;
  SoftAcc<1 + 2 * D> A; A.init();
  SoftAcc<1> Q, R2, R3; Q.init(); R2.init(); R3.init();
  for (int b0 = tid; b0 < a.blocks_per_col; b0 += 4 * ST_NT) {
    // up to four records per thread, all loads in flight before the first merge (one L2 round trip instead of one per record)
    float4 v0[4], v1[4], v2[4];
    float v3[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int b = b0 + r * ST_NT;
      if (b < a.blocks_per_col) {  // L2 loads: the records were written by other blocks of a kernel that may still be running
        const Partial* pp = a.partials + (int64_t)col * a.blocks_per_col + b;
        const float4* q = reinterpret_cast<const float4*>(pp);
        v0[r] = __ldcg(q); v1[r] = __ldcg(q + 1); v2[r] = __ldcg(q + 2);
        v3[r] = __ldcg(reinterpret_cast<const float*>(pp) + 12);
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int b = b0 + r * ST_NT;
      if (b >= a.blocks_per_col) break;
      Partial p;
      p.m1 = v0[r].x; p.z1 = v0[r].y; p.zz1 = v0[r].z; p.sx[0] = v0[r].w; p.sx[1] = v1[r].x; p.sx[2] = v1[r].y; p.sxx[0] = v1[r].z;
      p.sxx[1] = v1[r].w; p.sxx[2] = v2[r].x; p.m2 = v2[r].y; p.z2 = v2[r].z; p.m3 = v2[r].w; p.z3 = v3[r];
      SoftAcc<1 + 2 * D> o; o.m = p.m1; o.s[0] = p.z1;
#pragma unroll
      for (int d = 0; d < D; ++d) { o.s[1 + d] = p.sx[d]; o.s[1 + D + d] = p.sxx[d]; }
      SoftAcc<1> q; q.m = (p.m1 == -INFINITY) ? -INFINITY : 2.f * p.m1; q.s[0] = p.zz1;
      SoftAcc<1> o2; o2.m = p.m2; o2.s[0] = p.z2;
      SoftAcc<1> o3; o3.m = p.m3; o3.s[0] = p.z3;
      A.merge(o); Q.merge(q); R2.merge(o2); R3.merge(o3);
    }
  }
  softacc4_block_reduce(A, Q, R2, R3, fs.f4);
  if (tid == 0) {
    ColStats st = pre.st;
    const float nf = (float)a.n;
    if (mode == FIN_PREWEIGHT) {
      const bool observed = pre.observed;
      if (observed) {
        st.m_rw = R2.m; st.z_rw = R2.s[0]; st.inv_z_rw = 1.0f / R2.s[0];
        st.ll_aux = logf(R2.s[0]) + (R2.m - st.m_lw) - logf(st.z_lw);   // log sum W exp(g), W = softmax(lw)
      }
      st.resample = observed ? 1 : 0;
      st.fold_valid = 1;
      a.stats[col] = st;
    } else {
      const float ll = fin_apply<D, OD, ALG>(a, col, mode, t, A, Q, R2, R3, pre, st);
      if (mode == FIN_STEP) smcb_exchange_publish(a.xch, col, ll, pre.ll_total + ll);
    }
    if (mode == FIN_STEP) {
      if (a.B == 1 || atomicAdd(&a.ctrl->ticket, 1) == a.B - 1) {  // every column is finalized, hence every block has read ctrl->t
        a.ctrl->ticket = 0;
        a.ctrl->t = t + 1;
      }
    }
  }
}

template <int D, int OD, int ALG>
__global__ void __launch_bounds__(ST_NT) finalize_kernel(StepArgs a) {
  __shared__ FinSmem<D> fs;
  pdl_trigger();  // (launched behind move_kernel with programmatic serialisation: the next move's blocks may take their places now)
  pdl_wait();
  if (a.fin_host) {  // everything the fold needs besides the partial records is known or loadable at once
    const int t = a.t_host;
    FinPre pre;
    float y[OD];
    pre.st = a.stats[blockIdx.x];
    pre.observed = (a.fin_mode != FIN_STATE) && st_load_obs<OD>(a.y_t, y);
    pre.fold = (a.fin_mode == FIN_STEP) && a.fold && st_load_obs<OD>(a.y_next, y);
    pre.ll_total = a.ll_total[blockIdx.x];
    finalize_column<D, OD, ALG>(a, blockIdx.x, a.fin_mode, t, fs, pre);
    return;
  }
  const int t = a.ctrl->t;
  const FinPre pre = fin_preload<OD>(a, blockIdx.x, a.fin_mode, t);
  finalize_column<D, OD, ALG>(a, blockIdx.x, a.fin_mode, t, fs, pre);
}

// ---- the fused step ----------------------------------------------------------------------------------------------------------
#define SMCB_L2E 1.4426950408889634f
__device__ __forceinline__ float st_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// nan_to_num of utils.py:57 for values that only feed the accumulators and the stored log-weights
__device__ __forceinline__ float st_sanitize(float w) { return (w != w || w == INFINITY) ? -INFINITY : fmaxf(w, -SMCB_FLT_MAX); }

// running soft-max sums of one thread: reference point m, sum e, sum e^2, sum e (x - shift), sum e (x - shift)^2, e = exp(lw - m)
template <int D>
struct StepAcc {
  float m, s0, q, sx[D], sxx[D];
  __device__ __forceinline__ void init() {
    m = -INFINITY; s0 = 0.f; q = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) { sx[d] = 0.f; sxx[d] = 0.f; }
  }
  __device__ __forceinline__ void add4(const float (&lw)[4], const float (&x)[D][4], const float (&shift)[D]) {
    const float mx = fmaxf(fmaxf(lw[0], lw[1]), fmaxf(lw[2], lw[3]));
    if (mx == -INFINITY) return;
    if (mx > m) {  // rare after the first few groups
      const float sc = (m == -INFINITY) ? 0.f : st_ex2((m - mx) * SMCB_L2E);
      s0 *= sc; q *= sc * sc;
#pragma unroll
      for (int d = 0; d < D; ++d) { sx[d] *= sc; sxx[d] *= sc; }
      m = mx;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float e = st_ex2((lw[k] - m) * SMCB_L2E);
      s0 += e;
      q = fmaf(e, e, q);
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const float c = x[d][k] - shift[d];
        const float ec = e * c;
        sx[d] += ec;
        sxx[d] = fmaf(ec, c, sxx[d]);
      }
    }
  }
  __device__ __forceinline__ void to_softacc(SoftAcc<1 + 2 * D>& A, SoftAcc<1>& Q) const {
    A.m = m; A.s[0] = s0;
#pragma unroll
    for (int d = 0; d < D; ++d) { A.s[1 + d] = sx[d]; A.s[1 + D + d] = sxx[d]; }
    Q.m = (m == -INFINITY) ? -INFINITY : 2.f * m;
    Q.s[0] = q;
  }
};
// sum of w * exp(v - m) with a running reference point (w = 1: plain soft-max sum)
struct StepAcc1 {
  float m, z;
  __device__ __forceinline__ void init() { m = -INFINITY; z = 0.f; }
  __device__ __forceinline__ void add4(const float (&v)[4], const float (&w)[4]) {
    const float mx = fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3]));
    if (mx == -INFINITY) return;
    if (mx > m) { z *= (m == -INFINITY) ? 0.f : st_ex2((m - mx) * SMCB_L2E); m = mx; }
#pragma unroll
    for (int k = 0; k < 4; ++k) z = fmaf(w[k], st_ex2((v[k] - m) * SMCB_L2E), z);
  }
  __device__ __forceinline__ void to_softacc(SoftAcc<1>& R) const { R.m = m; R.s[0] = z; }
};

template <int MODEL, int PROP, int ALG>
__global__ void __launch_bounds__(ST_NT, SMCB_ST_MINB) step_kernel(StepArgs a) {
  typedef Model<MODEL> M;
  constexpr int D = M::D, OD = M::OD;
  __shared__ float Ps[SMCB_NPARAM];
  __shared__ FinSmem<D> fin_smem;
  __shared__ FinPre fin_pre;
  SoftAcc<1 + 2 * D>* sA = fin_smem.A;
  SoftAcc<1>* sQ = fin_smem.Q;
  const int col = blockIdx.y, tid = threadIdx.x;
  if (tid < SMCB_NPARAM) Ps[tid] = a.P[(int64_t)col * SMCB_NPARAM + tid];  // parameters do not depend on the predecessor kernel
  pdl_wait();
  if (a.dbg && tid == 0) atomicMin((unsigned long long*)&a.dbg[11], (unsigned long long)st_now());
  __syncthreads();
  const int t = a.t_host;
  float y[OD], yn[OD];
  const bool observed = st_load_obs<OD>(a.y_t, y);
  const bool fold = (ALG == SMCB_ALG_APF) && a.fold && st_load_obs<OD>(a.y_next, yn);
  const ColStats st = a.stats[col];
  if (tid == 0) { fin_pre.st = st; fin_pre.observed = observed; fin_pre.fold = fold; fin_pre.ll_total = a.ll_total[col]; }  // for the finalizing block's tail
  if (a.dbg && tid == 0 && blockIdx.x == 0) a.dbg[14] = st_now();
  // SISR resamples when the ESS test fired (sisr.py:19-26), the APF on every observed step (apf.py:29-34, filters/base.py:213)
  const bool resampled = (ALG == SMCB_ALG_APF) ? observed : (st.resample != 0);
  const int32_t n = (int32_t)a.n;
  const float inv_n = 1.0f / (float)a.n;
  const float one4[4] = {1.f, 1.f, 1.f, 1.f};
  float shift[D];
  const float* xprev[D];
  float* xnext[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    shift[d] = st.shift[d];
    xprev[d] = a.xbuf[t & 1] + ((int64_t)d * a.B + col) * a.ld;
    xnext[d] = a.xbuf[(t + 1) & 1] + ((int64_t)d * a.B + col) * a.ld;
  }
  const int32_t* ancrow = a.anc + (int64_t)col * a.ld;
  int32_t* pirow = a.prev_inds + (int64_t)col * a.ld;
  const float* lwin = a.lw + (int64_t)col * a.ld;
  float* lwrow = a.lw_out + (int64_t)col * a.ld;
  float* rwrow = a.rw_out + (int64_t)col * a.ld;
  // keep the row pointers in registers: re-deriving them from (column, pitch, buffer index) inside the loop costs ~5 instructions per access
#pragma unroll
  for (int d = 0; d < D; ++d) { asm volatile("" : "+l"(xprev[d])); asm volatile("" : "+l"(xnext[d])); }
  asm volatile("" : "+l"(ancrow)); asm volatile("" : "+l"(pirow)); asm volatile("" : "+l"(lwrow)); asm volatile("" : "+l"(rwrow));
  asm volatile("" : "+l"(lwin));

  StepAcc<D> mom; mom.init();
  StepAcc1 r2; r2.init();   // APF: folded resampling weights
  StepAcc1 r3; r3.init();   // SISR: likelihood increment

  const int32_t stride = a.blocks_per_col * (ST_NT * ST_VEC);
  const int32_t ilast = (int32_t)a.ld - 4;  // last group of the padded row
  // The steady state (observed move, resampled, look-ahead folded, no injected or dumped noise) gets its own copy of the loop with
  // those facts as compile-time constants: the per-particle `if (observed)` / `fold ? .. : ..` are warp-uniform but the compiler
  // cannot know it and wraps each in a convergence region.
  const bool observed_rt = observed, fold_rt = fold, resampled_rt = resampled;
  auto run_loop = [&](auto fast_tag) {
    constexpr bool FAST = decltype(fast_tag)::value;
    const bool observed = FAST ? true : observed_rt;
    const bool fold = FAST ? (ALG == SMCB_ALG_APF) : fold_rt;
    const bool resampled = FAST ? true : resampled_rt;
    int32_t i0 = (blockIdx.x * ST_NT + tid) * ST_VEC;
    int4 ancq = make_int4(i0, i0 + 1, i0 + 2, i0 + 3);
    if (resampled && i0 < n) ancq = *reinterpret_cast<const int4*>(ancrow + i0);
    for (int it = 0; it < a.iters; ++it, i0 += stride) {
      if (i0 >= n) break;
      const bool full = i0 + 4 <= n;
      int anc[4] = {i0, i0 + 1, i0 + 2, i0 + 3};
      if (resampled) {  // fetch the next group's ancestors while this one is processed: clamped into the padded row, no branch per lane
        anc[0] = ancq.x; anc[1] = ancq.y; anc[2] = ancq.z; anc[3] = ancq.w;
        ancq = *reinterpret_cast<const int4*>(ancrow + min(i0 + stride, ilast));
      }
      if (resampled) {
        *reinterpret_cast<int4*>(pirow + i0) = make_int4(anc[0], anc[1], anc[2], anc[3]);
        if (!full) {
#pragma unroll
          for (int k = 0; k < 4; ++k) if (i0 + k >= n) anc[k] = 0;
        }
      } else if (ALG == SMCB_ALG_APF) {
        *reinterpret_cast<int4*>(pirow + i0) = make_int4(anc[0], anc[1], anc[2], anc[3]);  // apf.py:18-23 arange
      }
      float lwp[4] = {0.f, 0.f, 0.f, 0.f};
      if (!resampled) {  // weights carry over (sisr.py:52 without the reset of :34; particle/state.py:42)
        const float4 q = *reinterpret_cast<const float4*>(lwin + i0);
        lwp[0] = q.x; lwp[1] = q.y; lwp[2] = q.z; lwp[3] = q.w;
      }
      float xa[D][4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int d = 0; d < D; ++d) xa[d][k] = __ldg(xprev[d] + anc[k]);
      }
      float z[D][4];
      st_noise4<D, FAST>(a, col, i0, t, SMCB_RNG_TRANSITION, z);

      float xn[D][4], lwn[4], rwn[4], gnx[4], inc4[4], wprev[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float xk[D], zk[D], xo[D], inc, g_anc;
#pragma unroll
        for (int d = 0; d < D; ++d) { xk[d] = xa[d][k]; zk[d] = z[d][k]; }
        prop_sample_and_weight<MODEL, PROP>(a, col, (int64_t)i0 + k, t, y, xk, zk, Ps, observed, xo, inc, g_anc);
#pragma unroll
        for (int d = 0; d < D; ++d) xn[d][k] = xo[d];
        float lw;
        if (!observed) lw = lwp[k];
        else if (ALG == SMCB_ALG_APF) lw = __fsub_rn(inc, g_anc);   // apf.py:43
        else lw = __fadd_rn(inc, lwp[k]);                           // sisr.py:52
        lwn[k] = lw;
        inc4[k] = inc;
        if (ALG == SMCB_ALG_SISR) wprev[k] = resampled ? inv_n : smcb_weight(lwp[k], st.m_lw, st.inv_z_lw);
        gnx[k] = fold ? Proposal<MODEL, PROP>::pre_weight(yn, xo, Ps) : 0.f;
      }
      {  // nan_to_num (utils.py:57) only when something in the group is not finite: a sum of four finite floats can overflow at worst
        const float chk = fabsf(lwn[0]) + fabsf(lwn[1]) + fabsf(lwn[2]) + fabsf(lwn[3]);
        if (!(chk < INFINITY)) {
#pragma unroll
          for (int k = 0; k < 4; ++k) lwn[k] = st_sanitize(lwn[k]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) rwn[k] = __fadd_rn(gnx[k], lwn[k]);
        if (fold) {
          const float chk2 = fabsf(rwn[0]) + fabsf(rwn[1]) + fabsf(rwn[2]) + fabsf(rwn[3]);
          if (!(chk2 < INFINITY)) {
#pragma unroll
            for (int k = 0; k < 4; ++k) rwn[k] = st_sanitize(rwn[k]);
          }
        }
      }
#pragma unroll
      for (int d = 0; d < D; ++d) *reinterpret_cast<float4*>(xnext[d] + i0) = make_float4(xn[d][0], xn[d][1], xn[d][2], xn[d][3]);
      if (!fold || a.store_lw) *reinterpret_cast<float4*>(lwrow + i0) = make_float4(lwn[0], lwn[1], lwn[2], lwn[3]);
      if (fold) *reinterpret_cast<float4*>(rwrow + i0) = make_float4(rwn[0], rwn[1], rwn[2], rwn[3]);

      if (!full) {  // the tail of the column: padding contributes nothing
#pragma unroll
        for (int k = 0; k < 4; ++k) if (i0 + k >= n) { lwn[k] = -INFINITY; rwn[k] = -INFINITY; inc4[k] = -INFINITY; }
      }
      mom.add4(lwn, xn, shift);
      if (fold) r2.add4(rwn, one4);
      if (ALG == SMCB_ALG_SISR && observed) r3.add4(inc4, wprev);
    }
  };
  if (observed_rt && resampled_rt && (ALG != SMCB_ALG_APF || fold_rt) && !a.eps_in && !a.eps_out) run_loop(std::true_type{});
  else run_loop(std::false_type{});
  if (a.dbg && tid == 0) atomicMax((unsigned long long*)&a.dbg[15], (unsigned long long)st_now());
  SoftAcc<1 + 2 * D> A;
  SoftAcc<1> Q, R2, R3;
  mom.to_softacc(A, Q); r2.to_softacc(R2); r3.to_softacc(R3);
  softacc4_block_reduce(A, Q, R2, R3, fin_smem.f4);
  pdl_trigger();  // the successor may be scheduled while the last block folds the partials
  __shared__ int is_last;
  if (tid == 0) {
    Partial& p = a.partials[(int64_t)col * a.blocks_per_col + blockIdx.x];
    st_write_partial1(p, A, Q);
    p.m2 = R2.m; p.z2 = R2.s[0];
    p.m3 = R3.m; p.z3 = R3.s[0];
    __threadfence();
    is_last = (atomicAdd(&a.col_ticket[col], 1) == a.blocks_per_col - 1);
  }
  __syncthreads();
  if (is_last) {  // this block completed the column: fold the partials here instead of launching another kernel
    if (tid == 0) a.col_ticket[col] = 0;
    __threadfence();
    if (a.dbg && tid == 0) a.dbg[12] = st_now();
    finalize_column<D, OD, ALG>(a, col, FIN_STEP, t, fin_smem, fin_pre);
    if (a.dbg && tid == 0) a.dbg[13] = st_now();
  }
}
