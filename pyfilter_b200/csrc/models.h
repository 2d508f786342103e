// Compiled model zoo: the four state-space models of BASELINE.json, as device functions.
//
// In the reference the model is a Python callable (stochproc AffineProcess.mean_scale / StateSpaceModel.build_density,
// SURVEY.md Appendix C); a fused kernel cannot call that, so each model is a struct with
//     loc_scale(x_prev, P) -> (loc[D], scale)        x_t = loc + scale * (inc_scale * z)            [AffineProcess.propagate]
//     obs_lp(y, x, P)                                 log p(y | x)                                    [build_density().log_prob]
// P is the per-column parameter row (raw parameters first, then host-precomputed constants), see smcb_param_layout in
// smcb_api.cu and pyfilter_b200/timeseries.py.  Arithmetic is written with explicit round-to-nearest intrinsics so that every
// call site produces identical bits (the APF recomputes the look-ahead weight of the ancestor instead of gathering it).
#pragma once
#include <cuda_runtime.h>

#define SMCB_NPARAM 40
#define SMCB_LOG_SQRT_2PI 0.9189385332046727f

enum { SMCB_MODEL_LG_AR1 = 0, SMCB_MODEL_SINE_EM = 1, SMCB_MODEL_SV_AR1 = 2, SMCB_MODEL_LORENZ63_EM = 3, SMCB_MODEL_USER = 4, SMCB_NUM_MODELS = 5 };
enum { SMCB_PROPOSAL_BOOTSTRAP = 0, SMCB_PROPOSAL_LINEAR_GAUSS = 1, SMCB_PROPOSAL_LINEARIZED = 2, SMCB_PROPOSAL_NESTED = 3 };
enum { SMCB_ALG_SISR = 0, SMCB_ALG_APF = 1 };
enum { SMCB_RESAMPLE_SYSTEMATIC = 0, SMCB_RESAMPLE_MULTINOMIAL = 1 };

// Parameter row layout (per column).  Slots 0..7: model parameters and observation constants; 8..19: constants of the
// LinearGaussianObservations proposal; 20..26: increment scale and initial distribution.
// linear-Gaussian observation slots (lg_ar1, sine_em):   y = b + a x + s nu
#define P_OBS_A 3
#define P_OBS_B 4
#define P_OBS_S 5
#define P_OBS_INV2VAR 6    // 1 / (2 s^2)
#define P_OBS_LOGNORM 7    // log s + log sqrt(2 pi)
// LinearGaussianObservations constants (proposals/linear.py:38-86, proposals/utils.py:219-267), scalar state/observation
#define P_LGO_HVI 8          // sigma^-2
#define P_LGO_COV 9          // P = 1 / (sigma^-2 + a^2 s^-2)
#define P_LGO_KSTD 10        // sqrt(P)
#define P_LGO_K_INV2VAR 11   // 1 / (2 P)
#define P_LGO_K_LOGNORM 12   // log sqrt(P) + log sqrt(2 pi)
#define P_LGO_PRE_INV2VAR 13 // 1 / (2 (s^2 + a^2 sigma^2))
#define P_LGO_PRE_LOGNORM 14 // log sqrt(s^2 + a^2 sigma^2) + log sqrt(2 pi)
#define P_LGO_INC_INV2VAR 15 // 1 / (2 inc_scale^2)
#define P_LGO_INC_LOGNORM 16 // log inc_scale + log sqrt(2 pi) + log sigma   (increment density + log|d inc / d x|)
#define P_LGO_INV_SIGMA 17   // 1 / sigma
#define P_LGO_OVI 18         // s^-2
#define P_INC_SCALE 20       // std of the increment distribution (1 or sqrt(dt))
#define P_X0_LOC 21          // .. +2 (three dims)
#define P_X0_SCALE 24        // .. +2
// multi-dimensional LinearGaussianObservations of the Lorenz model (examples/lorenz.ipynb:214): with y = a (x1, x3) + s nu the matrices
// of proposals/utils.py:219-267 are DIAGONAL - P = diag(p, 1/sigma^-2, p), p = 1/(sigma^-2 + a^2 s^-2) - so slots 8..18 serve the two
// observed coordinates with (a, s) = (obs_a, obs_s) and these four the unobserved one
#define P_LGO_COV1 27        // fl32(1 / sigma^-2)
#define P_LGO_KSTD1 28       // sqrt of it
#define P_LGO_K1_INV2VAR 29
#define P_LGO_K1_LOGNORM 30
#define P_LGO_OBS_S 31        // Lorenz: obs_s itself (slots 6, 7 hold the derived constants of its density)
// Linearized proposal (proposals/linearized.py:22, proposals/utils.py:30-146): its three settings and the constants of the transition
// density log p(x' | x) = sum_d -((x'_d - loc_d) / scale)^2 / (2 inc^2) - log(inc |scale|) - log sqrt(2 pi)
#define P_LIN_STEPS 32        // n_steps (as float)
#define P_LIN_ALPHA 33        // alpha
#define P_LIN_SECOND 34       // use_second_order (0 / 1)
#define P_LIN_T_INV2VAR 35    // 1 / (2 (scale inc)^2)
#define P_LIN_T_LOGNORM 36    // log(inc |scale|) + log sqrt(2 pi)
#define P_LIN_T_INVVAR 37     // 1 / (scale inc)^2
#define P_NESTED_M 38         // NestedProposal: num_samples (as float), proposals/nested.py:17
#define SMCB_NESTED_MAX 256   // largest num_samples (the per-particle log-densities of the inner samples sit in local memory)

// sin(v) for the drift of the sine diffusion: explicit two-constant reduction to [-pi, pi] (exact products through fma), then the SFU.
// Absolute error <= 2^-20.9 ~ 5e-7 on the reduced argument (PTX sin.approx.ftz.f32) against ~27 instructions of sinf() with its slow
// path; the drift is multiplied by dt, so the state moves by < 1e-7 - an ulp of x - and the parity tolerances (2e-6 on x_t) hold.
// One function for every kernel: the APF recomputes g(x_anc) from the gathered state and needs identical bits at every call site.
__device__ __forceinline__ float smcb_sin(float v) {
  const float k = __fadd_rn(fmaf(v, 0.15915494309189535f, 12582912.f), -12582912.f);  // nearest integer to v / (2 pi)
  float r = fmaf(k, -6.2831854820251465f, v);                                           // v - k * fl32(2 pi)
  r = fmaf(k, 1.7484555314695172e-7f, r);                                               // ... - k * (2 pi - fl32(2 pi))
  return __sinf(r);
}

__device__ __forceinline__ float smcb_normal_lp(float v, float loc, float inv2var, float lognorm) {
  float d = __fsub_rn(v, loc);
  return __fsub_rn(__fmul_rn(-__fmul_rn(d, d), inv2var), lognorm);
}
// Normal(loc, scale).log_prob(v) with scale given directly (used where scale is not a per-column constant)
__device__ __forceinline__ float smcb_normal_lp_scale(float v, float loc, float scale) {
  float d = __fsub_rn(v, loc);
  float var2 = __fmul_rn(2.0f, __fmul_rn(scale, scale));
  return __fsub_rn(__fsub_rn(__fdiv_rn(-__fmul_rn(d, d), var2), logf(scale)), SMCB_LOG_SQRT_2PI);
}

template <int MODEL> struct Model;

// ---- config 1: x_t = alpha + beta x_{t-1} + sigma eps;  y = b + a x + s nu          (reference tests/filters/models.py:12-16)
template <> struct Model<SMCB_MODEL_LG_AR1> {
  static constexpr int D = 1, OD = 1;
  static constexpr bool LINEAR_OBS = true;
  __device__ static __forceinline__ void loc_scale(const float* x, const float* P, float* loc, float& scale) {
    loc[0] = __fadd_rn(P[0], __fmul_rn(P[1], x[0]));
    scale = P[2];
  }
  __device__ static __forceinline__ float obs_lp(const float* y, const float* x, const float* P) {
    return smcb_normal_lp(y[0], __fadd_rn(P[P_OBS_B], __fmul_rn(P[P_OBS_A], x[0])), P[P_OBS_INV2VAR], P[P_OBS_LOGNORM]);
  }
  // d/dx and d2/dx2 of obs_lp (Linearized proposal: proposals/utils.py:52-62 takes them by automatic differentiation)
  __device__ static __forceinline__ void obs_grad_hess(const float* y, const float* x, const float* P, float* g, float* h) {
    const float r = __fsub_rn(y[0], __fadd_rn(P[P_OBS_B], __fmul_rn(P[P_OBS_A], x[0])));
    const float ivar = __fmul_rn(2.0f, P[P_OBS_INV2VAR]);
    g[0] = __fmul_rn(__fmul_rn(P[P_OBS_A], r), ivar);
    h[0] = -__fmul_rn(__fmul_rn(P[P_OBS_A], P[P_OBS_A]), ivar);
  }
  // y = b + a x + s v   (sample of build_density(x): ParticleFilterCorrection.predict_path, particle/state.py:173-174)
  __device__ static __forceinline__ void obs_sample(const float* x, const float* v, const float* P, float* y) {
    y[0] = __fadd_rn(__fadd_rn(P[P_OBS_B], __fmul_rn(P[P_OBS_A], x[0])), __fmul_rn(P[P_OBS_S], v[0]));
  }
};

// ---- configs 2 and 5: Euler-Maruyama of dx = sin(x - gamma) dt + sigma dW;  y = b + a x + s nu      (reference README.md:44-67)
template <> struct Model<SMCB_MODEL_SINE_EM> {
  static constexpr int D = 1, OD = 1;
  static constexpr bool LINEAR_OBS = true;
  __device__ static __forceinline__ void loc_scale(const float* x, const float* P, float* loc, float& scale) {
    loc[0] = __fadd_rn(x[0], __fmul_rn(smcb_sin(__fsub_rn(x[0], P[0])), P[2]));
    scale = P[1];
  }
  __device__ static __forceinline__ float obs_lp(const float* y, const float* x, const float* P) {
    return smcb_normal_lp(y[0], __fadd_rn(P[P_OBS_B], __fmul_rn(P[P_OBS_A], x[0])), P[P_OBS_INV2VAR], P[P_OBS_LOGNORM]);
  }
  // d/dx and d2/dx2 of obs_lp (Linearized proposal: proposals/utils.py:52-62 takes them by automatic differentiation)
  __device__ static __forceinline__ void obs_grad_hess(const float* y, const float* x, const float* P, float* g, float* h) {
    const float r = __fsub_rn(y[0], __fadd_rn(P[P_OBS_B], __fmul_rn(P[P_OBS_A], x[0])));
    const float ivar = __fmul_rn(2.0f, P[P_OBS_INV2VAR]);
    g[0] = __fmul_rn(__fmul_rn(P[P_OBS_A], r), ivar);
    h[0] = -__fmul_rn(__fmul_rn(P[P_OBS_A], P[P_OBS_A]), ivar);
  }
  __device__ static __forceinline__ void obs_sample(const float* x, const float* v, const float* P, float* y) {
    y[0] = __fadd_rn(__fadd_rn(P[P_OBS_B], __fmul_rn(P[P_OBS_A], x[0])), __fmul_rn(P[P_OBS_S], v[0]));
  }
};

// ---- config 3: x_t = mu + phi (x_{t-1} - mu) + sigma_v eps;  y ~ N(0, exp(x/2))
template <> struct Model<SMCB_MODEL_SV_AR1> {
  static constexpr int D = 1, OD = 1;
  static constexpr bool LINEAR_OBS = false;
  __device__ static __forceinline__ void loc_scale(const float* x, const float* P, float* loc, float& scale) {
    loc[0] = __fadd_rn(P[0], __fmul_rn(P[1], __fsub_rn(x[0], P[0])));
    scale = P[2];
  }
  __device__ static __forceinline__ float obs_lp(const float* y, const float* x, const float* P) {
    // Normal(0, exp(x/2)).log_prob(y) = -y^2 / (2 exp(x)) - x/2 - log sqrt(2 pi); exp(-x) through the SFU (ex2.approx.ftz, 2 ulp)
    const float hy2 = __fmul_rn(0.5f, __fmul_rn(y[0], y[0]));
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__fmul_rn(-x[0], 1.4426950408889634f)));
    return fmaf(-hy2, e, fmaf(-0.5f, x[0], -SMCB_LOG_SQRT_2PI));  // explicit fma: the same bits at every call site
  }
  __device__ static __forceinline__ void obs_grad_hess(const float* y, const float* x, const float* P, float* g, float* h) {
    const float hy2e = __fmul_rn(__fmul_rn(0.5f, __fmul_rn(y[0], y[0])), __expf(-x[0]));   // y^2 exp(-x) / 2
    g[0] = __fsub_rn(hy2e, 0.5f);
    h[0] = -hy2e;
  }
  __device__ static __forceinline__ void obs_sample(const float* x, const float* v, const float* P, float* y) {
    y[0] = __fmul_rn(__expf(__fmul_rn(0.5f, x[0])), v[0]);
  }
};

// ---- config 4: Euler-Maruyama Lorenz-63, y = obs_a (x^1, x^3) + obs_s nu                      (reference examples/lorenz.ipynb:53-117)
//      P: 0 s, 1 r, 2 b, 3 sigma, 4 dt, 5 obs_a, 6 inv2var, 7 lognorm (per component)
template <> struct Model<SMCB_MODEL_LORENZ63_EM> {
  static constexpr int D = 3, OD = 2;
  static constexpr bool LINEAR_OBS = true;   // y = a (x1, x3) + s nu: a LinearStateSpaceModel in the notebook (lorenz.ipynb:105-117)
  __device__ static __forceinline__ void loc_scale(const float* x, const float* P, float* loc, float& scale) {
    float f0 = __fmul_rn(-P[0], __fsub_rn(x[0], x[1]));
    float f1 = __fsub_rn(__fsub_rn(__fmul_rn(P[1], x[0]), x[1]), __fmul_rn(x[0], x[2]));
    float f2 = __fsub_rn(__fmul_rn(x[0], x[1]), __fmul_rn(P[2], x[2]));
    loc[0] = __fadd_rn(x[0], __fmul_rn(f0, P[4]));
    loc[1] = __fadd_rn(x[1], __fmul_rn(f1, P[4]));
    loc[2] = __fadd_rn(x[2], __fmul_rn(f2, P[4]));
    scale = P[3];
  }
  __device__ static __forceinline__ float obs_lp(const float* y, const float* x, const float* P) {
    float l0 = smcb_normal_lp(y[0], __fmul_rn(P[5], x[0]), P[6], P[7]);
    float l1 = smcb_normal_lp(y[1], __fmul_rn(P[5], x[2]), P[6], P[7]);
    return __fadd_rn(l0, l1);
  }
  __device__ static __forceinline__ void obs_grad_hess(const float* y, const float* x, const float* P, float* g, float* h) {
    const float ivar = __fmul_rn(2.0f, P[6]);   // 1 / obs_s^2; the Hessian of this model is diagonal (each observation reads one coordinate)
    g[0] = __fmul_rn(__fmul_rn(P[5], __fsub_rn(y[0], __fmul_rn(P[5], x[0]))), ivar);
    g[1] = 0.f;
    g[2] = __fmul_rn(__fmul_rn(P[5], __fsub_rn(y[1], __fmul_rn(P[5], x[2]))), ivar);
    h[0] = -__fmul_rn(__fmul_rn(P[5], P[5]), ivar); h[1] = 0.f; h[2] = h[0];
  }
  __device__ static __forceinline__ void obs_sample(const float* x, const float* v, const float* P, float* y) {
    y[0] = __fadd_rn(__fmul_rn(P[5], x[0]), __fmul_rn(P[P_LGO_OBS_S], v[0]));
    y[1] = __fadd_rn(__fmul_rn(P[5], x[2]), __fmul_rn(P[P_LGO_OBS_S], v[1]));
  }
};

// ---- a user-supplied model: the reference takes ANY stochproc callables; here the user writes the same two functions
//      as device code and the library is compiled once more with that header (pyfilter_b200.timeseries.compile_user_model: nvcc at run
//      time, the shared object cached by the hash of the source).  The header defines
//          struct UserModel {
//            static constexpr int D = ..., OD = ..., NRAW = ...;            // state / observation dimensions, raw parameters per column
//            static constexpr bool LINEAR_OBS = false;
//            __device__ static void loc_scale(const float* x, const float* P, float* loc, float& scale);     // mean_scale
//            __device__ static float obs_lp(const float* y, const float* x, const float* P);                 // build_density().log_prob
//            __device__ static void obs_sample(const float* x, const float* v, const float* P, float* y);    // build_density().sample
//            static void derive(const double* raw, float* P);   // HOST: raw parameters -> the row P (P[0 .. NRAW) + any constants), incl.
//          };                                                   //       P[P_INC_SCALE], P[P_X0_LOC + d], P[P_X0_SCALE + d]
#ifdef SMCB_USER_MODEL_HEADER
#include SMCB_USER_MODEL_HEADER
template <> struct Model<SMCB_MODEL_USER> : UserModel {};
#endif
