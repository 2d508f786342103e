// Test fixture: the stochastic-volatility model of BASELINE.json configs[2] written as a USER model (csrc/models.h contract) - the same
// arithmetic as the built-in Model<SMCB_MODEL_SV_AR1>, so a filter on it must reproduce the built-in bit for bit.
struct UserModel {
  static constexpr int D = 1, OD = 1, NRAW = 3;   // mu, phi, sigma_v
  static constexpr bool LINEAR_OBS = false;
  __device__ static __forceinline__ void loc_scale(const float* x, const float* P, float* loc, float& scale) {
    loc[0] = __fadd_rn(P[0], __fmul_rn(P[1], __fsub_rn(x[0], P[0])));
    scale = P[2];
  }
  __device__ static __forceinline__ float obs_lp(const float* y, const float* x, const float* P) {
    const float hy2 = __fmul_rn(0.5f, __fmul_rn(y[0], y[0]));
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__fmul_rn(-x[0], 1.4426950408889634f)));
    return fmaf(-hy2, e, fmaf(-0.5f, x[0], -SMCB_LOG_SQRT_2PI));
  }
  __device__ static __forceinline__ void obs_sample(const float* x, const float* v, const float* P, float* y) {
    y[0] = __fmul_rn(__expf(__fmul_rn(0.5f, x[0])), v[0]);
  }
  static void derive(const double* r, float* P) {
    for (int i = 0; i < 3; ++i) P[i] = (float)r[i];
    P[P_INC_SCALE] = 1.f;
    P[P_X0_LOC] = (float)r[0];
    P[P_X0_SCALE] = (float)(r[2] / sqrt(1.0 - r[1] * r[1]));
  }
};
