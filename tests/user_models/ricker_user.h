// Test fixture: a model that is NOT in the compiled zoo - a noisy Ricker map on the log scale with a Student-t-free, heteroscedastic
// Gaussian observation:   x' = x + r (1 - exp(x) / K) + sigma e,    y ~ N(exp(x), (tau exp(x / 2))^2),   x_0 ~ N(log K, 0.5)
struct UserModel {
  static constexpr int D = 1, OD = 1, NRAW = 4;   // r, K, sigma, tau
  static constexpr bool LINEAR_OBS = false;
  __device__ static __forceinline__ void loc_scale(const float* x, const float* P, float* loc, float& scale) {
    loc[0] = x[0] + P[0] * (1.0f - expf(x[0]) / P[1]);
    scale = P[2];
  }
  __device__ static __forceinline__ float obs_lp(const float* y, const float* x, const float* P) {
    const float m = expf(x[0]), s = P[3] * expf(0.5f * x[0]);
    const float d = (y[0] - m) / s;
    return -0.5f * d * d - logf(s) - SMCB_LOG_SQRT_2PI;
  }
  __device__ static __forceinline__ void obs_sample(const float* x, const float* v, const float* P, float* y) {
    y[0] = expf(x[0]) + P[3] * expf(0.5f * x[0]) * v[0];
  }
  static void derive(const double* r, float* P) {
    for (int i = 0; i < 4; ++i) P[i] = (float)r[i];
    P[P_INC_SCALE] = 1.f;
    P[P_X0_LOC] = (float)log(r[1]);
    P[P_X0_SCALE] = 0.5f;
  }
};
