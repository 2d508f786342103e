// The model of the reference's examples/stochastic-volatility.ipynb:60-83 as a user model: a Verhulst volatility process, Euler-Maruyama
// with step dt, observed every 1/dt steps through a sinh-arcsinh transformed standard normal,
//     V' = V + kappa V (gamma - V) dt + sigma V sqrt(dt) e,        Y = mu + V sinh((asinh(W) + nu) tau),   W ~ N(0, 1),   V_0 ~ N(gamma, sigma)
// (the process and the transform are classes of the absent stochproc package; the equations are those of the notebook's first cell).
struct UserModel {
  static constexpr int D = 1, OD = 1, NRAW = 7;   // kappa, gamma, sigma, mu, nu (skew), tau (tail weight), dt
  static constexpr bool LINEAR_OBS = false;
  __device__ static __forceinline__ void loc_scale(const float* x, const float* P, float* loc, float& scale) {
    loc[0] = x[0] + P[0] * x[0] * (P[1] - x[0]) * P[6];
    scale = P[2] * x[0];
  }
  // log density of y = mu + v s(W):  w = sinh(asinh(r) / tau - nu) with r = (y - mu) / v,  p(y) = phi(w) cosh(asinh(r) / tau - nu) / (tau sqrt(1 + r^2) |v|)
  __device__ static __forceinline__ float obs_lp(const float* y, const float* x, const float* P) {
    const float r = (y[0] - P[3]) / x[0];
    const float a = asinhf(r) / P[5] - P[4];
    const float w = sinhf(a);
    return -0.5f * w * w - SMCB_LOG_SQRT_2PI + logf(coshf(a)) - logf(P[5]) - 0.5f * log1pf(r * r) - logf(fabsf(x[0]));
  }
  __device__ static __forceinline__ void obs_sample(const float* x, const float* v, const float* P, float* y) {
    y[0] = P[3] + x[0] * sinhf((asinhf(v[0]) + P[4]) * P[5]);
  }
  static void derive(const double* r, float* P) {
    for (int i = 0; i < 7; ++i) P[i] = (float)r[i];
    P[P_INC_SCALE] = (float)sqrt(r[6]);
    P[P_X0_LOC] = (float)r[1];
    P[P_X0_SCALE] = (float)r[2];
  }
};
