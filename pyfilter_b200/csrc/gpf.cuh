// Gaussian particle filter (reference filters/particle/gpf.py:10-36 with its default GaussianProposal, proposals/approximate.py:13-37;
// Kotecha & Djuric).  One move:
//   gpf_predict_kernel   x~ = propagate(x_{t-1})  (stochproc AffineProcess.propagate), stored as the next state, and the block sums of
//                        W, W (x~ - shift), W (x~ - shift)(x~ - shift)^T with W the normalised weights of the previous state
//                        (ParticleFilterPrediction.get_predictive_density(approximate=True), particle/state.py:59-66)
//   gpf_moments_kernel   mean = sum W x~, cov = sum W (x~ - mean)(x~ - mean)^T  (particle/utils.py:42-58, covariance=True), Cholesky factor
//   gpf_correct_kernel   x_t ~ N(mean, cov) for every particle, log w_t = log p(y_t | x_t) (the weights are REPLACED, approximate.py:27-35);
//                        an all-NaN y_t keeps x~ and the old weights (filters/base.py:213-214, particle/state.py:38-42);
//                        soft-max partial records in the layout of step_kernel, so that finalize_kernel<.., SISR> folds them:
//                        ll_t = log mean exp(log w_t) (gpf.py:35, particle/utils.py:7-22 with uniform weights), moments, ESS, history rows.
// Nothing resamples: previous_indices stay what they were (gpf.py:29-30).
#pragma once
#include "step.cuh"

#define SMCB_RNG_GPF 0x60u   // + state dimension: the draws from the Gaussian approximation

template <int D>
struct GpfLayout {
  static constexpr int NC = D * (D + 1) / 2;   // upper triangle of the second moments
  static constexpr int NP = 1 + D + NC;        // one block's sums
  static constexpr int ND = D + D * D;         // per column: mean, Cholesky factor (row-major, lower)
};

struct GpfArgs {
  StepArgs s;
  float* partial;   // (B, blocks_per_col, NP)
  float* dist;      // (B, ND)
};

template <int NV>
__device__ __forceinline__ void gpf_block_sum(float (&v)[NV], float* scratch /* (ST_NT / 32) * NV */) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) scratch[warp * NV + k] = v[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < ST_NT / 32; ++w) {
#pragma unroll
      for (int k = 0; k < NV; ++k) v[k] += scratch[w * NV + k];
    }
  }
}

template <int MODEL>
__global__ void __launch_bounds__(ST_NT) gpf_predict_kernel(GpfArgs g) {
  typedef Model<MODEL> M;
  constexpr int D = M::D, NC = GpfLayout<D>::NC, NP = GpfLayout<D>::NP;
  const StepArgs& a = g.s;
  __shared__ float Ps[SMCB_NPARAM];
  __shared__ float scratch[(ST_NT / 32) * NP];
  const int col = blockIdx.y, tid = threadIdx.x;
  if (tid < SMCB_NPARAM) Ps[tid] = a.P[(int64_t)col * SMCB_NPARAM + tid];
  __syncthreads();
  const int t = a.t_host;
  const float* xcur = a.xbuf[t & 1];
  float* xnext = a.xbuf[(t + 1) & 1];
  const ColStats st = a.stats[col];
  float acc[NP];
#pragma unroll
  for (int k = 0; k < NP; ++k) acc[k] = 0.f;
  for (int it = 0; it < a.iters; ++it) {
    const int64_t i0 = ((int64_t)(it * a.blocks_per_col + blockIdx.x) * ST_NT + tid) * ST_VEC;
    if (i0 >= a.n) continue;
    float x[D][4], z[D][4];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const float4 q = *reinterpret_cast<const float4*>(xcur + ((int64_t)d * a.B + col) * a.ld + i0);
      x[d][0] = q.x; x[d][1] = q.y; x[d][2] = q.z; x[d][3] = q.w;
    }
    const float4 l = *reinterpret_cast<const float4*>(a.lw + (int64_t)col * a.ld + i0);
    const float lw[4] = {l.x, l.y, l.z, l.w};
    st_noise4<D>(a, col, i0, t, SMCB_RNG_TRANSITION, z);
    float xp[D][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float xs[D], loc[D], sc;
#pragma unroll
      for (int d = 0; d < D; ++d) xs[d] = x[d][k];
      M::loc_scale(xs, Ps, loc, sc);
#pragma unroll
      for (int d = 0; d < D; ++d) xp[d][k] = __fadd_rn(loc[d], __fmul_rn(sc, __fmul_rn(z[d][k], Ps[P_INC_SCALE])));
      if (i0 + k < a.n) {
        const float W = smcb_weight(lw[k], st.m_lw, st.inv_z_lw);
        float dx[D];
#pragma unroll
        for (int d = 0; d < D; ++d) dx[d] = xp[d][k] - st.shift[d];
        acc[0] += W;
        int c = 0;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          acc[1 + d] += W * dx[d];
#pragma unroll
          for (int e = d; e < D; ++e) acc[1 + D + c++] += W * dx[d] * dx[e];
        }
      }
    }
#pragma unroll
    for (int d = 0; d < D; ++d)
      *reinterpret_cast<float4*>(xnext + ((int64_t)d * a.B + col) * a.ld + i0) = make_float4(xp[d][0], xp[d][1], xp[d][2], xp[d][3]);
  }
  (void)NC;
  gpf_block_sum<NP>(acc, scratch);
  if (tid == 0) {
    float* p = g.partial + ((int64_t)col * a.blocks_per_col + blockIdx.x) * NP;
#pragma unroll
    for (int k = 0; k < NP; ++k) p[k] = acc[k];
  }
}

template <int D>
__global__ void __launch_bounds__(ST_NT) gpf_moments_kernel(GpfArgs g) {
  constexpr int NP = GpfLayout<D>::NP, ND = GpfLayout<D>::ND;
  const StepArgs& a = g.s;
  __shared__ float scratch[(ST_NT / 32) * NP];
  const int col = blockIdx.x, tid = threadIdx.x;
  float acc[NP];
#pragma unroll
  for (int k = 0; k < NP; ++k) acc[k] = 0.f;
  for (int b = tid; b < a.blocks_per_col; b += ST_NT) {
    const float* p = g.partial + ((int64_t)col * a.blocks_per_col + b) * NP;
#pragma unroll
    for (int k = 0; k < NP; ++k) acc[k] += p[k];
  }
  gpf_block_sum<NP>(acc, scratch);
  if (tid != 0) return;
  const ColStats st = a.stats[col];
  float dm[D], cov[D][D];
  float* out = g.dist + (int64_t)col * ND;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    dm[d] = acc[1 + d] + st.shift[d] * (acc[0] - 1.f);   // mean - shift, with mean = sum W x~ as the reference forms it
    out[d] = st.shift[d] + dm[d];
  }
  int c = 0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
#pragma unroll
    for (int e = d; e < D; ++e) {
      const float v = acc[1 + D + c++] - dm[d] * acc[1 + e] - dm[e] * acc[1 + d] + acc[0] * dm[d] * dm[e];
      cov[d][e] = v; cov[e][d] = v;
    }
  }
  // Cholesky factor (MultivariateNormal(mean, covariance_matrix=cov), Normal(mean, sqrt(var)) for a scalar state)
  float L[D][D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
#pragma unroll
    for (int j = 0; j < D; ++j) L[i][j] = 0.f;
  }
#pragma unroll
  for (int j = 0; j < D; ++j) {
    float s = cov[j][j];
#pragma unroll
    for (int k = 0; k < j; ++k) s -= L[j][k] * L[j][k];
    L[j][j] = sqrtf(s);
#pragma unroll
    for (int i = j + 1; i < D; ++i) {
      float v = cov[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k];
      L[i][j] = v / L[j][j];
    }
  }
#pragma unroll
  for (int i = 0; i < D; ++i) {
#pragma unroll
    for (int j = 0; j < D; ++j) out[D + i * D + j] = L[i][j];
  }
}

template <int MODEL>
__global__ void __launch_bounds__(ST_NT) gpf_correct_kernel(GpfArgs g) {
  typedef Model<MODEL> M;
  constexpr int D = M::D, OD = M::OD, ND = GpfLayout<D>::ND;
  const StepArgs& a = g.s;
  __shared__ float Ps[SMCB_NPARAM];
  __shared__ float dist[ND];
  __shared__ SoftAcc<1 + 2 * D> sA[ST_NT / 32];
  __shared__ SoftAcc<1> sQ[ST_NT / 32];
  __shared__ SoftAcc<1> sR[ST_NT / 32];
  const int col = blockIdx.y, tid = threadIdx.x;
  if (tid < SMCB_NPARAM) Ps[tid] = a.P[(int64_t)col * SMCB_NPARAM + tid];
  if (tid < ND) dist[tid] = g.dist[(int64_t)col * ND + tid];
  __syncthreads();
  const int t = a.t_host;
  float y[OD];
  const bool observed = st_load_obs<OD>(a.y_t, y);
  float* xnext = a.xbuf[(t + 1) & 1];
  float shift[D];
#pragma unroll
  for (int d = 0; d < D; ++d) shift[d] = a.stats[col].shift[d];
  const float inv_n = 1.0f / (float)a.n;
  Moments<D> mom; mom.init();
  SoftAcc<1> r3; r3.init();
  for (int it = 0; it < a.iters; ++it) {
    const int64_t i0 = ((int64_t)(it * a.blocks_per_col + blockIdx.x) * ST_NT + tid) * ST_VEC;
    if (i0 >= a.n) continue;
    bool valid[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) valid[k] = i0 + k < a.n;
    float xn[D][4], lw[4];
    if (observed) {
      float z[D][4];
#pragma unroll
      for (int d = 0; d < D; ++d) {
        if (a.nest_z) {   // parity hook: the N(0, 1) draws behind predictive_distribution.sample(), layout (1, D, B, ld)
          const float4 q = *reinterpret_cast<const float4*>(a.nest_z + ((int64_t)d * a.B + col) * a.ld + i0);
          z[d][0] = q.x; z[d][1] = q.y; z[d][2] = q.z; z[d][3] = q.w;
        } else {
          const Philox4 r = philox4x32_10_keys((uint32_t)(i0 >> 2), (uint32_t)(col + a.col0), (uint32_t)t, SMCB_RNG_GPF + d, a.pkeys);
          smcb_normal4(r, z[d]);
        }
      }
      float mx = -INFINITY;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float xs[D];
#pragma unroll
        for (int d = 0; d < D; ++d) {
          float v = dist[d];
#pragma unroll
          for (int e = 0; e <= d; ++e) v += dist[D + d * D + e] * z[e][k];
          xs[d] = v;
          xn[d][k] = v;
        }
        lw[k] = st_sanitize(M::obs_lp(y, xs, Ps));
        if (valid[k]) mx = fmaxf(mx, lw[k]);
      }
      if (mx > -INFINITY) {
        r3.raise(mx);
#pragma unroll
        for (int k = 0; k < 4; ++k) if (valid[k]) r3.s[0] += inv_n * __expf(lw[k] - r3.m);
      }
#pragma unroll
      for (int d = 0; d < D; ++d)
        *reinterpret_cast<float4*>(xnext + ((int64_t)d * a.B + col) * a.ld + i0) = make_float4(xn[d][0], xn[d][1], xn[d][2], xn[d][3]);
    } else {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const float4 q = *reinterpret_cast<const float4*>(xnext + ((int64_t)d * a.B + col) * a.ld + i0);
        xn[d][0] = q.x; xn[d][1] = q.y; xn[d][2] = q.z; xn[d][3] = q.w;
      }
      const float4 l = *reinterpret_cast<const float4*>(a.lw + (int64_t)col * a.ld + i0);
      lw[0] = l.x; lw[1] = l.y; lw[2] = l.z; lw[3] = l.w;
    }
    *reinterpret_cast<float4*>(a.lw_out + (int64_t)col * a.ld + i0) = make_float4(lw[0], lw[1], lw[2], lw[3]);
    mom.add4(lw, xn, shift, valid);
  }
  softacc_block_reduce(mom.a, sA);
  softacc_block_reduce(mom.q, sQ);
  softacc_block_reduce(r3, sR);
  if (tid == 0) {
    Partial& p = a.partials[(int64_t)col * a.blocks_per_col + blockIdx.x];
    st_write_partial1(p, mom.a, mom.q);
    p.m2 = -INFINITY; p.z2 = 0.f;
    p.m3 = r3.m; p.z3 = r3.s[0];
  }
}
