// The callers and plug-ins either side of the fused move: kernels behind the stand-alone entry points of the
// proposal plug-in surface (proposals/base.py:52-85), ParticleFilterCorrection.predict_path (particle/state.py:173-174), batched_gather
// (filters/utils.py:4-21), fixed-lag smoothing (filters/particle/base.py:130-146), the theta-level column operations of SMC2 / PMMH
// (FilterResult.resample / exchange, filters/result.py:76-117; particle/state.py:150-168) and the residual resampler's deterministic
// part (resampling.py:68-105).  They use the SAME Proposal<> / Model<> device functions and Philox counters as the fused kernels.
#pragma once
#include "step.cuh"

// ---- proposal plug-in: pre_weight / sample_and_weight over a caller's particles ------------------------------------------------------
struct ProposalOpArgs {
  StepArgs s;            // n, ld, B, P, eps_in (optional injected N(0,1) draws), pkeys, col0
  const float* x_in;     // (D, B, ld)
  const float* y;        // (OD) observation on the device
  float* x_out;          // (D, B, ld)  sample_and_weight: the proposed particles
  float* w_out;          // (B, ld)     log-weights: pre_weight -> log p(y | .), sample_and_weight -> the weight increment
  int32_t mode;          // 0 pre_weight, 1 sample_and_weight
  int32_t t;             // move index of the Philox counter
};

template <int MODEL, int PROP>
__global__ void __launch_bounds__(ST_NT) proposal_op_kernel(ProposalOpArgs c) {
  typedef Model<MODEL> M;
  constexpr int D = M::D, OD = M::OD;
  __shared__ float Ps[SMCB_NPARAM];
  const StepArgs& a = c.s;
  const int col = blockIdx.y, tid = threadIdx.x;
  if (tid < SMCB_NPARAM) Ps[tid] = a.P[(int64_t)col * SMCB_NPARAM + tid];
  __syncthreads();
  float y[OD];
  const bool observed = st_load_obs<OD>(c.y, y);
  const int64_t i0 = ((int64_t)blockIdx.x * ST_NT + tid) * ST_VEC;
  if (i0 >= a.n) return;
  float x[D][4];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const float4 q = *reinterpret_cast<const float4*>(c.x_in + ((int64_t)d * a.B + col) * a.ld + i0);
    x[d][0] = q.x; x[d][1] = q.y; x[d][2] = q.z; x[d][3] = q.w;
  }
  float w[4];
  if (c.mode == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float xk[D];
#pragma unroll
      for (int d = 0; d < D; ++d) xk[d] = x[d][k];
      w[k] = observed ? Proposal<MODEL, PROP>::pre_weight(y, xk, Ps) : 0.f;
    }
  } else {
    float z[D][4], xn[D][4];
    st_noise4<D>(a, col, i0, c.t, SMCB_RNG_TRANSITION, z);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float xk[D], zk[D], xo[D], inc, g_anc;
#pragma unroll
      for (int d = 0; d < D; ++d) { xk[d] = x[d][k]; zk[d] = z[d][k]; }
      prop_sample_and_weight<MODEL, PROP>(a, col, i0 + k, c.t, y, xk, zk, Ps, observed, xo, inc, g_anc);
#pragma unroll
      for (int d = 0; d < D; ++d) xn[d][k] = xo[d];
      w[k] = inc;
    }
#pragma unroll
    for (int d = 0; d < D; ++d)
      *reinterpret_cast<float4*>(c.x_out + ((int64_t)d * a.B + col) * a.ld + i0) = make_float4(xn[d][0], xn[d][1], xn[d][2], xn[d][3]);
  }
  *reinterpret_cast<float4*>(c.w_out + (int64_t)col * a.ld + i0) = make_float4(w[0], w[1], w[2], w[3]);
}

// ---- predict_path (particle/state.py:173-174 -> model.sample_states(num_steps, x_0)): every particle simulated forward -------------
// x_out (steps, D, B, ld), y_out (steps, OD, B, ld); row s is the state after s + 1 transitions and its observation.  Philox counters:
// (particle group, column, t0 + s, SMCB_RNG_PATH + dimension) - a stream of its own, so a path never replays the filter's noise.
#define SMCB_RNG_PATH 24u
struct PathArgs {
  StepArgs s;
  const float* x_in;   // (D, B, ld)
  float* x_out;
  float* y_out;
  int32_t steps, t0;
};
template <int MODEL>
__global__ void __launch_bounds__(ST_NT) predict_path_kernel(PathArgs c) {
  typedef Model<MODEL> M;
  constexpr int D = M::D, OD = M::OD;
  __shared__ float Ps[SMCB_NPARAM];
  const StepArgs& a = c.s;
  const int col = blockIdx.y, tid = threadIdx.x;
  if (tid < SMCB_NPARAM) Ps[tid] = a.P[(int64_t)col * SMCB_NPARAM + tid];
  __syncthreads();
  const int64_t i0 = ((int64_t)blockIdx.x * ST_NT + tid) * ST_VEC;
  if (i0 >= a.n) return;
  float x[D][4];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const float4 q = *reinterpret_cast<const float4*>(c.x_in + ((int64_t)d * a.B + col) * a.ld + i0);
    x[d][0] = q.x; x[d][1] = q.y; x[d][2] = q.z; x[d][3] = q.w;
  }
  const int64_t plane = (int64_t)a.B * a.ld;
  for (int s = 0; s < c.steps; ++s) {
    float z[D][4], v[OD][4];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const Philox4 r = philox4x32_10_keys((uint32_t)(i0 >> 2), (uint32_t)(col + a.col0), (uint32_t)(c.t0 + s), SMCB_RNG_PATH + d, a.pkeys);
      smcb_normal4(r, z[d]);
    }
#pragma unroll
    for (int d = 0; d < OD; ++d) {
      const Philox4 r = philox4x32_10_keys((uint32_t)(i0 >> 2), (uint32_t)(col + a.col0), (uint32_t)(c.t0 + s), SMCB_RNG_PATH + 4u + d, a.pkeys);
      smcb_normal4(r, v[d]);
    }
    float yv[OD][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float xk[D], loc[D], sc, vk[OD], yo[OD];
#pragma unroll
      for (int d = 0; d < D; ++d) xk[d] = x[d][k];
      M::loc_scale(xk, Ps, loc, sc);
#pragma unroll
      for (int d = 0; d < D; ++d) { x[d][k] = __fadd_rn(loc[d], __fmul_rn(sc, __fmul_rn(z[d][k], Ps[P_INC_SCALE]))); xk[d] = x[d][k]; }
#pragma unroll
      for (int d = 0; d < OD; ++d) vk[d] = v[d][k];
      M::obs_sample(xk, vk, Ps, yo);
#pragma unroll
      for (int d = 0; d < OD; ++d) yv[d][k] = yo[d];
    }
#pragma unroll
    for (int d = 0; d < D; ++d)
      *reinterpret_cast<float4*>(c.x_out + ((int64_t)s * D + d) * plane + (int64_t)col * a.ld + i0) = make_float4(x[d][0], x[d][1], x[d][2], x[d][3]);
#pragma unroll
    for (int d = 0; d < OD; ++d)
      *reinterpret_cast<float4*>(c.y_out + ((int64_t)s * OD + d) * plane + (int64_t)col * a.ld + i0) = make_float4(yv[d][0], yv[d][1], yv[d][2], yv[d][3]);
  }
}

// ---- batched_gather (filters/utils.py:4-21) and one backward step of fixed-lag smoothing (filters/particle/base.py:130-146) ------------
// Reference layout, contiguous: x (N, B, D) float32, idx (N, B) int64 -> out[i, b, :] = x[idx[i, b], b, :].
// With `prev` (N, B) int64 the index is first pushed one generation back: lineage[i, b] <- prev[lineage[i, b], b] (written back), the step
// `prev_inds = batched_gather(latest_state.previous_indices, prev_inds)` followed by the gather of the older state's particles.
__global__ void __launch_bounds__(256) gather_lineage_kernel(const float* __restrict__ x, int64_t n, int B, int D, int64_t* __restrict__ lineage,
                                                              const int64_t* __restrict__ prev, float* __restrict__ out, int* bad) {
  const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;   // element of the (N, B) index matrix
  if (e >= n * B) return;
  const int b = (int)(e % B);
  int64_t l = lineage[e];
  if (l < 0 || l >= n) { if (bad) atomicExch(bad, 1); return; }
  if (prev) {
    l = prev[l * B + b];
    if (l < 0 || l >= n) { if (bad) atomicExch(bad, 1); return; }
    lineage[e] = l;
  }
  const float* src = x + (l * B + b) * D;
  float* dst = out + e * D;
  for (int d = 0; d < D; ++d) dst[d] = __ldg(src + d);
}

// ---- theta-level column operations on the resident state of a handle (SMC2 / PMMH) --------------------------------------------------
// gather: dst column b <- src column idx[b] (FilterResult.resample, filters/result.py:76-95; particle/state.py:150-158)
// masked copy: dst column b <- src column b where mask[b] (FilterResult.exchange, filters/result.py:97-117; particle/state.py:160-168)
// rows: `planes` planes of (B, ld) elements of 4 bytes each (state dimensions, weights, ancestors), 128-bit copies.
__global__ void __launch_bounds__(256) column_gather_kernel(const float4* __restrict__ src, float4* __restrict__ dst, int B, int64_t ld4, const int64_t* __restrict__ idx,
                                                             const uint8_t* __restrict__ mask, int* bad) {
  const int b = blockIdx.y, plane = blockIdx.z;
  int64_t from = b;
  if (idx) {
    from = idx[b];
    if (from < 0 || from >= B) { if (bad && threadIdx.x == 0 && blockIdx.x == 0) atomicExch(bad, 1); return; }
  }
  if (mask && !mask[b]) return;
  const float4* s = src + ((int64_t)plane * B + from) * ld4;
  float4* d = dst + ((int64_t)plane * B + b) * ld4;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < ld4; i += (int64_t)gridDim.x * 256) d[i] = __ldg(s + i);
}
// the small per-column records (ColStats, running log-likelihood, latest moments, history rows): `rec` floats per column and row
__global__ void column_gather_small_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int rec, int rows, const int64_t* __restrict__ idx,
                                           const uint8_t* __restrict__ mask) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)rows * B * rec) return;
  const int k = (int)(e % rec);
  const int b = (int)((e / rec) % B);
  const int64_t r = e / ((int64_t)rec * B);
  int64_t from = idx ? idx[b] : b;
  if (from < 0 || from >= B) return;
  if (mask && !mask[b]) return;
  dst[(r * B + b) * rec + k] = src[(r * B + from) * rec + k];
}

// ---- residual resampling (resampling.py:68-105): the deterministic copies ------------------------------------------------------------
// mw = fl32(n w), floored = floor(mw); particle j is copied floored_j times to out[cum_{j-1} .. cum_j); the fractional parts
// (mw - floored) / k (k = sum floored, float32 division) are the weights of the multinomial part, drawn by the multinomial pipeline.
// Three passes over tiles of 4096 particles: counts + tile sums, exclusive scan of the tile sums (one block per column), expansion.
// n <= 2^24 (torch.multinomial's limit), so every count and sum is an exact float32 / int32.
#define RES_NT 256
#define RES_ITEMS 16
#define RES_TILE (RES_NT * RES_ITEMS)
__global__ void __launch_bounds__(RES_NT) residual_counts_kernel(const float* __restrict__ w, int64_t n, int64_t ld, int tiles, int32_t* __restrict__ counts,
                                                                 float* __restrict__ frac, int32_t* __restrict__ tile_sum) {
  __shared__ int32_t scratch[33];
  const int tile = blockIdx.x, col = blockIdx.y;
  const float nf = (float)n;
  int32_t local = 0;
#pragma unroll
  for (int j = 0; j < RES_ITEMS; ++j) {
    const int64_t i = (int64_t)tile * RES_TILE + j * RES_NT + threadIdx.x;
    if (i < n) {
      const float mw = __fmul_rn(nf, w[(int64_t)col * ld + i]);
      const float fl = floorf(mw);
      counts[(int64_t)col * ld + i] = (int32_t)fl;
      frac[(int64_t)col * ld + i] = __fsub_rn(mw, fl);
      local += (int32_t)fl;
    }
  }
  const int32_t k = block_allreduce<RES_NT>(local, 0, OpSumI(), scratch);
  if (threadIdx.x == 0) tile_sum[(int64_t)col * tiles + tile] = k;
}
// exclusive scan of a column's tile sums (in place) and its total k; one block per column
__global__ void __launch_bounds__(1024) residual_scan_kernel(int32_t* __restrict__ tile_sum, int tiles, int32_t* __restrict__ ksum) {
  __shared__ int32_t wsum[32];
  __shared__ int32_t carry_s;
  const int col = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < tiles; base += 1024) {
    const int i = base + tid;
    const int32_t v = (i < tiles) ? tile_sum[(int64_t)col * tiles + i] : 0;
    int32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    int32_t off = carry_s;
    for (int k = 0; k < wid; ++k) off += wsum[k];
    if (i < tiles) tile_sum[(int64_t)col * tiles + i] = off + inc - v;
    __syncthreads();
    if (tid == 1023) carry_s = off + inc;
    __syncthreads();
  }
  if (tid == 0) ksum[col] = carry_s;
}
// expansion of one tile: out[offset_of_tile + exclusive_scan(counts) + r] = j, and the fractions are divided by k
__global__ void __launch_bounds__(RES_NT) residual_expand_kernel(const int32_t* __restrict__ counts, float* __restrict__ frac, int64_t n, int64_t ld, int tiles,
                                                                 const int32_t* __restrict__ tile_off, const int32_t* __restrict__ ksum, int32_t* __restrict__ out) {
  __shared__ int32_t wsum[RES_NT / 32];
  const int tile = blockIdx.x, col = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const float kf = (float)ksum[col];
  int32_t c[RES_ITEMS];
  int32_t s = 0;
  const int64_t i0 = (int64_t)tile * RES_TILE + (int64_t)tid * RES_ITEMS;   // blocked: a thread owns 16 consecutive particles
#pragma unroll
  for (int j = 0; j < RES_ITEMS; ++j) {
    const int64_t i = i0 + j;
    c[j] = (i < n) ? counts[(int64_t)col * ld + i] : 0;
    if (i < n) frac[(int64_t)col * ld + i] = __fdiv_rn(frac[(int64_t)col * ld + i], kf);
    s += c[j];
  }
  int32_t inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[wid] = inc;
  __syncthreads();
  int32_t pos = tile_off[(int64_t)col * tiles + tile] + inc - s;
  for (int k = 0; k < wid; ++k) pos += wsum[k];
#pragma unroll
  for (int j = 0; j < RES_ITEMS; ++j) {
    for (int r = 0; r < c[j]; ++r) out[(int64_t)col * ld + pos + r] = (int32_t)(i0 + j);
    pos += c[j];
  }
}

// ---- FFBS, one backward step (filters/particle/base.py:105-128) -----------------------------------------------------------------------
// For every smoothed particle i (its value at the later time is xnext[i]) an index is drawn from the categorical distribution with
// logits  log w_j + log p(xnext_i | x_j)  over the filter's particles j of the earlier time (Categorical(logits=...).sample() in the
// reference), by inversion: first j whose cumulative unnormalised probability reaches U_i * total.  One block per smoothed particle,
// two passes over j (online max / sum, then the ordered search with a block scan per chunk of 256) - O(N^2) like the reference.
// Non-batched filters only, like the reference's working branch.  Layout: the reference's, contiguous: x (N, D), lw (N).
#define SMCB_RNG_FFBS 40u
struct FfbsArgs {
  const float* P;        // parameter row of column 0
  const float* x;        // (N, D) particles of the earlier state
  const float* lw;       // (N) their log-weights
  const float* xnext;    // (N, D) smoothed particles of the later state
  const double* U;       // optional injected uniforms (N)
  int64_t* idx;          // (N) out
  float* xout;           // (N, D) out: x[idx]
  int64_t n;
  uint64_t seed;
  int32_t t;
};
template <int MODEL>
__device__ __forceinline__ float ffbs_logit(const float* xj, float lwj, const float* xn, const float* Ps) {
  typedef Model<MODEL> M;
  float loc[M::D], sc;
  M::loc_scale(xj, Ps, loc, sc);
  const float tot = __fmul_rn(sc, Ps[P_INC_SCALE]);   // std of the transition
  const float inv = 1.0f / tot;
  float q = 0.f;
#pragma unroll
  for (int d = 0; d < M::D; ++d) {
    const float e = __fmul_rn(__fsub_rn(xn[d], loc[d]), inv);
    q = fmaf(e, e, q);
  }
  return lwj - 0.5f * q - (float)M::D * (logf(fabsf(tot)) + SMCB_LOG_SQRT_2PI);
}
template <int MODEL>
__global__ void __launch_bounds__(256) ffbs_step_kernel(FfbsArgs a) {
  typedef Model<MODEL> M;
  constexpr int D = M::D;
  __shared__ float Ps[SMCB_NPARAM];
  __shared__ float sred[2][8];
  __shared__ float wsum[8];
  __shared__ int found;
  __shared__ float carry_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t i = blockIdx.x;
  if (tid < SMCB_NPARAM) Ps[tid] = a.P[tid];
  if (tid == 0) { found = 256; carry_s = 0.f; }
  __syncthreads();
  float xn[D];
#pragma unroll
  for (int d = 0; d < D; ++d) xn[d] = a.xnext[i * D + d];
  // pass 1: max and sum of exp
  float m = -INFINITY, z = 0.f;
  for (int64_t j = tid; j < a.n; j += 256) {
    float xj[D];
#pragma unroll
    for (int d = 0; d < D; ++d) xj[d] = __ldg(a.x + j * D + d);
    const float l = ffbs_logit<MODEL>(xj, __ldg(a.lw + j), xn, Ps);
    if (l > m) { z = z * ((m == -INFINITY) ? 0.f : __expf(m - l)) + 1.f; m = l; }
    else if (l > -INFINITY) z += __expf(l - m);
  }
  float mb = warp_max(m);
  if (lane == 0) sred[0][wid] = mb;
  __syncthreads();
  mb = sred[0][0];
#pragma unroll
  for (int k = 1; k < 8; ++k) mb = fmaxf(mb, sred[0][k]);
  z = (m == -INFINITY) ? 0.f : z * __expf(m - mb);
  z = warp_sum(z);
  if (lane == 0) sred[1][wid] = z;
  __syncthreads();
  float Z = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) Z += sred[1][k];
  double U;
  if (a.U) U = a.U[i];
  else {
    const Philox4 r = philox4x32_10((uint32_t)i, 0u, (uint32_t)a.t, SMCB_RNG_FFBS, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
    U = smcb_u01_double(r.x, r.y);
  }
  const float target = (float)(U * (double)Z);
  // pass 2: first j with cumulative sum >= target, in index order
  for (int64_t base = 0; base < a.n; base += 256) {
    const int64_t j = base + tid;
    float e = 0.f;
    if (j < a.n) {
      float xj[D];
#pragma unroll
      for (int d = 0; d < D; ++d) xj[d] = __ldg(a.x + j * D + d);
      e = __expf(ffbs_logit<MODEL>(xj, __ldg(a.lw + j), xn, Ps) - mb);
    }
    float inc = e;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    float off = carry_s;
    for (int k = 0; k < wid; ++k) off += wsum[k];
    const float cum = off + inc;
    if (j < a.n && cum >= target) atomicMin(&found, tid);
    __syncthreads();
    if (found < 256) {   // block-uniform after the barrier
      if (tid == found) {
        a.idx[i] = j;
#pragma unroll
        for (int d = 0; d < D; ++d) a.xout[i * D + d] = __ldg(a.x + j * D + d);
      }
      return;
    }
    if (tid == 255) carry_s = cum;
    __syncthreads();
  }
  if (tid == 0) {  // rounding left the target above the last cumulative sum: the last particle (searchsorted clamps the same way)
    const int64_t j = a.n - 1;
    a.idx[i] = j;
#pragma unroll
    for (int d = 0; d < D; ++d) a.xout[i * D + d] = __ldg(a.x + j * D + d);
  }
}

// ---- columns of a handle as self-contained records (cross-rank theta-resampling of a sharded SMC2 run) -----------------------------------
// Every per-column array of a handle is described by (pointer, rows, elements per row, stride between rows, stride between columns, all in
// 4-byte elements).  pack: record[column] = the column's slices of all arrays, back to back; unpack: column b of the handle <- record
// [idx[b]] of a buffer that may hold the records of MANY handles (the all-gather of the ranks' packed shards).
#define SMCB_MAX_COLDESC 16
struct ColDesc { uint32_t* ptr; int32_t rows, inner; int64_t row_stride, col_stride; int64_t offset; };   // offset: position inside the record
struct ColPackArgs {
  ColDesc d[SMCB_MAX_COLDESC];
  int32_t ndesc, B;
  int64_t record;          // elements per record
  uint32_t* buf;           // (columns, record)
  const int64_t* idx;      // unpack: source record of every local column (NULL: identity)
  int32_t n_records;       // unpack: records in buf (range check)
  int32_t* bad;
};
template <bool PACK>
__global__ void __launch_bounds__(256) column_pack_kernel(ColPackArgs a) {
  const int b = blockIdx.y;
  int64_t rec = b;
  if (!PACK && a.idx) {
    rec = a.idx[b];
    if (rec < 0 || rec >= a.n_records) { if (a.bad && threadIdx.x == 0 && blockIdx.x == 0) atomicExch(a.bad, 1); return; }
  }
  uint32_t* r = a.buf + rec * a.record;
  for (int k = 0; k < a.ndesc; ++k) {
    const ColDesc d = a.d[k];
    const int64_t count = (int64_t)d.rows * d.inner;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < count; e += (int64_t)gridDim.x * 256) {
      const int64_t row = e / d.inner, in = e - row * d.inner;
      uint32_t* p = d.ptr + row * d.row_stride + (int64_t)b * d.col_stride + in;
      if (PACK) r[d.offset + e] = *p;
      else *p = r[d.offset + e];
    }
  }
}
