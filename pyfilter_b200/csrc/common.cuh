// Shared device-side plumbing: per-column statistics, block partial records, warp/block reductions, acquire/release.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "models.h"

#define SMCB_FLT_MAX 3.402823466e+38f

// number of SMs of the current device (148 on B200), queried once per process
static inline int smcb_sm_count() {
  static int v = 0;
  if (v <= 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) {
      cudaGetLastError();
      v = 148;
    }
  }
  return v;
}

// Per-column summary written by the finalize kernel and read by the next step's kernels (all on device, no host sync).
struct ColStats {
  float m_lw, z_lw, inv_z_lw;   // max, sum exp(lw - max) and its reciprocal for the log-weights (filters/particle/state.py:148)
  float m_rw, z_rw, inv_z_rw;   // the same for the APF resampling log-weights  g + lw   (filters/particle/apf.py:29)
  float ess;                    // 1 / sum W^2                                            (utils.py:8-20)
  int32_t resample;             // 1 when the coming step resamples this column            (sisr.py:19 / apf.py:31)
  float shift[3];               // moment shift (= latest filter mean) used to accumulate the variance in one pass
  float ll_aux;                 // log sum_i W_i exp(g_i)  (second term of apf.py:44), valid for the coming step
  int32_t fold_valid;           // 1 when rw already holds g_{t+1} + lw_t for the coming step (look-ahead folded by the step kernel)
  float pad[3];
};

// One record per (column, block) of the step / pre-weight / init kernels: soft-max partials.
struct __align__(16) Partial {  // 64 bytes: read back with 128-bit loads
  float m1, z1, zz1;            // over lw:  max, sum e, sum e^2          (e = exp(lw - m1))
  float sx[3], sxx[3];          // sum e (x - shift), sum e (x - shift)^2
  float m2, z2;                 // over rw (APF look-ahead folded)
  float m3, z3;                 // SISR likelihood increment: max inc, sum W_prev exp(inc - m3)   (filters/particle/utils.py:16-22)
  float pad[3];
};
static_assert(sizeof(Partial) == 64, "Partial must stay 64 bytes");

// Device control block: lets one captured CUDA graph serve every time step (no per-step kernel arguments change).
struct Ctrl {
  int32_t t;            // number of completed filter moves (row index into the moment history is t + 1)
  int32_t y_base;       // time index of y[0]
  int32_t y_count;      // observations available at y
  int32_t ticket;       // last-block-done counter of the finalize kernel
  uint32_t epoch;       // unused (the slot tag of resample_fused_kernel is a kernel argument: ResampleArgs.epoch_host)
  uint32_t tile_counter;  // unused (tile id = block id)
  int32_t slow_tiles;   // diagnostics: tiles that took the sequential fallback of the exact scan
  int32_t lb_fail;      // reserved
  const float* y;       // (y_count, OD) observations on device
  int64_t lb_windows;   // reserved
};

// Exchange of the per-column log-likelihood values between the GPUs that hold shards of ONE batch of filters (theta-sharded SMC2 /
// NESS): the kernel that finalises a column stores (value, sequence tag) pairs straight into EVERY rank's buffer
// over NVLink peer memory - one 8-byte store per rank and value: the pair is written atomically, so no fence, no flag and no
// collective launch is needed; the reader polls the tags (exchange_wait_kernel).  Buffers: torch symmetric memory (plumbing),
// layout  slot[parity = seq & 1][kind: 0 = increment of the last move, 1 = running total][global column]  of 8 bytes each.
#define SMCB_MAX_PEERS 8
struct ExchangeArgs {
  unsigned long long* peer[SMCB_MAX_PEERS];  // base of every rank's buffer (this rank's own included), NULL when no exchange is attached
  int32_t world;
  int32_t total;        // columns of the whole batch
  int32_t lo;           // global index of this rank's first column
  uint32_t seq;         // sequence number of the exchange this launch publishes (>= 1); 0: do not publish
};
__device__ __forceinline__ void smcb_exchange_publish(const ExchangeArgs& x, int col, float ll, float ll_total) {
  if (!x.seq) return;
  const int64_t base = (int64_t)(x.seq & 1u) * 2 * x.total + x.lo + col;
  const unsigned long long a = ((unsigned long long)x.seq << 32) | (unsigned long long)__float_as_uint(ll);
  const unsigned long long b = ((unsigned long long)x.seq << 32) | (unsigned long long)__float_as_uint(ll_total);
  for (int r = 0; r < x.world; ++r) {
    unsigned long long* p = x.peer[r] + base;
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(a) : "memory");
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p + x.total), "l"(b) : "memory");
  }
}

// Programmatic dependent launch (sm_90+): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start
// while its predecessor in the stream drains; pdl_wait() blocks until the predecessor has completed and its writes are visible
// (a no-op for ordinary launches), pdl_trigger() lets the successor's blocks be scheduled once every block has called it or exited.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float smcb_sanitize(float w) {
  // utils.py:57 nan_to_num_(nan=-inf, posinf=-inf) with neginf left at its default (lowest finite float)
  if (w != w || w == INFINITY) return -INFINITY;
  if (w == -INFINITY) return -SMCB_FLT_MAX;
  return w;
}

// normalised weight exactly as every kernel of this library evaluates it (must be ONE function: every kernel has to see identical
// bits).  exp(v) = 2^t (1 + r ln 2) with t = fl(v log2 e) and r the exact residual of that product plus the low part of log2 e: one
// SFU ex2 (2 ulp) and five FMA-pipe instructions, ~3e-7 relative for any v <= 0 (tolerance of the parity tests: 3e-6).
__device__ __forceinline__ float smcb_weight(float lw, float m, float inv_z) {
  const float v = fmaxf(__fsub_rn(lw, m), -200.f);                         // exp(-200) = 0 in float32; keeps -inf out of the residual
  const float t = __fmul_rn(v, 1.4426950216293335f);                       // log2(e), high part
  float r = fmaf(v, 1.4426950216293335f, -t);                              // exact residual of the product
  r = fmaf(v, 1.9259629911766e-8f, r);                                     // + v * low part of log2(e)
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
  e = fmaf(e, __fmul_rn(r, 0.6931471805599453f), e);
  return __fmul_rn(e, inv_z);
}

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// warp maximum in ONE instruction (sm_100a: redux.sync.max.f32 -> CREDUX.MAX.F32); inputs are never NaN here
__device__ __forceinline__ float warp_redux_max(float v) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide max / sum for blockDim.x == NT (multiple of 32, <= 1024); result broadcast to every thread.  `scratch` holds
// at least 33 elements.  Deterministic (fixed tree).
template <int NT, typename T, typename Op>
__device__ __forceinline__ T block_allreduce(T v, T ident, Op op, T* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();  // protect scratch reuse
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  T r = (lane < NT / 32) ? scratch[lane] : ident;
#pragma unroll
  for (int o = 16; o; o >>= 1) r = op(r, __shfl_xor_sync(0xffffffffu, r, o));
  return r;
}
struct OpMaxF { __device__ __forceinline__ float operator()(float a, float b) const { return fmaxf(a, b); } };
struct OpSumF { __device__ __forceinline__ float operator()(float a, float b) const { return a + b; } };
struct OpSumD { __device__ __forceinline__ double operator()(double a, double b) const { return a + b; } };
struct OpSumI { __device__ __forceinline__ int operator()(int a, int b) const { return a + b; } };
