// Launch helpers and the small kernels behind the stand-alone operators (pyfilter.utils.normalize / get_ess,
// pyfilter.resampling.systematic / multinomial) - the same resampling kernels the filter handle uses, fed from a caller tensor
// with arbitrary (particle, column) strides.
#pragma once
#include "resample.cuh"

// ---- layout adapters ---------------------------------------------------------------------------------------------------------
// in[i*sn + b*sb]  ->  rows[b*ld + i]   (32x32 shared-memory transpose so both sides stay coalesced for particle-major input)
__global__ void gather_rows_kernel(const float* __restrict__ in, int64_t n, int B, int64_t sn, int64_t sb, float* __restrict__ rows, int64_t ld) {
  __shared__ float tile[32][33];
  const int64_t i0 = (int64_t)blockIdx.x * 32;
  const int b0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // (32, 8)
  if (sn == 1) {  // rows already contiguous: plain copy, x along particles
    for (int k = ty; k < 32; k += 8) {
      const int b = b0 + k;
      const int64_t i = i0 + tx;
      if (b < B && i < n) rows[(int64_t)b * ld + i] = in[i + (int64_t)b * sb];
    }
    return;
  }
  for (int k = ty; k < 32; k += 8) {  // read with x along columns
    const int64_t i = i0 + k;
    const int b = b0 + tx;
    if (i < n && b < B) tile[k][tx] = in[i * sn + (int64_t)b * sb];
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {  // write with x along particles
    const int b = b0 + k;
    const int64_t i = i0 + tx;
    if (b < B && i < n) rows[(int64_t)b * ld + i] = tile[tx][k];
  }
}

// rows[b*ld + i] (int32)  ->  out[i*sn + b*sb] (int64)     (torch.searchsorted / torch.multinomial return int64)
__global__ void scatter_i64_kernel(const int32_t* __restrict__ rows, int64_t n, int B, int64_t ld, int64_t* __restrict__ out, int64_t sn, int64_t sb) {
  __shared__ int32_t tile[32][33];
  const int64_t i0 = (int64_t)blockIdx.x * 32;
  const int b0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  if (sn == 1) {
    for (int k = ty; k < 32; k += 8) {
      const int b = b0 + k;
      const int64_t i = i0 + tx;
      if (b < B && i < n) out[i + (int64_t)b * sb] = (int64_t)rows[(int64_t)b * ld + i];
    }
    return;
  }
  for (int k = ty; k < 32; k += 8) {
    const int b = b0 + k;
    const int64_t i = i0 + tx;
    if (b < B && i < n) tile[k][tx] = rows[(int64_t)b * ld + i];
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int64_t i = i0 + k;
    const int b = b0 + tx;
    if (i < n && b < B) out[i * sn + (int64_t)b * sb] = (int64_t)tile[tx][k];
  }
}

// Columns that are contiguous in the caller's tensor (a 1-D tensor, or the (1, N)-strided result layout): plain coalesced copies, one
// block per 4096 elements of a column - the 32 x 32 transposing kernels above would move 32 elements per block for a single column.
__global__ void __launch_bounds__(256) copy_rows_contig_kernel(const float* __restrict__ in, int64_t n, int64_t sb, float* __restrict__ rows, int64_t ld) {
  const int b = blockIdx.y;
  const int64_t base = (int64_t)blockIdx.x * 4096;
#pragma unroll 4
  for (int j = 0; j < 16; ++j) {
    const int64_t i = base + j * 256 + threadIdx.x;
    if (i < n) rows[(int64_t)b * ld + i] = __ldg(in + i + (int64_t)b * sb);
  }
}
__global__ void __launch_bounds__(256) scatter_i64_contig_kernel(const int32_t* __restrict__ rows, int64_t n, int64_t ld, int64_t* __restrict__ out, int64_t sb) {
  const int b = blockIdx.y;
  const int64_t base = (int64_t)blockIdx.x * 4096;
#pragma unroll 4
  for (int j = 0; j < 16; ++j) {
    const int64_t i = base + j * 256 + threadIdx.x;
    if (i < n) out[i + (int64_t)b * sb] = (int64_t)rows[(int64_t)b * ld + i];
  }
}
static inline void op_launch_gather_rows(const float* in, int64_t n, int B, int64_t sn, int64_t sb, float* rows, int64_t ld, cudaStream_t s) {
  if (sn == 1) {
    copy_rows_contig_kernel<<<dim3((unsigned)((n + 4095) / 4096), B), 256, 0, s>>>(in, n, sb, rows, ld);
    return;
  }
  dim3 g((unsigned)((n + 31) / 32), (unsigned)((B + 31) / 32));
  gather_rows_kernel<<<g, dim3(32, 8), 0, s>>>(in, n, B, sn, sb, rows, ld);
}
static inline void op_launch_scatter_i64(const int32_t* rows, int64_t n, int B, int64_t ld, int64_t* out, int64_t sn, int64_t sb, cudaStream_t s) {
  if (sn == 1) {
    scatter_i64_contig_kernel<<<dim3((unsigned)((n + 4095) / 4096), B), 256, 0, s>>>(rows, n, ld, out, sb);
    return;
  }
  dim3 g((unsigned)((n + 31) / 32), (unsigned)((B + 31) / 32));
  scatter_i64_kernel<<<g, dim3(32, 8), 0, s>>>(rows, n, B, ld, out, sn, sb);
}

// ---- normalisers of a matrix of log-weights (utils.py:49-64, 8-20) --------------------------------------------------------------
struct NormPartial { float m, z, zz; };

__global__ void __launch_bounds__(RS_NT) colstats_partial_kernel(const float* __restrict__ rows, int64_t n, int64_t ld, int nblk, NormPartial* parts) {
  __shared__ float scratch[33];
  const int col = blockIdx.y, blk = blockIdx.x;
  const float* src = rows + (int64_t)col * ld + (int64_t)blk * RS_TILE;
  float v[RS_ITEMS];
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {  // striped: coalesced
    const int64_t i = (int64_t)blk * RS_TILE + j * RS_NT + threadIdx.x;
    v[j] = (i < n) ? smcb_sanitize(src[j * RS_NT + threadIdx.x]) : -INFINITY;
    m = fmaxf(m, v[j]);
  }
  m = block_allreduce<RS_NT>(m, -INFINITY, OpMaxF(), scratch);
  float z = 0.f, zz = 0.f;
  if (m > -INFINITY) {
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
      const float e = (v[j] == -INFINITY) ? 0.f : expf(v[j] - m);
      z += e; zz += e * e;
    }
  }
  z = block_allreduce<RS_NT>(z, 0.f, OpSumF(), scratch);
  zz = block_allreduce<RS_NT>(zz, 0.f, OpSumF(), scratch);
  if (threadIdx.x == 0) { NormPartial p; p.m = m; p.z = z; p.zz = zz; parts[(int64_t)col * nblk + blk] = p; }
}

__global__ void __launch_bounds__(128) colstats_final_kernel(const NormPartial* parts, int nblk, int64_t n, ColStats* stats) {
  __shared__ float scratch[33];
  const int col = blockIdx.x;
  float m = -INFINITY;
  for (int b = threadIdx.x; b < nblk; b += 128) m = fmaxf(m, parts[(int64_t)col * nblk + b].m);
  m = block_allreduce<128>(m, -INFINITY, OpMaxF(), scratch);
  float z = 0.f, zz = 0.f;
  for (int b = threadIdx.x; b < nblk; b += 128) {
    const NormPartial p = parts[(int64_t)col * nblk + b];
    if (p.m > -INFINITY) {
      const float sc = expf(p.m - m);
      z += p.z * sc; zz += p.zz * sc * sc;
    }
  }
  z = block_allreduce<128>(z, 0.f, OpSumF(), scratch);
  zz = block_allreduce<128>(zz, 0.f, OpSumF(), scratch);
  if (threadIdx.x == 0) {
    ColStats st;
    memset(&st, 0, sizeof(st));
    st.m_lw = m; st.z_lw = z; st.inv_z_lw = 1.0f / z;
    st.ess = (z * z) / zz;
    st.resample = 1;
    stats[col] = st;
  }
}

// out[i*sn + b*sb] = normalised weight; all-(-inf) columns (max == -inf, i.e. every entry NaN/+inf) give NaN like the reference
__global__ void apply_weights_kernel(const float* __restrict__ rows, const ColStats* stats, int64_t n, int B, int64_t ld, float* out, int64_t sn, int64_t sb) {
  __shared__ float tile[32][33];
  const int64_t i0 = (int64_t)blockIdx.x * 32;
  const int b0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int k = ty; k < 32; k += 8) {
    const int b = b0 + k;
    const int64_t i = i0 + tx;
    if (b < B && i < n) {
      const ColStats st = stats[b];
      tile[k][tx] = smcb_weight(smcb_sanitize(rows[(int64_t)b * ld + i]), st.m_lw, st.inv_z_lw);
    }
  }
  __syncthreads();
  if (sn == 1) {
    for (int k = ty; k < 32; k += 8) {
      const int b = b0 + k;
      const int64_t i = i0 + tx;
      if (b < B && i < n) out[i + (int64_t)b * sb] = tile[k][tx];
    }
  } else {
    for (int k = ty; k < 32; k += 8) {
      const int64_t i = i0 + k;
      const int b = b0 + tx;
      if (i < n && b < B) out[i * sn + (int64_t)b * sb] = tile[tx][k];
    }
  }
}

__global__ void copy_ess_kernel(const ColStats* st, float* out, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) out[b] = st[b].ess;
}

// get_ess(W, normalized=True) (utils.py:8-20): 1 / sum_i W_i^2 of every column, from weights with arbitrary strides
__global__ void __launch_bounds__(256) ess_normalized_kernel(const float* __restrict__ w, int64_t n, int64_t sn, int64_t sb, float* out) {
  __shared__ double scratch[33];
  const int col = blockIdx.x;
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 256) {
    const double v = (double)w[i * sn + (int64_t)col * sb];
    s += v * v;
  }
  s = block_allreduce<256>(s, 0.0, OpSumD(), scratch);
  if (threadIdx.x == 0) out[col] = (float)(1.0 / s);
}

static inline void op_launch_colstats(const float* rows, int64_t n, int B, int64_t ld, int nblk, NormPartial* parts, ColStats* stats, cudaStream_t s) {
  colstats_partial_kernel<<<dim3(nblk, B), RS_NT, 0, s>>>(rows, n, ld, nblk, parts);
  colstats_final_kernel<<<B, 128, 0, s>>>(parts, nblk, n, stats);
}
__global__ void __launch_bounds__(256) apply_weights_contig_kernel(const float* __restrict__ rows, const ColStats* stats, int64_t n, int64_t ld, float* __restrict__ out, int64_t sb) {
  const int b = blockIdx.y;
  const ColStats st = stats[b];
  const int64_t base = (int64_t)blockIdx.x * 4096;
#pragma unroll 4
  for (int j = 0; j < 16; ++j) {
    const int64_t i = base + j * 256 + threadIdx.x;
    if (i < n) out[i + (int64_t)b * sb] = smcb_weight(smcb_sanitize(rows[(int64_t)b * ld + i]), st.m_lw, st.inv_z_lw);
  }
}
static inline void op_launch_apply_weights(const float* rows, const ColStats* stats, int64_t n, int B, int64_t ld, float* out, int64_t sn, int64_t sb, cudaStream_t s) {
  if (sn == 1) {
    apply_weights_contig_kernel<<<dim3((unsigned)((n + 4095) / 4096), B), 256, 0, s>>>(rows, stats, n, ld, out, sb);
    return;
  }
  dim3 g((unsigned)((n + 31) / 32), (unsigned)((B + 31) / 32));
  apply_weights_kernel<<<g, dim3(32, 8), 0, s>>>(rows, stats, n, B, ld, out, sn, sb);
}
static inline void op_launch_copy_ess(const ColStats* st, float* out, int B, cudaStream_t s) {
  copy_ess_kernel<<<(B + 127) / 128, 128, 0, s>>>(st, out, B);
}

// ---- reader of the peer-memory exchange (common.cuh: smcb_exchange_publish) ------------------------------------------------------------
// Polls the (value, tag) pairs of exchange `seq` in this rank's buffer until every column of the batch has arrived (strong system-scope
// loads: the pairs are written by other GPUs) and writes the values out densely: out[0 .. total) increments, out[total .. 2 total) totals.
__global__ void __launch_bounds__(256) exchange_wait_kernel(const unsigned long long* buf, int total, uint32_t seq, float* out, long long* spins) {
  const unsigned long long* base = buf + (int64_t)(seq & 1u) * 2 * total;
  long long n = 0;
  for (int i = threadIdx.x; i < 2 * total; i += blockDim.x) {
    unsigned long long v;
    for (;;) {
      asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(base + i) : "memory");
      if ((uint32_t)(v >> 32) == seq) break;
      __nanosleep(100);
      ++n;
    }
    out[i] = __uint_as_float((uint32_t)v);
  }
  if (spins && n) atomicAdd((unsigned long long*)spins, (unsigned long long)n);
}

// ---- systematic ------------------------------------------------------------------------------------------------------------------
static inline void op_launch_systematic(const ResampleArgs& r, cudaStream_t s) {
  const dim3 g(r.tiles_per_col, r.B);
  normalize_kernel<<<g, RS_NT, 0, s>>>(r);
  describe_kernel<53><<<g, RS_NT, 0, s>>>(r);
  expand_kernel<53, RS_OUT_ANCESTORS><<<g, RS_NT, 0, s>>>(r);
}

// ---- multinomial (resampling.py:55-65 -> ATen multinomial_with_replacement_kernel on CPU) --------------------------------------------
//   c32_k = fl32(c32_{k-1} + W_k)  (sequential float32 prefix, reproduced exactly by the transducer scan with a 24-bit mantissa),
//   cn_k = fl32(c32_k / c32_{n-1});  draw i:  ancestor = first k with (double)cn_k >= U_i,  U_i float64 uniforms in draw order.
struct MultinomialArgs {
  const float* c;          // (B, ld) sequential float32 prefix sums
  int64_t n, ld;
  int32_t B;
  const ColStats* stats;   // resample flags (may be NULL)
  const double* U;         // optional injected uniforms (B, U_pitch)
  int64_t U_pitch;
  uint64_t seed;
  const Ctrl* ctrl;
  int32_t* anc;            // (B, ld)
  int32_t stride;          // coarse table: last element of every `stride` cumulative weights (power of two)
  int32_t ncoarse;
  const float* mid;        // (B, nmid) packed middle level: last element of every MN_MID cumulative weights - the 32+ probes of a chunk
  int32_t nmid;            //   share one or two 128-byte lines instead of touching a different line of `c` each
  int32_t col0;            // global index of column 0 (Philox counter)
  const int32_t* draw_offset;  // residual resampling: the first draw_offset[col] outputs are the deterministic copies (NULL: none)
};

// Draw i picks the first k with (double) fl32(c_k / c_{n-1}) >= U_i.  The predicate is monotone in c_k, so it is turned into a
// threshold on c_k itself once per draw - no division inside the search: with Uc = U rounded up to float32 and Up its predecessor,
// fl32(c / total) >= Uc  <=>  c >= mid(Up, Uc) * total (exact in fp64: 25 x 24 bits), strictly when the tie rounds to the odd Up.
// Two-level search: a coarse table of every `stride`-th cumulative weight in shared memory, then one chunk in global memory.
#define MN_MID 64
__global__ void __launch_bounds__(256) multinomial_mid_kernel(const float* __restrict__ c, int64_t n, int64_t ld, float* __restrict__ mid, int nmid) {
  pdl_wait();
  const int col = blockIdx.y;
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j < nmid) mid[(int64_t)col * nmid + j] = __ldg(c + (int64_t)col * ld + min((int64_t)(j + 1) * MN_MID, n) - 1);
}
__global__ void __launch_bounds__(256) multinomial_draw_kernel(MultinomialArgs a) {
  extern __shared__ float coarse[];
  pdl_wait();
  const int col = blockIdx.y;
  if (a.stats && !a.stats[col].resample) return;
  const float* c = a.c + (int64_t)col * a.ld;
  const int64_t n = a.n;
  for (int i = threadIdx.x; i < a.ncoarse; i += blockDim.x) {
    const int64_t k = min((int64_t)(i + 1) * a.stride, n) - 1;
    coarse[i] = __ldg(c + k);
  }
  const double total = (double)__ldg(c + n - 1);
  const int t = a.ctrl->t;
  const int64_t first = a.draw_offset ? (int64_t)a.draw_offset[col] : 0;   // draws go to anc[first ..], uniforms are consumed from 0
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n - first; i += (int64_t)gridDim.x * blockDim.x) {
    double U;
    if (a.U) U = a.U[(int64_t)col * a.U_pitch + i];
    else {
      Philox4 r = philox4x32_10((uint32_t)i, (uint32_t)(col + a.col0), (uint32_t)t, SMCB_RNG_MULTINOMIAL, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
      U = smcb_u01_double(r.x, r.y);
    }
    const float Uc = __double2float_ru(U);
    float cthr = 0.f;  // U <= 0: every cumulative weight qualifies
    if (Uc > 0.f) {
      const uint32_t ub = __float_as_uint(Uc);
      const double mid = 0.5 * ((double)__uint_as_float(ub - 1u) + (double)Uc);
      double thr = mid * total;
      if (ub & 1u) thr = __longlong_as_double(__double_as_longlong(thr) + 1);  // a tie rounds to the even neighbour Up: need c/total > mid
      cthr = __double2float_ru(thr);
    }
    int lo = 0, hi = a.ncoarse;  // first chunk whose last element reaches the threshold
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (coarse[mid] < cthr) lo = mid + 1; else hi = mid;
    }
    int64_t ans = n - 1;
    if (lo < a.ncoarse) {
      int64_t k0 = (int64_t)lo * a.stride, k1 = min(k0 + a.stride, n) - 1;
      if (a.mid) {  // middle level: first group of MN_MID elements of the chunk whose last element reaches the threshold
        const float* midp = a.mid + (int64_t)col * a.nmid;
        int m0 = (int)(k0 / MN_MID), m1 = (int)(k1 / MN_MID);   // groups of the chunk; the last one always qualifies
        while (m0 < m1) {
          const int m = (m0 + m1) >> 1;
          if (__ldg(midp + m) < cthr) m0 = m + 1; else m1 = m;
        }
        k0 = (int64_t)m0 * MN_MID;
        k1 = min(k0 + MN_MID, n) - 1;
      }
      while (k0 < k1) {
        const int64_t m = (k0 + k1) >> 1;
        if (__ldg(c + m) < cthr) k0 = m + 1; else k1 = m;
      }
      ans = k0;
    }
    a.anc[(int64_t)col * a.ld + first + i] = (int32_t)ans;
  }
}

// r.c_out must point at a (B, ld) float scratch buffer owned by the caller; normalize_kernel has already been enqueued
static inline void op_launch_multinomial_after_normalize(const ResampleArgs& r, const double* U, int64_t U_pitch, cudaStream_t s) {
  const dim3 g(r.tiles_per_col, r.B);
  describe_kernel<24><<<g, RS_NT, 0, s>>>(r);
  expand_kernel<24, RS_OUT_CUMSUM><<<g, RS_NT, 0, s>>>(r);
  MultinomialArgs m;
  m.c = r.c_out; m.n = r.n; m.ld = r.ld; m.B = r.B; m.stats = r.stats; m.U = U; m.U_pitch = U_pitch;
  m.seed = r.seed; m.ctrl = r.ctrl; m.anc = r.anc; m.col0 = r.col0; m.draw_offset = r.draw_offset;
  int stride = 256;  // every block gathers the coarse table itself (one sector per entry): keep it to ~1k entries
  while ((r.n + stride - 1) / stride > 1024) stride *= 2;
  m.stride = stride; m.ncoarse = (int)((r.n + stride - 1) / stride);
  m.mid = nullptr; m.nmid = 0;
  if (r.mid_out && stride % MN_MID == 0) {
    m.nmid = (int)((r.n + MN_MID - 1) / MN_MID);
    multinomial_mid_kernel<<<dim3((m.nmid + 255) / 256, r.B), 256, 0, s>>>(r.c_out, r.n, r.ld, r.mid_out, m.nmid);
    m.mid = r.mid_out;
  }
  int bx = (int)((r.n + 255) / 256);
  const int cap = (smcb_sm_count() * 6 + r.B - 1) / r.B;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  multinomial_draw_kernel<<<dim3(bx, r.B), 256, (size_t)m.ncoarse * sizeof(float), s>>>(m);
}
static inline int op_launch_multinomial(const ResampleArgs& r, const double* U, int64_t U_pitch, cudaStream_t s) {
  if (!r.c_out) return SMCB_EINVAL;
  normalize_kernel<<<dim3(r.tiles_per_col, r.B), RS_NT, 0, s>>>(r);
  op_launch_multinomial_after_normalize(r, U, U_pitch, s);
  return SMCB_OK;
}
