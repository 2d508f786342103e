// Systematic resampling on the device, bit-exact against the reference's CPU path
//     probs = (arange(n) + u) / n ; cumsum = W.cumsum(-1) ; cumsum[-1] = 1 ; searchsorted(cumsum, probs)     (resampling.py:44-50)
//
//   normalize_kernel   normalised weights, their fp64 tile sums and exclusive tile prefix, the column verdict (benign or not)
//   describe_kernel    general columns only: every tile's transducer descriptor (scan_tile.h); the block that completes a column
//                      chains the descriptors into the EXACT state before every tile
//   expand_kernel      exact cumulative weights of the tile, then the ancestors by EXPANSION: particle j owns the output slots
//                      [count(c_{j-1}), count(c_j)); marks in shared memory, a "last mark" scan, coalesced 128-bit stores
// No kernel waits on another CTA: the dependencies between tiles are carried by the kernel boundaries and the small tails.
// Layout: one column (independent filter) is a contiguous row of `ld` floats; tiles never straddle columns.
#pragma once
#include "common.cuh"
#include "scan_tile.h"
#include "philox.h"

#define RS_NT 256
#define RS_ITEMS 16
#define RS_TILE (RS_NT * RS_ITEMS)
#define RS_MAXSEG 192

// has_special of a published XsDesc: 0 / 1 as in scan_tile.h, or
#define RS_KIND_TABLE 2   // more than one special element: the segment table sits in global memory (e1 = number of specials)
#define RS_KIND_RAW 3     // too many segments: the chain walks the raw weights
#define RS_KIND_NONE 5    // no (alternative) descriptor
#define RS_KIND_CROSS 6   // crossing descriptor: per-thread sums under both binades in the tile's table slot (a_s, b_s = the two totals)
#define RS_KIND_ABS 4     // the state after the tile is a_s whatever came before (tile 0 resolves itself: it starts from 0)

struct __align__(16) FusedSlot {  // written and read with one 128-bit access: the tag says the sum belongs to this launch
  double sum;
  unsigned long long tag;
};

struct SegTable {
  XsT agg[RS_MAXSEG];
  float wc[RS_MAXSEG];
  int32_t e[RS_MAXSEG];
};

struct ResampleArgs {
  const float* w;          // (B, ld) log-weights (or normalised weights when input_is_w)
  float* wn;               // (B, ld) normalised weights: written by normalize_kernel, consumed by the scan (== w when input_is_w)
  int64_t n;               // particles per column
  int64_t ld;              // row pitch (multiple of RS_TILE)
  int32_t B;
  int32_t tiles_per_col;
  int32_t input_is_w;      // 1: `w` already holds normalised weights (stand-alone operator, normalized=True)
  int32_t use_rw;          // 1: normalisers are (m_rw, inv_z_rw) (APF), 0: (m_lw, inv_z_lw) (SISR)
  const ColStats* stats;   // per column; also carries the per-column `resample` flag (NULL => every column resamples)
  const float* u_in;       // optional injected offsets (B)
  float* u_out;            // optional dump of the offsets used (B)
  uint64_t seed;
  double* tilesum;         // (B, tiles_per_col) fp64 sum of every tile
  double* prefix;          // (B, tiles_per_col) exclusive prefix of the tile sums (exact in a benign column)
  double* sin;             // (B, tiles_per_col) general columns: exact state before every tile (describe_kernel's chain)
  int32_t* tileflag;       // (B, tiles_per_col) bit 0: the tile's speculation failed, expand it sequentially
  XsDesc* desc;            // (B, tiles_per_col)
  XsDesc* desc2;           // (B, tiles_per_col) alternative descriptor of tiles next to a binade crossing (has_special = RS_KIND_NONE: absent)
  SegTable* tables;        // (B, tiles_per_col)
  int32_t* anc;            // (B, ld) ancestors out
  float* w_out;            // optional dump of the normalised weights used (B, ld)
  float* c_out;            // OUT_CUMSUM: the emulated sequential prefix sums (B, ld)
  Ctrl* ctrl;
  uint32_t* tilemin;       // (B, tiles_per_col) smallest non-zero weight of every tile (float bits; 0 when a weight is negative)
  int32_t* ncounter;       // (B) last-block-done tickets of normalize_kernel (self resetting)
  int32_t* dcounter;       // (B) last-block-done tickets of describe_kernel (self resetting)
  int32_t* verdict;        // (B) bit 0: the column is "benign" - its sequential fp64 prefix sum never rounds; bit 1: n and u allow
                           //     the lean probe count
  float* u_col;            // (B) the systematic offset of every column for this launch (injected or Philox)
  long long* dbg;          // optional diagnostics (SMCB_DEBUG_TIMELINE): globaltimer stamps / counters of describe_kernel's chain
  int32_t quantize;        // 1: round the weights derived from log-weights to multiples of 2^-52 (every column becomes benign)
  struct FusedSlot* fslots; // (B, tiles_per_col) resample_fused_kernel: tile sums tagged with the launch epoch
  int32_t presanitized;    // 1: the log-weights were stored by the step / state kernels, nan_to_num (utils.py:57) already applied
  int32_t force_benign;    // 1: the host skipped describe_kernel (quantised weights, n <= 2^23, Philox offsets): benign by construction
  int32_t t_host;          // resample_fused_kernel: the move index (== ctrl->t), passed by the host to keep it off the critical path
  unsigned long long epoch_host;  // resample_fused_kernel: launch tag of the slots, unique per launch and never 0
  int32_t col0;            // global index of column 0 (smcb_config.column_offset): the Philox counters use col0 + column
  float* mid_out;          // multinomial: (B, ceil(n / 64)) scratch for the packed middle level of the draw's search (NULL: two levels)
  const int32_t* draw_offset;  // multinomial draws of the residual resampler: column b draws n - draw_offset[b] ancestors into anc[draw_offset[b] ..]
};
__device__ __forceinline__ long long rs_now() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
enum { RS_OUT_ANCESTORS = 0, RS_OUT_CUMSUM = 1 };

// ---- pre-pass: normalised weights (written once, zero beyond n), their fp64 tile sums and the column verdict ---------------------
// A column is BENIGN when every non-zero weight is a multiple of 2^-52 (w >= 2^-29) and the weights sum to less than 2: then
// every partial sum is a multiple of 2^-52 below 2, i.e. exactly representable, so the reference's sequential fp64 prefix sum
// never rounds and equals the exact real prefix sum in ANY association.  Such columns need no binade labels, no transducers
// and no chaining: the exact state before a tile is the plain sum of the preceding tile sums.
struct OpMinU { __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a < b ? a : b; } };
#define RS_BENIGN_MIN_BITS 0x31000000u  // 2^-29

// exclusive block scan of one double per thread (blockDim.x == RS_NT); *total = sum over the block
template <int NT = RS_NT, bool PROTECT = true>  // PROTECT = false: `scratch` is known to be idle (saves a barrier)
__device__ __forceinline__ double rs_block_excl_scan_d(double v, double* scratch /*>= NT/32*/, double* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (PROTECT) __syncthreads();
  if (lane == 31) scratch[wid] = inc;
  __syncthreads();
  double off = 0.0, tot = 0.0;
  if (NT > 256) {  // one lane per warp: cost independent of the block size (exact in any order for the rounding-free weights)
    const double s = (lane < NT / 32) ? scratch[lane] : 0.0;
    tot = warp_sum(s);
    off = warp_sum((lane < wid) ? s : 0.0);
  } else {
#pragma unroll
    for (int k = 0; k < NT / 32; ++k) {
      const double s = scratch[k];
      off += (k < wid) ? s : 0.0;
      tot += s;
    }
  }
  *total = tot;
  return off + (inc - v);
}

__global__ void __launch_bounds__(RS_NT) normalize_kernel(ResampleArgs a) {
  __shared__ double scratch[33];
  __shared__ uint32_t uscratch[33];
  __shared__ int is_last;
  const int col = blockIdx.y, tile = blockIdx.x;
  const int64_t off = (int64_t)col * a.ld + (int64_t)tile * RS_TILE;
  const int64_t g0 = (int64_t)tile * RS_TILE;
  pdl_wait();
  // every independent global load first (striped float4: fully coalesced)
  float4 q4[RS_ITEMS / 4];
#pragma unroll
  for (int v = 0; v < RS_ITEMS / 4; ++v) q4[v] = __ldg(reinterpret_cast<const float4*>(a.w + off + (v * RS_NT + threadIdx.x) * 4));
  ColStats st;
  st.resample = 1; st.m_lw = 0.f; st.inv_z_lw = 1.f; st.m_rw = 0.f; st.inv_z_rw = 1.f;
  if (a.stats) st = a.stats[col];
  // what the tail of the last block needs, fetched now (a dependent global round trip in the serial tail costs ~1 us)
  int t_now = 0;
  float u_inj = 0.f;
  if (threadIdx.x == 0) { t_now = a.ctrl->t; if (a.u_in) u_inj = a.u_in[col]; }
  if (!st.resample) return;
  if (a.dbg && threadIdx.x == 0) atomicMin((unsigned long long*)&a.dbg[8], (unsigned long long)rs_now());
  float m = 0.f, iz = 1.f;
  if (!a.input_is_w) {
    m = a.use_rw ? st.m_rw : st.m_lw;
    iz = a.use_rw ? st.inv_z_rw : st.inv_z_lw;
  }
  double s = 0.0;
  uint32_t key = 0xFFFFFFFFu;
  const bool quant = a.quantize && !a.input_is_w && m == m && iz == iz && iz < 1e30f;  // finite normalisers only
  const bool inner = g0 + RS_TILE <= a.n;            // no padding in this tile
  const bool raw = !a.input_is_w && !a.presanitized;  // apply nan_to_num here
#pragma unroll
  for (int v = 0; v < RS_ITEMS / 4; ++v) {
    const int e = (v * RS_NT + threadIdx.x) * 4;
    float x[4] = {q4[v].x, q4[v].y, q4[v].z, q4[v].w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (raw) x[k] = smcb_sanitize(x[k]);
      if (!a.input_is_w) x[k] = smcb_weight(x[k], m, iz);
      if (!inner && g0 + e + k >= a.n) x[k] = 0.f;
      double xd = (double)x[k];
      if (quant) {  // RN to a multiple of 2^-52: exact in float32 (at most 23 significant bits below 2^-29), |dW| <= 2^-53
        xd = __dadd_rn(__dadd_rn(1.0, xd), -1.0);
        x[k] = (float)xd;
      } else {      // the verdict needs the smallest non-zero weight (quantised weights are multiples of 2^-52 by construction)
        const uint32_t b = __float_as_uint(x[k]);
        const uint32_t kk = (b == 0u) ? 0xFFFFFFFFu : ((b >> 31) ? 0u : b);
        key = min(key, kk);
      }
      s += xd;
    }
    if (!a.input_is_w || a.wn != a.w) *reinterpret_cast<float4*>(a.wn + off + e) = make_float4(x[0], x[1], x[2], x[3]);
    if (a.w_out) *reinterpret_cast<float4*>(a.w_out + off + e) = make_float4(x[0], x[1], x[2], x[3]);
  }
  s = block_allreduce<RS_NT>(s, 0.0, OpSumD(), scratch);
  key = block_allreduce<RS_NT>(key, 0xFFFFFFFFu, OpMinU(), uscratch);
  pdl_trigger();  // the successor may be scheduled while the last block runs the serial tail
  if (threadIdx.x == 0) {
    a.tilesum[(int64_t)col * a.tiles_per_col + tile] = s;
    a.tilemin[(int64_t)col * a.tiles_per_col + tile] = key;
    __threadfence();
    is_last = (atomicAdd(&a.ncounter[col], 1) == a.tiles_per_col - 1);
  }
  __syncthreads();
  if (!is_last) return;
  // ---- the block that completes a column: exclusive prefix of the tile sums, systematic offset, verdict
  __threadfence();
  if (a.dbg && threadIdx.x == 0) a.dbg[9] = rs_now();
  const int T = a.tiles_per_col;
  const int per = (T + RS_NT - 1) / RS_NT;
  const int q0 = min(T, (int)threadIdx.x * per), q1 = min(T, q0 + per);
  const double* ts = a.tilesum + (int64_t)col * T;
  double part = 0.0;
  uint32_t mk = 0xFFFFFFFFu;
  for (int q = q0; q < q1; ++q) {
    part += __ldcg(ts + q);
    mk = min(mk, __ldcg(a.tilemin + (int64_t)col * T + q));
  }
  double tot;
  double run = rs_block_excl_scan_d(part, scratch, &tot);
  for (int q = q0; q < q1; ++q) {
    a.prefix[(int64_t)col * T + q] = run;
    run += __ldcg(ts + q);
  }
  mk = block_allreduce<RS_NT>(mk, 0xFFFFFFFFu, OpMinU(), uscratch);
  if (threadIdx.x == 0) {
    float u;  // one uniform per column (resampling.py:41)
    if (a.u_in) u = u_inj;
    else {
      Philox4 r = philox4x32_10((uint32_t)(col + a.col0), 0u, (uint32_t)t_now, SMCB_RNG_SYSTEMATIC, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
      u = smcb_u01(r.x);
    }
    a.u_col[col] = u;
    if (a.u_out) a.u_out[col] = u;
    const bool u_ok = (u == 0.f) || (u >= 5.5e-20f && u < 1.0f);
    const bool fast_ok = a.n <= (1 << 23) && u_ok;  // xs_count_fast applies (exact_scan.h)
    // quantised weights are multiples of 2^-52 by construction (the smallest-weight proxy does not apply to them)
    const bool quant = a.quantize && !a.input_is_w;  // weights exp(.) * c >= 0, multiples of 2^-52 (NaN normalisers give tot = NaN)
    const bool benign = fast_ok && tot < 1.5 && (quant || mk >= RS_BENIGN_MIN_BITS);
    a.verdict[col] = (benign ? 1 : 0) | (fast_ok ? 2 : 0);
    a.ncounter[col] = 0;
    if (a.dbg) a.dbg[10] = rs_now();
  }
}

// ---- block-level pieces of the exact tile scan ------------------------------------------------------------------------------------
__device__ __forceinline__ void rs_cbar() { asm volatile("bar.sync 1, %0;" ::"n"(RS_NT) : "memory"); }

__device__ __forceinline__ XsSeg rs_shfl_up(const XsSeg& s, int o) {
  XsSeg r;
  r.t.s = __shfl_up_sync(0xffffffffu, s.t.s, o);
  r.t.d = __shfl_up_sync(0xffffffffu, s.t.d, o);
  r.cnt = __shfl_up_sync(0xffffffffu, s.cnt, o);
  return r;
}

// exclusive segmented scan of the per-thread contributions; *total = inclusive result of the whole block.
// `lab` = binade label at the start of the calling thread (needed when a tie makes a composition parity dependent).
template <int MB>
__device__ __forceinline__ XsSeg rs_block_excl_scan_seg(const XsSeg& v, int lab, XsSeg* scratch /*>=8*/, int* lab_scratch /*>=8*/,
                                                        XsSeg* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  XsSeg inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    XsSeg t = rs_shfl_up(inc, o);
    if (lane >= o) inc = xs_seg_combine<MB>(t, inc, lab);
  }
  XsSeg excl = rs_shfl_up(inc, 1);
  if (lane == 0) excl = xs_seg_identity();
  rs_cbar();
  if (lane == 31) scratch[wid] = inc;
  if (lane == 0) lab_scratch[wid] = lab;
  rs_cbar();
  XsSeg run = xs_seg_identity(), mine = xs_seg_identity();
#pragma unroll
  for (int k = 0; k < RS_NT / 32; ++k) {  // every thread folds the 8 warp totals itself (no second shuffle stage, no extra barrier)
    if (k == wid) mine = run;
    run = xs_seg_combine<MB>(run, scratch[k], lab_scratch[k]);
  }
  *total = run;
  return xs_seg_combine<MB>(mine, excl, lab);
}

__device__ __forceinline__ XsDesc rs_shfl_desc(const XsDesc& d, int src) {
  XsDesc r;
  r.a_s = __shfl_sync(0xffffffffu, d.a_s, src);
  r.b_s = __shfl_sync(0xffffffffu, d.b_s, src);
  r.wc = __shfl_sync(0xffffffffu, d.wc, src);
  int packed = ((int)(uint16_t)d.e0) | ((int)(uint16_t)d.e1 << 16);
  packed = __shfl_sync(0xffffffffu, packed, src);
  r.e0 = (int16_t)(packed & 0xffff);
  r.e1 = (int16_t)((uint32_t)packed >> 16);
  int p2 = ((int)(uint8_t)d.a_d) | ((int)(uint8_t)d.b_d << 8) | ((int)(uint8_t)d.has_special << 16);
  p2 = __shfl_sync(0xffffffffu, p2, src);
  r.a_d = (int8_t)(p2 & 0xff);
  r.b_d = (int8_t)((p2 >> 8) & 0xff);
  r.has_special = (int8_t)((p2 >> 16) & 0xff);
  r.pad = 0;
  return r;
}

// descriptors are written by other CTAs of the same launch: read them through L2
__device__ __forceinline__ XsDesc rs_read_desc(const XsDesc* p) {
  XsDesc d;
  const double2 q0 = __ldcg(reinterpret_cast<const double2*>(p));
  const int4 q1 = __ldcg(reinterpret_cast<const int4*>(p) + 1);
  d.a_s = q0.x; d.b_s = q0.y;
  d.wc = __int_as_float(q1.x);
  d.e0 = (int16_t)(q1.y & 0xffff);
  d.e1 = (int16_t)((uint32_t)q1.y >> 16);
  d.a_d = (int8_t)(q1.z & 0xff);
  d.b_d = (int8_t)((q1.z >> 8) & 0xff);
  d.has_special = (int8_t)((q1.z >> 16) & 0xff);
  d.pad = 0;
  return d;
}

struct RsTileSmem {
  double dscratch[33];
  XsSeg sscratch[8];
  int lscratch[8];
  int lab_end[RS_NT];
  XsT seg_agg[RS_MAXSEG];
  double base[RS_MAXSEG];
  float seg_wc[RS_MAXSEG];
  int seg_e[RS_MAXSEG];
  int X, e0, table_ok, ok, is_last;
  double S_out, sp_end;
};

// rare paths, kept out of line so that the hot loops stay small (instruction-cache footprint)
template <int MB>
__device__ __noinline__ void rs_thread_reduce_slow(const float (&w)[RS_ITEMS], double sp_thread, int lab_prev, int lab_end,
                                                   uint32_t* mask, XsSeg* contrib) {
  xs_thread_reduce<MB, RS_ITEMS>(w, sp_thread, lab_prev, lab_end, mask, contrib);
}
template <int MB>
__device__ __noinline__ void rs_thread_finalize_slow(const float (&w)[RS_ITEMS], uint32_t mask, int lab_prev, int s_base, XsT t_open,
                                                     const double* base, const int* seg_e, float* c /*RS_ITEMS, shared memory*/) {
  float cl[RS_ITEMS];
  xs_thread_finalize<MB, RS_ITEMS>(w, mask, lab_prev, s_base, t_open, base, seg_e, cl);
  for (int j = 0; j < RS_ITEMS; ++j) c[j] = cl[j];
}
template <int MB>
__device__ __noinline__ void rs_fill_table(const float (&w)[RS_ITEMS], uint32_t mask, double sp_thread, int lab_prev, int lab_end,
                                           XsSeg excl, RsTileSmem& sm) {
  int s = excl.cnt;
  XsT T = excl.t;
  int E = lab_prev;
  for (int j = 0; j < RS_ITEMS; ++j) {
    if (mask & (1u << j)) {
      sm.seg_agg[s] = T;
      ++s;
      sm.seg_wc[s] = w[j];
      E = xs_elem_label<RS_ITEMS>(w, sp_thread, lab_end, j);
      sm.seg_e[s] = E;
      T = xs_identity();
    } else {
      T = xs_compose<MB>(T, xs_elem<MB>(w[j], E), E);
    }
  }
}

// RN_q(w) for an even incoming state in binade E (M = 2^E; M = 0: the element is added unrounded)
template <int MB>
__device__ __forceinline__ double rs_round_q(float w, double M) {
  return (MB == 53) ? __dadd_rn(__dadd_rn(M, (double)w), -M) : (double)__fadd_rn(__fadd_rn((float)M, w), -(float)M);
}

// Phases A and B of scan_tile.h for one tile plus the block scan of the contributions.  Identical inputs give identical results
// in describe_kernel and expand_kernel (same code, no atomics), which is what lets the two kernels share the speculation.
template <int MB>
struct RsScan {
  double sp_thread;  // approximate state before this thread
  double M;          // 2^lab_prev (0 in the zero state)
  int lab_prev, lab_end;
  bool simple;       // one binade, no tie: the thread's contribution is the plain exact sum of RN_q(w_j)
  bool tile_simple;  // every thread of the tile is simple
  uint32_t mask;     // special elements (non-simple threads)
  XsSeg excl, total; // exclusive prefix of the contributions / whole tile
};

template <int MB>
__device__ __forceinline__ void rs_tile_scan(const float (&w)[RS_ITEMS], double sp0, RsTileSmem& sm, RsScan<MB>& r, double bias = 1.0) {
  const int tid = threadIdx.x;
  double tsum = 0.0;
  uint32_t key = 0xFFFFFFFFu;  // smallest non-zero weight of the thread (bits - 1; zeros wrap to the top)
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    tsum += (double)w[j];
    key = min(key, __float_as_uint(w[j]) - 1u);
  }
  // `bias` scales the predictor only (speculation): the float32-accumulated sum of torch.multinomial drifts below the fp64 prefix
  double tot_approx;
  const double sp_raw = sp0 + rs_block_excl_scan_d(tsum, sm.dscratch, &tot_approx);
  r.sp_thread = sp_raw * bias;
  const int e0 = xs_label(sp0 * bias);
  r.lab_end = xs_label((sp_raw + tsum) * bias);
  sm.lab_end[tid] = r.lab_end;
  if (tid == RS_NT - 1) sm.sp_end = sp_raw + tsum;
  __syncthreads();
  r.lab_prev = tid ? sm.lab_end[tid - 1] : e0;
  r.M = (r.lab_prev == XS_E_ZERO) ? 0.0 : xs_pow2(r.lab_prev);
  r.mask = 0;
  XsSeg contrib = xs_seg_identity();
  r.simple = (r.lab_prev == r.lab_end);
  if (r.simple && r.lab_prev != XS_E_ZERO) {
    // coarse thread: every non-zero weight is a multiple of the quantum 2^(E-(MB-1)), nothing rounds and tsum is exact
    const int ef = r.lab_prev - (MB == 53 ? 29 : 0) + 127;
    const bool coarse = (key == 0xFFFFFFFFu) || ef <= 0 || (ef < 255 && key + 1u >= ((uint32_t)ef << 23));
    if (coarse) contrib.t.s = tsum;
    else {
      const double hq = xs_pow2(r.lab_prev - MB);
      double acc = 0.0;
      bool tie = false;
#pragma unroll
      for (int j = 0; j < RS_ITEMS; ++j) {
        const double q = rs_round_q<MB>(w[j], r.M);
        tie |= (fabs(__dadd_rn((double)w[j], -q)) == hq);
        acc = __dadd_rn(acc, q);
      }
      contrib.t.s = acc;
      r.simple = !tie;
    }
  }
  if (!r.simple) rs_thread_reduce_slow<MB>(w, r.sp_thread, r.lab_prev, r.lab_end, &r.mask, &contrib);
  r.tile_simple = !__syncthreads_or(r.simple ? 0 : 1);
  int X = 0;
  bool table_ok = true;
  if (r.tile_simple) {
    double tot;
    r.excl.t.s = rs_block_excl_scan_d(contrib.t.s, sm.dscratch, &tot);  // exact: multiples of the quantum below 2^(E+1)
    r.excl.t.d = 0; r.excl.cnt = 0;
    r.total.t.s = tot; r.total.t.d = 0; r.total.cnt = 0;
  } else {
    r.excl = rs_block_excl_scan_seg<MB>(contrib, r.lab_prev, sm.sscratch, sm.lscratch, &r.total);
    X = r.total.cnt;
    table_ok = X < RS_MAXSEG;
    if (table_ok && r.mask) rs_fill_table<MB>(w, r.mask, r.sp_thread, r.lab_prev, r.lab_end, r.excl, sm);
  }
  if (tid == 0) {
    if (table_ok) sm.seg_agg[X] = r.total.t;
    sm.X = X; sm.e0 = e0; sm.table_ok = table_ok ? 1 : 0;
  }
  __syncthreads();
}

// the reference operation itself over one whole tile, by one warp (fallback when a speculation fails): S <- fl_MB(S + w_k);
// c (optional, shared memory) receives fl32 of every state.  The loads run one chunk ahead and the state stays in its own type, so
// the serial chain is one addition per element.
template <int MB>
__device__ __noinline__ double rs_warp_raw_walk(const float* wrow, double S, float* c) {
  const int lane = threadIdx.x & 31;
  float vn = __ldg(wrow + lane);
  float Sf = (float)S;  // MB == 24: the state is a float32
  for (int k0 = 0; k0 < RS_TILE; k0 += 32) {
    const float v = vn;
    if (k0 + 32 < RS_TILE) vn = __ldg(wrow + k0 + 32 + lane);
    float cl = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float x = __shfl_sync(0xffffffffu, v, i);
      if (MB == 24) { Sf = __fadd_rn(Sf, x); if (i == lane) cl = Sf; }
      else { S = __dadd_rn(S, (double)x); if (i == lane) cl = (float)S; }
    }
    if (c) c[k0 + lane] = cl;
  }
  return MB == 24 ? (double)Sf : S;
}

// Tile 0 is where the running sum climbs through a dozen binades, i.e. where all the special elements of a typical column sit.
// But while the sum is small its quantum is tiny: if every non-zero weight of tile 0 is a multiple of the quantum of the highest
// binade the tile can reach (one binade of margin), nothing in the tile rounds and - the tile starting from 0 - its states are the
// exact real prefix sums in any association, like a benign column.  Both kernels evaluate this same predicate from the tile sum
// and the smallest weight normalize_kernel recorded.
template <int MB>
__device__ __forceinline__ bool rs_tile0_exact(const ResampleArgs& a, int col, double* tile_sum) {
  if (MB != 53) return false;
  const double ts = a.tilesum[(int64_t)col * a.tiles_per_col];
  const uint32_t key = a.tilemin[(int64_t)col * a.tiles_per_col];
  *tile_sum = ts;
  if (!(ts >= 0.0 && ts < 2.0)) return false;
  if (key == 0xFFFFFFFFu) return true;  // all zero
  if (key == 0u) return false;          // a negative weight
  const int ef = xs_label(ts) + 1 - 29 + 127;
  return ef <= 0 || (ef < 255 && key >= ((uint32_t)ef << 23));
}

// ---- describe_kernel: one descriptor per tile; the last block of a column chains them -----------------------------------------------
// Chain semantics: S = 0; for every tile in order: sin[tile] = S; S = descriptor(S), every application verifying the speculation
// behind the descriptor (scan_tile.h).  Rounds of RS_NT tiles, three phases per round:
//   P1 (all warps)  descriptors -> integer transducers (scan_tile.h, XiT); each warp runs a segmented scan over its 32 tiles, a
//                   segment being a maximal run of descriptors without special element in one binade (any other descriptor is a
//                   segment of its own), and lists its segments;
//   P2 (warp 0)     walks the segments in order: a handful of integer instructions per run, one genuine IEEE addition per special
//                   element; a descriptor that does not verify is replaced by the reference operation over the raw tile and the
//                   tile is flagged so that expand_kernel takes the same sequential route;
//   P3 (all warps)  state before every tile = state at its segment start advanced by the tile's exclusive transducer.
struct ChainEnt {            // one descriptor in integer form
  XiT ta;                    // the tile's transducer (first one when it has a special element)
  XiT tb;                    // second transducer of a descriptor with one special element; .k = bits of the state for RS_KIND_ABS
  float wc;
  int16_t e0, e1;
  int8_t kind;               // has_special of the descriptor; -1: the conversion to integer form failed
};
struct ChainSmem {
  ChainEnt prim[RS_NT];      // the descriptor speculated from the fp64 prefix
  ChainEnt alt[RS_NT];       // the one speculated from the biased prefix (tiles next to a binade crossing), kind RS_KIND_NONE if absent
  XiT ex[RS_NT];             // exclusive transducer of the tile inside its segment
  uint8_t seg_of[RS_NT];     // local segment index of the tile inside its warp
  XiT seg_t[RS_NT];          // per (warp, local segment): aggregate transducer of a run
  int16_t seg_tile[RS_NT];   // tile (index inside the round) of a single-descriptor segment, -1 for a run
  uint64_t seg_start[RS_NT]; // P2 -> P3: bit pattern of the state at the start of the segment
  int8_t seg_flag[RS_NT];    // 1: the single tile of the segment failed verification; 2: P2 wrote the tiles of the run itself
  int nseg[RS_NT / 32];
};

template <int MB>
__device__ __forceinline__ ChainEnt rs_chain_convert(const XsDesc& d, bool live) {
  ChainEnt c;
  c.ta = xi_identity(); c.tb = xi_identity(); c.wc = d.wc; c.e1 = d.e1;
  int kind = live ? (int)d.has_special : 0;
  const int e0 = live ? (int)d.e0 : XS_E_ZERO;
  if (live && kind <= 1) {
    bool ok = xi_from<MB>(d.a_s, d.a_d, e0, &c.ta);
    if (kind == 1) ok = ok && xi_from<MB>(d.b_s, d.b_d, (int)d.e1, &c.tb);
    if (!ok) kind = -1;
  }
  if (kind == RS_KIND_ABS) c.tb.k = (int64_t)xs_d2u(d.a_s);
  if (kind == RS_KIND_CROSS) { c.ta.k = (int64_t)xs_d2u(d.a_s); c.tb.k = (int64_t)xs_d2u(d.b_s); }
  c.e0 = (int16_t)e0; c.kind = (int8_t)kind;
  return c;
}

// one descriptor that is not part of a run (executed by a whole warp, uniformly); returns false when it does not verify
template <int MB>
__device__ __forceinline__ bool rs_chain_single(const ResampleArgs& a, int col, int tile, const ChainEnt& c, uint64_t S, uint64_t* out) {
  const int lane = threadIdx.x & 31;
  const int kind = c.kind;
  *out = S;
  if (kind == 0) return xi_apply<MB>(S, (int)c.e0, c.ta, out);
  if (kind == 1) {
    uint64_t s1;
    if (!xi_apply<MB>(S, (int)c.e0, c.ta, &s1)) return false;
    const double s2 = xs_add_special<MB>(xs_u2d(s1), c.wc);
    if (xs_label(s2) != (int)c.e1) return false;
    return xi_apply<MB>(xs_d2u(s2), (int)c.e1, c.tb, out);
  }
  if (kind == RS_KIND_ABS) { *out = (uint64_t)c.tb.k; return true; }
  if (kind == RS_KIND_TABLE) {  // walk the tile's segments; 32 table entries are fetched per round trip
    const SegTable* tb = a.tables + (int64_t)col * a.tiles_per_col + tile;
    const int X = (int)c.e1;
    double S2 = xs_u2d(S);
    bool ok = true;
    for (int r0 = 0; r0 <= X && ok; r0 += 32) {
      const int sl = min(r0 + lane, X);
      const double gs = __ldcg(&tb->agg[sl].s);
      const int gd = __ldcg(&tb->agg[sl].d);
      const float gw = __ldcg(&tb->wc[sl]);
      const int ge = __ldcg(&tb->e[sl]);
      const int m = min(32, X + 1 - r0);
      for (int i = 0; i < m && ok; ++i) {
        XsT ts;
        ts.s = __shfl_sync(0xffffffffu, gs, i);
        ts.d = __shfl_sync(0xffffffffu, gd, i);
        const float wc = __shfl_sync(0xffffffffu, gw, i);
        int es = __shfl_sync(0xffffffffu, ge, i);
        if (r0 + i == 0) es = (int)c.e0;
        else { S2 = xs_add_special<MB>(S2, wc); ok = (xs_label(S2) == es); }
        if (ok) ok = xs_apply<MB>(S2, es, ts, &S2);
      }
    }
    *out = xs_d2u(S2);
    return ok;
  }
  if (kind == RS_KIND_CROSS) {
    const int eA = (int)c.e0;
    const double Sd = xs_u2d(S);
    const double totA = xs_u2d((uint64_t)c.ta.k), totB = xs_u2d((uint64_t)c.tb.k);
    const int E = xs_label(Sd);
    XsT t; t.d = 0;
    double o;
    if (E == eA + 1) { t.s = totB; const bool ok = xs_apply<MB>(Sd, eA + 1, t, &o); *out = xs_d2u(o); return ok; }
    if (E != eA) return false;
    const double* cr = reinterpret_cast<const double*>(a.tables + (int64_t)col * a.tiles_per_col + tile);
    const double lim = __dadd_rn(xs_pow2(eA + 1), -Sd);  // exact: the sum leaves the lower binade once the increment reaches this
    int first = 8;  // this lane looks at the threads [8 lane, 8 lane + 8)
    double pa[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) pa[k] = __ldcg(cr + lane * 8 + k);
#pragma unroll
    for (int k = 7; k >= 0; --k) if (pa[k] >= lim) first = k;
    const uint32_t hit = __ballot_sync(0xffffffffu, first < 8);
    if (!hit) { t.s = totA; const bool ok = xs_apply<MB>(Sd, eA, t, &o); *out = xs_d2u(o); return ok; }
    const int hl = __ffs(hit) - 1;
    const int j = hl * 8 + __shfl_sync(0xffffffffu, first, hl);  // the thread whose particles take the sum across the boundary
    const double before = j ? __ldcg(cr + j - 1) : 0.0;
    const double upto = __ldcg(cr + RS_NT + j);
    double S2 = __dadd_rn(Sd, before);  // exact, still in the lower binade
    const float wv = (lane < RS_ITEMS) ? __ldg(a.wn + (int64_t)col * a.ld + (int64_t)tile * RS_TILE + j * RS_ITEMS + lane) : 0.f;
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) S2 = xs_add_special<MB>(S2, __shfl_sync(0xffffffffu, wv, k));
    if (xs_label(S2) != eA + 1) return false;
    t.s = __dadd_rn(totB, -upto);  // exact: both are sums of multiples of the upper quantum
    const bool ok = xs_apply<MB>(S2, eA + 1, t, &o);
    *out = xs_d2u(o);
    return ok;
  }
  return false;  // RS_KIND_RAW, RS_KIND_NONE, failed conversion
}

// diagnostics: why did a descriptor not verify (SMCB_DEBUG_TIMELINE)
template <int MB>
__device__ __noinline__ void rs_chain_why(const ResampleArgs& a, const ChainEnt& p, const ChainEnt& q, uint64_t S) {
  if (!a.dbg || (threadIdx.x & 31)) return;
  int why = 0;                                        // 0: other
  if (p.kind < 0) why = 1;                            // conversion failed
  else if (p.kind >= 2) why = 2;                      // table / raw
  else if ((int)(S >> 52) - 1023 < (int)p.e0) why = 3;  // state still below the speculated binade
  else if ((int)(S >> 52) - 1023 > (int)p.e0) why = 4;  // state already above
  else why = 5;                                       // right binade at the start, leaves it inside (or second part fails)
  atomicAdd((unsigned long long*)&a.dbg[16 + why], 1ull);
  atomicAdd((unsigned long long*)&a.dbg[24 + (q.kind == RS_KIND_NONE ? 0 : 1)], 1ull);
}

template <int MB>
__device__ void rs_chain(const ResampleArgs& a, int col, ChainSmem& cs) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int T = a.tiles_per_col;
  const XsDesc* desc = a.desc + (int64_t)col * T;
  double* sin = a.sin + (int64_t)col * T;
  int32_t* flag = a.tileflag + (int64_t)col * T;
  const float* wcol = a.wn + (int64_t)col * a.ld;
  uint64_t S_carry = 0;  // state carried from round to round (identical in every lane of warp 0, the only warp that uses it)
  long long t_raw = 0, t_p2 = 0, n_raw = 0, n_single = 0, n_run = 0;  // diagnostics (SMCB_DEBUG_TIMELINE)
  const XsDesc* desc2 = a.desc2 + (int64_t)col * T;
  XsDesc dn = {}, dn2 = {};
  if (tid < T) { dn = rs_read_desc(desc + tid); dn2 = rs_read_desc(desc2 + tid); }
  for (int base = 0; base < T; base += RS_NT) {
    const int tile = base + tid;
    const bool live = tile < T;
    const XsDesc d = dn, d2 = dn2;
    if (tile + RS_NT < T) { dn = rs_read_desc(desc + tile + RS_NT); dn2 = rs_read_desc(desc2 + tile + RS_NT); }  // in flight while this round runs
    // ---- P1
    const ChainEnt pe = rs_chain_convert<MB>(d, live);
    ChainEnt ae = rs_chain_convert<MB>(d2, live);
    if (!live || d2.has_special != RS_KIND_CROSS) ae.kind = RS_KIND_NONE;
    const XiT ta = pe.ta;
    const int kind = pe.kind, e0 = pe.e0;
    // a tile with an alternative descriptor sits next to a binade crossing: keep it out of the runs, so that a misplaced crossing
    // costs one descriptor (tried both ways) instead of the whole run around it
    const bool plain = live && kind == 0 && ae.kind == RS_KIND_NONE;
    const int e_left = __shfl_up_sync(0xffffffffu, e0, 1);
    const int plain_left = __shfl_up_sync(0xffffffffu, plain ? 1 : 0, 1);
    const bool head = (lane == 0) || !plain || !plain_left || e_left != e0;
    XiT inc = plain ? ta : xi_identity();
    bool hacc = head;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {  // segmented inclusive scan (Kogge-Stone with head flags)
      XiT t;
      t.k = __shfl_up_sync(0xffffffffu, inc.k, o);
      t.d = __shfl_up_sync(0xffffffffu, inc.d, o);
      const int th = __shfl_up_sync(0xffffffffu, hacc ? 1 : 0, o);
      if (lane >= o && !hacc) { inc = xi_compose<MB>(t, inc); hacc = th != 0; }
    }
    XiT ex;
    ex.k = __shfl_up_sync(0xffffffffu, inc.k, 1);
    ex.d = __shfl_up_sync(0xffffffffu, inc.d, 1);
    if (head) ex = xi_identity();
    const uint32_t heads = __ballot_sync(0xffffffffu, head);
    const int seg = __popc(heads & (0xffffffffu >> (31 - lane))) - 1;
    const bool tail = (lane == 31) || ((heads >> (lane + 1)) & 1u);  // last tile of its segment
    cs.prim[tid] = pe; cs.alt[tid] = ae; cs.ex[tid] = ex; cs.seg_of[tid] = (uint8_t)seg;
    if (tail) { cs.seg_t[wid * 32 + seg] = inc; cs.seg_tile[wid * 32 + seg] = plain ? (int16_t)-1 : (int16_t)tid; }
    if (lane == 0) cs.nseg[wid] = __popc(heads);
    __syncthreads();
    if (a.dbg && tid == 0 && base == 0) a.dbg[4] = rs_now();
    // ---- P2: warp 0 walks the segments; their records are fetched 32 at a time so that only the state itself is a serial chain
    if (wid == 0) {
      const long long t_p2_0 = a.dbg ? rs_now() : 0;
      uint64_t S = S_carry;
      const int nw = min(RS_NT / 32, (T - base + 31) / 32);
      int total = 0;
      for (int w = 0; w < nw; ++w) total += cs.nseg[w];
      for (int f0 = 0; f0 < total; f0 += 32) {
        // flat segment f0 + lane -> (warp, local index)
        int f = f0 + lane, w = 0;
        while (w < nw - 1 && f >= cs.nseg[w]) { f -= cs.nseg[w]; ++w; }
        const bool have = f0 + lane < total;
        const int si_l = have ? w * 32 + f : 0;
        const int st_l = have ? (int)cs.seg_tile[si_l] : -1;
        int t0_l = w * 32;  // first tile of a run
        if (have && st_l < 0) while (cs.seg_of[t0_l] != f) ++t0_l;
        const XiT t_l = cs.seg_t[si_l];
        const int e_l = (st_l < 0) ? (int)cs.prim[t0_l].e0 : 0;
        const int m = min(32, total - f0);
        for (int i = 0; i < m; ++i) {
          const int si = __shfl_sync(0xffffffffu, si_l, i);
          const int st = __shfl_sync(0xffffffffu, st_l, i);
          const int t0 = __shfl_sync(0xffffffffu, t0_l, i);
          XiT t;
          t.k = __shfl_sync(0xffffffffu, t_l.k, i);
          t.d = __shfl_sync(0xffffffffu, t_l.d, i);
          const int E = __shfl_sync(0xffffffffu, e_l, i);
          if (st >= 0 && base + st >= T) break;  // padding behind the last tile
          uint64_t S2 = S;
          int fl = 0;
          if (a.dbg) { if (st < 0) n_run++; else n_single++; }
          if (st < 0) {  // a run of plain descriptors in one binade
            if (!xi_apply<MB>(S, E, t, &S2)) {  // some speculation inside the run is wrong: tile by tile
              fl = 2;
              S2 = S;
              const int wq = t0 >> 5, jq = cs.seg_of[t0];
              for (int q = t0; q < wq * 32 + 32 && cs.seg_of[q] == jq && base + q < T; ++q) {
                uint64_t S3;
                int tf = 0;
                if (!xi_apply<MB>(S2, (int)cs.prim[q].e0, cs.prim[q].ta, &S3) &&
                    !rs_chain_single<MB>(a, col, base + q, cs.alt[q], S2, &S3)) {  // neither speculation holds: the reference operation
                  rs_chain_why<MB>(a, cs.prim[q], cs.alt[q], S2);
                  const long long t0r = a.dbg ? rs_now() : 0;
                  S3 = xs_d2u(rs_warp_raw_walk<MB>(wcol + (int64_t)(base + q) * RS_TILE, xs_u2d(S2), nullptr));
                  if (a.dbg) { t_raw += rs_now() - t0r; n_raw++; }
                  tf = 1;
                  if (lane == 0) atomicAdd(&a.ctrl->slow_tiles, 1);
                }
                if (lane == 0) { sin[base + q] = xs_u2d(S2); flag[base + q] = tf; }
                S2 = S3;
              }
            }
          } else if (!rs_chain_single<MB>(a, col, base + st, cs.prim[st], S, &S2) &&
                     !rs_chain_single<MB>(a, col, base + st, cs.alt[st], S, &S2)) {
            rs_chain_why<MB>(a, cs.prim[st], cs.alt[st], S);
            const long long t0r = a.dbg ? rs_now() : 0;
            S2 = xs_d2u(rs_warp_raw_walk<MB>(wcol + (int64_t)(base + st) * RS_TILE, xs_u2d(S), nullptr));
            if (a.dbg) { t_raw += rs_now() - t0r; n_raw++; }
            fl = 1;
            if (lane == 0) atomicAdd(&a.ctrl->slow_tiles, 1);
          }
          if (lane == 0) { cs.seg_start[si] = S; cs.seg_flag[si] = (int8_t)fl; }
          S = S2;
        }
      }
      S_carry = S;
      if (a.dbg) t_p2 += rs_now() - t_p2_0;
    }
    __syncthreads();

    // ---- P3
    if (live) {
      const int si = wid * 32 + seg;
      const int fl = cs.seg_flag[si];
      if (fl != 2) {
        uint64_t sb;
        xi_apply<MB>(cs.seg_start[si], e0, ex, &sb);
        sin[tile] = xs_u2d(sb);
        flag[tile] = fl;
      }
    }
    __syncthreads();  // the round's tables are reused

  }
  if (a.dbg && tid == 0) { a.dbg[26] = t_raw; a.dbg[27] = t_p2; a.dbg[28] = n_raw; a.dbg[29] = n_single; a.dbg[30] = n_run; }
}

template <int MB>
__global__ void __launch_bounds__(RS_NT, 4) describe_kernel(ResampleArgs a) {
  __shared__ __align__(16) RsTileSmem sm;
  __shared__ __align__(16) ChainSmem cs;
  const int tid = threadIdx.x;
  const int tile = blockIdx.x, col = blockIdx.y;
  const int T = a.tiles_per_col;
  pdl_wait();
  // every independent global load is issued before the first dependent use: a short-lived block cannot afford serialised round trips
  float w[RS_ITEMS];
  {
    const float4* src = reinterpret_cast<const float4*>(a.wn + (int64_t)col * a.ld + (int64_t)tile * RS_TILE + tid * RS_ITEMS);
#pragma unroll
    for (int v = 0; v < RS_ITEMS / 4; ++v) {
      const float4 q = __ldg(src + v);
      w[4 * v] = q.x; w[4 * v + 1] = q.y; w[4 * v + 2] = q.z; w[4 * v + 3] = q.w;
    }
  }
  const double sp0 = a.prefix[(int64_t)col * T + tile];
  const int resample = a.stats ? a.stats[col].resample : 1;
  const int vd = a.verdict[col];
  if (!resample) return;
  if (MB == 53 && (vd & 1)) return;  // benign column: nothing to chain
  long long t_start = 0;
  if (a.dbg && tid == 0) {
    t_start = rs_now();
    atomicMin((unsigned long long*)&a.dbg[0], (unsigned long long)t_start);
    atomicMax((unsigned long long*)&a.dbg[7], (unsigned long long)t_start);
  }
  double ts0 = 0.0;
  const bool exact0 = (tile == 0) && rs_tile0_exact<MB>(a, col, &ts0);  // block-uniform
  RsScan<MB> r;
  if (!exact0) rs_tile_scan<MB>(w, sp0, sm, r);
  if (a.dbg && tid == 0) atomicAdd((unsigned long long*)&a.dbg[3], (unsigned long long)(rs_now() - t_start));
  const int X = exact0 ? 0 : sm.X;
  const bool table_ok = exact0 ? true : (sm.table_ok != 0);
  if (table_ok && X > 1) {  // the segment table goes to global memory for the chain
    SegTable* tb = a.tables + (int64_t)col * T + tile;
    for (int s = tid; s <= X; s += RS_NT) { tb->agg[s] = sm.seg_agg[s]; tb->wc[s] = sm.seg_wc[s]; tb->e[s] = sm.seg_e[s]; }
  }
  if (tid == 0) {
    XsDesc d;
    d.a_s = 0.0; d.b_s = 0.0; d.wc = 0.f; d.e0 = (int16_t)sm.e0; d.e1 = 0; d.a_d = 0; d.b_d = 0; d.pad = 0;
    double S_abs;
    if (exact0) { d.has_special = RS_KIND_ABS; d.a_s = ts0; }
    else if (!table_ok) d.has_special = RS_KIND_RAW;
    else if (tile == 0 && X > 0 && xs_walk_segments<MB>(0.0, sm.e0, X, sm.seg_agg, sm.seg_wc, sm.seg_e, sm.base, &S_abs)) {
      d.has_special = RS_KIND_ABS; d.a_s = S_abs;
    }
    else if (X > 1) { d.has_special = RS_KIND_TABLE; d.e1 = (int16_t)X; }
    else {
      d.a_s = sm.seg_agg[0].s; d.a_d = (int8_t)sm.seg_agg[0].d; d.has_special = (int8_t)X;
      if (X) { d.wc = sm.seg_wc[1]; d.e1 = (int16_t)sm.seg_e[1]; d.b_s = sm.seg_agg[1].s; d.b_d = (int8_t)sm.seg_agg[1].d; }
    }
    a.desc[(int64_t)col * T + tile] = d;
  }
  // A tile next to a binade crossing also gets a CROSSING descriptor.  The sequential sum lags the fp64 prefix (by up to ~1 % for the
  // float32 accumulation of torch.multinomial), so WHICH element crosses cannot be predicted - but only one thread of the tile contains
  // it.  With eA the lower of the two binades: per thread the sum of RN_q(w) under "all in eA" and under "all in eA + 1", and their
  // inclusive scans over the threads (2 x 256 doubles, kept in the tile's table slot).  The chain then finds the crossing thread from
  // the EXACT state with one comparison per thread, walks that thread's 16 weights with genuine additions and adds the upper-binade
  // sums of the threads behind it (rs_chain_single, RS_KIND_CROSS).
  {
    constexpr double kBias = (MB == 24) ? 0.98 : (1.0 - 1e-9);
    __shared__ int need_alt;
    if (tid == 0) {
      const double spe = sm.sp_end;
      const int eA = xs_label(sp0 * kBias);
      need_alt = (!exact0 && tile > 0 && table_ok && X <= 1 && eA != XS_E_ZERO && xs_label(spe) <= eA + 1 &&
                  (eA != xs_label(sp0) || xs_label(spe * kBias) != xs_label(spe))) ? 1 : 0;
    }
    __syncthreads();
    XsDesc d2;
    memset(&d2, 0, sizeof(d2));
    d2.has_special = RS_KIND_NONE;
    if (need_alt) {
      const int eA = xs_label(sp0 * kBias);
      const double MA = xs_pow2(eA), MBv = xs_pow2(eA + 1), hqA = xs_pow2(eA - MB), hqB = xs_pow2(eA + 1 - MB);
      double accA = 0.0, accB = 0.0;
      bool tie = false;
#pragma unroll
      for (int k = 0; k < RS_ITEMS; ++k) {
        const double qa = rs_round_q<MB>(w[k], MA), qb = rs_round_q<MB>(w[k], MBv);
        tie |= (fabs(__dadd_rn((double)w[k], -qa)) == hqA) || (fabs(__dadd_rn((double)w[k], -qb)) == hqB);
        accA = __dadd_rn(accA, qa);
        accB = __dadd_rn(accB, qb);
      }
      const int any_tie = __syncthreads_or(tie ? 1 : 0);
      double totA, totB;
      const double exA = rs_block_excl_scan_d(accA, sm.dscratch, &totA);
      const double exB = rs_block_excl_scan_d(accB, sm.dscratch, &totB);
      if (!any_tie) {
        double* cr = reinterpret_cast<double*>(a.tables + (int64_t)col * T + tile);
        cr[tid] = exA + accA;            // inclusive over the threads, lower binade
        cr[RS_NT + tid] = exB + accB;    // ... upper binade
        d2.has_special = RS_KIND_CROSS; d2.e0 = (int16_t)eA; d2.a_s = totA; d2.b_s = totB;
        __threadfence();
      }
    }
    if (tid == 0) a.desc2[(int64_t)col * T + tile] = d2;
  }
  if (tid == 0 || (table_ok && X > 1)) __threadfence();  // table and descriptor stores precede the ticket
  __syncthreads();
  if (tid == 0) {
    sm.is_last = (atomicAdd(&a.dcounter[col], 1) == T - 1);
    if (a.dbg) {
      const long long dt = rs_now() - t_start;
      atomicAdd((unsigned long long*)&a.dbg[2], (unsigned long long)dt);
      atomicMax((unsigned long long*)&a.dbg[5], ((unsigned long long)dt << 20) | (unsigned long long)tile);
    }
  }
  __syncthreads();
  if (!sm.is_last) return;
  __threadfence();
  if (a.dbg && tid == 0) a.dbg[1] = rs_now();
  rs_chain<MB>(a, col, cs);
  if (tid == 0) a.dcounter[col] = 0;
  if (a.dbg && tid == 0) a.dbg[6] = rs_now();
}

// ---- probes at or below a cumulative weight ------------------------------------------------------------------------------------------
__device__ __noinline__ int32_t rs_count_slow(float c, float u, int32_t n, float nf) { return xs_count_le_t<int32_t>(c, u, n, nf); }

// general form: any n < 2^31, any c, any u
__device__ __forceinline__ int32_t rs_count_any(float c, float u, int32_t n, float nf, double nd, double nfd, bool fast_ok) {
  const uint32_t cb = __float_as_uint(c);
  if (fast_ok && (cb == 0u || (cb >= 0x00800000u && cb < 0x7F800000u))) return xs_count_fast(c, u, n, nd, nfd);
  return rs_count_slow(c, u, n, nf);
}

// ---- expand_kernel -------------------------------------------------------------------------------------------------------------------
#define FB_ROWS 5                              // int4 rows per warp and window
#define FB_WIN (RS_NT / 32 * FB_ROWS * 32 * 4) // 5120 output slots per window
struct ExpandSmem {
  int32_t stage[FB_WIN];      // global index of the particle whose FIRST offspring sits in this slot, -1 elsewhere
  float c_tile[RS_TILE];      // cumulative weights of the tile (general tiles only)
  RsTileSmem core;
  int32_t wtot[RS_NT / 32];
  int32_t carry, n_out;
};

// Counts of this thread's particles; every particle with offspring marks its first slot inside the window [wb, wb + FB_WIN).
// FAST: cumulative weights on the fly, c_j = fl32(S0 + sum_{i<=j} RN_q(w_i)); otherwise they are read from shared memory.
template <int MB, bool FAST, bool BENIGN, int ITEMS = RS_ITEMS, int WIN = FB_WIN, typename SM = ExpandSmem>
__device__ __forceinline__ int32_t rs_mark_pass(const float (&w)[ITEMS], double S0, double M, const float* c_thread, int32_t lo,
                                                int32_t gbase, int32_t wb, bool first, float u, int32_t n, float nf, double nd,
                                                double nfd, bool fast_ok, SM& sm) {
  double run = S0;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    float c;
    if (FAST) {
      run = __dadd_rn(run, BENIGN ? (double)w[j] : rs_round_q<MB>(w[j], M));
      c = (float)run;
    } else c = c_thread[j];
    int32_t hi = BENIGN ? xs_count_fast(c, u, n, nd, nfd) : rs_count_any(c, u, n, nf, nd, nfd, fast_ok);
    if (gbase + ITEMS >= n) hi = (gbase + j >= n - 1) ? n : hi;  // cumsum[..., -1] = 1.0 (resampling.py:49): every probe is <= 1
    const int32_t r = lo - wb;
    if (hi > lo && (uint32_t)r < (uint32_t)WIN) sm.stage[r] = gbase + j;
    if (!first && hi > lo && r < 0 && hi > wb) sm.carry = gbase + j;  // its slots began in an earlier window (one such particle at most)
    lo = max(lo, hi);
  }
  return lo;
}

// ---- ancestors of one tile from the marks: the tile owns the output slots [n_in, n_out) ----------------------------------------------
// `mark(wb, first)` runs the mark pass for the window starting at slot wb and returns the thread's last count.
template <int NT = RS_NT, int ROWS = FB_ROWS, typename SM = ExpandSmem, typename Mark>
__device__ __forceinline__ void rs_emit_ancestors(SM& sm, Mark mark, int32_t n_in, int32_t* anc) {
  constexpr int WIN = NT / 32 * ROWS * 32 * 4;  // output slots per window (FB_WIN for the defaults)
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int32_t wb0 = n_in & ~3;
  {
    const int32_t last = mark(wb0, true);
    if (tid == NT - 1) sm.n_out = last;
  }
  __syncthreads();
  const int32_t n_out = sm.n_out;

  int32_t carry = -1;
  for (int32_t wb = wb0; wb < n_out; wb += WIN) {
    if (wb != wb0) {  // rare: more than one window of offspring
      __syncthreads();
#pragma unroll
      for (int k = 0; k < ROWS; ++k) *reinterpret_cast<int4*>(&sm.stage[(k * NT + tid) * 4]) = make_int4(-1, -1, -1, -1);
      if (tid == 0) sm.carry = -1;
      __syncthreads();
      mark(wb, false);
      __syncthreads();
      carry = max(carry, sm.carry);
    }
    const int32_t wlen = min(WIN, n_out - wb);
    // last mark at or before every slot; warp `wid` owns ROWS rows of 32 int4, row k = int4 [(wid*ROWS + k)*32, +32)
    int4 m[ROWS];
    int32_t cin[ROWS];
    int32_t wrun = -1;  // last mark seen by this warp so far
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
      const int i4 = (wid * ROWS + k) * 32 + lane;
      m[k] = *reinterpret_cast<const int4*>(&sm.stage[i4 * 4]);
      const int32_t v = max(max(m[k].x, m[k].y), max(m[k].z, m[k].w));
      const uint32_t bal = __ballot_sync(0xffffffffu, v >= 0);
      const uint32_t before = bal & ((1u << lane) - 1u);
      const int32_t vb = __shfl_sync(0xffffffffu, v, before ? 31 - __clz(before) : 0);
      cin[k] = before ? vb : wrun;
      const int32_t vl = __shfl_sync(0xffffffffu, v, bal ? 31 - __clz(bal) : 0);
      if (bal) wrun = vl;
    }
    if (lane == 0) sm.wtot[wid] = wrun;
    __syncthreads();
    // last mark before this warp's rows / in the whole window: one lane per warp, integer redux
    const int32_t wt = (lane < NT / 32) ? sm.wtot[lane] : -1;
    const int32_t cw = max(carry, __reduce_max_sync(0xffffffffu, (lane < wid) ? wt : -1));
    carry = max(carry, __reduce_max_sync(0xffffffffu, wt));
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
      const int i4 = (wid * ROWS + k) * 32 + lane;
      const int32_t s0 = wb + i4 * 4;
      if (i4 * 4 < wlen) {
        int4 o;
        o.x = max(max(cw, cin[k]), m[k].x);
        o.y = max(o.x, m[k].y);
        o.z = max(o.y, m[k].z);
        o.w = max(o.z, m[k].w);
        if (s0 >= n_in && s0 + 4 <= n_out) *reinterpret_cast<int4*>(anc + s0) = o;
        else {
          if (s0 >= n_in && s0 < n_out) anc[s0] = o.x;
          if (s0 + 1 >= n_in && s0 + 1 < n_out) anc[s0 + 1] = o.y;
          if (s0 + 2 >= n_in && s0 + 2 < n_out) anc[s0 + 2] = o.z;
          if (s0 + 3 >= n_in && s0 + 3 < n_out) anc[s0 + 3] = o.w;
        }
      }
    }
  }
}

// "last mark at or before every slot" over a window of NT * PER slots in shared memory (sm.stage); the ancestors replace the marks
// in place.  A thread owns PER consecutive slots: running maximum in registers, one warp scan of the per-thread maxima, one cross-warp
// step (sm.wtot: NT / 32 ints).  `carry` = last mark of the windows before this one; returns it updated.  All NT threads call.
template <int NT, int PER, typename SM>
__device__ __forceinline__ int32_t rs_emit_blocked(SM& sm, int32_t carry) {
  static_assert(PER % 4 == 0, "a thread reads whole int4");
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int4 m[PER / 4];
#pragma unroll
  for (int k = 0; k < PER / 4; ++k) m[k] = *reinterpret_cast<const int4*>(&sm.stage[tid * PER + 4 * k]);
  int32_t v = -1;
#pragma unroll
  for (int k = 0; k < PER / 4; ++k) v = max(max(v, max(m[k].x, m[k].y)), max(m[k].z, m[k].w));
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v = max(v, t);
  }
  int32_t run = __shfl_up_sync(0xffffffffu, v, 1);
  if (lane == 0) run = -1;
  if (lane == 31) sm.wtot[wid] = v;
  __syncthreads();
  const int32_t wt = (lane < NT / 32) ? sm.wtot[lane] : -1;
  run = max(run, max(carry, __reduce_max_sync(0xffffffffu, (lane < wid) ? wt : -1)));
  carry = max(carry, __reduce_max_sync(0xffffffffu, wt));
#pragma unroll
  for (int k = 0; k < PER / 4; ++k) {
    m[k].x = run = max(run, m[k].x);
    m[k].y = run = max(run, m[k].y);
    m[k].z = run = max(run, m[k].z);
    m[k].w = run = max(run, m[k].w);
    *reinterpret_cast<int4*>(&sm.stage[tid * PER + 4 * k]) = m[k];
  }
  __syncthreads();
  return carry;
}

template <int MB, int OUT>
__global__ void __launch_bounds__(RS_NT, 4) expand_kernel(ResampleArgs a) {
  __shared__ __align__(16) ExpandSmem sm;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int tile = blockIdx.x, col = blockIdx.y;
  const int T = a.tiles_per_col;
  pdl_wait();
  // every independent global load is issued before the first dependent use
  float w[RS_ITEMS];
  {
    const float4* src = reinterpret_cast<const float4*>(a.wn + (int64_t)col * a.ld + (int64_t)tile * RS_TILE + tid * RS_ITEMS);
#pragma unroll
    for (int v = 0; v < RS_ITEMS / 4; ++v) {
      const float4 q = __ldg(src + v);
      w[4 * v] = q.x; w[4 * v + 1] = q.y; w[4 * v + 2] = q.z; w[4 * v + 3] = q.w;
    }
  }
  const double sp0 = a.prefix[(int64_t)col * T + tile];
  const double sin_t = a.sin[(int64_t)col * T + tile];
  const int resample = a.stats ? a.stats[col].resample : 1;
  const int vd = (OUT == RS_OUT_ANCESTORS) ? a.verdict[col] : 0;
  const float u = (OUT == RS_OUT_ANCESTORS) ? a.u_col[col] : 0.f;
  if (OUT == RS_OUT_ANCESTORS) {  // clear the first window while the loads are in flight
#pragma unroll
    for (int k = 0; k < FB_ROWS; ++k) *reinterpret_cast<int4*>(&sm.stage[(k * RS_NT + tid) * 4]) = make_int4(-1, -1, -1, -1);
    if (tid == 0) sm.carry = -1;
  }
  if (!resample) return;
  const bool benign = (MB == 53) && ((vd & 1) || a.force_benign);
  const bool fast_ok = (vd & 2) != 0;
  const int32_t n = (int32_t)a.n;
  const float nf = (float)a.n;
  const double nd = (double)n, nfd = (double)nf;

  // ---- exact state before the tile (S_in) and before this thread (S0)
  double S_in, S0, M = 0.0;
  bool fast = true;  // block-uniform: cumulative weights on the fly from (S0, M); else through sm.c_tile
  if (benign) {
    double tsum = 0.0, tot;
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) tsum += (double)w[j];
    S_in = sp0;
    S0 = S_in + rs_block_excl_scan_d(tsum, sm.core.dscratch, &tot);  // exact
  } else if (tile == 0 && rs_tile0_exact<MB>(a, col, &S_in)) {  // nothing rounds in the first tile: exact sums in any order
    double tsum = 0.0, tot;
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) tsum += (double)w[j];
    S_in = 0.0;
    S0 = rs_block_excl_scan_d(tsum, sm.core.dscratch, &tot);
  } else {
    // the exact state before the tile is the predictor here (better than the fp64 prefix the descriptors were speculated from),
    // so the speculation is verified again: by the apply below for a simple tile, by the segment walk otherwise
    S_in = sin_t;
    bool slow = false;
    RsScan<MB> r;
    S0 = S_in;
    {
      rs_tile_scan<MB>(w, S_in, sm.core, r);
      M = r.M;
      if (r.tile_simple) {
        if (tid == 0) { double o; sm.core.ok = xs_apply<MB>(S_in, sm.core.e0, r.total.t, &o) ? 1 : 0; }
        __syncthreads();
        if (!sm.core.ok) slow = true;
        else S0 = __dadd_rn(S_in, r.excl.t.s);  // exact: multiples of the quantum inside one binade
      } else {
        fast = false;
        if (tid == 0) {
          double S_out;
          sm.core.ok = (sm.core.table_ok && xs_walk_segments<MB>(S_in, sm.core.e0, sm.core.X, sm.core.seg_agg, sm.core.seg_wc,
                                                                 sm.core.seg_e, sm.core.base, &S_out)) ? 1 : 0;
        }
        __syncthreads();
        if (!sm.core.ok) slow = true;  // cannot happen after a verified chain; stay safe
        else if (r.simple) {
          const double b = sm.core.base[r.excl.cnt];
          double Ss = b;
          if (r.lab_prev != XS_E_ZERO) xs_apply<MB>(b, r.lab_prev, r.excl.t, &Ss);
          double acc = Ss;
#pragma unroll
          for (int j = 0; j < RS_ITEMS; ++j) {
            acc = __dadd_rn(acc, rs_round_q<MB>(w[j], r.M));
            sm.c_tile[tid * RS_ITEMS + j] = (float)acc;
          }
        } else rs_thread_finalize_slow<MB>(w, r.mask, r.lab_prev, r.excl.cnt, r.excl.t, sm.core.base, sm.core.seg_e,
                                           &sm.c_tile[tid * RS_ITEMS]);
      }
    }
    if (slow) {  // the reference operation itself, sequentially
      fast = false;
      __syncthreads();
      if (tid < 32) rs_warp_raw_walk<MB>(a.wn + (int64_t)col * a.ld + (int64_t)tile * RS_TILE, S_in, sm.c_tile);
    }
    if (!fast) __syncthreads();
  }

  const int32_t gbase = tile * RS_TILE + tid * RS_ITEMS;  // global index of this thread's first particle
  if (OUT == RS_OUT_CUMSUM) {  // torch.multinomial's prefix sums: the search over them happens in multinomial_draw_kernel
    float c[RS_ITEMS];
    double run = S0;
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
      if (fast) { run = __dadd_rn(run, rs_round_q<MB>(w[j], M)); c[j] = (float)run; }
      else c[j] = sm.c_tile[tid * RS_ITEMS + j];
    }
    float* dst = a.c_out + (int64_t)col * a.ld + gbase;
#pragma unroll
    for (int v = 0; v < RS_ITEMS / 4; ++v)
      reinterpret_cast<float4*>(dst)[v] = make_float4(c[4 * v], c[4 * v + 1], c[4 * v + 2], c[4 * v + 3]);
    return;
  }

  // ---- expansion: the tile owns the output slots [n_in, n_out)
  const float c_in = (float)S_in;  // cumulative weight just before the tile
  const float c_prev = fast ? (float)S0 : (tid ? sm.c_tile[tid * RS_ITEMS - 1] : c_in);
  const float* c_thread = &sm.c_tile[tid * RS_ITEMS];
  const int32_t n_in = (tile == 0) ? 0 : ((tile * RS_TILE - 1 >= n - 1) ? n : rs_count_any(c_in, u, n, nf, nd, nfd, fast_ok));
  int32_t lo_thread = n_in;
  if (tid) lo_thread = max(n_in, (gbase - 1 >= n - 1) ? n : rs_count_any(c_prev, u, n, nf, nd, nfd, fast_ok));
  auto mark = [&](int32_t wb, bool first) -> int32_t {
    if (benign) return rs_mark_pass<MB, true, true>(w, S0, 0.0, c_thread, lo_thread, gbase, wb, first, u, n, nf, nd, nfd, true, sm);
    if (fast) return rs_mark_pass<MB, true, false>(w, S0, M, c_thread, lo_thread, gbase, wb, first, u, n, nf, nd, nfd, fast_ok, sm);
    return rs_mark_pass<MB, false, false>(w, S0, M, c_thread, lo_thread, gbase, wb, first, u, n, nf, nd, nfd, fast_ok, sm);
  };
  rs_emit_ancestors(sm, mark, n_in, a.anc + (int64_t)col * a.ld);
}

// ---- resample_fused_kernel: normalise + prefix + expand in ONE pass (weights rounded to multiples of 2^-52) ---------------------
// With rounding-free weights the state before a tile is the plain sum of the preceding tile sums, in any order.  So one kernel can do
// everything: weights from the log-weights (registers only - they are never written), tile sum published with the launch epoch as
// tag, then ALL threads of the CTA poll the preceding tiles' slots in parallel (one 128-bit load per predecessor and round trip; tile
// ids are handed out in start order, so every predecessor is running or finished and nobody can wait on a tile that has not started),
// sum them exactly, and expand as expand_kernel does.  No normalised weights in memory, no serial tail, one launch per resampling.
__global__ void __launch_bounds__(RS_NT, 4) resample_fused_kernel(ResampleArgs a) {
  __shared__ __align__(16) ExpandSmem sm;
  const int tid = threadIdx.x;
  // tile id = block id: blocks of a 1-D grid are dispatched in increasing order, so every predecessor a tile polls is resident or
  // finished (the assumption CUB's decoupled look-back makes too) - no atomic ticket, hence no global round trip before the loads
  const int T = a.tiles_per_col;
  const int col = (int)blockIdx.x / T, tile = (int)blockIdx.x % T;
#pragma unroll
  for (int k = 0; k < FB_ROWS; ++k) *reinterpret_cast<int4*>(&sm.stage[(k * RS_NT + tid) * 4]) = make_int4(-1, -1, -1, -1);
  if (tid == 0) sm.carry = -1;  // (ordered before the marks by the barriers of the block scan below)
  pdl_wait();
  // independent loads first
  float w[RS_ITEMS];
  {
    const float4* src = reinterpret_cast<const float4*>(a.w + (int64_t)col * a.ld + (int64_t)tile * RS_TILE + tid * RS_ITEMS);
#pragma unroll
    for (int v = 0; v < RS_ITEMS / 4; ++v) {
      const float4 q = __ldg(src + v);
      w[4 * v] = q.x; w[4 * v + 1] = q.y; w[4 * v + 2] = q.z; w[4 * v + 3] = q.w;
    }
  }
  const ColStats st = a.stats[col];
  const unsigned long long epoch = a.epoch_host;
  const int t_now = a.t_host;
  if (!st.resample) return;
  const float m = a.use_rw ? st.m_rw : st.m_lw, iz = a.use_rw ? st.inv_z_rw : st.inv_z_lw;
  const int32_t n = (int32_t)a.n;
  const double nd = (double)n, nfd = (double)(float)a.n;
  const int32_t gbase = tile * RS_TILE + tid * RS_ITEMS;
  // weights, rounded to multiples of 2^-52 (exactly what normalize_kernel computes with `quantize`)
  double tsum = 0.0;
  const bool inner = gbase + RS_ITEMS <= n;  // no padding among this thread's particles
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    float x = smcb_weight(w[j], m, iz);
    if (!inner && gbase + j >= n) x = 0.f;
    const double xd = __dadd_rn(__dadd_rn(1.0, (double)x), -1.0);
    w[j] = (float)xd;
    tsum += xd;
  }
  if (a.w_out) {
    float* dst = a.w_out + (int64_t)col * a.ld + gbase;
#pragma unroll
    for (int v = 0; v < RS_ITEMS / 4; ++v)
      reinterpret_cast<float4*>(dst)[v] = make_float4(w[4 * v], w[4 * v + 1], w[4 * v + 2], w[4 * v + 3]);
  }
  double tot;
  const double ex = rs_block_excl_scan_d(tsum, sm.core.dscratch, &tot);
  FusedSlot* slots = a.fslots + (int64_t)col * T;
  if (tid == 0) {  // publish: one 128-bit store carries the sum and its tag
    double2 v;
    v.x = tot; v.y = __longlong_as_double((long long)epoch);
    __stcg(reinterpret_cast<double2*>(slots + tile), v);
  }
  // exact sum of the preceding tiles: every thread polls its share of the predecessors
  double part = 0.0;
  for (int q = tile - 1 - tid; q >= 0; q -= RS_NT) {
    double2 v;
    for (;;) {
      v = __ldcg(reinterpret_cast<const double2*>(slots + q));
      if ((unsigned long long)__double_as_longlong(v.y) == epoch) break;
      __nanosleep(64);
    }
    part += v.x;
  }
  const double S_in = block_allreduce<RS_NT>(part, 0.0, OpSumD(), sm.core.dscratch);
  const double S0 = S_in + ex;
  // one uniform per column (resampling.py:41)
  const Philox4 r4 = philox4x32_10((uint32_t)(col + a.col0), 0u, (uint32_t)t_now, SMCB_RNG_SYSTEMATIC, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
  const float u = smcb_u01(r4.x);
  if (a.u_out && tile == 0 && tid == 0) a.u_out[col] = u;
  const float nf = (float)a.n;
  const int32_t n_in = (tile == 0) ? 0 : xs_count_fast((float)S_in, u, n, nd, nfd);
  int32_t lo_thread = n_in;
  if (tid) lo_thread = max(n_in, (gbase - 1 >= n - 1) ? n : xs_count_fast((float)S0, u, n, nd, nfd));
  auto mark = [&](int32_t wb, bool first) -> int32_t {
    return rs_mark_pass<53, true, true>(w, S0, 0.0, nullptr, lo_thread, gbase, wb, first, u, n, nf, nd, nfd, true, sm);
  };
  rs_emit_ancestors(sm, mark, n_in, a.anc + (int64_t)col * a.ld);
}

