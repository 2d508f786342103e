// Systematic resampling on the device, bit-exact against the reference's CPU path
//     probs = (arange(n) + u) / n ; cumsum = W.cumsum(-1) ; cumsum[-1] = 1 ; searchsorted(cumsum, probs)     (resampling.py:44-50)
//
//   tile_sum_kernel   fp64 sum of every tile of normalised weights (approximate prefix -> binade labels)
//   systematic_kernel one pass per tile: exact transducer scan (scan_tile.h) chained across tiles by a decoupled look-back on
//                     exact states, then the ancestors are produced by EXPANSION: particle j owns the probes
//                     [count(c_{j-1}), count(c_j)), staged through shared memory and written with coalesced stores.
// Layout: one column (independent filter) is a contiguous row of `ld` floats; tiles never straddle columns.
#pragma once
#include "common.cuh"
#include "scan_tile.h"
#include "philox.h"

#define RS_NT 256
#define RS_ITEMS 16
#define RS_TILE (RS_NT * RS_ITEMS)
#define RS_MAXSEG 192
#define RS_LBSTACK 64
#define RS_BIGLIST 32
#define RS_BIG 96

struct __align__(16) TileSlot {
  XsDesc desc;       // 32 B
  double incl;       // exact state after the tile
  uint32_t status;   // (epoch << 2) | {0 none, 1 descriptor, 2 inclusive, 3 opaque}
  uint32_t pad;
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
  double* tilesum;         // (B, tiles_per_col)
  TileSlot* slots;         // (B, tiles_per_col)
  int32_t* anc;            // (B, ld) ancestors out
  float* w_out;            // optional dump of the normalised weights used (B, ld)
  float* c_out;            // OUT_CUMSUM: the emulated sequential prefix sums (B, ld)
  long long* dbg;          // optional per-tile timeline (8 x int64 globaltimer stamps per tile), diagnostics only
  int32_t approx;          // 1: skip the exact chaining (incoming state := fp64 approximate prefix); NOT bit-exact, diagnostics only
  Ctrl* ctrl;
  uint32_t* tilemin;       // (B, tiles_per_col) smallest non-zero weight of every tile (float bits; 0 when a weight is negative)
  int32_t* ncounter;       // (B) last-block-done tickets of normalize_kernel (self resetting)
  int32_t* verdict;        // (B) bit 0: the column is "benign" - its sequential fp64 prefix sum never rounds (benign kernel)
  float* u_col;            // (B) the systematic offset of every column for this launch (injected or Philox)
};
enum { RS_OUT_ANCESTORS = 0, RS_OUT_CUMSUM = 1 };

// ---- pre-pass: normalised weights (written once, zero beyond n), their fp64 tile sums and the column verdict ---------------------
// A column is BENIGN when every non-zero weight is a multiple of 2^-52 (w >= 2^-29) and the weights sum to less than 2: then
// every partial sum is a multiple of 2^-52 below 2, i.e. exactly representable, so the reference's sequential fp64 prefix sum
// never rounds and equals the exact real prefix sum in ANY association.  Such columns need no binade labels, no transducers
// and no chaining across tiles (systematic_benign_kernel); all others take the general exact scan (systematic_kernel).
struct OpMinU { __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a < b ? a : b; } };
#define RS_BENIGN_MIN_BITS 0x31000000u  // 2^-29

__global__ void __launch_bounds__(RS_NT) normalize_kernel(ResampleArgs a) {
  __shared__ double scratch[33];
  __shared__ uint32_t uscratch[33];
  __shared__ int is_last;
  const int col = blockIdx.y, tile = blockIdx.x;
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {  // arm the scan kernel that follows in stream order
    a.ctrl->tile_counter = 0;
    a.ctrl->epoch += 1;
  }
  if (a.stats && !a.stats[col].resample) return;
  const int64_t off = (int64_t)col * a.ld + (int64_t)tile * RS_TILE;
  const int64_t g0 = (int64_t)tile * RS_TILE;
  float m = 0.f, iz = 1.f;
  if (!a.input_is_w) {
    const ColStats& st = a.stats[col];
    m = a.use_rw ? st.m_rw : st.m_lw;
    iz = a.use_rw ? st.inv_z_rw : st.inv_z_lw;
  }
  double s = 0.0;
  uint32_t key = 0xFFFFFFFFu;
#pragma unroll
  for (int v = 0; v < RS_ITEMS / 4; ++v) {  // striped float4: fully coalesced
    const int e = (v * RS_NT + threadIdx.x) * 4;
    float4 q = __ldg(reinterpret_cast<const float4*>(a.w + off + e));
    float x[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!a.input_is_w) x[k] = smcb_weight(smcb_sanitize(x[k]), m, iz);
      if (g0 + e + k >= a.n) x[k] = 0.f;
      s += (double)x[k];
      const uint32_t b = __float_as_uint(x[k]);
      const uint32_t kk = (b == 0u) ? 0xFFFFFFFFu : ((b >> 31) ? 0u : b);
      key = min(key, kk);
    }
    if (!a.input_is_w || a.wn != a.w) *reinterpret_cast<float4*>(a.wn + off + e) = make_float4(x[0], x[1], x[2], x[3]);
    if (a.w_out) *reinterpret_cast<float4*>(a.w_out + off + e) = make_float4(x[0], x[1], x[2], x[3]);
  }
  s = block_allreduce<RS_NT>(s, 0.0, OpSumD(), scratch);
  key = block_allreduce<RS_NT>(key, 0xFFFFFFFFu, OpMinU(), uscratch);
  if (threadIdx.x == 0) {
    a.tilesum[(int64_t)col * a.tiles_per_col + tile] = s;
    a.tilemin[(int64_t)col * a.tiles_per_col + tile] = key;
    __threadfence();
    is_last = (atomicAdd(&a.ncounter[col], 1) == a.tiles_per_col - 1);
  }
  __syncthreads();
  if (!is_last) return;
  // ---- the block that completes a column settles its systematic offset and the verdict
  __threadfence();
  double tot = 0.0;
  uint32_t mk = 0xFFFFFFFFu;
  for (int q = threadIdx.x; q < a.tiles_per_col; q += RS_NT) {
    tot += __ldcg(a.tilesum + (int64_t)col * a.tiles_per_col + q);
    mk = min(mk, __ldcg(a.tilemin + (int64_t)col * a.tiles_per_col + q));
  }
  tot = block_allreduce<RS_NT>(tot, 0.0, OpSumD(), scratch);
  mk = block_allreduce<RS_NT>(mk, 0xFFFFFFFFu, OpMinU(), uscratch);
  if (threadIdx.x == 0) {
    float u;  // one uniform per column (resampling.py:41)
    if (a.u_in) u = a.u_in[col];
    else {
      Philox4 r = philox4x32_10((uint32_t)col, 0u, (uint32_t)a.ctrl->t, SMCB_RNG_SYSTEMATIC, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
      u = smcb_u01(r.x);
    }
    a.u_col[col] = u;
    if (a.u_out) a.u_out[col] = u;
    const bool u_ok = (u == 0.f) || (u >= 5.5e-20f && u < 1.0f);
    const bool benign = !a.approx && mk >= RS_BENIGN_MIN_BITS && tot < 1.5 && a.n <= (1 << 23) && u_ok;
    a.verdict[col] = benign ? 1 : 0;
    a.ncounter[col] = 0;
  }
}

// ---- block scans (the RS_NT compute threads synchronise on named barrier 1; the look-back warp is not involved) ------------
__device__ __forceinline__ void rs_cbar() { asm volatile("bar.sync 1, %0;" ::"n"(RS_NT) : "memory"); }

__device__ __forceinline__ double rs_block_sum(double v, double* scratch /*>=33*/) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  rs_cbar();
  if (lane == 0) scratch[wid] = v;
  rs_cbar();
  double r = 0.0;
#pragma unroll
  for (int k = 0; k < RS_NT / 32; ++k) r += scratch[k];
  return r;
}

__device__ __forceinline__ double rs_block_excl_scan(double v, double* scratch /*>=33*/) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  rs_cbar();
  if (lane == 31) scratch[wid] = inc;
  rs_cbar();
  double off = 0.0;
#pragma unroll
  for (int k = 0; k < RS_NT / 32; ++k) off += (k < wid) ? scratch[k] : 0.0;
  return off + (inc - v);
}

// exclusive max-scan of one int per thread
__device__ __forceinline__ int rs_block_excl_maxscan(int v, int* scratch /*>=8*/) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc = max(inc, t);
  }
  int excl = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane == 0) excl = 0;
  rs_cbar();
  if (lane == 31) scratch[wid] = inc;
  rs_cbar();
  int off = 0;
#pragma unroll
  for (int k = 0; k < RS_NT / 32; ++k) off = max(off, (k < wid) ? scratch[k] : 0);
  return max(off, excl);
}

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

// ---- decoupled look-back on exact states (warp 0) ---------------------------------------------------------------------------
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

__device__ __forceinline__ XsDesc rs_read_desc(const TileSlot* s) {
  XsDesc d;
  const double2 q0 = __ldcg(reinterpret_cast<const double2*>(&s->desc));
  const int4 q1 = __ldcg(reinterpret_cast<const int4*>(&s->desc) + 1);
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

__device__ __forceinline__ uint32_t rs_wait_state(const TileSlot* s, uint32_t epoch, uint32_t want_mask) {
  // spin until the slot carries this launch's epoch and a state whose bit is set in want_mask
  for (;;) {
    uint32_t st = ld_acquire_u32(&s->status);
    if ((st >> 2) == epoch && ((want_mask >> (st & 3u)) & 1u)) return st & 3u;
    __nanosleep(20);
  }
}

__device__ __forceinline__ XsT rs_shfl_down_t(const XsT& t, int o) {
  XsT r;
  r.s = __shfl_down_sync(0xffffffffu, t.s, o);
  r.d = __shfl_down_sync(0xffffffffu, t.d, o);
  return r;
}

// Exact incoming state of `tile`: walk back over the predecessors' descriptors (32 per round) until a tile with a published
// inclusive state is found, composing descriptors on the way.  Runs of descriptors without a special element in one binade
// compose associatively (xs_compose), so a whole window folds in five shuffle steps; descriptors with a special element are
// kept individually.  The collected entries are then applied forward to the exact state, each application verifying the
// speculation behind it; on any failure the tile simply waits for its direct predecessor's inclusive state.
template <int MB>
__device__ double rs_lookback(const TileSlot* slots, int tile, uint32_t epoch, XsDesc* stack /*RS_LBSTACK*/, Ctrl* ctrl) {
  const int lane = threadIdx.x & 31;
  int win_base = tile - 1, top = 0;
  bool have_run = false, overflow = false;
  XsT run = xs_identity();
  int runE = 0;
  double S = 0.0;
  auto flush_run = [&]() {
    if (!have_run) return;
    if (top < RS_LBSTACK) {
      XsDesc e;
      e.a_s = run.s; e.a_d = (int8_t)run.d; e.e0 = (int16_t)runE; e.has_special = 0;
      e.b_s = 0.0; e.b_d = 0; e.wc = 0.f; e.e1 = 0; e.pad = 0;
      if (lane == 0) stack[top] = e;
      ++top;
    } else overflow = true;
    have_run = false;
  };
  auto add_plain = [&](const XsT& T, int E) {  // T is EARLIER in the sequence than the current run
    if (have_run && runE == E) run = xs_compose<MB>(T, run, E);
    else { flush_run(); run = T; runE = E; have_run = true; }
  };
  for (;;) {
    const int idx = win_base - lane;
    uint32_t st = 4u;  // 4 = before the first tile
    if (idx >= 0) st = rs_wait_state(slots + idx, epoch, 0xEu);
    unsigned m2 = __ballot_sync(0xffffffffu, st == 2u), m3 = __ballot_sync(0xffffffffu, st == 3u);
    int l2 = m2 ? __ffs(m2) - 1 : 32, l3 = m3 ? __ffs(m3) - 1 : 32;
    if (l3 < l2) {  // an opaque tile sits in front of the nearest inclusive state: wait for it to finish
      if (lane == l3) st = rs_wait_state(slots + idx, epoch, 0x4u);
      __syncwarp();
      l2 = l3;
    }
    const int nd = l2;  // lanes [0, nd) hold descriptors
    XsDesc d = {};
    double incl = 0.0;
    if (lane < nd) d = rs_read_desc(slots + idx);
    if (lane == l2) incl = __ldcg(&slots[idx].incl);
    if (nd > 0) {
      const int e_first = __shfl_sync(0xffffffffu, (int)d.e0, 0);
      const bool simple = lane >= nd || (!d.has_special && (int)d.e0 == e_first);
      if (__all_sync(0xffffffffu, simple)) {
        XsT T;
        T.s = lane < nd ? d.a_s : 0.0;
        T.d = lane < nd ? (int)d.a_d : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          XsT up = rs_shfl_down_t(T, o);  // lane + o is a farther = earlier tile
          if (lane + o < 32) T = xs_compose<MB>(up, T, e_first);
        }
        XsT T0;
        T0.s = __shfl_sync(0xffffffffu, T.s, 0);
        T0.d = __shfl_sync(0xffffffffu, T.d, 0);
        add_plain(T0, e_first);
      } else {
        for (int l = 0; l < nd; ++l) {
          XsDesc dl = rs_shfl_desc(d, l);
          if (!dl.has_special) {
            XsT T; T.s = dl.a_s; T.d = dl.a_d;
            add_plain(T, (int)dl.e0);
          } else {
            flush_run();
            if (top < RS_LBSTACK) { if (lane == 0) stack[top] = dl; ++top; } else overflow = true;
          }
        }
      }
    }
    if (l2 < 32) { S = __shfl_sync(0xffffffffu, incl, l2); break; }
    win_base -= 32;
  }
  flush_run();
  __syncwarp();
  bool ok = !overflow;
  for (int k = top - 1; k >= 0 && ok; --k) {
    const XsDesc e = stack[k];
    double S2;
    ok = xs_apply_desc<MB>(S, e, &S2);
    S = S2;
  }
  if (lane == 0) atomicAdd((unsigned long long*)&ctrl->lb_windows, (unsigned long long)((tile - 1 - win_base) / 32 + 1));
  if (!ok) {  // some speculation on the way does not hold for the true state: take the predecessor's own result
    if (lane == 0) atomicAdd(&ctrl->lb_fail, 1);
    const TileSlot* q = slots + (tile - 1);
    rs_wait_state(q, epoch, 0x4u);
    S = __ldcg(&q->incl);
  }
  return S;
}

// ---- the scan + expansion kernel ----------------------------------------------------------------------------------------------
// Block = RS_NT compute threads (8 warps, 16 consecutive weights each) + ONE look-back warp.  The compute warps never wait for
// the exact incoming state of the tile: they run the whole expansion SPECULATIVELY from the approximate prefix (accurate to
// ~1e-13 relative, so the float32-rounded cumulative weights almost always come out identical), while the look-back warp
// chains the exact state across tiles.  When the exact state arrives every thread re-derives its exact cumulative weights
// (a handful of double additions) and compares bit for bit; only a tile where some value differs repeats the expansion.
#define RS_STAGE (2 * RS_TILE)
#define RS_THREADS (RS_NT + 32)
struct RsSmem {
  union {
    int32_t stage[RS_STAGE];     // local index + 1 of the particle owning each output slot (0 = not yet known)
    float c_slow[RS_TILE];       // cumulative weights from the sequential fallback
  };
  XsDesc stack[RS_LBSTACK];
  XsT seg_agg[RS_MAXSEG];
  double base[RS_MAXSEG];
  float seg_wc[RS_MAXSEG];
  int seg_e[RS_MAXSEG];
  double dscratch[33];
  XsSeg sscratch[8];
  int lscratch[8];
  int iscratch[8];
  int lab_end[RS_NT];
  int pre[RS_NT];
  int tile_id, ok, X, e0, table_ok, carry, n_out;
  double S_in, S_out, sp0;
};

__device__ __forceinline__ long long rs_now() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define RS_STAMP(k) do { if (a.dbg && (threadIdx.x & 31) == 0) a.dbg[((int64_t)col * a.tiles_per_col + tile) * 8 + (k)] = rs_now(); } while (0)
__device__ __forceinline__ void rs_bar_arrive(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(RS_THREADS) : "memory"); }
__device__ __forceinline__ void rs_bar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(RS_THREADS) : "memory"); }
__device__ __forceinline__ int rs_cbar_or(int pred) {
  int r;
  asm volatile("{\n .reg .pred p, q;\n setp.ne.s32 p, %1, 0;\n bar.red.or.pred q, 1, %2, p;\n selp.s32 %0, 1, 0, q;\n}\n"
               : "=r"(r) : "r"(pred), "n"(RS_NT) : "memory");
  return r;
}

// rare paths, kept out of line so that the hot loops stay small (instruction-cache footprint)
template <int MB>
__device__ __noinline__ void rs_thread_reduce_slow(const float (&w)[RS_ITEMS], double sp_thread, int lab_prev, int lab_end,
                                                   uint32_t* mask, XsSeg* contrib) {
  xs_thread_reduce<MB, RS_ITEMS>(w, sp_thread, lab_prev, lab_end, mask, contrib);
}
template <int MB>
__device__ __noinline__ void rs_thread_finalize_slow(const float (&w)[RS_ITEMS], uint32_t mask, int lab_prev, int s_base, XsT t_open,
                                                     const double* base, const int* seg_e, float (&c)[RS_ITEMS]) {
  xs_thread_finalize<MB, RS_ITEMS>(w, mask, lab_prev, s_base, t_open, base, seg_e, c);
}
template <int MB>
__device__ __noinline__ void rs_fill_table(const float (&w)[RS_ITEMS], uint32_t mask, double sp_thread, int lab_prev, int lab_end,
                                           XsSeg excl, RsSmem& sm) {
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
__device__ __noinline__ int32_t rs_count_slow(float c, float u, int32_t n, float nf) { return xs_count_le_t<int32_t>(c, u, n, nf); }

// #{ i in [0,n) : fl32(fl32(i + u) / nf) <= c }  (exact_scan.h explains the midpoint argument).  Branch-free for n < 2^24:
// with v = t - u, the probes i <= floor(v) - 1 always qualify and i >= floor(v) + 2 never do (fl32(i + u) is off by at most
// half an ulp <= 0.5), so two float compares at floor(v) and floor(v) + 1 settle the count.
__device__ __forceinline__ int32_t rs_count(float c, float u, int32_t n, float nf) {
  const uint32_t cb = __float_as_uint(c);
  const double t = (0.5 * ((double)c + (double)__uint_as_float(cb + 1u))) * (double)nf;
  float tf = __double2float_rd(t);                       // largest float <= t
  if ((cb & 1u) && (double)tf == t) tf = __uint_as_float(__float_as_uint(tf) - 1u);  // tie rounds away from c: need s < t
  if (n >= (1 << 24) || !(c == c) || !(tf > 0.f)) return rs_count_slow(c, u, n, nf);
  int i = __double2int_rd(t - (double)u);
  i = min(i, n - 1);                                     // i >= -1 because t > 0 and u < 1
  const bool ok0 = (i < 0) || (__fadd_rn((float)i, u) <= tf);
  const bool ok1 = (i + 1 < n) && (__fadd_rn((float)(i + 1), u) <= tf);
  return i + (ok0 ? 1 : 0) + ((ok0 && ok1) ? 1 : 0);
}

// RN_q(w) for an even incoming state in binade E (M = 2^E)
template <int MB>
__device__ __forceinline__ double rs_round_q(float w, double M) {
  return (MB == 53) ? __dadd_rn(__dadd_rn(M, (double)w), -M) : (double)__fadd_rn(__fadd_rn((float)M, w), -(float)M);
}

template <int MB, int OUT>
__global__ void __launch_bounds__(RS_THREADS, 3) systematic_kernel(ResampleArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RsSmem& sm = *reinterpret_cast<RsSmem*>(smem_raw);
  const int tid = threadIdx.x;

  // dynamic tile id: tiles start in id order, so every predecessor a tile waits on is running or finished
  if (tid == 0) sm.tile_id = (int)atomicAdd(&a.ctrl->tile_counter, 1u);
  __syncthreads();
  const int id = sm.tile_id;
  const int col = id / a.tiles_per_col, tile = id % a.tiles_per_col;
  if (col >= a.B) return;
  if (a.stats && !a.stats[col].resample) return;
  if (MB == 53 && OUT == RS_OUT_ANCESTORS && (a.verdict[col] & 1)) return;  // systematic_benign_kernel owns this column
  const uint32_t epoch = a.ctrl->epoch;
  TileSlot* slots = a.slots + (int64_t)col * a.tiles_per_col;

  if (tid >= RS_NT) {
    // ================================ look-back warp ================================
    rs_bar_sync(2);  // the compute warps have written the segment table
    RS_STAMP(5);
    const int lane = tid - RS_NT;
    const int X = sm.X, e0 = sm.e0;
    const bool table_ok = sm.table_ok != 0;
    if (lane == 0 && tile > 0) {  // publish what successors can use before our own incoming state is known
      TileSlot* me = slots + tile;
      if (table_ok && X <= 1) {
        XsDesc d;
        d.a_s = sm.seg_agg[0].s; d.a_d = (int8_t)sm.seg_agg[0].d;
        d.e0 = (int16_t)e0; d.has_special = (int8_t)X; d.pad = 0;
        d.b_s = 0.0; d.b_d = 0; d.wc = 0.f; d.e1 = 0;
        if (X) { d.wc = sm.seg_wc[1]; d.e1 = (int16_t)sm.seg_e[1]; d.b_s = sm.seg_agg[1].s; d.b_d = (int8_t)sm.seg_agg[1].d; }
        me->desc = d;
        st_release_u32(&me->status, (epoch << 2) | 1u);
      } else {
        st_release_u32(&me->status, (epoch << 2) | 3u);
      }
    }
    double S_in = 0.0;
    if (a.approx) S_in = sm.sp0;
    else if (tile > 0) S_in = rs_lookback<MB>(slots, tile, epoch, sm.stack, a.ctrl);
    if (lane == 0) {
      double S_out = S_in;
      bool ok = table_ok && xs_walk_segments<MB>(S_in, e0, X, sm.seg_agg, sm.seg_wc, sm.seg_e, sm.base, &S_out);
      if (!ok) {  // speculation failed verification (or too many segments): genuine sequential adds over the tile
        const float* wrow = a.wn + (int64_t)col * a.ld + (int64_t)tile * RS_TILE;
        double S = S_in;
        for (int k = 0; k < RS_TILE; ++k) S = xs_add_special<MB>(S, __ldg(wrow + k));
        S_out = S;
        atomicAdd(&a.ctrl->slow_tiles, 1);
      }
      sm.S_in = S_in;
      sm.S_out = S_out;
      sm.ok = ok ? 1 : 0;
      slots[tile].incl = S_out;
      st_release_u32(&slots[tile].status, (epoch << 2) | 2u);
    }
    RS_STAMP(6);
    __syncwarp();
    __threadfence_block();
    rs_bar_arrive(3);
    return;
  }

  // ================================ compute warps ================================
  const int32_t n = (int32_t)a.n;
  const float nf = (float)a.n;
  const float u = a.u_col[col];  // systematic offset of this column, settled by normalize_kernel

  if (tid == 0) RS_STAMP(0);
  // ---- this thread's 16 consecutive normalised weights (zero beyond n, written by normalize_kernel)
  float w[RS_ITEMS];
  {
    const float4* src = reinterpret_cast<const float4*>(a.wn + (int64_t)col * a.ld + (int64_t)tile * RS_TILE + tid * RS_ITEMS);
#pragma unroll
    for (int v = 0; v < RS_ITEMS / 4; ++v) {
      const float4 q = __ldg(src + v);
      w[4 * v] = q.x; w[4 * v + 1] = q.y; w[4 * v + 2] = q.z; w[4 * v + 3] = q.w;
    }
  }

  // ---- phase A: approximate prefix -> labels at the thread boundaries
  double sp0;
  {
    const double* ts = a.tilesum + (int64_t)col * a.tiles_per_col;
    double part = 0.0;
    for (int q = tid; q < tile; q += RS_NT) part += ts[q];
    sp0 = rs_block_sum(part, sm.dscratch);
  }
  double tsum = 0.0;
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) tsum += (double)w[j];
  const double sp_thread = sp0 + rs_block_excl_scan(tsum, sm.dscratch);
  const int e0 = xs_label(sp0);
  const int lab_end = xs_label(sp_thread + tsum);
  sm.lab_end[tid] = lab_end;
  rs_cbar();
  const int lab_prev = tid ? sm.lab_end[tid - 1] : e0;

  // ---- phase B: per-thread transducer.  Hot path: a clean thread without ties is a plain sum of RN_q(w_j).
  uint32_t mask = 0;
  XsSeg contrib = xs_seg_identity();
  bool simple = (lab_prev == lab_end);
  const double M = (lab_prev == XS_E_ZERO) ? 0.0 : xs_pow2(lab_prev);
  if (simple && lab_prev != XS_E_ZERO) {
    const double hq = xs_pow2(lab_prev - MB);
    double acc = 0.0;
    bool tie = false;
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
      const double r = rs_round_q<MB>(w[j], M);
      tie |= (fabs(__dadd_rn((double)w[j], -r)) == hq);
      acc = __dadd_rn(acc, r);
    }
    contrib.t.s = acc;
    simple = !tie;
  }
  if (!simple) rs_thread_reduce_slow<MB>(w, sp_thread, lab_prev, lab_end, &mask, &contrib);
  XsSeg total;
  const XsSeg excl = rs_block_excl_scan_seg<MB>(contrib, lab_prev, sm.sscratch, sm.lscratch, &total);
  const int X = total.cnt;
  const bool table_ok = X < RS_MAXSEG;
  if (table_ok && mask) rs_fill_table<MB>(w, mask, sp_thread, lab_prev, lab_end, excl, sm);
  if (tid == 0) {
    if (table_ok) sm.seg_agg[X] = total.t;
    sm.X = X; sm.e0 = e0; sm.table_ok = table_ok ? 1 : 0; sm.sp0 = sp0;
  }
  __threadfence_block();
  if (tid == 0) RS_STAMP(1);
  rs_bar_arrive(2);  // hand the table to the look-back warp and carry on

  // ---- phase D + expansion: first speculatively from the approximate prefix, then (rarely) again from the exact state
  const int64_t g0 = (int64_t)tile * RS_TILE + tid * RS_ITEMS;
  const int lidx0 = tid * RS_ITEMS;  // local index of this thread's first particle
  int32_t* anc = a.anc + (int64_t)col * a.ld;
  float c[RS_ITEMS];
  float c_prev = 0.f;      // float32 cumulative weight just before this thread's first particle
  int32_t n_in = 0;
  bool staged = false;     // the ancestors of the whole tile are sitting in sm.stage / sm.pre
  bool have_exact = false;
  double S_in = 0.0;
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1) {
      if (tid == 0) RS_STAMP(2);
      rs_bar_sync(3);
      if (tid == 0) RS_STAMP(3);  // exact incoming state, segment bases and the verdict of the segment walk are in shared memory
      have_exact = true;
      S_in = sm.S_in;
    }
    // ---- cumulative weights of this thread
    bool mismatch = false;
    if (have_exact && !sm.ok) {
      // sequential fallback: one compute thread redoes the genuine adds and leaves the results in shared memory
      rs_cbar();
      if (tid == 0) {
        const float* wrow = a.wn + (int64_t)col * a.ld + (int64_t)tile * RS_TILE;
        double S = S_in;
        for (int k = 0; k < RS_TILE; ++k) { S = xs_add_special<MB>(S, __ldg(wrow + k)); sm.c_slow[k] = (float)S; }
      }
      rs_cbar();
      mismatch = true;
      c_prev = tid ? sm.c_slow[lidx0 - 1] : (float)S_in;
      for (int k = 0; k < RS_ITEMS; ++k) c[k] = sm.c_slow[lidx0 + k];
      rs_cbar();
    } else if (simple) {
      double S0;
      if (have_exact) {
        const double b = sm.base[excl.cnt];
        S0 = b;
        if (lab_prev != XS_E_ZERO) {
          double inc = excl.t.s;
          if (excl.t.d && xs_parity<MB>(b)) inc = __dadd_rn(inc, (double)excl.t.d * xs_pow2(lab_prev - (MB - 1)));
          S0 = __dadd_rn(b, inc);
        }
      } else {
        S0 = sp_thread;
      }
      const float cp = (float)S0;
      mismatch |= have_exact && (__float_as_uint(cp) != __float_as_uint(c_prev));
      c_prev = cp;
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < RS_ITEMS; ++j) {
        if (lab_prev != XS_E_ZERO) acc = __dadd_rn(acc, rs_round_q<MB>(w[j], M));
        const float cj = (float)__dadd_rn(S0, acc);
        mismatch |= have_exact && (__float_as_uint(cj) != __float_as_uint(c[j]));
        c[j] = cj;
      }
    } else if (have_exact) {
      rs_thread_finalize_slow<MB>(w, mask, lab_prev, excl.cnt, excl.t, sm.base, sm.seg_e, c);
      int sb = excl.cnt;
      double b = sm.base[sb];
      double S0 = b;
      if (lab_prev != XS_E_ZERO) xs_apply<MB>(b, lab_prev, excl.t, &S0);
      c_prev = (float)S0;
      mismatch = true;
    } else {
      mismatch = true;  // threads with ties or specials do not speculate
#pragma unroll
      for (int j = 0; j < RS_ITEMS; ++j) c[j] = 0.f;
    }
    if (pass == 0) {
      if (rs_cbar_or(mismatch)) continue;        // somebody cannot speculate: wait for the exact state
    } else {
      if (tile == 0 && tid == 0) mismatch |= false;
      const int redo = rs_cbar_or(mismatch || !staged);
      if (!redo) break;                          // the speculative expansion was exact: its staged ancestors stand
    }

    if (OUT == RS_OUT_CUMSUM) {  // torch.multinomial's prefix sums: the search over them happens in multinomial_draw_kernel
      if (!have_exact) continue;
      float* dst = a.c_out + (int64_t)col * a.ld + g0;
#pragma unroll
      for (int v = 0; v < RS_ITEMS / 4; ++v)
        reinterpret_cast<float4*>(dst)[v] = make_float4(c[4 * v], c[4 * v + 1], c[4 * v + 2], c[4 * v + 3]);
      return;
    }

    // ---- expansion: particle j owns the probes [count(c_{j-1}), count(c_j))
    int32_t cnt[RS_ITEMS];
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j)
      cnt[j] = (g0 + j >= n - 1) ? n : rs_count(c[j], u, n, nf);  // cumsum[..., -1] = 1.0 (resampling.py:49): every probe is <= 1
    const int32_t lo_thread = (tile == 0 && tid == 0) ? 0 : ((g0 - 1 >= n - 1) ? n : rs_count(c_prev, u, n, nf));
    if (tid == 0) n_in = lo_thread;
    if (tid == RS_NT - 1) sm.n_out = cnt[RS_ITEMS - 1];
    if (tid == 0) { sm.carry = 0; sm.pre[0] = lo_thread; }
    rs_cbar();
    n_in = sm.pre[0];
    const int32_t n_out = sm.n_out;
    const int32_t len = n_out - n_in;
    if (!have_exact && (len > RS_STAGE || len < 0)) { staged = false; continue; }  // too long to hold speculatively
    rs_cbar();
    for (int32_t chunk = n_in; chunk < n_out; chunk += RS_STAGE) {
      const int32_t clen = min(RS_STAGE, n_out - chunk);
      const int kshift = clen > RS_TILE ? 5 : 4;  // slots per thread in the max-scan: 32 or 16
      // (a) clear, (b) every particle with offspring marks the first of its slots, (c) max-scan spreads the marks
      for (int i = tid * 4; i < clen; i += RS_NT * 4) *reinterpret_cast<int4*>(&sm.stage[i]) = make_int4(0, 0, 0, 0);
      rs_cbar();
      {
        int32_t lo = lo_thread;
#pragma unroll
        for (int j = 0; j < RS_ITEMS; ++j) {
          const int32_t hi = cnt[j];
          if (hi > lo) {
            if (lo >= chunk) { if (lo < chunk + clen) sm.stage[lo - chunk] = lidx0 + j + 1; }
            else if (hi > chunk) sm.carry = lidx0 + j + 1;  // its slots began in an earlier chunk (one such particle at most)
          }
          lo = hi;
        }
      }
      rs_cbar();
      int run = 0;
      {
        const int i0 = tid << kshift, i1 = min(i0 + (1 << kshift), clen);
        for (int i = i0; i < i1; ++i) { run = max(run, sm.stage[i]); sm.stage[i] = run; }
      }
      const int pre = max(rs_block_excl_maxscan(run, sm.iscratch), sm.carry);
      sm.pre[tid] = pre;
      rs_cbar();
      if (have_exact) {
        const int32_t tile_base = tile * RS_TILE - 1;
        for (int i = tid; i < clen; i += RS_NT) anc[chunk + i] = tile_base + max(sm.stage[i], sm.pre[i >> kshift]);
        rs_cbar();
        if (tid == 0) sm.carry = max(sm.stage[clen - 1], sm.pre[(clen - 1) >> kshift]);
        rs_cbar();
      }
    }
    if (have_exact) { if (tid == 0) RS_STAMP(7); return; }
    staged = true;
  }
  // ---- the speculative expansion was verified: write its staged ancestors
  {
    const int32_t n_out = sm.n_out;
    const int32_t clen = n_out - n_in;
    const int kshift = clen > RS_TILE ? 5 : 4;
    const int32_t tile_base = tile * RS_TILE - 1;
    for (int i = tid; i < clen; i += RS_NT) anc[n_in + i] = tile_base + max(sm.stage[i], sm.pre[i >> kshift]);
  }
  if (tid == 0) RS_STAMP(4);
}

// ---- benign columns: no rounding anywhere, hence no labels, no transducers, no chaining -----------------------------------------
// One CTA per tile, no dependency between CTAs: the exact state before the tile is the (exact) sum of the preceding tile sums.
// Expansion as in systematic_kernel: particle j owns the output slots [count(c_{j-1}), count(c_j)).  Every particle with
// offspring marks its first slot in a shared-memory window; a "last mark at or before me" scan (ancestors are sorted, so this is
// a running maximum) turns the marks into ancestors, which leave as coalesced 128-bit stores.
#define FB_ROWS 5                              // int4 rows per warp and window
#define FB_WIN (RS_NT / 32 * FB_ROWS * 32 * 4) // 5120 output slots per window
struct FbSmem {
  int32_t stage[FB_WIN];
  double dscratch[33];
  int32_t wtot[RS_NT / 32];
  int32_t carry, n_out;
};

// counts of this thread's particles, marking the first slot of every particle with offspring inside the window [wb, wb + FB_WIN)
__device__ __forceinline__ int32_t fb_mark_pass(const float (&w)[RS_ITEMS], double base, int32_t lo, int32_t gbase, int32_t wb, bool first,
                                                float u, int32_t n, double nd, double nfd, FbSmem& sm) {
  double run = base;
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    run += (double)w[j];
    int32_t hi = xs_count_fast((float)run, u, n, nd, nfd);
    hi = (gbase + j >= n - 1) ? n : hi;  // cumsum[..., -1] = 1.0 (resampling.py:49): every probe is <= 1
    const int32_t r = lo - wb;
    if (hi > lo && (uint32_t)r < (uint32_t)FB_WIN) sm.stage[r] = gbase + j;
    if (!first && hi > lo && r < 0 && hi > wb) sm.carry = gbase + j;  // its slots began in an earlier window (one such particle at most)
    lo = hi;
  }
  return lo;
}

__global__ void __launch_bounds__(RS_NT, 4) systematic_benign_kernel(ResampleArgs a) {
  __shared__ __align__(16) FbSmem sm;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int tile = blockIdx.x, col = blockIdx.y;
  if (a.stats && !a.stats[col].resample) return;
  if (!(a.verdict[col] & 1)) return;
  const int32_t n = (int32_t)a.n;
  const double nd = (double)n, nfd = (double)(float)a.n;
  const float u = a.u_col[col];

  float w[RS_ITEMS];
  {
    const float4* src = reinterpret_cast<const float4*>(a.wn + (int64_t)col * a.ld + (int64_t)tile * RS_TILE + tid * RS_ITEMS);
#pragma unroll
    for (int v = 0; v < RS_ITEMS / 4; ++v) {
      const float4 q = __ldg(src + v);
      w[4 * v] = q.x; w[4 * v + 1] = q.y; w[4 * v + 2] = q.z; w[4 * v + 3] = q.w;
    }
  }
  // clear the first window while the loads are in flight
#pragma unroll
  for (int k = 0; k < FB_ROWS; ++k) *reinterpret_cast<int4*>(&sm.stage[(k * RS_NT + tid) * 4]) = make_int4(-1, -1, -1, -1);
  if (tid == 0) sm.carry = -1;
  // exact state before the tile and before this thread (all additions are exact in a benign column)
  double S_in;
  {
    const double* ts = a.tilesum + (int64_t)col * a.tiles_per_col;
    double part = 0.0;
    for (int q = tid; q < tile; q += RS_NT) part += ts[q];
    S_in = block_allreduce<RS_NT>(part, 0.0, OpSumD(), sm.dscratch);
  }
  double tsum = 0.0;
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) tsum += (double)w[j];
  double base;
  {  // exclusive block scan of the thread sums
    double inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    __syncthreads();
    if (lane == 31) sm.dscratch[wid] = inc;
    __syncthreads();
    double off = 0.0;
#pragma unroll
    for (int k = 0; k < RS_NT / 32; ++k) off += (k < wid) ? sm.dscratch[k] : 0.0;
    base = S_in + (off + (inc - tsum));
  }
  // slots owned by the tile start at n_in = #probes at or below the cumulative weight before the tile
  const int32_t gbase = tile * RS_TILE + tid * RS_ITEMS;  // global index of this thread's first particle
  const int32_t n_in = (tile == 0) ? 0 : ((tile * RS_TILE - 1 >= n - 1) ? n : xs_count_fast((float)S_in, u, n, nd, nfd));
  int32_t lo_thread = n_in;
  if (tid) lo_thread = (gbase - 1 >= n - 1) ? n : xs_count_fast((float)base, u, n, nd, nfd);
  const int32_t wb0 = n_in & ~3;
  {
    const int32_t last = fb_mark_pass(w, base, lo_thread, gbase, wb0, true, u, n, nd, nfd, sm);
    if (tid == RS_NT - 1) sm.n_out = last;
  }
  __syncthreads();
  const int32_t n_out = sm.n_out;

  int32_t* anc = a.anc + (int64_t)col * a.ld;
  int32_t carry = -1;
  for (int32_t wb = wb0; wb < n_out; wb += FB_WIN) {
    if (wb != wb0) {  // rare: more than one window of offspring
      __syncthreads();
#pragma unroll
      for (int k = 0; k < FB_ROWS; ++k) *reinterpret_cast<int4*>(&sm.stage[(k * RS_NT + tid) * 4]) = make_int4(-1, -1, -1, -1);
      if (tid == 0) sm.carry = -1;
      __syncthreads();
      fb_mark_pass(w, base, lo_thread, gbase, wb, false, u, n, nd, nfd, sm);
      __syncthreads();
      carry = max(carry, sm.carry);
    }
    const int32_t wlen = min(FB_WIN, n_out - wb);
    // last mark at or before every slot; warp `wid` owns FB_ROWS rows of 32 int4, row k = int4 [(wid*FB_ROWS + k)*32, +32)
    int4 m[FB_ROWS];
    int32_t cin[FB_ROWS];
    int32_t wrun = -1;  // last mark seen by this warp so far
#pragma unroll
    for (int k = 0; k < FB_ROWS; ++k) {
      const int i4 = (wid * FB_ROWS + k) * 32 + lane;
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
    int32_t cw = carry;  // last mark before this warp's rows
    int32_t call = carry;
#pragma unroll
    for (int k = 0; k < RS_NT / 32; ++k) {
      const int32_t t = sm.wtot[k];
      if (k < wid) cw = max(cw, t);
      call = max(call, t);
    }
    carry = call;
#pragma unroll
    for (int k = 0; k < FB_ROWS; ++k) {
      const int i4 = (wid * FB_ROWS + k) * 32 + lane;
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
