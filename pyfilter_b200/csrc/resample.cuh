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
#define RS_MAXWIN 4
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
  Ctrl* ctrl;
};
enum { RS_OUT_ANCESTORS = 0, RS_OUT_CUMSUM = 1 };

__device__ __forceinline__ void rs_load_weights(const ResampleArgs& a, int col, int tile, float (&w)[RS_ITEMS]) {
  const float* src = a.w + (int64_t)col * a.ld + (int64_t)tile * RS_TILE + threadIdx.x * RS_ITEMS;
  const int64_t g0 = (int64_t)tile * RS_TILE + threadIdx.x * RS_ITEMS;
  float m = 0.f, iz = 1.f;
  if (!a.input_is_w) {
    const ColStats& s = a.stats[col];
    m = a.use_rw ? s.m_rw : s.m_lw;
    iz = a.use_rw ? s.inv_z_rw : s.inv_z_lw;
  }
#pragma unroll
  for (int v = 0; v < RS_ITEMS / 4; ++v) {
    float4 q = __ldg(reinterpret_cast<const float4*>(src) + v);
    float e[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float x = e[k];
      if (!a.input_is_w) x = smcb_weight(smcb_sanitize(x), m, iz);
      if (g0 + v * 4 + k >= a.n) x = 0.f;
      w[v * 4 + k] = x;
    }
  }
}

// ---- pre-pass: fp64 tile sums --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RS_NT) tile_sum_kernel(ResampleArgs a) {
  __shared__ double scratch[33];
  const int col = blockIdx.y, tile = blockIdx.x;
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {  // arm the scan kernel that follows in stream order
    a.ctrl->tile_counter = 0;
    a.ctrl->epoch += 1;
  }
  if (a.stats && !a.stats[col].resample) return;
  float w[RS_ITEMS];
  rs_load_weights(a, col, tile, w);
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) s += (double)w[j];
  s = block_allreduce<RS_NT>(s, 0.0, OpSumD(), scratch);
  if (threadIdx.x == 0) a.tilesum[(int64_t)col * a.tiles_per_col + tile] = s;
}

// ---- block scans ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double rs_block_excl_scan(double v, double* scratch /*>=33*/, double* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) scratch[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    double x = lane < RS_NT / 32 ? scratch[lane] : 0.0;
    double y = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      double t = __shfl_up_sync(0xffffffffu, y, o);
      if (lane >= o) y += t;
    }
    if (lane < RS_NT / 32) scratch[lane] = y - x;  // exclusive warp offsets
    if (lane == RS_NT / 32 - 1) scratch[32] = y;
  }
  __syncthreads();
  *total = scratch[32];
  return scratch[wid] + (inc - v);
}

__device__ __forceinline__ XsSeg rs_shfl_up(const XsSeg& s, int o) {
  XsSeg r;
  r.t.inc0 = __shfl_up_sync(0xffffffffu, (long long)s.t.inc0, o);
  r.t.d = __shfl_up_sync(0xffffffffu, s.t.d, o);
  r.cnt = __shfl_up_sync(0xffffffffu, s.cnt, o);
  return r;
}

// exclusive segmented scan of the per-thread contributions; *total = inclusive result of the whole block
__device__ __forceinline__ XsSeg rs_block_excl_scan_seg(const XsSeg& v, XsSeg* scratch /*>=33*/, XsSeg* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  XsSeg inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    XsSeg t = rs_shfl_up(inc, o);
    if (lane >= o) inc = xs_seg_combine(t, inc);
  }
  XsSeg excl = rs_shfl_up(inc, 1);
  if (lane == 0) excl = xs_seg_identity();
  __syncthreads();
  if (lane == 31) scratch[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    XsSeg x = lane < RS_NT / 32 ? scratch[lane] : xs_seg_identity();
    XsSeg y = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      XsSeg t = rs_shfl_up(y, o);
      if (lane >= o) y = xs_seg_combine(t, y);
    }
    XsSeg ye = rs_shfl_up(y, 1);
    if (lane == 0) ye = xs_seg_identity();
    if (lane == RS_NT / 32 - 1) scratch[32] = y;
    __syncwarp();
    if (lane < RS_NT / 32) scratch[lane] = ye;
  }
  __syncthreads();
  *total = scratch[32];
  return xs_seg_combine(scratch[wid], excl);
}

// ---- decoupled look-back on exact states (warp 0) ---------------------------------------------------------------------------
__device__ __forceinline__ XsDesc rs_shfl_desc(const XsDesc& d, int src) {
  XsDesc r;
  r.a_inc0 = __shfl_sync(0xffffffffu, (long long)d.a_inc0, src);
  r.b_inc0 = __shfl_sync(0xffffffffu, (long long)d.b_inc0, src);
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
  const longlong2 q0 = __ldcg(reinterpret_cast<const longlong2*>(&s->desc));
  const int4 q1 = __ldcg(reinterpret_cast<const int4*>(&s->desc) + 1);
  d.a_inc0 = q0.x; d.b_inc0 = q0.y;
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

template <int MB>
__device__ double rs_lookback(const TileSlot* slots, int tile, uint32_t epoch, XsDesc (*stack)[32]) {
  const int lane = threadIdx.x & 31;
  int win_base = tile - 1, nwin = 0;
  double S = 0.0;
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
    if (l2 == 32 && nwin == RS_MAXWIN) continue;  // stack full: keep polling this window until an inclusive state shows up
    XsDesc d = {};
    double incl = 0.0;
    if (idx >= 0 && lane < l2) d = rs_read_desc(slots + idx);
    if (lane == l2) incl = __ldcg(&slots[idx].incl);
    if (l2 == 32) {  // 32 more descriptors: stash, walk further back
      stack[nwin][lane] = d;
      ++nwin;
      win_base -= 32;
      __syncwarp();
      continue;
    }
    S = __shfl_sync(0xffffffffu, incl, l2);
    for (int l = l2 - 1; l >= 0; --l) {
      XsDesc dl = rs_shfl_desc(d, l);
      double S2;
      if (!xs_apply_desc<MB>(S, dl, &S2)) {  // that tile's speculation does not hold for the true state: use its own result
        const TileSlot* q = slots + (win_base - l);
        rs_wait_state(q, epoch, 0x4u);
        S2 = __ldcg(&q->incl);
      }
      S = S2;
    }
    for (int wdx = nwin - 1; wdx >= 0; --wdx) {
      const int wb = tile - 1 - 32 * wdx;
      for (int l = 31; l >= 0; --l) {
        XsDesc dl = stack[wdx][l];
        double S2;
        if (!xs_apply_desc<MB>(S, dl, &S2)) {
          const TileSlot* q = slots + (wb - l);
          rs_wait_state(q, epoch, 0x4u);
          S2 = __ldcg(&q->incl);
        }
        S = S2;
      }
    }
    return S;
  }
}

// ---- the scan + expansion kernel ----------------------------------------------------------------------------------------------
struct RsSmem {
  union {
    int32_t stage[RS_TILE];      // ancestors staged for coalesced stores
    float c_slow[RS_TILE];       // cumulative weights from the sequential fallback
  };
  XsDesc stack[RS_MAXWIN][32];
  XsT seg_agg[RS_MAXSEG];
  double base[RS_MAXSEG];
  float seg_wc[RS_MAXSEG];
  int seg_e[RS_MAXSEG];
  double dscratch[33];
  XsSeg sscratch[33];
  int lab_end[RS_NT];
  int32_t cnt_last[RS_NT];
  int64_t big_lo[RS_BIGLIST], big_hi[RS_BIGLIST];
  int32_t big_val[RS_BIGLIST];
  int big_n;
  int tile_id;
  int ok;
  double S_in, S_out;
};

template <int MB, int OUT>
__global__ void __launch_bounds__(RS_NT) systematic_kernel(ResampleArgs a) {
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
  const uint32_t epoch = a.ctrl->epoch;
  TileSlot* slots = a.slots + (int64_t)col * a.tiles_per_col;
  const int64_t n = a.n;
  const float nf = (float)n;

  // systematic offset of this column (one uniform per column, resampling.py:41)
  float u;
  if (a.u_in) u = a.u_in[col];
  else {
    Philox4 r = philox4x32_10((uint32_t)col, 0u, (uint32_t)a.ctrl->t, SMCB_RNG_SYSTEMATIC, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
    u = smcb_u01(r.x);
  }
  if (a.u_out && tile == 0 && tid == 0) a.u_out[col] = u;

  float w[RS_ITEMS];
  rs_load_weights(a, col, tile, w);
  if (a.w_out) {
    float* dst = a.w_out + (int64_t)col * a.ld + (int64_t)tile * RS_TILE + tid * RS_ITEMS;
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) dst[j] = w[j];
  }

  // ---- phase A: approximate prefix -> labels
  double sp0;
  {
    const double* ts = a.tilesum + (int64_t)col * a.tiles_per_col;
    double part = 0.0;
    for (int q = tid; q < tile; q += RS_NT) part += ts[q];
    sp0 = block_allreduce<RS_NT>(part, 0.0, OpSumD(), sm.dscratch);
  }
  double tsum = 0.0;
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) tsum += (double)w[j];
  double tot;
  const double sp_thread = sp0 + rs_block_excl_scan(tsum, sm.dscratch, &tot);
  const int e0 = xs_label(sp0);
  sm.lab_end[tid] = xs_thread_end_label<RS_ITEMS>(w, sp_thread);
  __syncthreads();
  const int lab_prev = tid ? sm.lab_end[tid - 1] : e0;

  // ---- phase B: transducers, segmented scan, segment table
  uint32_t mask;
  XsSeg contrib;
  XsT pre;
  xs_thread_label_and_reduce<MB, RS_ITEMS>(w, sp_thread, lab_prev, &mask, &contrib, &pre);
  XsSeg total;
  const XsSeg excl = rs_block_excl_scan_seg(contrib, sm.sscratch, &total);
  const int X = total.cnt;
  const bool table_ok = X < RS_MAXSEG;
  if (table_ok && mask) {
    int s = excl.cnt;
    XsT T = excl.t;
    int E = lab_prev;
    double sp = sp_thread;
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
      sp += (double)w[j];
      if (mask & (1u << j)) {
        sm.seg_agg[s] = T;
        ++s;
        sm.seg_wc[s] = w[j];
        E = xs_label(sp);
        sm.seg_e[s] = E;
        T = xs_identity();
      } else {
        T = xs_compose(T, xs_elem<MB>(w[j], E));
      }
    }
  }
  if (tid == 0 && table_ok) sm.seg_agg[X] = total.t;
  __syncthreads();

  // ---- publish what successors can use before our own incoming state is known
  if (tid == 0 && tile > 0) {
    TileSlot* me = slots + tile;
    if (table_ok && X <= 1) {
      XsDesc d;
      d.a_inc0 = sm.seg_agg[0].inc0; d.a_d = (int8_t)sm.seg_agg[0].d;
      d.e0 = (int16_t)e0; d.has_special = (int8_t)X; d.pad = 0;
      d.b_inc0 = 0; d.b_d = 0; d.wc = 0.f; d.e1 = 0;
      if (X) { d.wc = sm.seg_wc[1]; d.e1 = (int16_t)sm.seg_e[1]; d.b_inc0 = sm.seg_agg[1].inc0; d.b_d = (int8_t)sm.seg_agg[1].d; }
      me->desc = d;
      st_release_u32(&me->status, (epoch << 2) | 1u);
    } else {
      st_release_u32(&me->status, (epoch << 2) | 3u);
    }
  }

  // ---- exact incoming state
  if (tid < 32) {
    double S_in = 0.0;
    if (tile > 0) S_in = rs_lookback<MB>(slots, tile, epoch, sm.stack);
    if (tid == 0) {
      // ---- phase C
      double S_out = S_in;
      bool ok = table_ok && xs_walk_segments<MB>(S_in, e0, X, sm.seg_agg, sm.seg_wc, sm.seg_e, sm.base, &S_out);
      sm.S_in = S_in;
      sm.ok = ok ? 1 : 0;
      if (ok) {
        sm.S_out = S_out;
        slots[tile].incl = S_out;
        st_release_u32(&slots[tile].status, (epoch << 2) | 2u);
      }
    }
  }
  __syncthreads();
  const double S_in = sm.S_in;
  float c[RS_ITEMS];
  if (sm.ok) {
    // ---- phase D
    xs_thread_finalize<MB, RS_ITEMS>(w, mask, lab_prev, excl.cnt, excl.t, sm.base, sm.seg_e, c);
  } else {
    // ---- sequential fallback (speculation failed verification, or too many segments): genuine adds in element order
    for (int k = 0; k < RS_ITEMS; ++k) sm.c_slow[tid * RS_ITEMS + k] = w[k];
    __syncthreads();
    if (tid == 0) {
      double S = S_in;
      for (int k = 0; k < RS_TILE; ++k) {
        S = xs_add_special<MB>(S, sm.c_slow[k]);
        sm.c_slow[k] = (float)S;
      }
      sm.S_out = S;
      slots[tile].incl = S;
      st_release_u32(&slots[tile].status, (epoch << 2) | 2u);
      atomicAdd(&a.ctrl->slow_tiles, 1);
    }
    __syncthreads();
    for (int k = 0; k < RS_ITEMS; ++k) c[k] = sm.c_slow[tid * RS_ITEMS + k];
    __syncthreads();
  }

  const int64_t g0 = (int64_t)tile * RS_TILE + tid * RS_ITEMS;
  if (OUT == RS_OUT_CUMSUM) {  // torch.multinomial's prefix sums: the search over them happens in multinomial_draw_kernel
    float* dst = a.c_out + (int64_t)col * a.ld + g0;
#pragma unroll
    for (int v = 0; v < RS_ITEMS / 4; ++v)
      reinterpret_cast<float4*>(dst)[v] = make_float4(c[4 * v], c[4 * v + 1], c[4 * v + 2], c[4 * v + 3]);
    return;
  }

  // ---- expansion: particle j owns probes [count(c_{j-1}), count(c_j))
  int32_t cnt[RS_ITEMS];
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    const int64_t g = g0 + j;
    cnt[j] = (int32_t)((g >= n - 1) ? n : xs_count_le(c[j], u, n, nf));  // cumsum[..., -1] = 1.0 (resampling.py:49)
  }
  sm.cnt_last[tid] = cnt[RS_ITEMS - 1];
  if (tid == 0) sm.big_n = 0;
  __syncthreads();
  const int64_t n_in = tile == 0 ? 0 : ((int64_t)tile * RS_TILE - 1 >= n - 1 ? n : xs_count_le((float)S_in, u, n, nf));
  const int64_t lo_thread = tid ? sm.cnt_last[tid - 1] : n_in;
  const int64_t n_out = sm.cnt_last[RS_NT - 1];
  int32_t* anc = a.anc + (int64_t)col * a.ld;

  // very prolific particles are filled cooperatively, straight to global memory
  {
    int64_t lo = lo_thread;
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
      const int64_t hi = cnt[j];
      if (hi - lo >= RS_BIG) {
        int slot = atomicAdd(&sm.big_n, 1);
        if (slot < RS_BIGLIST) { sm.big_lo[slot] = lo; sm.big_hi[slot] = hi; sm.big_val[slot] = (int32_t)(g0 + j); }
      }
      lo = hi > lo ? hi : lo;
    }
  }
  __syncthreads();
  const int nbig = sm.big_n < RS_BIGLIST ? sm.big_n : RS_BIGLIST;
  const bool big_overflow = sm.big_n > RS_BIGLIST;
  for (int b = 0; b < nbig; ++b) {
    const int64_t lo = sm.big_lo[b], hi = sm.big_hi[b];
    const int32_t v = sm.big_val[b];
    for (int64_t i = lo + tid; i < hi; i += RS_NT) anc[i] = v;
  }
  for (int64_t chunk = n_in; chunk < n_out; chunk += RS_TILE) {
    const int64_t chunk_hi = chunk + RS_TILE < n_out ? chunk + RS_TILE : n_out;
    int64_t lo = lo_thread;
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
      const int64_t hi = cnt[j];
      if (hi > lo) {
        if (hi - lo < RS_BIG || big_overflow) {
          const int64_t s0 = lo > chunk ? lo : chunk, s1 = hi < chunk_hi ? hi : chunk_hi;
          for (int64_t i = s0; i < s1; ++i) sm.stage[i - chunk] = (int32_t)(g0 + j);
        }
        lo = hi;
      }
    }
    __syncthreads();
    for (int64_t i = chunk + tid; i < chunk_hi; i += RS_NT) {
      // skip slots owned by a cooperatively filled particle (their stage entry is stale)
      bool owned = false;
      if (!big_overflow)
        for (int b = 0; b < nbig; ++b) owned |= (i >= sm.big_lo[b] && i < sm.big_hi[b]);
      if (!owned) anc[i] = sm.stage[i - chunk];
    }
    __syncthreads();
  }
}
