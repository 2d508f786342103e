// move_kernel: ONE kernel per filter move for columns of more than one tile (systematic resampling, rounding-free weights).
//
// The two-kernel pipeline (resample_fused_kernel -> step_kernel) writes the ancestors to global memory and reads them back to gather
// x_{t-1}; but a tile of 4096 particles OWNS the output slots [count(c_in), count(c_out)) and every ancestor of those slots is one of
// the tile's own particles.  So the move is done where the expansion happens:
//
//   log-weights of the tile -> weights (registers) -> exact fp64 tile sum, published with the launch epoch -> exact sum of the
//   preceding tiles (parallel polling) -> probe counts -> marks in the shared-memory window -> "last mark" scan = ancestors, in
//   place -> gather x_{t-1} from the tile's copy in shared memory -> Philox / proposal / densities / folded look-ahead (the SAME
//   Proposal<> and model functions, the same Philox counters per output slot as step_kernel) -> 128-bit stores of x_t and the
//   weight row -> soft-max partials per TILE -> the block that completes a column folds them (finalize_column).
//
// Global traffic per particle: resampling log-weight 4 B in, x_{t-1} 4 D in (coalesced: the tile's own particles), x_t 4 D and one
// weight row 4 B out - the ancestors never leave the chip (they are stored, with the plain log-weights of an APF move whose look-ahead
// is folded, only when the caller asks for the API-visible state: StepArgs.store_lw).
// A tile stores new weights at its output SLOTS while another tile may not have read its input PARTICLES yet, so the weight rows
// ping-pong with the move index like the state buffers (StepArgs.lw / rw in, lw_out / rw_out out).
//
// Scheduling: one block per tile; a block draws its tile id from a global ticket counter.  Tickets are handed out in start order, so
// every predecessor a tile polls has been started by a running block - no assumption about the order in which the hardware
// dispatches blocks, no co-residency requirement.  The statistics are reduced per TILE, so the result does not depend on which
// block drew which ticket: same seed, same bits.
// Columns that do not resample in this move (SISR below the ESS threshold, a missing observation) take the same loop with identity
// ancestors and their carried log-weights.
#pragma once
#include "step.cuh"
#include "resample.cuh"

#define MV_NT 256
#define MV_ITEMS 16
#define MV_TILE (MV_NT * MV_ITEMS)            // == RS_TILE: the row pitch is a multiple of it
#define MV_PER 20                             // window slots per thread in the "last mark" scan
#define MV_WIN (MV_NT * MV_PER)               // 5120 output slots per window
#define MV_U_HOST 8
#define MV_LB 8                               // look-back: slots per lane in flight
#ifndef SMCB_MV_MINB
#define SMCB_MV_MINB 4
#endif
static_assert(MV_TILE == RS_TILE, "tiles of the move kernel are the resampling tiles");

struct MoveArgs {
  StepArgs s;                 // buffers, parameters, history; s.partials = per-tile records (B, tiles_per_col), s.blocks_per_col = tiles_per_col
  int32_t tiles_per_col;
  int32_t total_tiles;        // B * tiles_per_col
  uint32_t ticket_base;       // value of *tile_counter when this launch starts (the counter is never reset: wrap-around arithmetic)
  uint32_t* tile_counter;
  unsigned long long* mslots; // (B, tiles_per_col) tile sums tagged with the launch epoch, ONE 64-bit word each: a sum of weights that are
                              // multiples of 2^-52 below 2 is a 54-bit integer count of quanta, the 10 bits above it carry the tag
  unsigned long long epoch;   // tag of this launch in [1, 1023], different from the previous launch's (every tile of every launch
                              // rewrites its slot, so a slot never holds a tag older than one launch)
  const float* u_in;          // optional injected systematic offsets (B)
  float* u_out;               // optional dump of the offsets used (B)
  float* w_out;               // optional dump of the normalised resampling weights (B, ld)
  long long* wd;              // watchdog / diagnostics: [0] polls that had to wait, [1] tiles with more than one window
  long long* tl;              // optional timeline (SMCB_DEBUG_TIMELINE): 8 globaltimer stamps per tile
  int32_t n_u_host;           // B when B <= MV_U_HOST and no offsets are injected: the host evaluated the Philox offsets (same function)
  float u_host[MV_U_HOST];
};

template <int D>
struct MoveSmem {
  int32_t stage[MV_WIN];      // marks, then (in place) the ancestors of the window; identity tiles: the carried log-weights
  double scan_a[MV_NT / 32];  // every collective of the tile has its own scratch area: no "protect the reuse" barriers
  double s_in;                // exact sum of the preceding tiles (look-back warp -> everybody)
  int32_t wtot[MV_NT / 32];
  int32_t carry;
  int32_t ticket, is_last;
  float u;
  float Ps[SMCB_NPARAM];
  FinSmem<D> fin;
  FinPre fin_pre;
};

// a block-uniform flag / value the compiler can SEE is uniform (vote / redux result): branches on it need no divergence handling
// and the collectives behind them no re-convergence
#define MV_STAMP(k) do { if (c.tl && tid == 0) c.tl[(int64_t)ticket * 8 + (k)] = st_now(); } while (0)
__device__ __forceinline__ bool mv_uniform(bool f) { return __all_sync(0xffffffffu, f) != 0; }
__device__ __forceinline__ int32_t mv_uniform(int32_t v) { return __reduce_max_sync(0xffffffffu, v); }

// exclusive block scan / block sum of one double per thread with ONE barrier (dedicated scratch).  Exact in any order here: the
// summands are multiples of 2^-52 below 2.
__device__ __forceinline__ double mv_block_excl_scan(double v, double* scratch, double* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) scratch[wid] = inc;
  __syncthreads();
  double off = 0.0, tot = 0.0;
#pragma unroll
  for (int k = 0; k < MV_NT / 32; ++k) {
    const double s = scratch[k];
    off += (k < wid) ? s : 0.0;
    tot += s;
  }
  *total = tot;
  return off + (inc - v);
}
// the look-back of one warp: exact sum of the tile sums published by the tiles before `tile` (MV_LB slots per lane in flight).
// Out of line: only warp 0 of a block pays for the registers of the loads in flight.
#define MV_SLOT_MASK ((1ull << 54) - 1ull)
__device__ __noinline__ double mv_lookback(const unsigned long long* slots, int tile, unsigned long long epoch, long long* wd) {
  const int lane = threadIdx.x & 31;
  unsigned long long part = 0ull;   // integer quanta: exact in any order
  int waited = 0;
  const long long t_begin = wd ? st_now() : 0;
  for (int q0 = tile - 1 - lane; q0 >= 0; q0 -= 32 * MV_LB) {
    unsigned long long v[MV_LB];
#pragma unroll
    for (int j = 0; j < MV_LB; ++j) {
      const int q = q0 - 32 * j;
      v[j] = (q >= 0) ? __ldcg(slots + q) : (epoch << 54);
    }
#pragma unroll
    for (int j = 0; j < MV_LB; ++j) {
      const int q = q0 - 32 * j;
      while ((v[j] >> 54) != epoch) {
        __nanosleep(20);
        v[j] = __ldcg(slots + q);
        ++waited;
      }
      part += v[j] & MV_SLOT_MASK;
    }
  }
  if (wd) {  // diagnostics: tiles that had to wait, re-polls, time spent looking back (ns, lane 0)
    if (waited) atomicAdd((unsigned long long*)&wd[2], (unsigned long long)waited);
    if (lane == 0) { atomicAdd((unsigned long long*)&wd[3], (unsigned long long)(st_now() - t_begin)); if (waited) atomicAdd((unsigned long long*)&wd[0], 1ull); }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  return (double)part * 2.220446049250313e-16;   // quanta * 2^-52: exact (the sum of a column is below 2)
}

// weights of this thread's 16 particles, rounded to multiples of 2^-52 (exactly what normalize_kernel / resample_fused_kernel compute)
template <bool INNER>
__device__ __forceinline__ double mv_weights(const float (&win)[MV_ITEMS], float m, float iz, int32_t gbase, int32_t n, double (&wq)[MV_ITEMS]) {
  double tsum = 0.0;
#pragma unroll
  for (int j = 0; j < MV_ITEMS; ++j) {
    float x = smcb_weight(win[j], m, iz);
    if (!INNER && gbase + j >= n) x = 0.f;
    wq[j] = __dadd_rn(__dadd_rn(1.0, (double)x), -1.0);
    tsum += wq[j];
  }
  return tsum;
}

// probes at or below a cumulative weight, clamped into [.., nmax]; LEAN: exact_scan.h xs_count_lean (n <= 2^23, u == 0 or u >= 2^-64)
template <bool LEAN>
__device__ __forceinline__ int32_t mv_count(float c, float u, int32_t n, int32_t nmax, double nfd) {
  if (LEAN) return min(xs_count_lean(c, u, n, nfd), nmax);
  return min(rs_count_slow(c, u, n, (float)n), nmax);
}

// counts of this thread's particles; every particle with offspring marks its first slot inside the window [wb, wb + MV_WIN).
// CHECK = false: the whole tile fits the window (n_out - wb <= MV_WIN), no range test per particle.
// A mark is the index of the particle INSIDE the tile (the gather of x_{t-1} needs nothing else).
template <bool INNER, bool CHECK, bool FIRST, bool LEAN, typename SM>
__device__ __forceinline__ void mv_mark(const double (&wq)[MV_ITEMS], double S0, int32_t lo, int32_t gbase, int32_t wb, float u,
                                        int32_t n, int32_t n_out, double nfd, SM& sm) {
  double run = S0;
  const int32_t lbase = (int32_t)threadIdx.x * MV_ITEMS;
#pragma unroll
  for (int j = 0; j < MV_ITEMS; ++j) {
    run = __dadd_rn(run, wq[j]);
    int32_t hi = mv_count<LEAN>((float)run, u, n, n_out, nfd);
    if (!INNER) hi = (gbase + j >= n - 1) ? n_out : hi;  // cumsum[..., -1] = 1.0 (resampling.py:49): every probe is <= 1 (n_out == n here)
    const int32_t r = lo - wb;
    if (CHECK) {
      if (hi > lo && (uint32_t)r < (uint32_t)MV_WIN) sm.stage[r] = lbase + j;
      if (!FIRST && hi > lo && r < 0 && hi > wb) sm.carry = lbase + j;  // its slots began in an earlier window (one such particle at most)
    } else {
      if (hi > lo) sm.stage[r] = lbase + j;
    }
    lo = max(lo, hi);
  }
}

// later windows of a tile with more than MV_WIN offspring, or any window when the lean count does not apply (rare: degenerate
// weights, odd injected offsets): the weights are derived again from the log-weights (same function, same bits) instead of being
// kept in registers across the propagation
template <typename SM>
__device__ __noinline__ void mv_remark(const float* wsrc, float m, float iz, double S0, int32_t lo, int32_t gbase, int32_t wb, bool first,
                                       bool lean, float u, int32_t n, int32_t n_out, double nfd, SM& sm) {
  float win[MV_ITEMS];
#pragma unroll
  for (int v = 0; v < MV_ITEMS / 4; ++v) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(wsrc) + v);
    win[4 * v] = q.x; win[4 * v + 1] = q.y; win[4 * v + 2] = q.z; win[4 * v + 3] = q.w;
  }
  double wq[MV_ITEMS];
  mv_weights<false>(win, m, iz, gbase, n, wq);
  if (lean) {
    if (first) mv_mark<false, true, true, true>(wq, S0, lo, gbase, wb, u, n, n_out, nfd, sm);
    else mv_mark<false, true, false, true>(wq, S0, lo, gbase, wb, u, n, n_out, nfd, sm);
  } else {
    if (first) mv_mark<false, true, true, false>(wq, S0, lo, gbase, wb, u, n, n_out, nfd, sm);
    else mv_mark<false, true, false, false>(wq, S0, lo, gbase, wb, u, n, n_out, nfd, sm);
  }
}

// "last mark at or before every slot" over the window in shared memory; the ancestors replace the marks in place.  A thread owns
// MV_PER consecutive slots: running maximum in registers, one warp scan of the per-thread maxima, one cross-warp step.
template <typename SM>
__device__ __forceinline__ int32_t mv_emit(SM& sm, int32_t carry) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int4 m[MV_PER / 4];
#pragma unroll
  for (int k = 0; k < MV_PER / 4; ++k) m[k] = *reinterpret_cast<const int4*>(&sm.stage[tid * MV_PER + 4 * k]);
  int32_t v = -1;
#pragma unroll
  for (int k = 0; k < MV_PER / 4; ++k) v = max(max(v, max(m[k].x, m[k].y)), max(m[k].z, m[k].w));
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v = max(v, t);
  }
  int32_t run = __shfl_up_sync(0xffffffffu, v, 1);
  if (lane == 0) run = -1;
  if (lane == 31) sm.wtot[wid] = v;
  __syncthreads();
  const int32_t wt = (lane < MV_NT / 32) ? sm.wtot[lane] : -1;
  run = max(run, max(carry, __reduce_max_sync(0xffffffffu, (lane < wid) ? wt : -1)));
  carry = max(carry, __reduce_max_sync(0xffffffffu, wt));
#pragma unroll
  for (int k = 0; k < MV_PER / 4; ++k) {
    m[k].x = run = max(run, m[k].x);
    m[k].y = run = max(run, m[k].y);
    m[k].z = run = max(run, m[k].z);
    m[k].w = run = max(run, m[k].w);
    *reinterpret_cast<int4*>(&sm.stage[tid * MV_PER + 4 * k]) = m[k];
  }
  __syncthreads();
  return carry;
}

template <int MODEL, int PROP, int ALG>
__global__ void __launch_bounds__(MV_NT, SMCB_MV_MINB) move_kernel(MoveArgs c) {
  typedef Model<MODEL> M;
  constexpr int D = M::D, OD = M::OD;
  extern __shared__ __align__(16) float mv_x[];   // (D, MV_TILE): x_{t-1} of the tile
  __shared__ __align__(16) MoveSmem<D> sm;
  const StepArgs& a = c.s;
  const int tid = threadIdx.x;
  const int T = c.tiles_per_col;
  const int t = a.t_host;
  // ---- everything that does not depend on the previous kernel, before the grid dependency is awaited
  if (tid == 0) {
    // tile ticket: tickets are handed out in start order, so every predecessor this tile will poll belongs to a started block.
    // The counter is quiet: the blocks of the previous move kernel drew their tickets before they let this launch start.
    const int32_t ticket = (int32_t)(atomicAdd(c.tile_counter, 1u) - c.ticket_base);
    sm.ticket = ticket;
    const int col = ticket / T;
    float u;  // one uniform per column (resampling.py:41)
    if (c.u_in) u = c.u_in[col];
    else if (col < c.n_u_host) u = c.u_host[col];
    else {
      const Philox4 r4 = philox4x32_10((uint32_t)(col + a.col0), 0u, (uint32_t)t, SMCB_RNG_SYSTEMATIC, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
      u = smcb_u01(r4.x);
    }
    sm.u = u;
    sm.carry = -1;
  }
#pragma unroll
  for (int k = 0; k < MV_PER / 4; ++k) *reinterpret_cast<int4*>(&sm.stage[(k * MV_NT + tid) * 4]) = make_int4(-1, -1, -1, -1);
  float y[OD], yn[OD];
  const bool observed = mv_uniform(st_load_obs<OD>(a.y_t, y));
  const bool fold = mv_uniform((ALG == SMCB_ALG_APF) && a.fold && st_load_obs<OD>(a.y_next, yn));
  const int32_t n = (int32_t)a.n;
  const float inv_n = 1.0f / (float)a.n;
  const double nfd = (double)(float)a.n;
  const float one4[4] = {1.f, 1.f, 1.f, 1.f};
  const float* Ps = sm.Ps;
  const bool plain_noise = !a.eps_in && !a.eps_out;
  __syncthreads();
  const int ticket = mv_uniform(sm.ticket);
  const int col = ticket / T, tile = ticket - col * T;
  const float u = sm.u;
  if (tid < SMCB_NPARAM) sm.Ps[tid] = a.P[(int64_t)col * SMCB_NPARAM + tid];  // (first read behind several barriers)
  const int32_t tile_base = tile * MV_TILE;
  const int32_t gbase = tile_base + tid * MV_ITEMS;
  const int64_t rowoff = (int64_t)col * a.ld;
  float* const lwrow = a.lw_out + rowoff;
  float* const rwrow = a.rw_out + rowoff;
  int32_t* const pirow = a.prev_inds + rowoff;
  const float* xprev[D];
  float* xnext[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    xprev[d] = a.xbuf[t & 1] + ((int64_t)d * a.B + col) * a.ld;
    xnext[d] = a.xbuf[(t + 1) & 1] + ((int64_t)d * a.B + col) * a.ld;
  }
  const bool use_rw = (ALG == SMCB_ALG_APF) && observed;
  const float* const winrow = (use_rw ? a.rw : a.lw) + rowoff;
  // the lean probe count needs u == 0 or u >= 2^-64 (exact_scan.h); Philox offsets are multiples of 2^-24
  const bool lean = mv_uniform((u == 0.f) || (u >= 5.5e-20f && u < 1.0f));

  MV_STAMP(0);
  pdl_wait();
  MV_STAMP(1);
  // ---- the tile comes on chip: resampling log-weights (blocked: a thread owns 16 consecutive particles), x_{t-1} (striped),
  //      the column's normalisers (every thread, one broadcast transaction per warp)
  float win[MV_ITEMS];
#pragma unroll
  for (int v = 0; v < MV_ITEMS / 4; ++v) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(winrow + gbase) + v);
    win[4 * v] = q.x; win[4 * v + 1] = q.y; win[4 * v + 2] = q.z; win[4 * v + 3] = q.w;
  }
  const float4* stp = reinterpret_cast<const float4*>(a.stats + col);
  const float4 st0 = stp[0], st1 = stp[1], st2 = stp[2];   // m_lw z_lw inv_z_lw m_rw | z_rw inv_z_rw ess resample | shift[3] ll_aux
#pragma unroll
  for (int d = 0; d < D; ++d) {
#pragma unroll
    for (int v = 0; v < MV_ITEMS / 4; ++v) {
      const int e = (v * MV_NT + tid) * 4;
      *reinterpret_cast<float4*>(mv_x + d * MV_TILE + e) = __ldg(reinterpret_cast<const float4*>(xprev[d] + tile_base + e));
    }
  }
  if (tid < (int)(sizeof(ColStats) / 4)) reinterpret_cast<float*>(&sm.fin_pre.st)[tid] = reinterpret_cast<const float*>(a.stats + col)[tid];
  if (tid == 32) { sm.fin_pre.observed = observed; sm.fin_pre.fold = fold; sm.fin_pre.ll_total = a.ll_total[col]; }
  const float st_m_lw = st0.x, st_inv_z_lw = st0.z;
  // SISR resamples when the ESS test fired (sisr.py:19-26), the APF on every observed step (apf.py:29-34, filters/base.py:213)
  const bool resampled = (ALG == SMCB_ALG_APF) ? observed : mv_uniform(__float_as_int(st1.w) != 0);
  const float wm = use_rw ? st0.w : st0.x;
  const float wiz = use_rw ? st1.y : st0.z;
  const float shift3[3] = {st2.x, st2.y, st2.z};
  float shift[D];
#pragma unroll
  for (int d = 0; d < D; ++d) shift[d] = shift3[d];
  if (c.u_out && resampled && tile == 0 && tid == 0) c.u_out[col] = u;

  int32_t n_in, n_out, wb0;
  double S0 = 0.0;
  int32_t lo_thread = 0;
  if (resampled) {
    const bool inner = mv_uniform(tile_base + MV_TILE <= n - 1);  // neither padding nor the last particle of the column in this tile
    double wq[MV_ITEMS];
    const double tsum = inner ? mv_weights<true>(win, wm, wiz, gbase, n, wq) : mv_weights<false>(win, wm, wiz, gbase, n, wq);
    if (c.w_out) {
      float* dst = c.w_out + rowoff + gbase;
#pragma unroll
      for (int v = 0; v < MV_ITEMS / 4; ++v)
        reinterpret_cast<float4*>(dst)[v] = make_float4((float)wq[4 * v], (float)wq[4 * v + 1], (float)wq[4 * v + 2], (float)wq[4 * v + 3]);
    }
    double tot;
    const double ex = mv_block_excl_scan(tsum, sm.scan_a, &tot);
    const unsigned long long* slots = c.mslots + (int64_t)col * T;
    if (tid == 0)  // publish: ONE 64-bit exchange at the L2 (no store lingering in the SM's write path, no fence) carries sum and tag
      atomicExch(c.mslots + (int64_t)col * T + tile, (c.epoch << 54) | (__double2ull_rn(tot * 4503599627370496.0) & MV_SLOT_MASK));
    MV_STAMP(2);
    // exact sum of the preceding tiles (any order: the weights are multiples of 2^-52).  ONE warp looks back, eight slots per lane in
    // flight at a time: 256 polling threads per block times ~600 resident blocks queue up on the few cache lines that hold the slots
    // (measured: 5 us between the last publication and the last tile knowing its prefix), a single round of coalesced 512-byte
    // reads does not.
    if (tid < 32) {
      const double part = mv_lookback(slots, tile, c.epoch, c.wd);
      if (tid == 0) sm.s_in = part;
    }
    __syncthreads();
    const double S_in = sm.s_in;
    MV_STAMP(3);
    S0 = S_in + ex;
    // the tile owns the output slots [n_in, n_out): known before the marks, so the common single-window tile marks without range tests
    n_in = 0;
    if (tile) n_in = max(0, mv_uniform(lean ? mv_count<true>((float)S_in, u, n, n, nfd) : mv_count<false>((float)S_in, u, n, n, nfd)));
    n_out = n;
    if (inner) n_out = mv_uniform(lean ? mv_count<true>((float)(S_in + tot), u, n, n, nfd) : mv_count<false>((float)(S_in + tot), u, n, n, nfd));
    n_out = max(n_out, n_in);
    lo_thread = n_in;
    if (tid) {
      const int32_t cnt = (gbase - 1 >= n - 1) ? n_out : (lean ? mv_count<true>((float)S0, u, n, n_out, nfd) : mv_count<false>((float)S0, u, n, n_out, nfd));
      lo_thread = max(n_in, cnt);
    }
    wb0 = n_in & ~3;
    if (mv_uniform(n_out - wb0 <= MV_WIN) && lean) {
      if (inner) mv_mark<true, false, true, true>(wq, S0, lo_thread, gbase, wb0, u, n, n_out, nfd, sm);
      else mv_mark<false, false, true, true>(wq, S0, lo_thread, gbase, wb0, u, n, n_out, nfd, sm);
    } else {
      mv_remark(winrow + gbase, wm, wiz, S0, lo_thread, gbase, wb0, true, lean, u, n, n_out, nfd, sm);
    }
    __syncthreads();
  } else {  // identity ancestors; the carried log-weights travel through the window
    if (tid == 0) atomicExch(c.mslots + (int64_t)col * T + tile, c.epoch << 54);  // (keeps the slot's tag one launch old at most)
    n_in = tile_base;
    n_out = min(tile_base + MV_TILE, n);
    wb0 = tile_base;
#pragma unroll
    for (int v = 0; v < MV_ITEMS / 4; ++v)
      *reinterpret_cast<float4*>(&sm.stage[tid * MV_ITEMS + 4 * v]) = make_float4(win[4 * v], win[4 * v + 1], win[4 * v + 2], win[4 * v + 3]);
    __syncthreads();
  }

  MV_STAMP(4);
  // ---- the move itself over the tile's output slots [n_in, n_out), window by window
  StepAcc<D> mom; mom.init();
  StepAcc1 r2; r2.init();   // APF: folded resampling weights
  StepAcc1 r3; r3.init();   // SISR: likelihood increment
  const bool observed_rt = observed, fold_rt = fold, resampled_rt = resampled;
  auto run_window = [&](auto fast_tag, int32_t wb, int32_t wlen) {
    constexpr bool FAST = decltype(fast_tag)::value;
    const bool observed = FAST ? true : observed_rt;
    const bool fold = FAST ? (ALG == SMCB_ALG_APF) : fold_rt;
    const bool resampled = FAST ? true : resampled_rt;
    for (int32_t g4 = tid * 4; g4 < wlen; g4 += MV_NT * 4) {
      const int32_t s0 = wb + g4;
      const bool full = s0 >= n_in && s0 + 4 <= n_out;
      float lwp[4] = {0.f, 0.f, 0.f, 0.f};
      // ancestors as indices inside the tile: marks are in [0, MV_TILE), "no mark yet" is -1 (only in slots of a neighbouring tile that
      // share a group of four - computed, never stored -, or in a dead column whose normalisers are NaN): the gather stays inside the
      // block's shared memory either way
      int4 aq = make_int4(g4, g4 + 1, g4 + 2, g4 + 3);
      if (resampled) aq = *reinterpret_cast<const int4*>(&sm.stage[g4]);
      else {  // weights carry over (sisr.py:52 without the reset of :34; particle/state.py:42)
        const float4 q = *reinterpret_cast<const float4*>(&sm.stage[g4]);
        lwp[0] = q.x; lwp[1] = q.y; lwp[2] = q.z; lwp[3] = q.w;
      }
      const int l[4] = {aq.x, aq.y, aq.z, aq.w};
      if (resampled ? (ALG == SMCB_ALG_SISR || a.store_lw) : (ALG == SMCB_ALG_APF && a.store_lw)) {  // sisr.py:32 / apf.py:18-23,46
        if (full) *reinterpret_cast<int4*>(pirow + s0) = make_int4(aq.x + tile_base, aq.y + tile_base, aq.z + tile_base, aq.w + tile_base);
        else {
#pragma unroll
          for (int k = 0; k < 4; ++k) if (s0 + k >= n_in && s0 + k < n_out) pirow[s0 + k] = l[k] + tile_base;
        }
      }
      float xa[D][4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int d = 0; d < D; ++d) xa[d][k] = mv_x[d * MV_TILE + l[k]];
      }
      float z[D][4];
      st_noise4<D, FAST>(a, col, (int64_t)s0, t, SMCB_RNG_TRANSITION, z);

      float xn[D][4], lwn[4], rwn[4], gnx[4], inc4[4], wprev[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float xk[D], zk[D], xo[D], inc, g_anc;
#pragma unroll
        for (int d = 0; d < D; ++d) { xk[d] = xa[d][k]; zk[d] = z[d][k]; }
        Proposal<MODEL, PROP>::sample_and_weight(y, xk, zk, Ps, observed, xo, inc, g_anc);
#pragma unroll
        for (int d = 0; d < D; ++d) xn[d][k] = xo[d];
        float lw;
        if (!observed) lw = lwp[k];
        else if (ALG == SMCB_ALG_APF) lw = __fsub_rn(inc, g_anc);   // apf.py:43
        else lw = __fadd_rn(inc, lwp[k]);                           // sisr.py:52
        lwn[k] = lw;
        inc4[k] = inc;
        wprev[k] = 0.f;
        if (ALG == SMCB_ALG_SISR) wprev[k] = resampled ? inv_n : smcb_weight(lwp[k], st_m_lw, st_inv_z_lw);
        gnx[k] = fold ? Proposal<MODEL, PROP>::pre_weight(yn, xo, Ps) : 0.f;
        rwn[k] = __fadd_rn(gnx[k], lw);
      }
      {  // nan_to_num (utils.py:57) only when something in the group is not finite: a sum of four finite floats can overflow at worst.
         // With the look-ahead folded one test serves both rows: g + lw is finite only if lw is.
        const float chk = fold ? fabsf(rwn[0]) + fabsf(rwn[1]) + fabsf(rwn[2]) + fabsf(rwn[3])
                               : fabsf(lwn[0]) + fabsf(lwn[1]) + fabsf(lwn[2]) + fabsf(lwn[3]);
        if (!(chk < INFINITY)) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            lwn[k] = st_sanitize(lwn[k]);
            rwn[k] = st_sanitize(__fadd_rn(gnx[k], lwn[k]));
          }
        }
      }
      const bool keep_lw = !fold || a.store_lw;
      if (full) {
#pragma unroll
        for (int d = 0; d < D; ++d) *reinterpret_cast<float4*>(xnext[d] + s0) = make_float4(xn[d][0], xn[d][1], xn[d][2], xn[d][3]);
        if (keep_lw) *reinterpret_cast<float4*>(lwrow + s0) = make_float4(lwn[0], lwn[1], lwn[2], lwn[3]);
        if (fold) *reinterpret_cast<float4*>(rwrow + s0) = make_float4(rwn[0], rwn[1], rwn[2], rwn[3]);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (s0 + k >= n_in && s0 + k < n_out) {
#pragma unroll
            for (int d = 0; d < D; ++d) xnext[d][s0 + k] = xn[d][k];
            if (keep_lw) lwrow[s0 + k] = lwn[k];
            if (fold) rwrow[s0 + k] = rwn[k];
          } else {  // not this tile's slot: contributes nothing
            lwn[k] = -INFINITY; rwn[k] = -INFINITY; inc4[k] = -INFINITY;
          }
        }
      }
      mom.add4(lwn, xn, shift);
      if (fold) r2.add4(rwn, one4);
      if (ALG == SMCB_ALG_SISR && observed) r3.add4(inc4, wprev);
    }
  };
  if (n_out > n_in) {
    const bool fast = observed_rt && resampled_rt && (ALG != SMCB_ALG_APF || fold_rt) && plain_noise;
    int32_t carry = -1;
    for (int32_t wb = wb0; wb < n_out; wb += MV_WIN) {
      const int32_t wlen = min(MV_WIN, n_out - wb);
      if (resampled) {
        if (wb != wb0) {  // rare: more than one window of offspring
          __syncthreads();
#pragma unroll
          for (int k = 0; k < MV_PER / 4; ++k) *reinterpret_cast<int4*>(&sm.stage[(k * MV_NT + tid) * 4]) = make_int4(-1, -1, -1, -1);
          if (tid == 0) { sm.carry = -1; if (c.wd && wb == wb0 + MV_WIN) atomicAdd((unsigned long long*)&c.wd[1], 1ull); }
          __syncthreads();
          mv_remark(winrow + gbase, wm, wiz, S0, lo_thread, gbase, wb, false, lean, u, n, n_out, nfd, sm);
          __syncthreads();
          carry = max(carry, sm.carry);
        }
        carry = mv_emit(sm, carry);
        if (wb == wb0) MV_STAMP(5);
      }
      if (fast) run_window(std::true_type{}, wb, wlen);
      else run_window(std::false_type{}, wb, wlen);
    }
  }
  MV_STAMP(6);
  pdl_trigger();  // the successor may be scheduled while the last block folds the partials

  // ---- per-tile partial record; the block that completes the column folds them
  SoftAcc<1 + 2 * D> A;
  SoftAcc<1> Q, R2, R3;
  mom.to_softacc(A, Q); r2.to_softacc(R2); r3.to_softacc(R3);
  softacc4_block_reduce<1 + 2 * D, ST_NT, false>(A, Q, R2, R3, sm.fin.f4);
  if (tid == 0) {
    Partial& p = a.partials[(int64_t)col * T + tile];
    st_write_partial1(p, A, Q);
    p.m2 = R2.m; p.z2 = R2.s[0];
    p.m3 = R3.m; p.z3 = R3.s[0];
    __threadfence();
    sm.is_last = (atomicAdd(&a.col_ticket[col], 1) == T - 1);
  }
  __syncthreads();
  if (sm.is_last) {
    if (tid == 0) a.col_ticket[col] = 0;
    __threadfence();
    finalize_column<D, OD, ALG>(a, col, FIN_STEP, t, sm.fin, sm.fin_pre);
  }
  MV_STAMP(7);
}
