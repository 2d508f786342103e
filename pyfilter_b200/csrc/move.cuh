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

#ifndef MV_NT
#define MV_NT 256
#endif
#define MV_ITEMS 16
#define MV_TILE (MV_NT * MV_ITEMS)            // == RS_TILE: the row pitch is a multiple of it
#define MV_PER 20                             // window slots per thread in the "last mark" scan (largest tile class)
#define MV_WIN (MV_NT * MV_PER)               // 5120 output slots per window
// Tile classes.  A tile is MV_NT threads x ITEMS consecutive particles with ITEMS in {16, 4}; its window holds MV_NT * (ITEMS + 4)
// output slots.  A column is cut into t1 tiles of items1 particles per thread followed by tiles of items2 <= items1 (MoveArgs).  Measured
// (tools/geom_sweep.py, profiles/README.md): a tile costs a fixed part worth about 8 particles per thread, and the machine behaves like a
// throughput device once it is full - 4,000,000 particles as 977 tiles of 4096 take 40.6 us, as 592 x 4096 + 513 x 3072 ("two full waves")
// 42.4 us, as 1303 x 3072 47.2 us - so the host keeps 4096-particle tiles and only cuts the remainder of a column finer when the last
// wave would leave more than three quarters of the block slots idle (2,000,000 x 3-D particles: 489 tiles on 444 slots, 38.2 -> 36.8 us).
#define MV_WIN_OF(items) (MV_NT * ((items) + 4))
#define MV_U_HOST 8
#define MV_LB 4                               // look-back: slots per thread in flight
// resident blocks per SM the register budget is sized for.  Scalar state (37 KB of shared memory per block): 5 blocks of 51 registers
// measured 2 % faster than 4 of 64 and 13 % faster than 3 of 85 (tools/variants.py: 39.9 / 40.7 / 46.0 us per 4M-particle move) - the
// kernel hides latency with warps, not with registers.  Three state dimensions: 70 KB per block, three blocks fit whatever the registers.
#ifndef SMCB_MV_MINB
#define SMCB_MV_MINB 5
#endif
#ifndef SMCB_MV_MINB3
#define SMCB_MV_MINB3 3
#endif
static_assert(RS_TILE % MV_TILE == 0, "the row pitch (a multiple of RS_TILE) is a multiple of the largest tile");

struct MoveArgs {
  StepArgs s;                 // buffers, parameters, history; s.partials = per-tile records (B, tiles_per_col), s.blocks_per_col = tiles_per_col
  int32_t tiles_per_col;      // t1 + (tiles of the second class)
  int32_t total_tiles;        // B * tiles_per_col
  int32_t t1, items1, items2; // tile classes of a column: tiles [0, t1) hold MV_NT * items1 particles each, the others MV_NT * items2
  uint32_t ticket_base;       // value of *tile_counter when this launch starts (the counter is never reset: wrap-around arithmetic)
  uint32_t* tile_counter;
  unsigned long long* mslots; // (B, tiles_per_col) tile sums tagged with the launch epoch, ONE 64-bit word each: a sum of weights that are
                              // multiples of 2^-52 below 2 is a 54-bit integer count of quanta, the 10 bits above it carry the tag
  unsigned long long* gwords; // (2, B, ceil(tiles_per_col / 32) * MV_GPAD) group words (member count << 58 | quanta), two sets: launch k
                              // uses set k & 1 and clears the other one (both zero after smcb_filter_initialize)
  uint32_t launch_index;      // move launches since the handle's state was initialised
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
  unsigned long long lb;      // look-back: quanta of the tiles before this one (warp 0 -> everybody)
  int32_t wtot[MV_NT / 32];
  int32_t carry;
  int32_t ticket, is_last;
  float u;
  float Ps[SMCB_NPARAM];
  Fin4Scratch<1 + 2 * D, MV_NT> f4;
  FinPre fin_pre;
};

// a block-uniform flag / value the compiler can SEE is uniform (vote / redux result): branches on it need no divergence handling
// and the collectives behind them no re-convergence
#define MV_STAMP(k) do { if (c.tl && tid == 0) c.tl[(int64_t)ticket * 16 + (k)] = st_now(); } while (0)
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
// ---- publishing a tile sum and looking back, two levels -------------------------------------------------------------------------
// A tile sum is a count of quanta (2^-52) in 54 bits.  Tiles form groups of 32.  A tile publishes twice, with two fire-and-forget
// memory operations and NO return value to wait for: (1) its own slot = (launch tag << 54) | quanta, one 64-bit store; (2) a 64-bit
// reduction  (1 << 58) | quanta  onto its group's word - the word counts its members in the top six bits and adds their quanta up below, so
// "the group total is complete" is readable from the word itself.  A tile then needs one group word per earlier group and the slots of its
// group mates before it: two loads per lane of ONE warp for up to 1024 tiles, all issued right behind the publication - the look-back of
// a tile whose predecessors are done is one L2 round trip.  The group words exist twice; launch k uses set k & 1 and the first tile of
// every group clears the other set's word for launch k + 1 (every block of launch k - 1, the last user of that set, had finished before
// launch k passed its grid dependency).  (Measured on the way: a counter whose old value picks a "closing" tile that adds the group up
// and publishes the total costs three dependent round trips, 4-5 us per tile; 256 polling threads per block, or a warp polling every
// predecessor, flood the L2 with strong loads and delay the very publications they wait for - up to 20 us per tile.)
#define MV_SLOT_MASK ((1ull << 54) - 1ull)
#define MV_GROUP 32
#define MV_GPAD 16   // a group's word sits alone in a 128-byte line: hundreds of tiles poll and bump them at the same time
#define MV_GCOUNT_SHIFT 58
__device__ __forceinline__ unsigned long long mv_ld_strong(const unsigned long long* p) {  // a STRONG load at gpu scope
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mv_st_strong(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void mv_red_add(unsigned long long* p, unsigned long long v) {
  asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long mv_wait_slot(const unsigned long long* p, unsigned long long v, unsigned long long epoch, int& waited) {
  while ((v >> 54) != epoch) {
#ifdef MV_SLEEP
    __nanosleep(MV_SLEEP);
#endif
    v = mv_ld_strong(p);
    ++waited;
  }
  return v & MV_SLOT_MASK;
}
__device__ __forceinline__ unsigned long long mv_wait_group(const unsigned long long* p, unsigned long long v, int& waited) {
  while ((v >> MV_GCOUNT_SHIFT) != (unsigned long long)MV_GROUP) {   // every group BEFORE a tile's own is full
#ifdef MV_SLEEP
    __nanosleep(MV_SLEEP);
#endif
    v = mv_ld_strong(p);
    ++waited;
  }
  return v & MV_SLOT_MASK;
}
__device__ __forceinline__ unsigned long long mv_warp_sum_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// executed by warp 0 of the block; `quanta` = this tile's sum (valid in lane 0); returns the sum of all tiles before `tile` (every lane).
// `gw` = this launch's set of group words of the column, `gw_next` = the other set.
__device__ __forceinline__ unsigned long long mv_publish_lookback(unsigned long long* slots, unsigned long long* gw, unsigned long long* gw_next,
                                                                   int tile, unsigned long long quanta, unsigned long long epoch,
                                                                   long long* wd, long long* tlrow) {
  const int lane = threadIdx.x & 31;
#define MV_LSTAMP(k) do { if (tlrow && lane == 0) tlrow[k] = st_now(); } while (0)
  const int g = tile / MV_GROUP, r = tile - g * MV_GROUP;
  if (lane == 0) {
    mv_st_strong(slots + tile, (epoch << 54) | (quanta & MV_SLOT_MASK));
    mv_red_add(gw + (int64_t)g * MV_GPAD, (1ull << MV_GCOUNT_SHIFT) | (quanta & MV_SLOT_MASK));
    if (r == 0) mv_st_strong(gw_next + (int64_t)g * MV_GPAD, 0ull);
  }
  MV_LSTAMP(8);
  const unsigned long long* pm = slots + g * MV_GROUP + lane;
  unsigned long long vm = (lane < r) ? mv_ld_strong(pm) : (epoch << 54);
  unsigned long long acc = 0ull;
  int waited = 0;
  for (int g0 = 0; g0 < g; g0 += 32) {   // one round per 32 earlier groups (1024 tiles)
    const unsigned long long* pg = gw + (int64_t)(g0 + lane) * MV_GPAD;
    const unsigned long long vg = (g0 + lane < g) ? mv_ld_strong(pg) : ((unsigned long long)MV_GROUP << MV_GCOUNT_SHIFT);
    acc += mv_wait_group(pg, vg, waited);
  }
  MV_LSTAMP(11);
  acc += mv_wait_slot(pm, vm, epoch, waited);
  MV_LSTAMP(12);
  if (tlrow && lane == 0) tlrow[13] = (long long)waited;
  if (tlrow && wd && waited) atomicAdd((unsigned long long*)&wd[2], (unsigned long long)waited);   // diagnostics: re-polls
  return mv_warp_sum_u64(acc);
}
// a tile that does not resample only keeps the group words in step: the other set is cleared for the next launch
__device__ __forceinline__ void mv_publish_idle(unsigned long long* gw_next, int tile) {
  const int g = tile / MV_GROUP;
  if (tile - g * MV_GROUP == 0) mv_st_strong(gw_next + (int64_t)g * MV_GPAD, 0ull);
}

// weights of this thread's 16 particles, rounded to multiples of 2^-52 (exactly what normalize_kernel / resample_fused_kernel compute)
template <bool INNER, int ITEMS>
__device__ __forceinline__ double mv_weights(const float (&win)[ITEMS], float m, float iz, int32_t gbase, int32_t n, float (&wq)[ITEMS]) {
  double tsum = 0.0;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    float x = smcb_weight(win[j], m, iz);
    if (!INNER && gbase + j >= n) x = 0.f;
    const double xd = __dadd_rn(__dadd_rn(1.0, (double)x), -1.0);
    wq[j] = (float)xd;   // exact: a multiple of 2^-52 below 2^-29 has at most 23 significant bits (kept as float: registers)
    tsum += xd;
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
template <bool INNER, bool CHECK, bool FIRST, bool LEAN, int ITEMS, typename SM>
__device__ __forceinline__ void mv_mark(const float (&wq)[ITEMS], double S0, int32_t lo, int32_t gbase, int32_t wb, float u,
                                        int32_t n, int32_t n_out, double nfd, SM& sm) {
  double run = S0;
  const int32_t lbase = (int32_t)threadIdx.x * ITEMS;
  constexpr int32_t WIN = MV_WIN_OF(ITEMS);
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    run = __dadd_rn(run, (double)wq[j]);
    int32_t hi = mv_count<LEAN>((float)run, u, n, n_out, nfd);
    if (!INNER) hi = (gbase + j >= n - 1) ? n_out : hi;  // cumsum[..., -1] = 1.0 (resampling.py:49): every probe is <= 1 (n_out == n here)
    const int32_t r = lo - wb;
    if (CHECK) {
      if (hi > lo && (uint32_t)r < (uint32_t)WIN) sm.stage[r] = lbase + j;
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
template <int ITEMS, typename SM>
__device__ __noinline__ void mv_remark(const float* wsrc, float m, float iz, double S0, int32_t lo, int32_t gbase, int32_t wb, bool first,
                                       bool lean, float u, int32_t n, int32_t n_out, double nfd, SM& sm) {
  float win[ITEMS];
#pragma unroll
  for (int v = 0; v < ITEMS / 4; ++v) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(wsrc) + v);
    win[4 * v] = q.x; win[4 * v + 1] = q.y; win[4 * v + 2] = q.z; win[4 * v + 3] = q.w;
  }
  float wq[ITEMS];
  mv_weights<false, ITEMS>(win, m, iz, gbase, n, wq);
  if (lean) {
    if (first) mv_mark<false, true, true, true, ITEMS>(wq, S0, lo, gbase, wb, u, n, n_out, nfd, sm);
    else mv_mark<false, true, false, true, ITEMS>(wq, S0, lo, gbase, wb, u, n, n_out, nfd, sm);
  } else {
    if (first) mv_mark<false, true, true, false, ITEMS>(wq, S0, lo, gbase, wb, u, n, n_out, nfd, sm);
    else mv_mark<false, true, false, false, ITEMS>(wq, S0, lo, gbase, wb, u, n, n_out, nfd, sm);
  }
}
template <typename SM>
__device__ __forceinline__ void mv_remark_any(int items, const float* wsrc, float m, float iz, double S0, int32_t lo, int32_t gbase, int32_t wb,
                                              bool first, bool lean, float u, int32_t n, int32_t n_out, double nfd, SM& sm) {
  if (items == 16) mv_remark<16>(wsrc, m, iz, S0, lo, gbase, wb, first, lean, u, n, n_out, nfd, sm);
  else mv_remark<4>(wsrc, m, iz, S0, lo, gbase, wb, first, lean, u, n, n_out, nfd, sm);
}
template <typename SM>
__device__ __forceinline__ int32_t mv_emit_any(int items, SM& sm, int32_t carry) {
  if (items == 16) return rs_emit_blocked<MV_NT, 20>(sm, carry);
  return rs_emit_blocked<MV_NT, 8>(sm, carry);
}

template <int MODEL, int PROP, int ALG>
__global__ void __launch_bounds__(MV_NT, (Model<MODEL>::D == 1 ? SMCB_MV_MINB : SMCB_MV_MINB3)) move_kernel(MoveArgs c) {
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
  bool observed, fold;
  {  // only the flags now; the values are fetched again behind the marks (registers are scarce while the weights are live)
    float y0[OD], y1[OD];
    observed = mv_uniform(st_load_obs<OD>(a.y_t, y0));
    fold = mv_uniform((ALG == SMCB_ALG_APF) && a.fold && st_load_obs<OD>(a.y_next, y1));
  }
  const int32_t n = (int32_t)a.n;
  const double nfd = (double)(float)a.n;
  __syncthreads();
  const int ticket = mv_uniform(sm.ticket);
  const int col = ticket / T, tile = ticket - col * T;
  const float u = sm.u;
  if (tid < SMCB_NPARAM) sm.Ps[tid] = a.P[(int64_t)col * SMCB_NPARAM + tid];  // (first read behind several barriers)
  // tile class: particles per thread, first particle of the tile, slots per window
  const int items = mv_uniform(tile < c.t1 ? c.items1 : c.items2);
  const int32_t tile_base = (tile < c.t1) ? tile * (MV_NT * c.items1) : c.t1 * (MV_NT * c.items1) + (tile - c.t1) * (MV_NT * c.items2);
  const int32_t tile_len = MV_NT * items;
  const int32_t win_slots = MV_NT * (items + 4);
  const int32_t gbase = tile_base + tid * items;
  const int64_t rowoff = (int64_t)col * a.ld;
  const bool use_rw = (ALG == SMCB_ALG_APF) && observed;
  const float* const winrow = (use_rw ? a.rw : a.lw) + rowoff;
  // the lean probe count needs u == 0 or u >= 2^-64 (exact_scan.h); Philox offsets are multiples of 2^-24
  const bool lean = mv_uniform((u == 0.f) || (u >= 5.5e-20f && u < 1.0f));

  MV_STAMP(0);
  pdl_wait();
  MV_STAMP(1);
  // ---- the tile comes on chip: x_{t-1} (striped), the column's normalisers (every thread, one broadcast transaction per warp), then
  //      (per tile class) the resampling log-weights: blocked, a thread owns `items` consecutive particles.  Rows are allocated with a
  //      tile of slack behind them: the last tile of a column may reach past the row (its surplus particles have index >= n: weight 0)
  const float4* stp = reinterpret_cast<const float4*>(a.stats + col);
  const float4 st0 = stp[0], st1 = stp[1];   // m_lw z_lw inv_z_lw m_rw | z_rw inv_z_rw ess resample
#pragma unroll
  for (int d = 0; d < D; ++d) {
#pragma unroll
    for (int v = 0; v < MV_ITEMS / 4; ++v) {
      const int e = (v * MV_NT + tid) * 4;
      if (e < tile_len)
        *reinterpret_cast<float4*>(mv_x + d * MV_TILE + e) =
            __ldg(reinterpret_cast<const float4*>(a.xbuf[t & 1] + ((int64_t)d * a.B + col) * a.ld + tile_base + e));
    }
  }
  if (tid < (int)(sizeof(ColStats) / 4)) reinterpret_cast<float*>(&sm.fin_pre.st)[tid] = reinterpret_cast<const float*>(a.stats + col)[tid];
  if (tid == 32) { sm.fin_pre.observed = observed; sm.fin_pre.fold = fold; sm.fin_pre.ll_total = a.ll_total[col]; }
  // SISR resamples when the ESS test fired (sisr.py:19-26), the APF on every observed step (apf.py:29-34, filters/base.py:213)
  const bool resampled = (ALG == SMCB_ALG_APF) ? observed : mv_uniform(__float_as_int(st1.w) != 0);
  const float wm = use_rw ? st0.w : st0.x;
  const float wiz = use_rw ? st1.y : st0.z;
  if (c.u_out && resampled && tile == 0 && tid == 0) c.u_out[col] = u;

  int32_t n_in = 0, n_out = 0, wb0 = 0;
  double S0 = 0.0;
  int32_t lo_thread = 0;
  auto front = [&](auto items_tag) {
    constexpr int ITEMS = decltype(items_tag)::value;
    constexpr int32_t WIN = MV_WIN_OF(ITEMS);
    float win[ITEMS];
#pragma unroll
    for (int v = 0; v < ITEMS / 4; ++v) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(winrow + gbase) + v);
      win[4 * v] = q.x; win[4 * v + 1] = q.y; win[4 * v + 2] = q.z; win[4 * v + 3] = q.w;
    }
    if (resampled) {
      const bool inner = mv_uniform(tile_base + MV_NT * ITEMS <= n - 1);  // neither padding nor the last particle of the column in this tile
      float wq[ITEMS];
      const double tsum = inner ? mv_weights<true, ITEMS>(win, wm, wiz, gbase, n, wq) : mv_weights<false, ITEMS>(win, wm, wiz, gbase, n, wq);
      if (c.w_out) {
        float* dst = c.w_out + rowoff + gbase;
#pragma unroll
        for (int v = 0; v < ITEMS / 4; ++v)
          if (gbase + 4 * v + 4 <= a.ld) reinterpret_cast<float4*>(dst)[v] = make_float4(wq[4 * v], wq[4 * v + 1], wq[4 * v + 2], wq[4 * v + 3]);
      }
      double tot;
      const double ex = mv_block_excl_scan(tsum, sm.scan_a, &tot);
      const int G = (T + MV_GROUP - 1) / MV_GROUP;
      MV_STAMP(2);
      if (tid < 32) {
        unsigned long long* const gw0 = c.gwords + (int64_t)col * G * MV_GPAD;
        const int64_t gset = (int64_t)a.B * G * MV_GPAD;
        const unsigned long long before = mv_publish_lookback(c.mslots + (int64_t)col * T, gw0 + (c.launch_index & 1u) * gset, gw0 + ((c.launch_index + 1u) & 1u) * gset,
                                                              tile, __double2ull_rn(tot * 4503599627370496.0), c.epoch, c.wd, c.tl ? c.tl + (int64_t)ticket * 16 : nullptr);
        if (tid == 0) sm.lb = before;
      }
      __syncthreads();
      const double S_in = (double)sm.lb * 2.220446049250313e-16;   // quanta * 2^-52: exact (the sum of a column is below 2)
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
      if (mv_uniform(n_out - wb0 <= WIN) && lean) {
        if (inner) mv_mark<true, false, true, true, ITEMS>(wq, S0, lo_thread, gbase, wb0, u, n, n_out, nfd, sm);
        else mv_mark<false, false, true, true, ITEMS>(wq, S0, lo_thread, gbase, wb0, u, n, n_out, nfd, sm);
      } else {
        mv_remark<ITEMS>(winrow + gbase, wm, wiz, S0, lo_thread, gbase, wb0, true, lean, u, n, n_out, nfd, sm);
      }
      __syncthreads();
    } else {  // identity ancestors; the carried log-weights travel through the window
      if (tid == 0) {
        const int G = (T + MV_GROUP - 1) / MV_GROUP;
        mv_publish_idle(c.gwords + (int64_t)col * G * MV_GPAD + ((c.launch_index + 1u) & 1u) * ((int64_t)a.B * G * MV_GPAD), tile);
      }
      n_in = tile_base;
      n_out = max(n_in, min(tile_base + MV_NT * ITEMS, n));
      wb0 = tile_base;
#pragma unroll
      for (int v = 0; v < ITEMS / 4; ++v)
        *reinterpret_cast<float4*>(&sm.stage[tid * ITEMS + 4 * v]) = make_float4(win[4 * v], win[4 * v + 1], win[4 * v + 2], win[4 * v + 3]);
      __syncthreads();
    }
  };
  if (items == 16) front(std::integral_constant<int, 16>{});   // (two classes are compiled: every further one costs instruction-cache
  else front(std::integral_constant<int, 4>{});                // misses in all of them - 12 and 8 were measured and never chosen)

  MV_STAMP(4);
  // ---- the move itself over the tile's output slots [n_in, n_out), window by window.  What only this phase needs is fetched now.
  float y[OD], yn[OD];
  st_load_obs<OD>(a.y_t, y);
  if (fold) st_load_obs<OD>(a.y_next, yn);
  const float* Ps = sm.Ps;
  const float inv_n = 1.0f / (float)a.n;
  const float one4[4] = {1.f, 1.f, 1.f, 1.f};
  const bool plain_noise = !a.eps_in && !a.eps_out;
  const ColStats& stc = sm.fin_pre.st;   // the column's statistics before the move (shared-memory copy made with the tile loads)
  const float st_m_lw = stc.m_lw, st_inv_z_lw = stc.inv_z_lw;
  float shift[D];
#pragma unroll
  for (int d = 0; d < D; ++d) shift[d] = stc.shift[d];
  float* const lwrow = a.lw_out + rowoff;
  float* const rwrow = a.rw_out + rowoff;
  int32_t* const pirow = a.prev_inds + rowoff;
  float* xnext[D];
#pragma unroll
  for (int d = 0; d < D; ++d) xnext[d] = a.xbuf[(t + 1) & 1] + ((int64_t)d * a.B + col) * a.ld;
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
        prop_sample_and_weight<MODEL, PROP>(a, col, (int64_t)s0 + k, t, y, xk, zk, Ps, observed, xo, inc, g_anc);
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
          } else {  // not this tile's slot: contributes nothing (its "ancestor" may be an unset mark pointing at shared memory nobody wrote:
                    // the weight is zero, but 0 * NaN would still poison the moment sums)
            lwn[k] = -INFINITY; rwn[k] = -INFINITY; inc4[k] = -INFINITY;
#pragma unroll
            for (int d = 0; d < D; ++d) xn[d][k] = 0.f;
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
    for (int32_t wb = wb0; wb < n_out; wb += win_slots) {
      const int32_t wlen = min(win_slots, n_out - wb);
      if (resampled) {
        if (wb != wb0) {  // rare: more than one window of offspring
          __syncthreads();
#pragma unroll
          for (int k = 0; k < MV_PER / 4; ++k) *reinterpret_cast<int4*>(&sm.stage[(k * MV_NT + tid) * 4]) = make_int4(-1, -1, -1, -1);
          if (tid == 0) { sm.carry = -1; if (c.wd && wb == wb0 + win_slots) atomicAdd((unsigned long long*)&c.wd[1], 1ull); }
          __syncthreads();
          mv_remark_any(items, winrow + gbase, wm, wiz, S0, lo_thread, gbase, wb, false, lean, u, n, n_out, nfd, sm);
          __syncthreads();
          carry = max(carry, sm.carry);
        }
        carry = mv_emit_any(items, sm, carry);
        if (wb == wb0) MV_STAMP(5);
      }
      if (fast) run_window(std::true_type{}, wb, wlen);
      else run_window(std::false_type{}, wb, wlen);
    }
  }
  MV_STAMP(6);
  pdl_trigger();  // the successor may be scheduled while the last block folds the partials

  // ---- per-tile partial record.  The records are folded by finalize_kernel, launched right behind this kernel (programmatic
  //      dependent launch): a block neither fences nor takes a ticket before it leaves (measured: 1.5 us per block)
  SoftAcc<1 + 2 * D> A;
  SoftAcc<1> Q, R2, R3;
  mom.to_softacc(A, Q); r2.to_softacc(R2); r3.to_softacc(R3);
  softacc4_block_reduce<1 + 2 * D, MV_NT, false>(A, Q, R2, R3, sm.f4);
  if (tid == 0) {
    Partial& p = a.partials[(int64_t)col * T + tile];
    st_write_partial1(p, A, Q);
    p.m2 = R2.m; p.z2 = R2.s[0];
    p.m3 = R3.m; p.z3 = R3.s[0];
  }
  MV_STAMP(7);
}
