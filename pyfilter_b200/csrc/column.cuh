// column_kernel: ONE block owns ONE column (one independent filter of at most RS_TILE = 4096 particles) for a whole run of moves.
//
// The multi-kernel pipeline (resample_fused_kernel -> step_kernel, step.cuh / resample.cuh) pays two launches and about ten dependent
// global round trips per move (normalisers, tile sums, tickets, partials): ~9 us per move however small the column is.  The batch of
// independent filters of SMC2 / NESS (BASELINE.json configs[4]: 4096 state particles x 1024 theta, filters/base.py:93-119) is exactly
// the regime where that floor dominates.  Here the column never leaves the SM between moves: particles in shared memory, log-weights
// in registers, normalisers in shared memory; a move is  weights -> block scan -> systematic expansion (the SAME rs_mark_pass /
// rs_emit_ancestors as resample_fused_kernel, single tile, so n_in = 0) -> gather from shared memory -> proposal / weights (the SAME
// Proposal<> and model functions as step_kernel, same Philox counters) -> block reduction -> fin_apply.  Global memory sees the
// history rows every move and the state once at the end.  The host launches it instead of the pipeline whenever it applies
// (smcb_api.cu: column_path_ok); results differ from the pipeline only through the order of the float32 reductions.
#pragma once
#include "step.cuh"
#include "resample.cuh"

struct ColumnArgs {
  StepArgs s;           // buffers, parameters, history (s.t_host = first move of this launch)
  int32_t steps;        // moves to run
  int32_t y_avail;      // observations available from y (>= steps); move k may fold the next look-ahead iff k + 1 < y_avail
  const float* y;       // (y_avail, OD): observation of move s.t_host + k at y + k * OD
  const float* u_in;    // optional injected systematic offsets (B)
  float* u_out;         // optional dump of the offsets used (B)
  float* w_out;         // optional dump of the normalised resampling weights (B, ld)
  int32_t quantize;     // must be 1: rounding-free weights
  float* xch_out;           // with an exchange attached: the block that finishes LAST on this rank also plays the reader - it polls this rank's
  int32_t* xch_ticket;      // buffer until the values of every column of the batch (all ranks) have arrived and writes them out densely
  int32_t xch_rank;         // ((2, total) floats) - no reader kernel behind the move (one launch per move in the online loop)
  int32_t preweight_first;  // APF: the look-ahead of the FIRST move was not folded into the stored weights by the previous launch (online
                            // use: y_{t+1} was unknown then) - the block evaluates apf.py:27-29 itself before the first move instead of
                            // the host launching preweight_kernel + finalize_kernel in front (two launches per move saved)
};

template <int NT, int D>
struct ColumnSmem {
  int32_t stage[RS_TILE];       // marks of the expansion, then (in place: single window) the ancestors
  double dscratch[33];          // block scan
  int32_t wtot[NT / 32];
  int32_t carry, n_out;
  float Ps[SMCB_NPARAM];
  Fin4Scratch<1 + 2 * D, NT> f4;
  ColStats st;
};

// NT threads per column, ITEMS = RS_TILE / NT consecutive particles per thread; MINB resident blocks per SM bound the registers
// (512 x 8 particles: two blocks per SM at 64 registers for large batches, one block at 128 registers when there are fewer columns
// than SMs - a lone block is latency bound and wants the instruction-level parallelism).  Threads whose particles are all padding
// (n < RS_TILE) skip the arithmetic.
template <int MODEL, int PROP, int ALG, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) column_kernel(ColumnArgs c) {
  typedef Model<MODEL> M;
  constexpr int D = M::D, OD = M::OD;
  constexpr int ITEMS = RS_TILE / NT, ROWS = ITEMS / 4;
  static_assert(ITEMS % 4 == 0 && ROWS >= 1, "a thread owns whole groups of four particles (one Philox block each)");
  __shared__ __align__(16) ColumnSmem<NT, D> cs;
  extern __shared__ __align__(16) float ck_xs[];  // (D, RS_TILE) particles of the column
  const StepArgs& a = c.s;
  const int tid = threadIdx.x, col = blockIdx.x;
  const int32_t n = (int32_t)a.n;
  const int32_t gbase = tid * ITEMS;
  int32_t* anc_s = cs.stage;
  // launched with programmatic stream serialisation: the blocks of the NEXT launch (the online loop is one launch per move) may take
  // their places while this one runs; everything a predecessor kernel writes is read behind pdl_wait()
  pdl_trigger();
  if (tid < SMCB_NPARAM) cs.Ps[tid] = a.P[(int64_t)col * SMCB_NPARAM + tid];
  pdl_wait();
  if (tid == 0) cs.st = a.stats[col];
  const float* Ps = cs.Ps;
  float ll_total = a.ll_total[col];
  const float nf = (float)a.n, inv_n = 1.0f / (float)a.n;
  const double nd = (double)n, nfd = (double)(float)a.n;
  const float one4[4] = {1.f, 1.f, 1.f, 1.f};
  const float* lwrow = a.lw + (int64_t)col * a.ld;          // rows of the state at the first move ...
  const float* rwrow = a.rw + (int64_t)col * a.ld;
  float* lwrow_out = a.lw_out + (int64_t)col * a.ld;        // ... and of the state after the last one (parity of the move index)
  float* rwrow_out = a.rw_out + (int64_t)col * a.ld;
  int32_t* pirow = a.prev_inds + (int64_t)col * a.ld;

  // ---- the column comes on chip
  float lw[ITEMS], rw[ITEMS];
  {
    const int t0 = a.t_host;
#pragma unroll
    for (int v = 0; v < ITEMS / 4; ++v) {
      const float4 q = *reinterpret_cast<const float4*>(lwrow + gbase + 4 * v);
      lw[4 * v] = q.x; lw[4 * v + 1] = q.y; lw[4 * v + 2] = q.z; lw[4 * v + 3] = q.w;
      const float4 r = *reinterpret_cast<const float4*>(rwrow + gbase + 4 * v);
      rw[4 * v] = r.x; rw[4 * v + 1] = r.y; rw[4 * v + 2] = r.z; rw[4 * v + 3] = r.w;
#pragma unroll
      for (int d = 0; d < D; ++d)
        *reinterpret_cast<float4*>(ck_xs + d * RS_TILE + gbase + 4 * v) =
            *reinterpret_cast<const float4*>(a.xbuf[t0 & 1] + ((int64_t)d * a.B + col) * a.ld + gbase + 4 * v);
    }
  }
  __syncthreads();

  if (ALG == SMCB_ALG_APF && c.preweight_first) {
    // rw = g(y_t | x_{t-1}) + lw and its normalisers, exactly what the folded look-ahead of a previous move would have left: the same
    // pre_weight function, the same per-thread accumulation order and the same block reduction (bit-identical to a batch run)
    float y0[OD];
    const bool observed0 = st_load_obs<OD>(c.y, y0);
    StepAcc1 r2; r2.init();
    if (observed0) {
#pragma unroll
      for (int g = 0; g < ITEMS / 4; ++g) {
        const int32_t i0 = gbase + 4 * g;
        if (i0 >= n) continue;
        float rwn[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float xk[D];
#pragma unroll
          for (int d = 0; d < D; ++d) xk[d] = ck_xs[d * RS_TILE + i0 + q];
          rwn[q] = st_sanitize(__fadd_rn(Proposal<MODEL, PROP>::pre_weight(y0, xk, Ps), lw[4 * g + q]));
          rw[4 * g + q] = rwn[q];
          if (i0 + q >= n) rwn[q] = -INFINITY;
        }
        r2.add4(rwn, one4);
      }
    }
    SoftAcc<1 + 2 * D> A0; A0.init();
    SoftAcc<1> Q0, R2, R30; Q0.init(); R30.init();
    r2.to_softacc(R2);
    softacc4_block_reduce<1 + 2 * D, NT, false>(A0, Q0, R2, R30, cs.f4);
    if (tid == 0) {
      ColStats stp = cs.st;
      if (observed0) {
        stp.m_rw = R2.m; stp.z_rw = R2.s[0]; stp.inv_z_rw = 1.0f / R2.s[0];
        stp.ll_aux = logf(R2.s[0]) + (R2.m - stp.m_lw) - logf(stp.z_lw);   // log sum W exp(g), W = softmax(lw)  (apf.py:44)
      }
      stp.resample = observed0 ? 1 : 0;
      stp.fold_valid = 1;
      cs.st = stp;
    }
    __syncthreads();
  }

  for (int k = 0; k < c.steps; ++k) {
    const int t = a.t_host + k;
    const ColStats st = cs.st;
    float y[OD], yn[OD];
    const bool observed = st_load_obs<OD>(c.y + (int64_t)k * OD, y);
    const bool fold = (ALG == SMCB_ALG_APF) && a.fold && (k + 1 < c.y_avail) && st_load_obs<OD>(c.y + (int64_t)(k + 1) * OD, yn);
    const bool resampled = (ALG == SMCB_ALG_APF) ? observed : (st.resample != 0);

    // ---- systematic resampling of the column (resample_fused_kernel with a single tile)
    if (resampled && st.resample) {
#pragma unroll
      for (int r = 0; r < ROWS; ++r) *reinterpret_cast<int4*>(&cs.stage[(r * NT + tid) * 4]) = make_int4(-1, -1, -1, -1);
      const float m = (ALG == SMCB_ALG_APF) ? st.m_rw : st.m_lw, iz = (ALG == SMCB_ALG_APF) ? st.inv_z_rw : st.inv_z_lw;
      float w[ITEMS];
      double tsum = 0.0;
      const bool tlive = gbase < n;   // some of this thread's particles exist
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) w[j] = 0.f;
      if (tlive) {
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
          float x = smcb_weight((ALG == SMCB_ALG_APF) ? rw[j] : lw[j], m, iz);
          if (gbase + j >= n) x = 0.f;
          const double xd = __dadd_rn(__dadd_rn(1.0, (double)x), -1.0);  // multiple of 2^-52: the column is benign
          w[j] = (float)xd;
          tsum += xd;
        }
      }
      if (c.w_out) {
        float* dst = c.w_out + (int64_t)col * a.ld + gbase;
#pragma unroll
        for (int v = 0; v < ITEMS / 4; ++v)
          reinterpret_cast<float4*>(dst)[v] = make_float4(w[4 * v], w[4 * v + 1], w[4 * v + 2], w[4 * v + 3]);
      }
      double tot;
      const double S0 = rs_block_excl_scan_d<NT, false>(tsum, cs.dscratch, &tot);   // (dscratch was last read several barriers ago)
      float u;
      if (c.u_in) u = c.u_in[col];
      else {
        const Philox4 r4 = philox4x32_10((uint32_t)(col + a.col0), 0u, (uint32_t)t, SMCB_RNG_SYSTEMATIC, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
        u = smcb_u01(r4.x);
      }
      if (c.u_out && tid == 0) c.u_out[col] = u;
      // counts of this thread's particles (exact_scan.h: xs_count_lean; the column is one window, so a mark needs no range test) and
      // the "last mark" scan in place: the ancestors replace the marks
      int32_t lo = 0;
      if (tid) lo = (gbase - 1 >= n - 1) ? n : max(0, min(xs_count_lean((float)S0, u, n, nfd), n));
      if (tlive) {
        double run = S0;
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
          run = __dadd_rn(run, (double)w[j]);
          int32_t hi = min(xs_count_lean((float)run, u, n, nfd), n);
          hi = (gbase + j >= n - 1) ? n : hi;   // cumsum[..., -1] = 1.0 (resampling.py:49): every probe is <= 1
          if (hi > lo) cs.stage[lo] = gbase + j;
          lo = max(lo, hi);
        }
      }
      __syncthreads();
      rs_emit_blocked<NT, ITEMS>(cs, -1);
    }

    // ---- the move itself (step_kernel's body on this thread's ITEMS particles)
    StepAcc<D> mom; mom.init();
    StepAcc1 r2; r2.init();
    StepAcc1 r3; r3.init();
    float shift[D];
#pragma unroll
    for (int d = 0; d < D; ++d) shift[d] = st.shift[d];
    float xnew[D][ITEMS];
    // the steady state (observed, resampled, look-ahead folded, plain Philox noise) runs a copy of the loop with those facts as
    // compile-time constants (same reason as in step_kernel: the uniform per-particle branches otherwise become convergence regions)
    const bool observed_rt = observed, fold_rt = fold, resampled_rt = resampled;
    auto groups = [&](auto fast_tag) {
      constexpr bool FAST = decltype(fast_tag)::value;
      const bool observed = FAST ? true : observed_rt;
      const bool fold = FAST ? (ALG == SMCB_ALG_APF) : fold_rt;
      const bool resampled = FAST ? true : resampled_rt;
#pragma unroll
      for (int g = 0; g < ITEMS / 4; ++g) {
        const int32_t i0 = gbase + 4 * g;
        int anc[4] = {i0, i0 + 1, i0 + 2, i0 + 3};
        if (resampled) {
          const int4 q = *reinterpret_cast<const int4*>(anc_s + i0);
          anc[0] = q.x; anc[1] = q.y; anc[2] = q.z; anc[3] = q.w;
        }
        const bool live = i0 < n, full = i0 + 4 <= n;
        if (!live) {  // padding only
#pragma unroll
          for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int d = 0; d < D; ++d) xnew[d][4 * g + q] = 0.f;
          }
          continue;
        }
        if (resampled || ALG == SMCB_ALG_APF)   // sisr.py:32 / apf.py:18-23
          *reinterpret_cast<int4*>(pirow + i0) = make_int4(anc[0], anc[1], anc[2], anc[3]);
        if (!full) {
#pragma unroll
          for (int q = 0; q < 4; ++q) if (i0 + q >= n) anc[q] = 0;
        }
        float lwp[4] = {0.f, 0.f, 0.f, 0.f};
        if (!resampled) {
#pragma unroll
          for (int q = 0; q < 4; ++q) lwp[q] = lw[4 * g + q];
        }
        float xa[D][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
          for (int d = 0; d < D; ++d) xa[d][q] = ck_xs[d * RS_TILE + anc[q]];
        }
        float z[D][4];
        st_noise4<D, FAST>(a, col, i0, t, SMCB_RNG_TRANSITION, z);
        float xn[D][4], lwn[4], rwn[4], gnx[4], inc4[4], wprev[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float xk[D], zk[D], xo[D], inc, g_anc;
#pragma unroll
          for (int d = 0; d < D; ++d) { xk[d] = xa[d][q]; zk[d] = z[d][q]; }
          prop_sample_and_weight<MODEL, PROP>(a, col, (int64_t)i0 + q, t, y, xk, zk, Ps, observed, xo, inc, g_anc);
#pragma unroll
          for (int d = 0; d < D; ++d) xn[d][q] = xo[d];
          float lwv;
          if (!observed) lwv = lwp[q];
          else if (ALG == SMCB_ALG_APF) lwv = __fsub_rn(inc, g_anc);   // apf.py:43
          else lwv = __fadd_rn(inc, lwp[q]);                           // sisr.py:52
          lwn[q] = lwv;
          inc4[q] = inc;
          wprev[q] = 0.f;
          if (ALG == SMCB_ALG_SISR) wprev[q] = resampled ? inv_n : smcb_weight(lwp[q], st.m_lw, st.inv_z_lw);
          gnx[q] = fold ? Proposal<MODEL, PROP>::pre_weight(yn, xo, Ps) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          lwn[q] = st_sanitize(lwn[q]);                                // utils.py:57
          rwn[q] = fold ? st_sanitize(__fadd_rn(gnx[q], lwn[q])) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          lw[4 * g + q] = lwn[q];
          if (fold) rw[4 * g + q] = rwn[q];
#pragma unroll
          for (int d = 0; d < D; ++d) xnew[d][4 * g + q] = xn[d][q];
        }
        if (!full) {  // padding contributes nothing
#pragma unroll
          for (int q = 0; q < 4; ++q) if (i0 + q >= n) { lwn[q] = -INFINITY; rwn[q] = -INFINITY; inc4[q] = -INFINITY; }
        }
        mom.add4(lwn, xn, shift);
        if (fold) r2.add4(rwn, one4);
        if (ALG == SMCB_ALG_SISR && observed) r3.add4(inc4, wprev);
      }
    };
    if (observed_rt && resampled_rt && (ALG != SMCB_ALG_APF || fold_rt) && !a.eps_in && !a.eps_out) groups(std::true_type{});
    else groups(std::false_type{});
    SoftAcc<1 + 2 * D> A;
    SoftAcc<1> Q, R2, R3;
    mom.to_softacc(A, Q); r2.to_softacc(R2); r3.to_softacc(R3);
    softacc4_block_reduce<1 + 2 * D, NT, false>(A, Q, R2, R3, cs.f4);   // its barriers also separate this move's gathers from the stores below
    if (tid == 0) {
      FinPre pre;
      pre.st = st; pre.observed = observed; pre.fold = fold; pre.ll_total = ll_total;
      ColStats stn = st;
      const float ll = fin_apply<D, OD, ALG>(a, col, FIN_STEP, t, A, Q, R2, R3, pre, stn);
      ll_total += ll;
      cs.st = stn;
      if (k == c.steps - 1) smcb_exchange_publish(a.xch, col, ll, ll_total);   // the values of the launch's last move go to every rank
    }
#pragma unroll
    for (int v = 0; v < ITEMS / 4; ++v) {
#pragma unroll
      for (int d = 0; d < D; ++d)
        *reinterpret_cast<float4*>(ck_xs + d * RS_TILE + gbase + 4 * v) =
            make_float4(xnew[d][4 * v], xnew[d][4 * v + 1], xnew[d][4 * v + 2], xnew[d][4 * v + 3]);
    }
    __syncthreads();
  }

  // ---- the column goes back to global memory
  {
    const int t1 = a.t_host + c.steps;
    const bool rw_valid = (ALG == SMCB_ALG_APF) && cs.st.fold_valid;
#pragma unroll
    for (int v = 0; v < ITEMS / 4; ++v) {
      *reinterpret_cast<float4*>(lwrow_out + gbase + 4 * v) = make_float4(lw[4 * v], lw[4 * v + 1], lw[4 * v + 2], lw[4 * v + 3]);
      if (rw_valid) *reinterpret_cast<float4*>(rwrow_out + gbase + 4 * v) = make_float4(rw[4 * v], rw[4 * v + 1], rw[4 * v + 2], rw[4 * v + 3]);
#pragma unroll
      for (int d = 0; d < D; ++d)
        *reinterpret_cast<float4*>(a.xbuf[t1 & 1] + ((int64_t)d * a.B + col) * a.ld + gbase + 4 * v) =
            *reinterpret_cast<const float4*>(ck_xs + d * RS_TILE + gbase + 4 * v);
    }
    if (col == 0 && tid == 0) a.ctrl->t = t1;
  }
  if (a.xch.seq && c.xch_out) {   // the last block of this rank gathers the exchange (its own rank's values were published above)
    __shared__ int last_block;
    if (tid == 0) {
      const int old = atomicAdd(c.xch_ticket, 1);
      last_block = (old == (int)gridDim.x - 1);
      if (last_block) *c.xch_ticket = 0;
    }
    __syncthreads();
    if (last_block) {
      const unsigned long long* base = a.xch.peer[c.xch_rank] + (int64_t)(a.xch.seq & 1u) * 2 * a.xch.total;
      for (int i = tid; i < 2 * a.xch.total; i += NT) {
        unsigned long long v;
        for (;;) {
          asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(base + i) : "memory");
          if ((uint32_t)(v >> 32) == a.xch.seq) break;
          __nanosleep(100);
        }
        c.xch_out[i] = __uint_as_float((uint32_t)v);
      }
    }
  }
}
