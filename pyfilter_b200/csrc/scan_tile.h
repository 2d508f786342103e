// Per-thread phases of the exact tile scan (see exact_scan.h for the arithmetic and DESIGN.md for the kernel structure).
//
// A tile is TILE = NT * ITEMS consecutive weights of one column; thread t owns elements [t*ITEMS, (t+1)*ITEMS).
//   phase A  approximate (fp64, any association) prefix at every thread boundary -> binade label at the start and end of
//            each thread.  A thread whose two labels agree is "clean": all its elements are regular in that binade.  In a
//            "dirty" thread the label of every element is derived; an element whose label differs from its predecessor's is
//            "special" (speculated binade crossing) and opens a new segment.
//   phase B  every regular element becomes a transducer in the quantum of its segment; a segmented block scan composes them.
//   phase C  one thread walks the (few) segments with the EXACT incoming state: special elements are applied with a genuine
//            IEEE add, segment aggregates with xs_apply; every speculation is verified on the way.
//   phase D  every thread turns (segment base state, transducer prefix) into the exact S_k, then c_k = fl32(S_k).
// The functions below are the thread-local parts; the block-wide scans live in the kernel (and in tests/host/ serially).
#pragma once
#include "exact_scan.h"

// Segmented-scan payload: number of specials seen, and the transducer of the OPEN segment (regular elements since the last
// special, or since the range start when cnt == 0).
struct XsSeg {
  XsT t;
  int32_t cnt;
};
XS_HD XsSeg xs_seg_identity() { XsSeg s; s.t = xs_identity(); s.cnt = 0; return s; }
// a THEN b; E = binade label at the start of b's range (only used when b has no special, i.e. the open segment continues)
template <int MB>
XS_HD XsSeg xs_seg_combine(const XsSeg& a, const XsSeg& b, int E) {
  XsSeg r;
  r.cnt = a.cnt + b.cnt;
  r.t = b.cnt ? b.t : xs_compose<MB>(a.t, b.t, E);
  return r;
}

// Tile descriptor published for the decoupled look-back (valid when the tile holds at most one special element).
struct XsDesc {
  double a_s;       // segment 0 aggregate
  double b_s;       // segment 1 aggregate (after the special), if any
  float wc;         // the special element
  int16_t e0;       // speculated label of the incoming state
  int16_t e1;       // speculated label right after the special
  int8_t a_d, b_d;
  int8_t has_special;
  int8_t pad;
};

template <int MB>
XS_HD bool xs_apply_desc(double S, const XsDesc& d, double* out) {
  XsT a; a.s = d.a_s; a.d = d.a_d;
  double s1;
  if (!xs_apply<MB>(S, (int)d.e0, a, &s1)) { *out = S; return false; }
  if (!d.has_special) { *out = s1; return true; }
  double s2 = xs_add_special<MB>(s1, d.wc);
  if (xs_label(s2) != (int)d.e1) { *out = S; return false; }
  XsT b; b.s = d.b_s; b.d = d.b_d;
  return xs_apply<MB>(s2, (int)d.e1, b, out);
}

// ---- phase A/B, thread-local -------------------------------------------------------------------------------------------
// In:  w[ITEMS], approximate prefix at thread start (sp_start), label of the element before this thread (lab_prev) and the
//      label after this thread's last element (lab_end = label(sp_start + sum(w)), the value the next thread starts from).
// Out: bit mask of special elements, the thread's segmented-scan contribution.
template <int MB, int ITEMS>
XS_HD void xs_thread_reduce(const float (&w)[ITEMS], double sp_start, int lab_prev, int lab_end, uint32_t* special_mask,
                            XsSeg* contrib) {
  XsSeg c = xs_seg_identity();
  uint32_t mask = 0;
  if (lab_prev == lab_end) {  // clean thread: one binade, no label work
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) c.t = xs_compose<MB>(c.t, xs_elem<MB>(w[j], lab_prev), lab_prev);
  } else {
    double p = 0.0;
    int lab = lab_prev;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      p += (double)w[j];
      int l = (j == ITEMS - 1) ? lab_end : xs_label(sp_start + p);
      if (l != lab) {
        mask |= 1u << j;
        c.cnt += 1;
        c.t = xs_identity();
        lab = l;
      } else {
        c.t = xs_compose<MB>(c.t, xs_elem<MB>(w[j], lab), lab);
      }
    }
  }
  *special_mask = mask;
  *contrib = c;
}

// label after element j of a dirty thread (same expression as in xs_thread_reduce)
template <int ITEMS>
XS_HD int xs_elem_label(const float (&w)[ITEMS], double sp_start, int lab_end, int j) {
  if (j == ITEMS - 1) return lab_end;
  double p = 0.0;
  for (int k = 0; k <= j; ++k) p += (double)w[k];
  return xs_label(sp_start + p);
}

// ---- phase C: walk the segments of one tile with the exact incoming state (one thread) ---------------------------------
// seg_agg[s]: aggregate of the regular elements of segment s (s = 0..X); seg_wc[s], seg_e[s] (s = 1..X): the special element
// opening segment s and the speculated label right after it.  Writes base[s] = exact state at the start of segment s's
// regular elements and *S_out = exact state after the tile.  Returns false when any speculation fails verification.
template <int MB>
XS_HD bool xs_walk_segments(double S_in, int e0, int X, const XsT* seg_agg, const float* seg_wc, const int* seg_e,
                            double* base, double* S_out) {
  double S = S_in;
  base[0] = S;
  bool ok = xs_apply<MB>(S, e0, seg_agg[0], &S);
  for (int s = 1; s <= X && ok; ++s) {
    S = xs_add_special<MB>(S, seg_wc[s]);
    ok = (xs_label(S) == seg_e[s]);
    base[s] = S;
    if (ok) ok = xs_apply<MB>(S, seg_e[s], seg_agg[s], &S);
  }
  *S_out = S;
  return ok;
}

// ---- phase D: exact S_k of every element of one thread ------------------------------------------------------------------
template <int MB, int ITEMS>
XS_HD void xs_thread_finalize(const float (&w)[ITEMS], uint32_t special_mask, int lab_prev, int s_base, XsT t_open,
                              const double* base, const int* seg_e, float (&c)[ITEMS]) {
  int s = s_base;
  int E = lab_prev;
  XsT T = t_open;
  double b = base[s];
  int par = xs_parity<MB>(b);
  double q = (E == XS_E_ZERO) ? 0.0 : xs_pow2(E - (MB - 1));
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    double S;
    if (special_mask & (1u << j)) {
      ++s;
      b = base[s];
      par = xs_parity<MB>(b);
      E = seg_e[s];
      q = xs_pow2(E - (MB - 1));
      T = xs_identity();
      S = b;
    } else {
      T = xs_compose<MB>(T, xs_elem<MB>(w[j], E), E);
      S = xs_dadd(b, (T.d && par) ? xs_dadd(T.s, (double)T.d * q) : T.s);
    }
    c[j] = (float)S;
  }
}

// ---- integer form of the transducers (used by the chain of describe_kernel) ---------------------------------------------
// Inside one binade the bit pattern of the state is an integer counter: adding one quantum 2^(E-(MB-1)) adds
// XI_UNIT = 1 << (53 - MB) to the pattern.  A transducer becomes (k, d): k pattern units for an even state, k + d*XI_UNIT for an
// odd one.  Applying and composing are then a handful of integer instructions with no floating-point latency chain.
struct XiT {
  int64_t k;
  int32_t d;
};
XS_HD XiT xi_identity() { XiT t; t.k = 0; t.d = 0; return t; }

// (s, d) in binade E -> integer form; false when s is not an increment that can stay inside the binade
template <int MB>
XS_HD bool xi_from(double s, int d, int E, XiT* out) {
  out->k = 0; out->d = 0;
  if (E == XS_E_ZERO) return s == 0.0 && d == 0;
  const double M = xs_pow2(E);
  if (!(s >= 0.0 && s < M)) return false;
  out->k = (int64_t)(xs_d2u(xs_dadd(M, s)) - xs_d2u(M));
  out->d = d;
  return true;
}

// a THEN b, both in the same binade
template <int MB>
XS_HD XiT xi_compose(const XiT& a, const XiT& b) {
  constexpr int SH = 53 - MB;
  const int64_t U = (int64_t)1 << SH;
  XiT r;
  const int64_t a1 = a.k + (int64_t)a.d * U;
  const int p0 = (int)((a.k >> SH) & 1);        // parity of the state after a, incoming even
  const int p1 = (int)((a1 >> SH) & 1) ^ 1;     // ... incoming odd
  const int64_t r0 = a.k + b.k + (p0 ? (int64_t)b.d * U : 0);
  const int64_t r1 = a1 + b.k + (p1 ? (int64_t)b.d * U : 0);
  r.k = r0;
  r.d = r1 > r0 ? 1 : (r1 < r0 ? -1 : 0);
  return r;
}

// bit pattern of a state in binade E (or 0 with E == XS_E_ZERO) advanced by a transducer; false when it would leave the binade
template <int MB>
XS_HD bool xi_apply(uint64_t bits, int E, const XiT& t, uint64_t* out) {
  constexpr int SH = 53 - MB;
  *out = bits;
  if (E == XS_E_ZERO) return bits == 0 && t.k == 0 && t.d == 0;
  if ((int)(bits >> 52) != E + 1023) return false;
  const int64_t inc = t.k + ((t.d && ((bits >> SH) & 1)) ? (int64_t)t.d * ((int64_t)1 << SH) : 0);
  if (inc < 0) return false;
  const uint64_t nb = bits + (uint64_t)inc;
  *out = nb;
  return (nb >> 52) == (bits >> 52);
}
