// Per-thread phases of the exact tile scan (see exact_scan.h for the arithmetic and DESIGN.md for the kernel structure).
//
// A tile is TILE = NT * ITEMS consecutive weights of one column; thread t owns elements [t*ITEMS, (t+1)*ITEMS).
//   phase A  approximate (fp64, any association) inclusive prefix S'_k -> binade label of every element; an element whose
//            label differs from its predecessor's is "special" (speculated binade crossing) and opens a new segment.
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
// a THEN b
XS_HD XsSeg xs_seg_combine(const XsSeg& a, const XsSeg& b) {
  XsSeg r;
  r.cnt = a.cnt + b.cnt;
  r.t = b.cnt ? b.t : xs_compose(a.t, b.t);
  return r;
}

// Tile descriptor published for the decoupled look-back (valid when the tile holds at most one special element).
struct XsDesc {
  int64_t a_inc0;   // segment 0 aggregate
  int64_t b_inc0;   // segment 1 aggregate (after the special), if any
  float wc;         // the special element
  int16_t e0;       // speculated label of the incoming state
  int16_t e1;       // speculated label right after the special
  int8_t a_d, b_d;
  int8_t has_special;
  int8_t pad;
};

template <int MB>
XS_HD bool xs_apply_desc(double S, const XsDesc& d, double* out) {
  XsT a; a.inc0 = d.a_inc0; a.d = d.a_d;
  double s1;
  int lab = xs_label(S);
  if (lab != (int)d.e0) { *out = S; return false; }
  if (!xs_apply<MB>(S, lab, a, &s1)) { *out = S; return false; }
  if (!d.has_special) { *out = s1; return true; }
  double s2 = xs_add_special<MB>(s1, d.wc);
  if (xs_label(s2) != (int)d.e1) { *out = S; return false; }
  XsT b; b.inc0 = d.b_inc0; b.d = d.b_d;
  return xs_apply<MB>(s2, (int)d.e1, b, out);
}

// ---- phase A/B, thread-local -------------------------------------------------------------------------------------------
// In:  w[ITEMS], approximate prefix at thread start (sp_start) and the label of the element before this thread (lab_prev).
// Out: bit mask of special elements, label after the last element, the thread's segmented-scan contribution, and the
//      transducer of the regular elements BEFORE the first special (== whole thread when there is none).
template <int MB, int ITEMS>
XS_HD void xs_thread_label_and_reduce(const float (&w)[ITEMS], double sp_start, int lab_prev, uint32_t* special_mask,
                                      XsSeg* contrib, XsT* pre) {
  double sp = sp_start;
  int lab = lab_prev;
  uint32_t mask = 0;
  XsSeg c = xs_seg_identity();
  XsT p = xs_identity();
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    sp += (double)w[j];
    int l = xs_label(sp);
    if (l != lab) {
      mask |= 1u << j;
      if (c.cnt == 0) p = c.t;
      c.cnt += 1;
      c.t = xs_identity();
      lab = l;
    } else {
      c.t = xs_compose(c.t, xs_elem<MB>(w[j], lab));
    }
  }
  if (c.cnt == 0) p = c.t;
  *special_mask = mask;
  *contrib = c;
  *pre = p;
}

// Labels only (to learn the label at the end of each thread before phase B can start).
template <int ITEMS>
XS_HD int xs_thread_end_label(const float (&w)[ITEMS], double sp_start) {
  double sp = sp_start;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) sp += (double)w[j];
  return xs_label(sp);
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
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    double S;
    if (special_mask & (1u << j)) {
      ++s;
      b = base[s];
      par = xs_parity<MB>(b);
      E = seg_e[s];
      T = xs_identity();
      S = b;
    } else {
      T = xs_compose(T, xs_elem<MB>(w[j], E));
      S = (E == XS_E_ZERO) ? b : b + (double)xs_inc(T, par) * xs_pow2(E - (MB - 1));
    }
    c[j] = (float)S;
  }
}
