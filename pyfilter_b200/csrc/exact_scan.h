// Exact, order-independent emulation of a SEQUENTIAL floating-point prefix sum.
//
// pyfilter.resampling.systematic (reference resampling.py:47) calls torch.cumsum on float32 weights; on CPU that is
//     S_k = fl64(S_{k-1} + (double)w_k),  c_k = fl32(S_k)                       (ATen cumsum_cpu_kernel, acc_type<float>=double)
// and torch.multinomial's CPU kernel (resampling.py:65) builds  S_k = fl32(S_{k-1} + w_k).
// A parallel scan in any floating type re-associates the additions and flips ancestors.  This header
// holds the arithmetic that lets a parallel scan reproduce the sequential result bit for bit:
//
//  * while the running sum stays inside one binade [2^E, 2^(E+1)), it is a multiple of the quantum q = 2^(E-(MB-1))
//    (MB = 53 or 24 mantissa bits) and adding w is  S <- S + RN_q(w)  with ties-to-even decided by the parity of S/q.
//    RN_q(w) for an even state is obtained with two IEEE additions, (2^E + w) - 2^E; a tie (remainder exactly q/2) makes the
//    increment depend on the incoming parity.  Each element therefore is a 2-state transducer "increment if S/q even /
//    increment if S/q odd"; composition of such transducers is associative, so any scan order gives the same bits, and while
//    no tie is involved composition is a plain (exact) double addition;
//  * an element that moves the sum into another binade ("special" element) is applied with one genuine IEEE addition; the
//    caller speculates where those are from an approximate prefix and VERIFIES the speculation on the exact state
//    (xs_apply* return false when the speculation was wrong, the caller then falls back to a sequential walk).
//
// Everything here is __host__ __device__ so that tests/host/ can exercise it with g++ (no GPU in the build container).
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define XS_HD __host__ __device__ __forceinline__
#else
#define XS_HD inline
#endif

#define XS_E_ZERO (-4000)  // "binade" label of the state S == 0 (no element added yet / only zeros so far)

XS_HD uint64_t xs_d2u(double x) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}
XS_HD double xs_u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double x; memcpy(&x, &u, 8); return x;
#endif
}
XS_HD uint32_t xs_f2u(float x) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(x);
#else
  uint32_t u; memcpy(&u, &x, 4); return u;
#endif
}
XS_HD float xs_u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float x; memcpy(&x, &u, 4); return x;
#endif
}

// IEEE add/sub that the compiler may not contract or re-associate
XS_HD double xs_dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  volatile double r = a + b; return r;
#endif
}
XS_HD float xs_fadd(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  volatile float r = a + b; return r;
#endif
}

// 2^e as a double, e in [-1022, 1023] (labels of float32-derived sums are >= -149)
XS_HD double xs_pow2(int e) { return xs_u2d((uint64_t)(e + 1023) << 52); }

// Binade label of a non-negative finite double: floor(log2(x)), XS_E_ZERO for 0.
XS_HD int xs_label(double x) {
  uint64_t u = xs_d2u(x);
  if ((u << 1) == 0) return XS_E_ZERO;
  return (int)((u >> 52) & 0x7ff) - 1023;
}

// One transducer: increment when the incoming state S/q is even (s), or odd (s + d*q), d in {-1,0,1}.  s is a multiple of q.
struct XsT {
  double s;
  int32_t d;
};
XS_HD XsT xs_identity() { XsT t; t.s = 0.0; t.d = 0; return t; }

// parity of x/q for a multiple x of q = 2^(E-(MB-1)) with 0 <= x <= 2^E  (read off the mantissa of 2^E + x)
template <int MB>
XS_HD int xs_par_inc(double x, double M) { return (int)((xs_d2u(xs_dadd(M, x)) >> (53 - MB)) & 1ull); }
// parity of S/q for a state S in binade E
template <int MB>
XS_HD int xs_parity(double S) { return (int)((xs_d2u(S) >> (53 - MB)) & 1ull); }

// a THEN b, both in binade E
template <int MB>
XS_HD XsT xs_compose(const XsT& a, const XsT& b, int E) {
  XsT r;
  if ((a.d | b.d) == 0) { r.s = xs_dadd(a.s, b.s); r.d = 0; return r; }
  const double M = xs_pow2(E), q = xs_pow2(E - (MB - 1));
  const double a1 = xs_dadd(a.s, (double)a.d * q);
  const int p0 = xs_par_inc<MB>(a.s, M), p1 = xs_par_inc<MB>(a1, M) ^ 1;
  const double r0 = xs_dadd(a.s, xs_dadd(b.s, p0 ? (double)b.d * q : 0.0));
  const double r1 = xs_dadd(a1, xs_dadd(b.s, p1 ? (double)b.d * q : 0.0));
  r.s = r0;
  const double diff = xs_dadd(r1, -r0);
  r.d = diff > 0.0 ? 1 : (diff < 0.0 ? -1 : 0);
  return r;
}

// Transducer of one regular element w >= 0 (float32) while the running sum lives in binade E (MB mantissa bits).
// If w is too large for the binade the result is wrong but >= 2^E, so the verification in xs_apply fails as it must.
template <int MB>
XS_HD XsT xs_elem(float w, int E) {
  XsT t; t.s = 0.0; t.d = 0;
  if (E == XS_E_ZERO) return t;  // in the zero state every regular element is 0 (the first non-zero one is special)
  double r;
  if (MB == 53) {
    const double M = xs_pow2(E);
    r = xs_dadd(xs_dadd(M, (double)w), -M);
  } else {
    const float Mf = (float)xs_pow2(E);
    r = (double)xs_fadd(xs_fadd(Mf, w), -Mf);
  }
  const double rem = xs_dadd((double)w, -r);  // exact
  const double hq = xs_pow2(E - MB);          // q / 2
  t.s = r;
  t.d = (rem == hq) ? 1 : ((rem == -hq) ? -1 : 0);
  return t;
}

// S (exact state in binade E, or S == 0 with E == XS_E_ZERO) advanced by a transducer.  Returns false when the result leaves
// the binade, i.e. when some element inside the aggregated range was in fact a crossing element.
template <int MB>
XS_HD bool xs_apply(double S, int E, const XsT& t, double* out) {
  if (E == XS_E_ZERO) { *out = S; return S == 0.0 && t.s == 0.0 && t.d == 0; }
  if (xs_label(S) != E) { *out = S; return false; }
  const double inc = (t.d && xs_parity<MB>(S)) ? xs_dadd(t.s, (double)t.d * xs_pow2(E - (MB - 1))) : t.s;
  const double r = xs_dadd(S, inc);  // exact: multiple of q below 2^(E+1) unless the check below fires
  *out = r;
  return r < xs_pow2(E + 1) && inc >= 0.0;
}

// The reference operation itself for one element (always valid): S <- fl_MB(S + w).
template <int MB>
XS_HD double xs_add_special(double S, float w) {
  if (MB == 53) return xs_dadd(S, (double)w);
  return (double)xs_fadd((float)S, w);
}

// ---------------------------------------------------------------------------------------------------------------------------
// Systematic probes (reference resampling.py:44-46): p_i = fl32(fl32(i + u) / n), true IEEE division (torch CPU div).
// xs_count_le(c) = #{ i in [0,n) : p_i <= c }  - the number of offspring slots whose probe lies at or below the cumulative
// weight c; particle j's offspring are the probes [count(c_{j-1}), count(c_j)).
// The division is never executed:  fl32(s / nf) <= c  <=>  s <= m * nf (or < when the tie rounds away from c), with m the
// midpoint between c and the next float above it; m * nf is exact in double (25 x 24 bits).
// ---------------------------------------------------------------------------------------------------------------------------
XS_HD float xs_probe(int64_t i, float u, float nf) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(__fadd_rn((float)i, u), nf);
#else
  volatile float s = (float)i + u;
  volatile float r = s / nf;
  return r;
#endif
}

template <typename I>
XS_HD I xs_count_le_t(float c, float u, I n, float nf) {
  const uint32_t cb = xs_f2u(c);
  const double m = 0.5 * ((double)c + (double)xs_u2f(cb + 1u));
  const double t = m * (double)nf;
  const bool incl = !(cb & 1u);  // a probe landing exactly on the midpoint rounds to the even neighbour
  double est = floor(t - (double)u);
  if (est < -1.0) est = -1.0;
  if (est > (double)(n - 1)) est = (double)(n - 1);
  I i = (I)est;
#define XS_PROBE_OK(ii) (incl ? ((double)xs_fadd((float)(ii), u) <= t) : ((double)xs_fadd((float)(ii), u) < t))
  while (i + 1 < n && XS_PROBE_OK(i + 1)) ++i;
  while (i >= 0 && !XS_PROBE_OK(i)) --i;
#undef XS_PROBE_OK
  return i + 1;
}
XS_HD int64_t xs_count_le(float c, float u, int64_t n, float nf) { return xs_count_le_t<int64_t>(c, u, n, nf); }

// Lean variant for the kernels' hot path (two float->double conversions, 4 double-precision operations, no division).  Requirements: n <= 2^23,
// c == 0 or a normal float, u == 0 or u >= 2^-64.  With t = (c + ulp(c)/2) * n exact in double and tt = t, or the double just
// below t when the midpoint tie rounds away from c, probe i qualifies iff fl32(i + u) <= tt.  For K = floor(tt): every
// i <= K - 1 has i + u < K <= tt, hence fl32(i + u) <= K <= tt (rounding is monotone, K is a float); every i >= K + 1 has
// fl32(i + u) >= K + 1 > tt.  Only probe K itself has to be evaluated.
XS_HD int32_t xs_count_fast(float c, float u, int32_t n, double nd /* (double)n */, double nfd /* (double)(float)n */) {
  const uint32_t cb = xs_f2u(c);
#if defined(__CUDA_ARCH__)
  // (double)c + half an ulp of c: the conversion leaves the low 29 mantissa bits clear, bit 28 is half a float32 ulp
  const double cd = (double)c;
  const double t = __hiloint2double(__double2hiint(cd), __double2loint(cd) | 0x10000000) * nfd;  // exact: 25 x 24 bits
#else
  const uint64_t mb = ((uint64_t)((cb >> 3) + 0x38000000u) << 32) | (uint64_t)((cb << 29) | 0x10000000u);
  const double t = xs_u2d(mb) * nfd;
#endif
  const double tt = xs_u2d(xs_d2u(t) - (uint64_t)(cb & 1u));           // t > 0
  const double tc = (tt < nd) ? tt : nd;                               // also catches NaN / inf
#if defined(__CUDA_ARCH__)
  const int32_t K = __double2loint(__dadd_rd(tc, 4503599627370496.0)); // floor(tc) read off the mantissa of 2^52 + tc
#else
  const int32_t K = (int32_t)floor(tc);
#endif
  const float Kf = xs_fadd(xs_u2f(0x4B000000u | (uint32_t)K), -8388608.0f);   // exact for K <= 2^23
#if defined(__CUDA_ARCH__)
  const double sd = (double)xs_fadd(Kf, u);
#else
  const uint32_t sb = xs_f2u(xs_fadd(Kf, u));
  const uint32_t shi = sb ? (sb >> 3) + 0x38000000u : 0u;
  const double sd = xs_u2d(((uint64_t)shi << 32) | (uint64_t)(sb << 29));
#endif
  const int32_t r = K + (sd <= tt ? 1 : 0);
  return r < n ? r : n;
}

// The same count with fewer instructions for the move kernel (no upper clamp: the caller clamps into [lo, n_out]).  tt = t, or the double
// just below t when the midpoint tie rounds away from c, comes out of ONE fused multiply-add rounded down: the product (c + ulp/2) * n is
// exact, the addend is -0 for an even mantissa and minus the smallest denormal for an odd one.  A garbage c (NaN normalisers of a dead
// column) gives a garbage but harmless count: the caller's clamp keeps every index in range.
XS_HD int32_t xs_count_lean(float c, float u, int32_t n, double nfd /* (double)(float)n */) {
#if defined(__CUDA_ARCH__)
  const uint32_t cb = __float_as_uint(c);
  const double cd = (double)c;
  const double cdh = __hiloint2double(__double2hiint(cd), __double2loint(cd) | 0x10000000);
  const double adj = __hiloint2double((int)0x80000000u, (int)(cb & 1u));
  const double tt = __fma_rd(cdh, nfd, adj);
  const int32_t K = __double2loint(__dadd_rd(tt, 4503599627370496.0));   // floor(tt) read off the mantissa of 2^52 + tt (tt < 2^31)
  const float Kf = __fadd_rn(__uint_as_float(0x4B000000u | (uint32_t)K), -8388608.0f);
  const double sd = (double)__fadd_rn(Kf, u);
  return K + (sd <= tt ? 1 : 0);
#else
  return xs_count_fast(c, u, n, (double)n, nfd);
#endif
}

