// Exact, order-independent emulation of a SEQUENTIAL floating-point prefix sum.
//
// pyfilter.resampling.systematic (reference resampling.py:47) calls torch.cumsum on float32 weights; on CPU that is
//     S_k = fl64(S_{k-1} + (double)w_k),  c_k = fl32(S_k)                       (ATen cumsum_cpu_kernel, acc_type<float>=double)
// and torch.multinomial's CPU kernel (resampling.py:65) builds  S_k = fl32(S_{k-1} + w_k).
// A parallel scan in any floating type re-associates the additions and flips ancestors.  This header
// holds the arithmetic that lets a parallel scan reproduce the sequential result bit for bit:
//
//  * while the running sum stays inside one binade [2^E, 2^(E+1)), it is an integer multiple P*q of the quantum
//    q = 2^(E-(MB-1)) (MB = 53 or 24 mantissa bits) and adding w is  P <- P + RN_q(w)  with ties-to-even decided by the parity
//    of P.  Each element therefore is a 2-state transducer  "increment if P even / increment if P odd"; composition of such
//    transducers is associative, so any scan order gives the same bits;
//  * an element that moves the sum into another binade ("special" element) is applied with one genuine IEEE addition; the
//    caller speculates where those are from an approximate prefix and VERIFIES the speculation on the exact state
//    (xs_apply_* return false when the speculation was wrong, the caller then falls back to a sequential walk).
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

// 2^e as a double, e in [-1074, 1023]
XS_HD double xs_pow2(int e) {
  if (e >= -1022) return xs_u2d((uint64_t)(e + 1023) << 52);
  return xs_u2d(1ull << (e + 1074));
}

// Binade label of a non-negative finite double: floor(log2(x)), XS_E_ZERO for 0.  (Denormal doubles cannot occur: the
// smallest positive float32 is 2^-149.)
XS_HD int xs_label(double x) {
  uint64_t u = xs_d2u(x);
  if ((u << 1) == 0) return XS_E_ZERO;
  return (int)((u >> 52) & 0x7ff) - 1023;
}

// One transducer: increment (in quanta) when the incoming integer state P is even (inc0) or odd (inc0 + d), d in {-1,0,1}.
struct XsT {
  int64_t inc0;
  int32_t d;
};
XS_HD XsT xs_identity() { XsT t; t.inc0 = 0; t.d = 0; return t; }
XS_HD int64_t xs_inc(const XsT& t, int parity) { return t.inc0 + (parity ? (int64_t)t.d : 0); }
// a THEN b
XS_HD XsT xs_compose(const XsT& a, const XsT& b) {
  int64_t a1 = a.inc0 + a.d;
  int64_t r0 = a.inc0 + xs_inc(b, (int)(a.inc0 & 1));
  int64_t r1 = a1 + xs_inc(b, (int)((a1 + 1) & 1));
  XsT r; r.inc0 = r0; r.d = (int32_t)(r1 - r0); return r;
}

// Transducer of one element w >= 0 (float32) when the running sum lives in the binade labelled E with MB mantissa bits.
// Saturates (inc0 = 2^61) when w is so large relative to the binade that it must be a binade-crossing element: the
// verification in xs_apply then fails, which is the intended outcome of a wrong speculation.
template <int MB>
XS_HD XsT xs_elem(float w, int E) {
  XsT t; t.inc0 = 0; t.d = 0;
  uint32_t b = xs_f2u(w) & 0x7fffffffu;
  if (b == 0u || E == XS_E_ZERO) return t;  // zeros never change the sum; in the zero state every regular element is 0
  int eb = (int)(b >> 23);
  uint32_t mw = (b & 0x7fffffu) | (eb ? 0x800000u : 0u);
  int le = (eb ? eb : 1) - 150;        // w = mw * 2^le
  int sh = (E - (MB - 1)) - le;        // quantum exponent minus lsb exponent
  if (sh <= 0) {
    if (-sh > 37) { t.inc0 = (int64_t)1 << 61; return t; }
    t.inc0 = (int64_t)((uint64_t)mw << (-sh));
    return t;
  }
  if (sh >= 25) return t;              // w < q/2: rounds away entirely, never a tie
  uint32_t f = mw >> sh;
  uint32_t rem = mw & ((1u << sh) - 1u);
  uint32_t half = 1u << (sh - 1);
  if (rem == half) {                   // tie: round so that P becomes even
    if (f & 1u) { t.inc0 = (int64_t)f + 1; t.d = -1; }
    else        { t.inc0 = (int64_t)f;     t.d = +1; }
  } else {
    t.inc0 = (int64_t)f + (rem > half ? 1 : 0);
  }
  return t;
}

// Parity of the integer state P = S / q for a state S in binade E (S is a double that is exactly representable with MB bits).
template <int MB>
XS_HD int xs_parity(double S) {
  return (int)((xs_d2u(S) >> (53 - MB)) & 1ull);
}

// S (exact state, binade label E, or S == 0 with E == XS_E_ZERO) advanced by a transducer.  Returns false when the result
// leaves the binade, i.e. when some element inside the aggregated range was in fact a crossing element.
template <int MB>
XS_HD bool xs_apply(double S, int E, const XsT& t, double* out) {
  if (E == XS_E_ZERO) { *out = S; return S == 0.0 && t.inc0 == 0 && t.d == 0; }
  if (xs_label(S) != E) { *out = S; return false; }
  int64_t inc = xs_inc(t, xs_parity<MB>(S));
  if (inc < 0 || inc >= ((int64_t)1 << 54)) { *out = S; return false; }
  double r = S + (double)inc * xs_pow2(E - (MB - 1));  // exact: multiple of q below 2^(E+1) unless the check below fires
  *out = r;
  return r < xs_pow2(E + 1);
}

// The reference operation itself for one element (always valid): S <- fl_MB(S + w).
template <int MB>
XS_HD double xs_add_special(double S, float w) {
  if (MB == 53) return S + (double)w;
#if defined(__CUDA_ARCH__)
  return (double)__fadd_rn((float)S, w);
#else
  volatile float r = (float)S + w;
  return (double)r;
#endif
}

// ---------------------------------------------------------------------------------------------------------------------------
// Systematic probes (reference resampling.py:44-46): p_i = fl32(fl32(i + u) / n), true IEEE division (torch CPU div).
// xs_count_le(c) = #{ i in [0,n) : p_i <= c }  - the number of offspring slots whose probe lies at or below the cumulative
// weight c; particle j's offspring are the probes [count(c_{j-1}), count(c_j)).
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

XS_HD int64_t xs_count_le(float c, float u, int64_t n, float nf) {
  // estimate the last probe index at or below c, then settle with exact float32 probe evaluations (monotone in i)
  double est = (double)c * (double)n - (double)u;
  int64_t i = (int64_t)floor(est);
  if (i < -1) i = -1;
  if (i > n - 1) i = n - 1;
  while (i + 1 < n && xs_probe(i + 1, u, nf) <= c) ++i;
  while (i >= 0 && xs_probe(i, u, nf) > c) --i;
  return i + 1;
}
