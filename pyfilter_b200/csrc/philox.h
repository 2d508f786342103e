// Counter-based Philox4x32-10 kept entirely in registers (no curand state in memory): the stream is a pure function of
// (seed, particle group, time step, purpose), so results do not depend on grid shape and can be re-generated for parity dumps.
#pragma once
#include <stdint.h>
#include "exact_scan.h"

struct Philox4 { uint32_t x, y, z, w; };

XS_HD uint32_t smcb_mulhi(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

XS_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = smcb_mulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = smcb_mulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  Philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3; return o;
}

// the same generator with the ten round keys precomputed (kernel arguments: they become constant-bank operands of the xor instead of
// twenty uniform additions per call)
XS_HD void philox_round_keys(uint32_t k0, uint32_t k1, uint32_t* keys /*20*/) {
  for (int r = 0; r < 10; ++r) { keys[2 * r] = k0; keys[2 * r + 1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
}
XS_HD Philox4 philox4x32_10_keys(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t* keys) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = smcb_mulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = smcb_mulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ keys[2 * r], n2 = hi0 ^ c3 ^ keys[2 * r + 1];
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
  }
  Philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3; return o;
}

// purposes (counter word 3)
#define SMCB_RNG_TRANSITION 0u   // + state dimension
#define SMCB_RNG_INIT 8u         // + state dimension
#define SMCB_RNG_SYSTEMATIC 16u
#define SMCB_RNG_MULTINOMIAL 17u

// uniform in [0,1) with 24 random bits (what torch's float32 uniform_ produces)
XS_HD float smcb_u01(uint32_t r) { return (float)(r >> 8) * 5.9604644775390625e-08f; }
// uniform in (0,1), never 0: safe for log
XS_HD float smcb_u01_open(uint32_t r) { return ((float)(r >> 8) + 0.5f) * 5.9604644775390625e-08f; }
// uniform double in [0,1) with 53 random bits (what torch.multinomial's CPU path draws)
XS_HD double smcb_u01_double(uint32_t hi, uint32_t lo) {
  uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;
  return (double)v * 1.1102230246251565e-16;
}

#if defined(__CUDACC__)
// four N(0,1) draws from one Philox block (two Box-Muller pairs)
__device__ __forceinline__ void smcb_normal4(const Philox4& r, float (&z)[4]) {
  float u0 = smcb_u01_open(r.x), u1 = smcb_u01(r.y), u2 = smcb_u01_open(r.z), u3 = smcb_u01(r.w);
  float ra, rb, l0, l2;  // sqrt(-2 ln u) = sqrt(-2 ln2 * lg2 u): two SFU operations each
  // u0, u2 >= 2^-25 are normal numbers: the .ftz form returns the same bits as __log2f without its scaling code for denormals
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l0) : "f"(u0));
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u2));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(ra) : "f"(-1.3862943611198906f * l0));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rb) : "f"(-1.3862943611198906f * l2));
  float s0, c0, s1, c1;
  __sincosf(6.283185307179586f * u1, &s0, &c0);
  __sincosf(6.283185307179586f * u3, &s1, &c1);
  z[0] = ra * c0; z[1] = ra * s0; z[2] = rb * c1; z[3] = rb * s1;
}
#endif
