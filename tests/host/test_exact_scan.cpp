// Host-side check of the exact-scan arithmetic (pyfilter_b200/csrc/{exact_scan,scan_tile}.h): emulates the tile algorithm
// serially (same per-thread functions as the CUDA kernel, block scans replaced by loops, approximate prefixes summed in a
// DIFFERENT association than the sequential reference) and compares with the plain sequential prefix sum bit for bit.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <random>
#include <algorithm>
#include "../../pyfilter_b200/csrc/scan_tile.h"
#include "../../pyfilter_b200/csrc/philox.h"

template <int MB>
static std::vector<float> seq_cumsum(const std::vector<float>& w) {
  std::vector<float> c(w.size());
  double S = 0;
  for (size_t k = 0; k < w.size(); ++k) { S = xs_add_special<MB>(S, w[k]); c[k] = (float)S; }
  return c;
}

struct Stats { long tiles = 0, slow = 0, specials = 0, desc_ok = 0, desc_fail = 0, opaque = 0; };

static double pairwise(const double* a, int n) { return n == 1 ? a[0] : pairwise(a, n / 2) + pairwise(a + n / 2, n - n / 2); }

template <int MB, int NT, int ITEMS>
static bool run(const std::vector<float>& win, Stats& st, double approx_noise, std::mt19937_64& rng) {
  const int TILE = NT * ITEMS;
  size_t n = win.size();
  size_t ntiles = (n + TILE - 1) / TILE;
  std::vector<float> w(ntiles * TILE, 0.f);
  std::copy(win.begin(), win.end(), w.begin());
  std::vector<float> ref = seq_cumsum<MB>(w), got(w.size());
  // tile sums (pre-kernel): per-thread sequential, pairwise across threads
  std::vector<double> tsum(ntiles);
  for (size_t b = 0; b < ntiles; ++b) {
    double th[NT];
    for (int t = 0; t < NT; ++t) { double s = 0; for (int j = 0; j < ITEMS; ++j) s += (double)w[b * TILE + t * ITEMS + j]; th[t] = s; }
    tsum[b] = pairwise(th, NT);
  }
  double S_in = 0.0;           // exact state chained tile to tile
  double S_desc = 0.0;         // exact state obtained purely through descriptors (look-back path)
  bool desc_chain_valid = true;
  std::uniform_real_distribution<double> un(-1.0, 1.0);
  for (size_t b = 0; b < ntiles; ++b) {
    st.tiles++;
    double sp0 = 0; for (size_t q = 0; q < b; ++q) sp0 += tsum[q];
    sp0 *= (1.0 + approx_noise * un(rng));   // emulate a different association / inaccurate predictor
    float wt[NT][ITEMS];
    for (int t = 0; t < NT; ++t) for (int j = 0; j < ITEMS; ++j) wt[t][j] = w[b * TILE + t * ITEMS + j];
    // phase A: thread sums, exclusive scan (serial here), labels at the thread boundaries
    double tp[NT], tsm[NT]; double acc = 0;
    for (int t = 0; t < NT; ++t) { tp[t] = acc; double s = 0; for (int j = 0; j < ITEMS; ++j) s += (double)wt[t][j]; tsm[t] = s; acc += s; }
    int lab_end[NT], lab_prev[NT];
    for (int t = 0; t < NT; ++t) lab_end[t] = xs_label((sp0 + tp[t]) + tsm[t]);
    int e0 = xs_label(sp0);
    for (int t = 0; t < NT; ++t) lab_prev[t] = t ? lab_end[t - 1] : e0;
    // phase B
    uint32_t mask[NT]; XsSeg contrib[NT]; XsSeg excl[NT];
    XsSeg run_ = xs_seg_identity();
    for (int t = 0; t < NT; ++t) {
      xs_thread_reduce<MB, ITEMS>(wt[t], sp0 + tp[t], lab_prev[t], lab_end[t], &mask[t], &contrib[t]);
      excl[t] = run_; run_ = xs_seg_combine<MB>(run_, contrib[t], lab_prev[t]);
    }
    int X = run_.cnt; st.specials += X;
    const int MAXSEG = 4096;
    static XsT seg_agg[MAXSEG]; static float seg_wc[MAXSEG]; static int seg_e[MAXSEG]; static double base[MAXSEG];
    if (X >= MAXSEG) { fprintf(stderr, "too many segments\n"); return false; }
    for (int t = 0; t < NT; ++t) {
      if (!mask[t]) continue;
      int s = excl[t].cnt; XsT T = excl[t].t; int E = lab_prev[t];
      for (int j = 0; j < ITEMS; ++j) {
        if (mask[t] & (1u << j)) { seg_agg[s] = T; ++s; seg_wc[s] = wt[t][j]; E = xs_elem_label<ITEMS>(wt[t], sp0 + tp[t], lab_end[t], j); seg_e[s] = E; T = xs_identity(); }
        else T = xs_compose<MB>(T, xs_elem<MB>(wt[t][j], E), E);
      }
    }
    seg_agg[X] = run_.t;
    // descriptor
    bool has_desc = X <= 1;
    XsDesc d{};
    if (has_desc) {
      d.e0 = (int16_t)e0; d.a_s = seg_agg[0].s; d.a_d = (int8_t)seg_agg[0].d; d.has_special = (int8_t)X;
      if (X) { d.wc = seg_wc[1]; d.e1 = (int16_t)seg_e[1]; d.b_s = seg_agg[1].s; d.b_d = (int8_t)seg_agg[1].d; }
    } else st.opaque++;
    // phase C
    double S_out; bool ok = xs_walk_segments<MB>(S_in, e0, X, seg_agg, seg_wc, seg_e, base, &S_out);
    if (ok) {
      for (int t = 0; t < NT; ++t) {
        float c[ITEMS];
        xs_thread_finalize<MB, ITEMS>(wt[t], mask[t], lab_prev[t], excl[t].cnt, excl[t].t, base, seg_e, c);
        for (int j = 0; j < ITEMS; ++j) got[b * TILE + t * ITEMS + j] = c[j];
      }
    } else {
      st.slow++;
      double S = S_in;
      for (int k = 0; k < TILE; ++k) { S = xs_add_special<MB>(S, w[b * TILE + k]); got[b * TILE + k] = (float)S; }
      S_out = S;
    }
    // descriptor path must agree whenever it claims success
    if (desc_chain_valid && has_desc) {
      double o; bool dk = xs_apply_desc<MB>(S_desc, d, &o);
      if (dk) { st.desc_ok++; if (o != S_out) { fprintf(stderr, "descriptor state mismatch at tile %zu\n", b); return false; } S_desc = o; }
      else { st.desc_fail++; if (ok) { fprintf(stderr, "descriptor failed but walk succeeded at tile %zu\n", b); return false; } S_desc = S_out; }
    } else S_desc = S_out;
    S_in = S_out;
  }
  for (size_t k = 0; k < w.size(); ++k)
    if (got[k] != ref[k]) { fprintf(stderr, "MISMATCH at %zu: got %.9g ref %.9g\n", k, got[k], ref[k]); return false; }
  return true;
}

static std::vector<float> make_weights(size_t n, double logstd, int kind, std::mt19937_64& rng) {
  std::normal_distribution<double> nd(0, 1);
  std::vector<double> lw(n);
  for (auto& v : lw) v = nd(rng) * logstd;
  if (kind == 1) for (size_t i = 0; i < n; i += 3) lw[i] = -1e30;            // exact zeros
  if (kind == 2) for (size_t i = 0; i < n / 2; ++i) lw[i] -= 40;             // tiny leading weights
  if (kind == 3) for (size_t i = 0; i < n; ++i) lw[i] = 0;                   // uniform
  if (kind == 4) { for (auto& v : lw) v = -1e30; lw[n / 3] = 0; lw[n - 1] = 0; }  // two-point
  if (kind == 5) for (size_t i = 0; i < n; ++i) lw[i] = -0.00002 * (double)i;    // geometric decay
  double m = *std::max_element(lw.begin(), lw.end());
  std::vector<float> e(n); double z = 0;
  for (size_t i = 0; i < n; ++i) { e[i] = expf((float)(lw[i] - m)); z += e[i]; }
  float zf = (float)z;
  for (auto& v : e) v = v / zf;
  return e;
}

// ancestors through xs_count_le vs. a lower_bound over probes
static bool check_counts(const std::vector<float>& w, float u, std::mt19937_64&) {
  size_t n = w.size();
  std::vector<float> c = seq_cumsum<53>(w); c[n - 1] = 1.0f;
  float nf = (float)n;
  std::vector<float> p(n);
  for (size_t i = 0; i < n; ++i) p[i] = xs_probe((int64_t)i, u, nf);
  int64_t prev = 0;
  std::vector<int64_t> anc(n, -1);
  for (size_t j = 0; j < n; ++j) {
    int64_t cnt = xs_count_le(c[j], u, (int64_t)n, nf);
    int64_t expect = std::upper_bound(p.begin(), p.end(), c[j]) - p.begin();
    if (cnt != expect) { fprintf(stderr, "count mismatch j=%zu got %ld expect %ld\n", j, (long)cnt, (long)expect); return false; }
    for (int64_t i = prev; i < cnt; ++i) anc[i] = (int64_t)j;
    if (cnt > prev) prev = cnt;
  }
  for (size_t i = 0; i < n; ++i) {
    int64_t e = std::lower_bound(c.begin(), c.end(), p[i]) - c.begin();
    if (anc[i] != e) { fprintf(stderr, "ancestor mismatch i=%zu got %ld expect %ld\n", i, (long)anc[i], (long)e); return false; }
  }
  return true;
}

// xs_count_fast against xs_count_le (itself checked against lower_bound over the probes in check_counts)
static bool check_count_fast(std::mt19937_64& rng) {
  int ns[] = {1, 2, 3, 7, 300, 1000, 4097, 1000003, 4000000, (1 << 22), (1 << 23) - 1, (1 << 23)};
  for (int n : ns) {
    float nf = (float)n; double nfd = (double)nf;
    for (int rep = 0; rep < 200000; ++rep) {
      float u = (float)(rng() >> 40) * 5.9604644775390625e-08f;
      if (rep % 17 == 0) u = 0.f;
      float c;
      int mode = rep % 5;
      if (mode == 0) c = (float)((rng() >> 11) * 1.1102230246251565e-16);
      else if (mode == 1) { long i = rng() % (unsigned long)n; c = xs_probe(i, u, nf); int d = (int)(rng() % 5) - 2; c = xs_u2f(xs_f2u(c) + d); if (!(c >= 0)) c = 0; }
      else if (mode == 2) c = (float)((rng() >> 11) * 1.1102230246251565e-16) * 1e-3f;
      else if (mode == 3) c = 1.0f - (float)(rng() % 64) * 5.96e-8f;
      else c = (rng() % 3 == 0) ? 0.f : (float)((rng() >> 11) * 1.1102230246251565e-16) * 1.0001f;
      if (c != 0.f && c < 1.2e-38f) continue;
      int a = (int)xs_count_le(c, u, (int64_t)n, nf), b = xs_count_fast(c, u, n, (double)n, nfd);
      if (a != b) { fprintf(stderr, "count_fast mismatch n=%d c=%.9g u=%.9g ref=%d fast=%d\n", n, c, u, a, b); return false; }
    }
  }
  return true;
}

// integer transducers (XiT) against the double ones: sequential and composed application, validity included
template <int MB>
static bool check_integer_transducers(std::mt19937_64& rng) {
  std::uniform_real_distribution<double> un(0, 1);
  for (int rep = 0; rep < 200000; ++rep) {
    int E = -1 - (int)(rng() % 30);
    double q = xs_pow2(E - (MB - 1));
    double S = xs_pow2(E) + floor(un(rng) * 0.98 * (xs_pow2(E) / q)) * q;
    int ng = 1 + rng() % 6;
    std::vector<XsT> ts; std::vector<XiT> ti; bool conv_ok = true;
    for (int g = 0; g < ng; ++g) {
      XsT T = xs_identity();
      int ne = 1 + rng() % 20;
      for (int j = 0; j < ne; ++j) {
        int ew = E - MB - 4 + (int)(rng() % 12);
        if (rng() % 40 == 0) ew = E - 3;  // large element: may leave the binade
        float w = (float)ldexp(1.0 + floor(un(rng) * 16) / 16.0, ew);
        if (rng() % 7 == 0) w = (float)ldexp(1.0, E - MB);  // half a quantum: tie
        T = xs_compose<MB>(T, xs_elem<MB>(w, E), E);
      }
      ts.push_back(T);
      XiT I; conv_ok = conv_ok && xi_from<MB>(T.s, T.d, E, &I); ti.push_back(I);
    }
    if (!conv_ok) continue;
    double Sd = S; uint64_t Sb = xs_d2u(S); bool okd = true, oki = true;
    for (int g = 0; g < ng && okd; ++g) okd = xs_apply<MB>(Sd, E, ts[g], &Sd);
    for (int g = 0; g < ng && oki; ++g) oki = xi_apply<MB>(Sb, E, ti[g], &Sb);
    XsT Tc = xs_identity(); XiT Ic = xi_identity();
    for (int g = 0; g < ng; ++g) { Tc = xs_compose<MB>(Tc, ts[g], E); Ic = xi_compose<MB>(Ic, ti[g]); }
    double Sc; uint64_t Sbc;
    bool okc_d = xs_apply<MB>(S, E, Tc, &Sc), okc_i = xi_apply<MB>(xs_d2u(S), E, Ic, &Sbc);
    if (okd != oki || okc_d != okc_i) { fprintf(stderr, "XiT validity mismatch MB=%d\n", MB); return false; }
    if (okd && (xs_d2u(Sd) != Sb)) { fprintf(stderr, "XiT state mismatch MB=%d\n", MB); return false; }
    if (okc_d && (xs_d2u(Sc) != Sbc)) { fprintf(stderr, "XiT composed state mismatch MB=%d\n", MB); return false; }
  }
  return true;
}

// philox4x32_10_keys (round keys precomputed on the host and passed as kernel arguments) is the same generator as philox4x32_10
static bool check_philox_keys(std::mt19937_64& rng) {
  for (int it = 0; it < 200000; ++it) {
    const uint32_t c0 = (uint32_t)rng(), c1 = (uint32_t)rng(), c2 = (uint32_t)rng(), c3 = (uint32_t)rng();
    const uint64_t seed = rng();
    uint32_t keys[20];
    philox_round_keys((uint32_t)seed, (uint32_t)(seed >> 32), keys);
    const Philox4 a = philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
    const Philox4 b = philox4x32_10_keys(c0, c1, c2, c3, keys);
    if (a.x != b.x || a.y != b.y || a.z != b.z || a.w != b.w) { fprintf(stderr, "FAILED philox keys at %d\n", it); return false; }
  }
  // known answer of Philox4x32-10 (Random123 kat_vectors: counter = key = 0)
  const Philox4 z = philox4x32_10(0u, 0u, 0u, 0u, 0u, 0u);
  if (z.x != 0x6627e8d5u || z.y != 0xe169c58du || z.z != 0xbc57ac4cu || z.w != 0x9b00dbd8u) { fprintf(stderr, "FAILED philox KAT %08x %08x %08x %08x\n", z.x, z.y, z.z, z.w); return false; }
  return true;
}

int main(int argc, char** argv) {
  std::mt19937_64 rng(12345);
  if (!check_philox_keys(rng)) return 1;
  if (!check_count_fast(rng)) return 1;
  if (!check_integer_transducers<53>(rng) || !check_integer_transducers<24>(rng)) return 1;
  size_t big = argc > 1 ? (size_t)atol(argv[1]) : (size_t)1 << 20;
  Stats st;
  bool ok = true;
  size_t sizes[] = {1, 2, 31, 32, 33, 1000, 4096, 65537, big};
  for (size_t n : sizes)
    for (int kind = 0; kind <= 5; ++kind)
      for (double ls : {0.3, 2.0, 6.0}) {
        auto w = make_weights(n, ls, kind, rng);
        ok = ok && run<53, 8, 4>(w, st, 0.0, rng);
        ok = ok && run<53, 256, 16>(w, st, 0.0, rng);
        ok = ok && run<53, 32, 8>(w, st, 1e-13, rng);     // sloppy predictor: must still be exact (via verification)
        ok = ok && run<53, 32, 8>(w, st, 1e-3, rng);      // useless predictor: exactness must survive
        ok = ok && run<24, 32, 8>(w, st, 0.0, rng);       // float32-accumulated prefix (torch.multinomial): fp64 predictor is poor
        ok = ok && run<24, 256, 16>(w, st, 0.0, rng);
        if (!ok) { fprintf(stderr, "FAILED n=%zu kind=%d logstd=%g\n", n, kind, ls); return 1; }
        if (n <= 70000) for (float u : {0.0f, 0.37f, 0.99999994f}) ok = ok && check_counts(w, u, rng);
        if (!ok) { fprintf(stderr, "FAILED counts n=%zu kind=%d\n", n, kind); return 1; }
      }
  printf("OK tiles=%ld slow=%ld specials=%ld desc_ok=%ld desc_fail=%ld opaque=%ld\n", st.tiles, st.slow, st.specials,
         st.desc_ok, st.desc_fail, st.opaque);
  return 0;
}
