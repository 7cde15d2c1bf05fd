// Host logic test for bayesssm_b200/csrc/bssm_exact.cuh (no GPU needed).
// Emulates, tile by tile, exactly the phases the CUDA exact-scan kernels run
// (approximate scan -> binade/crossing detection -> parity-map tile reduction ->
// serial chain over tiles -> in-tile parity-map scan) and compares the result with
// the sequential double cumsum of src/resampling.cpp:25.  Built and run by
// tests/test_exact_scan_host.py.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../bayesssm_b200/csrc/bssm_exact.cuh"

using namespace bssm;

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t next_u64() {
  uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static double next_unit() { return ((double)(next_u64() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

struct Stats { long n_fallback = 0, n_serial_tiles = 0, n_tiles = 0; };

// Near a power of two the approximate scan may name the wrong binade for a whole run of tiles
// (e.g. a long tail of tiny weights while the cdf sits at 1 -/+ a few ulp).  Such tiles carry
// a second map for the neighbouring binade; the chain picks by the exact value.
static int alt_binade(double first_prev, double last, int be) {
  const u64 M = (1ull << 52) - 1, NEAR = 1ull << 26;
  u64 f0 = dbits(first_prev) & M, f1 = dbits(last) & M;
  if (f1 >= M + 1 - NEAR || f0 >= M + 1 - NEAR) return be + 1;
  if (f0 < NEAR || f1 < NEAR) return be - 1;
  return -1;
}

// always exact: a tile whose assumed binade fails verification is redone serially by the chain
static bool exact_scan_emulated(const std::vector<double>& v, int tile, std::vector<double>& out, Stats& st,
                                double perturb) {
  const int n = (int)v.size();
  const int ntiles = (n + tile - 1) / tile;
  out.assign(n, 0.0);
  // phase A: approximate scan (pairwise inside the tile so the rounding order differs from sequential)
  std::vector<double> part(ntiles), approx(n);
  for (int t = 0; t < ntiles; t++) {
    int lo = t * tile, hi = std::min(n, lo + tile);
    double s0 = 0, s1 = 0;
    for (int i = lo; i < hi; i += 2) { s0 += v[i]; if (i + 1 < hi) s1 += v[i + 1]; }
    part[t] = s0 + s1;
  }
  std::vector<double> tprefix(ntiles);
  { double run = 0; for (int t = 0; t < ntiles; t++) { tprefix[t] = run; run += part[t]; } }
  for (int t = 0; t < ntiles; t++) {
    int lo = t * tile, hi = std::min(n, lo + tile);
    double run = tprefix[t] * (1.0 + perturb);  // optional deliberate error to stress the verification
    for (int i = lo; i < hi; i++) { run += v[i]; approx[i] = run; }
  }
  // phase B: tile classification + tile maps
  std::vector<int> tile_be(ntiles), tile_alt(ntiles), used_be(ntiles);
  std::vector<char> regular(ntiles);
  std::vector<ParFn> tfn(ntiles), tfn_alt(ntiles);
  for (int t = 0; t < ntiles; t++) {
    int lo = t * tile, hi = std::min(n, lo + tile);
    double prev = (lo == 0) ? 0.0 : approx[lo - 1];
    int be = biased_exp(prev);
    bool reg = be >= 2 && be <= 2044;
    for (int i = lo; i < hi && reg; i++) if (biased_exp(approx[i]) != be) reg = false;
    ParFn f = parfn_identity(), g = parfn_identity();
    int alt = reg ? alt_binade(prev, approx[hi - 1], be) : -1;
    for (int i = lo; i < hi && reg; i++) {
      f = parfn_compose(f, parfn_element(v[i], be));
      if (alt >= 0) g = parfn_compose(g, parfn_element(v[i], alt));
    }
    tile_be[t] = be; regular[t] = reg; tfn[t] = f; tile_alt[t] = alt; tfn_alt[t] = g;
  }
  // phase C: serial chain over tiles
  std::vector<double> cstart(ntiles);
  double c = 0.0;
  for (int t = 0; t < ntiles; t++) {
    int lo = t * tile, hi = std::min(n, lo + tile);
    cstart[t] = c;
    st.n_tiles++;
    int use = -1;
    if (regular[t]) {
      int bc = biased_exp(c);
      if (bc == tile_be[t]) use = tile_be[t];
      else if (tile_alt[t] >= 0 && bc == tile_alt[t]) use = tile_alt[t];
    }
    if (use >= 0) {
      i64 C = to_units(c, use);
      i64 C2 = parfn_apply(use == tile_be[t] ? tfn[t] : tfn_alt[t], C);
      if (units_ok_start(C) && units_ok_end(C2)) c = from_units(C2, use); else use = -1;
    }
    used_be[t] = use;
    if (use < 0) {
      st.n_serial_tiles++;
      if (regular[t]) st.n_fallback++;
      for (int i = lo; i < hi; i++) { c = (i == 0) ? v[0] : c + v[i]; out[i] = c; }
    }
  }
  // phase D: in-tile scan of the maps
  for (int t = 0; t < ntiles; t++) {
    if (used_be[t] < 0) continue;
    int lo = t * tile, hi = std::min(n, lo + tile);
    i64 C0 = to_units(cstart[t], used_be[t]);
    ParFn f = parfn_identity();
    for (int i = lo; i < hi; i++) {
      f = parfn_compose(f, parfn_element(v[i], used_be[t]));
      out[i] = from_units(parfn_apply(f, C0), used_be[t]);
    }
  }
  return true;
}

static long check(const std::vector<double>& v, int tile, Stats& st, double perturb, const char* name) {
  std::vector<double> ref(v.size()), got;
  double run = 0;
  for (size_t i = 0; i < v.size(); i++) { run = (i == 0) ? v[0] : run + v[i]; ref[i] = run; }
  bool ok = exact_scan_emulated(v, tile, got, st, perturb);
  if (!ok) return 1;
  long bad = 0;
  for (size_t i = 0; i < v.size(); i++)
    if (memcmp(&ref[i], &got[i], 8) != 0) {
      if (bad < 3) fprintf(stderr, "[%s] mismatch at %zu: ref %.17g got %.17g\n", name, i, ref[i], got[i]);
      bad++;
    }
  return bad;
}

int main(int argc, char** argv) {
  int reps = argc > 1 ? atoi(argv[1]) : 20;
  long bad = 0;
  Stats st;
  // parity-map algebra: compose/apply agree with elementwise application
  for (int r = 0; r < 200000; r++) {
    int be = 1000 + (int)(next_u64() % 40);
    ParFn f = parfn_identity();
    i64 C = (i64)((1ull << 52) + (next_u64() >> 13));
    i64 Cseq = C;
    int k = 1 + (int)(next_u64() % 6);
    for (int j = 0; j < k; j++) {
      double p = ldexp(next_unit(), be - 1023 - 1 - (int)(next_u64() % 60));
      if (next_u64() % 4 == 0) {  // force a tie: p = (Q + 1/2) ulp
        u64 Q = next_u64() % 1000;
        p = ldexp((double)(2 * Q + 1), be - 1075 - 1);
      }
      ParFn e = parfn_element(p, be);
      Cseq = parfn_apply(e, Cseq);
      f = parfn_compose(f, e);
    }
    if (parfn_apply(f, C) != Cseq) { bad++; if (bad < 5) fprintf(stderr, "compose mismatch\n"); }
  }
  // single addition: map == real floating-point add
  for (int r = 0; r < 2000000; r++) {
    int be = 1 + (int)(next_u64() % 2040);
    u64 frac = next_u64() & 0xFFFFFFFFFFFFFull;
    double c = bits_d(((u64)be << 52) | frac);
    int down = (int)(next_u64() % 70);
    double p = ldexp(next_unit(), be - 1023 - down);
    if (r % 5 == 0) p = ldexp((double)(2 * (next_u64() % 4096) + 1), be - 1075 - 1);  // exact ties
    if (r % 11 == 0) p = 0.0;
    double s = c + p;
    if (!(s < ldexp(1.0, be - 1022)) && s != ldexp(1.0, be - 1022)) continue;  // left the binade
    i64 C2 = parfn_apply(parfn_element(p, be), to_units(c, be));
    if (!units_ok_end(C2)) { if (s == ldexp(1.0, be - 1022)) { /* allowed */ } else { bad++; continue; } }
    double got = from_units(C2, be);
    if (memcmp(&got, &s, 8) != 0) { bad++; if (bad < 5) fprintf(stderr, "add mismatch c=%a p=%a s=%a got=%a\n", c, p, s, got); }
  }
  const int sizes[] = {1, 2, 5, 31, 1000, 4096, 65536, 100003, 1 << 20};
  const int tiles[] = {32, 256, 1024};
  for (int rep = 0; rep < reps; rep++) {
    for (int n : sizes) {
      if (n == (1 << 20) && rep >= 3) continue;
      int tile = tiles[rep % 3];
      std::vector<double> v(n);
      // (1) normalised random weights (what the cumsum sees in the reference)
      double tot = 0;
      for (int i = 0; i < n; i++) { v[i] = next_unit(); tot += v[i]; }
      for (int i = 0; i < n; i++) v[i] /= tot;
      bad += check(v, tile, st, 0.0, "uniform");
      // (2) exp-distributed log-weights (particle-filter like), many near-zero
      tot = 0;
      for (int i = 0; i < n; i++) { double z = 6.0 * (next_unit() - 0.5); v[i] = exp(-0.5 * z * z * 9.0); tot += v[i]; }
      for (int i = 0; i < n; i++) v[i] /= tot;
      bad += check(v, tile, st, 0.0, "gauss");
      // (3) dyadic weights: sums hit powers of two exactly
      for (int i = 0; i < n; i++) v[i] = 1.0 / (double)(1 << 20);
      bad += check(v, tile, st, 0.0, "dyadic");
      // (4) raw (unnormalised) weights with a huge dynamic range, zeros and subnormals
      for (int i = 0; i < n; i++) {
        int k = (int)(next_u64() % 8);
        v[i] = k == 0 ? 0.0 : ldexp(next_unit(), -(int)(next_u64() % 1100));
        if (k == 1) v[i] = 4.9406564584124654e-324 * (double)(next_u64() % 5);
      }
      bad += check(v, tile, st, 0.0, "range");
      // (5) integers with forced ties: 2^53-scale sums
      for (int i = 0; i < n; i++) v[i] = (double)(next_u64() % 3) * 0.5 + (i == 0 ? 9007199254740992.0 / 4 : 0.0);
      bad += check(v, tile, st, 0.0, "ties");
      // (7) one dominant weight then a long tail of tiny ones: the cdf sits at 1 -/+ ulps for many tiles
      tot = 0;
      for (int i = 0; i < n; i++) { v[i] = (i < n / 8) ? next_unit() : 1e-22 * next_unit(); tot += v[i]; }
      for (int i = 0; i < n; i++) v[i] /= tot;
      { long f0 = st.n_fallback; bad += check(v, tile, st, 0.0, "tail");
        if (st.n_fallback - f0 > 2) { bad++; fprintf(stderr, "tail: %ld demoted tiles (n=%d)\n", st.n_fallback - f0, n); } }
      // (6) deliberately wrong approximate scan: verification must catch it or the result must still be exact
      tot = 0;
      for (int i = 0; i < n; i++) { v[i] = next_unit(); tot += v[i]; }
      for (int i = 0; i < n; i++) v[i] /= tot;
      bad += check(v, tile, st, 1e-3, "perturbed");
    }
  }
  printf("mismatches=%ld tiles=%ld serial_tiles=%ld demoted_regular_tiles=%ld\n", bad, st.n_tiles, st.n_serial_tiles, st.n_fallback);
  return bad == 0 ? 0 : 1;
}
