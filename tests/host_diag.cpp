// Host logic test for bayesssm_b200/csrc/bssm_diag.cuh (no GPU needed): runs the per-thread bodies of the five
// diagnostics kernels for every thread index the launchers would create, in launch order, on draws read from stdin
// ([k][m_total][p] doubles, raw), and prints ess / rhat / flags per parameter.  tests/test_diag_host.py compares
// the output with the numpy restatement of R/ESS.R and R/rhat.R in oracle/mcmc_diag.py.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../bayesssm_b200/csrc/bssm_diag.cuh"

using namespace bssm;

int main(int argc, char** argv) {
  if (argc < 6) { fprintf(stderr, "usage: host_diag k m_total p burn_in want_ess < draws\n"); return 2; }
  const int k = atoi(argv[1]), m_total = atoi(argv[2]), p = atoi(argv[3]), burn = atoi(argv[4]), want_ess = atoi(argv[5]);
  const int m = m_total - burn;
  std::vector<double> x((size_t)k * m_total * p);
  if (fread(x.data(), sizeof(double), x.size(), stdin) != x.size()) { fprintf(stderr, "short read\n"); return 2; }
  std::vector<double> xc((size_t)p * k * m), mom((size_t)p * k * 6), acov((size_t)p * k * m), par((size_t)p * 2),
      rho((size_t)p * m), ess(p), rhat(p);
  std::vector<int> flags(p);
  DiagArgs a;
  a.x = x.data() + (size_t)burn * p;
  a.chain_stride = (long long)m_total * p;
  a.iter_stride = p;
  a.k = k; a.m = m; a.p = p;
  a.xc = xc.data(); a.mom = mom.data(); a.acov = acov.data(); a.par = par.data(); a.rho = rho.data();
  a.ess = want_ess ? ess.data() : nullptr;
  a.rhat = rhat.data(); a.flags = flags.data();
  // threads of one launch are independent, so any order is a valid schedule; run them backwards to catch a
  // body that silently relies on a lower thread index having run first
  for (long long t = diag_n_moments(a) - 1; t >= 0; t--) diag_moments_thread(a, t);
  if (want_ess) for (long long t = diag_n_acov(a) - 1; t >= 0; t--) diag_acov_thread(a, t);
  for (long long t = p - 1; t >= 0; t--) diag_between_thread(a, t);
  if (want_ess) {
    for (long long t = diag_n_rho(a) - 1; t >= 0; t--) diag_rho_thread(a, t);
    for (long long t = p - 1; t >= 0; t--) diag_geyer_thread(a, t);
  }
  for (int j = 0; j < p; j++) printf("%.17g %.17g %d\n", want_ess ? ess[j] : 0.0, rhat[j], flags[j]);
  return 0;
}
