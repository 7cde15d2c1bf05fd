// CPU logic test of the device-resident PMMH: the kernel text of bayesssm_b200/csrc/bssm_pmmh.cuh (one thread per
// chain: start, first draw, proposal, accept / reject, pilot statistics, replicate set-up, tuning) driven in the order
// bssm_pmmh_run() drives it -- pilot chain, pilot run, tuned main chain (R/pmmh_tuning.R:111-317, :29-64,
// R/pmmh.R:345-505) -- with every batched filter pass run by the persistent kernel's text (bssm_fast.cuh), all of it
// over the SIMT emulation of tests/simt_emu.h.  Bootstrap filter, nonlinear AR model (README.md:137-146), f64.
// Prints the chains; tests/test_pmmh_host.py compares them draw by draw with the oracle's orc_pmmh_chain.
//
// usage: host_pmmh C T pilot_n pilot_m pilot_reps m seed chain_id_base pilot_ralg pilot_rfn fixed_n G
//        < y[T] init[C][3] prior_kind[3] prior_a[3] prior_b[3] transform[3] pilot_sd[3]      (all doubles)
#include "simt_emu.h"

#include "../bayesssm_b200/csrc/bssm_fast.cuh"
#include "../bayesssm_b200/csrc/bssm_pmmh.cuh"

using namespace bssm;

// one batched filter pass: filter_reset() + fast_launch() of the library, restated for the emulation
struct FilterState {
  std::vector<double> sd, ess, se, llh;
  std::vector<int> si;
  void attach(FilterDev& f) {
    const size_t C = f.C, T = f.T;
    sd.assign(C * 4, 0.0); si.assign(C * 6, 0);
    ess.assign(C * (T + 1), 0.0); se.assign(C * (T + 1), 0.0); llh.assign(C * T, 0.0);
    f.M = sd.data(); f.S = sd.data() + C; f.loglike = sd.data() + 2 * C; f.cur_ess = sd.data() + 3 * C;
    f.alive = si.data(); f.resample = si.data() + C; f.status = si.data() + 2 * C; f.early_exit = si.data() + 3 * C;
    f.n_resampled = si.data() + 4 * C; f.cur = si.data() + 5 * C;
    f.ess = ess.data(); f.state_est = se.data(); f.loglike_history = llh.data();
  }
};
static void run_filters(const FilterDev& f, int resample_fn, const int* active, int G) {
  constexpr int PPT = 8;
  const size_t C = f.C, T = f.T;
  std::fill(f.M, f.M + 4 * C, 0.0);
  std::fill(f.alive, f.alive + 6 * C, 0);
  for (size_t c = 0; c < C; c++) f.alive[c] = active[c];
  std::fill(f.ess, f.ess + C * (T + 1), 0.0);
  std::fill(f.state_est, f.state_est + C * (T + 1), 0.0);
  std::fill(f.loglike_history, f.loglike_history + C * T, 0.0);
  int nb_max = (f.N + G - 1) / G;
  nb_max = (nb_max + PPT - 1) / PPT * PPT;
  int threads = (nb_max / PPT + 31) / 32 * 32;
  if (threads < 32) threads = 32;
  const int cap = threads * fast_spt(PPT);
  const size_t smem = (size_t)((5 * G + 1) & ~1) * sizeof(double) + 5 * 32 * sizeof(double) + (size_t)cap * sizeof(double) +
                      (size_t)cap * sizeof(unsigned int) + (size_t)cap * sizeof(unsigned int) + (size_t)threads * PPT * sizeof(double);
  const int ngroups = (int)std::min<size_t>(C, 4);
  FastParams P;
  memset(&P, 0, sizeof(P));
  P.f = f; P.G = G; P.ngroups = ngroups; P.resample_fn = resample_fn; P.nb_max = nb_max; P.cap = cap;
  std::vector<FastRec> rec((size_t)ngroups * 2 * G);
  memset((void*)rec.data(), 0, sizeof(FastRec) * rec.size());
  std::vector<unsigned long long> xnew((size_t)ngroups * G * nb_max * 2, 0ull);
  P.rec = rec.data(); P.xnew = xnew.data(); P.timing = nullptr;
  const FastParams Pc = P;
  emu_launch_cooperative((unsigned int)(ngroups * G), (unsigned int)threads, smem, [&] { k_fast_bpf<ModelArSin, double, 8, true>(Pc); });
}

int main(int argc, char** argv) {
  if (argc < 13) { fprintf(stderr, "usage: see the header of tests/host_pmmh.cpp\n"); return 2; }
  int a = 1;
  const int C = atoi(argv[a++]), T = atoi(argv[a++]), pilot_n = atoi(argv[a++]), pm = atoi(argv[a++]), reps = atoi(argv[a++]), m = atoi(argv[a++]);
  const unsigned long long seed = strtoull(argv[a++], nullptr, 10);
  const unsigned int chain_id_base = (unsigned int)atoi(argv[a++]);
  const int pilot_ralg = atoi(argv[a++]), pilot_rfn = atoi(argv[a++]), fixed_n = atoi(argv[a++]), G = atoi(argv[a++]);
  const int p = 3, ts = 3;
  std::vector<double> y(T), init((size_t)C * p), cfgv(15);
  if (fread(y.data(), 8, T, stdin) != (size_t)T || fread(init.data(), 8, init.size(), stdin) != init.size() ||
      fread(cfgv.data(), 8, 15, stdin) != 15) return 2;

  std::vector<double> cur((size_t)C * p), prop((size_t)C * p), cur_ll(C), lp_prop(C), theta_full((size_t)C * ts),
      pilot_chain((size_t)C * pm * p), pilot_ll((size_t)C * pm), mean((size_t)C * p), cov((size_t)C * p * p), chol((size_t)C * p * p),
      chain((size_t)C * m * p), ll_chain((size_t)C * m), theta_rep((size_t)C * reps * ts), rep_ll((size_t)C * reps);
  std::vector<int> valid(C, 0), alive(C, 0), status(C, 0), n_accept(C, 0), target_n(C, 0), active_rep((size_t)C * reps, 0), moved(C, 0);
  std::vector<unsigned int> ids((size_t)2 * C), ids_rep((size_t)2 * C * reps);

  PmmhDev P;
  memset(&P, 0, sizeof(P));
  P.C = C; P.p = p; P.nconst = 0; P.theta_stride = ts; P.seed = seed; P.chain_id_base = chain_id_base;
  for (int j = 0; j < p; j++) {
    P.prior_kind[j] = (int)cfgv[j]; P.prior_a[j] = cfgv[3 + j]; P.prior_b[j] = cfgv[6 + j];
    P.transform[j] = (int)cfgv[9 + j]; P.pilot_sd[j] = cfgv[12 + j];
  }
  P.cur = cur.data(); P.prop = prop.data(); P.cur_ll = cur_ll.data(); P.lp_prop = lp_prop.data(); P.theta_full = theta_full.data();
  P.valid = valid.data(); P.alive = alive.data(); P.status = status.data(); P.n_accept = n_accept.data(); P.moved = moved.data();
  P.stream = ids.data(); P.run_id = ids.data() + C;

  FilterDev f;
  memset(&f, 0, sizeof(f));
  f.C = C; f.T = T; f.dy = 1; f.d = 1; f.theta = theta_full.data(); f.theta_stride = ts; f.y = y.data();
  f.stream = P.stream; f.run_id = P.run_id; f.seed = seed; f.algorithm = 0; f.threshold = -1.0;
  FilterState fs, fsr;
  const unsigned int gb = (unsigned int)(C + 127) / 128;
  auto launch = [&](unsigned int grid, auto body) { emu_launch(grid, 128, body); };

  // ---- pilot chain (R/pmmh_tuning.R:111-317) ----
  f.N = pilot_n; f.n_per = nullptr; f.ralg = pilot_ralg;
  fs.attach(f);
  P.f_loglike = f.loglike; P.f_status = f.status;
  { const PmmhDev Pc = P; launch(gb, [&] { k_pm_start(Pc, init.data(), PH_PILOT, 1, 1); }); }
  run_filters(f, pilot_rfn, P.valid, G);
  { const PmmhDev Pc = P; launch(gb, [&] { k_pm_first(Pc, pilot_chain.data(), pilot_ll.data(), pm); }); }
  for (int it = 1; it < pm; it++) {
    const PmmhDev Pc = P;
    launch(gb, [&] { k_pm_propose(Pc, PH_PILOT, it, nullptr); });
    run_filters(f, pilot_rfn, P.valid, G);
    launch(gb, [&] { k_pm_accept(Pc, PH_PILOT, it, pilot_chain.data(), pilot_ll.data(), pm); });
  }
  { const PmmhDev Pc = P; launch(gb, [&] { k_pm_pilot_stats(Pc, pilot_chain.data(), pm, mean.data(), cov.data()); }); }
  // ---- .pilot_run (R/pmmh_tuning.R:29-64): reps replicate filters per chain at the pilot mean, with the pilot's resampler settings ----
  {
    FilterDev fr = f;
    fr.C = C * reps; fr.theta = theta_rep.data(); fr.stream = ids_rep.data(); fr.run_id = ids_rep.data() + (size_t)C * reps; fr.ralg = pilot_ralg;
    fsr.attach(fr);
    const PmmhDev Pc = P;
    launch((unsigned int)(C * reps + 127) / 128, [&] { k_pm_reps_setup(Pc, mean.data(), reps, theta_rep.data(), ids_rep.data(), ids_rep.data() + (size_t)C * reps, active_rep.data()); });
    run_filters(fr, pilot_rfn, active_rep.data(), G);
    launch(gb, [&] { k_pm_tune(Pc, fr.loglike, fr.status, reps, pilot_n, fixed_n, mean.data(), cov.data(), rep_ll.data(), target_n.data(), chol.data()); });
  }
  // ---- main chain (R/pmmh.R:395-500); SISAR + stratified whatever the caller asked for (quirk A10) ----
  int nmax = 1;
  for (int c = 0; c < C; c++) nmax = std::max(nmax, target_n[c]);
  bool ragged = false;
  for (int c = 0; c < C; c++) ragged = ragged || target_n[c] != nmax;
  f.N = nmax; f.n_per = ragged ? target_n.data() : nullptr; f.ralg = 2;
  fs.attach(f);
  P.f_loglike = f.loglike; P.f_status = f.status;
  { const PmmhDev Pc = P; launch(gb, [&] { k_pm_start(Pc, mean.data(), PH_MAIN, 0, 0); }); }
  run_filters(f, 0, P.valid, G);
  { const PmmhDev Pc = P; launch(gb, [&] { k_pm_first(Pc, chain.data(), ll_chain.data(), m); }); }
  for (int it = 1; it < m; it++) {
    const PmmhDev Pc = P;
    launch(gb, [&] { k_pm_propose(Pc, PH_MAIN, it, chol.data()); });
    run_filters(f, 0, P.valid, G);
    launch(gb, [&] { k_pm_accept(Pc, PH_MAIN, it, chain.data(), ll_chain.data(), m); });
  }

  auto row = [](const char* name, const double* v, size_t n) { printf("%s", name); for (size_t i = 0; i < n; i++) printf(" %.17g", v[i]); printf("\n"); };
  for (int c = 0; c < C; c++) {
    printf("chain %d status %d target_n %d n_accept %d\n", c, status[c], target_n[c], n_accept[c]);
    row("pilot_theta_chain", pilot_chain.data() + (size_t)c * pm * p, (size_t)pm * p);
    row("pilot_loglike_chain", pilot_ll.data() + (size_t)c * pm, pm);
    row("pilot_theta_mean", mean.data() + (size_t)c * p, p);
    row("pilot_theta_cov", cov.data() + (size_t)c * p * p, (size_t)p * p);
    row("pilot_loglikes", rep_ll.data() + (size_t)c * reps, reps);
    row("proposal_chol", chol.data() + (size_t)c * p * p, (size_t)p * p);
    row("theta_chain", chain.data() + (size_t)c * m * p, (size_t)m * p);
    row("loglike_chain", ll_chain.data() + (size_t)c * m, m);
  }
  return 0;
}
