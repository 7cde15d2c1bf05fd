// CPU logic test of the persistent bootstrap-filter kernel: the kernel text of bayesssm_b200/csrc/bssm_fast.cuh
// (k_fast_bpf: the whole T loop in one cooperative launch, CTAs exchanging epoch-tagged records and particles
// through global memory) compiled by g++ over the SIMT emulation of tests/simt_emu.h -- every CTA of the group on its
// own OS thread, its threads as fibers -- and launched with the geometry fast_launch() in bssm_fast.cu computes.
// Prints loglike / n_resampled / status / early_exit / ess / state_est per filter; tests/test_fast_host.py compares
// them with the oracle's Philox-mode filter.
//
// usage: host_fast model variant G ngroups N T C resample_fn ralg threshold seed run_id stream_base [n_0 ... n_{C-1}]
//   (optional trailing particle counts: a ragged batch, FilterDev::n_per; N is then the maximum)
//                  < y (T doubles) theta (C x 3 doubles)
//   variant: 0 = <double, 8, HEADS>, 1 = <double, 8, scatter loops>, 2 = <float, 8, HEADS>, 3 = <float, 16, scatter loops>
#include "simt_emu.h"

#include "../bayesssm_b200/csrc/bssm_fast.cuh"

using namespace bssm;

template <typename Model, typename Real, int PPT, bool HEADS>
static int run(int argc, char** argv) {
  int a = 3;
  const int G = atoi(argv[a++]), ngroups_req = atoi(argv[a++]), N = atoi(argv[a++]), T = atoi(argv[a++]), C = atoi(argv[a++]);
  const int rfn = atoi(argv[a++]), ralg = atoi(argv[a++]);
  const double threshold = atof(argv[a++]);
  const unsigned long long seed = strtoull(argv[a++], nullptr, 10);
  const unsigned int run_id = (unsigned int)atoi(argv[a++]), stream_base = (unsigned int)atoi(argv[a++]);
  std::vector<int> n_per;
  if (argc >= a + C) for (int c = 0; c < C; c++) n_per.push_back(atoi(argv[a++]));
  std::vector<double> y(T), theta((size_t)C * 3);
  if (T && fread(y.data(), 8, T, stdin) != (size_t)T) return 2;
  if (fread(theta.data(), 8, theta.size(), stdin) != theta.size()) return 2;
  std::vector<unsigned int> stream(C), runid(C, run_id);
  for (int c = 0; c < C; c++) stream[c] = stream_base + c;
  std::vector<double> M(C, 0), S(C, 0), loglike(C, 0), ess((size_t)C * (T + 1), 0), se((size_t)C * (T + 1), 0), llh((size_t)C * std::max(T, 1), 0);
  std::vector<int> alive(C, 1), status(C, 0), early(C, 0), nres(C, 0);

  // optional observation times (gaps between them are extra transitions, R/particle_filter_core.R:70-71,124-136)
  std::vector<int> obs_times;
  if (const char* e = getenv("EMU_OBS_TIMES")) { for (const char* q = e; *q;) { obs_times.push_back((int)strtol(q, (char**)&q, 10)); if (*q == ',') q++; } }
  if ((int)obs_times.size() != T) obs_times.clear();

  // geometry: fast_launch() of bssm_fast.cu
  int nb_max = (N + G - 1) / G;
  nb_max = (nb_max + PPT - 1) / PPT * PPT;
  if (nb_max > FAST_MAX_NB) { fprintf(stderr, "slice too large\n"); return 2; }
  int threads = (nb_max / PPT + 31) / 32 * 32;
  if (threads < 32) threads = 32;
  const int cap = HEADS ? threads * fast_spt(PPT) : (nb_max + FAST_SLACK + 31) / 32 * 32;
  size_t smem = (size_t)((5 * G + 1) & ~1) * sizeof(double) + 5 * 32 * sizeof(double) + (size_t)cap * sizeof(Real) + (size_t)cap * sizeof(unsigned int);
  if (HEADS) smem += (size_t)cap * sizeof(unsigned int) + (size_t)threads * PPT * sizeof(Real);
  const int ngroups = std::max(1, std::min(ngroups_req, C));

  FastParams P;
  memset(&P, 0, sizeof(P));
  FilterDev& f = P.f;
  f.C = C; f.N = N; f.T = T; f.dy = 1; f.d = 1;
  f.theta = theta.data(); f.theta_stride = 3; f.y = y.data();
  f.stream = stream.data(); f.run_id = runid.data(); f.seed = seed;
  f.n_per = n_per.empty() ? nullptr : n_per.data();
  f.obs_times = obs_times.empty() ? nullptr : obs_times.data();
  f.M = M.data(); f.S = S.data(); f.loglike = loglike.data();
  f.alive = alive.data(); f.status = status.data(); f.early_exit = early.data(); f.n_resampled = nres.data();
  f.ess = ess.data(); f.state_est = se.data(); f.loglike_history = llh.data();
  f.algorithm = 0; f.ralg = ralg; f.threshold = threshold;
  P.G = G; P.ngroups = ngroups; P.resample_fn = rfn; P.nb_max = nb_max; P.cap = cap;
  std::vector<FastRec> rec((size_t)ngroups * 2 * G);
  memset((void*)rec.data(), 0, sizeof(FastRec) * rec.size());
  const size_t xbytes = (size_t)ngroups * G * nb_max * (sizeof(Real) == 4 ? 8 : 16);
  std::vector<unsigned long long> xnew(xbytes / 8, 0ull);   // exactly what fast_launch() allocates (AddressSanitizer runs rely on it)
  P.rec = rec.data(); P.xnew = xnew.data(); P.timing = nullptr;
  const FastParams Pc = P;
  emu_launch_cooperative((unsigned int)(ngroups * G), (unsigned int)threads, smem, [&] { k_fast_bpf<Model, Real, PPT, HEADS>(Pc); });
  for (int c = 0; c < C; c++) {
    printf("rank 0 filter %d loglike %.17g n_resampled %d status %d early_exit %d\n", c, loglike[c], nres[c], status[c], early[c]);
    printf("ess");
    for (int t = 0; t <= T; t++) printf(" %.17g", ess[(size_t)c * (T + 1) + t]);
    printf("\nstate_est");
    for (int t = 0; t <= T; t++) printf(" %.17g", se[(size_t)c * (T + 1) + t]);
    printf("\nloglike_history");
    for (int t = 0; t < T; t++) printf(" %.17g", llh[(size_t)c * T + t]);
    printf("\n");
  }
  return 0;
}

template <typename Model> static int by_variant(int argc, char** argv) {
  switch (atoi(argv[2])) {
    case 0: return run<Model, double, 8, true>(argc, argv);
    case 1: return run<Model, double, 8, false>(argc, argv);
    case 2: return run<Model, float, 8, true>(argc, argv);
    case 3: return run<Model, float, 16, false>(argc, argv);
  }
  return 2;
}

int main(int argc, char** argv) {
  if (argc < 14) { fprintf(stderr, "usage: see the header of tests/host_fast.cpp\n"); return 2; }
  switch (atoi(argv[1])) {
    case 0: return by_variant<ModelArSin>(argc, argv);
    case 1: return by_variant<ModelLG>(argc, argv);
    case 2: return by_variant<ModelRwDrift>(argc, argv);
    case 4: return by_variant<ModelArCos>(argc, argv);
  }
  return 2;
}
