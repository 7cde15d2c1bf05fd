// CPU logic test of the general engine: the kernel text of bayesssm_b200/csrc/bssm_filter.cuh (k_init, k_weight,
// k_finalize, k_search_gather, k_flip, k_post) and bssm_resample.cuh (the cdf pipeline k_tile_sums / k_tile_scan / k_chain
// / k_tile_exact with the bit-exact sequential cumsum of bssm_exact.cuh) compiled by g++ over the SIMT emulation of
// tests/simt_emu.h and driven as run_filter_steps() / resample_stage() / resample_cdf() in bssm_engine.cu drive it:
// bootstrap, auxiliary and resample-move filters, every built-in model, Philox noise, f64 with exact resampling.
// Prints loglike / n_resampled / status / early_exit / ess / state_est per filter; tests/test_general_host.py compares them
// with the oracle's Philox-mode filter.
//
// usage: host_general model algorithm N T C resample_fn ralg threshold seed run_id stream_base exact
//        < y[T] theta[C][NTHETA + NCONST]     (all doubles)
#include "simt_emu.h"

#include "../include/bayesssm_b200.h"
#include "../bayesssm_b200/csrc/bssm_filter.cuh"
#include "../bayesssm_b200/csrc/bssm_resample.cuh"

using namespace bssm;

struct RsArgs {   // as in bssm_engine.cuh
  int nseg, n; const int* n_per; const int* enable; double* cdf; size_t cdf_stride; int* status; int validate; int exact;
  long long* n_serial; double* total;
};
static void k_zero_sum_check(const double* total, int nseg, int* status, const int* enable) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseg || !seg_on(enable, s)) return;
  if (status && status[s] == 0 && total[s] == 0.0) status[s] = BSSM_ERR_ZERO_SUM;
}

// resample_cdf() of bssm_engine.cu
template <typename Src, typename MakeNorm>
static void resample_cdf(const Src& src, MakeNorm make_norm, const RsArgs& a) {
  const int ntiles = (a.n + RS_TILE - 1) / RS_TILE;
  std::vector<double> part((size_t)a.nseg * ntiles), cstart((size_t)a.nseg * ntiles), total_own(a.nseg);
  std::vector<TileRec> rec((size_t)a.nseg * ntiles);
  std::vector<int> used((size_t)a.nseg * ntiles);
  double* total = a.total ? a.total : total_own.data();
  // large inputs scan the tile totals once per pass (k_tile_prefix); EMU_RS_PREFIX_TILES moves the threshold, as BSSM_RS_PREFIX_TILES does
  const int prefix_from = getenv("EMU_RS_PREFIX_TILES") ? atoi(getenv("EMU_RS_PREFIX_TILES")) : RS_PREFIX_TILES;
  std::vector<double> pref_v(ntiles > prefix_from ? (size_t)a.nseg * ntiles : 0, -7.0);
  double* const pref = pref_v.empty() ? nullptr : pref_v.data();
  auto tile_prefix = [&] { if (pref) emu_launch(a.nseg, 256, [&] { k_tile_prefix(part.data(), a.n, a.n_per, ntiles, pref, a.enable); }); };
  emu_launch2d(a.nseg, ntiles, RS_THREADS, [&] { k_tile_sums<Src>(src, a.n, a.n_per, ntiles, part.data(), a.status, a.validate, a.enable); });
  tile_prefix();
  if (!a.exact) {
    emu_launch2d(a.nseg, ntiles, RS_THREADS, [&] { k_tile_scan<Src>(src, a.n, a.n_per, ntiles, part.data(), a.cdf, a.cdf_stride, rec.data(), 0, a.enable, pref); });
    return;
  }
  emu_launch2d(a.nseg, ntiles, RS_THREADS, [&] { k_tile_scan<Src>(src, a.n, a.n_per, ntiles, part.data(), nullptr, 0, rec.data(), 1, a.enable, pref); });
  emu_launch(a.nseg, 32, [&] { k_chain<Src>(src, a.n, a.n_per, ntiles, rec.data(), cstart.data(), used.data(), total, nullptr, 0, nullptr, a.enable); });
  if (a.status) emu_launch((a.nseg + 127) / 128, 128, [&] { k_zero_sum_check(total, a.nseg, a.status, a.enable); });
  auto srcn = make_norm(total);
  typedef decltype(srcn) SrcN;
  emu_launch2d(a.nseg, ntiles, RS_THREADS, [&] { k_tile_sums<SrcN>(srcn, a.n, a.n_per, ntiles, part.data(), nullptr, 0, a.enable); });
  tile_prefix();
  emu_launch2d(a.nseg, ntiles, RS_THREADS, [&] { k_tile_scan<SrcN>(srcn, a.n, a.n_per, ntiles, part.data(), nullptr, 0, rec.data(), 1, a.enable, pref); });
  emu_launch(a.nseg, 32, [&] { k_chain<SrcN>(srcn, a.n, a.n_per, ntiles, rec.data(), cstart.data(), used.data(), part.data(), a.cdf, a.cdf_stride, a.n_serial, a.enable); });
  emu_launch2d(a.nseg, ntiles, RS_THREADS, [&] { k_tile_exact<SrcN>(srcn, a.n, a.n_per, ntiles, cstart.data(), used.data(), a.cdf, a.cdf_stride, a.enable); });
}

// resample_stage() of bssm_engine.cu
template <typename Real>
static void resample_stage(FilterDev& f, int resample_fn, int exact, int obs, int aux_stage, double* cdf) {
  RsArgs a;
  a.nseg = f.C; a.n = f.N; a.n_per = f.n_per; a.enable = f.resample; a.cdf = cdf; a.cdf_stride = (size_t)f.N;
  a.status = nullptr; a.validate = 0; a.exact = exact; a.n_serial = nullptr; a.total = nullptr;
  const Real* lw = (const Real*)(aux_stage ? f.lw_aux : f.lw);
  const double *M = f.M, *S = f.S;
  size_t stride = (size_t)f.N;
  SrcLogW<Real> src{lw, stride, M, S};
  resample_cdf(src, [=](const double* total) { return SrcLogWNorm<Real>{lw, stride, M, S, total}; }, a);
  USrcFilter us;
  us.buf = nullptr; us.N = f.N; us.injected = 0; us.seed = f.seed; us.run_id = f.run_id; us.stream = f.stream;
  us.tag = aux_stage ? TAG_RESAMP_AUX_U : TAG_RESAMP_U; us.obs = obs;
  const FilterDev fc = f;
  emu_launch2d(f.C, f.nblk, FT_THREADS, [&] { k_search_gather<Real>(fc, us, resample_fn, obs, aux_stage, cdf); });
  emu_launch((f.C + 127) / 128, 128, [&] { k_flip(fc); });
}

template <typename Model>
static int run(char** argv) {
  typedef double Real;
  int a = 2;
  const int algorithm = atoi(argv[a++]), N = atoi(argv[a++]), T = atoi(argv[a++]), C = atoi(argv[a++]);
  const int rfn = atoi(argv[a++]), ralg = atoi(argv[a++]);
  const double threshold = atof(argv[a++]);
  const unsigned long long seed = strtoull(argv[a++], nullptr, 10);
  const unsigned int run_id = (unsigned int)atoi(argv[a++]), stream_base = (unsigned int)atoi(argv[a++]);
  const int exact = atoi(argv[a++]);
  const int d = Model::D, ts = Model::NTHETA + Model::NCONST;
  std::vector<double> y(T), theta((size_t)C * ts);
  if (T && fread(y.data(), 8, T, stdin) != (size_t)T) return 2;
  if (fread(theta.data(), 8, theta.size(), stdin) != theta.size()) return 2;
  std::vector<unsigned int> stream(C), runid(C, run_id);
  for (int c = 0; c < C; c++) stream[c] = stream_base + c;

  // optional observation times (gaps between them are extra transitions, R/particle_filter_core.R:70-71,124-136)
  std::vector<int> obs_times;
  if (const char* e = getenv("EMU_OBS_TIMES")) { for (const char* q = e; *q;) { obs_times.push_back((int)strtol(q, (char**)&q, 10)); if (*q == ',') q++; } }
  if ((int)obs_times.size() != T) obs_times.clear();

  // filter_setup() / filter_reset() of bssm_engine.cu
  FilterDev f;
  memset(&f, 0, sizeof(f));
  f.C = C; f.N = N; f.T = T; f.dy = 1; f.d = d; f.theta = theta.data(); f.theta_stride = ts; f.y = y.data();
  f.carry = getenv("EMU_CARRY") ? 1 : 0;        // carried-weights mode (bssm_filter_config::carry_weights)
  f.obs_times = obs_times.empty() ? nullptr : obs_times.data();
  f.stream = stream.data(); f.run_id = runid.data(); f.seed = seed; f.algorithm = algorithm; f.ralg = ralg; f.threshold = threshold;
  f.nblk = std::max(1, std::min(1024, (N + FT_THREADS - 1) / FT_THREADS));
  std::vector<Real> xa((size_t)C * d * N, NAN), xb((size_t)C * d * N, NAN), lw((size_t)C * N, NAN), lw_aux((size_t)C * N, NAN), auxg((size_t)C * N, NAN);
  std::vector<double> part((size_t)C * f.nblk * PART_W, NAN), cdf((size_t)C * N, NAN), sd((size_t)C * 4, 0.0), ess((size_t)C * (T + 1), 0.0),
      se((size_t)C * (T + 1) * d, 0.0), llh((size_t)C * std::max(T, 1), 0.0);
  std::vector<int> si((size_t)C * 6, 0);
  f.xa = xa.data(); f.xb = xb.data(); f.lw = lw.data(); f.lw_aux = lw_aux.data(); f.auxg = auxg.data(); f.part = part.data();
  f.M = sd.data(); f.S = sd.data() + C; f.loglike = sd.data() + 2 * C; f.cur_ess = sd.data() + 3 * C;
  f.alive = si.data(); f.resample = si.data() + C; f.status = si.data() + 2 * C; f.early_exit = si.data() + 3 * C;
  f.n_resampled = si.data() + 4 * C; f.cur = si.data() + 5 * C;
  for (int c = 0; c < C; c++) f.alive[c] = 1;
  f.ess = ess.data(); f.state_est = se.data(); f.loglike_history = llh.data();

  // run_filter_steps() of bssm_engine.cu
  if ((algorithm == BSSM_APF && !Model::HAS_AUX) || (algorithm == BSSM_RMPF && !Model::HAS_MOVE)) return 3;
  const FilterDev fc = f;
  auto finalize = [&](int obs, int kind) { emu_launch(C, 128, [&] { k_finalize<sizeof(Real) == 4>(fc, obs, kind); }); };
  auto weight = [&](int obs, int flags, int wkind) { emu_launch2d(C, f.nblk, FT_THREADS, [&] { k_weight<Model, Real>(fc, obs, flags, wkind); }); };
  emu_launch2d(C, f.nblk, FT_THREADS, [&] { k_init<Model, Real>(fc); });
  finalize(0, 0);
  const bool may_resample = algorithm == BSSM_RMPF || ralg != BSSM_SIS;
  for (int obs = 0; obs < T; obs++) {
    if (algorithm == BSSM_APF) {
      weight(obs, WF_GAP, 1);
      finalize(obs, 2);
      resample_stage<Real>(f, rfn, exact, obs, 1, cdf.data());
      weight(obs, WF_SECOND, 2);
    } else {
      weight(obs, WF_GAP, 0);
    }
    finalize(obs, 1);
    if (may_resample) {
      resample_stage<Real>(f, rfn, exact, obs, 0, cdf.data());
      emu_launch2d(C, f.nblk, FT_THREADS, [&] { k_post<Model, Real>(fc, obs); });
      finalize(obs, 3);
    }
  }
  for (int c = 0; c < C; c++) {
    printf("rank 0 filter %d loglike %.17g n_resampled %d status %d early_exit %d\n", c, f.loglike[c], f.n_resampled[c], f.status[c], f.early_exit[c]);
    printf("ess");
    for (int t = 0; t <= T; t++) printf(" %.17g", ess[(size_t)c * (T + 1) + t]);
    printf("\nstate_est");
    for (int t = 0; t <= T; t++) for (int k = 0; k < d; k++) printf(" %.17g", se[((size_t)c * (T + 1) + t) * d + k]);
    printf("\nloglike_history");
    for (int t = 0; t < T; t++) printf(" %.17g", llh[(size_t)c * T + t]);
    printf("\n");
  }
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 13) { fprintf(stderr, "usage: see the header of tests/host_general.cpp\n"); return 2; }
  switch (atoi(argv[1])) {
    case BSSM_MODEL_AR_SIN: return run<ModelArSin>(argv);
    case BSSM_MODEL_LG: return run<ModelLG>(argv);
    case BSSM_MODEL_RW_DRIFT: return run<ModelRwDrift>(argv);
    case BSSM_MODEL_SIR_CB: return run<ModelSirCB>(argv);
    case BSSM_MODEL_AR_COS: return run<ModelArCos>(argv);
    case BSSM_MODEL_RW2D: return run<ModelRw2D>(argv);
    case BSSM_MODEL_SIR_GILLESPIE: return run<ModelSirGillespie>(argv);
  }
  return 2;
}
