// CPU logic test of the streaming bootstrap-filter engine: the kernel text of bayesssm_b200/csrc/bssm_stream.cuh
// (k_st_setup, k_st_init, k_st_step, k_st_merge, k_st_resample, k_st_flush, k_st_flush_merge) compiled by g++ over
// the SIMT emulation of tests/simt_emu.h and driven exactly as stream_launch() in bssm_stream.cu drives it --
// including the particle-sharded form, where `world` emulated ranks run side by side and the per-observation
// ncclAllGather of the 64-byte records is a memcpy.  Prints loglike / n_resampled / status / early_exit / ess /
// state_est per filter; tests/test_stream_host.py compares them with the oracle's Philox-mode filter.
//
// usage: host_stream model precision threads N T C bpc resample_fn ralg threshold seed run_id stream_base world
//                    capacity_factor block_order [n_0 ... n_{C-1}] < y (T doubles) theta (C x 3 doubles)
//   the optional trailing particle counts make the batch ragged (FilterDev::n_per; N is then the stride / maximum)
#include "simt_emu.h"

#include "../bayesssm_b200/csrc/bssm_stream.cuh"

#include <algorithm>
#include <random>

using namespace bssm;

struct Rank {
  StreamParams P;
  std::vector<unsigned char> x0, x1;
  std::vector<double> pref, bsum, blk, M, S, loglike, ess, state_est, llh, mn_pos, mn_tsum, mn_total;
  std::vector<unsigned int> counter, epoch, bar2;
  std::vector<int> res, alive, status, early, nres;
  std::vector<StSeg> seg;
  std::vector<StRec> rec_local, rec_all;
  long long goff0; int nloc0;
};

template <typename Model, typename Real, int PPT, int THREADS>
static int run(int argc, char** argv) {
  constexpr int TS = THREADS * PPT;
  int a = 4;
  const int N = atoi(argv[a++]), T = atoi(argv[a++]), C = atoi(argv[a++]), bpc_req = atoi(argv[a++]);
  const int rfn = atoi(argv[a++]), ralg = atoi(argv[a++]);
  const double threshold = atof(argv[a++]);
  const unsigned long long seed = strtoull(argv[a++], nullptr, 10);
  const unsigned int run_id = (unsigned int)atoi(argv[a++]), stream_base = (unsigned int)atoi(argv[a++]);
  const int world = atoi(argv[a++]);
  const double capf = atof(argv[a++]);
  const int order_mode = atoi(argv[a++]);   // 0 ascending, 1 descending, 2 shuffled block order
  std::vector<int> n_per;
  if (argc >= a + C) for (int c = 0; c < C; c++) n_per.push_back(atoi(argv[a++]));
  std::vector<double> y(T), theta((size_t)C * 3);
  if (T && fread(y.data(), 8, T, stdin) != (size_t)T) return 2;
  if (fread(theta.data(), 8, theta.size(), stdin) != theta.size()) return 2;
  std::vector<unsigned int> stream(C), runid(C, run_id);
  for (int c = 0; c < C; c++) stream[c] = stream_base + c;

  // optional observation times (gaps between them are extra transitions, R/particle_filter_core.R:70-71,124-136)
  std::vector<int> obs_times;
  if (const char* e = getenv("EMU_OBS_TIMES")) { for (const char* q = e; *q;) { obs_times.push_back((int)strtol(q, (char**)&q, 10)); if (*q == ',') q++; } }
  if ((int)obs_times.size() != T) obs_times.clear();
  std::vector<Rank> ranks(world);
  const bool sharded = world > 1;
  for (int g = 0; g < world; g++) {
    Rank& R = ranks[g];
    StreamParams& P = R.P;
    memset(&P, 0, sizeof(P));
    // the partition and capacity of bssm_shard.cu (shard_partition, bssm_filter_run_sharded): whole Philox quads per rank
    const long long quads = ((long long)N + 3) / 4;
    const long long qa = std::min<long long>(N, quads * g / world * 4), qb = std::min<long long>(N, quads * (g + 1) / world * 4);
    R.nloc0 = sharded ? (int)(qb - qa) : N;
    R.goff0 = sharded ? qa : 0;
    const int cap = sharded ? (int)std::min<double>((double)N, capf * ((double)N / world) + 1024.0) : N;
    R.M.assign(C, 0); R.S.assign(C, 0); R.loglike.assign(C, 0);
    R.ess.assign((size_t)C * (T + 1), 0); R.state_est.assign((size_t)C * (T + 1), 0); R.llh.assign((size_t)C * std::max(T, 1), 0);
    R.alive.assign(C, 1); R.status.assign(C, 0); R.early.assign(C, 0); R.nres.assign(C, 0);
    FilterDev& f = P.f;
    f.C = C; f.N = sharded ? cap : N; f.T = T; f.dy = 1; f.d = 1;
    f.theta = theta.data(); f.theta_stride = 3; f.y = y.data();
    f.stream = stream.data(); f.run_id = runid.data(); f.seed = seed;
    f.n_per = n_per.empty() ? nullptr : n_per.data();
    f.obs_times = obs_times.empty() ? nullptr : obs_times.data();
    f.M = R.M.data(); f.S = R.S.data(); f.loglike = R.loglike.data();
    f.alive = R.alive.data(); f.status = R.status.data(); f.early_exit = R.early.data(); f.n_resampled = R.nres.data();
    f.ess = R.ess.data(); f.state_est = R.state_est.data(); f.loglike_history = R.llh.data();
    f.algorithm = 0; f.ralg = ralg; f.threshold = threshold;
    P.resample_fn = rfn;
    st_fill_round_keys(P);
    P.sharded = sharded; P.rank = g; P.world = world; P.n_glob = sharded ? N : 0;
    P.cap = cap;
    P.log_n = n_per.empty() ? log((double)N) : nan("");
    P.nt = (cap + 4 + TS - 1) / TS;
    P.xstride = (size_t)P.nt * TS;
    P.bpc = std::max(1, std::min(bpc_req, P.nt));
    R.x0.assign((size_t)C * P.xstride * sizeof(Real), 0xFF);   // NaN-ish garbage: nothing may rely on zeroed storage
    R.x1.assign((size_t)C * P.xstride * sizeof(Real), 0xFF);
    P.x0 = R.x0.data(); P.x1 = R.x1.data();
    R.pref.assign((size_t)C * (P.bpc + 1), -1.0); R.bsum.assign((size_t)C * P.bpc, -1.0); R.blk.assign((size_t)C * P.bpc * 4, -1.0);
    P.pref = R.pref.data(); P.bsum = R.bsum.data();
    P.blk_m = R.blk.data(); P.blk_s = P.blk_m + (size_t)C * P.bpc; P.blk_q = P.blk_s + (size_t)C * P.bpc; P.blk_x = P.blk_q + (size_t)C * P.bpc;
    R.counter.assign(C, 77u); R.res.assign((size_t)2 * C, -1); R.seg.resize((size_t)2 * C);
    P.counter = R.counter.data(); P.res = R.res.data(); P.seg = R.seg.data();
    if (getenv("EMU_CHAIN")) { R.epoch.assign(C, 99u); R.bar2.assign(C, 98u); P.epoch = R.epoch.data(); P.bar2 = R.bar2.data(); }   // stream_launch(): only for k_st_chain
    R.rec_local.resize(C); R.rec_all.resize((size_t)C * world);
    P.rec_local = R.rec_local.data(); P.rec_all = R.rec_all.data();
    P.dbg = nullptr;
    if (rfn == 2) {   // multinomial: positions of the output slots + the scan of the spacings (stream_launch, bssm_stream.cu)
      P.mn_nt = (cap + 1 + MN_TILE - 1) / MN_TILE + 1;
      P.mn_ahead = getenv("EMU_MN_AHEAD") ? 1 : 0;      // one GPU: the arrays doubled by parity, laid out for every observation
      const size_t rows = P.mn_ahead ? (size_t)2 * C : (size_t)C;
      R.mn_pos.assign(rows * P.xstride, -1.0); R.mn_tsum.assign(rows * P.mn_nt, -1.0); R.mn_total.assign(rows, -1.0);
      P.mn_pos = R.mn_pos.data(); P.mn_tsum = R.mn_tsum.data(); P.mn_total = R.mn_total.data();
    }
  }
  std::mt19937 shuf(12345);
  auto block_order = [&](unsigned int grid) {
    std::vector<unsigned int> o(grid);
    for (unsigned int i = 0; i < grid; i++) o[i] = order_mode == 1 ? grid - 1 - i : i;
    if (order_mode == 2) std::shuffle(o.begin(), o.end(), shuf);
    return o;
  };
  auto allgather = [&]() {   // shard_allgather(): rec_all[g][c] on every rank
    for (int r = 0; r < world; r++)
      for (int g = 0; g < world; g++) memcpy(ranks[r].rec_all.data() + (size_t)g * C, ranks[g].rec_local.data(), sizeof(StRec) * C);
  };
  for (int g = 0; g < world; g++) {
    Rank& R = ranks[g];
    const StreamParams P = R.P;
    emu_launch((C + 127) / 128, 128, [&] { k_st_setup(P, R.goff0, R.nloc0); });
    const unsigned int grid = (unsigned int)P.bpc * C;
    const auto o = block_order(grid);
    emu_launch(grid, THREADS, [&] { k_st_init<Model, Real, PPT, THREADS>(P); }, &o);
  }
  // chain-persistent kernel (stream_launch(): batches on one GPU): cooperative launches over groups of filters, every
  // observation inside; EMU_CHAIN = filters per launch
  const int chain_group = getenv("EMU_CHAIN") ? std::max(1, atoi(getenv("EMU_CHAIN"))) : 0;
  if (chain_group && T > 0) {
    if (sharded || rfn == 2) { fprintf(stderr, "EMU_CHAIN: one rank, stratified / systematic only\n"); return 2; }
    const StreamParams P = ranks[0].P;
    for (int c0 = 0; c0 < C; c0 += chain_group) {
      const int cb = std::min(chain_group, C - c0);
      emu_launch_cooperative((unsigned int)(cb * P.bpc), THREADS, 0, [&] { k_st_chain<Model, Real, PPT, THREADS>(P, c0, T); });
    }
  }
  for (int obs = 0; obs < T && !chain_group; obs++) {
    for (int g = 0; g < world; g++) {
      const StreamParams P = ranks[g].P;
      const auto o = block_order((unsigned int)P.bpc * C);
      emu_launch((unsigned int)P.bpc * C, THREADS, [&] { k_st_step<Model, Real, PPT, THREADS>(P, obs); }, &o);
    }
    if (sharded) {
      allgather();
      for (int g = 0; g < world; g++) { const StreamParams P = ranks[g].P; emu_launch((C + 127) / 128, 128, [&] { k_st_merge(P, obs); }); }
    }
    if (ralg != 0) {
      for (int g = 0; g < world; g++) {
        const StreamParams P = ranks[g].P;
        const auto o = block_order((unsigned int)P.bpc * C);
        if (rfn == 2) {
          emu_launch2d((unsigned int)C, (unsigned int)P.mn_nt, MN_THREADS, [&] { k_st_mn_sums(P, obs); });
          emu_launch((unsigned int)C, MN_THREADS, [&] { k_st_mn_scan(P, obs); });
          emu_launch2d((unsigned int)C, (unsigned int)P.mn_nt, MN_THREADS, [&] { k_st_mn_positions(P, obs); });
        }
        emu_launch((unsigned int)P.bpc * C, THREADS, [&] { k_st_resample<Model, Real, PPT, THREADS>(P, obs); }, &o);
      }
    }
  }
  for (int g = 0; g < world; g++) { const StreamParams P = ranks[g].P; emu_launch(C, 256, [&] { k_st_flush<TS>(P, T); }); }
  if (sharded) {
    allgather();
    for (int g = 0; g < world; g++) { const StreamParams P = ranks[g].P; emu_launch((C + 127) / 128, 128, [&] { k_st_flush_merge(P, T); }); }
  }
  for (int g = 0; g < world; g++) {
    const Rank& R = ranks[g];
    for (int c = 0; c < C; c++) {
      printf("rank %d filter %d loglike %.17g n_resampled %d status %d early_exit %d\n", g, c, R.loglike[c], R.nres[c], R.status[c], R.early[c]);
      printf("ess");
      for (int t = 0; t <= T; t++) printf(" %.17g", R.ess[(size_t)c * (T + 1) + t]);
      printf("\nstate_est");
      for (int t = 0; t <= T; t++) printf(" %.17g", R.state_est[(size_t)c * (T + 1) + t]);
      printf("\nloglike_history");
      for (int t = 0; t < T; t++) printf(" %.17g", R.llh[(size_t)c * T + t]);
      printf("\n");
    }
  }
  return 0;
}

template <typename Model> static int by_shape(int argc, char** argv) {
  const int prec = atoi(argv[2]), threads = atoi(argv[3]);
  if (prec == 64) return threads == 128 ? run<Model, double, 4, 128>(argc, argv) : run<Model, double, 4, 256>(argc, argv);
  return threads == 128 ? run<Model, float, 8, 128>(argc, argv) : run<Model, float, 8, 256>(argc, argv);
}

int main(int argc, char** argv) {
  if (argc < 17) { fprintf(stderr, "usage: see the header of tests/host_stream.cpp\n"); return 2; }
  switch (atoi(argv[1])) {
    case 0: return by_shape<ModelArSin>(argc, argv);
    case 1: return by_shape<ModelLG>(argc, argv);
    case 2: return by_shape<ModelRwDrift>(argc, argv);
    case 4: return by_shape<ModelArCos>(argc, argv);
  }
  return 2;
}
