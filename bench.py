#!/usr/bin/env python
"""Benchmark of the particle-filter hot path (BASELINE.json metric: particle-timesteps/sec).

  python bench.py --gpus N --steps K --warmup W            our CUDA engine
  python bench.py --impl reference --gpus N ...            the reference algorithm on the host cores (oracle port)

A "step" is one full pass of the hot path over one batch: one bootstrap filter of the README
nonlinear-AR model, T=1000 observations, N=2^20 particles, SISAR (threshold 0.5 N) + stratified
resampling -- BASELINE.json configs[1].  With N GPUs every rank runs its own independent filter
(independent units, no data-path collective; weak scaling): that is `value`.

The same line carries three more blocks, measured in the same run (skip them with --no-extras):
  "pmmh"     BASELINE configs[4]: 1024 chains x N=65536 x T=1000, the chains SHARDED over the N ranks by global id, one
             final NCCL gather of the draws (strong scaling): iterations/s, particle-timesteps/s, roofline fraction
             from the run's own resampling count
  "sharded"  ONE filter of 2^28 particles, its particles sharded over the N ranks (one all-gather of a 64-byte record per
             observation, fused into the filter kernel through peer memory -- BSSM_SHARD_EXCHANGE=nccl: ncclAllGather; strong scaling), with an inline parity check: a small f64 filter run sharded and on one
             GPU must agree to 1e-9
  "f64"      configs[1] in the reference's own precision (fp64 throughout) on the persistent kernel
`--workload pmmh|sharded` times one of them alone as the line's `value`.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

THETA = (0.8, 1.0, 0.5)  # README.md:97-114
MODEL_AR = 0


def simulate_y(T, seed=1405):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal()
    ys = np.empty(T)
    for t in range(T):
        x = THETA[0] * x + np.sin(x) + THETA[1] * rng.standard_normal()
        ys[t] = x + THETA[2] * rng.standard_normal()
    return ys


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi SM clock / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.stop, self.th = index, [], threading.Event(), None

    def _run(self):
        while not self.stop.is_set():
            try:
                r = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5)
                f = [x.strip() for x in r.stdout.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.samples[0][1]),
                "reasons": reasons, "samples": len(self.samples)}


def algorithmic_bytes(N, T, n_resampled, sx=4, sw=4, sc=8, d=1):
    """B_P = 2 d s_x + s_w every step, B_R = s_w + 2 s_c + 2 d s_x on resampled steps."""
    return N * (T * (2 * d * sx + sw) + n_resampled * (sw + 2 * sc + 2 * d * sx))


def cpu_sample(N, T, seed=1405):
    import oracle
    y = simulate_y(T)
    secs, ll, nres = oracle.bench_bootstrap_filter(MODEL_AR, N, y, THETA, resample_algorithm=2, resample_fn=0,
                                                   threshold=0.5 * N, seed=seed)
    return secs, ll, nres


def run_reference(args, rank, world):
    """The reference algorithm (oracle port; R itself is not installable here) on all host cores: one
    independent filter per core, the package's only parallelism (one chain per worker, R/pmmh.R:512-535)."""
    if rank != 0:
        return
    import oracle
    oracle.lib()
    cores = os.cpu_count() or 1
    N, T = 1 << 17, 25
    from concurrent.futures import ThreadPoolExecutor

    def one(i):
        return cpu_sample(N, T, seed=1405 + i)[0]

    times = []
    with ThreadPoolExecutor(max_workers=cores) as ex:
        for s in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            list(ex.map(one, range(cores)))
            dt = time.perf_counter() - t0
            if s >= args.warmup:
                times.append(dt)
    total = sum(times)
    value = cores * N * T * args.steps / total
    sample = (f"{cores} independent bootstrap filters (one per core), N=2^17, T=25, SISAR+stratified, "
              "C oracle restating R/particle_filter_core.R + src/resampling.cpp with R's "
              "Mersenne-Twister/inversion RNG; R itself is not installable in this image")
    metric, unit, scaling, config = "particle-timesteps/sec", "particle-timesteps/s", "weak", workload_config(args)
    # what this arm actually runs: a bounded sample of that workload's algorithm, one filter per host core
    config["workload"] = ("bootstrap_filter nonlinear-AR (README model) SISAR threshold=0.5N %s resampling, CPU port: %d independent filters "
                          "(one per host core) of N=%d T=%d per step -- a bounded sample of the GPU arm's N=%d T=%d" % (args.resample_fn, cores, N, T, args.N, args.T))
    config["N"], config["T"], config["filters"], config["gpu_arm"] = N, T, cores, {"N": args.N, "T": args.T}
    config["precision"], config["engine"], config["l2"] = "f64", "oracle port (C)", "n/a (host)"
    if args.workload == "pmmh":
        # one PMMH iteration = one filter of N x T particle-timesteps per chain (R/pmmh.R:445-457); proposal and accept
        # are a few scalar operations beside it, so the port's iteration rate is its filter throughput over that work
        per_iter = float(args.chains) * args.pmmh_N * args.pmmh_T
        value = value / per_iter
        metric, unit, scaling = "pmmh-iterations/sec", "iter/s", "strong"
        config = {"workload": f"pmmh nonlinear-AR {args.chains} chains x N={args.pmmh_N} x T={args.pmmh_T}, chain-sharded, pilot skipped, "
                              "final NCCL gather of draws", "chains": args.chains, "N": args.pmmh_N, "T": args.pmmh_T, "engine": args.engine}
        sample += f"; converted to iterations of {args.chains} chains x N={args.pmmh_N} x T={args.pmmh_T} ({per_iter:.3g} particle-timesteps each)"
    elif args.workload == "sharded":
        scaling = "strong"
        config = {"workload": f"ONE bootstrap filter nonlinear-AR N={args.shard_N} T={args.shard_T} SISAR stratified, particle-sharded over the ranks",
                  "N": args.shard_N, "T": args.shard_T, "engine": "stream"}
    line = {"impl": "reference", "metric": metric, "value": value, "unit": unit,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {"workload": "bootstrap_filter nonlinear-AR (README model) T=%d N=%d SISAR threshold=0.5N %s resampling, "
                        "one filter per GPU" % (args.T, args.N, args.resample_fn),
            "T": args.T, "N": args.N, "resample_fn": args.resample_fn, "precision": args.precision,
            "engine": args.engine, "l2": "flushed between timed steps (256 MiB write)"}


def run_pmmh_workload(args, ctx, rank, local_rank, world, steps=None, warmup=None):
    """BASELINE configs[4]: PMMH on the nonlinear AR model, `--chains` chains x N = 65536 x T = 1000, chains split
    across the ranks by global id (strong scaling), pilot skipped (fixed proposal factor and particle count), one
    final NCCL gather of the draws.  A step is one PMMH iteration of ALL chains."""
    import torch
    import torch.distributed as dist

    from bayesssm_b200 import distributed as D
    from bayesssm_b200 import models, priors
    from bayesssm_b200 import _native as nat
    from bayesssm_b200.pmmh import default_tune_control, run_chains
    T, N, Ctot = args.pmmh_T, args.pmmh_N, args.chains
    steps = args.steps if steps is None else steps
    warmup = args.warmup if warmup is None else warmup
    y = simulate_y(T)
    base, count = D.shard_chains(Ctot, rank, world)
    m = models.nonlinear_ar()
    pri = [priors.uniform(0, 1), priors.exponential(1), priors.exponential(1)]
    init = np.tile(np.array(THETA), (count, 1))
    chol = np.tile(np.diag([0.05, 0.05, 0.05]), (count, 1, 1))
    prec = nat.F32 if args.precision == "f32" else nat.F64
    engine = {"auto": nat.ENGINE_AUTO, "general": nat.ENGINE_GENERAL, "persistent": nat.ENGINE_PERSISTENT, "stream": nat.ENGINE_STREAM}[args.engine]

    def run(iters, seed):
        return run_chains(ctx, m, nat.BPF, y, init, pri, [nat.TR_LOGIT, nat.TR_LOG, nat.TR_LOG], default_tune_control(),
                          iters + 1, seed, chain_id_base=base, fixed_num_particles=N, precision=prec, skip_pilot=True,
                          proposal_chol=chol, engine=engine)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.synchronize()

    run(max(warmup, 1), 1)
    barrier()
    l0 = ctx.launch_count()
    with ClockSampler(local_rank) as clocks:
        t0 = time.perf_counter()
        out = run(steps, 2)
        main_ms = out["main_ms"] * steps / (steps + 1)   # the first filter of the phase is not an iteration
        barrier()
        wall_ms = 1e3 * (time.perf_counter() - t0)
    launches = ctx.launch_count() - l0
    gathered = D.gather_chain_arrays({"theta_chain": out["theta_chain"], "n_accept": out["n_accept"]}, Ctot, rank, world,
                                     device=torch.device("cuda", local_rank) if world > 1 else None)
    r_frac = float(out["main_resampled_fraction"])
    if world > 1:
        t = torch.tensor([main_ms, wall_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        main_ms, wall_ms = float(t[0]), float(t[1])
        t = torch.tensor([r_frac], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        r_frac = float(t[0]) / world
    iters_per_s = steps / (main_ms * 1e-3)
    pts = Ctot * N * T * iters_per_s
    peak, peak_src = measured_peak_gbs()
    bpp = 12.0 + 28.0 * r_frac     # 12 B every step + 28 B on the steps that resampled (f32 state / weights, f64 cdf)
    return {"metric": "pmmh-iterations/sec", "value": iters_per_s, "unit": "iter/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": main_ms / steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": f"pmmh nonlinear-AR {Ctot} chains x N={N} x T={T}, chain-sharded, pilot skipped, final NCCL gather of draws",
                       "chains": Ctot, "N": N, "T": T, "engine": args.engine, "l2": "working set 2 x 256 MiB of particles per GPU at 1024 chains: larger than L2"},
            "particle_timesteps_per_s": pts,
            "e2e": {"value": (steps + 1) / (wall_ms * 1e-3), "unit": "iter/s",   # the call runs steps + 1 filter passes: draw 1 is the filter at the start value (R/pmmh.R:402-423)
                    "h2d_bytes_per_step": int(8 * T / steps),
                    "d2h_bytes_per_step": int(out["theta_chain"].nbytes / steps)},
            "gpu_launches": int(launches), "clocks": clocks.summary(),
            "roofline": {"bound": "hbm", "achieved": pts * bpp / 1e9, "peak": peak * world, "unit": "GB/s",
                         "frac": pts * bpp / 1e9 / (peak * world), "traffic": None, "peak_source": peak_src,
                         "kernel": "batched filter kernels of one iteration (AUTO: streaming engine k_st_step + k_st_resample at this size)",
                         "algorithmic_bytes_per_particle_timestep": bpp, "resampled_fraction": r_frac,
                         "note": "12 B + 28 B x the share of filter steps that resampled in THIS run (bssm_pmmh_result.main_resampled_fraction)"},
            "acceptance_rate": float(gathered["n_accept"].mean() / max(steps, 1)),
            "draws_gathered_shape": list(gathered["theta_chain"].shape)}


def run_sharded_workload(args, ctx, rank, local_rank, world, steps=None, warmup=None, parity=False):
    """One bootstrap filter of `--shard-N` particles (default 2^28) sharded over the ranks (third row):
    per observation ONE all-gather of a 64-byte record (config.exchange says how); no particle crosses NVLink.  Strong scaling: the
    filter is fixed, the ranks split its particles.  A step is one whole filter of `--shard-T` observations."""
    import torch
    import torch.distributed as dist

    from bayesssm_b200 import models, sharding as S
    N, T = args.shard_N, args.shard_T
    steps = args.steps if steps is None else steps
    warmup = args.warmup if warmup is None else warmup
    y = simulate_y(T)
    grp = S.ShardGroup(ctx, rank=rank, world=world, device=torch.device("cuda", local_rank) if world > 1 else None)
    m = models.nonlinear_ar()
    exch_text = {"peer": "fused into the filter kernel: stores into every rank's inbox through CUDA-IPC peer memory over NVLink",
                 "nccl": "ncclAllGather + merge kernel between the filter's two kernels" + (f" [{grp.exchange_note}]" if grp.exchange_note else ""),
                 "none": "one rank: no exchange"}[grp.exchange]

    def run(seed):
        return S.sharded_bootstrap_filter(y, N, m.init_fn, m.transition_fn, m.log_likelihood_fn, grp,
                                          resample_algorithm="SISAR", resample_fn=args.resample_fn, threshold=0.5 * N,
                                          precision=args.precision, seed=seed, capacity_factor=args.capacity_factor,
                                          phi=THETA[0], sigma_x=THETA[1], sigma_y=THETA[2])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.synchronize()

    for i in range(warmup):
        run(i)
    barrier()
    l0 = ctx.launch_count()
    ms, e2e = [], []
    with ClockSampler(local_rank) as clocks:
        for i in range(steps):
            barrier()
            t1 = time.perf_counter()
            r = run(100 + i)
            e2e.append(1e3 * (time.perf_counter() - t1))
            ms.append(r["kernel_ms"])
        barrier()
    launches = ctx.launch_count() - l0
    tot, tot_e2e = float(sum(ms)), float(sum(e2e))
    if world > 1:
        t = torch.tensor([tot, tot_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot, tot_e2e = float(t[0]), float(t[1])
    value = N * T * steps / (tot * 1e-3)
    peak, peak_src = measured_peak_gbs()
    nbytes = algorithmic_bytes(N, T, r["n_resampled"])
    achieved = nbytes / (tot / steps * 1e-3) / 1e9
    line = {"metric": "particle-timesteps/sec", "value": value, "unit": "particle-timesteps/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": tot / steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": f"ONE bootstrap filter nonlinear-AR N={N} T={T} SISAR threshold=0.5N {args.resample_fn}, particles "
                                   f"sharded over {world} GPU(s), one all-gather of a 64-byte record per observation ({exch_text})",
                       "N": N, "T": T, "engine": "stream (sharded)", "exchange": grp.exchange, "capacity_factor": args.capacity_factor,
                       "l2": "working set (8 B/particle x 2 buffers per rank) far larger than L2"},
            "e2e": {"value": N * T * steps / (tot_e2e * 1e-3), "unit": "particle-timesteps/s",
                    "h2d_bytes_per_step": 8 * T + 24, "d2h_bytes_per_step": 8 * (3 * T + 3) + 12},
            "gpu_launches": int(launches), "clocks": clocks.summary(),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak * world, "unit": "GB/s", "frac": achieved / (peak * world),
                         "traffic": None, "peak_source": peak_src, "kernel": "k_st_step + k_st_resample (all launches of one filter)",
                         "algorithmic_bytes_per_launch": nbytes, "resampled_steps": int(r["n_resampled"]), "T": T},
            "loglike": r["loglike"], "n_local_final": r["n_local_final"]}
    if parity:
        # inline parity check: a small filter in the parity precision, sharded over these ranks, against the SAME filter on one
        # GPU (this rank's streaming engine).  Philox streams are keyed by the global particle index, so the two differ only in
        # the summation order of the normaliser: 1e-9 relative
        from bayesssm_b200 import _native as nat
        from bayesssm_b200 import bootstrap_filter
        Np, Tp = 100003, 20
        yp = simulate_y(Tp, seed=7)
        rs = S.sharded_bootstrap_filter(yp, Np, m.init_fn, m.transition_fn, m.log_likelihood_fn, grp, resample_algorithm="SISAR",
                                        resample_fn=args.resample_fn, threshold=0.5 * Np, precision="f64", seed=11,
                                        capacity_factor=max(args.capacity_factor, 2.0), phi=THETA[0], sigma_x=THETA[1], sigma_y=THETA[2])
        r1 = bootstrap_filter(yp, Np, m.init_fn, m.transition_fn, m.log_likelihood_fn, resample_algorithm="SISAR",
                              resample_fn=args.resample_fn, threshold=0.5 * Np, return_particles=False, precision="f64", seed=11,
                              ctx=ctx, engine=nat.ENGINE_STREAM, phi=THETA[0], sigma_x=THETA[1], sigma_y=THETA[2])
        rel = abs(rs["loglike"] - r1["loglike"]) / abs(r1["loglike"])
        ok = bool(rel <= 1e-9 and rs["n_resampled"] == r1["n_resampled"])
        if world > 1:
            t = torch.tensor([0.0 if ok else 1.0], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ok = bool(t.item() == 0.0)
        line["parity_ok"] = ok
        line["parity"] = {"N": Np, "T": Tp, "precision": "f64", "loglike_sharded": rs["loglike"], "loglike_one_gpu": r1["loglike"],
                          "rel_diff": rel, "tolerance": 1e-9, "n_resampled": [int(rs["n_resampled"]), int(r1["n_resampled"])]}
    grp.close()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--N", type=int, default=1 << 20)
    ap.add_argument("--T", type=int, default=1000)
    ap.add_argument("--resample-fn", dest="resample_fn", default="stratified", choices=["stratified", "systematic", "multinomial"])
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--engine", default="auto", choices=["auto", "general", "persistent", "stream"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="filter", choices=["filter", "pmmh", "sharded"],
                    help="filter: BASELINE configs[1] (default); pmmh: configs[4], 1024 chains x N=65536 x T=1000 chain-sharded; "
                         "sharded: one filter of --shard-N particles sharded over the ranks")
    ap.add_argument("--shard-N", dest="shard_N", type=int, default=1 << 28)
    ap.add_argument("--shard-T", dest="shard_T", type=int, default=50)
    ap.add_argument("--capacity-factor", dest="capacity_factor", type=float, default=1.5)
    ap.add_argument("--chains", type=int, default=1024)
    ap.add_argument("--pmmh-N", dest="pmmh_N", type=int, default=65536)
    ap.add_argument("--pmmh-T", dest="pmmh_T", type=int, default=1000)
    ap.add_argument("--no-extras", action="store_true", help="default workload only: skip the pmmh / sharded / f64 blocks of the line")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from bayesssm_b200 import _native as nat
    from bayesssm_b200 import bootstrap_filter, models

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = nat.Context(local_rank)
    lib = ctx.lib
    if args.workload in ("pmmh", "sharded"):
        line = (run_pmmh_workload(args, ctx, rank, local_rank, world) if args.workload == "pmmh"
                else run_sharded_workload(args, ctx, rank, local_rank, world, parity=True))
        if rank == 0:
            print(json.dumps(line), flush=True)
        ctx.close()
        if world > 1:
            dist.destroy_process_group()
        return
    N, T = args.N, args.T
    y = simulate_y(T)
    fn = nat.RESAMPLE_FNS[args.resample_fn]
    prec = nat.F32 if args.precision == "f32" else nat.F64
    engine = {"auto": nat.ENGINE_AUTO, "general": nat.ENGINE_GENERAL, "persistent": nat.ENGINE_PERSISTENT, "stream": nat.ENGINE_STREAM}[args.engine]

    # device-resident inputs (torch owns the memory; the engine gets raw pointers)
    d_y = torch.tensor(y, dtype=torch.float64, device="cuda")
    d_theta = torch.tensor([THETA], dtype=torch.float64, device="cuda")
    d_ll = torch.zeros(1, dtype=torch.float64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    cfg = nat.FilterConfig()
    cfg.model, cfg.algorithm, cfg.resample_algorithm, cfg.resample_fn = MODEL_AR, nat.BPF, nat.SISAR, fn
    cfg.threshold = 0.5 * N
    cfg.num_particles, cfg.num_obs, cfg.dy = N, T, 1
    cfg.num_filters, cfg.precision = 1, prec
    cfg.seed, cfg.run_id, cfg.stream_base = 1405, 0, rank
    cfg.return_particles, cfg.exact_resampling, cfg.engine = 0, -1, engine

    def device_step(i):
        cfg.run_id = i
        ms = C.c_float()
        nat.check(lib.bssm_filter_run_device(ctx.handle, C.byref(cfg), d_y.data_ptr(), d_theta.data_ptr(),
                                             d_ll.data_ptr(), C.byref(ms)))
        return ms.value

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.synchronize()

    for i in range(args.warmup):
        device_step(i)
    barrier()
    launches0 = ctx.launch_count()
    step_ms = []
    with ClockSampler(local_rank) as clocks:
        for i in range(args.steps):
            flush.fill_(i & 0xFF)  # L2 flush between timed iterations (outside the timed region)
            torch.cuda.synchronize()
            step_ms.append(device_step(args.warmup + i))  # CUDA events on the engine's stream around the whole filter
        barrier()
    launches = ctx.launch_count() - launches0
    total_ms = float(sum(step_ms))
    per_rank_ms = [total_ms / args.steps]
    if world > 1:
        # every rank's own time as well (each runs its own filter, no data-path collective): the spread between the GPUs of a box
        # is what the max over ranks -- the line's ms_per_step -- loses against one GPU
        g = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(g, torch.tensor([total_ms / args.steps], dtype=torch.float64, device="cuda"))
        per_rank_ms = [float(v.item()) for v in g]
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = world * N * T * args.steps / (total_ms * 1e-3)

    # number of resampled steps of the timed configuration (for the algorithmic bytes): one host-API run
    m = models.nonlinear_ar()
    api = lambda seed: bootstrap_filter(y, N, m.init_fn, m.transition_fn, m.log_likelihood_fn,
                                        resample_algorithm="SISAR", resample_fn=args.resample_fn, threshold=0.5 * N,
                                        return_particles=False, precision=args.precision, seed=seed, ctx=ctx, engine=engine,
                                        phi=THETA[0], sigma_x=THETA[1], sigma_y=THETA[2])
    r0 = api(1405)
    n_res = r0["n_resampled"]

    # e2e: the user-facing call with HOST buffers (y up, state_est / ess / loglike_history / loglike back)
    barrier()
    t0 = time.perf_counter()
    e2e_ms = []
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        api(2000 + i)
        e2e_ms.append(1e3 * (time.perf_counter() - t1))
    e2e_total = float(sum(e2e_ms))
    if world > 1:
        t = torch.tensor([e2e_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_total = float(t.item())
    e2e_value = world * N * T * args.steps / (e2e_total * 1e-3)
    h2d = 8 * T + 8 * 3
    d2h = 8 * ((T + 1) * 2 + T + 1) + 4 * 3

    peak, peak_src = measured_peak_gbs()
    traffic = None   # DRAM bytes per launch of the persistent kernel from the committed ncu capture (same configuration only)
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "persistent_traffic.json")))
        if (N, T, args.resample_fn, args.precision) == (1 << 20, 1000, "stratified", "f32") and launches / max(args.steps, 1) <= 4:
            traffic = tj["dram_bytes_per_launch"]
    except Exception:
        pass
    bytes_per_launch = algorithmic_bytes(N, T, n_res, sx=4 if prec == nat.F32 else 8, sw=4 if prec == nat.F32 else 8)
    ms_per_launch = total_ms / args.steps
    achieved = bytes_per_launch / (ms_per_launch * 1e-3) / 1e9
    line = {
        "metric": "particle-timesteps/sec", "value": value, "unit": "particle-timesteps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": workload_config(args),
        "e2e": {"value": e2e_value, "unit": "particle-timesteps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "per_rank_ms": per_rank_ms,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src,
                     "kernel": "persistent filter kernel" if launches / max(args.steps, 1) <= 4 else
                               ("k_st_step + k_st_resample (streaming engine, all launches of one filter)" if launches / max(args.steps, 1) <= 2 * T + 8
                                else "whole filter pass (general engine, all launches of one step)"),
                     "algorithmic_bytes_per_launch": bytes_per_launch, "resampled_steps": int(n_res), "T": T},
        "loglike": r0["loglike"],
    }
    if rank == 0 and not args.no_cpu_baseline:
        Ns, Ts = 1 << 20, 200   # ~25 s of one host core: a bounded sample of the same workload (resampling steps included)
        secs, _, _ = cpu_sample(Ns, Ts)
        line["cpu_baseline"] = {"value": Ns * Ts / secs, "unit": "particle-timesteps/s", "cores": 1, "kind": "port",
                                "sample": f"same model/config, N=2^20, first {Ts} of the {T} observations, 1 thread (the R "
                                          "interpreter is single-threaded), C oracle with R's RNG cost model"}
    if not args.no_extras:
        # the north star's other numbers, in the same run and on the same line (the driver records this line for every N)
        def keep(d, keys):
            return {k: d[k] for k in keys if k in d}
        px = run_pmmh_workload(args, ctx, rank, local_rank, world, steps=max(3, min(args.steps, 5)), warmup=1)
        line["pmmh"] = keep(px, ["metric", "value", "unit", "n_gpus", "steps", "ms_per_step", "scaling", "dtype", "config",
                                 "particle_timesteps_per_s", "e2e", "gpu_launches", "roofline", "acceptance_rate", "draws_gathered_shape"])
        sx = run_sharded_workload(args, ctx, rank, local_rank, world, steps=max(2, min(args.steps, 4)), warmup=1, parity=True)
        line["sharded"] = keep(sx, ["metric", "value", "unit", "n_gpus", "steps", "ms_per_step", "scaling", "dtype", "config", "e2e",
                                    "gpu_launches", "roofline", "loglike", "n_local_final", "parity_ok", "parity"])
        # configs[1] in the reference's precision: fp64 state, weights, cdf -- 64 B per resampled particle-timestep, 24 B otherwise
        cfg.precision, cfg.engine = nat.F64, nat.ENGINE_PERSISTENT
        d_ll64 = torch.zeros(1, dtype=torch.float64, device="cuda")

        def f64_step(i):
            cfg.run_id = i
            ms = C.c_float()
            nat.check(lib.bssm_filter_run_device(ctx.handle, C.byref(cfg), d_y.data_ptr(), d_theta.data_ptr(), d_ll64.data_ptr(), C.byref(ms)))
            return ms.value
        f64_step(0)
        barrier()
        n64 = max(2, min(args.steps, 3))
        ms64 = []
        for i in range(n64):
            flush.fill_(i & 0xFF)
            torch.cuda.synchronize()
            ms64.append(f64_step(10 + i))
        barrier()
        tot64 = float(sum(ms64))
        if world > 1:
            t = torch.tensor([tot64], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tot64 = float(t.item())
        r64 = bootstrap_filter(y, N, m.init_fn, m.transition_fn, m.log_likelihood_fn, resample_algorithm="SISAR", resample_fn=args.resample_fn,
                               threshold=0.5 * N, return_particles=False, precision="f64", seed=1405, ctx=ctx, engine=nat.ENGINE_PERSISTENT,
                               phi=THETA[0], sigma_x=THETA[1], sigma_y=THETA[2])
        b64 = algorithmic_bytes(N, T, r64["n_resampled"], sx=8, sw=8)
        a64 = b64 / (tot64 / n64 * 1e-3) / 1e9
        line["f64"] = {"metric": "particle-timesteps/sec", "value": world * N * T * n64 / (tot64 * 1e-3), "unit": "particle-timesteps/s",
                       "n_gpus": world, "steps": n64, "ms_per_step": tot64 / n64, "scaling": "weak", "dtype": "f64",
                       "config": {"workload": "the default workload in the reference's precision (fp64 state, weights, cdf, libm sin / exp / log), "
                                              "persistent kernel k_fast_bpf<double>", "T": T, "N": N, "engine": "persistent"},
                       "roofline": {"bound": "hbm", "achieved": a64, "peak": peak, "unit": "GB/s", "frac": a64 / peak,
                                    "algorithmic_bytes_per_launch": b64, "resampled_steps": int(r64["n_resampled"]),
                                    "note": "24 B per particle-timestep + 40 B on resampled steps (all-fp64 model of SURVEY 8d)"},
                       "loglike": r64["loglike"], "loglike_f32": r0["loglike"]}
    if world > 1:
        dist.barrier()
    if rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
