#!/usr/bin/env python
"""Randomised CPU campaign: the streaming and persistent kernels' text on the SIMT emulation (tests/host_stream.cpp,
tests/host_fast.cpp, and in the parity precision tests/host_general.cpp, over tests/simt_emu.h) against the oracle's
Philox-mode filter, over random sizes, geometries,
resamplers, thresholds, models, seeds and (streaming) emulated shard counts / block orders.  Parity precision (default):
any difference above 1e-8 is a logic bug.  --f32: the throughput-precision instantiations; scratch memory starts as NaN
bit patterns, so a slot that is never written poisons the sums -- the check is status 0, finite outputs and, from 2000
particles up, closeness to the oracle at Monte-Carlo scale.
usage: python scripts/fuzz_emulated_kernels.py [--cases 200] [--seed 1] [--max-n 30000] [--f32]"""
import argparse
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from test_filter_gpu import THETA, sim_y  # noqa: E402

CXX = ["g++", "-O1", "-std=c++20", "-ffp-contract=off", "-Wno-unknown-pragmas", "-pthread"]


def build(tmp, name):
    exe = os.path.join(tmp, name)
    subprocess.run(CXX + ["-o", exe, os.path.join(ROOT, "tests", name + ".cpp")], check=True)
    return exe


def call(exe, args, y, thetas, env=None):
    th = np.zeros((len(thetas), 3))
    for c, t in enumerate(thetas):
        th[c, :len(t)] = t
    r = subprocess.run([exe] + [str(a) for a in args], input=np.ascontiguousarray(y, dtype=np.float64).tobytes() + th.tobytes(),
                       capture_output=True, timeout=900, env=env)
    if r.returncode != 0:
        return None, r.stderr.decode()[-500:]
    lines, recs = r.stdout.decode().strip().splitlines(), []
    for i in range(0, len(lines), 4):
        h = lines[i].split()
        recs.append({"rank": int(h[1]), "filter": int(h[3]), "loglike": float(h[5]), "n_resampled": int(h[7]), "status": int(h[9]),
                     "early_exit": int(h[11]), "ess": np.array(lines[i + 1].split()[1:], float),
                     "state_est": np.array(lines[i + 2].split()[1:], float)})
    return recs, ""


def differs_f32(rec, ref, n):
    if rec["status"] != 0:
        return "status"
    if ref["early_exit"] or rec["early_exit"]:
        return "" if rec["early_exit"] == ref["early_exit"] else "early_exit"
    if not (np.isfinite(rec["loglike"]) and np.isfinite(rec["ess"]).all() and np.isfinite(rec["state_est"]).all()):
        return "non-finite"
    if n >= 2000 and (abs(rec["loglike"] - ref["loglike"]) > 0.5 or np.abs(rec["state_est"] - ref["state_est"][:, 0]).max() > 0.5):
        return "far from the oracle"
    return ""


def differs(rec, ref, tol=1e-8):
    if rec["status"] != 0 or rec["early_exit"] != ref["early_exit"] or rec["n_resampled"] != ref["n_resampled"]:
        return "flags"
    if ref["early_exit"]:
        return ""
    if abs(rec["loglike"] - ref["loglike"]) > tol * max(1.0, abs(ref["loglike"])):
        return "loglike"
    if not np.allclose(rec["ess"], ref["ess"], rtol=tol) or not np.allclose(rec["state_est"], ref["state_est"][:, 0], rtol=tol, atol=tol):
        return "ess/state_est"
    return ""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=200)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--max-n", dest="max_n", type=int, default=30000)
    ap.add_argument("--f32", action="store_true")
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    bad = 0
    with tempfile.TemporaryDirectory() as tmp:
        hs, hf, hg = build(tmp, "host_stream"), build(tmp, "host_fast"), build(tmp, "host_general")
        for case in range(args.cases):
            model = int(rng.choice([0, 1, 2, 4]))
            base = np.array(THETA[0] if model == 4 else THETA[model])
            if rng.random() < 0.25:
                base = base * np.array([1.0, 1.0, rng.choice([0.01, 0.1, 3.0])])[:len(base)]   # sharp / flat likelihoods
            N = int(rng.choice([rng.integers(1, 40), rng.integers(40, 3000), rng.integers(3000, args.max_n)]))
            T = int(rng.integers(0, 9))
            C = int(rng.choice([1, 1, 2, 3]))
            rfn, ralg = int(rng.integers(0, 2)), int(rng.integers(0, 3))
            thr = float(rng.choice([-1.0, -1.0, rng.uniform(0.1, 1.0) * N]))
            seed, run_id, sb = int(rng.integers(0, 2**31)), int(rng.integers(0, 100)), int(rng.integers(0, 100))
            y = sim_y(0 if model == 4 else model, T, rng) if T else np.zeros(0)
            thetas = [list(base * (1 + 0.03 * c)) for c in range(C)]
            ot = None
            if T and rng.random() < 0.3:                               # observation times with gaps (extra transitions)
                ot = np.cumsum(rng.integers(1, 4, size=T)).tolist()
            env = dict(os.environ, EMU_OBS_TIMES=",".join(map(str, ot))) if ot else None
            ragged = C > 1 and thr < 0 and rng.random() < 0.5          # per-filter particle counts (FilterDev::n_per)
            ns = [int(rng.integers(1, N + 1)) for _ in range(C)] if ragged else [N] * C
            if ragged:
                ns[int(rng.integers(0, C))] = N
            tail = ns if ragged else []
            refs = [oracle.particle_filter(model, 0, ralg, rfn, ns[c], y, thetas[c], threshold=thr, obs_times=ot, seed=seed, run_id=run_id, stream=sb + c)
                    for c in range(C)]
            # streaming engine
            threads = int(rng.choice([128, 256]))
            world = int(rng.choice([1, 1, 2, 3, 4])) if C == 1 and N >= 64 else 1
            bpc, order = int(rng.integers(1, 7)), int(rng.integers(0, 3))
            sargs = [model, 32 if args.f32 else 64, threads, N, T, C, bpc, rfn, ralg, thr, seed, run_id, sb, world, 4.0, order] + tail
            recs, err = call(hs, sargs, y, thetas, env)
            cmp = (lambda r, ref: differs_f32(r, ref, ns[r["filter"]])) if args.f32 else differs
            what = err or next((d for r in recs if (d := cmp(r, refs[r["filter"]]))), "")
            if what:
                bad += 1
                print("STREAM MISMATCH", what, sargs, thetas, flush=True)
            # the same configuration on the chain-persistent kernel (cooperative launches over groups of filters, concurrent blocks)
            if world == 1 and T > 0:
                cenv = dict(env or os.environ, EMU_CHAIN=str(int(rng.integers(1, C + 1))))
                recs, err = call(hs, sargs, y, thetas, cenv)
                what = err or next((d for r in recs if (d := cmp(r, refs[r["filter"]]))), "")
                if what:
                    bad += 1
                    print("CHAIN KERNEL MISMATCH", what, sargs, cenv["EMU_CHAIN"], thetas, flush=True)
            # multinomial by sorted uniforms on the streaming engine (oracle: resample_fn 3 restates the same variant)
            if world == 1 and ralg != 0 and rng.random() < 0.5:
                mrefs = [oracle.particle_filter(model, 0, ralg, 3, ns[c], y, thetas[c], threshold=thr, obs_times=ot, seed=seed, run_id=run_id, stream=sb + c)
                         for c in range(C)]
                margs = list(sargs); margs[7] = 2
                recs, err = call(hs, margs, y, thetas, env)
                what = err or next((d for r in recs if (d := cmp(r, mrefs[r["filter"]]))), "")
                if what:
                    bad += 1
                    print("MULTINOMIAL MISMATCH", what, margs, thetas, flush=True)
            # persistent kernel
            variant = int(rng.integers(0, 2)) + (2 if args.f32 else 0)
            G = int(rng.integers(1, 6))
            if (N + G - 1) // G > 7168:
                G = (N + 7167) // 7168
            fargs = [model, variant, G, int(rng.integers(1, C + 1)), N, T, C, rfn, ralg, thr, seed, run_id, sb] + tail
            recs, err = call(hf, fargs, y, thetas, env)
            what = err or next((d for r in recs if (d := cmp(r, refs[r["filter"]]))), "")
            if what:
                bad += 1
                print("PERSISTENT MISMATCH", what, fargs, thetas, flush=True)
            # general engine (parity precision only): any algorithm, any built-in model, any resampler, smaller sizes
            if not args.f32 and not ragged:
                gmodel = int(rng.choice([0, 1, 2, 3, 4, 5]))
                galg = int(rng.choice([0, 1, 2])) if gmodel != 5 else 0
                grfn = int(rng.integers(0, 3))
                gN = min(N, 4000)
                nth = {0: 3, 1: 3, 2: 2, 3: 2, 4: 3, 5: 1}[gmodel]
                gthetas = [list(np.array(THETA[gmodel][:nth]) * (1 + 0.03 * c)) + list(THETA[gmodel][nth:]) for c in range(C)]
                gy = sim_y(gmodel, T, rng) if T else np.zeros(0)
                gthr = thr if thr < 0 else min(thr, 0.9 * gN)
                gargs = [gmodel, galg, gN, T, C, grfn, ralg, gthr, seed, run_id, sb, 1]
                r = subprocess.run([hg] + [str(a) for a in gargs], input=np.ascontiguousarray(gy, dtype=np.float64).tobytes()
                                   + np.ascontiguousarray(gthetas, dtype=np.float64).tobytes(), capture_output=True, timeout=900, env=env)
                what = ""
                if r.returncode != 0:
                    what = r.stderr.decode()[-300:]
                else:
                    lines = r.stdout.decode().strip().splitlines()
                    for i in range(0, len(lines), 4):
                        h = lines[i].split()
                        c = int(h[3])
                        rec = {"status": int(h[9]), "early_exit": int(h[11]), "n_resampled": int(h[7]), "loglike": float(h[5]),
                               "ess": np.array(lines[i + 1].split()[1:], float), "state_est": np.array(lines[i + 2].split()[1:], float)}
                        ref = oracle.particle_filter(gmodel, galg, ralg, grfn, gN, gy, gthetas[c] if gmodel == 3 else gthetas[c][:nth],
                                                     threshold=gthr, obs_times=ot, seed=seed, run_id=run_id, stream=sb + c)
                        ref = dict(ref, state_est=ref["state_est"].reshape(-1, 1))
                        what = what or differs(rec, ref)
                if what:
                    bad += 1
                    print("GENERAL MISMATCH", what, gargs, gthetas, flush=True)
            if (case + 1) % 20 == 0:
                print(f"{case + 1} cases, {bad} mismatches", flush=True)
    print(f"done: {args.cases} cases, {bad} mismatches")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
