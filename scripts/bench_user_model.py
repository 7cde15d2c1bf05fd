"""Throughput of an NVRTC user model (stochastic volatility snippet of tests/test_nvrtc_gpu.py) on the general kernels and
on the streaming engine: python scripts/bench_user_model.py"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import engine_helpers as eh  # noqa: E402
from bayesssm_b200 import _native as nat  # noqa: E402
from test_nvrtc_gpu import SV_SNIPPET  # noqa: E402

ctx = nat.Context(0)
mid = C.c_int()
nat.check(ctx.lib.bssm_model_compile(ctx.handle, SV_SNIPPET.encode(), C.byref(mid)))
rng = np.random.default_rng(2)
x, ys = -1.0, []
for _ in range(200):
    x = -1.0 + 0.95 * (x + 1.0) + 0.25 * rng.standard_normal()
    ys.append(np.exp(x / 2) * rng.standard_normal())
y = np.array(ys)
for Cn, N in ((1, 1 << 20), (1, 1 << 24), (256, 1 << 16)):
    row = {"filters": Cn, "N": N, "T": len(y)}
    for name, eng in (("general", nat.ENGINE_GENERAL), ("stream", nat.ENGINE_STREAM)):
        best = None
        for s in range(3):
            r = eh.filter_run(ctx, mid.value, 0, 2, 0, N, y, [-1.0, 0.95, 0.25], seed=s, num_filters=Cn, precision=nat.F32, engine=eng)
            best = r["kernel_ms"] if best is None else min(best, r["kernel_ms"])
        row[name + "_G_pts_per_s"] = round(Cn * N * len(y) / (best * 1e-3) / 1e9, 2)
        row[name + "_loglike"] = float(r["loglike"][0])
    print(json.dumps(row), flush=True)
