"""Engine comparison on batched / single bootstrap filters (device-resident inputs, CUDA-event time of the whole
filter): python scripts/bench_engines.py C N T [model]  -> G particle-timesteps/s per engine.  Used to place the
AUTO thresholds of resolve_engine (csrc/bssm_engine.cu)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bayesssm_b200 import _native as nat  # noqa: E402
from bench import simulate_y  # noqa: E402

Cn, N, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
model = int(sys.argv[4]) if len(sys.argv) > 4 else 0
ctx = nat.Context(0)
y = torch.tensor(simulate_y(T), dtype=torch.float64, device="cuda")
theta = torch.tensor([[0.8, 1.0, 0.5]] * Cn, dtype=torch.float64, device="cuda")
ll = torch.zeros(Cn, dtype=torch.float64, device="cuda")
res = {}
for name, eng in (("persistent", nat.ENGINE_PERSISTENT), ("stream", nat.ENGINE_STREAM), ("auto", nat.ENGINE_AUTO)):
    cfg = nat.FilterConfig()
    cfg.model, cfg.algorithm, cfg.resample_algorithm, cfg.resample_fn = model, nat.BPF, nat.SISAR, nat.STRATIFIED
    cfg.threshold = -1.0
    cfg.num_particles, cfg.num_obs, cfg.dy = N, T, 1
    cfg.num_filters, cfg.precision = Cn, nat.F32
    cfg.seed, cfg.exact_resampling, cfg.engine = 1405, -1, eng
    best = None
    try:
        for i in range(4):
            cfg.run_id = i
            ms = C.c_float()
            nat.check(ctx.lib.bssm_filter_run_device(ctx.handle, C.byref(cfg), y.data_ptr(), theta.data_ptr(), ll.data_ptr(), C.byref(ms)))
            if i and (best is None or ms.value < best):
                best = ms.value
        res[name] = Cn * N * T / (best * 1e-3) / 1e9
    except nat.EngineError as e:
        res[name] = float("nan")
print(f"C={Cn} N={N} T={T} model={model}: " + "  ".join(f"{k} {v:.1f} G/s" for k, v in res.items()), flush=True)
