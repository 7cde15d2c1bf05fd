"""Particle-sharded single filter: one bootstrap filter whose particles are block-partitioned over the GPUs
of one box, one process per GPU (third row; C ABI: bssm_shard_* / bssm_filter_run_sharded).

Per observation the ranks exchange one 64-byte record (local max log-weight, sum e, sum e^2, sum e*x,
pending state sum) -- through peer memory from inside the propagate / weight kernel when the ranks could
map each other's inboxes with CUDA IPC (`ShardGroup.exchange == "peer"`, the default on one box), else with
ncclAllGather between the kernels; every rank derives the same global normaliser / ESS / decision and
its own cdf offset (exclusive prefix of the per-rank weight totals), and resamples the offspring of its
own particles -- the output slots [F(A_g / S), F(A_{g+1} / S)).  No particle crosses NVLink.

`ShardGroup` bootstraps the NCCL communicator of the engine through whatever torch.distributed process
group the host already has (the 128-byte unique id is broadcast from rank 0).  `exchange_step` is the
host-side statement of one exchange in numpy (same algebra as csrc/bssm_stream.cuh:st_global); the
world_size-2 gloo test drives it on CPU and checks the ancestors against the unsharded resampler."""
from __future__ import annotations

import ctypes as C
import glob
import os

import numpy as np

from . import _native as nat


def nccl_library_path():
    """The NCCL the process already uses (torch's bundled libnccl.so.2), else None (loader search path)."""
    env = os.environ.get("BSSM_NCCL_LIB")
    if env:
        return env
    try:
        import nvidia.nccl  # type: ignore
        for base in list(getattr(nvidia.nccl, "__path__", [])):
            hits = sorted(glob.glob(os.path.join(base, "lib", "libnccl.so*")))
            if hits:
                return hits[0]
    except Exception:
        pass
    return None


def partition(n: int, world: int, rank: int):
    """Initial block partition (boundaries at multiples of 4 = Philox quads): (goff, nloc).  Host-only."""
    lib = nat.load_library()
    g, nl = C.c_int64(), C.c_int()
    nat.check(lib.bssm_shard_partition(int(n), int(world), int(rank), C.byref(g), C.byref(nl)))
    return int(g.value), int(nl.value)


class ShardGroup:
    """The engine's NCCL communicator over the ranks of torch.distributed's default group."""

    def __init__(self, ctx: nat.Context, rank: int | None = None, world: int | None = None, device=None,
                 exchange: str | None = None):
        import torch
        import torch.distributed as dist
        self.ctx = ctx
        if world is None:
            world = dist.get_world_size() if dist.is_initialized() else 1
        if rank is None:
            rank = dist.get_rank() if dist.is_initialized() else 0
        self.rank, self.world = int(rank), int(world)
        path = nccl_library_path()
        cpath = path.encode() if path else None
        uid = (C.c_ubyte * 128)()
        if self.world > 1:
            if self.rank == 0:
                nat.check(ctx.lib.bssm_shard_unique_id(cpath, uid))
            t = torch.tensor(list(bytes(uid)), dtype=torch.uint8)
            if device is not None:
                t = t.to(device)
            dist.broadcast(t, src=0)
            raw = bytes(t.cpu().tolist())
            uid = (C.c_ubyte * 128).from_buffer_copy(raw)
        nat.check(ctx.lib.bssm_shard_init(ctx.handle, cpath, self.rank, self.world, uid))
        self.exchange = "nccl" if self.world > 1 else "none"
        self.exchange_note = ""
        if exchange is None:
            exchange = os.environ.get("BSSM_SHARD_EXCHANGE", "peer")
        if exchange not in ("peer", "nccl"):
            raise ValueError("exchange must be 'peer' or 'nccl'")
        if self.world > 1 and exchange == "peer":
            self._attach_peers(device)

    def _attach_peers(self, device):
        """The per-observation exchange through peer memory (CUDA IPC over NVLink, fused into the filter kernel) instead
        of ncclAllGather.  All ranks attach or none does: the outcome of every stage is agreed with an all_reduce(MIN)."""
        import torch
        import torch.distributed as dist
        lib, h = self.ctx.lib, self.ctx.handle

        def agreed(ok: bool) -> bool:
            t = torch.tensor([1 if ok else 0], dtype=torch.int32)
            if device is not None:
                t = t.to(device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            return bool(int(t.cpu()[0]))

        mine = (C.c_ubyte * 64)()
        st = lib.bssm_shard_peer_export(h, mine)
        note = "" if st == nat.OK else nat.last_error()
        if not agreed(st == nat.OK):
            lib.bssm_shard_peer_detach(h)
            self.exchange_note = "peer export failed on a rank: " + (note or "another rank")
            return
        t = torch.tensor(list(bytes(mine)), dtype=torch.uint8)
        if device is not None:
            t = t.to(device)
        parts = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(parts, t)
        raw = b"".join(bytes(q.cpu().tolist()) for q in parts)
        buf = (C.c_ubyte * (64 * self.world)).from_buffer_copy(raw)
        st = lib.bssm_shard_peer_attach(h, buf)
        note = "" if st == nat.OK else nat.last_error()
        if not agreed(st == nat.OK):
            lib.bssm_shard_peer_detach(h)
            self.exchange_note = "peer attach failed on a rank: " + (note or "another rank")
            return
        self.exchange = "peer"

    def close(self):
        if self.ctx is not None and self.ctx.handle:
            if self.exchange == "peer":
                import torch.distributed as dist
                self.ctx.lib.bssm_shard_peer_detach(self.ctx.handle)   # waits for this rank's stream
                if dist.is_initialized():
                    dist.barrier()                                        # nobody frees an inbox a peer still has mapped
            self.ctx.lib.bssm_shard_finalize(self.ctx.handle)
        self.ctx = None


def sharded_bootstrap_filter(y, num_particles, init_fn, transition_fn, log_likelihood_fn, group: ShardGroup,
                             obs_times=None, resample_algorithm="SISAR", resample_fn="stratified", threshold=None,
                             *, precision="f32", seed=0, capacity_factor=1.5, run_id=0, stream=0, **params):
    """bootstrap_filter (R/bootstrap_filter.R:129-171) for ONE filter of `num_particles` (global count) sharded
    over `group`.  Collective: every rank calls it with the same arguments and gets the same result."""
    from .filters import _match_arg, _precision, _theta_vector
    from .models import resolve_model
    resample_algorithm = _match_arg(resample_algorithm, ("SISAR", "SISR", "SIS"), "resample_algorithm")
    resample_fn = _match_arg(resample_fn, ("stratified", "systematic"), "resample_fn")
    model = resolve_model(init_fn, transition_fn, log_likelihood_fn)
    y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(len(y), -1))
    T, dy = y.shape
    theta = np.ascontiguousarray(_theta_vector(model, params)[None, :])
    ctx = group.ctx
    cfg = nat.FilterConfig()
    cfg.model, cfg.algorithm = model.model_id, nat.BPF
    cfg.resample_algorithm, cfg.resample_fn = nat.RESAMPLE_ALGORITHMS[resample_algorithm], nat.RESAMPLE_FNS[resample_fn]
    cfg.threshold = -1.0 if threshold is None else float(threshold)
    cfg.num_particles, cfg.num_obs, cfg.dy = int(num_particles), T, dy
    ot = None
    if obs_times is not None:
        ot = np.ascontiguousarray(obs_times, dtype=np.int32)
        cfg.obs_times = ot.ctypes.data_as(nat.c_int_p)
    cfg.num_filters, cfg.precision = 1, _precision(precision)
    cfg.seed, cfg.run_id, cfg.stream_base = int(seed), int(run_id), int(stream)
    cfg.exact_resampling, cfg.engine = 0, nat.ENGINE_STREAM
    d = model.dim
    out = {"loglike": np.zeros(1), "loglike_history": np.zeros((1, T)), "ess": np.zeros((1, T + 1)),
           "state_est": np.zeros((1, T + 1, d)), "status": np.zeros(1, dtype=np.int32),
           "early_exit": np.zeros(1, dtype=np.int32), "n_resampled": np.zeros(1, dtype=np.int32)}
    res = nat.FilterResult()
    for k in ("loglike", "loglike_history", "ess", "state_est"):
        setattr(res, k, out[k].ctypes.data_as(nat.c_double_p))
    for k in ("status", "early_exit", "n_resampled"):
        setattr(res, k, out[k].ctypes.data_as(nat.c_int32_p))
    n_local = C.c_int()
    nat.check(ctx.lib.bssm_filter_run_sharded(ctx.handle, C.byref(cfg), y.ctypes.data_as(nat.c_double_p),
                                              theta.ctypes.data_as(nat.c_double_p), float(capacity_factor), C.byref(res),
                                              C.byref(n_local)))
    if out["status"][0] == nat.ERR_CAPACITY:
        raise nat.EngineError(nat.ERR_CAPACITY, "a rank's share of the offspring outgrew its storage; raise capacity_factor")
    if out["status"][0] == nat.ERR_NCCL:
        raise nat.EngineError(nat.ERR_NCCL, "a rank's record never arrived (peer-memory exchange timed out): the group is out of step, re-create it")
    if out["status"][0] == nat.ERR_NAN_WEIGHT:
        raise ValueError("missing value where TRUE/FALSE needed")
    se = out["state_est"][0]
    r = {"state_est": se[:, 0].copy() if d == 1 else se.copy(), "ess": out["ess"][0].copy(),
         "loglike": float(out["loglike"][0]), "loglike_history": out["loglike_history"][0].copy(), "algorithm": "BPF",
         "n_resampled": int(out["n_resampled"][0]), "kernel_ms": float(res.kernel_ms), "n_local_final": int(n_local.value)}
    if not out["early_exit"][0]:
        r["resample_algorithm"] = resample_algorithm
    return r


# ---- host-side statement of one exchange (numpy; the algebra of st_global in csrc/bssm_stream.cuh) ----
def count_le(c: float, n: int, u) -> int:
    """F(c) = #{ i : (i + u_i) / n <= c } in closed form: with t = c n and i = floor(t) only slot i needs a look.
    u: array of n stratified uniforms, or a scalar (systematic)."""
    t = c * n
    if not (t > 0.0):
        return 0
    if t >= n:
        return n
    i = int(t)
    ui = float(u) if np.ndim(u) == 0 else float(u[i])
    return i + (1 if (i + ui) <= t else 0)


def local_record(lw_local: np.ndarray):
    """(local max, sum exp(lw - max)) of a rank's log-weights."""
    if lw_local.size == 0:
        return -np.inf, 0.0
    m = float(np.max(lw_local))
    if m == -np.inf:
        return m, 0.0
    return m, float(np.sum(np.exp(lw_local - m)))


def exchange_step(records, rank: int, n: int, u):
    """From the all-gathered records [(m_g, s_g)] in rank order: global max M, normaliser S, this rank's cdf
    numerator interval [abase, aend) and scale exp(m_rank - M), and the output slots [o_lo, o_hi) it serves."""
    world = len(records)
    M = max(m for m, _ in records)
    S, abase, aend, gscale = 0.0, 0.0, 0.0, 0.0
    bounds = [0.0]
    for g, (m, s) in enumerate(records):
        sc = 0.0 if (m == -np.inf or M == -np.inf) else float(np.exp(m - M))
        if g == rank:
            abase, gscale = S, sc
        S = S + s * sc
        if g == rank:
            aend = S
        bounds.append(S)
    o = [count_le(b / S, n, u) for b in bounds]
    o[0], o[-1] = 0, n
    return {"M": M, "S": S, "abase": abase, "aend": aend, "gscale": gscale, "o_lo": o[rank], "o_hi": o[rank + 1],
            "all_slots": o, "loglike_increment": M + np.log(S) - np.log(n), "world": world}


def local_ancestors(lw_local: np.ndarray, goff: int, ex: dict, n: int, u):
    """Global ancestor index of every output slot this rank serves: first j with cdf[j] >= pos, clamp (the tie
    rule of src/resampling.cpp:32-37), on the rank's own stretch of the cdf."""
    m, _ = local_record(lw_local)
    e = np.exp(lw_local - m) if m > -np.inf else np.zeros_like(lw_local)
    cdf = (ex["abase"] + np.cumsum(e) * ex["gscale"]) / ex["S"]
    slots = np.arange(ex["o_lo"], ex["o_hi"])
    uu = np.full(slots.shape, float(u)) if np.ndim(u) == 0 else np.asarray(u)[slots]
    pos = (slots + uu) / n
    j = np.searchsorted(cdf, pos, side="left")
    j = np.minimum(j, len(cdf) - 1)
    return goff + j
