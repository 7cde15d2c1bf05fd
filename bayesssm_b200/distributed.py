"""Chain-sharded PMMH across the GPUs of one node: one process per GPU, chains split by GLOBAL id, no
communication until one final gather of the draws (replaces the future_lapply fan-out / fan-in of
R/pmmh.R:512-535).  Philox streams are keyed by the global chain id, so the gathered result is identical
to a single-GPU run of all chains."""
from __future__ import annotations

import numpy as np


def shard_chains(num_chains: int, rank: int, world: int):
    """Contiguous block partition: ranks [0, num_chains % world) get one extra chain.  Returns (base, count)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    q, r = divmod(int(num_chains), int(world))
    count = q + (1 if rank < r else 0)
    base = rank * q + min(rank, r)
    return base, count


def gather_chain_arrays(local: dict, num_chains: int, rank: int, world: int, device=None):
    """all_gather of per-chain arrays (first axis = local chain) into global chain order.  Uses the default
    torch.distributed process group (NCCL with CUDA tensors on GPUs, gloo with CPU tensors in tests)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return {k: np.asarray(v) for k, v in local.items()}
    counts = [shard_chains(num_chains, r, world)[1] for r in range(world)]
    cmax = max(counts)
    out = {}
    for k in sorted(local):
        a = np.ascontiguousarray(local[k])
        tail = a.shape[1:]
        pad = np.zeros((cmax,) + tail, dtype=a.dtype)
        pad[:a.shape[0]] = a
        t = torch.from_numpy(pad)
        if device is not None:
            t = t.to(device)
        bufs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(bufs, t)
        parts = [b.cpu().numpy()[:counts[r]] for r, b in enumerate(bufs)]
        out[k] = np.concatenate(parts, axis=0)
    return out


def pmmh_sharded(run_local, num_chains: int, rank: int, world: int, device=None,
                 keys=("theta_chain", "loglike_chain", "n_accept", "target_n", "status")):
    """run_local(chain_id_base, count) -> dict of per-chain arrays for this rank's shard (e.g. a closure over
    bayesssm_b200.pmmh.run_chains).  Returns the gathered dict on every rank."""
    if num_chains < world:     # checked on EVERY rank before any work: a rank left without chains must not leave the others in the gather
        raise ValueError("more ranks than chains")
    base, count = shard_chains(num_chains, rank, world)
    local = run_local(base, count)
    return gather_chain_arrays({k: local[k] for k in keys if k in local}, num_chains, rank, world, device)


def replicate_filters_sharded(y, num_particles, model, num_filters: int, rank: int, world: int, device=None, ctx=None,
                              keys=("loglike", "n_resampled", "status"), **kwargs):
    """`num_filters` replicate bootstrap filters of one model (config C3) split across the ranks by GLOBAL filter id
    (Philox stream = filter id, so a filter's estimate does not depend on where it runs); one all_gather of the
    per-filter results at the end.  kwargs go to filters.batched_bootstrap_filter (resample_algorithm, precision,
    seed, engine, model parameters ...)."""
    from .filters import _particle_filter_core

    def run_local(base, count):
        params = {k: v for k, v in kwargs.items() if k not in ("resample_algorithm", "resample_fn", "threshold", "precision", "seed", "engine")}
        from . import _native as nat
        out = _particle_filter_core(y, num_particles, model, "BPF", kwargs.get("resample_algorithm", "SISAR"),
                                    kwargs.get("resample_fn", "stratified"), kwargs.get("threshold"), False, None, params,
                                    kwargs.get("precision", "f32"), kwargs.get("seed", 0), ctx, num_filters=count,
                                    engine=kwargs.get("engine", nat.ENGINE_AUTO), stream_base=base)
        return out
    return pmmh_sharded(run_local, num_filters, rank, world, device, keys=keys)
