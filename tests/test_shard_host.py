"""Host-side logic of the particle-sharded filter, without a GPU: the block partition exported by the C ABI and
the per-observation exchange (records -> global normaliser, cdf offsets, output-slot ranges) run by two CPU
processes over gloo, checked against the unsharded resampler of the oracle."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_covers_and_aligns():
    from bayesssm_b200 import sharding as S
    for n, w in ((1 << 20, 8), (1000, 3), (7, 2), (5, 8), (1 << 28, 8)):
        parts = [S.partition(n, w, r) for r in range(w)]
        assert parts[0][0] == 0 and sum(c for _, c in parts) == n
        for (g0, c0), (g1, _) in zip(parts, parts[1:]):
            assert g0 + c0 == g1
        assert all(g % 4 == 0 or g == n for g, _ in parts)      # Philox quads stay whole


def test_count_le_is_the_number_of_positions_below():
    from bayesssm_b200 import sharding as S
    rng = np.random.default_rng(3)
    n = 257
    u = rng.random(n)
    pos = (np.arange(n) + u) / n
    for c in list(rng.random(200)) + [0.0, 1.0, float(pos[17]), float(np.nextafter(pos[17], 0))]:
        assert S.count_le(c, n, u) == int(np.sum(pos <= c))
    us = 0.37
    for c in rng.random(50):
        assert S.count_le(c, n, us) == int(np.sum((np.arange(n) + us) / n <= c))


def test_exchange_world_size_2_gloo(tmp_path):
    """Two CPU processes (gloo), each holding a block of the log-weights: all_gather of the records, every rank
    derives the same normaliser and its own slot range, the concatenated ancestors equal the unsharded
    stratified / systematic resampler (oracle) and the log-likelihood increment equals the unsharded one."""
    script = tmp_path / "worker.py"
    script.write_text(f'''
import os, sys
sys.path.insert(0, {ROOT!r}); sys.path.insert(0, os.path.join({ROOT!r}, "tests"))
import numpy as np, torch, torch.distributed as dist
from bayesssm_b200 import sharding as S
import oracle
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
for n, seed, fn in ((4096, 1, "stratified"), (10007, 2, "stratified"), (10007, 3, "systematic"), (13, 4, "stratified")):
    rng = np.random.default_rng(seed)                       # same stream on both ranks
    lw = -0.5 * (2.5 * rng.standard_normal(n)) ** 2 + 3.0 * (np.arange(n) >= n // 3)   # uneven weight shares
    u = rng.random(n) if fn == "stratified" else float(rng.random())
    goff, nloc = S.partition(n, world, rank)
    mine = lw[goff:goff + nloc]
    rec = torch.tensor([S.local_record(mine)], dtype=torch.float64)
    bufs = [torch.empty_like(rec) for _ in range(world)]
    dist.all_gather(bufs, rec)
    records = [(float(b[0, 0]), float(b[0, 1])) for b in bufs]
    ex = S.exchange_step(records, rank, n, u)
    anc = S.local_ancestors(mine, goff, ex, n, u)
    assert len(anc) == ex["o_hi"] - ex["o_lo"]
    # gather the ancestors (ragged) and compare with the unsharded resampler
    cnt = torch.tensor([len(anc)]); cnts = [torch.empty_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    pad = torch.full((n,), -1, dtype=torch.int64); pad[:len(anc)] = torch.from_numpy(anc)
    allp = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(allp, pad)
    full = np.concatenate([p.numpy()[:int(c)] for p, c in zip(allp, cnts)])
    assert len(full) == n and ex["all_slots"][0] == 0 and ex["all_slots"][-1] == n
    w = np.exp(lw - lw.max())
    ref = oracle.resample(fn, w, np.atleast_1d(u)) - 1
    bad = np.flatnonzero(full != ref)
    assert len(bad) <= 2 and all(abs(int(full[i]) - int(ref[i])) == 1 for i in bad), (n, fn, len(bad))
    inc = lw.max() + np.log(np.sum(np.exp(lw - lw.max()))) - np.log(n)
    assert abs(ex["loglike_increment"] - inc) <= 1e-12 * abs(inc)
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
''')
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_closed_form_offspring_counts_equal_the_reference_resampler(orc):
    """The persistent kernel and the streaming engine never search: source j owns the output slots
    [F(c_{j-1}), F(c_j)) with F(c) = #{ i : (i + u_i) / n <= c } in closed form.  On the reference's own cdf
    (sequential cumsum of w / sum(w), src/resampling.cpp:20-25) the offspring counts F(c_j) - F(c_{j-1}) must equal
    the histogram of the reference's ancestors (first j with cdf[j] >= pos, clamp), slot for slot."""
    from bayesssm_b200 import sharding as S
    rng = np.random.default_rng(8)
    for n in (1, 2, 7, 1000, 4097):
        for fam in ("lognormal", "few_heavy", "uniform"):
            w = {"lognormal": np.exp(2.0 * rng.standard_normal(n)), "uniform": np.ones(n),
                 "few_heavy": np.where(rng.random(n) < 0.02, 1.0, 1e-9) + 1e-300}[fam]
            for fn in ("stratified", "systematic"):
                u = rng.random(n) if fn == "stratified" else float(rng.random())
                anc = orc.resample(fn, w, np.atleast_1d(u)) - 1
                total = 0.0
                for v in w:                      # Rcpp sugar sum: plain sequential double loop
                    total += v
                cdf = np.cumsum(w / total)       # sequential cumsum of the normalised weights
                F = np.array([S.count_le(float(c), n, u) for c in cdf])
                F[-1] = n                        # clamp: the last source takes what is left (src/resampling.cpp:33,60)
                F = np.maximum.accumulate(F)
                counts = np.diff(np.concatenate([[0], F]))
                assert counts.sum() == n
                assert np.array_equal(counts, np.bincount(anc, minlength=n)), (n, fam, fn)
                # the slots of source j are contiguous and in order: expanding the counts reproduces the ancestor vector
                assert np.array_equal(np.repeat(np.arange(n), counts), anc)
