"""GPU parity of the resamplers (through the C ABI) against the oracle restating src/resampling.cpp:
ancestor indices and the cdf are BIT-EXACT given identical weights and uniforms."""
import numpy as np
import pytest

import engine_helpers as eh
from bayesssm_b200 import _native as nat

pytestmark = pytest.mark.gpu

SIZES = [1, 2, 3, 5, 31, 32, 33, 1000, 1023, 1024, 1025, 4097, 65536, 100003, 1 << 20]


def _weights(kind, n, rng):
    if kind == "uniform":
        return rng.random(n)
    if kind == "pf":  # particle-filter-like: exp of Gaussian log-weights, many tiny
        z = rng.standard_normal(n) * 3
        return np.exp(-0.5 * z * z * 4)
    if kind == "dyadic":
        return np.full(n, 1.0 / (1 << 20))
    if kind == "degenerate":
        w = np.zeros(n)
        w[n // 2] = 1.0
        return w
    if kind == "tail":  # one heavy block then a long run of tiny weights: cdf sits at 1 -/+ ulps
        w = rng.random(n) * 1e-22
        w[: max(1, n // 8)] = rng.random(max(1, n // 8))
        return w
    if kind == "range":
        return np.ldexp(rng.random(n), -rng.integers(0, 600, n)) * (rng.random(n) > 0.2)
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["uniform", "pf", "dyadic", "degenerate", "tail", "range"])
def test_cdf_and_total_bit_exact(engine, orc, kind):
    rng = np.random.default_rng(11)
    for n in SIZES:
        w = _weights(kind, n, rng)
        if w.sum() == 0:
            w[0] = 1.0
        ref_cdf, ref_tot = orc.cdf(w)
        got_cdf, got_tot, n_serial = eh.cdf(engine, w)
        assert got_tot == ref_tot, (kind, n)
        bad = np.flatnonzero(got_cdf.view(np.int64) != ref_cdf.view(np.int64))
        assert bad.size == 0, (kind, n, bad[:5], got_cdf[bad[:5]], ref_cdf[bad[:5]])
        if kind in ("uniform", "pf") and n == 1 << 20:
            # the serial chain only walks the ~log2(n) tiles in which the running sum changes binade
            assert n_serial < 80 * 1024, n_serial


@pytest.mark.parametrize("fn", ["stratified", "systematic", "multinomial"])
@pytest.mark.parametrize("kind", ["uniform", "pf", "dyadic", "degenerate", "tail"])
def test_ancestors_bit_exact(engine, orc, fn, kind):
    rng = np.random.default_rng(23)
    for n in SIZES:
        w = _weights(kind, n, rng)
        u = rng.random(1 if fn == "systematic" else n)
        ref = orc.resample(fn, w, u)
        got = eh.resample(engine, fn, w, u)
        assert np.array_equal(ref, got), (fn, kind, n, np.flatnonzero(ref != got)[:5])


def test_reference_known_answers(engine):
    # tests/testthat/test-resampling.R:48-68,190-202 on the CUDA path
    rng = np.random.default_rng(7)
    w = np.array([0.1, 0.5, 0.1, 0.15, 0.15])
    for _ in range(200):
        s = eh.resample(engine, "stratified", w, rng.random(5))
        assert s[1] == 2 and s[2] == 2
        y = eh.resample(engine, "systematic", w, rng.random(1))
        assert y[1] == 2 and y[2] == 2
        assert (y[0] != 1 or y[3] == 3) and (y[0] != 2 or y[3] == 4)
    for fn in ("stratified", "systematic", "multinomial"):
        assert (eh.resample(engine, fn, [0, 0, 1, 0, 0], rng.random(5)) == 3).all()
    assert eh.resample(engine, "systematic", [0.25] * 4, [0.0]).tolist() == [1, 1, 2, 3]


def test_reference_error_strings(engine):
    # tests/testthat/test-resampling.R:2-28
    for fn in ("stratified", "systematic", "multinomial"):
        with pytest.raises(nat.EngineError, match="Weights must be non-negative") as e:
            eh.resample(engine, fn, [-1, 1, 2], [0.5, 0.5, 0.5])
        assert e.value.status == nat.ERR_NEGATIVE_WEIGHT
        with pytest.raises(nat.EngineError, match="Sum of weights must be greater than 0") as e:
            eh.resample(engine, fn, [0, 0, 0], [0.5, 0.5, 0.5])
        assert e.value.status == nat.ERR_ZERO_SUM


def test_proportions(engine):
    # tests/testthat/test-resampling.R:29-47
    rng = np.random.default_rng(1405)
    w = np.array([0.1, 0.2, 0.3, 0.2, 0.2])
    for fn in ("stratified", "systematic", "multinomial"):
        counts = np.zeros(5)
        for _ in range(2000):
            counts += np.bincount(eh.resample(engine, fn, w, rng.random(5)) - 1, minlength=5)
        np.testing.assert_allclose(counts / 10000, w, atol=0.03)


@pytest.mark.parametrize("kind", ["pf", "tail"])
def test_large_inputs_scan_the_tile_totals_once(engine, orc, kind, monkeypatch):
    # more than RS_PREFIX_TILES (2048) tiles: k_tile_prefix scans the tile totals once per pass instead of every tile summing its
    # predecessors (O(tiles^2) loads); the exact pipeline still reproduces the sequential cumsum bit for bit.  Also forced at small n.
    rng = np.random.default_rng(5)
    for n, force in ((1 << 22, None), (100003, "0"), (4097, "1")):
        if force is None:
            monkeypatch.delenv("BSSM_RS_PREFIX_TILES", raising=False)
        else:
            monkeypatch.setenv("BSSM_RS_PREFIX_TILES", force)
        w = _weights(kind, n, rng)
        ref_cdf, ref_tot = orc.cdf(w)
        got_cdf, got_tot, _ = eh.cdf(engine, w)
        assert got_tot == ref_tot and np.array_equal(got_cdf.view(np.int64), ref_cdf.view(np.int64)), (kind, n)
        u = rng.random(n)
        assert np.array_equal(eh.resample(engine, "stratified", w, u), orc.resample("stratified", w, u))
