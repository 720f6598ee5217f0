"""Tensor-core (tcgen05) descriptor matching: the screen + exact re-score path must return the
same indices as the reference loop (oracle) and as the exact CUDA-core kernel, bit for bit."""
import importlib

import numpy as np
import pytest

syn = importlib.import_module("3dvision_b200.synthetic")
pytestmark = pytest.mark.gpu

EXACT, TC = 1, 2


def run_match(ctx, mode, sd, td, rows=None):
    n_src, n_tgt = sd.shape[0], td.shape[0]
    ctx.set_clouds(np.zeros((n_src, 3), np.float32), np.zeros((n_tgt, 3), np.float32))
    ctx.set_features(sd, td)
    ctx.set_match_mode(mode)
    try:
        if rows is None:
            ctx.match_features()
        else:
            ctx.set_correspondences(np.zeros(n_src, np.uint32))
            ctx.match_features(*rows)
        return ctx.get_correspondences()
    finally:
        ctx.set_match_mode(0)


@pytest.mark.parametrize("n_src,n_tgt", [(1, 1), (7, 5), (128, 256), (129, 257), (300, 1000), (1000, 300), (3000, 2500)])
def test_tc_match_equals_oracle(ctx, oracle, n_src, n_tgt):
    rng = np.random.default_rng(17 * n_src + n_tgt)
    sd = syn.histograms(n_src, rng); td = syn.histograms(n_tgt, rng)
    assert np.array_equal(run_match(ctx, TC, sd, td), oracle.match_features(sd, td))


def test_tc_match_near_duplicates_and_exact_ties(ctx, oracle):
    """Rows that tie exactly (planar scenes) and rows that differ in the last bits: the screen must let
    every contender through and the exact key must pick the reference's (lowest) index."""
    rng = np.random.default_rng(23)
    base = syn.histograms(40, rng)
    td = base[rng.integers(0, 40, 1500)].copy()
    jitter = rng.integers(0, 3, td.shape).astype(np.float32) * np.float32(2 ** -24)      # +-1 ulp scale perturbations
    td = (td + jitter * (rng.random(td.shape) < 0.3)).astype(np.float32)
    sd = base[rng.integers(0, 40, 700)].copy()
    want = oracle.match_features(sd, td)
    assert np.array_equal(run_match(ctx, TC, sd, td), want)
    assert np.array_equal(run_match(ctx, EXACT, sd, td), want)


def test_tc_match_realistic_noise(ctx, oracle):
    c = syn.ransac_case(n_src=2500, n_tgt=4000, seed=31, max_iterations=10)
    want = oracle.match_features(c.source_desc, c.target_desc)
    assert np.array_equal(run_match(ctx, TC, c.source_desc, c.target_desc), want)


def test_tc_match_unnormalised_and_negative_descriptors(ctx, oracle):
    """The API does not require L1-normalised rows; the error band scales with the norms."""
    rng = np.random.default_rng(29)
    sd = (rng.standard_normal((600, 33)) * 37.0).astype(np.float32)
    td = (rng.standard_normal((900, 33)) * 37.0).astype(np.float32)
    td[100:140] = sd[:40] + rng.normal(0, 1e-3, (40, 33)).astype(np.float32)
    assert np.array_equal(run_match(ctx, TC, sd, td), oracle.match_features(sd, td))
    sd *= np.float32(1e-12); td *= np.float32(1e-12)          # tiny magnitudes
    assert np.array_equal(run_match(ctx, TC, sd, td), oracle.match_features(sd, td))


def test_tc_match_non_finite_input_falls_back_to_exact_rescoring(ctx, oracle):
    rng = np.random.default_rng(37)
    sd = syn.histograms(200, rng); td = syn.histograms(400, rng)
    td[17, 3] = np.inf; td[250, 0] = np.nan; sd[5, 5] = np.float32(3e30)
    assert np.array_equal(run_match(ctx, TC, sd, td), oracle.match_features(sd, td))


def test_tc_match_row_range(ctx, oracle):
    rng = np.random.default_rng(41)
    sd = syn.histograms(1000, rng); td = syn.histograms(800, rng)
    want = oracle.match_features(sd, td)
    got = run_match(ctx, TC, sd, td, rows=(130, 777))
    assert np.array_equal(got[130:777], want[130:777]) and not got[:130].any() and not got[777:].any()


def test_tc_match_equals_exact_kernel_at_scale(ctx):
    """40k x 60k (2.4e9 pairs): too big for the CPU oracle in a unit test; the two GPU kernels, which share no
    arithmetic except the final exact distance, must agree on every index."""
    c = syn.ransac_case(n_src=40_000, n_tgt=60_000, seed=43, max_iterations=10)
    a = run_match(ctx, TC, c.source_desc, c.target_desc)
    b = run_match(ctx, EXACT, c.source_desc, c.target_desc)
    assert np.array_equal(a, b)
    inl = c.true_match >= 0
    assert (a[inl] == c.true_match[inl]).mean() > 0.99
