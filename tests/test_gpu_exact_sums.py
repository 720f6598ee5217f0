"""The device's exact-sequential-sum passes (terms -> block summaries -> TMA-fed warp walk, csrc/b3d_ess.cuh) against
native in-order fp32 addition, through the C-ABI diagnostic b3d_sequential_sum.

tests/test_ess_host.py hammers the shared host/device arithmetic on the CPU; this file runs the real kernels — staging
ring, segmented summary scan, super-block walk, term-by-term fallbacks — on inputs built to hit every exit, at sizes
around every chunk / super-block / ring boundary.  The default ICP mode and the RANSAC rmse ride on these passes
(src/registration.cpp:277, 351-357, 377-391), so the bar is the same BITS as `for (x : terms) s += x`."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _seq(x):
    """fp32 in-order sum: np.add.accumulate is a plain left-to-right loop in the array's dtype."""
    x = np.ascontiguousarray(x, np.float32)
    if x.size == 0:
        return np.float32(0.0)
    with np.errstate(all="ignore"):
        return np.add.accumulate(x, dtype=np.float32)[-1]


def _same_bits(a, b):
    a, b = np.float32(a), np.float32(b)
    return a.view(np.uint32) == b.view(np.uint32) or (np.isnan(a) and np.isnan(b))


@pytest.fixture(scope="module")
def ctx(b3d):
    with b3d.Context(0) as c:
        yield c


def test_accumulate_is_the_sequential_loop():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(5000).astype(np.float32)
    s = np.float32(0.0)
    for v in x:
        s = np.float32(s + v)
    assert _same_bits(s, _seq(x))


SIZES = [0, 1, 2, 31, 32, 33, 1023, 1024, 1025, 4095, 4096, 4097, 5 * 4096, 5 * 4096 + 1, 6 * 4096, 6 * 4096 + 31, 7 * 4096 - 1,
         12 * 4096 + 1024, 100_003, 300_000]


@pytest.mark.parametrize("n", SIZES)
def test_positive_terms_at_every_boundary(ctx, n):
    rng = np.random.default_rng(n)
    x = (rng.random(n, dtype=np.float32) ** 2) * np.float32(1e-6)          # like squared distances
    got, st = ctx.sequential_sum(x)
    assert _same_bits(got, _seq(x))
    if n >= 100_000:
        assert st[1] < 400                                                 # a monotone sum folds nearly every block


@pytest.mark.parametrize("n", [33, 4097, 6 * 4096 + 31, 100_003])
def test_signed_terms_at_boundaries(ctx, n):
    rng = np.random.default_rng(1000 + n)
    x = rng.standard_normal(n).astype(np.float32)
    got, _ = ctx.sequential_sum(x)
    assert _same_bits(got, _seq(x))


@pytest.mark.parametrize("seed", range(8))
def test_signed_random_walk(ctx, seed):
    rng = np.random.default_rng(100 + seed)
    n = 200_000 + 37 * seed
    x = rng.standard_normal(n).astype(np.float32) * np.float32(10.0 ** rng.integers(-6, 3))
    got, _ = ctx.sequential_sum(x)
    assert _same_bits(got, _seq(x))


def test_sum_returning_to_zero(ctx):
    rng = np.random.default_rng(7)
    h = rng.standard_normal(50_000).astype(np.float32)
    x = np.concatenate([h, -h[::-1], h[:5000]])
    got, _ = ctx.sequential_sum(x)
    assert _same_bits(got, _seq(x))


def test_exact_ties_and_power_of_two_terms(ctx):
    x = np.empty(20_000, np.float32)
    x[0] = 1.0
    x[1:] = np.float32(2.0 ** -24)                                         # every add is a tie
    got, _ = ctx.sequential_sum(x)
    assert _same_bits(got, _seq(x)) and got == np.float32(1.0)
    x[1:] = np.float32(2.0 ** -24) * np.float32(1.5)                       # just above the tie: every add moves the sum
    got, _ = ctx.sequential_sum(x)
    assert _same_bits(got, _seq(x))
    rng = np.random.default_rng(8)
    x = (np.float32(2.0) ** rng.integers(-30, 4, 100_000)).astype(np.float32) * rng.choice(np.float32([1, -1, 1.5, 0.75]), 100_000)
    got, _ = ctx.sequential_sum(x)
    assert _same_bits(got, _seq(x))


def test_mixed_magnitudes_and_zeros(ctx):
    rng = np.random.default_rng(9)
    n = 150_000
    x = rng.standard_normal(n).astype(np.float32) * (np.float32(10.0) ** rng.integers(-12, 6, n).astype(np.float32))
    x[rng.random(n) < 0.2] = 0.0
    x[rng.random(n) < 0.01] = -0.0
    got, _ = ctx.sequential_sum(x)
    assert _same_bits(got, _seq(x))


def test_mostly_skipped_records(ctx):
    """ICP at a tight threshold: most records contribute +0, the rest are spread thin (zero-on-zero blocks, then sparse)."""
    rng = np.random.default_rng(12)
    n = 300_000
    x = np.zeros(n, np.float32)
    got, st = ctx.sequential_sum(x)
    assert _same_bits(got, np.float32(0.0)) and st[1] == 0
    hit = rng.random(n) < 0.003
    x[hit] = rng.standard_normal(int(hit.sum())).astype(np.float32) * np.float32(1e-3)
    got, _ = ctx.sequential_sum(x)
    assert _same_bits(got, _seq(x))
    x[: n // 2] = 0.0                                                      # a long zero prefix, then terms
    got, _ = ctx.sequential_sum(x)
    assert _same_bits(got, _seq(x))


def test_integer_valued_terms(ctx):
    rng = np.random.default_rng(10)
    x = rng.integers(-3, 4, 400_000).astype(np.float32) * np.float32(0.5)
    got, _ = ctx.sequential_sum(x)
    assert _same_bits(got, _seq(x))
    x = np.full(100_000, 1.0, np.float32)
    x[::3] = 16777216.0                                                    # +1 becomes a tie, then is lost
    got, _ = ctx.sequential_sum(x)
    assert _same_bits(got, _seq(x))


def test_tiny_and_subnormal_sums(ctx):
    rng = np.random.default_rng(11)
    for scale in (1e-38, 1e-30, 1e-44):
        x = rng.standard_normal(50_000).astype(np.float32) * np.float32(scale)
        got, _ = ctx.sequential_sum(x)
        assert _same_bits(got, _seq(x))


def test_huge_terms_overflow_and_non_finite(ctx):
    rng = np.random.default_rng(13)
    x = rng.standard_normal(40_000).astype(np.float32) * np.float32(1e37)
    got, _ = ctx.sequential_sum(x)                                         # may overflow to +-inf or inf - inf = nan on the way
    assert _same_bits(got, _seq(x))
    x = rng.standard_normal(40_000).astype(np.float32)
    x[12_345] = np.inf
    got, _ = ctx.sequential_sum(x)
    assert _same_bits(got, _seq(x)) and np.isinf(got)
    x[30_000] = -np.inf
    got, _ = ctx.sequential_sum(x)
    assert np.isnan(got) and np.isnan(_seq(x))
    x = rng.standard_normal(10_000).astype(np.float32)
    x[77] = np.nan
    got, _ = ctx.sequential_sum(x)
    assert np.isnan(got)


def test_cancellation_between_large_neighbours(ctx):
    """ICP's ATb chains: large terms of both signs that nearly cancel, so the running sum hops across many binades."""
    rng = np.random.default_rng(14)
    n = 120_000
    big = rng.standard_normal(n // 2).astype(np.float32) * np.float32(100.0)
    x = np.empty(n, np.float32)
    x[0::2] = big
    x[1::2] = -big + rng.standard_normal(n // 2).astype(np.float32) * np.float32(1e-4)
    got, _ = ctx.sequential_sum(x)
    assert _same_bits(got, _seq(x))


@pytest.mark.parametrize("seed", range(4))
def test_piecewise_regimes(ctx, seed):
    """Concatenated stretches of different character: each hands the next a running sum it was not summarised for."""
    rng = np.random.default_rng(200 + seed)
    parts = []
    for _ in range(12):
        m = int(rng.integers(1, 30_000))
        kind = int(rng.integers(0, 5))
        if kind == 0:
            parts.append(np.zeros(m, np.float32))
        elif kind == 1:
            parts.append((rng.random(m, dtype=np.float32) ** 2) * np.float32(10.0 ** rng.integers(-8, 2)))
        elif kind == 2:
            parts.append(rng.standard_normal(m).astype(np.float32) * np.float32(10.0 ** rng.integers(-8, 4)))
        elif kind == 3:
            parts.append(np.full(m, np.float32(2.0 ** int(rng.integers(-26, 2))), np.float32))
        else:
            h = rng.standard_normal(m).astype(np.float32)
            parts.append(np.concatenate([h, -h]))
    x = np.concatenate(parts)
    got, _ = ctx.sequential_sum(x)
    assert _same_bits(got, _seq(x))


def test_repeatable_and_context_reuse(ctx):
    rng = np.random.default_rng(15)
    a = rng.standard_normal(70_000).astype(np.float32)
    b = rng.standard_normal(3_000).astype(np.float32)                      # a shorter call after a longer one: stale buffers behind it
    ra, _ = ctx.sequential_sum(a)
    rb, _ = ctx.sequential_sum(b)
    ra2, _ = ctx.sequential_sum(a)
    assert _same_bits(ra, _seq(a)) and _same_bits(rb, _seq(b)) and _same_bits(ra, ra2)


@pytest.mark.parametrize("seed", [55067, 55824, 56985, 59046, 61030] + list(range(300, 420)))
def test_adversarial_sequences(ctx, seed):
    """The soak generator (scripts/fuzz_exact_sums.py); the first five seeds are the ones that found the margin-0 acceptance."""
    import importlib
    syn = importlib.import_module("3dvision_b200.synthetic")
    x = syn.adversarial_terms(np.random.default_rng(seed))
    got, _ = ctx.sequential_sum(x)
    assert _same_bits(got, _seq(x))


@pytest.mark.parametrize("seed", [17383] + list(range(501, 700, 2)))
def test_sparse_mixed_magnitude_sequences(ctx, seed):
    """ICP-like sparse hits with per-term magnitudes over many decades (synthetic.sparse_mixed_terms); see
    tests/test_ess_host.py::test_guess_equals_the_predecessors_next_guess for what they are after."""
    import importlib
    syn = importlib.import_module("3dvision_b200.synthetic")
    x = syn.sparse_mixed_terms(np.random.default_rng(seed))
    got, _ = ctx.sequential_sum(x)
    assert _same_bits(got, _seq(x))


def test_icp_sum_that_sat_on_a_rounding_tie(ctx):
    idx = [103, 226, 379, 633, 662, 786, 891, 917, 1242, 1340, 1404, 1497, 1551, 1640, 1694, 2041]
    val = [2.13406831e-02, 4.87659033e-03, -1.04450270e-12, -1.28960852e-02, -4.93540495e-12, -7.21113978e-13, 1.03771625e-12,
           2.30799165e-11, 6.08017258e-14, 8.29449818e-02, -1.04637252e-12, 2.15440299e-02, 5.47327101e-04, -2.79879256e-13,
           1.01686455e-02, -4.61007643e-13]
    x = np.zeros(2068, np.float32)
    x[idx] = np.float32(val)
    got, _ = ctx.sequential_sum(x)
    assert _same_bits(got, _seq(x))
