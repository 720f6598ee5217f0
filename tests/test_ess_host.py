"""Exact-sequential-sum core (3dvision_b200/csrc/b3d_ess.cuh) against native in-order fp32 addition, on the CPU.

The kernels that replay the reference's one-point-at-a-time sums (src/registration.cpp:341-358, 374-386, 270-279)
skip the dependent add chain with precomputed integer block summaries; the result must be the same BITS as the chain.
The header's arithmetic is host/device code, so the very same functions are compiled here with g++ and hammered with
inputs built to hit every exit: binade changes, sign changes, exact ties, cancellation, huge/tiny/zero/non-finite terms."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "stubs", "ess_host_harness.cpp")
HDR = os.path.join(HERE, "..", "3dvision_b200", "csrc", "b3d_ess.cuh")


@pytest.fixture(scope="module")
def ess(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("ess") / "libess_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++", SRC, "-o", out], check=True)
    lib = ctypes.CDLL(out)
    fp = ctypes.POINTER(ctypes.c_float)
    lib.ess_seq_sum.restype = ctypes.c_float
    lib.ess_seq_sum.argtypes = [fp, ctypes.c_long]
    lib.ess_parallel_sum.restype = ctypes.c_float
    lib.ess_parallel_sum.argtypes = [fp, ctypes.c_long, ctypes.POINTER(ctypes.c_long)]
    lib.ess_parallel_sum_devicelike.restype = ctypes.c_float
    lib.ess_parallel_sum_devicelike.argtypes = [fp, ctypes.c_long, ctypes.POINTER(ctypes.c_long)]
    lib.ess_min_usable_margin.restype = ctypes.c_long
    lib.ess_min_usable_margin.argtypes = [fp, ctypes.c_long]
    return lib


def _both(lib, x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    p = x.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    st = (ctypes.c_long * 3)()
    a = np.float32(lib.ess_seq_sum(p, x.size))
    b = np.float32(lib.ess_parallel_sum(p, x.size, st))
    # the lane-by-lane replay of the device passes (fp64 butterflies, Hillis-Steele scan, unit-carrying walk) must agree too
    c = np.float32(lib.ess_parallel_sum_devicelike(p, x.size, None))
    assert _same_bits(a, c), (a, c)
    return a, b, tuple(st)


def _same_bits(a, b):
    return a.view(np.uint32) == b.view(np.uint32) or (np.isnan(a) and np.isnan(b))


def test_header_is_the_product_header():
    assert os.path.exists(HDR) and "block_summary" in open(HDR).read()


@pytest.mark.parametrize("n", [0, 1, 2, 31, 32, 33, 1023, 1024, 1025, 4097, 300_000])
def test_positive_terms_take_the_fast_path(ess, n):
    rng = np.random.default_rng(n)
    x = (rng.random(n, dtype=np.float32) ** 2) * np.float32(1e-6)          # like squared distances / J_a^2
    a, b, st = _both(ess, x)
    assert _same_bits(a, b)
    if n >= 100_000:                                                       # monotone sum: ~20 binade changes in all
        assert st[0] > 0.97 * (n // 32) and st[1] < 200


@pytest.mark.parametrize("seed", range(6))
def test_signed_random_walk(ess, seed):
    rng = np.random.default_rng(100 + seed)
    n = 200_000 + 37 * seed
    x = rng.standard_normal(n).astype(np.float32) * np.float32(10.0 ** rng.integers(-6, 3))
    a, b, st = _both(ess, x)
    assert _same_bits(a, b)
    assert st[0] > 0.8 * (n // 32)                                         # ties and binade changes included, most blocks fold


def test_sum_returning_to_zero(ess):
    rng = np.random.default_rng(7)
    h = rng.standard_normal(50_000).astype(np.float32)
    x = np.concatenate([h, -h[::-1], h[:5000]])                            # crosses zero, every binade twice
    a, b, _ = _both(ess, x)
    assert _same_bits(a, b)


def test_exact_ties_and_power_of_two_terms(ess):
    # terms that are exact half-ulps of the running sum: round-to-even depends on the sum's parity
    x = np.empty(20_000, np.float32)
    x[0] = 1.0
    x[1:] = np.float32(2.0 ** -24)                                         # every add is a tie at first
    a, b, _ = _both(ess, x)
    assert _same_bits(a, b)
    rng = np.random.default_rng(8)
    x = (np.float32(2.0) ** rng.integers(-30, 4, 100_000)).astype(np.float32) * rng.choice(np.float32([1, -1, 1.5, 0.75]), 100_000)
    a, b, _ = _both(ess, x)
    assert _same_bits(a, b)


def test_mixed_magnitudes_and_zeros(ess):
    rng = np.random.default_rng(9)
    n = 150_000
    x = rng.standard_normal(n).astype(np.float32) * (np.float32(10.0) ** rng.integers(-12, 6, n).astype(np.float32))
    x[rng.random(n) < 0.2] = 0.0
    x[rng.random(n) < 0.01] = -0.0
    a, b, _ = _both(ess, x)
    assert _same_bits(a, b)


def test_integer_valued_terms_are_all_ties_or_exact(ess):
    rng = np.random.default_rng(10)
    x = rng.integers(-3, 4, 400_000).astype(np.float32) * np.float32(0.5)  # sum grows past 2^24 * 0.5? no — stays exact
    a, b, _ = _both(ess, x)
    assert _same_bits(a, b)
    x = np.full(40_000_0, 1.0, np.float32)
    x[::3] = 16777216.0                                                    # jumps the sum where +1 becomes a tie / is lost
    a, b, _ = _both(ess, x[:100_000])
    assert _same_bits(a, b)


def test_tiny_and_subnormal_sums(ess):
    rng = np.random.default_rng(11)
    x = rng.standard_normal(50_000).astype(np.float32) * np.float32(1e-38)
    a, b, _ = _both(ess, x)
    assert _same_bits(a, b)
    x = rng.standard_normal(50_000).astype(np.float32) * np.float32(1e-30)
    a, b, _ = _both(ess, x)
    assert _same_bits(a, b)


def test_non_finite_terms(ess):
    rng = np.random.default_rng(12)
    for bad in (np.inf, -np.inf, np.nan, 3e38):
        x = rng.random(10_000, dtype=np.float32)
        x[5000] = bad
        x[7000] = bad
        a, b, _ = _both(ess, x)
        assert _same_bits(a, b)


def test_icp_like_normal_equation_terms(ess):
    """Products J_a*J_b and J_a*r of a registration near convergence: the 28 sums of registration.cpp:343-354."""
    rng = np.random.default_rng(13)
    n = 120_000
    p = rng.random((n, 3), dtype=np.float32) * np.float32(0.3)
    nrm = rng.standard_normal((n, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    J = np.concatenate([np.cross(p, nrm).astype(np.float32), nrm], axis=1)
    r = (rng.standard_normal(n) * 5e-4).astype(np.float32)
    fast = 0
    for a_ in range(6):
        for b_ in range(a_, 6):
            s0, s1, st = _both(ess, J[:, a_] * J[:, b_])
            assert _same_bits(s0, s1)
            fast += st[0]
        s0, s1, st = _both(ess, J[:, a_] * r)
        assert _same_bits(s0, s1)
    assert fast > 0


def test_many_short_random_cases(ess):
    rng = np.random.default_rng(14)
    for _ in range(400):
        n = int(rng.integers(1, 5000))
        scale = np.float32(10.0 ** rng.integers(-8, 8))
        bias = np.float32(rng.standard_normal() * rng.choice([0.0, 0.1, 1.0, 10.0]))
        x = (rng.standard_normal(n).astype(np.float32) + bias) * scale
        a, b, _ = _both(ess, x)
        assert _same_bits(a, b)


def test_leading_and_embedded_zero_runs(ess):
    """Skipped records are +0 terms: long zero runs before the first term (sum still +0) and inside the sequence."""
    rng = np.random.default_rng(15)
    x = np.zeros(200_000, np.float32)
    x[150_000:] = rng.random(50_000, dtype=np.float32) * np.float32(1e-6)
    x[160_000:170_000] = 0.0
    a, b, st = _both(ess, x)
    assert _same_bits(a, b) and st[1] < 50                                  # the zero runs are folded, not added one by one
    x = np.zeros(100_000, np.float32); x[99_999] = np.float32(3.5)
    a, b, st = _both(ess, x)
    assert _same_bits(a, b) and a == np.float32(3.5) and st[1] <= 2
    x = np.zeros(70_000, np.float32); x[::7000] = np.float32(-0.0)          # -0 terms keep a +0 sum at +0
    a, b, _ = _both(ess, x)
    assert _same_bits(a, b) and a.view(np.uint32) == 0
    h = rng.standard_normal(3000).astype(np.float32)                        # exact cancellation back to +0, then zeros, then more terms
    x = np.concatenate([h, -h[::-1], np.zeros(5000, np.float32), h])
    a, b, _ = _both(ess, x)
    assert _same_bits(a, b)


# Seeds of synthetic.adversarial_terms that a 60 000-case soak of the device passes found (scripts/fuzz_exact_sums.py): a block
# whose partial sums come within (8, 9) units of a frame border got margin 0, and the walk's unsigned one-compare range test
# accepted it for every offset.  block_summary() now rejects margin < 1.
SOAK_FINDS = [55067, 55824, 56985, 59046, 61030]


@pytest.mark.parametrize("seed", SOAK_FINDS + list(range(40)))
def test_adversarial_sequences(ess, seed):
    import importlib
    syn = importlib.import_module("3dvision_b200.synthetic")
    x = syn.adversarial_terms(np.random.default_rng(seed))
    a, b, _ = _both(ess, x)
    assert _same_bits(a, b)


def test_margin_is_never_zero_on_a_usable_block(ess):
    """A usable summary must cover at least the exact guess: margin >= 1 whenever the tag is not kFail."""
    import importlib
    syn = importlib.import_module("3dvision_b200.synthetic")
    for seed in SOAK_FINDS:
        x = np.ascontiguousarray(syn.adversarial_terms(np.random.default_rng(seed)), np.float32)
        assert ess.ess_min_usable_margin(x.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), x.size) >= 1


def test_guess_equals_the_predecessors_next_guess(ess):
    """Sum 9 of the 11th ICP iteration of scripts/fuzz_pipeline.py seed 856 (16 hits in 2068 records): the fp64 prefix after
    the second hit sits on a float rounding tie, and an exclusive prefix formed as `inclusive - own` rounded the other way
    than the predecessor's inclusive prefix — the block map then tied two different guesses together and the walk was off
    by one ulp.  Guess and next-guess now come from the same fp64 expression."""
    idx = [103, 226, 379, 633, 662, 786, 891, 917, 1242, 1340, 1404, 1497, 1551, 1640, 1694, 2041]
    val = [2.13406831e-02, 4.87659033e-03, -1.04450270e-12, -1.28960852e-02, -4.93540495e-12, -7.21113978e-13, 1.03771625e-12,
           2.30799165e-11, 6.08017258e-14, 8.29449818e-02, -1.04637252e-12, 2.15440299e-02, 5.47327101e-04, -2.79879256e-13,
           1.01686455e-02, -4.61007643e-13]
    x = np.zeros(2068, np.float32)
    x[idx] = np.float32(val)
    a, b, _ = _both(ess, x)
    assert _same_bits(a, b)


@pytest.mark.parametrize("seed", [17383] + list(range(1, 80, 2)))
def test_sparse_mixed_magnitude_sequences(ess, seed):
    """synthetic.sparse_mixed_terms: ICP-like sparse hits with per-term magnitudes over many decades; 17383 is the seed a
    100 000-case soak of the lane-by-lane replay found against the old guess expression."""
    import importlib
    syn = importlib.import_module("3dvision_b200.synthetic")
    x = syn.sparse_mixed_terms(np.random.default_rng(seed))
    a, b, _ = _both(ess, x)
    assert _same_bits(a, b)


def test_lane_by_lane_replay_soak(ess):
    """2 000 seeded sequences (both families) through the lane-by-lane replay of the device passes; the long soaks
    (3.6M cases here, 700k on the device) are recorded in DESIGN.md."""
    import importlib
    syn = importlib.import_module("3dvision_b200.synthetic")
    fp = ctypes.POINTER(ctypes.c_float)
    for seed in range(200_000, 202_000):
        rng = np.random.default_rng(seed)
        x = np.ascontiguousarray(syn.sparse_mixed_terms(rng) if seed % 2 else syn.adversarial_terms(rng), np.float32)
        got = np.float32(ess.ess_parallel_sum_devicelike(x.ctypes.data_as(fp), x.size, None))
        with np.errstate(all="ignore"):
            want = np.add.accumulate(x, dtype=np.float32)[-1] if x.size else np.float32(0)
        assert _same_bits(got, want), seed
