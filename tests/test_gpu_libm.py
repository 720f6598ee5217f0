"""The ICP update's rotation on the device (b3d_euler_rotations: glibc's sinf / cosf as compiled by nvcc from
csrc/b3d_libm.cuh, Eigen's SSE quaternion product, toRotationMatrix) against the oracle, which calls the installed libm —
over every argument range the range reduction distinguishes: |x| < 2^-12 (returned as is), < pi/4 (polynomial only),
< 120 (FMA reduction), beyond (4/pi table), and non-finite.  tests/test_libm_host.py checks the same header, compiled by g++,
against libm on the whole float range; this file checks that the device build computes the same bits."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _angles(rng, n, lo_exp, hi_exp):
    mag = np.float32(2.0) ** rng.uniform(lo_exp, hi_exp, (n, 3)).astype(np.float32)
    return (mag * rng.choice(np.float32([1, -1]), (n, 3))).astype(np.float32)


@pytest.mark.parametrize("name,lo,hi", [("tiny", -40.0, -11.0), ("icp-like", -14.0, -4.0), ("below pi/4", -4.0, -0.35),
                                        ("moderate", -1.0, 7.9), ("large", 7.9, 30.0), ("huge", 30.0, 127.0)])
def test_update_rotation_bits(b3d, oracle, name, lo, hi):
    rng = np.random.default_rng(int(1000 * (hi - lo) + 7 * hi + 100))
    a = _angles(rng, 4000, lo, hi)                           # half-angles are what sinf / cosf see: same ranges, one binade down
    with b3d.Context(0) as ctx:
        got = ctx.euler_rotations(a)
    for i in range(a.shape[0]):
        want = oracle.euler_xyz(float(a[i, 0]), float(a[i, 1]), float(a[i, 2]))
        assert np.array_equal(got[i].view(np.uint32), want.view(np.uint32)), (name, a[i])


def test_update_rotation_special_arguments(b3d, oracle):
    a = np.float32([[0.0, 0.0, 0.0], [-0.0, 0.0, -0.0], [np.pi, -np.pi, np.pi / 2], [2.0 ** -11, -2.0 ** -11, 2.0 ** -12],
                    [1.5707964, 3.1415927, 6.2831855], [240.0, -239.99998, 240.00002], [1e-45, -1e-45, 1e-38],
                    [np.inf, 0.0, 0.0], [0.0, -np.inf, 0.0], [np.nan, 1.0, 2.0], [3.4e38, -3.4e38, 1.0]])
    with b3d.Context(0) as ctx:
        got = ctx.euler_rotations(a)
        assert ctx.euler_rotations(np.zeros((0, 3), np.float32)).shape == (0, 3, 3)
    with np.errstate(all="ignore"):
        for i in range(a.shape[0]):
            want = oracle.euler_xyz(float(a[i, 0]), float(a[i, 1]), float(a[i, 2]))
            same = (got[i].view(np.uint32) == want.view(np.uint32)) | (np.isnan(got[i]) & np.isnan(want))
            assert same.all(), a[i]
