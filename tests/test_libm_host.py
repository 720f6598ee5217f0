"""glibc's sinf / cosf as restated for the device (3dvision_b200/csrc/b3d_libm.cuh) against the installed libm, on the CPU.

The reference's ICP update goes through std::sin / std::cos on float (Eigen AngleAxisf -> Quaternionf,
src/registration.cpp:369-371); CUDA's sinf / cosf are a different function, so the device evaluates glibc's algorithm.
The header is host/device code: compiled here with g++ and compared bit for bit with sinf / cosf of this image's libm
(2.39) over a stride of the whole float range plus every argument of the ranges ICP actually visits.  The exhaustive
2^32-argument run (0 mismatches for both functions, 13 s on 8 cores) is recorded in DESIGN.md."""
import ctypes
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "stubs", "libm_host_harness.cpp")


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("libm") / "liblibm_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++", SRC, "-o", out, "-lm"], check=True)
    L = ctypes.CDLL(out)
    L.libm_compare.restype = ctypes.c_long
    L.libm_compare.argtypes = [ctypes.c_ulonglong] * 3 + [ctypes.c_int, ctypes.POINTER(ctypes.c_uint)]
    return L


def _cpu_has_fma():
    try:
        return " fma " in open("/proc/cpuinfo").read()
    except OSError:
        return True


@pytest.mark.parametrize("which", [0, 1], ids=["sinf", "cosf"])
def test_whole_float_range_strided(lib, which):
    if not _cpu_has_fma():
        pytest.skip("libm's non-FMA variant differs on 34 arguments with 17 < |x| < 120 (b3d_libm.cuh)")
    first_bad = ctypes.c_uint(0)
    bad = lib.libm_compare(0, 1 << 32, 251, which, ctypes.byref(first_bad))          # 17M arguments incl. inf / nan / subnormals
    assert bad == 0, hex(first_bad.value)


@pytest.mark.parametrize("which", [0, 1], ids=["sinf", "cosf"])
def test_every_argument_below_pi_over_4_in_a_binade_sample(lib, which):
    """|x| < pi/4 is where every half-angle of a converging ICP lies: whole binades around 2^-12 (the `return x` edge),
    2^-6 and the last one below pi/4, both signs."""
    first_bad = ctypes.c_uint(0)
    bad = 0
    for lo, hi in ((0x39000000, 0x3A000000), (0x3C800000, 0x3D000000), (0x3F000000, 0x3F490FDB + 64)):
        bad += lib.libm_compare(lo, hi, 1, which, ctypes.byref(first_bad))
        bad += lib.libm_compare(lo | 0x80000000, hi | 0x80000000, 1, which, ctypes.byref(first_bad))
    assert bad == 0, hex(first_bad.value)


def test_special_values(lib):
    import math
    lib.libm_sin.restype = ctypes.c_float; lib.libm_sin.argtypes = [ctypes.c_float]
    lib.libm_cos.restype = ctypes.c_float; lib.libm_cos.argtypes = [ctypes.c_float]
    assert lib.libm_sin(0.0) == 0.0 and math.copysign(1.0, lib.libm_sin(-0.0)) == -1.0
    assert lib.libm_cos(0.0) == 1.0 and lib.libm_cos(-0.0) == 1.0
    assert math.isnan(lib.libm_sin(float("inf"))) and math.isnan(lib.libm_cos(float("-inf"))) and math.isnan(lib.libm_sin(float("nan")))
