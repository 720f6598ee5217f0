"""The C-ABI library: loads without a GPU, exports exactly what include/b3d.h declares,
and refuses to compute (loudly) when no CUDA device exists — there is no CPU fallback."""
import ctypes
import importlib
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "b3d.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b3d_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(b3d):
    assert header_functions() == sorted(b3d._capi.SYMBOLS)


def test_library_exports_every_declared_symbol(b3d):
    lib = ctypes.CDLL(b3d._capi.LIB_PATH)
    for name in header_functions():
        assert hasattr(lib, name), f"libb3d.so does not export {name}"


def test_library_has_sm100a_code_only(b3d):
    out = subprocess.run(["cuobjdump", "--list-elf", b3d._capi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "b3d.h"\nint main(void){ return b3d_cuda_available() < 0; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", f"-I{ROOT}/include", str(src)], check=True)


def test_strerror_covers_status_codes(b3d):
    L = b3d._capi.lib()
    seen = {L.b3d_strerror(code).decode() for code in range(0, -7, -1)}
    assert len(seen) == 7 and "unknown status" not in seen


def test_no_device_means_error_not_fallback(b3d):
    """On a box without CUDA every compute entry point must fail with B3D_ERR_NO_DEVICE."""
    if b3d.cuda_available():
        pytest.skip("a CUDA device is present")
    assert b3d.GPURegistration.isCudaAvailable() is False
    with pytest.raises(b3d.B3DError) as e:
        b3d.Context(0)
    assert e.value.status == b3d._capi.B3D_ERR_NO_DEVICE
    pts = np.zeros((10, 3), np.float32)
    with pytest.raises(RuntimeError):                       # gpu_impl.cpp:258 throws std::runtime_error
        b3d.GPURegistration.icpRefine(b3d.PointCloud(points=pts), b3d.PointCloud(points=pts), np.eye(4), 0.01)
    with pytest.raises(RuntimeError):
        b3d.Registration.icpRefine(b3d.PointCloud(points=pts), b3d.PointCloud(points=pts), np.eye(4), 0.01)


def test_product_never_touches_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may use oracle/."""
    pkg = os.path.join(ROOT, "3dvision_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text and "oracle/" not in text.replace("# oracle/", ""), f


def test_point_cloud_mirror_semantics(b3d):
    pc = b3d.PointCloud(points=np.zeros((5, 3), np.float32))
    assert pc.size() == 5 and not pc.empty() and not pc.hasNormals() and not pc.hasColors()
    pc.normals = np.zeros((5, 3), np.float32)
    assert pc.hasNormals()
    assert b3d.PointCloud().empty() and b3d.PointCloud().hasNormals()      # 0 == 0, registration.hpp:17
    r = b3d.RegistrationResult()
    assert np.array_equal(r.transformation, np.eye(4)) and r.fitness == 0.0 and r.rmse == 0.0
    assert b3d.FPFHFeatures(np.zeros((7, 33), np.float32)).size() == 7
