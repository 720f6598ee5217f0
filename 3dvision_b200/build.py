"""In-tree build of libb3d.so (sm_100a only) with nvcc.

``python -m 3dvision_b200.build`` is not importable syntax (package name starts with a
digit); use ``python 3dvision_b200/build.py`` or ``__graft_entry__.build()``.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libb3d.so")
SOURCES = ["b3d_api.cu", "b3d_match.cu", "b3d_match_tc.cu", "b3d_ransac.cu", "b3d_icp.cu", "b3d_features.cu", "b3d_pose.cu", "b3d_dist.cu", "b3d_pool.cu"]
HEADERS = ["b3d_common.cuh", "b3d_linalg.cuh", "b3d_scan.cuh", "b3d_grid.cuh", "b3d_featmath.cuh", "b3d_ess.cuh", "b3d_libm.cuh", os.path.join("..", "..", "include", "b3d.h")]

# --fmad=false: the reference CPU build never contracts a*b+c (README.md:13, no -march), and
# bit-exact inlier counts / match indices depend on it.  Kernels that may fuse say so with
# explicit fmaf()/__fmaf_rn().  Division and sqrt stay IEEE (nvcc defaults; no --use_fast_math).
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "--fmad=false",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off",
    "-shared", "-cudart", "static", "-ldl", "-lpthread",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", OUT] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libb3d.so")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
