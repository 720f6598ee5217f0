"""Deterministic synthetic workloads for the registration hot path (numpy only).

These build the inputs of BASELINE.json's configs.  Seeds are
``1234 + config id``.  Nothing here touches the GPU or the oracle; tests and
bench.py feed the same arrays to both.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


def _rot(axis, angle_rad):
    axis = np.asarray(axis, np.float64)
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(angle_rad) * K + (1 - np.cos(angle_rad)) * (K @ K)


def rigid(axis, angle_deg, t) -> np.ndarray:
    T = np.eye(4)
    T[:3, :3] = _rot(axis, np.deg2rad(angle_deg))
    T[:3, 3] = t
    return T


def apply(T, pts):
    return (pts.astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)


def _bumpy_torus_xyz(u, v, R, r):
    rho = R * (1.0 + 0.20 * np.cos(2.0 * u) + 0.10 * np.sin(3.0 * u))      # major radius varies with u:
    rr = r * (1.0 + 0.30 * np.cos(u + 1.0))                                 # not a surface of revolution,
    x = (rho + rr * np.cos(v)) * np.cos(u)                                  # so the pose is fully observable
    y = (rho + rr * np.cos(v)) * np.sin(u)
    z = rr * np.sin(v) + 0.25 * R * np.sin(2.0 * u + 0.3)
    return np.stack([x, y, z], 1)


def torus(n, rng, R=0.08, r=0.03, center=(0.0, 0.0, 0.6)):
    """n points on an asymmetric ("bumpy") torus, metres, with unit outward normals.

    Normals are the normalised cross product of central-difference tangents in float64
    (accurate to ~1e-9), oriented away from the tube's centre curve.
    """
    u = rng.uniform(0, 2 * np.pi, n)
    v = rng.uniform(0, 2 * np.pi, n)
    P = _bumpy_torus_xyz(u, v, R, r)
    h = 1e-6
    du = (_bumpy_torus_xyz(u + h, v, R, r) - _bumpy_torus_xyz(u - h, v, R, r)) / (2 * h)
    dv = (_bumpy_torus_xyz(u, v + h, R, r) - _bumpy_torus_xyz(u, v - h, R, r)) / (2 * h)
    nrm = np.cross(du, dv)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    axis_pt = _bumpy_torus_xyz(u, v, R, 0.0)                                # centre curve of the tube
    flip = np.sum(nrm * (P - axis_pt), axis=1) < 0
    nrm[flip] *= -1.0
    return (P + np.asarray(center)).astype(np.float32), nrm.astype(np.float32)


_ROUGH = np.random.default_rng(977)
_ROUGH_F = _ROUGH.integers(1, 14, 32); _ROUGH_G = _ROUGH.integers(0, 7, 32)
_ROUGH_A = _ROUGH.uniform(0.3, 1.0, 32) / np.sqrt(32.0); _ROUGH_P = _ROUGH.uniform(0, 2 * np.pi, 32)


def rough_torus(n, rng, R=0.25, r=0.09, rough=0.18, center=(0.0, 0.0, 0.6)):
    """The bumpy torus with a fixed random Fourier relief on the tube radius (features of a few centimetres, amplitude
    ~rough*r): locally distinctive geometry, so real FPFH descriptors can tell places apart. Points only."""
    u = rng.uniform(0, 2 * np.pi, n)
    v = rng.uniform(0, 2 * np.pi, n)
    relief = np.zeros(n)
    for f, g, a, p in zip(_ROUGH_F, _ROUGH_G, _ROUGH_A, _ROUGH_P):
        relief += a * np.cos(f * u + g * v + p)
    rho = R * (1.0 + 0.20 * np.cos(2.0 * u) + 0.10 * np.sin(3.0 * u))
    rr = r * (1.0 + 0.30 * np.cos(u + 1.0)) * (1.0 + rough * relief)
    x = (rho + rr * np.cos(v)) * np.cos(u)
    y = (rho + rr * np.cos(v)) * np.sin(u)
    z = rr * np.sin(v) + 0.25 * R * np.sin(2.0 * u + 0.3)
    return (np.stack([x, y, z], 1) + np.asarray(center)).astype(np.float32)


def histograms(n, rng, sparsity=0.5):
    """n x 33 non-negative, L1-normalised descriptors shaped like FPFH rows (registration.cpp:192-194)."""
    d = rng.gamma(0.6, 1.0, (n, 33))
    d *= rng.random((n, 33)) > sparsity * rng.random((n, 1))
    d[:, 0] += 1e-3
    d /= d.sum(1, keepdims=True)
    return d.astype(np.float32)


@dataclass
class IcpCase:
    source: np.ndarray        # (Ns,3) scene points
    target: np.ndarray        # (Nt,3) model points
    target_normals: np.ndarray
    T_init: np.ndarray        # 4x4, coarse source->target
    T_true: np.ndarray        # 4x4, exact source->target (before noise)
    threshold: float
    iterations: int


def icp_case(n_model=100_000, n_scene=300_000, seed=1234 + 2, noise=0.0005, threshold=0.005,
             iterations=50, init_angle_deg=0.5, init_shift=0.0015) -> IcpCase:
    """configs[1]: model on an analytic surface + normals; scene = resampled model, moved, noisy."""
    rng = np.random.default_rng(seed)
    model, normals = torus(n_model, rng)
    scene_on_model, _ = torus(n_scene, rng)
    T_model_to_scene = rigid([0.3, -0.5, 0.8], 7.0, [0.012, -0.008, 0.015])
    scene = apply(T_model_to_scene, scene_on_model)
    scene = (scene + rng.normal(0, noise, scene.shape)).astype(np.float32)
    T_true = np.linalg.inv(T_model_to_scene)
    T_init = rigid([0.7, 0.2, -0.4], init_angle_deg, [init_shift, -init_shift * 0.5, init_shift * 0.3]) @ T_true
    return IcpCase(scene, model, normals, T_init.astype(np.float32), T_true.astype(np.float32), threshold, iterations)


@dataclass
class RansacCase:
    source: np.ndarray
    target: np.ndarray
    source_desc: np.ndarray
    target_desc: np.ndarray
    voxel_size: float
    max_iterations: int
    T_true: np.ndarray
    true_match: np.ndarray    # (Ns,) index into target, -1 for outliers
    target_normals: np.ndarray = None


def ransac_case(n_src=100_000, n_tgt=100_000, seed=1234 + 3, inlier_frac=0.7, voxel=0.001,
                noise=0.0003, desc_noise=0.002, max_iterations=1_000_000) -> RansacCase:
    """configs[2]: correspondences from descriptor matching, 70 % true matches, 30 % outliers."""
    rng = np.random.default_rng(seed)
    target, tnormals = torus(n_tgt, rng, R=0.25, r=0.09)
    tdesc = histograms(n_tgt, rng)
    T_true = rigid([0.2, 0.9, -0.3], 25.0, [0.05, -0.03, 0.08])         # source -> target
    T_inv = np.linalg.inv(T_true)
    match = rng.integers(0, n_tgt, n_src)
    inl = rng.random(n_src) < inlier_frac
    src = apply(T_inv, target[match]) + rng.normal(0, noise, (n_src, 3)).astype(np.float32)
    lo, hi = src.min(0), src.max(0)
    src[~inl] = rng.uniform(lo, hi, ((~inl).sum(), 3)).astype(np.float32)
    sdesc = np.abs(tdesc[match] + rng.normal(0, desc_noise, (n_src, 33)).astype(np.float32))
    sdesc /= sdesc.sum(1, keepdims=True)
    sdesc[~inl] = histograms(int((~inl).sum()), rng)
    true_match = np.where(inl, match, -1)
    return RansacCase(src.astype(np.float32), target, sdesc.astype(np.float32), tdesc, voxel, max_iterations,
                      T_true.astype(np.float32), true_match, tnormals)


def batch_cases(n_instances=64, seed=1234 + 4, n_src=30_000, n_tgt_lo=2_000, n_tgt_hi=10_000, max_iterations=100_000):
    """configs[3]: `n_instances` object instances of the demo's scale (~30k scene points against a
    2k-10k-point model), each with its own model sampling, pose, inlier fraction and seed."""
    rng = np.random.default_rng(seed)
    cases = []
    for i in range(n_instances):
        n_tgt = int(rng.integers(n_tgt_lo, n_tgt_hi + 1))
        cases.append(ransac_case(n_src=n_src, n_tgt=n_tgt, seed=seed * 1000 + i, inlier_frac=float(rng.uniform(0.5, 0.8)),
                                 voxel=0.002, noise=0.0004, max_iterations=max_iterations))
    return cases


def rotation_error(Ta, Tb) -> float:
    """Frobenius norm of the rotation-block difference (north_star tolerance: 1e-5)."""
    return float(np.linalg.norm(np.asarray(Ta, np.float64)[:3, :3] - np.asarray(Tb, np.float64)[:3, :3]))


def translation_error(Ta, Tb) -> float:
    """Euclidean norm of the translation difference in metres (north_star tolerance: 1e-6)."""
    return float(np.linalg.norm(np.asarray(Ta, np.float64)[:3, 3] - np.asarray(Tb, np.float64)[:3, 3]))


def adversarial_terms(rng) -> np.ndarray:
    """A term sequence for the exact-sequential-sum checks (csrc/b3d_ess.cuh): up to 15 stretches of different character —
    zeros, squared-distance-like positives, signed noise over 15 decades, constant powers of two (ties), mirrored walks that
    return to the start, per-term random magnitudes, sparse hits among zeros, integer grids — each handing the next a
    running sum it was not summarised for.  Seeded, so a failing case is a seed."""
    parts = []
    for _ in range(int(rng.integers(1, 16))):
        m = int(rng.integers(1, 40_000))
        kind = int(rng.integers(0, 8))
        if kind == 0:
            p = np.zeros(m, np.float32)
        elif kind == 1:
            p = (rng.random(m, dtype=np.float32) ** 2) * np.float32(10.0 ** rng.integers(-10, 3))
        elif kind == 2:
            p = rng.standard_normal(m).astype(np.float32) * np.float32(10.0 ** rng.integers(-10, 5))
        elif kind == 3:
            p = np.full(m, np.float32(2.0 ** int(rng.integers(-28, 3))) * np.float32(rng.choice([1.0, -1.0, 1.5, 3.0])), np.float32)
        elif kind == 4:
            h = rng.standard_normal(m).astype(np.float32) * np.float32(10.0 ** rng.integers(-4, 3))
            p = np.concatenate([h, -h[::-1]])
        elif kind == 5:
            p = rng.standard_normal(m).astype(np.float32) * (np.float32(10.0) ** rng.integers(-12, 6, m).astype(np.float32))
        elif kind == 6:
            p = np.zeros(m, np.float32)
            hit = rng.random(m) < 10.0 ** rng.uniform(-3.5, -0.5)
            p[hit] = rng.standard_normal(int(hit.sum())).astype(np.float32) * np.float32(10.0 ** rng.integers(-6, 1))
        else:
            p = rng.integers(-4, 5, m).astype(np.float32) * np.float32(2.0 ** int(rng.integers(-6, 20)))
        parts.append(p)
    return np.concatenate(parts)


def random_icp_case(seed):
    """One case of the ICP soak (scripts/fuzz_registration.py): random sizes, noise, threshold, start pose, iteration cap
    and error metric.  Returns (IcpCase, point_to_plane).  Tight thresholds leave a handful of correspondences, so the
    6x6 system is singular and the update angles are large — the cases that exposed CUDA-vs-glibc sinf / cosf."""
    rng = np.random.default_rng(seed)
    n_model = int(rng.integers(40, 6000))
    n_scene = int(rng.integers(40, 12000))
    noise = float(10.0 ** rng.uniform(-4.5, -2.5))
    thr = float(10.0 ** rng.uniform(-3.3, -1.7))
    iters = int(rng.choice([1, 2, 3, 5, 10, 30, 60]))
    plane = bool(rng.integers(0, 2))
    c = icp_case(n_model=n_model, n_scene=n_scene, seed=seed, noise=noise, threshold=thr, iterations=iters,
                 init_angle_deg=float(rng.uniform(0.0, 4.0)), init_shift=float(rng.uniform(0.0, 0.006)))
    return c, plane


def random_ransac_case(seed):
    """One case of the ransacRegistration soak: random sizes, inlier ratio, voxel, noise, hypothesis count, confidence.
    Returns (RansacCase, confidence)."""
    rng = np.random.default_rng(seed)
    n_src = int(rng.integers(30, 5000))
    n_tgt = int(rng.integers(30, 5000))
    H = int(rng.choice([1, 10, 300, 1000, 4000]))
    conf = float(rng.choice([0.05, 0.5, 0.999, 2.0]))
    c = ransac_case(n_src=n_src, n_tgt=n_tgt, seed=seed, inlier_frac=float(rng.uniform(0.05, 0.95)),
                    voxel=float(10.0 ** rng.uniform(-3.3, -2.0)), noise=float(10.0 ** rng.uniform(-4.5, -3.0)), max_iterations=H)
    return c, conf


def random_cloud(seed):
    """One cloud of the front-end soak (scripts/fuzz_features.py): a random shape (surface, box, clusters with exact
    duplicates, plane, line, lattice — the degenerate ones make the covariance eigen-solver and the Darboux frames hit
    their edge cases), a random extent, and the stage parameters to run it with.
    Returns (xyz float32 (n,3), voxel, k, radius)."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 4000))
    extent = float(10.0 ** rng.uniform(-1.5, 0.7))
    kind = int(rng.integers(0, 7))
    if kind == 0:
        xyz, _ = torus(n, rng, R=0.4 * extent, r=0.15 * extent)
        xyz = xyz + rng.normal(0, 0.002 * extent, xyz.shape)
    elif kind == 1:
        xyz = rng.uniform(-extent, extent, (n, 3))
    elif kind == 2:
        centres = rng.uniform(-extent, extent, (max(1, n // 40), 3))
        xyz = centres[rng.integers(0, len(centres), n)] + rng.normal(0, 0.01 * extent, (n, 3))
        dup = rng.random(n) < 0.2
        xyz[dup] = xyz[rng.integers(0, n, int(dup.sum()))]
    elif kind == 3:
        xyz = np.zeros((n, 3))
        xyz[:, :2] = rng.uniform(-extent, extent, (n, 2))
        xyz[:, 2] = 0.25 * extent
    elif kind == 4:
        t = rng.uniform(-extent, extent, n)
        xyz = np.stack([t, 0.5 * t, -0.25 * t], 1)
    elif kind == 5:
        side = max(1, int(round(n ** (1.0 / 3.0))))
        g = np.stack(np.meshgrid(np.arange(side), np.arange(side), np.arange(side), indexing="ij"), -1).reshape(-1, 3)
        xyz = (g[rng.permutation(len(g))] - side / 2.0) * (2.0 * extent / side)
    else:
        xyz = rng.normal(0, extent, (n, 3)) * np.array([1.0, 0.05, 0.001])
    xyz = np.ascontiguousarray(xyz, np.float32)
    voxel = float(extent * 10.0 ** rng.uniform(-2.0, -0.3))
    k = int(rng.choice([3, 5, 10, 30, 30, 30, 64, 128]))
    radius = float(extent * 10.0 ** rng.uniform(-1.6, 0.3))
    return xyz, voxel, k, radius


def random_descriptors(seed):
    """Source / target descriptor sets of the matcher soak (scripts/fuzz_match.py): random sizes and a mix of populations
    — histograms of random sparsity, noisy copies of target rows, exact duplicates inside the target (the lowest index must
    win, src/registration.cpp:222-229 keeps the first strict minimum), rows scaled far above the rest (wide screen band),
    all-zero rows.  Returns (source (ns,33), target (nt,33)) float32."""
    rng = np.random.default_rng(seed)
    ns, nt = int(rng.integers(1, 3000)), int(rng.integers(1, 3000))
    td = histograms(nt, rng, sparsity=float(rng.uniform(0.0, 0.9)))
    if rng.random() < 0.5 and nt > 4:                       # exact duplicates inside the target
        m = int(rng.integers(1, nt // 2 + 1))
        td[rng.integers(0, nt, m)] = td[rng.integers(0, nt, m)]
    if rng.random() < 0.3:                                  # a few rows far from the rest
        td[rng.integers(0, nt, max(1, nt // 50))] *= np.float32(rng.uniform(3.0, 300.0))
    if rng.random() < 0.3:
        td[rng.integers(0, nt, max(1, nt // 100))] = 0.0
    sd = histograms(ns, rng, sparsity=float(rng.uniform(0.0, 0.9)))
    near = rng.random(ns) < rng.uniform(0.0, 1.0)           # noisy (or exact) copies of target rows
    noise = float(10.0 ** rng.uniform(-7.0, -1.5)) if rng.random() < 0.8 else 0.0
    pick = rng.integers(0, nt, int(near.sum()))
    sd[near] = np.abs(td[pick] + rng.normal(0, noise, (int(near.sum()), 33))).astype(np.float32)
    if rng.random() < 0.2:
        sd[rng.integers(0, ns, max(1, ns // 100))] = 0.0
    if rng.random() < 0.2:
        scale = np.float32(10.0 ** rng.uniform(-3.0, 3.0))  # FPFH of this code base sums to ~300 per row, not 1
        sd, td = sd * scale, td * scale
    return np.ascontiguousarray(sd, np.float32), np.ascontiguousarray(td, np.float32)


def random_scene_case(seed) -> dict:
    """One case of the whole-pipeline soak (scripts/fuzz_pipeline.py): a model cloud, a scene that is a moved, noisy,
    sometimes partial / cluttered re-sampling of it (or, rarely, something degenerate), and the parameters of
    Pipeline::processInstance (src/pipeline.cpp:86-129): voxel, normals k, FPFH radius, RANSAC budget / confidence, ICP
    threshold / iteration cap / error metric."""
    rng = np.random.default_rng(seed)
    n_model, n_scene = int(rng.integers(200, 6000)), int(rng.integers(200, 8000))
    kind = int(rng.integers(0, 6))
    if kind <= 2:
        model = rough_torus(n_model, rng)
        base = rough_torus(n_scene, rng)
    elif kind == 3:
        model, _ = torus(n_model, rng, R=0.25, r=0.09)
        base, _ = torus(n_scene, rng, R=0.25, r=0.09)
    elif kind == 4:                                           # a box shell: planar faces, ambiguous registration
        def shell(n):
            p = rng.uniform(-0.2, 0.2, (n, 3)); ax = rng.integers(0, 3, n)
            p[np.arange(n), ax] = np.where(rng.random(n) < 0.5, -0.2, 0.2)
            return p + np.array([0.0, 0.0, 0.6])
        model, base = shell(n_model), shell(n_scene)
    else:                                                     # a flat patch (rank-deficient plane ICP, degenerate normals at the rim)
        model = np.concatenate([rng.uniform(-0.3, 0.3, (n_model, 2)), np.full((n_model, 1), 0.6)], 1)
        base = np.concatenate([rng.uniform(-0.3, 0.3, (n_scene, 2)), np.full((n_scene, 1), 0.6)], 1)
    T = rigid(rng.normal(size=3), float(rng.uniform(0.0, 40.0)), rng.uniform(-0.1, 0.1, 3))
    scene = apply(np.linalg.inv(T), np.asarray(base, np.float32)) + rng.normal(0, float(10.0 ** rng.uniform(-4.5, -2.5)), (n_scene, 3))
    if rng.random() < 0.4:                                    # partial view
        keep = scene[:, int(rng.integers(0, 3))] > np.quantile(scene[:, 0], rng.uniform(0.1, 0.6))
        if keep.sum() >= 50:
            scene = scene[keep]
    if rng.random() < 0.4:                                    # clutter
        lo, hi = scene.min(0), scene.max(0)
        scene = np.concatenate([scene, rng.uniform(lo, hi, (int(rng.integers(1, max(2, len(scene) // 3))), 3))])
    voxel = float(10.0 ** rng.uniform(-2.2, -1.3))
    return {"model": np.ascontiguousarray(model, np.float32), "scene": np.ascontiguousarray(scene, np.float32), "voxel": voxel,
            "k": int(rng.choice([5, 10, 30, 30, 60])), "radius": float(voxel * rng.choice([5.0, 5.0, 2.0, 10.0, 0.7])),
            "H": int(rng.choice([50, 1000, 5000])), "conf": float(rng.choice([0.3, 0.999, 0.999, 2.0])),
            "icp_thr": float(voxel * rng.choice([0.4, 0.4, 1.0, 3.0])), "icp_iters": int(rng.choice([1, 5, 30, 200])),
            "plane": bool(rng.random() < 0.7), "T_true": T.astype(np.float32)}


def sparse_mixed_terms(rng) -> np.ndarray:
    """A second family for the exact-sum checks, shaped like the ICP normal-equation terms at a tight threshold: mostly
    +0 records with a few hits whose magnitudes are drawn PER TERM over many decades (products of coordinates, normals
    and residuals), both signs.  fp64 prefix sums of such terms are inexact and land on float rounding ties — the case in
    which a block's guess and its predecessor's next-guess, computed by two different fp64 expressions, once disagreed
    (found by scripts/fuzz_pipeline.py, seed 856)."""
    n = int(rng.integers(1, 60_000))
    x = np.zeros(n, np.float32)
    density = 10.0 ** rng.uniform(-3.5, 0.0)
    hit = rng.random(n) < density
    m = int(hit.sum())
    lo = int(rng.integers(-14, -2)); hi = lo + int(rng.integers(1, 14))
    mag = 10.0 ** rng.uniform(lo, hi, m)
    sign = np.where(rng.random(m) < rng.uniform(0.0, 1.0), -1.0, 1.0)
    x[hit] = (sign * mag).astype(np.float32)
    if rng.random() < 0.3 and m:                              # a handful of exact repeats and exact cancellations
        idx = np.flatnonzero(hit)
        j = rng.integers(0, m, max(1, m // 10)); k = rng.integers(0, m, max(1, m // 10))
        x[idx[j]] = x[idx[k]] * np.float32(rng.choice([1.0, -1.0, 0.5, 2.0]))
    return x
