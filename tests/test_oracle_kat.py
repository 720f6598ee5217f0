"""Known-answer tests that pin the CPU oracle (no GPU needed).

The reference has no tests or golden vectors for this path, so these KATs are
authored from the semantics of src/registration.cpp and from
properties any correct restatement must have.
"""
import numpy as np
import pytest

import importlib

syn = importlib.import_module("3dvision_b200.synthetic")


# ---------------------------------------------------------------- RNG
def test_mt19937_raw_known_answers(oracle):
    raw = oracle.mt19937_raw(42, 10000)
    assert list(raw[:4]) == [1608637542, 3421126067, 4083286876, 787846414]
    assert raw[9999] == 1399405940
    # the C++11 standard's check value: 10000th output of a default-seeded mt19937
    assert oracle.mt19937_raw(5489, 10000)[9999] == 4123659995


@pytest.mark.parametrize("n,first9", [
    (1000, [374, 796, 950, 183, 731, 779, 598, 596, 156]),
    (31337, [11736, 24961, 29792, 5748, 22938, 24433, 18760, 18703, 4889]),
    (100000, [37454, 79654, 95071, 18343, 73199, 77969, 59865, 59685, 15601]),
    (1000000, [374540, 796542, 950714, 183434, 731993, 779690, 598658, 596850, 156018]),
])
def test_uniform_index_known_answers(oracle, n, first9):
    assert list(oracle.uniform_indices_std(42, n, 9)) == first9


@pytest.mark.parametrize("n,rejections", [(1000, 0), (31337, 9), (100000, 47), (1000000, 648), (3, None), (1, None),
                                          (3 * 2**30, None), (2**31 + 12345, None)])
def test_lemire_mapping_equals_libstdcxx(oracle, n, rejections):
    """The hand-rolled multiply-shift + rejection mapping (what the CUDA path implements)
    reproduces std::uniform_int_distribution<size_t> draw for draw, rejections included."""
    count = 3_000_000 if rejections is not None else 200_000
    std = oracle.uniform_indices_std(42, n, count)
    lem, used = oracle.uniform_indices_lemire(42, n, count)
    assert np.array_equal(std, lem)
    if rejections is not None:
        assert used - count == rejections
    assert std.max() < n


# ---------------------------------------------------------------- small linear algebra
def test_svd3_properties(oracle):
    rng = np.random.default_rng(0)
    for k in range(200):
        M = rng.standard_normal((3, 3)).astype(np.float32) * np.float32(10.0 ** rng.integers(-6, 3))
        if k % 5 == 0:
            M[:, 2] = M[:, 0] * 2                      # rank deficient
        U, S, V = oracle.svd3(M)
        scale = np.abs(M).max()
        assert np.abs(U @ np.diag(S) @ V.T - M).max() <= 4e-6 * scale
        assert np.abs(U.T @ U - np.eye(3)).max() < 2e-6 and np.abs(V.T @ V - np.eye(3)).max() < 2e-6
        assert S[0] >= S[1] >= S[2] >= 0
        assert np.allclose(S, np.linalg.svd(M.astype(np.float64), compute_uv=False), rtol=2e-5, atol=2e-6 * scale)


def test_svd3_zero_and_diagonal(oracle):
    U, S, V = oracle.svd3(np.zeros((3, 3), np.float32))
    assert np.array_equal(U, np.eye(3)) and np.array_equal(V, np.eye(3)) and not S.any()
    U, S, V = oracle.svd3(np.diag([1.0, -3.0, 2.0]).astype(np.float32))
    assert list(S) == [3.0, 2.0, 1.0]
    assert np.array_equal(np.abs(U), np.abs(V))         # no rotation applied; only a sign flip and the sort
    assert np.array_equal(U @ np.diag(S) @ V.T, np.diag([1.0, -3.0, 2.0]))


def test_kabsch_recovers_exact_rigid_motion(oracle):
    rng = np.random.default_rng(1)
    for _ in range(50):
        T = syn.rigid(rng.standard_normal(3), rng.uniform(-170, 170), rng.uniform(-1, 1, 3))
        s = rng.uniform(-1, 1, (3, 3)).astype(np.float32)
        q = syn.apply(T, s)
        R, t = oracle.kabsch3(s, q)
        assert np.abs(R - T[:3, :3]).max() < 2e-5 * max(1.0, 1.0 / np.linalg.svd(s - s.mean(0))[1][1])
        assert abs(np.linalg.det(R.astype(np.float64)) - 1) < 1e-5


def test_kabsch_degenerate_target_gives_identity(oracle):
    """All three target points equal => H = 0 => JacobiSVD returns U = V = I => R = I (Appendix A)."""
    s = np.array([[0.1, 0.2, 0.3], [0.5, -0.1, 0.0], [0.0, 0.3, 0.9]], np.float32)
    q = np.tile(np.array([[0.25, -0.5, 1.5]], np.float32), (3, 1))
    R, t = oracle.kabsch3(s, q)
    assert np.array_equal(R, np.eye(3, dtype=np.float32))
    assert np.allclose(t, q[0] - s.mean(0), atol=1e-6)


def test_ldlt6(oracle):
    rng = np.random.default_rng(2)
    for _ in range(100):
        J = rng.standard_normal((40, 6)).astype(np.float32)
        A = (J.T @ J).astype(np.float32); b = rng.standard_normal(6).astype(np.float32)
        x = oracle.ldlt6_solve(A, b)
        ref = np.linalg.solve(A.astype(np.float64), b.astype(np.float64))
        assert np.abs(x - ref).max() < 5e-5 * max(1.0, np.abs(ref).max()) * np.linalg.cond(A.astype(np.float64)) ** 0.5
    assert not oracle.ldlt6_solve(np.zeros((6, 6), np.float32), np.ones(6, np.float32)).any()


def test_ldlt6_planar_target_zeroes_unobservable_dofs(oracle):
    """Planar target with normals (0,0,1): J[2]=J[3]=J[4]=0 exactly, so rows/cols 2..4 of ATA are zero and
    the pseudo-inverse of D gives x[2]=x[3]=x[4]=0 exactly (Appendix B; config 0 exercises this)."""
    rng = np.random.default_rng(3)
    p = rng.uniform(-1, 1, (200, 3)).astype(np.float32)
    n = np.array([0, 0, 1], np.float32)
    A = np.zeros((6, 6), np.float32); b = np.zeros(6, np.float32)
    for pi in p:
        J = np.concatenate([np.cross(pi, n), n]).astype(np.float32)
        r = np.float32(pi[2] * 0.01)
        A += np.outer(J, J).astype(np.float32); b += J * r
    x = oracle.ldlt6_solve(A, -b)
    assert x[2] == 0 and x[3] == 0 and x[4] == 0
    assert np.abs(A @ x + b)[[0, 1, 5]].max() < 1e-4


def test_euler_xyz_is_rx_ry_rz(oracle):
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(4)
    for _ in range(50):
        a, b, g = rng.uniform(-0.5, 0.5, 3)
        R = oracle.euler_xyz(a, b, g)
        assert np.abs(R - Rotation.from_euler("XYZ", [a, b, g]).as_matrix()).max() < 3e-7
    assert np.array_equal(oracle.euler_xyz(0, 0, 0), np.eye(3, dtype=np.float32))


# ---------------------------------------------------------------- feature matching
def test_match_strict_less_keeps_lowest_index(oracle):
    rng = np.random.default_rng(5)
    base = syn.histograms(5, rng)
    td = base[[0, 1, 2, 1, 0, 3, 4, 2]]
    sd = base[[1, 0, 2, 4]]
    assert list(oracle.match_features(sd, td)) == [1, 0, 2, 6]


def test_match_agrees_with_float64_argmin_when_unambiguous(oracle):
    rng = np.random.default_rng(6)
    sd = syn.histograms(200, rng); td = syn.histograms(300, rng)
    d = ((sd[:, None, :].astype(np.float64) - td[None].astype(np.float64)) ** 2).sum(-1)
    part = np.partition(d, 1, axis=1)
    clear = (part[:, 1] - part[:, 0]) > 1e-6
    got = oracle.match_features(sd, td)
    assert clear.sum() > 150 and np.array_equal(got[clear], d.argmin(1)[clear])


# ---------------------------------------------------------------- RANSAC semantics
def test_ransac_exact_motion_exits_at_first_valid_triple(oracle):
    rng = np.random.default_rng(7)
    tgt = rng.uniform(-0.2, 0.2, (500, 3)).astype(np.float32)
    T = syn.rigid([0.3, 0.2, 0.9], 40.0, [0.1, -0.2, 0.05])
    src = syn.apply(np.linalg.inv(T), tgt)
    corr = np.arange(500, dtype=np.uint32)
    r = oracle.ransac(src, tgt, corr, 0.001, 1000, 0.999, want_counts=True)
    counts = r.extra["counts"]
    first_valid = int(np.flatnonzero(counts != -1)[0])
    assert r.extra["best_iter"] == first_valid and r.extra["iters_run"] == first_valid + 1
    assert counts[first_valid] == 500 and r.fitness == 1.0
    assert (counts[first_valid + 1:] == -2).all()            # never executed after the break
    assert syn.rotation_error(r.transformation, T) < 1e-4


def test_ransac_degenerate_triples_consume_an_iteration(oracle):
    """n_src = 4: most triples repeat an index; they `continue` (count -1) but still use 3 draws and an id."""
    rng = np.random.default_rng(8)
    pts = rng.uniform(-1, 1, (4, 3)).astype(np.float32)
    corr = np.arange(4, dtype=np.uint32)
    r = oracle.ransac(pts, pts, corr, 0.001, 200, 2.0, want_counts=True)
    idx = oracle.uniform_indices_std(42, 4, 600).reshape(200, 3)
    degenerate = (idx[:, 0] == idx[:, 1]) | (idx[:, 1] == idx[:, 2]) | (idx[:, 0] == idx[:, 2])
    assert np.array_equal(r.extra["counts"] == -1, degenerate)
    assert (r.extra["counts"][~degenerate] == 4).all()


def test_ransac_inlier_test_is_strict_and_first_best_wins(oracle):
    # identity motion; one pair sits exactly on the threshold sphere => not an inlier (err < thr is strict)
    voxel = np.float32(0.5); thr = voxel * np.float32(1.5)      # 0.75, exactly representable
    src = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [5, 5, 5]], np.float32)
    tgt = src.copy(); tgt[4] = src[4] + np.array([thr, 0, 0], np.float32)
    corr = np.arange(5, dtype=np.uint32)
    r = oracle.ransac(src, tgt, corr, float(voxel), 300, 2.0, want_counts=True)
    c = r.extra["counts"]
    tri = oracle.uniform_indices_std(42, 5, 900).reshape(300, 3)
    clean = np.array([len(set(t)) == 3 and 4 not in t for t in tri])
    assert (c[clean] == 4).all()                                 # point 4 at distance == thr is rejected
    assert r.extra["best_iter"] == int(np.flatnonzero(c == c.max())[0])   # strict '>' => earliest maximum
    assert r.fitness == np.float32(c.max()) / np.float32(5)


def test_ransac_zero_inlier_hypotheses_never_win(oracle):
    rng = np.random.default_rng(9)
    src = rng.uniform(-1, 1, (50, 3)).astype(np.float32)
    tgt = rng.uniform(100, 200, (50, 3)).astype(np.float32)
    r = oracle.ransac(src, tgt, rng.integers(0, 50, 50).astype(np.uint32), 1e-9, 100, 0.999, want_counts=True)
    if r.extra["counts"].max() <= 0:
        assert np.array_equal(r.transformation, np.eye(4, dtype=np.float32)) and r.fitness == 0.0 and r.rmse == 0.0


# ---------------------------------------------------------------- ICP semantics
def test_icp_identical_clouds(oracle):
    rng = np.random.default_rng(10)
    pts, nrm = syn.torus(1500, rng)
    for plane in (True, False):
        r = oracle.icp(pts, pts, nrm, np.eye(4, dtype=np.float32), 0.002, 50, plane)
        assert r.extra["iters_run"] == 2 and r.fitness == 1.0
        # plane: residuals are exactly 0 so delta == I; p2p: the SVD of H returns I only to ~1e-7
        assert r.rmse == 0.0 if plane else r.rmse < 1e-7
        assert syn.rotation_error(r.transformation, np.eye(4)) < 1e-6


def test_icp_threshold_is_inclusive_and_ties_take_lowest_index(oracle):
    src = np.array([[0, 0, 0], [10, 0, 0], [20, 0, 0]], np.float32)
    tgt = np.array([[0.5, 0, 0], [-0.5, 0, 0], [10, 0.5, 0], [20.25, 0, 0], [0.5, 0, 0]], np.float32)
    r = oracle.icp(src, tgt, None, np.eye(4, dtype=np.float32), 0.5, 1, False, want_nn0=True)
    assert list(r.extra["nn_idx0"]) == [0, 2, 3]                 # 0 beats its duplicates 1 and 4 (strict '<')
    assert r.extra["ncorr"][0] == 3                              # d == thr is kept (`d > thr` skips)
    r2 = oracle.icp(src, tgt, None, np.eye(4, dtype=np.float32), np.nextafter(np.float32(0.5), np.float32(0)), 1, False)
    assert r2.extra["ncorr"][0] == 1 and r2.extra["iters_run"] == 0


def test_icp_fewer_than_three_correspondences_returns_initial(oracle):
    rng = np.random.default_rng(11)
    src = rng.uniform(-1, 1, (100, 3)).astype(np.float32); tgt = rng.uniform(5, 6, (80, 3)).astype(np.float32)
    T0 = syn.rigid([1, 0, 0], 5.0, [0.01, 0.02, 0.03]).astype(np.float32)
    r = oracle.icp(src, tgt, None, T0, 0.01, 10, True)
    assert np.array_equal(r.transformation, T0) and r.fitness == 0.0 and r.rmse == 0.0 and r.extra["iters_run"] == 0


def test_icp_point_to_plane_needs_normals(oracle):
    """point_to_plane && target.hasNormals() (registration.cpp:343): without normals the p2p branch runs."""
    c = syn.icp_case(n_model=800, n_scene=900, seed=12)
    a = oracle.icp(c.source, c.target, None, c.T_init, c.threshold, 5, True)
    b = oracle.icp(c.source, c.target, None, c.T_init, c.threshold, 5, False)
    assert np.array_equal(a.transformation, b.transformation)
    d = oracle.icp(c.source, c.target, c.target_normals, c.T_init, c.threshold, 5, True)
    assert not np.array_equal(a.transformation, d.transformation)


def test_icp_converges_to_ground_truth(oracle):
    c = syn.icp_case(n_model=4000, n_scene=3000, seed=13, noise=0.0)
    r = oracle.icp(c.source, c.target, c.target_normals, c.T_init, c.threshold, 40, True)
    assert syn.rotation_error(r.transformation, c.T_true) < 5e-3
    assert syn.translation_error(r.transformation, c.T_true) < 2e-3


def test_depth_to_cloud_oracle_reproduces_the_demo_scene_and_known_pixels(oracle):
    """pipeline.cpp:38-84 restated for any depth image: (a) on the procedural configs[0] images it must give exactly the
    demo-scene builder's cloud; (b) hand-computed pixels: z = float(d) * float(1/1000) (OpenCV scales 16u -> 32f in float),
    x = (u-cx) z / fx, mask <= 10 and z > clip dropped — 1500 * float(0.001) rounds to 1.5000001 and is therefore clipped."""
    w, h = 1280, 720
    u = np.arange(w)[None, :]; v = np.arange(h)[:, None]
    depth = np.where((np.abs(u - w / 2.0) < 100) & (np.abs(v - h / 2.0) < 100), 800, 1000).astype(np.uint16)
    mask = (((u >= w // 2 - 100) & (u <= w // 2 + 100) & (v >= h // 2 - 100) & (v <= h // 2 + 100)) * 255).astype(np.uint8)
    xyz, _ = oracle.depth_to_cloud(depth, mask, 1000.0, 1.5, 900.0, 900.0, w / 2.0, h / 2.0)
    assert np.array_equal(xyz.view(np.uint32), oracle.demo_scene_points().view(np.uint32))
    d = np.array([[500, 0, 2000], [1500, 1501, 750]], np.uint16)
    m = np.array([[255, 255, 255], [11, 255, 10]], np.uint8)
    bgr = np.arange(18, dtype=np.uint8).reshape(2, 3, 3)
    xyz, rgb = oracle.depth_to_cloud(d, m, 1000.0, 1.5, 2.0, 4.0, 1.0, 0.5, bgr=bgr)
    # kept: (v=0,u=0) z=0.5; dropped: zero depth, z=2.0 > clip, 1500 -> 1.5000001 > clip, 1501 > clip, mask == 10
    a = np.float32(1.0 / 1000.0)
    assert np.float32(1500) * a > np.float32(1.5) and np.float32(500) * a == np.float32(0.5)
    want = np.array([[(0 - 1.0) * 0.5 / 2.0, (0 - 0.5) * 0.5 / 4.0, 0.5]], np.float32)
    assert np.array_equal(xyz, want)
    assert np.allclose(rgb, np.array([[2, 1, 0]], np.float32) / 255.0, atol=0, rtol=0)
    xyz2, _ = oracle.depth_to_cloud(d, m, 1000.0, 1.6, 2.0, 4.0, 1.0, 0.5)           # a looser clip keeps the 1.5000001 m and 1.501 m pixels
    assert xyz2.shape[0] == 3 and xyz2[1, 2] == np.float32(1500) * a and xyz2[2, 2] == np.float32(1501) * a


# ------------------------------------------------------------------ pose post-processing, mask resize (pipeline.cpp:38-41, 136-137, 153-180)
def _pose(x, y, z):
    T = np.eye(4, dtype=np.float32); T[:3, 3] = (x, y, z); return T


def test_filter_duplicates_oracle_follows_the_reference_rule(oracle):
    """pipeline.cpp:153-180: first-come slots, replace by the pose nearer the origin, compare only against kept poses."""
    wps = [_pose(1, 0, 0), _pose(1.01, 0, 0), _pose(0.99, 0, 0), _pose(2, 0, 0), _pose(0.985, 0, 0)]
    out = oracle.filter_duplicates(wps, 0.02)
    assert len(out) == 2
    assert np.array_equal(out[0], wps[4]) and np.array_equal(out[1], wps[3])      # 0.985 replaced the slot of the first pose
    assert oracle.filter_duplicates([], 0.1) == []
    # strict '<': a pose exactly min_distance away is NOT a duplicate; an equally distant duplicate does not replace
    out = oracle.filter_duplicates([_pose(1, 0, 0), _pose(1.5, 0, 0), _pose(-1, 0, 0)], 0.5)
    assert len(out) == 3
    out = oracle.filter_duplicates([_pose(1, 0, 0), _pose(0, 1, 0)], 2.0)
    assert len(out) == 1 and np.array_equal(out[0], _pose(1, 0, 0))


def test_mat4_inverse_oracle_is_an_inverse_and_exact_on_easy_cases(oracle):
    """Eigen's SSE 4x4 inverse restated (2x2-block cofactors).  It must (a) invert, (b) be exact where every product is."""
    rng = np.random.default_rng(0)
    for _ in range(100):
        M = rng.standard_normal((4, 4)).astype(np.float32)
        ref = np.linalg.inv(M.astype(np.float64))
        assert np.abs(oracle.mat4_inverse(M) - ref).max() <= 2e-4 * max(1.0, np.abs(ref).max() ** 2)
    assert np.array_equal(oracle.mat4_inverse(np.eye(4, dtype=np.float32)), np.eye(4, dtype=np.float32))
    D = np.diag(np.float32([2, 4, 0.5, 8]))
    assert np.array_equal(oracle.mat4_inverse(D), np.diag(np.float32([0.5, 0.25, 2, 0.125])))
    P = np.eye(4, dtype=np.float32)[[2, 0, 3, 1]]                                    # a permutation: inverse = transpose, exactly
    assert np.array_equal(oracle.mat4_inverse(P), P.T)
    T = _pose(0.25, -0.5, 2.0); T[:3, :3] = np.float32([[0, -1, 0], [1, 0, 0], [0, 0, 1]])
    Ti = oracle.mat4_inverse(T)
    assert np.array_equal(Ti @ T, np.eye(4, dtype=np.float32))
    ext = _pose(0.5, 0.25, -1.0)
    assert np.array_equal(oracle.world_pose(T, ext), (ext @ Ti).astype(np.float32)) and np.array_equal(oracle.world_pose(T), Ti)


def test_mask_resize_oracle_equals_opencv(oracle):
    """cv::resize(..., INTER_NEAREST) restated; OpenCV's Python module is present in this image, so this one IS pinned
    against the third-party implementation the reference calls (pipeline.cpp:40)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    for sh, sw, dh, dw in ((240, 320, 720, 1280), (721, 1283, 720, 1280), (100, 100, 37, 53), (5, 7, 480, 640), (720, 1280, 720, 1280), (1, 1, 9, 4)):
        m = rng.integers(0, 256, (sh, sw)).astype(np.uint8)
        assert np.array_equal(oracle.resize_mask_nearest(m, dw, dh), cv2.resize(m, (dw, dh), interpolation=cv2.INTER_NEAREST)), (sh, sw, dh, dw)
