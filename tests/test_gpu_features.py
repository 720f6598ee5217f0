"""Parity of the stages that feed the hot path (voxelDownsample, estimateNormals, computeFPFH)
against the CPU oracle — bit-identical arrays, order included — and, at configs[0] full size, against the committed
golden digests of the demo scene (tests/golden/demo_scene.json)."""
import hashlib
import importlib
import json
import os

import numpy as np
import pytest

syn = importlib.import_module("3dvision_b200.synthetic")
reg = importlib.import_module("3dvision_b200.registration")

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "demo_scene.json")


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def same_bits(a, b):
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


# ------------------------------------------------------------------ voxelDownsample
@pytest.mark.parametrize("n,voxel,scale,seed", [(1, 0.01, 1.0, 0), (50, 10.0, 1.0, 1), (5000, 0.05, 1.0, 2), (60000, 0.004, 0.3, 3),
                                                 (20000, 0.001, 0.2, 4)])
def test_voxel_downsample_bit_identical_in_reference_order(ctx, oracle, n, voxel, scale, seed):
    rng = np.random.default_rng(seed)
    xyz = (rng.uniform(-scale, scale, (n, 3)) + rng.normal(0, 0.01 * scale, (n, 3))).astype(np.float32)
    want = oracle.voxel_downsample(xyz, voxel)
    got, _ = ctx.voxel_downsample(xyz, voxel)
    assert same_bits(got, want)


@pytest.mark.parametrize("n_vox", [1, 12, 13, 14, 29, 30, 59, 60, 127, 541, 2357, 5087, 5088, 20753, 20754, 100000])
def test_voxel_order_device_replay_equals_host_container(ctx, oracle, n_vox):
    """Voxel counts straddling libstdc++'s rehash points (13, 29, 59, ... buckets): the device replay of the container's
    insertion / rehash rules must give the order of the real std::unordered_map (the oracle runs the real container)."""
    rng = np.random.default_rng(n_vox)
    cells = rng.permutation(400000)[:n_vox]
    xyz = np.stack([cells % 100 - 50, (cells // 100) % 100 - 50, cells // 10000 - 20], 1).astype(np.float32) * np.float32(0.01) + np.float32(0.005)
    xyz = np.concatenate([xyz, xyz[rng.integers(0, n_vox, n_vox // 2)] + np.float32(0.001)])
    dev, _ = ctx.voxel_downsample(xyz, 0.01)
    host = oracle.voxel_downsample(xyz, 0.01)
    assert dev.shape[0] == n_vox and same_bits(dev, host)


def test_voxel_downsample_surface_cloud_and_colors(ctx, oracle):
    rng = np.random.default_rng(11)
    pts, _ = syn.torus(80000, rng)
    col = rng.random((80000, 3)).astype(np.float32)
    want = oracle.voxel_downsample(pts, 0.003)
    got, got_col = ctx.voxel_downsample(pts, 0.003, col)
    assert same_bits(got, want)
    # colours are averaged over the same members in the same order: check against a direct per-voxel replay
    inv = np.float32(1.0) / np.float32(0.003)
    keys = np.floor(pts * inv).astype(np.int64)
    order = {}
    for i, k in enumerate(map(tuple, keys)):
        order.setdefault(k, []).append(i)
    by_key = {}
    for k, idx in order.items():
        acc = np.zeros(3, np.float32)
        for i in idx:
            acc = acc + col[i]
        by_key[k] = acc / np.float32(len(idx))
    got_keys = np.floor(got * inv)     # a mean can round across a voxel face; match through the point means instead
    means = {}
    for k, idx in order.items():
        acc = np.zeros(3, np.float32)
        for i in idx:
            acc = acc + pts[i]
        means[(acc / np.float32(len(idx))).tobytes()] = k
    for row, c in zip(got, got_col):
        assert np.array_equal(by_key[means[row.tobytes()]], c)


def test_voxel_downsample_out_of_range_and_bad_voxel(ctx, b3d):
    xyz = np.array([[0, 0, 0], [3e6, 0, 0]], np.float32)
    with pytest.raises(b3d.B3DError):
        ctx.voxel_downsample(xyz, 1.0)          # |coordinate / voxel| >= 2^20
    with pytest.raises(b3d.B3DError):
        ctx.voxel_downsample(xyz, 0.0)
    got, _ = ctx.voxel_downsample(np.zeros((0, 3), np.float32), 0.01)
    assert got.shape == (0, 3)


# ------------------------------------------------------------------ estimateNormals
def _clouds():
    rng = np.random.default_rng(21)
    torus, _ = syn.torus(2500, rng)
    cube = rng.uniform(-0.1, 0.1, (1200, 3)).astype(np.float32)
    outliers = rng.uniform(-3.0, 3.0, (40, 3)).astype(np.float32)                   # isolated: the grid cannot settle them
    dup = np.concatenate([cube[:300], cube[:300], cube[:50]])                        # exact ties in d2 -> index order decides
    return {"torus": torus, "cube+outliers": np.concatenate([cube, outliers]), "duplicates": dup,
            "tiny": cube[:10], "one": cube[:1], "voxelised": None}


@pytest.mark.parametrize("name,k", [("torus", 30), ("cube+outliers", 30), ("duplicates", 30), ("tiny", 30), ("one", 30),
                                    ("torus", 7), ("cube+outliers", 100), ("voxelised", 30)])
def test_estimate_normals_bit_identical(ctx, oracle, name, k):
    pts = _clouds()[name]
    if pts is None:
        rng = np.random.default_rng(5)
        raw, _ = syn.torus(40000, rng)
        pts = oracle.voxel_downsample(raw, 0.01)
    want = oracle.estimate_normals(pts, k)
    got = ctx.estimate_normals(pts, k)
    assert same_bits(got, want)


def test_estimate_normals_rejects_large_k(ctx, b3d):
    with pytest.raises(b3d.B3DError):
        ctx.estimate_normals(np.zeros((10, 3), np.float32), 129)


# ------------------------------------------------------------------ computeFPFH
@pytest.mark.parametrize("n,radius,seed", [(2500, 0.03, 1), (2500, 0.08, 2), (600, 0.005, 3), (1, 0.1, 4), (3000, 10.0, 5)])
def test_compute_fpfh_bit_identical(ctx, oracle, n, radius, seed):
    """radius 0.03: ~20-40 neighbours; 0.08: well over the 100 cap; 0.005: mostly empty lists; 10.0: every point is a
    neighbour of every point (cap + grid bypass)."""
    rng = np.random.default_rng(seed)
    pts, _ = syn.torus(n, rng)
    nrm = oracle.estimate_normals(pts, 30)
    want = oracle.compute_fpfh(pts, nrm, radius)
    got = ctx.compute_fpfh(pts, nrm, radius)
    assert same_bits(got, want)


def test_compute_fpfh_duplicate_points_and_zero_normals(ctx, oracle):
    rng = np.random.default_rng(8)
    base = rng.uniform(-0.05, 0.05, (400, 3)).astype(np.float32)
    pts = np.concatenate([base, base[:100]])                                        # dist < 1e-8 pairs are skipped
    nrm = rng.normal(size=pts.shape).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm[::17] = 0.0
    want = oracle.compute_fpfh(pts, nrm, 0.02)
    got = ctx.compute_fpfh(pts, nrm, 0.02)
    assert same_bits(got, want)


# ------------------------------------------------------------------ configs[0] front end at full size vs golden digests
def test_demo_scene_front_end_matches_golden_digests(ctx, oracle):
    g = json.load(open(GOLDEN))
    voxel = g["voxel"]
    scene = oracle.demo_scene_points()            # deterministic procedural scene builder (no arithmetic under test)
    model = oracle.demo_model_points()
    R = reg.Registration
    src = R.voxelDownsample(reg.PointCloud(scene), voxel)
    tgt = R.voxelDownsample(reg.PointCloud(model), voxel)
    assert src.size() == g["n_src"] and tgt.size() == g["n_tgt"]
    assert digest(src.points) == g["sha256"]["src"] and digest(tgt.points) == g["sha256"]["tgt"]
    R.estimateNormals(src, 30); R.estimateNormals(tgt, 30)
    assert digest(tgt.normals) == g["sha256"]["tgt_normals"]
    src_f = R.computeFPFH(src, voxel * 5.0); tgt_f = R.computeFPFH(tgt, voxel * 5.0)
    assert digest(tgt_f.descriptors) == g["sha256"]["tgt_fpfh"]
    assert digest(src_f.descriptors) == g["sha256"]["src_fpfh"]
    coarse = R.ransacRegistration(src, tgt, src_f, tgt_f, voxel, g["ransac_max_iterations"], g["confidence"])
    assert np.array_equal(coarse.transformation.reshape(-1), np.asarray(g["ransac"]["T"], np.float32))
    assert np.float32(coarse.fitness) == np.float32(g["ransac"]["fitness"]) and np.float32(coarse.rmse) == np.float32(g["ransac"]["rmse"])


# ------------------------------------------------------------------ the whole per-instance body as one resident call
def test_register_scene_equals_the_five_separate_calls(b3d, ctx):
    rng = np.random.default_rng(77)
    voxel = 0.006
    model_raw = syn.rough_torus(60000, rng)
    T_true = syn.rigid([0.2, 0.9, -0.3], 25.0, [0.05, -0.03, 0.08])
    scene_raw = (syn.apply(np.linalg.inv(T_true), syn.rough_torus(60000, rng)) + rng.normal(0, 0.0003, (60000, 3))).astype(np.float32)
    R = reg.Registration
    tgt = R.voxelDownsample(reg.PointCloud(model_raw), voxel); R.estimateNormals(tgt, 30); tgt_f = R.computeFPFH(tgt, voxel * 5.0)
    src = R.voxelDownsample(reg.PointCloud(scene_raw), voxel); R.estimateNormals(src, 30); src_f = R.computeFPFH(src, voxel * 5.0)
    coarse = R.ransacRegistration(src, tgt, src_f, tgt_f, voxel, 20000, 0.999)
    fine = R.icpRefine(src, tgt, coarse.transformation, voxel * 0.4, 50, True)
    c2 = b3d.Context(0)
    try:
        assert c2.prepare_model(model_raw, voxel) == tgt.size()
        for _ in range(2):                                      # the model stays resident across scenes
            out = c2.register_scene(scene_raw, voxel, ransac_max_iterations=20000, icp_max_iterations=50)
            assert out["n_source_points"] == src.size()
            T0, f0, r0, _ = out["coarse"]; T1, f1, r1, _ = out["refined"]
            assert np.array_equal(T0, coarse.transformation) and f0 == coarse.fitness and r0 == coarse.rmse
            assert np.array_equal(T1, fine.transformation) and f1 == fine.fitness and r1 == fine.rmse
        assert syn.rotation_error(T1, T_true) < 2e-3 and syn.translation_error(T1, T_true) < 1e-3      # and it registered
        import torch                                            # device-resident input: same result, no host hop
        d_scene = torch.from_numpy(scene_raw).cuda()
        torch.cuda.synchronize()
        dev = c2.register_scene_device(d_scene.data_ptr(), scene_raw.shape[0], voxel, ransac_max_iterations=20000, icp_max_iterations=50)
        assert np.array_equal(dev["refined"][0], T1) and dev["refined"][1:3] == (f1, r1)
        empty = c2.register_scene(np.zeros((0, 3), np.float32), voxel)
        assert np.array_equal(empty["refined"][0], np.eye(4, dtype=np.float32)) and empty["refined"][1] == 0.0
    finally:
        c2.close()


def test_register_scene_requires_a_model(b3d):
    c2 = b3d.Context(0)
    try:
        with pytest.raises(b3d.B3DError):
            c2.register_scene(np.zeros((10, 3), np.float32), 0.01)
    finally:
        c2.close()


# ------------------------------------------------------------------ depth image -> cloud (row f-4)
def _depth_case(seed, h=240, w=320):
    rng = np.random.default_rng(seed)
    depth = rng.integers(0, 2600, (h, w)).astype(np.uint16)
    depth[rng.random((h, w)) < 0.2] = 0
    mask = rng.choice(np.array([0, 5, 10, 11, 128, 255], np.uint8), (h, w))
    bgr = rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
    return depth, mask, bgr


@pytest.mark.parametrize("use_mask,use_bgr", [(True, True), (False, False), (True, False)])
def test_depth_to_cloud_bit_identical_in_raster_order(ctx, oracle, use_mask, use_bgr):
    depth, mask, bgr = _depth_case(3)
    args = (1000.0, 1.5, 611.5, 609.25, 159.5, 121.25)
    want_xyz, want_rgb = oracle.depth_to_cloud(depth, mask if use_mask else None, *args, bgr=bgr if use_bgr else None)
    got_xyz, got_rgb = ctx.depth_to_cloud(depth, mask if use_mask else None, *args, bgr=bgr if use_bgr else None)
    assert same_bits(got_xyz, want_xyz) and want_xyz.shape[0] > 1000
    assert (got_rgb is None) == (not use_bgr)
    if use_bgr:
        assert same_bits(got_rgb, want_rgb)


def test_depth_to_cloud_reproduces_the_demo_scene(ctx, oracle):
    """configs[0]: floor at 1.0 m, 200 x 200 px box at 0.8 m, rectangular mask (pipeline.cpp:211-257) -> the golden's 40401 points."""
    w, h = 1280, 720
    u = np.arange(w)[None, :]; v = np.arange(h)[:, None]
    depth = np.where((np.abs(u - w / 2.0) < 100) & (np.abs(v - h / 2.0) < 100), 800, 1000).astype(np.uint16)
    mask = (((u >= w // 2 - 100) & (u <= w // 2 + 100) & (v >= h // 2 - 100) & (v <= h // 2 + 100)) * 255).astype(np.uint8)
    got, _ = ctx.depth_to_cloud(depth, mask, 1000.0, 1.5, 900.0, 900.0, w / 2.0, h / 2.0)
    want = oracle.demo_scene_points()
    assert got.shape[0] == json.load(open(GOLDEN))["n_raw_points"] and same_bits(got, want)


def test_depth_to_cloud_empty_and_bad_arguments(ctx, b3d):
    z = np.zeros((8, 8), np.uint16)
    xyz, _ = ctx.depth_to_cloud(z, None, 1000.0, 1.5, 100.0, 100.0, 4.0, 4.0)
    assert xyz.shape == (0, 3)
    with pytest.raises(b3d.B3DError):
        ctx.depth_to_cloud(z, None, 0.0, 1.5, 100.0, 100.0, 4.0, 4.0)


def test_register_depth_equals_deprojection_then_register_scene(b3d):
    rng = np.random.default_rng(91)
    voxel = 0.004
    h, w, f = 300, 400, 500.0
    uu, vv = np.meshgrid(np.arange(w), np.arange(h))
    zz = 0.6 + 0.08 * np.sin(uu / 23.0) * np.cos(vv / 17.0) + 0.05 * np.cos((uu + vv) / 31.0) + 0.03 * np.sin(uu / 7.0 + 1.0) * np.sin(vv / 9.0)
    depth = np.round(zz * 1000.0).astype(np.uint16)
    mask = np.full((h, w), 255, np.uint8); mask[:20] = 0
    c2 = b3d.Context(0)
    try:
        cloud, _ = c2.depth_to_cloud(depth, mask, 1000.0, 1.5, f, f, w / 2.0, h / 2.0)
        T = syn.rigid([0.1, 0.3, 0.9], 12.0, [0.02, -0.01, 0.03])
        model = syn.apply(T, cloud[rng.permutation(cloud.shape[0])[: cloud.shape[0] * 3 // 4]])
        assert c2.prepare_model(model, voxel) > 1000
        a = c2.register_scene(cloud, voxel, ransac_max_iterations=5000, icp_max_iterations=30)
        b = c2.register_depth(depth, mask, 1000.0, 1.5, f, f, w / 2.0, h / 2.0, voxel, ransac_max_iterations=5000, icp_max_iterations=30)
        assert a["n_source_points"] == b["n_source_points"] > 1000
        assert np.array_equal(a["coarse"][0], b["coarse"][0]) and np.array_equal(a["refined"][0], b["refined"][0])
        assert a["refined"][1:] == b["refined"][1:]
    finally:
        c2.close()


# ------------------------------------------------------------------ rest of row f-4: mask resize, world pose, duplicate filter
@pytest.mark.parametrize("mh,mw", [(60, 80), (181, 321), (360, 640), (7, 5)])
def test_depth_to_cloud_resizes_the_mask_like_the_reference(ctx, oracle, mh, mw):
    """A mask smaller / larger than the depth image goes through cv::resize(..., INTER_NEAREST) first (pipeline.cpp:39-41);
    the device reads the mask through the same source-pixel map instead of materialising the resized image."""
    rng = np.random.default_rng(mh * 1000 + mw)
    h, w = 180, 320
    depth = rng.integers(300, 1400, (h, w)).astype(np.uint16)
    mask = (rng.random((mh, mw)) < 0.6).astype(np.uint8) * 255
    args = (1000.0, 1.5, 300.0, 310.0, w / 2.0, h / 2.0)
    want, _ = oracle.depth_to_cloud(depth, oracle.resize_mask_nearest(mask, w, h), *args)
    got, _ = ctx.depth_to_cloud(depth, mask, *args)
    assert got.shape[0] > 1000 and np.array_equal(got, want)


def test_world_poses_bit_identical(ctx, oracle):
    """pipeline.cpp:136-137 on the device == the oracle's restatement of Eigen's SSE inverse + packet product."""
    rng = np.random.default_rng(9)
    Ts = np.stack([syn.rigid(rng.standard_normal(3), float(rng.uniform(-170, 170)), rng.uniform(-1, 1, 3)).astype(np.float32) for _ in range(70)])
    Ts[5] = rng.standard_normal((4, 4)).astype(np.float32)                        # not rigid: the general 4x4 path
    ext = syn.rigid([0.1, 0.2, 0.9], 33.0, [0.4, -0.2, 1.1]).astype(np.float32)
    got = ctx.world_poses(Ts, ext)
    for T, g in zip(Ts, got):
        assert np.array_equal(g, oracle.world_pose(T, ext))
    got = ctx.world_poses(Ts)
    for T, g in zip(Ts, got):
        assert np.array_equal(g, oracle.world_pose(T))
    assert ctx.world_poses(np.zeros((0, 4, 4), np.float32)).shape == (0, 4, 4)


def test_filter_duplicates_bit_identical(ctx, oracle, b3d):
    rng = np.random.default_rng(10)
    centres = rng.uniform(-0.5, 0.5, (12, 3))
    wps = []
    for i in range(64):                                                           # clusters of near-duplicate detections
        T = np.eye(4, dtype=np.float32); T[:3, 3] = centres[rng.integers(0, 12)] + rng.normal(0, 0.004, 3)
        wps.append(T)
    for min_d in (0.0, 0.005, 0.02, 0.3, 10.0):
        want = oracle.filter_duplicates(wps, min_d)
        got = ctx.filter_duplicates(wps, min_d)
        assert len(got) == len(want) and all(np.array_equal(a, b) for a, b in zip(got, want)), min_d
    assert ctx.filter_duplicates([], 0.1) == []
    pipe = importlib.import_module("3dvision_b200.pipeline")                      # the orchestrator mirror goes through the same entry
    assert all(np.array_equal(a, b) for a, b in zip(pipe.filter_duplicates(wps, 0.02), oracle.filter_duplicates(wps, 0.02)))
    assert np.array_equal(pipe.world_pose(wps[3]), oracle.world_pose(wps[3]))
