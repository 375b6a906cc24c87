"""The oracle's colour path (ProjectiveColorIntegrator + SphereTracer) against the reference's own tests:
tests/test_color_integrator.cpp and tests/test_sphere_tracing.cpp, restated on the same scenes and constants."""
import numpy as np
import pytest

from helpers import (points_on_a_sphere, rotation_y, sphere_in_box_signed_distance, sphere_scene_tsdf_layer, spheres_distance,
                     tsdf_layer_from_distance, voxel_at_position)
from isaac_ros_nvblox_b200 import synthetic as syn
from oracle import oracle as orc

RED, GREEN, BLUE = (255, 0, 0), (0, 255, 0), (0, 0, 255)
W, H = 640, 480


def _cam(f=300.0):
    return orc.Camera(f, f, W / 2.0, H / 2.0, W, H)


def _solid(color):
    img = np.zeros((H, W, 3), np.uint8)
    img[:] = color
    return img


def _map_with_layer(voxel, idx, vox):
    m = orc.OracleMap(voxel)
    for k, v in zip(idx, vox):
        m.set_tsdf_block(k, v)
    return m


def test_half_rounding_matches_ieee_binary16():
    """blendTwoArrays rounds its weights through __half (projective_appearance_integrator.cu:287-306)."""
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-3, 3, 4000), rng.uniform(-1e-4, 1e-4, 4000), rng.uniform(-7e4, 7e4, 2000),
                        [0.0, 0.8, 0.2, 1.0 - 0.8, 6.1e-5, 5.96e-8, 2.9e-8, 65504.0, 65519.9]]).astype(np.float32)
    with np.errstate(over="ignore"):
        ref = x.astype(np.float16).astype(np.float32)
    got = np.array([orc.round_through_half(v) for v in x], np.float32)
    assert np.array_equal(got, ref)


@pytest.fixture(scope="module")
def gt_sphere_layer():
    return sphere_scene_tsdf_layer(voxel_size=0.1, truncation_m=0.4)  # ColorIntegrationTest fixture (:38-70)


def test_truncation_band_selection(gt_sphere_layer):
    """TruncationBandTest (:173-223): every updated block touches the band, and an allocated block in view that is left out
    does not (the band is |distance| < truncation on voxels with weight > 0)."""
    idx, vox = gt_sphere_layer
    m = _map_with_layer(0.1, idx, vox)
    T = syn.circle_trajectory(80)[0]
    updated = m.integrate_color(_solid(RED), T, _cam(), orc.default_color_params(truncation_distance_vox=1.0))
    assert len(updated) > 10
    layer = dict(zip(map(tuple, idx.tolist()), vox))
    for k in map(tuple, updated.tolist()):
        b = layer[k]
        assert ((b["weight"] > 0) & (np.abs(b["distance"]) < np.float32(0.1))).any()
    assert set(map(tuple, m.color_block_indices().tolist())) == set(map(tuple, updated.tolist()))


def test_integrate_color_to_ground_truth_distance_field(gt_sphere_layer):
    """IntegrateColorToGroundTruthDistanceField (:225-331): a red image from 80 poses on the circle; every observed voxel is
    red, more than half of the sphere's surface points have weight >= 1, colour blocks exist only where TSDF blocks do."""
    idx, vox = gt_sphere_layer
    m = _map_with_layer(0.1, idx, vox)
    img, cam = _solid(RED), _cam()
    touched = set()
    for T in syn.circle_trajectory(80):
        touched |= set(map(tuple, m.integrate_color(img, T, cam).tolist()))
    layer = m.color_layer()
    assert touched == set(layer)
    for blk in layer.values():
        seen = blk["weight"] > 0.0
        assert np.all(blk["color"][seen] == RED)
        assert np.all(blk["color"][~seen] == 127)  # ColorVoxel(): Gray, weight 0
    pts = points_on_a_sphere(2.0, (0.0, 0.0, 2.0))
    vs = [voxel_at_position(layer, p, 0.1) for p in pts]
    assert all(v is not None for v in vs)
    assert np.mean([v["weight"] >= 1.0 for v in vs]) > 0.5
    tsdf_blocks = set(map(tuple, idx.tolist()))
    assert set(layer) <= tsdf_blocks
    assert max(float(b["weight"].max()) for b in layer.values()) == 5.0  # max_weight


def _check_sphere_color(layer, center, color, voxel=0.1):
    """checkSphereColor (:109-141) -> ratio of surface points observed (weight >= 1e-3), all of which have `color`."""
    n_obs, pts = 0, points_on_a_sphere(2.0, center)
    for p in pts:
        v = voxel_at_position(layer, p, voxel)
        assert v is not None
        if v["weight"] >= 1e-3:
            assert tuple(v["color"]) == color
            n_obs += 1
    return n_obs / len(pts)


def test_colored_spheres():
    """ColoredSpheres (:333-440): three spheres, three cameras looking along +x, three solid images."""
    centers = [(5.0, 0.0, 0.0), (5.0, 5.0, 0.0), (5.0, 10.0, 0.0)]
    idx, vox = tsdf_layer_from_distance(spheres_distance(centers, 2.0), (-5.0, -5.0, -5.0), (10.0, 15.0, 5.0), 0.1, 0.2)
    m = _map_with_layer(0.1, idx, vox)
    cam = _cam(450.0)
    for y, color in zip((0.0, 5.0, 10.0), (RED, GREEN, BLUE)):
        T = rotation_y(np.pi / 2)
        T[:3, 3] = (0.0, y, 0.0)
        m.integrate_color(_solid(color), T, cam)
    layer = m.color_layer()
    for c, color in zip(centers, (RED, GREEN, BLUE)):
        assert _check_sphere_color(layer, c, color) > 0.2


def test_occlusion():
    """OcclusionTesting (:442-508): the sphere behind the first one gets no colour."""
    c1, c2 = (5.0, 0.0, 0.0), (10.0, 0.0, 0.0)
    idx, vox = tsdf_layer_from_distance(spheres_distance([c1, c2], 2.0), (-5.0, -5.0, -5.0), (15.0, 15.0, 5.0), 0.1, 0.4)
    m = _map_with_layer(0.1, idx, vox)
    m.integrate_color(_solid(RED), rotation_y(np.pi / 2), _cam())
    layer = m.color_layer()
    assert _check_sphere_color(layer, c1, RED) > 0.2
    for p in points_on_a_sphere(2.0, c2):
        v = voxel_at_position(layer, p, 0.1)
        if v is not None:
            assert v["weight"] == 0.0


def test_measurement_weight():
    """WeightingFunction (:510-570): a plane 5 m ahead, measurement weight 0.3: every observed voxel has weight 0.3."""
    idx, vox = tsdf_layer_from_distance(lambda P: 5.0 - P[..., 2], (-10.0, -10.0, -10.0), (10.0, 10.0, 10.0), 0.1, 5.0)
    m = _map_with_layer(0.1, idx, vox)
    updated = m.integrate_color(_solid(RED), np.eye(4, dtype=np.float32), _cam(),
                                orc.default_color_params(measurement_weight=0.3))
    assert len(updated) > 0 and len(m.color_block_indices()) > 0
    n = 0
    for blk in m.color_layer().values():
        w = blk["weight"][blk["weight"] > 1e-4]
        assert np.all(np.abs(w - 0.3) < 1e-4)
        n += w.size
    assert n > 0


def test_blend_is_the_fp16_weighted_sum():
    """UpdateAppearanceVoxelFunctor (:312-348): first observation copies the colour, later ones blend with weights rounded to
    binary16: round(old * h(0.2) + new * h(0.8)); the weight accumulates up to max_weight."""
    idx, vox = tsdf_layer_from_distance(lambda P: 5.0 - P[..., 2], (-2.0, -2.0, 4.0), (2.0, 2.0, 6.0), 0.1, 0.4)
    m = _map_with_layer(0.1, idx, vox)
    T, cam = np.eye(4, dtype=np.float32), _cam()
    c1, c2 = (200, 10, 77), (13, 250, 78)
    m.integrate_color(_solid(c1), T, cam)
    m.integrate_color(_solid(c2), T, cam)
    h_old = np.float32(np.float16(np.float32(1.0) - np.float32(0.8)))
    h_new = np.float32(np.float16(np.float32(0.8)))
    exp = tuple(int(np.floor(np.float32(a) * h_old + np.float32(b) * h_new + np.float32(0.5))) for a, b in zip(c1, c2))
    seen = 0
    for blk in m.color_layer().values():
        two = np.abs(blk["weight"] - np.float32(1.6)) < 1e-6
        assert np.all(blk["color"][two] == exp)
        seen += int(two.sum())
    assert seen > 1000


def test_mask_and_interpolation():
    """The measured colour is the bilinear interpolation of the 4 neighbours rounded per channel (interpolation_2d_impl.h:26-48,
    152-199); masked-out pixels leave the voxel untouched."""
    idx, vox = tsdf_layer_from_distance(lambda P: 5.0 - P[..., 2], (-3.0, -3.0, 4.0), (3.0, 3.0, 6.0), 0.1, 0.4)
    cam, T = _cam(), np.eye(4, dtype=np.float32)
    img = np.zeros((H, W, 3), np.uint8)
    img[..., 0] = (np.arange(W) % 256)[None, :]
    img[..., 1] = (np.arange(H) % 256)[:, None]
    img[..., 2] = 50
    m = _map_with_layer(0.1, idx, vox)
    m.integrate_color(img, T, cam)
    n = 0
    for (bx, by, bz), blk in m.color_layer().items():
        for vx, vy, vz in np.argwhere(blk["weight"] > 0)[::7]:
            p = (np.array([bx, by, bz], np.float32) * np.float32(0.8) + np.array([vx, vy, vz], np.float32) * np.float32(0.1)
                 + np.float32(0.05))
            u = p[0] / p[2] * np.float32(300.0) + np.float32(320.0)
            v = p[1] / p[2] * np.float32(300.0) + np.float32(240.0)
            # a horizontal ramp interpolates to u - 0.5 (away from the 255 -> 0 wrap), a vertical one to v - 0.5
            if 2 < (u - 0.5) % 256 < 253:
                assert abs(int(blk["color"][vx, vy, vz][0]) - float((u - 0.5) % 256)) <= 0.5 + 1e-3
            if 2 < (v - 0.5) % 256 < 253:
                assert abs(int(blk["color"][vx, vy, vz][1]) - float((v - 0.5) % 256)) <= 0.5 + 1e-3
            assert blk["color"][vx, vy, vz][2] == 50
            n += 1
    assert n > 300
    # left half masked out (mask value 0 = inactive in the non-inverted mode)
    mask = np.zeros((H, W), np.uint8)
    mask[:, W // 2:] = 1
    m2 = _map_with_layer(0.1, idx, vox)
    m2.integrate_color(img, T, cam, mask=mask)
    for (bx, by, bz), blk in m2.color_layer().items():
        x = (bx * 8 + np.arange(8) + 0.5) * 0.1
        assert np.all(blk["weight"][x < -0.1] == 0.0)
    assert sum(float(b["weight"].sum()) for b in m2.color_layer().values()) > 100


@pytest.mark.parametrize("scale", [0.9, 1.0, 1.1])
def test_sphere_tracing_plane(scale):
    """PlaneTest (test_sphere_tracing.cpp:178-276): rays towards the plane x = 0 through a (scaled) ground-truth distance field
    converge within one voxel of the true intersection."""
    voxel, half = 0.05, 5.0
    idx, vox = tsdf_layer_from_distance(lambda P: -P[..., 0], (-half,) * 3, (half,) * 3, voxel, 5.0)
    vox = vox.copy()
    vox["distance"] *= np.float32(scale)
    m = _map_with_layer(voxel, idx, vox)
    rng = np.random.default_rng(3)
    bad = 0
    for _ in range(1000):
        p_plane = np.array([0.0, rng.uniform(-4.5, 4.5), rng.uniform(-4.5, 4.5)], np.float32)
        p_vol = np.array([rng.uniform(-4.5, -1.0), rng.uniform(-4.5, 4.5), rng.uniform(-4.5, 4.5)], np.float32)
        d = p_plane - p_vol
        d = (d / np.float32(np.linalg.norm(d))).astype(np.float32)
        ok, t = m.sphere_trace_ray(p_vol, d, 4 * voxel)
        assert ok
        t_gt = -p_vol[0] / d[0]
        bad += abs(t - t_gt) > voxel
    assert bad * 100.0 / 1000 < 0.5


def test_sphere_tracing_sphere_scene_ground_truth():
    """SphereSceneTests, ground-truth field (:293-365): rendered depth images from random viewpoints: > 99.5 % of the rays
    converge, < 2 % of the pixels are off by more than 4 voxels."""
    voxel = 0.05
    scene = syn.sphere_in_box()
    idx, vox = tsdf_layer_from_distance(sphere_in_box_signed_distance, (-6.0, -6.0, -1.0), (6.0, 6.0, 6.0), voxel, 4 * voxel)
    m = _map_with_layer(voxel, idx, vox)
    cs = syn.PinholeCamera(300.0, 300.0, W / 2.0, H / 2.0, W, H)
    cam = _cam()
    rng = np.random.default_rng(5)
    for _ in range(3):
        while True:
            p = np.array([rng.uniform(-4.75, 4.75), rng.uniform(-4.75, 4.75), rng.uniform(0.25, 4.75)])
            if np.linalg.norm(p - (0, 0, 2)) > 2.0:
                break
        axis = rng.normal(size=3)
        axis /= np.linalg.norm(axis)
        ang = rng.uniform(-np.pi, np.pi)
        K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
        R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
        T = np.eye(4, dtype=np.float32)
        T[:3, :3], T[:3, 3] = R, p
        img = m.sphere_trace_image(T, cam, 4 * voxel)
        gt = syn.render_depth(scene, cs, T, max_dist=15.0)
        conv = img > 0.0
        assert conv.mean() * 100.0 > 99.5
        err = conv & (np.abs(img - gt) > 4 * voxel)
        assert err.mean() * 100.0 < 2.0


def test_sphere_tracing_subsampled_image_is_the_strided_one():
    """sphereTracingKernel (:134-173): ray (r, c) of the f-subsampled image goes through pixel coordinate f * (c, r) + f / 2."""
    voxel = 0.1
    idx, vox = sphere_scene_tsdf_layer(voxel_size=voxel, truncation_m=0.4)
    m = _map_with_layer(voxel, idx, vox)
    T, cam = syn.circle_trajectory(80)[3], _cam()
    full = m.sphere_trace_image(T, cam, 0.4, maximum_ray_length_m=7.0)
    sub = m.sphere_trace_image(T, cam, 0.4, maximum_ray_length_m=7.0, ray_subsampling_factor=4)
    assert sub.shape == (H // 4, W // 4)
    # rays through (4c + 2, 4r + 2) vs the full image's rays through (c' + 0.5, r' + 0.5): neighbours, not identical
    both = (sub > 0) & (full[2::4, 2::4] > 0)
    assert both.mean() > 0.25 and ((sub > 0) == (full[2::4, 2::4] > 0)).mean() > 0.97  # (the far walls are beyond 7 m)
    assert np.median(np.abs(sub[both] - full[2::4, 2::4][both])) < 0.05
