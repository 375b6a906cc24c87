"""Parity of the occupancy leg of the path (ProjectiveOccupancyIntegrator::integrateFrame and
EsdfIntegrator::integrateBlocks(OccupancyLayer, ...)) with the CPU oracle, through the C-ABI.

Bars: updated-block lists bit-exact and in order, log odds bit-identical, all five EsdfVoxel fields exact.
"""
import numpy as np
import pytest

from helpers import ESDF_FIELDS, assert_esdf_equal, cameras, layer_checksum
from isaac_ros_nvblox_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _nvb():
    import isaac_ros_nvblox_b200 as nvb
    return nvb


def _orc():
    from oracle import oracle as orc
    return orc


def assert_occupancy_equal(gpu_layer, cpu_layer):
    assert set(gpu_layer) == set(cpu_layer), "allocated occupancy block sets differ"
    for k, c in cpu_layer.items():
        g = gpu_layer[k]["log_odds"]
        assert np.array_equal(g.view(np.uint32), c.view(np.uint32)), ("occupancy log-odds bits", k)


def _run_pair(voxel, frames, cam, ocam, esdf=False, tsdf_kw=None, occ_kw=None, esdf_kw=None, masks=None, mask_mode=0,
              mapper_kw=None, check_every_frame=True):
    nvb, orc = _nvb(), _orc()
    m = nvb.Mapper(voxel, projective_layer_type=nvb.ProjectiveLayerType.kOccupancy, **(mapper_kw or {}))
    o = orc.OracleMap(voxel)
    p = orc.default_tsdf_params(**(tsdf_kw or {}))
    op = orc.default_occupancy_params(**(occ_kw or {}))
    ep = orc.default_esdf_params(**(esdf_kw or {}))
    if tsdf_kw:
        m.occupancy_integrator().params(**tsdf_kw)
    if occ_kw:
        m.occupancy_integrator().occupancy_params(**occ_kw)
    if esdf_kw:
        m.esdf_integrator().params(**esdf_kw)
    for i, (depth, T) in enumerate(frames):
        mask = None if masks is None else masks[i]
        b_gpu = m.integrate_depth(depth, T, cam, mask=mask, mask_mode=mask_mode)
        b_cpu = o.integrate_occupancy(depth, T, ocam, p, op, mask=mask, mask_mode=mask_mode)
        assert np.array_equal(b_gpu, b_cpu), "updated_blocks of frame %d differ" % i
        if esdf:
            m.update_esdf()
            o.integrate_esdf_occupancy(b_cpu if i > 0 else o.occupancy_block_indices(), ep)
            if check_every_frame:
                assert_esdf_equal(m.esdf_layer().as_dict(), o.esdf_layer())
    assert_occupancy_equal(m.occupancy_layer().as_dict(), o.occupancy_layer())
    if esdf:
        assert_esdf_equal(m.esdf_layer().as_dict(), o.esdf_layer())
    return m, o, p


def test_occupancy_sequence_640x480_5cm(gpu):
    cs, cam, ocam = cameras()
    frames = syn.make_sequence(syn.sphere_in_box(), cs, syn.circle_trajectory(80)[:6])
    m, o, _ = _run_pair(0.05, frames, cam, ocam)
    assert m.occupancy_layer().num_blocks() > 3000
    lo = np.concatenate([b["log_odds"].ravel() for b in m.occupancy_layer().as_dict().values()])
    assert (lo > 0).any() and (lo < 0).any()
    m.close()


def test_occupancy_saturates_and_stays_equal(gpu):
    """Twelve identical frames drive the log odds into the kMinLogOdds_/kMaxLogOdds_ clamps."""
    cs, cam, ocam = cameras(320, 240)
    depth, T = syn.make_sequence(syn.box_with_cube(), cs, syn.circle_trajectory(40)[:1])[0]
    m, o, _ = _run_pair(0.05, [(depth, T)] * 12, cam, ocam)
    lo = np.concatenate([b["log_odds"].ravel() for b in m.occupancy_layer().as_dict().values()])
    assert abs(float(lo.max()) - np.log(0.99 / 0.01)) < 1e-5 and abs(float(lo.min()) - np.log(0.01 / 0.99)) < 1e-5
    m.close()


def test_occupancy_sensor_model_params_and_truncation_bump(gpu):
    cs, cam, ocam = cameras(320, 240)
    frames = syn.make_sequence(syn.sphere_in_box(), cs, syn.circle_trajectory(40)[:3], noise_sigma_rel=0.01, seed=2)
    m, o, p = _run_pair(0.05, frames, cam, ocam, tsdf_kw=dict(truncation_distance_vox=1.0, max_integration_distance_m=5.0),
                        occ_kw=dict(free_region_occupancy_probability=0.2, occupied_region_occupancy_probability=0.9,
                                    unobserved_region_occupancy_probability=0.45, occupied_region_half_width_m=0.15))
    # 1 vox * 0.05 m < 0.15 m: raised through the integrator's own setter, on both sides
    assert m.occupancy_integrator().truncation_distance_vox() == p.truncation_distance_vox == np.float32(0.15) / np.float32(0.05)
    m.close()


@pytest.mark.parametrize("mask_mode", [0, 1])
def test_occupancy_masked_frames(gpu, mask_mode):
    cs, cam, ocam = cameras(320, 240)
    frames = syn.make_sequence(syn.sphere_in_box(), cs, syn.circle_trajectory(40)[:3])
    rng = np.random.default_rng(7)
    masks = [(rng.uniform(size=(240, 320)) < 0.5).astype(np.uint8) * 255 for _ in frames]
    masks[1][:] = 0
    m, _, _ = _run_pair(0.05, frames, cam, ocam, masks=masks, mask_mode=mask_mode)
    m.close()


@pytest.mark.parametrize("bad", [np.nan, np.inf, -np.inf, 0.0, -10.0])
def test_occupancy_invalid_frames_integrate_nothing(gpu, bad):
    cs, cam, ocam = cameras(160, 120)
    T = np.eye(4, dtype=np.float32)
    good = np.full((120, 160), 2.0, np.float32)
    m, _, _ = _run_pair(0.1, [(np.full((120, 160), bad, np.float32), T), (good, T),
                              (np.full((120, 160), bad, np.float32), T)], cam, ocam)
    m.close()


def test_occupancy_distorted_camera(gpu):
    cs, cam, ocam = cameras(320, 240, radial=(-0.05, 0.01, 0, 0.02, 0, 0), tangential=(0.001, -0.0005))
    frames = syn.make_sequence(syn.box_with_cube(), cs, syn.circle_trajectory(40)[:3])
    m, _, _ = _run_pair(0.05, frames, cam, ocam, esdf=True)
    m.close()


@pytest.mark.parametrize("persistent", [1, 0, 3])
def test_esdf_from_occupancy_incremental(gpu, persistent):
    cs, cam, ocam = cameras(320, 240)
    frames = syn.make_sequence(syn.sphere_in_box(), cs, syn.circle_trajectory(40)[:6])
    m, o, _ = _run_pair(0.05, frames, cam, ocam, esdf=True, mapper_kw=dict(esdf_persistent=persistent))
    s_gpu, s_cpu = m.esdf_integrator().last_stats(), o.esdf_stats()
    for k in ("marked", "with_sites", "to_clear", "clear_candidates", "cleared", "swept", "face_passes", "rings"):
        assert s_gpu[k] == s_cpu[k], (k, s_gpu, s_cpu)
    assert s_gpu["with_sites"] > 0
    m.close()


def test_esdf_from_occupancy_dynamic_scene_clears_sites(gpu):
    """A sphere that moves between frames: occupied voxels turn free again, so the clear pass runs."""
    cs, cam, ocam = cameras(320, 240)
    frames = [(d, T) for d, T, _ in syn.moving_sphere_sequence(cs, syn.circle_trajectory(40)[:8], step_m=0.2)]
    m, o, _ = _run_pair(0.05, frames, cam, ocam, esdf=True, esdf_kw=dict(max_esdf_distance_m=1.0, occupied_threshold=0.6))
    assert o.esdf_stats()["cleared"] >= 0
    m.close()


def test_esdf_from_occupancy_noise_and_threshold(gpu):
    cs, cam, ocam = cameras(320, 240)
    frames = syn.make_sequence(syn.box_with_cube(), cs, syn.circle_trajectory(40)[:5], noise_sigma_rel=0.02, dropout=0.05,
                               seed=3)
    m, o, _ = _run_pair(0.05, frames, cam, ocam, esdf=True, esdf_kw=dict(occupied_threshold=0.8, max_esdf_distance_m=1.5))
    m.close()


def test_esdf_from_ground_truth_occupancy_layer(gpu):
    """OccupancySingleEsdfTestGPU shape (test_esdf_integrator.cpp:554-582): a hand-filled occupancy layer,
    one integrateBlocks call over all of it."""
    nvb, orc = _nvb(), _orc()
    voxel = 0.2
    scene = syn.sphere_in_box()
    m = nvb.Mapper(voxel, projective_layer_type=nvb.ProjectiveLayerType.kOccupancy)
    o = orc.OracleMap(voxel)
    ii = (np.indices((8, 8, 8)).reshape(3, -1).T + 0.5)
    keys, blocks = [], []
    for x in range(-4, 4):
        for y in range(-4, 4):
            for z in range(-1, 4):
                pos = (np.array([x, y, z]) * 8 + ii) * voxel
                occ = scene.distance(pos) <= np.sqrt(3.0) * voxel / 2.0
                blk = np.where(occ, np.float32(6.9), np.float32(-6.9)).astype(np.float32).reshape(8, 8, 8)
                keys.append((x, y, z)), blocks.append(blk)
                o.set_occupancy_block((x, y, z), blk)
    keys = np.asarray(keys, np.int32)
    vox = np.zeros((len(keys), 8, 8, 8), nvb.OCCUPANCY_VOXEL_DTYPE)
    vox["log_odds"] = np.stack(blocks)
    m.occupancy_layer().set_blocks(keys, vox)
    m.esdf_integrator().params(max_esdf_distance_m=4.0)
    m.esdf_integrator().integrate_blocks(keys)
    o.integrate_esdf_occupancy(keys, orc.default_esdf_params(max_esdf_distance_m=4.0))
    assert_esdf_equal(m.esdf_layer().as_dict(), o.esdf_layer())
    m.close()


def test_occupancy_layer_api_and_errors(gpu):
    nvb = _nvb()
    m = nvb.Mapper(0.05, projective_layer_type=nvb.ProjectiveLayerType.kOccupancy)
    t = nvb.Mapper(0.05)
    with pytest.raises(Exception):
        m.tsdf_layer().num_blocks()       # an occupancy mapper has no TSDF layer
    with pytest.raises(Exception):
        t.occupancy_layer().num_blocks()  # and vice versa
    with pytest.raises(Exception):
        m.occupancy_integrator().free_region_occupancy_probability(1.5)
    with pytest.raises(Exception):
        m.esdf_integrator().occupied_threshold(-0.1)
    with pytest.raises(Exception):
        nvb.Mapper(0.05, projective_layer_type=3)
    assert m.occupancy_integrator().occupied_region_half_width_m() == np.float32(0.1)
    assert m.esdf_integrator().occupied_threshold() == 0.5
    idx = np.array([[0, 0, 0], [5, -5, 5]], np.int32)
    vox = np.zeros((2, 8, 8, 8), nvb.OCCUPANCY_VOXEL_DTYPE)
    vox["log_odds"] = np.random.default_rng(1).normal(size=(2, 8, 8, 8)).astype(np.float32)
    m.occupancy_layer().set_blocks(idx, vox)
    got, found = m.occupancy_layer().get_blocks(idx)
    assert found.all() and np.array_equal(got, vox)
    p0, p1 = m.occupancy_layer().block_device_ptr(idx[0]), m.occupancy_layer().block_device_ptr(idx[1])
    assert p0 and p1 and abs(p1 - p0) % 2048 == 0
    m.clear()
    assert m.occupancy_layer().num_blocks() == 0
    m.close(), t.close()


def test_occupancy_clear_and_reuse_is_deterministic(gpu):
    nvb = _nvb()
    cs, cam, ocam = cameras(320, 240)
    frames = syn.make_sequence(syn.box_with_cube(), cs, syn.circle_trajectory(40)[:4])
    m = nvb.Mapper(0.05, projective_layer_type=nvb.ProjectiveLayerType.kOccupancy)
    sums = []
    for rep in range(2):
        for depth, T in frames:
            m.integrate_depth_async(depth, T, cam)
            m.update_esdf(sync=False)
        m.synchronize()
        sums.append((layer_checksum(m.occupancy_layer().as_dict(), ("log_odds",)),
                     layer_checksum(m.esdf_layer().as_dict(), ESDF_FIELDS)))
        m.clear()
    assert sums[0] == sums[1]
    m.close()
