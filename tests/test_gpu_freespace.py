"""Parity of the freespace leg (FreespaceIntegrator::updateFreespaceLayer, Mapper::updateFreespace, and the ESDF
integrator's freespace overload) with the CPU oracle, through the C-ABI. All three FreespaceVoxel fields exact."""
import numpy as np
import pytest

from helpers import assert_esdf_equal, assert_tsdf_equal, cameras
from isaac_ros_nvblox_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
FIELDS = ("last_occupied_timestamp_ms", "consecutive_occupancy_duration_ms", "is_high_confidence_freespace")


def _nvb():
    import isaac_ros_nvblox_b200 as nvb
    return nvb


def _orc():
    from oracle import oracle as orc
    return orc


def assert_freespace_equal(g, c):
    assert set(g) == set(c), "allocated freespace block sets differ"
    for k in c:
        for f in FIELDS:
            assert np.array_equal(g[k][f], c[k][f]), (f, k)


def _mapper(voxel=0.05, **kw):
    nvb = _nvb()
    return nvb.Mapper(voxel, projective_layer_type=nvb.ProjectiveLayerType.kTsdfWithFreespace, **kw)


def test_freespace_mapper_sequence_with_view_exclusion_and_esdf(gpu):
    """Mapper::integrateDepth + updateFreespace (tracker, view exclusion with 2 x truncation) + updateEsdf with the freespace
    overload, over a dynamic scene so that voxels go free -> occupied -> free."""
    orc = _orc()
    cs, cam, ocam = cameras(320, 240)
    frames = [(d, T) for d, T, _ in syn.moving_sphere_sequence(cs, syn.circle_trajectory(40)[:10], step_m=0.25)]
    m, o = _mapper(), orc.OracleMap(0.05)
    kw = dict(max_unobserved_to_keep_consecutive_occupancy_ms=250, min_duration_since_occupied_for_freespace_ms=300,
              min_consecutive_occupancy_duration_for_reset_ms=400)
    m.freespace_integrator().params(**kw)
    fp_ = orc.default_freespace_params(**kw)
    n_hc = 0
    for i, (d, T) in enumerate(frames):
        t_ms = 1000 + 100 * i
        b = m.integrate_depth(d, T, cam)
        o.integrate_depth(d, T, ocam)
        m.update_freespace(t_ms, depth=d, T_L_C=T, camera=cam)
        o.update_freespace(o.tsdf_block_indices() if i == 0 else b, t_ms, fp_, depth=d, T_L_C=T, cam=ocam,
                           max_view_distance_m=7.0, truncation_distance_m=2 * 4 * 0.05)
        assert_freespace_equal(m.freespace_layer().as_dict(), o.freespace_layer())
        m.update_esdf()
        o.integrate_esdf_with_freespace(o.tsdf_block_indices() if i == 0 else b)
        assert_esdf_equal(m.esdf_layer().as_dict(), o.esdf_layer())
    n_hc = sum(int(v["is_high_confidence_freespace"].sum()) for v in m.freespace_layer().as_dict().values())
    assert n_hc > 1000
    assert_tsdf_equal(m.tsdf_layer().as_dict(), o.tsdf_layer())
    m.close()


def test_freespace_plane_state_machine_on_block_lists(gpu):
    """FreespacePlane (test_freespace_integrator.cpp:107-290) through FreespaceIntegrator::updateFreespaceLayer on explicit
    lists, no view exclusion, check_neighborhood off and on."""
    orc = _orc()
    cs, cam, ocam = cameras()
    T = np.eye(4, dtype=np.float32)
    voxel, step = 0.1, 100
    for check_nb in (0, 1):
        m, o = _mapper(voxel), orc.OracleMap(voxel)
        m.tsdf_integrator().params(truncation_distance_vox=4.0, max_integration_distance_m=3.2)
        tp = orc.default_tsdf_params(truncation_distance_vox=4.0, max_integration_distance_m=3.2)
        kw = dict(max_tsdf_distance_for_occupancy_m=0.3, max_unobserved_to_keep_consecutive_occupancy_ms=2 * step,
                  min_duration_since_occupied_for_freespace_ms=5 * step, min_consecutive_occupancy_duration_for_reset_ms=10 * step,
                  check_neighborhood=check_nb)
        m.freespace_integrator().params(**kw)
        fp_ = orc.default_freespace_params(**kw)
        depth = syn.render_depth(syn.plane_scene(4.0), cs, np.eye(4), max_dist=8.0)
        b = m.integrate_depth(depth, T, cam)
        o.integrate_depth(depth, T, ocam, tp)
        times = [42, 142, 342, 542]
        for t in times:
            m.freespace_integrator().update_freespace_layer(b, t)
            o.update_freespace(b, t, fp_)
            assert_freespace_equal(m.freespace_layer().as_dict(), o.freespace_layer())
        depth2 = syn.render_depth(syn.plane_scene(3.2), cs, np.eye(4), max_dist=8.0)
        b = m.integrate_depth(depth2, T, cam)
        o.integrate_depth(depth2, T, ocam, tp)
        for t in (1042, 1242, 1442, 1642, 1842, 2042):
            dup = np.vstack([b, b[:7]])  # duplicates in the caller's list
            m.freespace_integrator().update_freespace_layer(dup, t)
            o.update_freespace(b, t, fp_)
            assert_freespace_equal(m.freespace_layer().as_dict(), o.freespace_layer())
        m.close()


def test_freespace_initialize_high_confidence_and_full_layer(gpu):
    orc = _orc()
    cs, cam, ocam = cameras(320, 240)
    frames = syn.make_sequence(syn.sphere_in_box(), cs, syn.circle_trajectory(40)[:3])
    m, o = _mapper(), orc.OracleMap(0.05)
    m.freespace_integrator().params(initialize_to_high_confidence_freespace=1)
    fp_ = orc.default_freespace_params(initialize_to_high_confidence_freespace=1)
    for d, T in frames:
        m.integrate_depth(d, T, cam)
        o.integrate_depth(d, T, ocam)
    m.update_freespace(500)  # first query of the tracker: all blocks, no view
    o.update_freespace(o.tsdf_block_indices(), 500, fp_)
    assert_freespace_equal(m.freespace_layer().as_dict(), o.freespace_layer())
    m.update_freespace(900)  # nothing new: empty list
    m.update_freespace(1300, update_full_layer=True)
    o.update_freespace(o.tsdf_block_indices(), 1300, fp_)
    assert_freespace_equal(m.freespace_layer().as_dict(), o.freespace_layer())
    m.update_esdf()
    o.integrate_esdf_with_freespace(o.tsdf_block_indices())
    assert_esdf_equal(m.esdf_layer().as_dict(), o.esdf_layer())
    assert sum(int(v["is_site"].sum()) for v in m.esdf_layer().as_dict().values()) == 0  # every voxel is high-confidence free
    m.close()


def test_freespace_with_decay_deallocation(gpu):
    """Decay on a TSDF-with-freespace mapper: deallocated blocks leave the freespace layer too (clearBlocksInLayers)."""
    orc = _orc()
    cs, cam, ocam = cameras(320, 240)
    frames = syn.make_sequence(syn.box_with_cube(), cs, syn.circle_trajectory(40)[:4])
    m, o = _mapper(), orc.OracleMap(0.05)
    fp_ = orc.default_freespace_params()
    for i, (d, T) in enumerate(frames[:3]):
        b = m.integrate_depth(d, T, cam)
        o.integrate_depth(d, T, ocam)
        m.update_freespace(100 * (i + 1))
        o.update_freespace(o.tsdf_block_indices() if i == 0 else b, 100 * (i + 1), fp_)
    m.tsdf_decay_integrator().params(decay_factor=0.05)
    dp = orc.default_tsdf_decay_params(decay_factor=0.05)
    d, T = frames[2]
    for _ in range(4):
        r_gpu = m.decay(depth=d, T_L_C=T, camera=cam)
        r_cpu = o.decay_tsdf(dp, depth=d, T_L_C=T, cam=ocam, max_view_distance_m=7.0, truncation_distance_m=0.2)
        assert set(map(tuple, r_gpu.tolist())) == set(map(tuple, r_cpu.tolist()))
    assert_freespace_equal(m.freespace_layer().as_dict(), o.freespace_layer())
    d, T = frames[3]
    b = m.integrate_depth(d, T, cam)
    o.integrate_depth(d, T, ocam)
    m.update_freespace(2000, depth=d, T_L_C=T, camera=cam)  # tracker was reset by the decay: all blocks
    o.update_freespace(o.tsdf_block_indices(), 2000, fp_, depth=d, T_L_C=T, cam=ocam, max_view_distance_m=7.0,
                       truncation_distance_m=0.4)
    assert_freespace_equal(m.freespace_layer().as_dict(), o.freespace_layer())
    m.update_esdf()
    o.integrate_esdf_with_freespace(o.tsdf_block_indices())
    assert_esdf_equal(m.esdf_layer().as_dict(), o.esdf_layer())
    m.close()


def test_freespace_api_errors(gpu):
    nvb = _nvb()
    plain = nvb.Mapper(0.05)
    with pytest.raises(Exception):
        plain.update_freespace(100)
    with pytest.raises(Exception):
        plain.freespace_layer().num_blocks()
    fs = _mapper()
    assert fs.freespace_layer().num_blocks() == 0
    assert fs.freespace_integrator().params().min_duration_since_occupied_for_freespace_ms == 1000
    fs.update_freespace(100)  # empty map: nothing to do
    plain.close(), fs.close()
