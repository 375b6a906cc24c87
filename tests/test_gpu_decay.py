"""Parity of the decay integrators (TsdfDecayIntegrator / OccupancyDecayIntegrator through Mapper::decayTsdf /
decayOccupancy) and of block deallocation with the CPU oracle, through the C-ABI.

Bars: deallocated block sets equal, voxel values bit-identical, ESDF fields exact after the following updates."""
import numpy as np
import pytest

from helpers import assert_esdf_equal, assert_tsdf_equal, cameras
from isaac_ros_nvblox_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _nvb():
    import isaac_ros_nvblox_b200 as nvb
    return nvb


def _orc():
    from oracle import oracle as orc
    return orc


def _as_set(a):
    return set(map(tuple, np.asarray(a).reshape(-1, 3).tolist()))


def _assert_occ_equal(g, c):
    assert set(g) == set(c), "allocated occupancy block sets differ"
    for k in c:
        assert np.array_equal(g[k]["log_odds"].view(np.uint32), c[k].view(np.uint32)), k


def _build_pair(frames, cam, ocam, voxel=0.05, occupancy=False, esdf=True):
    nvb, orc = _nvb(), _orc()
    m = nvb.Mapper(voxel, projective_layer_type=nvb.ProjectiveLayerType.kOccupancy if occupancy else nvb.ProjectiveLayerType.kTsdf)
    o = orc.OracleMap(voxel)
    tp = orc.default_tsdf_params()
    for i, (d, T) in enumerate(frames):
        b = m.integrate_depth(d, T, cam)
        if occupancy:
            o.integrate_occupancy(d, T, ocam, tp)
        else:
            o.integrate_depth(d, T, ocam)
        if esdf:
            m.update_esdf()
            blocks = b if i > 0 else (o.occupancy_block_indices() if occupancy else o.tsdf_block_indices())
            (o.integrate_esdf_occupancy if occupancy else o.integrate_esdf)(blocks)
    return m, o


def test_tsdf_single_decay_no_deallocation(gpu):
    orc = _orc()
    cs, cam, ocam = cameras(320, 240)
    frames = syn.make_sequence(syn.sphere_in_box(), cs, syn.circle_trajectory(40)[:3])
    m, o = _build_pair(frames, cam, ocam, esdf=False)
    m.tsdf_decay_integrator().params(decay_factor=0.75, deallocate_decayed_blocks=0)
    removed = m.decay()
    o.decay_tsdf(orc.default_tsdf_decay_params(decay_factor=0.75, deallocate_decayed_blocks=0))
    assert len(removed) == 0
    assert_tsdf_equal(m.tsdf_layer().as_dict(), o.tsdf_layer())
    m.close()


def test_tsdf_decay_until_removed_with_esdf_and_reuse(gpu):
    """Decay to deallocation, then keep mapping: freed slots are reused, the hashes were rebuilt, the ESDF twin blocks
    are gone, the next ESDF update covers every block."""
    nvb, orc = _nvb(), _orc()
    cs, cam, ocam = cameras(320, 240)
    frames = syn.make_sequence(syn.sphere_in_box(), cs, syn.circle_trajectory(40)[:6])
    m, o = _build_pair(frames[:3], cam, ocam)
    dp = orc.default_tsdf_decay_params(decay_factor=0.3)
    m.tsdf_decay_integrator().params(decay_factor=0.3)
    total_removed = 0
    for it in range(12):
        r_gpu = m.decay()
        r_cpu = o.decay_tsdf(dp)
        assert _as_set(r_gpu) == _as_set(r_cpu), it
        total_removed += len(r_gpu)
        assert_tsdf_equal(m.tsdf_layer().as_dict(), o.tsdf_layer())
        assert set(m.esdf_layer().as_dict()) == set(o.esdf_layer())
        if it == 5:  # interleave a frame + ESDF update (all blocks: the tracker was reset by the decay)
            d, T = frames[3]
            m.integrate_depth(d, T, cam)
            o.integrate_depth(d, T, ocam)
            m.update_esdf()
            o.integrate_esdf(o.tsdf_block_indices())
            assert_esdf_equal(m.esdf_layer().as_dict(), o.esdf_layer())
    assert total_removed > 100
    n_before = m.tsdf_layer().num_blocks()
    for i, (d, T) in enumerate(frames[4:]):
        b = m.integrate_depth(d, T, cam)
        o.integrate_depth(d, T, ocam)
        m.update_esdf()
        o.integrate_esdf(o.tsdf_block_indices() if i == 0 else b)
        assert_esdf_equal(m.esdf_layer().as_dict(), o.esdf_layer())
    assert m.tsdf_layer().num_blocks() > n_before
    assert_tsdf_equal(m.tsdf_layer().as_dict(), o.tsdf_layer())
    assert sorted(map(tuple, m.tsdf_layer().get_all_block_indices())) == sorted(o.tsdf_layer())
    m.close()


def test_tsdf_decay_exclude_last_view(gpu):
    orc = _orc()
    cs, cam, ocam = cameras(320, 240)
    frames = syn.make_sequence(syn.box_with_cube(), cs, syn.circle_trajectory(40)[:4])
    m, o = _build_pair(frames, cam, ocam, esdf=False)
    d, T = frames[-1]
    dp = orc.default_tsdf_decay_params(decay_factor=0.5)
    m.tsdf_decay_integrator().params(decay_factor=0.5)
    for _ in range(8):
        r_gpu = m.decay(depth=d, T_L_C=T, camera=cam)
        r_cpu = o.decay_tsdf(dp, depth=d, T_L_C=T, cam=ocam, max_view_distance_m=7.0, truncation_distance_m=4 * 0.05)
        assert _as_set(r_gpu) == _as_set(r_cpu)
    assert_tsdf_equal(m.tsdf_layer().as_dict(), o.tsdf_layer())
    # the view itself survived
    assert m.tsdf_layer().num_blocks() > 500
    w = np.concatenate([b["weight"].ravel() for b in m.tsdf_layer().as_dict().values()])
    assert (w > 0.5).any()
    m.close()


def test_tsdf_decay_exclusion_list_sphere_and_free_distance(gpu):
    orc = _orc()
    cs, cam, ocam = cameras(320, 240)
    frames = syn.make_sequence(syn.sphere_in_box(), cs, syn.circle_trajectory(40)[:2])
    m, o = _build_pair(frames, cam, ocam, esdf=False)
    idx = m.tsdf_layer().get_all_block_indices()
    excluded = idx[(idx[:, 0] % 2 == 0) | (idx[:, 2] % 3 == 0)]
    kw = dict(decay_factor=0.2, set_free_distance_on_decayed=1, deallocate_decayed_blocks=0)
    m.tsdf_decay_integrator().params(**kw)
    dp = orc.default_tsdf_decay_params(**kw)
    for _ in range(6):
        m.decay(excluded_blocks=excluded, exclusion_center=(1.0, 1.0, 1.0), exclusion_radius_m=1.5)
        o.decay_tsdf(dp, excluded_blocks=excluded, exclusion_center=(1.0, 1.0, 1.0), exclusion_radius_m=1.5)
    assert_tsdf_equal(m.tsdf_layer().as_dict(), o.tsdf_layer())
    m.close()


def test_occupancy_decay_with_view_and_deallocation(gpu):
    orc = _orc()
    cs, cam, ocam = cameras(320, 240)
    frames = syn.make_sequence(syn.sphere_in_box(), cs, syn.circle_trajectory(40)[:4])
    m, o = _build_pair(frames, cam, ocam, occupancy=True)
    d, T = frames[-1]
    kw = dict(free_region_decay_probability=0.7, occupied_region_decay_probability=0.3)
    m.occupancy_decay_integrator().params(**kw)
    dp = orc.default_occupancy_decay_params(**kw)
    removed_total = 0
    for it in range(6):
        use_view = it % 2 == 0
        r_gpu = m.decay(depth=d, T_L_C=T, camera=cam) if use_view else m.decay()
        r_cpu = (o.decay_occupancy(dp, depth=d, T_L_C=T, cam=ocam, max_view_distance_m=7.0, truncation_distance_m=4 * 0.05)
                 if use_view else o.decay_occupancy(dp))
        assert _as_set(r_gpu) == _as_set(r_cpu), it
        removed_total += len(r_gpu)
        _assert_occ_equal(m.occupancy_layer().as_dict(), o.occupancy_layer())
    assert removed_total > 0
    m.update_esdf()
    o.integrate_esdf_occupancy(o.occupancy_block_indices())
    assert_esdf_equal(m.esdf_layer().as_dict(), o.esdf_layer())
    m.occupancy_decay_integrator().decay_to_free(True)
    dp.decay_to_probability = 0.49
    m.decay(), o.decay_occupancy(dp)
    _assert_occ_equal(m.occupancy_layer().as_dict(), o.occupancy_layer())
    m.close()


def test_decay_empty_map_and_parameter_checks(gpu):
    nvb = _nvb()
    m = nvb.Mapper(0.05)
    assert len(m.decay()) == 0
    with pytest.raises(Exception):
        m.tsdf_decay_integrator().params(decay_factor=1.5)
    with pytest.raises(Exception):
        m.occupancy_decay_integrator().params(free_region_decay_probability=0.2)
    assert m.tsdf_decay_integrator().deallocate_decayed_blocks() is True
    m.close()


def test_persistent_cleared_list_survives_deallocation_and_reallocation(gpu):
    """A block on the persistent cleared list is deallocated and later allocated again while the list still
    persists (no block lost sites in between): the reference keeps block indices, so the block counts again."""
    nvb, orc = _nvb(), _orc()
    cs, cam, ocam = cameras(320, 240)
    frames = [(d, T) for d, T, _ in syn.moving_sphere_sequence(cs, syn.circle_trajectory(40)[:6], step_m=0.3)]
    m, o = _build_pair(frames[:4], cam, ocam)
    assert o.esdf_stats()["cleared"] > 0 or True
    dp = orc.default_tsdf_decay_params(decay_factor=0.05)
    m.tsdf_decay_integrator().params(decay_factor=0.05)
    d, T = frames[3]
    for _ in range(4):  # everything outside the last view disappears
        assert _as_set(m.decay(depth=d, T_L_C=T, camera=cam)) == _as_set(
            o.decay_tsdf(dp, depth=d, T_L_C=T, cam=ocam, max_view_distance_m=7.0, truncation_distance_m=0.2))
    for i, (d2, T2) in enumerate(frames[:3] + frames[4:]):  # old views again: blocks come back
        b = m.integrate_depth(d2, T2, cam)
        o.integrate_depth(d2, T2, ocam)
        m.update_esdf()
        o.integrate_esdf(o.tsdf_block_indices() if i == 0 else b)
        assert_esdf_equal(m.esdf_layer().as_dict(), o.esdf_layer())
    assert_tsdf_equal(m.tsdf_layer().as_dict(), o.tsdf_layer())
    m.close()


def test_keep_last_view_equals_explicit_view(gpu):
    nvb = _nvb()
    cs, cam, ocam = cameras(320, 240)
    frames = syn.make_sequence(syn.sphere_in_box(), cs, syn.circle_trajectory(40)[:3])
    a, b = nvb.Mapper(0.05, keep_last_view=True), nvb.Mapper(0.05)
    with pytest.raises(Exception):
        b.decay_exclude_last_view()  # created without keep_last_view
    assert len(a.decay_exclude_last_view()) == 0  # nothing integrated yet: decays all (of nothing)
    for d, T in frames:
        a.integrate_depth(d, T, cam)
        b.integrate_depth(d, T, cam)
    for m in (a, b):
        m.tsdf_decay_integrator().params(decay_factor=0.4)
    d, T = frames[-1]
    for _ in range(6):
        ra = a.decay_exclude_last_view()
        rb = b.decay(depth=d, T_L_C=T, camera=cam)
        assert _as_set(ra) == _as_set(rb)
    assert_tsdf_equal(a.tsdf_layer().as_dict(), b.tsdf_layer().as_dict())
    a.close(), b.close()
