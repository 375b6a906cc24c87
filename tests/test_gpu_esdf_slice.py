"""Parity of the 2-D ESDF (EsdfIntegrator::integrateSlice with the constant-z slice, Mapper::updateEsdfSlice) with the
CPU oracle, through the C-ABI. All five EsdfVoxel fields exact, block sets equal."""
import numpy as np
import pytest

from helpers import assert_esdf_equal, cameras
from isaac_ros_nvblox_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _nvb():
    import isaac_ros_nvblox_b200 as nvb
    return nvb


def _orc():
    from oracle import oracle as orc
    return orc


@pytest.mark.parametrize("mode", ["tsdf", "occupancy", "tsdf_freespace"])
def test_esdf_slice_incremental_sequence(gpu, mode):
    nvb, orc = _nvb(), _orc()
    cs, cam, ocam = cameras(320, 240)
    frames = [(d, T) for d, T, _ in syn.moving_sphere_sequence(cs, syn.circle_trajectory(40)[:7], step_m=0.25)]
    ltype = {"tsdf": nvb.ProjectiveLayerType.kTsdf, "occupancy": nvb.ProjectiveLayerType.kOccupancy,
             "tsdf_freespace": nvb.ProjectiveLayerType.kTsdfWithFreespace}[mode]
    m, o = nvb.Mapper(0.05, projective_layer_type=ltype), orc.OracleMap(0.05)
    z = dict(slice_min_height_m=0.25, slice_max_height_m=1.45, slice_height_m=0.9)
    m.esdf_integrator().slice_params(**z)
    tp = orc.default_tsdf_params()
    fkw = dict(min_duration_since_occupied_for_freespace_ms=200)
    if mode == "tsdf_freespace":
        m.freespace_integrator().params(**fkw)
    for i, (d, T) in enumerate(frames):
        b = m.integrate_depth(d, T, cam)
        if mode == "occupancy":
            o.integrate_occupancy(d, T, ocam, tp)
        else:
            o.integrate_depth(d, T, ocam)
        if mode == "tsdf_freespace":
            m.update_freespace(1000 + 100 * i)
            o.update_freespace(o.tsdf_block_indices() if i == 0 else b, 1000 + 100 * i, orc.default_freespace_params(**fkw))
        m.update_esdf_slice()
        blocks = b if i > 0 else (o.occupancy_block_indices() if mode == "occupancy" else o.tsdf_block_indices())
        o.integrate_esdf_slice(blocks, z_min_m=0.25, z_max_m=1.45, z_output_m=0.9, from_occupancy=(mode == "occupancy"),
                               use_freespace=(mode == "tsdf_freespace"))
        assert_esdf_equal(m.esdf_layer().as_dict(), o.esdf_layer())
    s_gpu, s_cpu = m.esdf_integrator().last_stats(), o.esdf_stats()
    for k in ("marked", "with_sites", "to_clear", "cleared"):
        assert s_gpu[k] == s_cpu[k], (k, s_gpu, s_cpu)
    layer = m.esdf_layer().as_dict()
    assert len(layer) > 100 and len({k[2] for k in layer}) == 1  # one z layer of blocks
    assert sum(int(v["is_site"].sum()) for v in layer.values()) > 100
    m.close()


def test_esdf_slice_explicit_lists_mode_check_and_decay(gpu):
    nvb, orc = _nvb(), _orc()
    cs, cam, ocam = cameras(320, 240)
    frames = syn.make_sequence(syn.sphere_in_box(), cs, syn.circle_trajectory(40)[:4])
    m, o = nvb.Mapper(0.05), orc.OracleMap(0.05)
    for d, T in frames[:3]:
        b = m.integrate_depth(d, T, cam)
        o.integrate_depth(d, T, ocam)
        lst = np.vstack([b, b[:5]])
        m.esdf_integrator().integrate_slice(lst)  # defaults: band 0..1 m, output at 1 m
        o.integrate_esdf_slice(lst)
        assert_esdf_equal(m.esdf_layer().as_dict(), o.esdf_layer())
    with pytest.raises(Exception):
        m.update_esdf()  # the ESDF layer is 2-D now (EsdfMode)
    # decay: a slice block disappears only when its whole column is gone
    dp = orc.default_tsdf_decay_params(decay_factor=0.05)
    m.tsdf_decay_integrator().params(decay_factor=0.05)
    d, T = frames[2]
    for _ in range(4):
        r_gpu = m.decay(depth=d, T_L_C=T, camera=cam)
        r_cpu = o.decay_tsdf(dp, depth=d, T_L_C=T, cam=ocam, max_view_distance_m=7.0, truncation_distance_m=0.2, clear_esdf=2)
        assert set(map(tuple, r_gpu.tolist())) == set(map(tuple, r_cpu.tolist()))
        assert set(m.esdf_layer().as_dict()) == set(o.esdf_layer())
    d, T = frames[3]
    m.integrate_depth(d, T, cam)
    o.integrate_depth(d, T, ocam)
    m.update_esdf_slice()  # tracker reset by the decay: all blocks
    o.integrate_esdf_slice(o.tsdf_block_indices())
    assert_esdf_equal(m.esdf_layer().as_dict(), o.esdf_layer())
    m.clear()
    m.integrate_depth(d, T, cam)
    m.update_esdf()  # after clear() the mode is unset again
    m.close()


@pytest.mark.parametrize("two_d", [False, True])
def test_esdf_slicer_distance_image_and_occupancy_grid(gpu, two_d):
    """EsdfSlicer::sliceLayerToDistanceImage / occupancyGridFromSliceImage of a 3-D ESDF and of a 2-D slice ESDF: AABB, image
    size, every pixel and every grid cell equal to the oracle's."""
    nvb, orc = _nvb(), _orc()
    cs, cam, ocam = cameras(320, 240)
    frames = syn.make_sequence(syn.sphere_in_box(), cs, syn.circle_trajectory(40)[:4])
    m, o = nvb.Mapper(0.05), orc.OracleMap(0.05)
    for i, (d, T) in enumerate(frames):
        b = m.integrate_depth(d, T, cam)
        o.integrate_depth(d, T, ocam)
        if two_d:
            m.update_esdf_slice()
            o.integrate_esdf_slice(b if i else o.tsdf_block_indices())
        else:
            m.update_esdf()
            o.integrate_esdf(b if i else o.tsdf_block_indices())
    for h in (1.0, 0.93, 2.2, 40.0):
        aabb_g, img_g, grid_g = nvb.EsdfSlicer(m).slice_layer_to_distance_image(h, 1000.0, with_occupancy_grid=True)
        aabb_c, img_c, grid_c = o.esdf_slice_image(h, 1000.0)
        assert img_g.shape == img_c.shape
        if img_c.size:
            assert np.array_equal(aabb_g, aabb_c)
            assert np.array_equal(img_g.view(np.uint32), img_c.view(np.uint32))
            assert np.array_equal(grid_g, grid_c)
    aabb, img = nvb.EsdfSlicer(m).slice_layer_to_distance_image(1.0)
    assert img.size > 10000 and (img == 1000.0).any() and (np.abs(img) < 0.01).any() and (img[img != 1000.0] > 0.5).any()
    m.close()


@pytest.mark.parametrize("plane", [(0.0, 0.0, 1.0, -0.5), (0.0599, -0.0399, 0.9974, -0.3), (1.0, 0.0, 0.0, 0.0)])
def test_esdf_planar_slice(gpu, plane):
    """integrateSlice with a PlanarSliceDescription: per-column bounds from the ground plane (a level one, a tilted one, and a
    vertical one that falls back to z = 0)."""
    nvb, orc = _nvb(), _orc()
    n = np.asarray(plane[:3], np.float64)
    n = n / np.linalg.norm(n)
    pl = np.array([n[0], n[1], n[2], plane[3]], np.float32)
    cs, cam, ocam = cameras(320, 240)
    frames = syn.make_sequence(syn.box_with_cube(), cs, syn.circle_trajectory(40)[:4])
    m, o = nvb.Mapper(0.05), orc.OracleMap(0.05)
    m.esdf_integrator().slice_params(slice_height_above_plane_m=0.3, slice_height_thickness_m=0.85, slice_height_m=1.0)
    for i, (d, T) in enumerate(frames):
        b = m.integrate_depth(d, T, cam)
        o.integrate_depth(d, T, ocam)
        m.update_esdf_slice(ground_plane=pl)
        o.integrate_esdf_slice_planar(b if i else o.tsdf_block_indices(), pl, above_plane_m=0.3, thickness_m=0.85, z_output_m=1.0)
        assert_esdf_equal(m.esdf_layer().as_dict(), o.esdf_layer())
    m.esdf_integrator().integrate_slice(b[::2], ground_plane=pl)  # explicit list
    o.integrate_esdf_slice_planar(b[::2], pl, above_plane_m=0.3, thickness_m=0.85, z_output_m=1.0)
    assert_esdf_equal(m.esdf_layer().as_dict(), o.esdf_layer())
    assert sum(int(v["is_site"].sum()) for v in m.esdf_layer().as_dict().values()) > 50
    m.close()


def _slice_both(m, o, lst, planar, kw):
    if planar:
        m.esdf_integrator().slice_params(slice_height_above_plane_m=kw["above_plane_m"],
                                         slice_height_thickness_m=kw["thickness_m"], slice_height_m=kw["z_output_m"])
        m.esdf_integrator().integrate_slice(lst, ground_plane=kw["plane"])
        o.integrate_esdf_slice_planar(lst, kw["plane"], above_plane_m=kw["above_plane_m"], thickness_m=kw["thickness_m"],
                                      z_output_m=kw["z_output_m"])
    else:
        m.esdf_integrator().slice_params(slice_min_height_m=kw["z_min_m"], slice_max_height_m=kw["z_max_m"],
                                         slice_height_m=kw["z_output_m"])
        m.esdf_integrator().integrate_slice(lst)
        o.integrate_esdf_slice(lst, **kw)


@pytest.mark.parametrize("planar", [False, True])
@pytest.mark.parametrize("name", ["single_block", "across_block", "45_degree"])
def test_slicing_reference_cases(gpu, name, planar):
    """The reference's own slicing cases (tests/test_esdf_integrator_slicing.cu SingleBlock, AcrossBlock, 45DegreeSlice) on
    hand-set TSDF blocks: the expected site columns, and all fields equal to the oracle's."""
    from helpers import check_slicing_sites, slicing_case
    nvb, orc = _nvb(), _orc()
    blocks, kw, exp = slicing_case(name, planar)
    m, o = nvb.Mapper(0.05), orc.OracleMap(0.05)
    lst = np.asarray(list(blocks), np.int32)
    m.tsdf_layer().set_blocks(lst, np.stack([blocks[tuple(k)] for k in lst]))
    for idx, vox in blocks.items():
        o.set_tsdf_block(idx, vox)
    _slice_both(m, o, lst, planar, kw)
    layer = m.esdf_layer().as_dict()
    check_slicing_sites(layer, exp)
    assert_esdf_equal(layer, o.esdf_layer())
    m.close()


@pytest.mark.parametrize("planar", [False, True])
def test_slicing_sphere_scene(gpu, planar):
    """TestScene (:485-555): one-voxel band through the ground-truth TSDF of the sphere-in-a-box scene (8 788 blocks)."""
    from helpers import check_sphere_scene_slice, sphere_scene_tsdf_layer, unit_plane
    nvb, orc = _nvb(), _orc()
    idx, vox = sphere_scene_tsdf_layer()
    m, o = nvb.Mapper(0.05), orc.OracleMap(0.05)
    m.tsdf_layer().set_blocks(idx, vox)
    for k, v in zip(idx, vox):
        o.set_tsdf_block(k, v)
    zmax = float(np.float32(2.0) + np.float32(0.05))
    kw = (dict(plane=unit_plane((-1, 0, 1), (0, 0, 2)), above_plane_m=0.0, thickness_m=float(np.float32(zmax) - np.float32(2.0)),
               z_output_m=0.0) if planar else dict(z_min_m=2.0, z_max_m=zmax, z_output_m=0.0))
    _slice_both(m, o, idx, planar, kw)
    layer = m.esdf_layer().as_dict()
    assert len(layer) == 26 * 26
    assert check_sphere_scene_slice(layer, planar) > 3000
    assert_esdf_equal(layer, o.esdf_layer())
    m.close()


def test_esdf_slicer_combined_image_of_two_mappers(gpu):
    """EsdfSlicer::sliceLayersToCombinedDistanceImage (esdf_slicer.h:78-118): two mappers (what MultiMapper's static and dynamic
    maps are), different scenes and slice heights: merged box, image and occupancy grid equal to the oracle's, bit for bit; one
    empty layer; two empty layers."""
    nvb, orc = _nvb(), _orc()
    cs, cam, ocam = cameras(320, 240)
    poses = syn.circle_trajectory(40)
    pairs = []
    for sel, scene in ((poses[:3], syn.sphere_in_box()), (poses[14:17], syn.box_with_cube())):
        m, o = nvb.Mapper(0.05), orc.OracleMap(0.05)
        for i, (d, T) in enumerate(syn.make_sequence(scene, cs, sel)):
            b = m.integrate_depth(d, T, cam)
            o.integrate_depth(d, T, ocam)
            m.update_esdf()
            o.integrate_esdf(b if i else o.tsdf_block_indices())
        pairs.append((m, o))
    (m1, o1), (m2, o2) = pairs
    for h1, h2 in ((1.0, 1.0), (0.93, 1.7), (1.0, 40.0)):
        aabb_g, img_g, grid_g = nvb.EsdfSlicer(m1).slice_layers_to_combined_distance_image(m2, h1, h2, 1000.0, with_occupancy_grid=True)
        aabb_c, img_c = orc.combined_slice_image(o1, o2, h1, h2, 1000.0)
        assert np.array_equal(aabb_g, aabb_c) and img_g.shape == img_c.shape and img_c.size > 0
        assert np.array_equal(img_g.view(np.uint32), img_c.view(np.uint32))
        assert np.array_equal(grid_g == 100, img_c < np.float32(1e-2)) and np.array_equal(grid_g == -1, np.abs(img_c - 1000.0) < 1e-2)
        box = nvb.EsdfSlicer(m1).get_aabb_of_layer_at_height(h1)
        assert np.array_equal(box, o1.esdf_slice_aabb(h1))
    assert nvb.EsdfSlicer(m1).slice_layers_to_combined_distance_image(m2, 40.0, 41.0) == (None, None)
    assert nvb.EsdfSlicer(m1).get_aabb_of_layer_at_height(40.0) is None
    for m, _ in pairs:
        m.close()
