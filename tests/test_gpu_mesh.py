"""Mesh integrator on the GPU (isaac_ros_nvblox_b200/csrc/nvb_mesh.cu) vs the CPU oracle: block sets, vertices, normals,
triangle indices and colours, exactly (both emit a block's triangles in x-major voxel order); plus the reference's own
known answers for the analytic plane scene (nvblox/tests/test_mesh.cpp:181-250)."""
import numpy as np
import pytest

import mesh_cases as mc
from helpers import cameras, sort_rows
from isaac_ros_nvblox_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _nvb():
    import isaac_ros_nvblox_b200 as nvb
    return nvb


def _orc():
    from oracle import oracle as orc
    return orc


def assert_mesh_equal(gpu, cpu, colors=False):
    assert set(gpu) == set(cpu), "mesh block sets differ"
    for k, c in cpu.items():
        g = gpu[k]
        for f in ("vertices", "normals", "triangles") + (("colors",) if colors else ()):
            assert g[f].shape == c[f].shape, (k, f, g[f].shape, c[f].shape)
            assert np.array_equal(g[f], c[f]), (k, f)


def _pair(layer, voxel=mc.VOXEL):
    nvb, orc = _nvb(), _orc()
    m, o = nvb.Mapper(voxel), orc.OracleMap(voxel)
    keys = np.array(list(layer.keys()), np.int32)
    m.tsdf_layer().set_blocks(keys, np.stack([layer[tuple(k)] for k in keys]))
    for k, v in layer.items():
        o.set_tsdf_block(k, v)
    return m, o


@pytest.mark.parametrize("weld", [False, True])
def test_plane_scene_known_answer_and_parity(gpu, weld):
    m, o = _pair(mc.plane_scene())
    mi = m.mesh_integrator()
    mi.weld_vertices(weld)
    assert mi.weld_vertices() is weld and abs(mi.min_weight() - 1e-4) < 1e-9
    mi.integrate_mesh_from_distance_field()
    o.integrate_mesh(weld_vertices=weld)
    mesh = m.mesh_layer().as_dict()
    assert_mesh_equal(mesh, o.mesh_layer())
    assert 0 < len(mesh) <= m.tsdf_layer().num_blocks()
    for k, b in mesh.items():
        assert len(b["vertices"]) > 0 and len(b["normals"]) == len(b["vertices"]) and len(b["triangles"]) > 0
        assert np.all(np.abs(b["vertices"][:, 0]) < 1e-4)
        assert np.all(np.abs(b["normals"] - np.array([-1.0, 0.0, 0.0], np.float32)) < 1e-4)
    assert sum(len(b["triangles"]) for b in mesh.values()) == 3 * 2 * 63 * 31
    m.close()


def test_welding_scene_parity_incremental_and_repeatable(gpu):
    layer = mc.welding_scene()
    m, o = _pair(layer)
    o.integrate_mesh()
    want = o.mesh_layer()
    mi = m.mesh_integrator()
    mi.integrate_mesh_from_distance_field()
    assert_mesh_equal(m.mesh_layer().as_dict(), want)
    mi.integrate_mesh_from_distance_field()  # RepeatabilityTest (:349-421): the second run re-meshes every block in place
    assert_mesh_equal(m.mesh_layer().as_dict(), want)
    # IncrementalMesh (:252-347): a few blocks per call on a fresh mapper, duplicates and absent blocks in the lists
    m2, _ = _pair(layer)
    keys = np.array(sorted(layer.keys()), np.int32)
    for i in range(0, len(keys), 7):
        chunk = np.concatenate([keys[i:i + 7], keys[i:i + 1], np.array([[99, 99, 99]], np.int32)])
        m2.mesh_integrator().integrate_blocks(chunk)
    assert_mesh_equal(m2.mesh_layer().as_dict(), want)
    # unwelded, and a block with >= 2560 vertices that the weld leaves alone
    rng = np.random.default_rng(0)
    key = next(iter(want))
    noisy = {k: v.copy() for k, v in layer.items()}
    noisy[key]["distance"] = rng.uniform(-0.4, 0.4, (8, 8, 8)).astype(np.float32)
    m3, o3 = _pair(noisy)
    m3.mesh_integrator().integrate_mesh_from_distance_field()
    o3.integrate_mesh()
    g3 = m3.mesh_layer().as_dict()
    assert_mesh_equal(g3, o3.mesh_layer())
    assert len(g3[key]["vertices"]) >= 2560 and len(g3[key]["vertices"]) == len(g3[key]["triangles"])
    for mm in (m, m2, m3):
        mm.close()


def test_holes_missing_neighbours_and_min_weight(gpu):
    layer = mc.plane_scene()
    key = (-1, 0, 0)
    hole = {k: v.copy() for k, v in layer.items() if k != (-1, 1, 0)}
    hole[key]["weight"][7, 3, 3] = 0.0
    hole[(-1, -1, 1)]["weight"][:, :, 4:] = 5e-3
    for min_weight in (1e-4, 1e-2):
        m, o = _pair(hole)
        m.mesh_integrator().min_weight(min_weight)
        m.mesh_integrator().integrate_mesh_from_distance_field()
        o.integrate_mesh(min_weight=min_weight)
        assert_mesh_equal(m.mesh_layer().as_dict(), o.mesh_layer())
        assert not m.mesh_layer().is_block_allocated((-3, 0, 0))
        m.close()


def test_mapper_update_mesh_sequence_with_colour_matches_oracle(gpu):
    """Mapper::updateColorMesh over a depth + colour sequence: incremental updates (the tracker's blocks only), then a full
    update; vertices, normals, indices and per-vertex colours equal to the oracle's at every step. The arena starts small,
    so it is repacked (grown + garbage-collected) several times on the way."""
    nvb, orc = _nvb(), _orc()
    cs, cam, ocam = cameras(320, 240)
    frames = syn.make_sequence(syn.sphere_in_box(), cs, syn.circle_trajectory(40)[:6])
    yy, xx = np.mgrid[0:240, 0:320]
    rgb = np.stack([xx % 256, yy % 256, (xx // 8 + yy // 8) % 2 * 200 + 20], axis=-1).astype(np.uint8)
    m, o = nvb.Mapper(0.05), orc.OracleMap(0.05)
    pending, first = [], True
    for i, (d, T) in enumerate(frames):
        b = m.integrate_depth(d, T, cam)
        assert np.array_equal(b, o.integrate_depth(d, T, ocam))
        m.integrate_color(np.roll(rgb, 5 * i, axis=1), T, cam)
        o.integrate_color(np.roll(rgb, 5 * i, axis=1), T, ocam)
        pending.append(b)
        if i % 2 == 1:
            continue  # two frames' blocks accumulate in the tracker
        m.update_mesh()
        todo = o.tsdf_block_indices() if first else np.unique(np.concatenate(pending), axis=0)
        pending, first = [], False
        o.integrate_mesh(blocks=todo)
        o.update_mesh_color(blocks=todo)
        assert_mesh_equal(m.mesh_layer().as_dict(), o.mesh_layer(), colors=True)
    m.update_mesh(update_full_layer=True)
    o.integrate_mesh()
    o.update_mesh_color()
    got = m.mesh_layer().as_dict()
    assert_mesh_equal(got, o.mesh_layer(), colors=True)
    st = m.mesh_layer().arena_stats()
    assert st["capacity"] >= st["used"] > 0
    assert sum(len(b["triangles"]) for b in got.values()) <= st["used"]
    m.close()


def test_occupancy_mapper_has_no_mesh(gpu):
    nvb = _nvb()
    m = nvb.Mapper(0.05, projective_layer_type=nvb.ProjectiveLayerType.kOccupancy)
    m.update_mesh()  # "Mesh is only updated for Tsdf layers" (src/mapper/mapper.cpp:380-383)
    assert m.mesh_layer().num_blocks() == 0
    with pytest.raises(RuntimeError):
        m.mesh_integrator().integrate_blocks([[0, 0, 0]])
    m.close()


def test_decay_removes_mesh_blocks_with_their_tsdf_blocks(gpu):
    """Mapper::clearBlocksInLayers (src/mapper/mapper.cpp:546-557): blocks the decay deallocates leave the mesh layer too."""
    nvb, orc = _nvb(), _orc()
    cs, cam, ocam = cameras(320, 240)
    frames = syn.make_sequence(syn.sphere_in_box(), cs, syn.circle_trajectory(40)[:2])
    m = nvb.Mapper(0.05)
    for d, T in frames:
        m.integrate_depth(d, T, cam)
    m.update_mesh()
    before = m.mesh_layer().num_blocks()
    assert before > 0
    m.tsdf_decay_integrator().params(decay_factor=1e-6, decayed_weight_threshold=1e-3)
    gone = m.decay_tsdf()
    assert len(gone) > 0
    assert m.tsdf_layer().num_blocks() == 0 and m.mesh_layer().num_blocks() == 0
    # a new frame after the wipe meshes again from scratch
    m.integrate_depth(frames[0][0], frames[0][1], cam)
    m.update_mesh()
    assert 0 < m.mesh_layer().num_blocks() <= before
    m.close()
    # a PARTIAL removal, side by side with the oracle: blocks outside a sphere decay away, the mesh blocks of the rest stay as
    # they are (vertices, normals, indices), later mesh updates reuse the freed header slots and arena space
    m, o = nvb.Mapper(0.05), orc.OracleMap(0.05)
    for d, T in frames:
        assert np.array_equal(m.integrate_depth(d, T, cam), o.integrate_depth(d, T, ocam))
    m.update_mesh()
    o.integrate_mesh()
    dp = orc.default_tsdf_decay_params(decay_factor=1e-6, decayed_weight_threshold=1e-3)
    m.tsdf_decay_integrator().params(decay_factor=1e-6, decayed_weight_threshold=1e-3)
    center = tuple(float(c) for c in frames[0][1][:3, 3] + 2.5 * frames[0][1][:3, 2])
    gone_g = m.decay(exclusion_center=center, exclusion_radius_m=1.2)
    gone_c = o.decay_tsdf(dp, exclusion_center=center, exclusion_radius_m=1.2)
    assert len(gone_c) > 0 and np.array_equal(sort_rows(gone_g), sort_rows(gone_c))
    assert 0 < m.tsdf_layer().num_blocks() == len(o.tsdf_block_indices())
    assert_mesh_equal(m.mesh_layer().as_dict(), o.mesh_layer())
    assert 0 < m.mesh_layer().num_blocks() < before
    b = m.integrate_depth(frames[1][0], frames[1][1], cam)
    assert np.array_equal(b, o.integrate_depth(frames[1][0], frames[1][1], ocam))
    m.update_mesh()  # after a deallocation the tracker restarts with every block (like the ESDF's)
    o.integrate_mesh()
    assert_mesh_equal(m.mesh_layer().as_dict(), o.mesh_layer())
    m.close()
