"""The oracle's mesh integrator (oracle/nvblox_oracle.c, MeshIntegrator::integrateBlocksGPU) against the known answers the
reference's own tests hold (nvblox/tests/test_mesh.cpp, nvblox/tests/test_mesh_appearance.cpp):
  * BlankMap (:85-99): an empty TSDF layer gives an empty mesh layer;
  * PlaneMesh / GPUPlaneTest (:101-155, :181-250): a plane through the origin facing -x -> no empty mesh blocks, vertices,
    normals and triangles of equal length, every vertex at x = 0 (1e-4), every normal (-1, 0, 0) (1e-4) -- which pins the
    table, the (c, b, a) winding and the interpolation;
  * ComplexScene (:157-179): a mesh comes out; IncrementalMesh (:252-347): meshing block by block == meshing at once;
  * RepeatabilityTest (:349-421): two runs agree;
  * WeldingTest / InPlaceWeldingTest (:423-505): welding shrinks every block's vertex list, the triangle list keeps its length,
    every index stays in range and every welded triangle still has its corners where they were (1 mm, the weld's quantum);
  * test_mesh_appearance.cpp: vertices take the colour of the voxel they fall into, gray (127) where there is no colour block.
"""
import numpy as np
import pytest

import mesh_cases as mc
from oracle import oracle as orc


def _oracle_with(layer):
    o = orc.OracleMap(mc.VOXEL)
    for k, v in layer.items():
        o.set_tsdf_block(k, v)
    return o


def test_blank_map():
    o = orc.OracleMap(mc.VOXEL)
    o.integrate_mesh()
    assert len(o.mesh_block_indices()) == 0


@pytest.mark.parametrize("weld", [False, True])
def test_plane_mesh_known_answer(weld):
    o = _oracle_with(mc.plane_scene())
    o.integrate_mesh(weld_vertices=weld)
    mesh = o.mesh_layer()
    assert 0 < len(mesh) <= len(o.tsdf_block_indices())
    for k, b in mesh.items():
        nv, nt = len(b["vertices"]), len(b["triangles"])
        assert nv > 0 and nt > 0 and len(b["normals"]) == nv
        if not weld:
            assert nv == nt and np.array_equal(b["triangles"], np.arange(nt))
        assert np.all(np.abs(b["vertices"][:, 0]) < 1e-4), k
        assert np.all(np.abs(b["normals"] - np.array([-1.0, 0.0, 0.0], np.float32)) < 1e-4), k
        assert k[0] == -1  # the surface belongs to the block whose cubes straddle it: x in [-0.8, 0)
    # the plane is covered exactly once: the blocks that touch the AABB span y -3.2..3.2, z 0..3.2, cube corners sit on voxel centres
    tri = sum(len(b["triangles"]) // 3 for b in mesh.values())
    v = np.concatenate([b["vertices"][b["triangles"]].reshape(-1, 3, 3) for b in mesh.values()])
    area = 0.5 * np.linalg.norm(np.cross(v[:, 1] - v[:, 0], v[:, 2] - v[:, 0]), axis=1).sum()
    assert tri == 2 * 63 * 31 and abs(area - 6.3 * 3.1) < 1e-3


def test_incremental_equals_batch_and_repeatable():
    layer = mc.welding_scene()
    a, b = _oracle_with(layer), _oracle_with(layer)
    a.integrate_mesh()
    a2 = a.mesh_layer()
    a.integrate_mesh()  # RepeatabilityTest
    for k, blk in a.mesh_layer().items():
        for f in ("vertices", "normals", "triangles"):
            assert np.array_equal(blk[f], a2[k][f])
    for k in b.tsdf_block_indices():  # IncrementalMesh: one block per call
        b.integrate_mesh(blocks=[k])
    bm = b.mesh_layer()
    assert bm.keys() == a2.keys() and len(bm) > 0
    for k in bm:
        for f in ("vertices", "normals", "triangles"):
            assert np.array_equal(bm[k][f], a2[k][f])


def test_welding_known_answer():
    layer = mc.welding_scene()
    o, w = _oracle_with(layer), _oracle_with(layer)
    o.integrate_mesh(weld_vertices=False)
    w.integrate_mesh(weld_vertices=True)
    plain, welded = o.mesh_layer(), w.mesh_layer()
    assert plain.keys() == welded.keys() and len(plain) > 0
    for k in plain:
        p, q = plain[k], welded[k]
        assert len(q["vertices"]) < len(p["vertices"]), k               # WeldingTest :458
        assert len(q["triangles"]) == len(p["triangles"])               # triangles keep their length (:668-686)
        assert len(q["normals"]) == len(q["vertices"])
        assert q["triangles"].min() >= 0 and q["triangles"].max() < len(q["vertices"])
        assert len(np.unique(q["vertices"], axis=0)) == len(q["vertices"])  # WeldingPartsTest :594: no duplicate left
        moved = np.abs(q["vertices"][q["triangles"]] - p["vertices"][p["triangles"]]).max()
        assert moved < 1.8e-3                                            # WeldingPartsTest :614: same corners, to the quantum
    # a block with 128 * 20 vertices or more is left alone (weldVerticesCubKernel :705-711)
    rng = np.random.default_rng(0)
    noisy = {k: v.copy() for k, v in layer.items()}
    key = next(iter(plain))
    noisy[key]["distance"] = rng.uniform(-0.4, 0.4, (8, 8, 8)).astype(np.float32)
    n, nw = _oracle_with(noisy), _oracle_with(noisy)
    n.integrate_mesh(weld_vertices=False), nw.integrate_mesh(weld_vertices=True)
    assert len(n.mesh_block(key)["vertices"]) >= 2560
    assert len(nw.mesh_block(key)["vertices"]) == len(n.mesh_block(key)["vertices"])


def test_unobserved_and_missing_neighbours_stop_the_mesh():
    layer = mc.plane_scene()
    key = (-1, 0, 0)
    o = _oracle_with(layer)
    o.integrate_mesh(weld_vertices=False)
    full = len(o.mesh_block(key)["vertices"])
    # a voxel below min_weight removes the 8 cubes that use it (:412-417)
    hole = {k: v.copy() for k, v in layer.items()}
    hole[key]["weight"][7, 3, 3] = 0.0
    o = _oracle_with(hole)
    o.integrate_mesh(weld_vertices=False)
    assert len(o.mesh_block(key)["vertices"]) < full
    # without the +y neighbour the cubes of the last y layer have no corners (:399-404)
    cut = {k: v for k, v in layer.items() if k != (-1, 1, 0)}
    o = _oracle_with(cut)
    o.integrate_mesh(weld_vertices=False)
    b = o.mesh_block(key)
    assert 0 < len(b["vertices"]) < full and b["vertices"][:, 1].max() <= 0.75 + 1e-6
    # a block that is nowhere near the surface is not meshable and gets no mesh block (:313-328)
    assert o.mesh_block((-3, 0, 0)) is None


def test_vertex_colours_follow_the_colour_layer():
    from helpers import cameras
    from isaac_ros_nvblox_b200 import synthetic as syn
    cs, _, ocam = cameras(320, 240)
    frames = syn.make_sequence(syn.sphere_in_box(), cs, syn.circle_trajectory(40)[:2])
    o = orc.OracleMap(0.05)
    rgb = np.zeros((240, 320, 3), np.uint8)
    rgb[..., 0], rgb[..., 1], rgb[..., 2] = 200, np.arange(320)[None, :] % 256, 30
    for d, T in frames:
        o.integrate_depth(d, T, ocam)
        o.integrate_color(rgb, T, ocam)
    o.integrate_mesh()
    o.update_mesh_color()
    mesh, col = o.mesh_layer(), o.color_layer()
    painted = gray = 0
    for k, b in mesh.items():
        assert len(b["colors"]) == len(b["vertices"])
        if k not in col:
            assert np.all(b["colors"][:, :3] == 127)
            gray += 1
            continue
        vi = np.clip(((b["vertices"] - np.array(k, np.float32) * np.float32(0.4)) / np.float32(0.05)).astype(int), 0, 7)
        want = col[k]["color"][vi[:, 0], vi[:, 1], vi[:, 2]] if "color" in col[k].dtype.names else None
        if want is not None:
            assert np.array_equal(b["colors"][:, :3], want[:, :3])
        painted += 1
    assert painted > 0
