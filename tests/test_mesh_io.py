"""MeshBlockLayer::getMesh and io::outputColorMeshLayerToPly (isaac_ros_nvblox_b200/io.py) on the oracle's mesh of the analytic
plane scene (CPU only): the monolithic mesh is the blocks appended in order with offset triangle indices; the PLY file has the
reference's header and round-trips."""
import numpy as np

import mesh_cases as mc
from isaac_ros_nvblox_b200 import io as nio
from oracle import oracle as orc


def _plane_mesh(color):
    o = orc.OracleMap(mc.VOXEL)
    for k, v in mc.plane_scene().items():
        o.set_tsdf_block(k, v)
    o.integrate_mesh()
    if color:
        o.update_mesh_color()  # no colour layer: Color::Gray() everywhere
    return o.mesh_layer()


def test_get_mesh_appends_blocks_with_offset_indices():
    layer = _plane_mesh(color=True)
    mesh = nio.get_mesh(layer)
    nv = sum(len(b["vertices"]) for b in layer.values())
    nt = sum(len(b["triangles"]) for b in layer.values())
    assert mesh["vertices"].shape == (nv, 3) and mesh["vertex_normals"].shape == (nv, 3) and mesh["vertex_appearances"].shape == (nv, 4)
    assert mesh["triangles"].shape == (nt,) and mesh["triangles"].min() == 0 and mesh["triangles"].max() == nv - 1
    off = 0
    t0 = 0
    for b in layer.values():
        n, t = len(b["vertices"]), len(b["triangles"])
        assert np.array_equal(mesh["vertices"][off:off + n], b["vertices"])
        assert np.array_equal(mesh["triangles"][t0:t0 + t], b["triangles"] + off)
        off, t0 = off + n, t0 + t
    # the triangles of the monolithic mesh are the plane: every corner at x = 0
    tri = mesh["vertices"][mesh["triangles"]].reshape(-1, 3, 3)
    assert np.all(np.abs(tri[..., 0]) < 1e-4) and len(tri) == 2 * 63 * 31
    assert nio.get_mesh({})["vertices"].shape == (0, 3)


def test_ply_round_trip(tmp_path):
    for color in (False, True):
        layer = _plane_mesh(color)
        path = str(tmp_path / ("mesh_%d.ply" % color))
        assert nio.output_color_mesh_layer_to_ply(layer, path)
        props, verts, faces = nio.read_ply(path)
        assert props == ["x", "y", "z", "nx", "ny", "nz"] + (["red", "green", "blue"] if color else [])
        mesh = nio.get_mesh(layer)
        assert verts.shape[0] == len(mesh["vertices"]) and faces.shape[0] == len(mesh["triangles"]) // 3
        assert np.allclose(verts[:, :3], mesh["vertices"], rtol=1e-5, atol=1e-6)  # 6 significant digits, like operator<<
        assert np.allclose(verts[:, 3:6], mesh["vertex_normals"], atol=1e-5)
        assert np.array_equal(faces.reshape(-1), mesh["triangles"])
        if color:
            assert np.all(verts[:, 6:9] == 127)
    assert nio.output_color_mesh_layer_to_ply({}, str(tmp_path / "empty.ply")) is False
    assert not (tmp_path / "empty.ply").exists()
