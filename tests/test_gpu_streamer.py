"""Serializers and streamers on real layers (GPU): nvblox/tests/test_layer_serializer_gpu.cpp, test_mesh_serializer.cpp and the
two layer tests of test_layer_streamer.cpp (SerializeNBytes :226-249, StreamNBytes :251-300)."""
import numpy as np
import pytest

import mesh_cases as mc
from isaac_ros_nvblox_b200 import streamer as st

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene(gpu):
    import isaac_ros_nvblox_b200 as nvb
    layer = mc.welding_scene()
    m = nvb.Mapper(mc.VOXEL)
    keys = np.array(sorted(layer.keys()), np.int32)
    m.tsdf_layer().set_blocks(keys, np.stack([layer[tuple(k)] for k in keys]))
    m.mesh_integrator().integrate_mesh_from_distance_field(update_color=True)
    yield m, layer, keys
    m.close()


def test_layer_serializer(scene):
    """serializeAllBlocks / serializeNoBlocks / serializeEmptyLayer (test_layer_serializer_gpu.cpp:43-95)."""
    m, layer, keys = scene
    s = st.serialize_voxel_layer(m.tsdf_layer(), keys)
    assert np.array_equal(s["block_indices"], keys) and len(s["block_offsets"]) == len(keys) + 1
    assert s["block_offsets"][0] == 0 and s["block_offsets"][-1] == len(s["voxels"]) == 512 * len(keys)
    for i, k in enumerate(keys):
        got = s["voxels"][s["block_offsets"][i]:s["block_offsets"][i + 1]].reshape(8, 8, 8)
        assert np.array_equal(got["distance"], layer[tuple(k)]["distance"]) and np.array_equal(got["weight"], layer[tuple(k)]["weight"])
    e = st.serialize_voxel_layer(m.tsdf_layer(), np.zeros((0, 3), np.int32))
    assert len(e["voxels"]) == 0 and len(e["block_indices"]) == 0 and e["block_offsets"].tolist() == [0]
    missing = np.array([[500, 500, 500], keys[0], [501, 0, 0]], np.int32)  # blocks that are not in the layer: size 0
    p = st.serialize_voxel_layer(m.tsdf_layer(), missing)
    assert p["block_offsets"].tolist() == [0, 0, 512, 512] and len(p["block_indices"]) == 3


def test_mesh_serializer(scene):
    """serializeAllBlocks, serializeSomeblocks, serializeFirstBlock, serializeLastBlock, serializeNoBlocks, serializeOneEmptyBlock
    (test_mesh_serializer.cpp:131-210): every serialized block equals the mesh block, offsets are consistent."""
    m, _, _ = scene
    ml = m.mesh_layer()
    idx = ml.get_all_block_indices()
    mesh = ml.as_dict()
    assert len(idx) > 4
    for sel in (idx, idx[::3], idx[:1], idx[-1:], np.zeros((0, 3), np.int32), np.array([[900, 0, 0]], np.int32)):
        s = st.serialize_mesh_layer(ml, sel)
        n = len(sel)
        assert len(s["vertex_block_offsets"]) == n + 1 and len(s["triangle_index_block_offsets"]) == n + 1
        assert s["vertex_block_offsets"][-1] == len(s["vertices"]) == len(s["vertex_appearances"])
        assert s["triangle_index_block_offsets"][-1] == len(s["triangle_indices"])
        for i, k in enumerate(sel):
            b = mesh.get(tuple(int(c) for c in k))
            v0, v1 = s["vertex_block_offsets"][i], s["vertex_block_offsets"][i + 1]
            t0, t1 = s["triangle_index_block_offsets"][i], s["triangle_index_block_offsets"][i + 1]
            if b is None:
                assert v0 == v1 and t0 == t1
                continue
            assert np.array_equal(s["vertices"][v0:v1], b["vertices"]) and np.array_equal(s["vertex_appearances"][v0:v1], b["colors"])
            assert np.array_equal(s["triangle_indices"][t0:t1], b["triangles"])


def test_streamers_on_layers(scene):
    """SerializeNBytes (TSDF layer, half of its bytes) and StreamNBytes (mesh layer: half the bytes within 5 %, then the rest)."""
    m, _, keys = scene
    tsdf = m.tsdf_layer()
    s = st.LayerStreamerOldestBlocks()
    s.mark_indices_candidates(keys)
    budget = len(keys) * 4096 // 2
    ser = s.get_n_bytes_of_serialized_blocks(budget, tsdf)
    assert 0 < len(ser["voxels"]) * 8 <= budget
    ml = m.mesh_layer()
    idx = ml.get_all_block_indices()
    sizes = {tuple(int(c) for c in k): st.size_in_bytes(ml, k) for k in idx}
    total = sum(sizes.values())
    ms = st.LayerStreamerOldestBlocks()
    ms.mark_indices_candidates(idx)
    half = ms.get_n_bytes_of_blocks(total // 2, ml)
    got = sum(sizes[tuple(k)] for k in half.tolist())
    assert got < (total + 1) // 2 and abs(got / total - 0.5) < 0.05
    rest = ms.get_n_bytes_of_blocks(total, ml)
    assert got + sum(sizes[tuple(k)] for k in rest.tolist()) == total and ms.num_candidates() == 0
    assert not ({tuple(k) for k in half.tolist()} & {tuple(k) for k in rest.tolist()})
    # exclusion by radius on a real layer + serialization of what is left
    ms.mark_indices_candidates(idx)
    p = st.BlockExclusionParams(exclusion_center_m=(-2.0, -2.0, 0.0), exclusion_radius_m=1.5, block_size_m=m.block_size())
    ser = ms.get_n_bytes_of_serialized_blocks(1 << 40, ml, p)
    c = 0.8 * (ser["block_indices"] + 0.5) - np.array([-2.0, -2.0, 0.0])
    assert len(ser["block_indices"]) > 0 and np.all((c * c).sum(axis=1) <= 1.5 ** 2 + 1e-5)
    assert ser["vertex_block_offsets"][-1] == len(ser["vertices"]) > 0
