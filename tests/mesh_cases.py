"""Analytic TSDF layers for the mesh tests, restating the scenes of the reference's mesh tests (nvblox/tests/test_mesh.cpp:
MeshTest::SetUp :42-58 -- 10 cm voxels, scene AABB (-3,-3,0)..(3,3,3); PlaneMesh :101-109; WeldingTest :423-434) the way
primitives::Scene::generateLayerFromScene fills a TsdfLayer: every voxel of every block that touches the AABB gets the scene's
signed distance at its centre, clipped to +-max_dist, weight 1."""
import numpy as np

from isaac_ros_nvblox_b200 import TSDF_VOXEL_DTYPE

VOXEL = 0.1
BLOCK = 8 * VOXEL
AABB = (np.array([-3.0, -3.0, 0.0]), np.array([3.0, 3.0, 3.0]))


def plane(center, normal):
    c, n = np.asarray(center, np.float64), np.asarray(normal, np.float64)
    n = n / np.linalg.norm(n)
    return lambda p: (p - c) @ n


def sphere(center, radius):
    c = np.asarray(center, np.float64)
    return lambda p: np.linalg.norm(p - c, axis=-1) - radius


def layer_from_scene(primitives, max_dist=4 * VOXEL, aabb=AABB, voxel=VOXEL):
    """{(bx, by, bz): (8,8,8) TSDF voxels}: min over the primitives' signed distances (a union of solids)."""
    block = 8 * voxel
    lo = np.floor(aabb[0] / block).astype(int)
    hi = np.floor((aabb[1] - 1e-6) / block).astype(int)
    g = (np.arange(8) + 0.5) * voxel
    out = {}
    for bx in range(lo[0], hi[0] + 1):
        for by in range(lo[1], hi[1] + 1):
            for bz in range(lo[2], hi[2] + 1):
                p = np.stack(np.meshgrid(bx * block + g, by * block + g, bz * block + g, indexing="ij"), axis=-1)
                d = np.min(np.stack([f(p) for f in primitives]), axis=0)
                v = np.zeros((8, 8, 8), TSDF_VOXEL_DTYPE)
                v["distance"] = np.clip(d, -max_dist, max_dist).astype(np.float32)
                v["weight"] = 1.0
                out[(bx, by, bz)] = v
    return out


def plane_scene():
    return layer_from_scene([plane((0.0, 0.0, 0.0), (-1, 0, 0))])


def welding_scene():
    return layer_from_scene([plane((0.0, 0.0, 0.0), (-1, 0, 0)), plane((2.1, 0.1, 0.1), (0, -1, 0)), sphere((-2, -2, 0), 2.0)])


def canonical_triangles(block):
    """A mesh block as an order-independent multiset: each triangle = its three vertex positions (quantised to 1e-5 m),
    rotated so the smallest comes first (the winding is kept). The parity bar for the reference's own output, whose
    triangle order within a block is an atomicAdd race."""
    v = block["vertices"][block["triangles"]].reshape(-1, 3, 3).astype(np.float64)
    q = np.round(v * 1e5).astype(np.int64)
    rows = []
    for t in q:
        keys = [tuple(p) for p in t]
        r = keys.index(min(keys))
        rows.append(keys[r] + keys[(r + 1) % 3] + keys[(r + 2) % 3])
    return sorted(rows)
