"""Shared helpers for the parity tests (CUDA path vs the CPU oracle)."""
import numpy as np

from isaac_ros_nvblox_b200 import synthetic as syn


def cameras(width=640, height=480, f=300.0, radial=None, tangential=None):
    """(synthetic, product, oracle) camera objects with the reference's test intrinsics
    (nvblox/tests/test_esdf_integrator.cpp:121), scaled to the image size; optional lens distortion
    for the product and oracle cameras (the synthetic renderer stays pinhole: it only makes inputs)."""
    import isaac_ros_nvblox_b200 as nvb
    from oracle import oracle as orc
    s = f * width / 640.0
    cs = syn.PinholeCamera(s, s, width / 2.0, height / 2.0, width, height)
    ocam = orc.Camera(cs.fu, cs.fv, cs.cu, cs.cv, width, height)
    if radial is not None or tangential is not None:
        ocam = ocam.with_distortion(k=radial or (0,) * 6, p=tangential or (0, 0))
    return cs, nvb.Camera(cs.fu, cs.fv, cs.cu, cs.cv, width, height, radial, tangential), ocam


def sort_rows(a):
    a = np.asarray(a).reshape(-1, 3)
    return a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]


def assert_tsdf_equal(gpu_layer, cpu_layer, atol=1e-4, exact=True):
    assert set(gpu_layer) == set(cpu_layer), "allocated TSDF block sets differ"
    worst = 0.0
    for k, c in cpu_layer.items():
        g = gpu_layer[k]
        if exact:
            # same operation order, no FMA on either side -> bit-identical
            assert np.array_equal(g["distance"].view(np.uint32), c["distance"].view(np.uint32)), ("tsdf distance bits", k)
            assert np.array_equal(g["weight"].view(np.uint32), c["weight"].view(np.uint32)), ("tsdf weight bits", k)
        d = float(np.max(np.abs(g["distance"] - c["distance"]))) if g.size else 0.0
        w = float(np.max(np.abs(g["weight"] - c["weight"]))) if g.size else 0.0
        worst = max(worst, d, w)
    assert worst <= atol, "TSDF differs by %g (tolerance %g, BASELINE.json north_star)" % (worst, atol)
    return worst


ESDF_FIELDS = ("squared_distance_vox", "parent_direction", "is_inside", "observed", "is_site")


def assert_esdf_equal(gpu_layer, cpu_layer):
    assert set(gpu_layer) == set(cpu_layer), "allocated ESDF block sets differ"
    for k, c in cpu_layer.items():
        g = gpu_layer[k]
        for f in ESDF_FIELDS:
            if not np.array_equal(g[f], c[f]):
                bad = np.argwhere(g[f] != c[f])
                raise AssertionError("ESDF field %s differs in block %s at %d voxels (first %s: gpu %s cpu %s)" %
                                     (f, k, len(bad), tuple(bad[0]), g[f][tuple(bad[0])], c[f][tuple(bad[0])]))


def layer_checksum(layer, fields):
    """Order-independent checksum of a {index: voxels} layer (sum of per-block CRCs)."""
    import zlib
    total = 0
    for k in sorted(layer):
        h = zlib.crc32(np.asarray(k, dtype=np.int32).tobytes())
        for f in fields:
            h = zlib.crc32(np.ascontiguousarray(layer[k][f]).tobytes(), h)
        total = (total + h) & 0xFFFFFFFFFFFF
    return total


# ---------------------------------------------------------------------------
# The reference's slicing known-answer cases (tests/test_esdf_integrator_slicing.cu), shared by the oracle KATs
# (test_oracle_kat.py) and the GPU parity tests (test_gpu_esdf_slice.py).
# ---------------------------------------------------------------------------
SLICING_VOXEL = np.float32(0.05)
_FREE, _OCC = 1.0, -0.025  # kFreeDistance, kOccupiedDistance (:53-54)
TSDF_DT = np.dtype([("distance", "<f4"), ("weight", "<f4")])


def _tsdf_block(distance):
    b = np.zeros((8, 8, 8), TSDF_DT)
    b["distance"], b["weight"] = distance, 1.0
    return b


def unit_plane(normal, point):
    """Plane(normal, point) (geometry/internal/impl/plane_impl.h:25-32) -> (nx, ny, nz, d), float32."""
    n = np.asarray(normal, np.float32)
    n = (n / np.float32(np.sqrt(np.float32(np.dot(n, n))))).astype(np.float32)
    d = -np.float32(np.dot(np.asarray(point, np.float32), n))
    return np.array([n[0], n[1], n[2], d], np.float32)


def slicing_case(name, planar):
    """-> (blocks {idx: tsdf block}, slice kwargs, expected) for SingleBlock (:109-208), AcrossBlock (:210-287) and
    45DegreeSlice (:353-483). expected = {'num_esdf_blocks': n, 'sites': {block idx: set of (x, y) site columns or 'all'}}.
    The output slice is z = 0 (voxel layer 0 of block layer 0)."""
    vs = float(SLICING_VOXEL)
    if name == "single_block":
        blocks = {(0, 0, 0): _tsdf_block(_OCC), (1, 0, 0): _tsdf_block(_FREE), (2, 0, 0): _tsdf_block(_FREE),
                  (3, 0, 0): _tsdf_block(_FREE)}
        blocks[(2, 0, 0)]["distance"][3, 3, 3] = _OCC
        blocks[(3, 0, 0)]["distance"][5, 5, 5] = _OCC  # just above the band
        zmin, zmax = 0.0, np.float32(3.5 * SLICING_VOXEL)
        plane = unit_plane((0, 0, 1), (0, 0, zmin))
        exp = {"num_esdf_blocks": 4, "sites": {(0, 0, 0): "all", (1, 0, 0): set(), (2, 0, 0): {(3, 3)}, (3, 0, 0): set()}}
    elif name == "across_block":
        blocks = {(0, 0, 0): _tsdf_block(_FREE), (0, 0, 1): _tsdf_block(_FREE)}
        blocks[(0, 0, 1)]["distance"][0, 0, 1] = _OCC
        zmin, zmax = 0.0, np.float32(9.5 * SLICING_VOXEL)  # 2nd voxel of the top block
        plane = unit_plane((0, 0, 1), (0, 0, zmin))
        exp = {"num_esdf_blocks": 1, "sites": {(0, 0, 0): {(0, 0)}}}
    elif name == "45_degree":
        blocks = {(0, 0, 0): _tsdf_block(_FREE)}
        for v in ((1, 1, 1), (2, 2, 3), (3, 3, 6)):
            blocks[(0, 0, 0)]["distance"][v] = _OCC
        zmin, zmax = 0.0, np.float32(2.0 * SLICING_VOXEL)
        plane = unit_plane((-1, 0, 1), (0, 0, zmin))
        exp = {"num_esdf_blocks": 1, "sites": {(0, 0, 0): {(1, 1), (2, 2)} if planar else {(1, 1)}}}
    else:
        raise KeyError(name)
    if planar:
        kw = dict(plane=plane, above_plane_m=0.0, thickness_m=float(np.float32(zmax) - np.float32(zmin)), z_output_m=0.0)
    else:
        kw = dict(z_min_m=float(zmin), z_max_m=float(zmax), z_output_m=0.0)
    return blocks, kw, exp


def check_slicing_sites(esdf_layer, exp):
    assert len(esdf_layer) == exp["num_esdf_blocks"]
    for idx, sites in exp["sites"].items():
        got = esdf_layer[idx]["is_site"][:, :, 0].astype(bool)
        want = np.ones((8, 8), bool) if sites == "all" else np.zeros((8, 8), bool)
        if sites != "all":
            for x, y in sites:
                want[x, y] = True
        assert np.array_equal(got, want), (idx, np.argwhere(got != want))


def sphere_scene_tsdf_layer(voxel_size=0.05, truncation_m=0.2):
    """Scene::generateLayerFromScene (primitives/internal/impl/scene_impl.h:100-140) for the sphere-in-a-box scene of the
    TestScene case (:485-500): every voxel of the AABB gets the truncated ground-truth distance and weight 1."""
    scene = syn.sphere_in_box()
    bs = 8 * voxel_size
    lo = [int(np.floor(-5.0 / bs)), int(np.floor(-5.0 / bs)), 0]  # getBlockIndicesTouchedByBoundingBox
    hi = [int(np.floor(5.0 / bs)), int(np.floor(5.0 / bs)), int(np.floor(5.0 / bs))]
    ax = [(np.arange(8 * (h - l + 1)) + 0.5) * voxel_size + l * bs for l, h in zip(lo, hi)]
    P = np.stack(np.meshgrid(*ax, indexing="ij"), axis=-1)
    D = np.clip(scene.distance(P), -truncation_m, truncation_m).astype(np.float32)
    inside = np.all((P >= (-5.0, -5.0, 0.0)) & (P <= (5.0, 5.0, 5.0)), axis=-1)  # voxels outside the AABB stay unset
    idx, vox = [], []
    for bx in range(hi[0] - lo[0] + 1):
        for by in range(hi[1] - lo[1] + 1):
            for bz in range(hi[2] - lo[2] + 1):
                sl = (slice(8 * bx, 8 * bx + 8), slice(8 * by, 8 * by + 8), slice(8 * bz, 8 * bz + 8))
                b = np.zeros((8, 8, 8), TSDF_DT)
                b["distance"] = np.where(inside[sl], D[sl], 0.0)
                b["weight"] = np.where(inside[sl], 1.0, 0.0)
                idx.append((bx + lo[0], by + lo[1], bz + lo[2]))
                vox.append(b)
    return np.asarray(idx, np.int32), np.stack(vox)


def check_sphere_scene_slice(esdf_layer, planar, voxel_size=0.05):
    """TestScene (:536-555): the inside voxels of the slice lie within the sphere's outline."""
    n_inside = 0
    for (bx, by, bz), blk in esdf_layer.items():
        ins = np.argwhere(blk["is_inside"].astype(bool))
        for vx, vy, vz in ins:
            x = (8 * bx + vx + 0.5) * voxel_size
            y = (8 * by + vy + 0.5) * voxel_size
            assert np.hypot(x, y) < 2.0
            assert x < (2.0 if planar else 2.0 * np.sqrt(2.0))
            n_inside += 1
    return n_inside
