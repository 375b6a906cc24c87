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
