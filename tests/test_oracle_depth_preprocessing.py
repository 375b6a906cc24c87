"""The oracle's DepthPreprocessor::dilateInvalidRegionsAsync (src/sensors/depth_preprocessing.cpp:36-58) against the reference's
own known answers (nvblox/tests/test_depth_image_preprocessing.cpp):
  * DilationNumberTests (lines 91-127): a 9x9 image of ones with a zero in the centre; after n dilations (n = 0..4) the image
    sums to 81 - (1 + 2n)^2;
  * NppThresholdDepthImage (lines 60-89): on the first 3DMatch frame (reference-held data, tests/golden/threedmatch_seq01.npz)
    every pixel that has an invalid (< 1e-2) pixel in its 3x3 neighbourhood is 0 after one dilation -- and, the converse the
    reference leaves implicit, every other pixel is untouched;
  * the Mapper wiring (mapper_impl.h:38-76): with do_depth_preprocessing the integrators see the dilated image.
"""
import os

import numpy as np
import pytest

from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("n", [0, 1, 2, 3, 4])
def test_dilation_number_known_answer(n):
    img = np.ones((9, 9), np.float32)
    img[4, 4] = 0.0
    out = orc.dilate_invalid(img, n)
    assert float(out.sum()) == 81 - (1 + 2 * n) ** 2
    w = 1 + 2 * n
    want = np.ones((9, 9), np.float32)
    want[4 - n:4 - n + w, 4 - n:4 - n + w] = 0.0
    assert np.array_equal(out, want)


def _box_any(mask, n):
    """Any invalid pixel within Chebyshev distance n (numpy, independent of the oracle's iteration)."""
    rows, cols = mask.shape
    pad = np.pad(mask, n, mode="constant")
    out = np.zeros_like(mask)
    for dy in range(2 * n + 1):
        for dx in range(2 * n + 1):
            out |= pad[dy:dy + rows, dx:dx + cols]
    return out


@pytest.mark.parametrize("n", [1, 4])
def test_threedmatch_frame_neighbourhood_property(n):
    fx = np.load(os.path.join(GOLDEN, "threedmatch_seq01.npz"))
    depth = (fx["depth_u16"][0].astype(np.float32) / np.float32(1000.0)).astype(np.float32)
    out = orc.dilate_invalid(depth, n)
    near_invalid = _box_any(depth < np.float32(1e-2), n)
    assert near_invalid.any() and not near_invalid.all()
    assert np.all(out[near_invalid] == 0.0)
    assert np.array_equal(out[~near_invalid], depth[~near_invalid])


def test_border_and_non_finite_pixels():
    """The replicated border adds nothing a clamped window does not hold; NaN does not compare below the threshold, -inf and
    negative depths do; custom threshold / value (DepthPreprocessor setters, depth_preprocessing.h:40-57)."""
    img = np.full((5, 7), 2.0, np.float32)
    img[0, 0], img[4, 6], img[2, 3] = 0.0, -np.inf, np.nan
    out = orc.dilate_invalid(img, 1)
    want = img.copy()
    want[0:2, 0:2] = 0.0
    want[3:5, 5:7] = 0.0
    assert np.array_equal(out, want, equal_nan=True)
    out = orc.dilate_invalid(np.full((3, 3), 0.5, np.float32), 0, threshold=1.0, value=-1.0)
    assert np.all(out == -1.0)


def test_mapper_integrates_the_dilated_image():
    from helpers import cameras
    _, _, ocam = cameras()
    depth = np.full((ocam.height, ocam.width), 2.0, np.float32)
    depth[100:140, 200:260] = 0.0
    T = np.eye(4, dtype=np.float32)
    a, b = orc.OracleMap(0.1), orc.OracleMap(0.1)
    a.depth_preprocessing(True, 3)
    la = a.integrate_depth(depth, T, ocam)
    lb = b.integrate_depth(orc.dilate_invalid(depth, 3), T, ocam)
    assert np.array_equal(la, lb) and len(la) > 0
    ta, tb = a.tsdf_layer(), b.tsdf_layer()
    assert ta.keys() == tb.keys()
    for k in ta:
        assert np.array_equal(ta[k]["distance"], tb[k]["distance"]) and np.array_equal(ta[k]["weight"], tb[k]["weight"])
    c = orc.OracleMap(0.1)
    c.integrate_depth(depth, T, ocam)
    wc = sum(float(v["weight"].sum()) for v in c.tsdf_layer().values())
    wa = sum(float(v["weight"].sum()) for v in ta.values())
    assert wa < wc  # the larger hole integrates fewer voxels
