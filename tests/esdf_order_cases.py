"""Hand-derived known answers for the ORDER-DEPENDENT parts of computeEsdf, written down from the reference's source
(nvblox/src/integrators/esdf_integrator.cu), not from the oracle: both the oracle (tests/test_oracle_esdf_order_kat.py) and the
CUDA path (tests/test_gpu_parity.py) are checked against them.

Case "last taker" -- sweepSingleBand (:542-600). The scan along a line carries ONE candidate site: the last voxel that was not
improved and has a valid distance replaces the candidate with its own parent even when the carried one would serve later voxels
better. A single block whose EsdfVoxels along the x line (., 0, 0) are preset (they point at sites outside the line, as left by
earlier updates); every other voxel is observed and empty (squared distance = max, parent 0). One real site at (7, 7, 7) makes the
block a member of computeEsdf's list; it is too far (>= 98 voxels^2) to touch the line. max_sq = (2.0 m / 0.05 m)^2 = 1600.

x sweep, forward (positions 0..7; "c" = carried site as (position on the line, perpendicular offset), d = squared distance):
  0: own sq 9, parent (0,3,0): nothing carried yet, valid -> c = (0, (3,0))
  1: own sq 9, parent (-3,0,0): d(c) = 1 + 9 = 10, 9 > 10 is false, valid -> c = (-2, (0,0))   [(0,(3,0)) is dropped here]
  2: own sq 14, parent (1,2,3): d(c) = 16, 14 > 16 false, valid -> c = (3, (2,3))      [the dropped site would give 4 + 9 = 13 < 14:
                                                                                         a "best so far" scan improves this voxel]
  3: empty (1600): d(c) = 0 + 13 -> sq 13, parent (0,2,3)
  4: empty: d(c) = 1 + 13 -> sq 14, parent (-1,2,3)
  5: own sq 20, parent (2,4,0): d(c) = 4 + 13 = 17 < 20 -> sq 17, parent (-2,2,3)
  6: own sq 10, parent (1,3,0): d(c) = 9 + 13 = 22, 10 > 22 false, valid -> c = (7, (3,0))
  7: empty: d(c) = 0 + 9 -> sq 9, parent (0,3,0)
backward (7..0):
  7: valid -> c = (7, (3,0));  6: d = 1 + 9 = 10, 10 > 10 false -> c = own = (7, (3,0));  5: d = 4 + 9 = 13 < 17 -> sq 13, parent (2,3,0)
  4: d = 9 + 9 = 18, 14 > 18 false -> c = own = (3, (2,3));  3: d = 13, 13 > 13 false -> c = own (same);  2: d = 1 + 13 = 14, 14 > 14 false
  -> c = own = (3, (2,3));  1: d = 4 + 13 = 17, 9 > 17 false -> c = own = (-2, (0,0));  0: d = 4 < 9 -> sq 4, parent (-2,0,0)
The y and z sweeps that follow only spread these sites to the rest of the block: every line through a voxel (x, 0, 0) starts at
that voxel, so what comes back to it in the backward pass is its own site at its own distance (not smaller: no update).

Case "across a face" -- updateNeighborBands / updateSingleNeighbor (:602-633, :1323-1386) + the sweeps after it. Blocks A = (0,0,0)
and B = (1,0,0), all voxels observed and free except one site at A(7,0,0). computeEsdf: sweep A (exact distances to the site inside
A); ring 1: the +x pass copies A's x = 7 face to B's x = 0 face with parent (0,-y,-z) - (+1 along x) = (-1,-y,-z), B is swept; ring 2:
B's faces give nobody anything better (A's face already holds the same site at distance y^2 + z^2 = (-1+1)^2 + y^2 + z^2). So
A(x,y,z): parent (7-x,-y,-z), B(x,y,z): parent (-1-x,-y,-z), squared distance = |parent|^2; 2 rings, 2 blocks swept after the seed... see
EXPECTED_STATS.
"""
import numpy as np

VOXEL = 0.05
MAX_SQ = np.float32((np.float32(2.0) / np.float32(VOXEL)) ** 2)  # 1600
FREE, SITE = 0.2, -0.01  # TSDF distances: observed free space; inside and within max_site_distance of the surface


def _tsdf_block(dtype, sites=()):
    b = np.zeros((8, 8, 8), dtype)
    b["distance"], b["weight"] = FREE, 1.0
    for v in sites:
        b["distance"][v] = SITE
    return b


def _empty_esdf(dtype):
    e = np.zeros((8, 8, 8), dtype)
    e["squared_distance_vox"], e["observed"] = MAX_SQ, 1
    return e


def last_taker_case(tsdf_dtype, esdf_dtype):
    """-> (tsdf blocks {idx: voxels}, preset esdf blocks {idx: voxels}, expected {voxel (x,y,z) of block (0,0,0): (sq, parent)})."""
    e = _empty_esdf(esdf_dtype)
    preset = {0: (9, (0, 3, 0)), 1: (9, (-3, 0, 0)), 2: (14, (1, 2, 3)), 5: (20, (2, 4, 0)), 6: (10, (1, 3, 0))}
    for x, (sq, p) in preset.items():
        e["squared_distance_vox"][x, 0, 0] = sq
        e["parent_direction"][x, 0, 0] = p
    expected = {
        (0, 0, 0): (4, (-2, 0, 0)), (1, 0, 0): (9, (-3, 0, 0)), (2, 0, 0): (14, (1, 2, 3)), (3, 0, 0): (13, (0, 2, 3)),
        (4, 0, 0): (14, (-1, 2, 3)), (5, 0, 0): (13, (2, 3, 0)), (6, 0, 0): (10, (1, 3, 0)), (7, 0, 0): (9, (0, 3, 0)),
        # spread by the y / z sweeps from the line's voxels (site of (0,0,0) is at (-2,0,0); of (3,0,0) at (3,2,3))
        (0, 1, 0): (5, (-2, -1, 0)), (3, 0, 1): (8, (0, 2, 2)),
        # the real site and its neighbourhood
        (7, 7, 7): (0, (0, 0, 0)), (6, 7, 7): (1, (1, 0, 0)), (7, 5, 6): (5, (0, 2, 1)),
    }
    return {(0, 0, 0): _tsdf_block(tsdf_dtype, sites=[(7, 7, 7)])}, {(0, 0, 0): e}, expected


def across_face_case(tsdf_dtype):
    """-> (tsdf blocks, expected parent function per block, expected stats)."""
    blocks = {(0, 0, 0): _tsdf_block(tsdf_dtype, sites=[(7, 0, 0)]), (1, 0, 0): _tsdf_block(tsdf_dtype)}
    x, y, z = np.meshgrid(np.arange(8), np.arange(8), np.arange(8), indexing="ij")
    parents = {(0, 0, 0): np.stack([7 - x, -y, -z], axis=-1), (1, 0, 0): np.stack([-1 - x, -y, -z], axis=-1)}
    return blocks, parents


ACROSS_FACE_STATS = {"with_sites": 1, "to_clear": 0, "swept": 2, "face_passes": 12, "rings": 2}


def check_across_face(esdf_layer, parents):
    assert set(esdf_layer) == set(parents)
    for idx, p in parents.items():
        blk = esdf_layer[idx]
        assert np.array_equal(blk["parent_direction"], p), idx
        assert np.array_equal(blk["squared_distance_vox"], (p.astype(np.int64) ** 2).sum(-1).astype(np.float32)), idx
        assert blk["observed"].all() and not blk["is_inside"][blk["is_site"] == 0].any()
    site = esdf_layer[(0, 0, 0)]["is_site"].astype(bool)
    assert site.sum() == 1 and site[7, 0, 0]


def check_last_taker(esdf_block, expected):
    for v, (sq, p) in expected.items():
        got = esdf_block[v]
        assert float(got["squared_distance_vox"]) == float(sq) and tuple(int(c) for c in got["parent_direction"]) == tuple(p), \
            (v, float(got["squared_distance_vox"]), tuple(got["parent_direction"]), sq, p)
