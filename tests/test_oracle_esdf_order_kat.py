"""The oracle's computeEsdf against hand-derived known answers (tests/esdf_order_cases.py): the scan's "last taker" rule and the
propagation across a block face. These pin the ORDER semantics by something other than the restatement itself."""
import numpy as np

from esdf_order_cases import ACROSS_FACE_STATS, across_face_case, check_across_face, check_last_taker, last_taker_case
from oracle import oracle as orc


def test_sweep_carries_the_last_taker_not_the_best_site():
    tsdf, esdf, expected = last_taker_case(orc.TSDF_VOXEL_DTYPE, orc.ESDF_VOXEL_DTYPE)
    o = orc.OracleMap(0.05)
    for k, v in tsdf.items():
        o.set_tsdf_block(k, v)
    for k, v in esdf.items():
        o.set_esdf_block(k, v)
    o.integrate_esdf(np.array(list(tsdf), np.int32))
    s = o.esdf_stats()
    assert s["with_sites"] == 1 and s["to_clear"] == 0 and s["swept"] == 1 and s["rings"] == 1
    check_last_taker(o.esdf_layer()[(0, 0, 0)], expected)


def test_propagation_across_a_face_and_ring_statistics():
    tsdf, parents = across_face_case(orc.TSDF_VOXEL_DTYPE)
    o = orc.OracleMap(0.05)
    for k, v in tsdf.items():
        o.set_tsdf_block(k, v)
    o.integrate_esdf(np.array(list(tsdf), np.int32))
    s = o.esdf_stats()
    for k, v in ACROSS_FACE_STATS.items():
        assert s[k] == v, (k, s)
    check_across_face(o.esdf_layer(), parents)
