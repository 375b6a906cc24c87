"""The oracle reproduces the committed golden fixture (guards against oracle drift). CPU only."""
import os

import numpy as np

from helpers import ESDF_FIELDS, layer_checksum
from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_oracle_reproduces_golden_fixture():
    g = np.load(os.path.join(GOLDEN, "c2_small.npz"))
    c = g["cam"]
    cam = orc.Camera(float(c[0]), float(c[1]), float(c[2]), float(c[3]), int(c[4]), int(c[5]))
    m = orc.OracleMap(float(g["voxel_size"]))
    for i in range(len(g["depth"])):
        b = m.integrate_depth(g["depth"][i], g["poses"][i], cam)
        assert np.array_equal(b, g["blocks_%d" % i])
        m.integrate_esdf(b)
    assert layer_checksum(m.tsdf_layer(), ("distance", "weight")) == int(g["tsdf_checksum"])
    assert layer_checksum(m.esdf_layer(), ESDF_FIELDS) == int(g["esdf_checksum"])


def test_oracle_thread_count_does_not_change_results():
    """OpenMP is only used where the reference's kernels are race-free; 1 thread == N threads."""
    from isaac_ros_nvblox_b200 import synthetic as syn
    cs = syn.PinholeCamera(75.0, 75.0, 80.0, 60.0, 160, 120)
    cam = orc.Camera(75.0, 75.0, 80.0, 60.0, 160, 120)
    frames = syn.make_sequence(syn.box_with_cube(), cs, syn.circle_trajectory(16)[:3])
    sums = []
    n_threads = orc.num_threads()
    for nt in (1, max(2, n_threads)):
        orc.set_num_threads(nt)
        m = orc.OracleMap(0.1)
        for d, T in frames:
            m.integrate_esdf(m.integrate_depth(d, T, cam))
        sums.append((layer_checksum(m.tsdf_layer(), ("distance", "weight")), layer_checksum(m.esdf_layer(), ESDF_FIELDS)))
    orc.set_num_threads(n_threads)
    assert sums[0] == sums[1]


def test_oracle_reproduces_the_f_rows_fixture():
    """tests/golden/f_rows_small.npz (make_golden_f_rows.py): occupancy, colour, decay, freespace, 2-D slice, mark-free."""
    from golden_f_rows import run_oracle
    g = np.load(os.path.join(GOLDEN, "c2_small.npz"))
    want = np.load(os.path.join(GOLDEN, "f_rows_small.npz"))
    got = run_oracle(g)
    assert set(got) == set(want.files)
    for k, v in got.items():
        assert int(want[k]) == v, k
