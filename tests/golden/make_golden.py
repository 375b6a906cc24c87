"""Generates tests/golden/c2_small.npz from the CPU oracle (the reference itself cannot be built
or imported in this environment -- see DESIGN.md "Oracle"). Run from the repo root:
    python tests/golden/make_golden.py
Inputs: 4 frames of the sphere-in-box scene on the reference's test circle, 160x120, 10 cm voxels.
Stored: the inputs, the per-frame updated_blocks lists, and order-independent checksums of the
final TSDF and ESDF layers."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import ESDF_FIELDS, layer_checksum  # noqa: E402
from isaac_ros_nvblox_b200 import synthetic as syn  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def main():
    cs = syn.PinholeCamera(75.0, 75.0, 80.0, 60.0, 160, 120)
    cam = orc.Camera(75.0, 75.0, 80.0, 60.0, 160, 120)
    frames = syn.make_sequence(syn.sphere_in_box(), cs, syn.circle_trajectory(16)[:4], noise_sigma_rel=0.005,
                               dropout=0.05, seed=7)
    voxel = 0.1
    m = orc.OracleMap(voxel)
    out = {"voxel_size": np.float32(voxel), "cam": np.array([75.0, 75.0, 80.0, 60.0, 160, 120], np.float32),
           "depth": np.stack([d for d, _ in frames]), "poses": np.stack([T for _, T in frames])}
    for i, (d, T) in enumerate(frames):
        b = m.integrate_depth(d, T, cam)
        m.integrate_esdf(b)
        out["blocks_%d" % i] = b
    out["tsdf_checksum"] = np.int64(layer_checksum(m.tsdf_layer(), ("distance", "weight")))
    out["esdf_checksum"] = np.int64(layer_checksum(m.esdf_layer(), ESDF_FIELDS))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "c2_small.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
