"""Deterministic launch sequence for ncu: N frames of the C2 sequence (device-resident depth),
integrateDepth + updateEsdf per frame. Used by the commands in profiles/README.md."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import __graft_entry__ as g
    g.build()
    import isaac_ros_nvblox_b200 as nvb
    from isaac_ros_nvblox_b200 import synthetic as syn
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    voxel = float(sys.argv[2]) if len(sys.argv) > 2 else 0.05
    cs = syn.PinholeCamera()
    cam = nvb.Camera(cs.fu, cs.fv, cs.cu, cs.cv, cs.width, cs.height)
    seq = syn.make_sequence(syn.sphere_in_box(), cs, syn.circle_trajectory(80)[:frames])
    depth = torch.from_numpy(np.stack([d for d, _ in seq])).cuda()
    m = nvb.Mapper(voxel)
    for i, (_, T) in enumerate(seq):
        m.integrate_depth_device(depth[i].data_ptr(), 480, 640, T, cam)
        m.update_esdf(sync=False)
    m.synchronize()
    print("frames", frames, "tsdf blocks", m.tsdf_layer().num_blocks(), "launches", m.kernel_launches())
    m.close()


if __name__ == "__main__":
    main()
