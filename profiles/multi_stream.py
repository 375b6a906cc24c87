"""Aggregate throughput of B independent camera streams (one Mapper = own map + own CUDA streams each)
driven round-robin from one host thread on ONE GPU: do the per-mapper kernel chains overlap?"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
g.build()
import isaac_ros_nvblox_b200 as nvb
from isaac_ros_nvblox_b200 import synthetic as syn

cs = syn.PinholeCamera(); cam = nvb.Camera(cs.fu, cs.fv, cs.cu, cs.cv, cs.width, cs.height)
F = 80
poses = syn.circle_trajectory(80)
for B in (1, 2, 4, 8):
    seqs = []
    for b in range(B):
        sh = (b * 80) // B
        p = poses[sh:] + poses[:sh]
        fr = syn.make_sequence(syn.sphere_in_box(), cs, p[:F])
        seqs.append((torch.from_numpy(np.stack([d for d, _ in fr])).cuda(), [T for _, T in fr]))
    ms = [nvb.Mapper(0.05) for _ in range(B)]
    def step():
        for m in ms:
            m.clear()
        for i in range(F):
            for b, m in enumerate(ms):
                m.integrate_depth_device(seqs[b][0][i].data_ptr(), 480, 640, seqs[b][1][i], cam)
                m.update_esdf(sync=False)
        for m in ms:
            m.synchronize()
    step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("streams %d: %.0f frames/s aggregate (%.0f per stream)" % (B, B * F * reps / dt, F * reps / dt))
    for m in ms:
        m.close()
