"""Prints the ESDF wavefront's time split (barrier / axis / sweep) per frame for a few frames."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
g.build()
import isaac_ros_nvblox_b200 as nvb
from isaac_ros_nvblox_b200 import synthetic as syn
cs = syn.PinholeCamera(); cam = nvb.Camera(cs.fu, cs.fv, cs.cu, cs.cv, cs.width, cs.height)
seq = syn.make_sequence(syn.sphere_in_box(), cs, syn.circle_trajectory(80)[:10])
depth = torch.from_numpy(np.stack([d for d, _ in seq])).cuda()
m = nvb.Mapper(0.05)
for i, (_, T) in enumerate(seq):
    m.integrate_depth_device(depth[i].data_ptr(), 480, 640, T, cam)
    m.update_esdf(sync=False)
    m.synchronize()
    sp = m.esdf_time_split()
    print(i, m.esdf_integrator().last_stats(), sp)
    if i >= 8:
        import ctypes as C
        n = int(sp["barriers"])
        arr = (C.c_int64 * 4000)()
        m._L.nvb_mapper_debug_phase_max(m._h, arr, 4000)
        print("   per-phase slowest-CTA work (ns):", [int(arr[q]) for q in range(min(n, 80))])
        print("   GES cta0/group0 (gather cycles, emulate cycles, sweep+store cycles, candidates, changed):", [int(arr[q]) for q in range(3990, 3995)])
        print("   sweep phases (total, scan, load+sweep+store, max k):", [(int(arr[q]), int(arr[1000+q]), int(arr[2000+q]), int(arr[3000+q])) for q in range(min(n, 80)) if arr[3000+q] or arr[2000+q]])
