"""Per-ring time split of the ESDF wavefront kernels on a few frames of the bench sequence.
usage: python profiles/wavex_split.py [mode=3] [first_frame=30] [frames=3]"""
import os, sys
import ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
g.build()
import isaac_ros_nvblox_b200 as nvb
import bench
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 3
first = int(sys.argv[2]) if len(sys.argv) > 2 else 30
nshow = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cs, seq = bench.make_frames(first + nshow, 0, 1)
cam = nvb.Camera(cs.fu, cs.fv, cs.cu, cs.cv, cs.width, cs.height)
depth = torch.from_numpy(np.stack([d for d, _ in seq])).cuda()
m = nvb.Mapper(0.05, esdf_persistent=mode)
m.enable_profiling(True)
for i, (_, T) in enumerate(seq):
    m.integrate_depth_device(depth[i].data_ptr(), 480, 640, T, cam)
    m.update_esdf(sync=False)
    m.synchronize()
    if i < first:
        m.stage_times(reset=True)
        continue
    sp = m.esdf_time_split()
    st = m.stage_times(reset=True)
    print("frame", i, m.esdf_integrator().last_stats(), sp, {k: round(v[0] * 1e3, 1) for k, v in st.items()})
    n = int(sp["barriers"])
    arr = (C.c_int64 * 4000)()
    m._L.nvb_mapper_debug_phase_max(m._h, arr, 4000)
    print("   per barrier (K, M, slowest-CTA work ns, CTA0 work ns):", [(int(arr[1000 + q]), int(arr[2000 + q]), int(arr[q]), int(arr[3000 + q])) for q in range(min(n, 60))])
    print("   cta0/group0 cycles (record, stamps+own, halo, replay, sweep, store+register), candidates, changed:", [int(arr[q]) for q in range(3990, 3998)])
