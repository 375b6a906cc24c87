import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import isaac_ros_nvblox_b200 as nvb
from isaac_ros_nvblox_b200 import multi_gpu, synthetic as syn
from helpers import cameras
m = nvb.Mapper(0.05)
cs, cam, ocam = cameras(320, 240)
frames = syn.make_sequence(syn.sphere_in_box(), cs, syn.circle_trajectory(40)[:6])
bm = multi_gpu.BatchMerger(m, cap_entries=8192)
per_frame = []
for i, (depth, T) in enumerate(frames):
    per_frame.append(m.integrate_depth(depth, T, cam))
    bm.append_last_frame()
    m.synchronize()
    print("frame", i, "list", len(per_frame[-1]), "segment count", int(bm.local[bm.cur][0].item()))
    if i % 3 == 2:
        seg = bm.local[bm.cur].clone()
        bm.merge(timed=True)
        got = bm.result().cpu().numpy()
        u = np.unique(np.concatenate(per_frame[i - 2:i + 1]), axis=0)
        want = u[np.lexsort((u[:, 0], u[:, 1], u[:, 2]))]
        n = int(seg[0].item())
        ent = seg[1:1 + 3 * n].cpu().numpy().reshape(-1, 3)
        print("batch end", i, "got", got.shape, "want", want.shape, "segment entries == concat:", np.array_equal(ent, np.concatenate(per_frame[i - 2:i + 1])))
        if got.shape == want.shape:
            print("equal", np.array_equal(got, want))
        else:
            gs, ws = set(map(tuple, got)), set(map(tuple, want))
            print("missing", len(ws - gs), "extra", len(gs - ws), list(ws - gs)[:3], list(gs - ws)[:3])
