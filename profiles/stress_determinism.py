"""Race detector: the same sequence integrated repeatedly must give bit-identical layers every time.
Prints the first frame at which a repetition diverges from the first one (TSDF / ESDF checksums per frame)."""
import os, sys, zlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import __graft_entry__ as g
g.build()
import isaac_ros_nvblox_b200 as nvb
from isaac_ros_nvblox_b200 import synthetic as syn
from helpers import ESDF_FIELDS, layer_checksum

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
nframes = int(sys.argv[2]) if len(sys.argv) > 2 else 20
mode = sys.argv[3] if len(sys.argv) > 3 else "sync"
cs = syn.PinholeCamera(); cam = nvb.Camera(cs.fu, cs.fv, cs.cu, cs.cv, cs.width, cs.height)
frames = syn.make_sequence(syn.sphere_in_box(), cs, syn.circle_trajectory(80)[::4][:nframes])
ref = None
bad = 0
for rep in range(reps):
    m = nvb.Mapper(0.05)
    sums = []
    for i, (d, T) in enumerate(frames):
        if mode == "sync":
            m.integrate_depth(d, T, cam); m.update_esdf()
        else:
            m.integrate_depth_async(d, T, cam); m.update_esdf(sync=False)
        if mode == "sync" or i == len(frames) - 1:
            m.synchronize()
            t = m.tsdf_layer().as_dict(); e = m.esdf_layer().as_dict()
            sums.append((i, layer_checksum(t, ("distance", "weight")), layer_checksum(e, ESDF_FIELDS), len(t), len(e)))
    if ref is None:
        ref = sums
    else:
        for a, b in zip(ref, sums):
            if a != b:
                bad += 1
                print("rep %d diverges at frame %d: ref %s got %s" % (rep, a[0], a[1:], b[1:]))
                break
    m.close()
print("stress", mode, "reps", reps, "frames", nframes, "divergent reps:", bad)
