"""Side study: the colour kernels on the bench's sequence (640x480, 5 cm voxels): depth + colour + ESDF per frame.
Run plain for a timing line, or under `ncu --metrics gpu__time_duration.sum -k regex:color|sphereTrace` for the launch list."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import isaac_ros_nvblox_b200 as nvb  # noqa: E402
from isaac_ros_nvblox_b200 import synthetic as syn  # noqa: E402

F = int(os.environ.get("FRAMES", "40"))
cs = syn.PinholeCamera()
frames = syn.make_sequence(syn.sphere_in_box(), cs, syn.circle_trajectory(80)[:F])
cam = nvb.Camera(cs.fu, cs.fv, cs.cu, cs.cv, cs.width, cs.height)
depth = torch.from_numpy(np.stack([d for d, _ in frames])).cuda()
yy, xx = np.mgrid[0:480, 0:640]
base = np.stack([xx * 255 // 639, yy * 255 // 479, ((xx // 16 + yy // 16) % 2) * 200 + 20], axis=-1).astype(np.uint8)
color = torch.from_numpy(np.stack([np.roll(base, 7 * i, axis=1) for i in range(F)])).cuda()
m = nvb.Mapper(0.05)
stream = torch.cuda.ExternalStream(m.cuda_stream())


def step(with_color):
    m.clear()
    for i, (_, T) in enumerate(frames):
        m.integrate_depth_device(depth[i].data_ptr(), 480, 640, T, cam)
        if with_color:
            m.integrate_color_device(color[i].data_ptr(), 480, 640, T, cam)
        m.update_esdf(sync=False)


out = {}
for name, wc in (("depth_esdf", False), ("depth_color_esdf", True)):
    for _ in range(3):
        step(wc)
    m.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(3):
        step(wc)
    m.join_streams()
    e1.record(stream)
    m.synchronize()
    out[name] = {"frames_per_s": 3 * F / (e0.elapsed_time(e1) * 1e-3), "ms_per_frame": e0.elapsed_time(e1) / (3 * F)}
out["color_blocks"] = m.color_layer().num_blocks()
out["last_color_frame_blocks"] = int(len(m.last_color_blocks()))
print(json.dumps(out))
