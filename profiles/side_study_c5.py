"""Side study, BASELINE config 5 shape: dynamic scene (a sphere moving through the sphere-in-box room) with its image-space
mask, 640x480, 5 cm voxels, occupancy mapper: ProjectiveOccupancyIntegrator (mask non-inverted: only the masked pixels are
active, the dynamic mapper's convention) + ESDF from occupancy every frame + OccupancyDecayIntegrator (exclude last view)
every 8 frames. Prints one JSON line: frames/s with device-resident inputs (CUDA events), the stage split, and the CPU port
on the first frames. Not the headline metric (bench.py)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import __graft_entry__ as g
    g.build()
    import isaac_ros_nvblox_b200 as nvb
    from isaac_ros_nvblox_b200 import synthetic as syn
    from oracle import oracle as orc
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 80
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    decay_every = 8
    cs = syn.PinholeCamera()
    cam = nvb.Camera(cs.fu, cs.fv, cs.cu, cs.cv, cs.width, cs.height)
    seq = syn.moving_sphere_sequence(cs, syn.circle_trajectory(F), step_m=0.05)
    depth = torch.from_numpy(np.stack([d for d, _, _ in seq])).cuda()
    mask = torch.from_numpy(np.stack([k for _, _, k in seq])).cuda()
    poses = [T for _, T, _ in seq]
    m = nvb.Mapper(0.05, projective_layer_type=nvb.ProjectiveLayerType.kOccupancy, keep_last_view=True)
    stream = torch.cuda.ExternalStream(m.cuda_stream())

    def step():
        m.clear()
        removed = 0
        for i in range(F):
            m.integrate_depth_device(depth[i].data_ptr(), 480, 640, poses[i], cam, mask_ptr=mask[i].data_ptr(), mask_mode=0)
            m.update_esdf(sync=False)
            if (i + 1) % decay_every == 0:
                removed += len(m.decay_exclude_last_view())
        return removed

    for _ in range(2):
        step()
    m.synchronize()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        removed = step()
    m.join_streams()
    ev1.record(stream)
    m.synchronize()
    ms = ev0.elapsed_time(ev1)
    m.enable_profiling(True)
    step()
    m.synchronize()
    stages = m.stage_times(reset=True)
    m.enable_profiling(False)
    blocks = m.occupancy_layer().num_blocks()
    # CPU port on the first frames
    n_cpu = min(12, F)
    orc.set_num_threads(max(1, min(os.cpu_count() or 1, 16)))
    ocam = orc.Camera(cs.fu, cs.fv, cs.cu, cs.cv, cs.width, cs.height)
    o = orc.OracleMap(0.05)
    tp, dp = orc.default_tsdf_params(), orc.default_occupancy_decay_params()
    t0 = time.perf_counter()
    for i in range(n_cpu):
        d, T, k = seq[i]
        b = o.integrate_occupancy(d, T, ocam, tp, mask=k, mask_mode=0)
        o.integrate_esdf_occupancy(b if i > 0 else o.occupancy_block_indices())
        if (i + 1) % decay_every == 0:
            o.decay_occupancy(dp, depth=d, T_L_C=T, cam=ocam, max_view_distance_m=tp.max_integration_distance_m,
                              truncation_distance_m=tp.truncation_distance_vox * 0.05)
    dt = time.perf_counter() - t0
    print(json.dumps({
        "study": "C5 dynamic scene: occupancy + mask + ESDF per frame, occupancy decay (exclude last view) every %d frames" % decay_every,
        "value": F * steps / (ms * 1e-3), "unit": "frames/s", "frames": F, "steps": steps, "ms_per_step": ms / steps,
        "occupancy_blocks_at_end": blocks, "blocks_deallocated_per_step": removed,
        "stages_ms_per_frame": {k: (ms_ / calls if calls else 0.0) for k, (ms_, calls) in stages.items()},
        "cpu_port": {"value": n_cpu / dt, "unit": "frames/s", "cores": orc.num_threads(), "sample": "first %d frames" % n_cpu}}))


if __name__ == "__main__":
    main()
