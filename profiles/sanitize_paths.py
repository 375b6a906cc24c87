"""Small workload that touches every kernel of the library once (for compute-sanitizer): TSDF / occupancy / freespace mappers,
3-D and 2-D ESDF (four-phase, host loop and gather-replay wavefronts), decay with deallocation and slot reuse, the slicer,
explicit block lists, layer read-back and growth, colour integration (with distortion and a mask) and the sphere tracer."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import __graft_entry__ as g
    g.build()
    import isaac_ros_nvblox_b200 as nvb
    from isaac_ros_nvblox_b200 import synthetic as syn
    cs = syn.PinholeCamera(75.0, 75.0, 80.0, 60.0, 160, 120)
    cam = nvb.Camera(75.0, 75.0, 80.0, 60.0, 160, 120)
    dcam = nvb.Camera(75.0, 75.0, 80.0, 60.0, 160, 120, radial=(-0.05, 0.01, 0, 0.02, 0, 0), tangential=(0.001, -0.0005))
    seq = syn.moving_sphere_sequence(cs, syn.circle_trajectory(16)[:4], step_m=0.3)
    voxel = 0.1
    total = 0
    for mode in (1, 0, 2):
        for ltype in (nvb.ProjectiveLayerType.kTsdf, nvb.ProjectiveLayerType.kOccupancy, nvb.ProjectiveLayerType.kTsdfWithFreespace):
            m = nvb.Mapper(voxel, esdf_persistent=mode, projective_layer_type=ltype, keep_last_view=True,
                           tsdf_capacity_blocks=256, esdf_capacity_blocks=256)
            for i, (d, T, k) in enumerate(seq):
                m.integrate_depth(d, T, dcam if i == 1 else cam, mask=k if i == 2 else None, mask_mode=1)
                if ltype == nvb.ProjectiveLayerType.kTsdfWithFreespace:
                    m.update_freespace(1000 + 400 * i, depth=d, T_L_C=T, camera=cam)
                if ltype != nvb.ProjectiveLayerType.kOccupancy:
                    img = np.full((120, 160, 3), 40 * i + 10, np.uint8)
                    m.integrate_color(img, T, dcam if i == 1 else cam, mask=k if i == 2 else None)
                m.update_esdf()
            m.decay_exclude_last_view()
            m.decay()
            d, T, _ = seq[0]
            b = m.integrate_depth(d, T, cam)
            m.update_esdf()
            m.esdf_integrator().integrate_blocks(b[:10])
            if ltype != nvb.ProjectiveLayerType.kOccupancy:
                m.integrate_color(np.zeros((120, 160, 3), np.uint8), T, cam)  # reuses deallocated colour slots
                m.color_integrator().render_depth(T, cam, 0.4, ray_subsampling_factor=2)
                m.color_layer().as_dict()
            nvb.EsdfSlicer(m).slice_layer_to_distance_image(1.0, with_occupancy_grid=True)
            total += m.esdf_layer().num_blocks()
            m.esdf_layer().as_dict()
            m.clear()
            for i, (d, T, k) in enumerate(seq[:2]):
                m.integrate_depth(d, T, cam)
                m.update_esdf_slice()
            m.decay()
            nvb.ViewCalculator(m).get_blocks_in_image_view_raycast(d, T, cam, 0.8, 0.4, 7.0)
            m.close()
    print("sanitize workload ok", total)


if __name__ == "__main__":
    main()
