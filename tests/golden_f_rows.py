"""The SURVEY.md 8(f) rows on the inputs of tests/golden/c2_small.npz (4 frames, 160x120, 10 cm voxels): one scenario, run on the
oracle (run_oracle) and on the CUDA path (run_gpu); both return the same dictionary of order-independent layer checksums.
tests/golden/make_golden_f_rows.py stores the oracle's in tests/golden/f_rows_small.npz."""
import numpy as np

from helpers import ESDF_FIELDS, layer_checksum, textured_image

FREESPACE_FIELDS = ("last_occupied_timestamp_ms", "consecutive_occupancy_duration_ms", "is_high_confidence_freespace")
SLICE = dict(z_min_m=0.3, z_max_m=1.7, z_output_m=1.0)
MARK_CENTER, MARK_RADIUS = (0.4, -0.3, 1.1), 1.3


def _occ_dict(layer):
    return {k: {"log_odds": np.asarray(v["log_odds"] if getattr(v.dtype, "names", None) else v, np.float32)} for k, v in layer.items()}


def _images(g):
    rows, cols = g["depth"].shape[1:]
    return [textured_image(rows, cols, seed=100 + i) for i in range(len(g["depth"]))]


def run_oracle(g):
    from oracle import oracle as orc
    c = g["cam"]
    cam = orc.Camera(float(c[0]), float(c[1]), float(c[2]), float(c[3]), int(c[4]), int(c[5]))
    voxel, depth, poses, imgs = float(g["voxel_size"]), g["depth"], g["poses"], _images(g)
    out = {}
    # occupancy + ESDF from occupancy
    m, tp = orc.OracleMap(voxel), orc.default_tsdf_params()
    for i in range(len(depth)):
        b = m.integrate_occupancy(depth[i], poses[i], cam, tp)
        m.integrate_esdf_occupancy(b if i else m.occupancy_block_indices())
    out["occupancy"] = layer_checksum(_occ_dict(m.occupancy_layer()), ("log_odds",))
    out["occupancy_esdf"] = layer_checksum(m.esdf_layer(), ESDF_FIELDS)
    # TSDF + colour + ESDF, then two decays with deallocation
    m = orc.OracleMap(voxel)
    for i in range(len(depth)):
        b = m.integrate_depth(depth[i], poses[i], cam)
        m.integrate_color(imgs[i], poses[i], cam)
        m.integrate_esdf(b if i else m.tsdf_block_indices())
    out["color"] = layer_checksum(m.color_layer(), ("color", "weight"))
    dp = orc.default_tsdf_decay_params(decay_factor=0.5)
    out["decay_removed"] = int(sum(len(m.decay_tsdf(dp)) for _ in range(2)))
    out["decay_tsdf"] = layer_checksum(m.tsdf_layer(), ("distance", "weight"))
    out["decay_color"] = layer_checksum(m.color_layer(), ("color", "weight"))
    out["decay_esdf"] = layer_checksum(m.esdf_layer(), ESDF_FIELDS)
    # freespace
    m, fp = orc.OracleMap(voxel), orc.default_freespace_params(min_duration_since_occupied_for_freespace_ms=200)
    for i in range(len(depth)):
        b = m.integrate_depth(depth[i], poses[i], cam)
        m.update_freespace(m.tsdf_block_indices() if i == 0 else b, 1000 + 150 * i, fp)
    out["freespace"] = layer_checksum(m.freespace_layer(), FREESPACE_FIELDS)
    # 2-D ESDF slice
    m = orc.OracleMap(voxel)
    for i in range(len(depth)):
        b = m.integrate_depth(depth[i], poses[i], cam)
        m.integrate_esdf_slice(b if i else m.tsdf_block_indices(), **SLICE)
    out["slice_esdf"] = layer_checksum(m.esdf_layer(), ESDF_FIELDS)
    # markUnobservedTsdfFreeInsideRadius
    m = orc.OracleMap(voxel)
    m.integrate_depth(depth[0], poses[0], cam)
    out["mark_free_blocks"] = int(len(m.mark_unobserved_free_inside_radius(MARK_CENTER, MARK_RADIUS)))
    out["mark_free_tsdf"] = layer_checksum(m.tsdf_layer(), ("distance", "weight"))
    return out


def run_gpu(g):
    import isaac_ros_nvblox_b200 as nvb
    c = g["cam"]
    cam = nvb.Camera(float(c[0]), float(c[1]), float(c[2]), float(c[3]), int(c[4]), int(c[5]))
    voxel, depth, poses, imgs = float(g["voxel_size"]), g["depth"], g["poses"], _images(g)
    out = {}
    m = nvb.Mapper(voxel, projective_layer_type=nvb.ProjectiveLayerType.kOccupancy)
    for i in range(len(depth)):
        m.integrate_depth(depth[i], poses[i], cam)
        m.update_esdf()
    out["occupancy"] = layer_checksum(_occ_dict(m.occupancy_layer().as_dict()), ("log_odds",))
    out["occupancy_esdf"] = layer_checksum(m.esdf_layer().as_dict(), ESDF_FIELDS)
    m.close()
    m = nvb.Mapper(voxel)
    for i in range(len(depth)):
        m.integrate_depth(depth[i], poses[i], cam)
        m.integrate_color(imgs[i], poses[i], cam)
        m.update_esdf()
    out["color"] = layer_checksum(m.color_layer().as_dict(), ("color", "weight"))
    m.tsdf_decay_integrator().params(decay_factor=0.5)
    out["decay_removed"] = int(sum(len(m.decay()) for _ in range(2)))
    out["decay_tsdf"] = layer_checksum(m.tsdf_layer().as_dict(), ("distance", "weight"))
    out["decay_color"] = layer_checksum(m.color_layer().as_dict(), ("color", "weight"))
    out["decay_esdf"] = layer_checksum(m.esdf_layer().as_dict(), ESDF_FIELDS)
    m.close()
    m = nvb.Mapper(voxel, projective_layer_type=nvb.ProjectiveLayerType.kTsdfWithFreespace)
    m.freespace_integrator().params(min_duration_since_occupied_for_freespace_ms=200)
    for i in range(len(depth)):
        m.integrate_depth(depth[i], poses[i], cam)
        m.update_freespace(1000 + 150 * i)
    out["freespace"] = layer_checksum(m.freespace_layer().as_dict(), FREESPACE_FIELDS)
    m.close()
    m = nvb.Mapper(voxel)
    m.esdf_integrator().slice_params(slice_min_height_m=SLICE["z_min_m"], slice_max_height_m=SLICE["z_max_m"],
                                     slice_height_m=SLICE["z_output_m"])
    for i in range(len(depth)):
        m.integrate_depth(depth[i], poses[i], cam)
        m.update_esdf_slice()
    out["slice_esdf"] = layer_checksum(m.esdf_layer().as_dict(), ESDF_FIELDS)
    m.close()
    m = nvb.Mapper(voxel)
    m.integrate_depth(depth[0], poses[0], cam)
    out["mark_free_blocks"] = int(len(m.mark_unobserved_tsdf_free_inside_radius(MARK_CENTER, MARK_RADIUS)))
    out["mark_free_tsdf"] = layer_checksum(m.tsdf_layer().as_dict(), ("distance", "weight"))
    m.close()
    return out
