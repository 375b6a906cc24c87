"""ctypes loader for the CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module (see oracle/nvblox_oracle.h).
Parity unpinned against a reference binary (none can be built here): the oracle is pinned to the reference's own tests.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libnvblox_oracle.so")

TSDF_VOXEL_DTYPE = np.dtype([("distance", "<f4"), ("weight", "<f4")])
ESDF_VOXEL_DTYPE = np.dtype(
    [
        ("squared_distance_vox", "<f4"),
        ("parent_direction", "<i4", (3,)),
        ("is_inside", "u1"),
        ("observed", "u1"),
        ("is_site", "u1"),
        ("pad", "u1"),
    ]
)
assert TSDF_VOXEL_DTYPE.itemsize == 8 and ESDF_VOXEL_DTYPE.itemsize == 20

WEIGHT_CONSTANT = 0
WEIGHT_CONSTANT_DROPOFF = 1
WEIGHT_INVERSE_SQUARE = 2
WEIGHT_INVERSE_SQUARE_DROPOFF = 3
WEIGHT_INVERSE_SQUARE_TSDF_DISTANCE_PENALTY = 4
WEIGHT_LINEAR_WITH_MAX = 5


class Camera(C.Structure):
    _fields_ = [("fu", C.c_float), ("fv", C.c_float), ("cu", C.c_float), ("cv", C.c_float),
                ("width", C.c_int32), ("height", C.c_int32), ("has_distortion", C.c_int32),
                ("k1", C.c_float), ("k2", C.c_float), ("k3", C.c_float), ("k4", C.c_float),
                ("k5", C.c_float), ("k6", C.c_float), ("p1", C.c_float), ("p2", C.c_float)]

    def with_distortion(self, k=(0, 0, 0, 0, 0, 0), p=(0, 0)):
        """RadialTangentialDistortionParams{radial k1..k6, tangential p1, p2}."""
        c = Camera(self.fu, self.fv, self.cu, self.cv, self.width, self.height, 1, *[float(v) for v in k],
                   *[float(v) for v in p])
        return c


class TsdfParams(C.Structure):
    _fields_ = [("truncation_distance_vox", C.c_float),
                ("max_integration_distance_m", C.c_float),
                ("max_weight", C.c_float),
                ("invalid_depth_decay_factor", C.c_float),
                ("weighting_type", C.c_int32),
                ("raycast_subsampling", C.c_int32),
                ("workspace_bounds_type", C.c_int32),
                ("workspace_min", C.c_float * 3),
                ("workspace_max", C.c_float * 3)]


class EsdfParams(C.Structure):
    _fields_ = [("max_esdf_distance_m", C.c_float),
                ("max_site_distance_vox", C.c_float),
                ("min_weight", C.c_float),
                ("occupied_threshold", C.c_float)]


class OccupancyParams(C.Structure):
    _fields_ = [("free_region_occupancy_probability", C.c_float),
                ("occupied_region_occupancy_probability", C.c_float),
                ("unobserved_region_occupancy_probability", C.c_float),
                ("occupied_region_half_width_m", C.c_float)]


FREESPACE_VOXEL_DTYPE = np.dtype([("last_occupied_timestamp_ms", "<i8"), ("consecutive_occupancy_duration_ms", "<i8"),
                                  ("is_high_confidence_freespace", "u1"), ("pad", "u1", (7,))])
assert FREESPACE_VOXEL_DTYPE.itemsize == 24


class FreespaceParams(C.Structure):
    _fields_ = [("max_tsdf_distance_for_occupancy_m", C.c_float),
                ("max_unobserved_to_keep_consecutive_occupancy_ms", C.c_int64),
                ("min_duration_since_occupied_for_freespace_ms", C.c_int64),
                ("min_consecutive_occupancy_duration_for_reset_ms", C.c_int64),
                ("check_neighborhood", C.c_int32), ("initialize_to_high_confidence_freespace", C.c_int32)]


COLOR_VOXEL_DTYPE = np.dtype([("color", "u1", (3,)), ("pad", "u1"), ("weight", "<f4")])
assert COLOR_VOXEL_DTYPE.itemsize == 8


class ColorParams(C.Structure):
    _fields_ = [("max_integration_distance_m", C.c_float), ("truncation_distance_vox", C.c_float),
                ("max_weight", C.c_float), ("measurement_weight", C.c_float),
                ("sphere_tracing_ray_subsampling_factor", C.c_int32), ("sphere_tracer_maximum_steps", C.c_int32),
                ("sphere_tracer_maximum_ray_length_m", C.c_float), ("sphere_tracer_surface_distance_epsilon_vox", C.c_float),
                ("workspace_bounds_type", C.c_int32), ("workspace_min", C.c_float * 3), ("workspace_max", C.c_float * 3)]


class MeshParams(C.Structure):
    """OrMeshParams: mesh_integrator_params.h:22-27, MeshIntegrator::cutoff_distance_vox_ (mesh_integrator.h:129)."""
    _fields_ = [("min_weight", C.c_float), ("weld_vertices", C.c_int32), ("cutoff_distance_vox", C.c_float)]


class TsdfDecayParams(C.Structure):
    _fields_ = [("decay_factor", C.c_float), ("decayed_weight_threshold", C.c_float),
                ("set_free_distance_on_decayed", C.c_int32), ("free_distance_vox", C.c_float),
                ("deallocate_decayed_blocks", C.c_int32)]


class OccupancyDecayParams(C.Structure):
    _fields_ = [("free_region_decay_probability", C.c_float), ("occupied_region_decay_probability", C.c_float),
                ("decay_to_probability", C.c_float), ("deallocate_decayed_blocks", C.c_int32)]


class DecayExclusion(C.Structure):
    _fields_ = [("excluded_blocks_xyz", C.POINTER(C.c_int32)), ("num_excluded_blocks", C.c_int32),
                ("has_exclusion_sphere", C.c_int32), ("exclusion_center", C.c_float * 3),
                ("exclusion_radius_m", C.c_float)]


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc, no FMA contraction)."""
    if force or not os.path.exists(_LIB_PATH) or (
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "nvblox_oracle.c"))):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    fp = C.POINTER(C.c_float)
    ip = C.POINTER(C.c_int32)
    u8p = C.POINTER(C.c_uint8)
    vp = C.c_void_p
    L.or_default_tsdf_params.argtypes = [C.POINTER(TsdfParams)]
    L.or_default_esdf_params.argtypes = [C.POINTER(EsdfParams)]
    L.or_map_create.argtypes = [C.c_float]
    L.or_map_create.restype = vp
    L.or_map_destroy.argtypes = [vp]
    L.or_map_clear.argtypes = [vp]
    L.or_raycast_cells.argtypes = [fp, fp, C.c_float, ip, C.c_int32]
    L.or_raycast_cells.restype = C.c_int32
    L.or_view_raycast.argtypes = [fp, C.c_int32, C.c_int32, fp, C.POINTER(Camera), C.c_float,
                                  C.c_float, C.POINTER(TsdfParams), ip, C.c_int32]
    L.or_view_raycast.restype = C.c_int32
    L.or_tsdf_integrate.argtypes = [vp, fp, u8p, C.c_int32, C.c_int32, C.c_int32, fp,
                                    C.POINTER(Camera), C.POINTER(TsdfParams), ip, C.c_int32]
    L.or_tsdf_integrate.restype = C.c_int32
    L.or_tsdf_integrate_blocks.argtypes = [vp, fp, u8p, C.c_int32, C.c_int32, C.c_int32, fp,
                                           C.POINTER(Camera), C.POINTER(TsdfParams), ip, C.c_int32]
    L.or_esdf_integrate.argtypes = [vp, ip, C.c_int32, C.POINTER(EsdfParams)]
    L.or_esdf_last_stats.argtypes = [vp, C.POINTER(C.c_int64)]
    L.or_default_occupancy_params.argtypes = [C.POINTER(OccupancyParams)]
    L.or_occupancy_integrate.argtypes = [vp, fp, u8p, C.c_int32, C.c_int32, C.c_int32, fp, C.POINTER(Camera),
                                         C.POINTER(TsdfParams), C.POINTER(OccupancyParams), ip, C.c_int32]
    L.or_occupancy_integrate.restype = C.c_int32
    L.or_esdf_integrate_occupancy.argtypes = [vp, ip, C.c_int32, C.POINTER(EsdfParams)]
    L.or_occupancy_num_blocks.argtypes = [vp]
    L.or_occupancy_num_blocks.restype = C.c_int32
    L.or_occupancy_block_indices.argtypes = [vp, ip, C.c_int32]
    L.or_occupancy_block_indices.restype = C.c_int32
    L.or_occupancy_get_block.argtypes = [vp, ip, vp]
    L.or_occupancy_get_block.restype = C.c_int32
    for name in ("or_tsdf_num_blocks", "or_esdf_num_blocks"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = C.c_int32
    for name in ("or_tsdf_block_indices", "or_esdf_block_indices"):
        getattr(L, name).argtypes = [vp, ip, C.c_int32]
        getattr(L, name).restype = C.c_int32
    L.or_tsdf_get_block.argtypes = [vp, ip, vp]
    L.or_tsdf_get_block.restype = C.c_int32
    L.or_esdf_get_block.argtypes = [vp, ip, vp]
    L.or_esdf_get_block.restype = C.c_int32
    L.or_tsdf_set_block.argtypes = [vp, ip, vp]
    L.or_map_cache_last_viewpoint.argtypes = [vp, C.c_int32]
    L.or_depth_dilate_invalid.argtypes = [C.POINTER(C.c_float), C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, C.POINTER(C.c_float)]
    L.or_depth_dilate_invalid.restype = None
    L.or_mesh_integrate_blocks.argtypes = [vp, ip, C.c_int32, C.POINTER(MeshParams)]
    L.or_mesh_integrate_blocks.restype = None
    L.or_mesh_update_color.argtypes = [vp, ip, C.c_int32]
    L.or_mesh_update_color.restype = None
    L.or_mesh_num_blocks.argtypes = [vp]
    L.or_mesh_block_indices.argtypes = [vp, ip, C.c_int32]
    L.or_mesh_block_sizes.argtypes = [vp, ip, C.POINTER(C.c_int32 * 3)]
    L.or_mesh_get_block.argtypes = [vp, ip, fp, fp, ip, u8p]
    L.or_esdf_set_block.argtypes = [vp, ip, vp]
    L.or_freespace_set_block.argtypes = [vp, ip, vp]
    L.or_freespace_set_block.restype = None
    L.or_occupancy_set_block.argtypes = [vp, ip, vp]
    L.or_occupancy_set_block.restype = None
    L.or_esdf_integrate_slice_planar.argtypes = [vp, C.c_int32, C.c_int32, ip, C.c_int32, C.POINTER(EsdfParams), fp, C.c_float,
                                                 C.c_float, C.c_float]
    L.or_esdf_integrate_slice_planar.restype = None
    u8p = C.POINTER(C.c_uint8)
    L.or_mark_unobserved_free_inside_radius.argtypes = [vp, C.c_int32, fp, C.c_float, C.c_float, ip, C.c_int32]
    L.or_mark_unobserved_free_inside_radius.restype = C.c_int32
    L.or_view_projection_blocks.argtypes = [C.c_float, fp, C.POINTER(Camera), C.c_float, ip, C.c_int32]
    L.or_view_projection_blocks.restype = C.c_int32
    L.or_weighting.argtypes = [C.c_int32, C.c_float, C.c_float, C.c_float]
    L.or_weighting.restype = C.c_float
    L.or_round_through_half.argtypes = [C.c_float]
    L.or_round_through_half.restype = C.c_float
    L.or_sphere_trace_ray.argtypes = [vp, fp, fp, C.c_float, C.c_int32, C.c_float, C.c_float, fp]
    L.or_sphere_trace_ray.restype = C.c_int32
    L.or_sphere_trace_image.argtypes = [vp, fp, C.POINTER(Camera), C.c_float, C.c_int32, C.c_float, C.c_float, C.c_int32, fp]
    L.or_sphere_trace_image.restype = None
    L.or_color_integrate.argtypes = [vp, u8p, u8p, C.c_int32, C.c_int32, C.c_int32, fp, C.POINTER(Camera),
                                     C.POINTER(ColorParams), ip, C.c_int32]
    L.or_color_integrate.restype = C.c_int32
    L.or_color_num_blocks.argtypes = [vp]
    L.or_color_num_blocks.restype = C.c_int32
    L.or_color_block_indices.argtypes = [vp, ip, C.c_int32]
    L.or_color_block_indices.restype = C.c_int32
    L.or_color_get_block.argtypes = [vp, ip, vp]
    L.or_color_get_block.restype = C.c_int32
    L.or_planar_column_bounds.argtypes = [C.c_float, fp, C.c_float, C.c_float, C.c_int32, C.c_int32, C.c_int32, C.c_int32, ip]
    L.or_planar_column_bounds.restype = None
    L.or_planar_num_blocks_in_column.argtypes = [C.c_float, C.c_float]
    L.or_planar_num_blocks_in_column.restype = C.c_int32
    L.or_block_and_voxel_from_1d.argtypes = [C.c_float, C.c_float, ip]
    L.or_block_and_voxel_from_1d.restype = None
    L.or_esdf_slice_image.argtypes = [vp, C.c_float, C.c_float, fp, fp, C.POINTER(C.c_int8), C.c_int32, ip, ip]
    L.or_esdf_slice_image.restype = C.c_int32
    L.or_esdf_slice_aabb.argtypes = [vp, C.c_float, fp]
    L.or_esdf_slice_aabb.restype = C.c_int32
    L.or_esdf_slice_image_in_aabb.argtypes = [vp, C.c_float, C.c_float, fp, fp, C.POINTER(C.c_int8), C.c_int32, ip, ip]
    L.or_esdf_slice_image_in_aabb.restype = C.c_int32
    L.or_esdf_integrate_slice.argtypes = [vp, C.c_int32, C.c_int32, ip, C.c_int32, C.POINTER(EsdfParams), C.c_float, C.c_float,
                                          C.c_float]
    L.or_esdf_integrate_slice.restype = None
    L.or_default_freespace_params.argtypes = [C.POINTER(FreespaceParams)]
    L.or_default_freespace_params.restype = None
    L.or_freespace_update.argtypes = [vp, ip, C.c_int32, C.c_int64, C.POINTER(FreespaceParams), fp, C.c_int32, C.c_int32, fp,
                                      C.POINTER(Camera), C.c_float, C.c_float]
    L.or_freespace_update.restype = None
    L.or_freespace_num_blocks.argtypes = [vp]
    L.or_freespace_num_blocks.restype = C.c_int32
    L.or_freespace_block_indices.argtypes = [vp, ip, C.c_int32]
    L.or_freespace_block_indices.restype = C.c_int32
    L.or_freespace_get_block.argtypes = [vp, ip, vp]
    L.or_freespace_get_block.restype = C.c_int32
    L.or_esdf_integrate_with_freespace.argtypes = [vp, ip, C.c_int32, C.POINTER(EsdfParams)]
    L.or_esdf_integrate_with_freespace.restype = None
    L.or_default_tsdf_decay_params.argtypes = [C.POINTER(TsdfDecayParams)]
    L.or_default_tsdf_decay_params.restype = None
    L.or_default_occupancy_decay_params.argtypes = [C.POINTER(OccupancyDecayParams)]
    L.or_default_occupancy_decay_params.restype = None
    for name, ptype in (("or_tsdf_decay", TsdfDecayParams), ("or_occupancy_decay", OccupancyDecayParams)):
        f = getattr(L, name)
        f.argtypes = [vp, C.POINTER(ptype), C.POINTER(DecayExclusion), fp, C.c_int32, C.c_int32, fp, C.POINTER(Camera),
                      C.c_float, C.c_float, C.c_int32, ip, C.c_int32]
        f.restype = C.c_int32
    L.or_camera_project.argtypes = [C.POINTER(Camera), fp, fp]
    L.or_camera_project.restype = C.c_int32
    L.or_camera_vector_from_image_plane.argtypes = [C.POINTER(Camera), C.c_float, C.c_float, fp]
    L.or_num_threads.restype = C.c_int32
    L.or_set_num_threads.argtypes = [C.c_int32]
    _lib = L
    return L


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def colmajor(T):
    """4x4 numpy transform -> 16 float32 in Eigen (column-major) order."""
    return np.ascontiguousarray(np.asarray(T, dtype=np.float32).T).reshape(16)


def dilate_invalid(depth, num_dilations, threshold=1e-2, value=0.0):
    """DepthPreprocessor::dilateInvalidRegionsAsync (src/sensors/depth_preprocessing.cpp:36-58)."""
    depth = np.ascontiguousarray(depth, dtype=np.float32)
    out = np.empty_like(depth)
    lib().or_depth_dilate_invalid(_fp(depth), depth.shape[0], depth.shape[1], int(num_dilations), float(threshold), float(value),
                                  _fp(out))
    return out


def combined_slice_image(map_1, map_2, slice_height_1, slice_height_2, unobserved_value=1000.0):
    """EsdfSlicer::sliceLayersToCombinedDistanceImage (src/integrators/esdf_slicer.cu:201-240): both layers sliced on the merged
    AABB of their slices (getCombinedAabbOfLayersAtHeight, :149-157), element-wise minimum. -> (aabb or None, image or None)."""
    boxes = [b for b in (map_1.esdf_slice_aabb(slice_height_1), map_2.esdf_slice_aabb(slice_height_2)) if b is not None]
    if not boxes:
        return None, None  # empty AABB: the reference returns without touching the output
    aabb = np.concatenate([np.min([b[:3] for b in boxes], axis=0), np.max([b[3:] for b in boxes], axis=0)]).astype(np.float32)
    i1 = map_1.esdf_slice_image_in_aabb(slice_height_1, aabb, unobserved_value)
    i2 = map_2.esdf_slice_image_in_aabb(slice_height_2, aabb, unobserved_value)
    return aabb, np.minimum(i1, i2)


def default_tsdf_params(**kw):
    p = TsdfParams()
    lib().or_default_tsdf_params(C.byref(p))
    for k, v in kw.items():
        if k in ("workspace_min", "workspace_max"):
            setattr(p, k, (C.c_float * 3)(*v))
        else:
            setattr(p, k, v)
    return p


def default_esdf_params(**kw):
    p = EsdfParams()
    lib().or_default_esdf_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def default_freespace_params(**kw):
    p = FreespaceParams()
    lib().or_default_freespace_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def view_projection_blocks(T_L_C, cam, block_size, max_distance, cap=1 << 20):
    """ViewCalculator::getBlocksInImageViewProjection(T_L_C, camera, block_size, max_distance) -> (n, 3) block indices."""
    T = colmajor(T_L_C)
    out = np.zeros((cap, 3), dtype=np.int32)
    n = lib().or_view_projection_blocks(float(block_size), _fp(T), C.byref(cam), float(max_distance), _ip(out), cap)
    assert n <= cap
    return out[:n].copy()


def default_color_params(**kw):
    p = ColorParams()
    lib().or_default_color_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def weighting(wtype, measured, voxel_depth, truncation_distance):
    """WeightingFunction(type)(measured depth, voxel depth, truncation distance)."""
    return float(lib().or_weighting(int(wtype), float(measured), float(voxel_depth), float(truncation_distance)))


def round_through_half(f):
    return float(lib().or_round_through_half(float(f)))


def default_tsdf_decay_params(**kw):
    p = TsdfDecayParams()
    lib().or_default_tsdf_decay_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def default_occupancy_decay_params(**kw):
    p = OccupancyDecayParams()
    lib().or_default_occupancy_decay_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def default_occupancy_params(**kw):
    p = OccupancyParams()
    lib().or_default_occupancy_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def raycast_cells(origin, dest, scale=1.0, cap=1 << 16):
    o = np.asarray(origin, dtype=np.float32)
    d = np.asarray(dest, dtype=np.float32)
    out = np.zeros((cap, 3), dtype=np.int32)
    n = lib().or_raycast_cells(_fp(o), _fp(d), scale, _ip(out), cap)
    return out[:n].copy()


def view_raycast(depth, T_L_C, cam, block_size, truncation_distance_m, params=None, cap=1 << 20):
    depth = np.ascontiguousarray(depth, dtype=np.float32)
    params = params or default_tsdf_params()
    T = colmajor(T_L_C)
    out = np.zeros((cap, 3), dtype=np.int32)
    n = lib().or_view_raycast(_fp(depth), depth.shape[0], depth.shape[1], _fp(T), C.byref(cam),
                              block_size, truncation_distance_m, C.byref(params), _ip(out), cap)
    assert n <= cap
    return out[:n].copy()


class OracleMap:
    """One TSDF layer + one ESDF layer + the EsdfIntegrator's carried-over state."""

    def __init__(self, voxel_size):
        self.voxel_size = float(voxel_size)
        self._h = lib().or_map_create(voxel_size)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().or_map_destroy(self._h)
            self._h = None

    def clear(self):
        lib().or_map_clear(self._h)

    def depth_preprocessing(self, enable, num_dilations=4):
        """Mapper::do_depth_preprocessing / depth_preprocessing_num_dilations (mapper_params.h:33-42; defaults off / 4):
        integrate_depth / integrate_occupancy then see the dilated image, like Mapper::integrateDepth (mapper_impl.h:38-42)."""
        self._pre = (bool(enable), int(num_dilations))

    def _preprocessed(self, depth):
        depth = np.ascontiguousarray(depth, dtype=np.float32)
        pre = getattr(self, "_pre", (False, 4))
        return dilate_invalid(depth, pre[1]) if pre[0] else depth

    def integrate_depth(self, depth, T_L_C, cam, params=None, mask=None, mask_mode=0, cap=1 << 20):
        depth = self._preprocessed(depth)
        params = params or default_tsdf_params()
        T = colmajor(T_L_C)
        out = np.zeros((cap, 3), dtype=np.int32)
        mp = None
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.uint8)
            mp = mask.ctypes.data_as(C.POINTER(C.c_uint8))
        n = lib().or_tsdf_integrate(self._h, _fp(depth), mp, mask_mode, depth.shape[0],
                                    depth.shape[1], _fp(T), C.byref(cam), C.byref(params),
                                    _ip(out), cap)
        assert n <= cap
        return out[:n].copy()

    def integrate_depth_blocks(self, depth, T_L_C, cam, blocks, params=None, mask=None, mask_mode=0):
        depth = np.ascontiguousarray(depth, dtype=np.float32)
        params = params or default_tsdf_params()
        T = colmajor(T_L_C)
        blocks = np.ascontiguousarray(blocks, dtype=np.int32).reshape(-1, 3)
        mp = None
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.uint8)
            mp = mask.ctypes.data_as(C.POINTER(C.c_uint8))
        lib().or_tsdf_integrate_blocks(self._h, _fp(depth), mp, mask_mode, depth.shape[0],
                                       depth.shape[1], _fp(T), C.byref(cam), C.byref(params),
                                       _ip(blocks), blocks.shape[0])

    def integrate_occupancy(self, depth, T_L_C, cam, params=None, occ_params=None, mask=None, mask_mode=0, cap=1 << 20):
        """ProjectiveOccupancyIntegrator::integrateFrame; `params` (TsdfParams) is updated in place when the
        truncation distance is raised to the occupied half width."""
        depth = self._preprocessed(depth)
        params = params or default_tsdf_params()
        occ_params = occ_params or default_occupancy_params()
        T = colmajor(T_L_C)
        out = np.zeros((cap, 3), dtype=np.int32)
        mp = None
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.uint8)
            mp = mask.ctypes.data_as(C.POINTER(C.c_uint8))
        n = lib().or_occupancy_integrate(self._h, _fp(depth), mp, mask_mode, depth.shape[0], depth.shape[1], _fp(T),
                                         C.byref(cam), C.byref(params), C.byref(occ_params), _ip(out), cap)
        assert n <= cap
        return out[:n].copy()

    def integrate_esdf_occupancy(self, blocks, params=None):
        params = params or default_esdf_params()
        blocks = np.ascontiguousarray(blocks, dtype=np.int32).reshape(-1, 3)
        lib().or_esdf_integrate_occupancy(self._h, _ip(blocks), blocks.shape[0], C.byref(params))

    def occupancy_block_indices(self):
        n = lib().or_occupancy_num_blocks(self._h)
        out = np.zeros((max(n, 1), 3), dtype=np.int32)
        lib().or_occupancy_block_indices(self._h, _ip(out), n)
        return out[:n].copy()

    def occupancy_layer(self):
        """{(x,y,z): (8,8,8) float32 log-odds} for every allocated occupancy block."""
        out = {}
        for k in self.occupancy_block_indices():
            blk = np.zeros((8, 8, 8), dtype=np.float32)
            lib().or_occupancy_get_block(self._h, _ip(np.ascontiguousarray(k, dtype=np.int32)), blk.ctypes.data)
            out[tuple(int(c) for c in k)] = blk
        return out

    def _decay(self, fn, params, depth, T_L_C, cam, max_view_distance_m, truncation_distance_m, excluded_blocks,
               exclusion_center, exclusion_radius_m, clear_esdf, cap):
        x = DecayExclusion()
        keep = None
        if excluded_blocks is not None and len(excluded_blocks):
            keep = np.ascontiguousarray(excluded_blocks, dtype=np.int32).reshape(-1, 3)
            x.excluded_blocks_xyz, x.num_excluded_blocks = _ip(keep), keep.shape[0]
        if exclusion_center is not None and exclusion_radius_m is not None:
            x.has_exclusion_sphere = 1
            x.exclusion_center = (C.c_float * 3)(*[float(v) for v in exclusion_center])
            x.exclusion_radius_m = float(exclusion_radius_m)
        out = np.zeros((cap, 3), dtype=np.int32)
        if depth is not None:
            depth = np.ascontiguousarray(depth, dtype=np.float32)
            T = colmajor(T_L_C)
            n = fn(self._h, C.byref(params), C.byref(x), _fp(depth), depth.shape[0], depth.shape[1], _fp(T), C.byref(cam),
                   float(max_view_distance_m), float(truncation_distance_m), int(clear_esdf), _ip(out), cap)
        else:
            n = fn(self._h, C.byref(params), C.byref(x), None, 0, 0, None, None, 0.0, 0.0, int(clear_esdf), _ip(out), cap)
        assert n <= cap
        return out[:n].copy()

    def decay_tsdf(self, params=None, depth=None, T_L_C=None, cam=None, max_view_distance_m=7.0,
                   truncation_distance_m=None, excluded_blocks=None, exclusion_center=None, exclusion_radius_m=None,
                   clear_esdf=True, cap=1 << 20):
        """TsdfDecayIntegrator::decay (+ Mapper::clearBlocksInLayers for the ESDF layer). depth=None decays every voxel;
        with a view (depth, T_L_C, cam) voxels that have a depth measurement are spared. Returns deallocated indices."""
        if truncation_distance_m is None:
            truncation_distance_m = 4.0 * self.voxel_size
        return self._decay(lib().or_tsdf_decay, params or default_tsdf_decay_params(), depth, T_L_C, cam,
                           max_view_distance_m, truncation_distance_m, excluded_blocks, exclusion_center,
                           exclusion_radius_m, clear_esdf, cap)

    def decay_occupancy(self, params=None, depth=None, T_L_C=None, cam=None, max_view_distance_m=7.0,
                        truncation_distance_m=None, excluded_blocks=None, exclusion_center=None, exclusion_radius_m=None,
                        clear_esdf=True, cap=1 << 20):
        if truncation_distance_m is None:
            truncation_distance_m = 4.0 * self.voxel_size
        return self._decay(lib().or_occupancy_decay, params or default_occupancy_decay_params(), depth, T_L_C, cam,
                           max_view_distance_m, truncation_distance_m, excluded_blocks, exclusion_center,
                           exclusion_radius_m, clear_esdf, cap)

    def update_freespace(self, blocks, update_time_ms, params=None, depth=None, T_L_C=None, cam=None,
                         max_view_distance_m=3.4028234663852886e38, truncation_distance_m=3.4028234663852886e38):
        """FreespaceIntegrator::updateFreespaceLayer; with a view only voxels that have a depth measurement are updated."""
        params = params or default_freespace_params()
        blocks = np.ascontiguousarray(blocks, dtype=np.int32).reshape(-1, 3)
        if depth is not None:
            depth = np.ascontiguousarray(depth, dtype=np.float32)
            T = colmajor(T_L_C)
            lib().or_freespace_update(self._h, _ip(blocks), blocks.shape[0], int(update_time_ms), C.byref(params), _fp(depth),
                                      depth.shape[0], depth.shape[1], _fp(T), C.byref(cam), float(max_view_distance_m),
                                      float(truncation_distance_m))
        else:
            lib().or_freespace_update(self._h, _ip(blocks), blocks.shape[0], int(update_time_ms), C.byref(params), None, 0, 0,
                                      None, None, 0.0, 0.0)

    def freespace_block_indices(self):
        n = lib().or_freespace_num_blocks(self._h)
        out = np.zeros((max(n, 1), 3), dtype=np.int32)
        lib().or_freespace_block_indices(self._h, _ip(out), n)
        return out[:n].copy()

    def freespace_layer(self):
        out = {}
        for k in self.freespace_block_indices():
            blk = np.zeros((8, 8, 8), dtype=FREESPACE_VOXEL_DTYPE)
            lib().or_freespace_get_block(self._h, _ip(np.ascontiguousarray(k, dtype=np.int32)), blk.ctypes.data)
            out[tuple(int(c) for c in k)] = blk
        return out

    def mark_unobserved_free_inside_radius(self, center, radius, occupancy=False, truncation_distance_m=None, cap=1 << 20):
        """Mapper::markUnobservedTsdfFreeInsideRadius(center, radius) -> the blocks inside the radius (all allocated now)."""
        if truncation_distance_m is None:
            truncation_distance_m = np.float32(4.0) * np.float32(self.voxel_size)
        c = np.ascontiguousarray(center, dtype=np.float32).reshape(3)
        out = np.zeros((cap, 3), dtype=np.int32)
        n = lib().or_mark_unobserved_free_inside_radius(self._h, 1 if occupancy else 0, _fp(c), float(radius),
                                                        float(truncation_distance_m), _ip(out), cap)
        assert n <= cap
        return out[:n].copy()

    def integrate_color(self, color, T_L_C, cam, params=None, mask=None, mask_mode=0, cap=1 << 20):
        """ProjectiveColorIntegrator::integrateFrame(color image (rows, cols, 3) uint8 RGB, T_L_C, camera, tsdf_layer,
        color_layer, &updated_blocks) -> updated blocks."""
        color = np.ascontiguousarray(color, dtype=np.uint8)
        assert color.ndim == 3 and color.shape[2] == 3
        params = params or default_color_params()
        T = colmajor(T_L_C)
        out = np.zeros((cap, 3), dtype=np.int32)
        u8p = C.POINTER(C.c_uint8)
        mp = None
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.uint8)
            mp = mask.ctypes.data_as(u8p)
        n = lib().or_color_integrate(self._h, color.ctypes.data_as(u8p), mp, mask_mode, color.shape[0], color.shape[1], _fp(T),
                                     C.byref(cam), C.byref(params), _ip(out), cap)
        assert n <= cap
        return out[:n].copy()

    def sphere_trace_image(self, T_L_C, cam, truncation_distance_m, maximum_steps=100, maximum_ray_length_m=15.0,
                           surface_distance_epsilon_m=None, ray_subsampling_factor=1):
        """SphereTracer::renderImageOnGPU -> (rows / f, cols / f) depth image, -1 where no surface was found."""
        f = int(ray_subsampling_factor)
        out = np.zeros((cam.height // f, cam.width // f), np.float32)
        if surface_distance_epsilon_m is None:
            surface_distance_epsilon_m = np.float32(0.1) * np.float32(self.voxel_size)  # sphere_tracer.h:218
        T = colmajor(T_L_C)
        lib().or_sphere_trace_image(self._h, _fp(T), C.byref(cam), float(truncation_distance_m), int(maximum_steps),
                                    float(maximum_ray_length_m), float(surface_distance_epsilon_m), f, _fp(out))
        return out

    def sphere_trace_ray(self, origin, direction, truncation_distance_m, maximum_steps=100, maximum_ray_length_m=15.0,
                         surface_distance_epsilon_m=None):
        """SphereTracer::castOnGPU(ray, layer, truncation_distance_m, &t) -> (converged, t)."""
        if surface_distance_epsilon_m is None:
            surface_distance_epsilon_m = np.float32(0.1) * np.float32(self.voxel_size)
        o = np.ascontiguousarray(origin, dtype=np.float32)
        d = np.ascontiguousarray(direction, dtype=np.float32)
        t = np.zeros(1, np.float32)
        ok = lib().or_sphere_trace_ray(self._h, _fp(o), _fp(d), float(truncation_distance_m), int(maximum_steps),
                                       float(maximum_ray_length_m), float(surface_distance_epsilon_m), _fp(t))
        return bool(ok), float(t[0])

    def color_block_indices(self):
        n = lib().or_color_num_blocks(self._h)
        out = np.zeros((max(n, 1), 3), dtype=np.int32)
        lib().or_color_block_indices(self._h, _ip(out), n)
        return out[:n].copy()

    def color_block(self, idx):
        k = np.asarray(idx, dtype=np.int32)
        out = np.zeros((8, 8, 8), dtype=COLOR_VOXEL_DTYPE)
        ok = lib().or_color_get_block(self._h, _ip(k), out.ctypes.data)
        return out if ok else None

    def color_layer(self):
        return {tuple(int(c) for c in k): self.color_block(k) for k in self.color_block_indices()}

    # ---- mesh (mesh/mesh_integrator.h)
    def integrate_mesh(self, blocks=None, min_weight=1e-4, weld_vertices=True, cutoff_distance_vox=5.0):
        """MeshIntegrator::integrateBlocksGPU (blocks) / integrateMeshFromDistanceField (blocks=None: every TSDF block)."""
        if blocks is None:
            blocks = self.tsdf_block_indices()
        blocks = np.ascontiguousarray(blocks, dtype=np.int32).reshape(-1, 3)
        p = MeshParams(float(min_weight), 1 if weld_vertices else 0, float(cutoff_distance_vox))
        lib().or_mesh_integrate_blocks(self._h, _ip(blocks), blocks.shape[0], C.byref(p))

    def update_mesh_color(self, blocks=None):
        """MeshIntegrator::updateAppearance for the colour layer (blocks=None: every mesh block)."""
        if blocks is None:
            blocks = self.mesh_block_indices()
        blocks = np.ascontiguousarray(blocks, dtype=np.int32).reshape(-1, 3)
        lib().or_mesh_update_color(self._h, _ip(blocks), blocks.shape[0])

    def mesh_block_indices(self):
        n = lib().or_mesh_num_blocks(self._h)
        out = np.zeros((max(n, 1), 3), dtype=np.int32)
        lib().or_mesh_block_indices(self._h, _ip(out), n)
        return out[:n]

    def mesh_block(self, idx):
        k = np.ascontiguousarray(idx, dtype=np.int32).reshape(3)
        sz = (C.c_int32 * 3)()
        if not lib().or_mesh_block_sizes(self._h, _ip(k), sz):
            return None
        v = np.zeros((sz[0], 3), np.float32)
        nrm = np.zeros((sz[0], 3), np.float32)
        t = np.zeros(sz[1], np.int32)
        c = np.zeros((sz[2], 4), np.uint8)
        lib().or_mesh_get_block(self._h, _ip(k), _fp(v) if sz[0] else None, _fp(nrm) if sz[0] else None, _ip(t) if sz[1] else None,
                                c.ctypes.data_as(C.POINTER(C.c_uint8)) if sz[2] else None)
        return {"vertices": v, "normals": nrm, "triangles": t, "colors": c}

    def mesh_layer(self):
        return {tuple(int(c) for c in k): self.mesh_block(k) for k in self.mesh_block_indices()}

    def integrate_esdf_slice(self, blocks, params=None, z_min_m=0.0, z_max_m=1.0, z_output_m=1.0, from_occupancy=False,
                             use_freespace=False):
        """EsdfIntegrator::integrateSlice with a constant-z slice (defaults = esdf_integrator_params.h:33-43)."""
        params = params or default_esdf_params()
        blocks = np.ascontiguousarray(blocks, dtype=np.int32).reshape(-1, 3)
        lib().or_esdf_integrate_slice(self._h, 1 if from_occupancy else 0, 1 if use_freespace else 0, _ip(blocks),
                                      blocks.shape[0], C.byref(params), float(z_min_m), float(z_max_m), float(z_output_m))

    def integrate_esdf_slice_planar(self, blocks, plane, params=None, above_plane_m=0.0, thickness_m=0.1, z_output_m=1.0,
                                    from_occupancy=False, use_freespace=False):
        """EsdfIntegrator::integrateSlice(layer, blocks, ground_plane, esdf); plane = (nx, ny, nz, d), unit normal."""
        params = params or default_esdf_params()
        blocks = np.ascontiguousarray(blocks, dtype=np.int32).reshape(-1, 3)
        pl = np.ascontiguousarray(plane, dtype=np.float32).reshape(4)
        lib().or_esdf_integrate_slice_planar(self._h, 1 if from_occupancy else 0, 1 if use_freespace else 0, _ip(blocks),
                                             blocks.shape[0], C.byref(params), _fp(pl), float(above_plane_m), float(thickness_m),
                                             float(z_output_m))

    def esdf_slice_image(self, slice_height, unobserved_value=1000.0):
        """EsdfSlicer::sliceLayerToDistanceImage + occupancyGridFromSliceImage -> (aabb(6), image (rows, cols), grid int8)."""
        aabb = np.zeros(6, np.float32)
        r, c = C.c_int32(0), C.c_int32(0)
        n = lib().or_esdf_slice_image(self._h, float(slice_height), float(unobserved_value), _fp(aabb), None, None, 0,
                                      C.byref(r), C.byref(c))
        img = np.zeros((max(r.value, 1), max(c.value, 1)), np.float32)
        grid = np.zeros((max(r.value, 1), max(c.value, 1)), np.int8)
        if n > 0:
            lib().or_esdf_slice_image(self._h, float(slice_height), float(unobserved_value), _fp(aabb), _fp(img),
                                      grid.ctypes.data_as(C.POINTER(C.c_int8)), n, C.byref(r), C.byref(c))
        return aabb, img[:r.value, :c.value], grid[:r.value, :c.value]

    def esdf_slice_aabb(self, slice_height):
        """EsdfSlicer::getAabbOfLayerAtHeight -> (6,) float32 or None if the layer has no block at that height."""
        aabb = np.zeros(6, np.float32)
        return aabb if lib().or_esdf_slice_aabb(self._h, float(slice_height), _fp(aabb)) else None

    def esdf_slice_image_in_aabb(self, slice_height, aabb, unobserved_value=1000.0):
        """EsdfSlicer::sliceLayerToDistanceImage on a given AABB -> image (rows, cols)."""
        box = np.ascontiguousarray(aabb, np.float32).reshape(6)
        r, c = C.c_int32(0), C.c_int32(0)
        n = lib().or_esdf_slice_image_in_aabb(self._h, float(slice_height), float(unobserved_value), _fp(box), None, None, 0,
                                              C.byref(r), C.byref(c))
        img = np.zeros((max(r.value, 1), max(c.value, 1)), np.float32)
        if n > 0:
            lib().or_esdf_slice_image_in_aabb(self._h, float(slice_height), float(unobserved_value), _fp(box), _fp(img), None, n,
                                              C.byref(r), C.byref(c))
        return img[:r.value, :c.value]

    def integrate_esdf_with_freespace(self, blocks, params=None):
        params = params or default_esdf_params()
        blocks = np.ascontiguousarray(blocks, dtype=np.int32).reshape(-1, 3)
        lib().or_esdf_integrate_with_freespace(self._h, _ip(blocks), blocks.shape[0], C.byref(params))

    def set_occupancy_block(self, idx, log_odds):
        k = np.asarray(idx, dtype=np.int32)
        v = np.ascontiguousarray(log_odds, dtype=np.float32).reshape(8, 8, 8)
        lib().or_occupancy_set_block(self._h, _ip(k), v.ctypes.data)

    def integrate_esdf(self, blocks, params=None):
        params = params or default_esdf_params()
        blocks = np.ascontiguousarray(blocks, dtype=np.int32).reshape(-1, 3)
        lib().or_esdf_integrate(self._h, _ip(blocks), blocks.shape[0], C.byref(params))

    def esdf_stats(self):
        out = (C.c_int64 * 8)()
        lib().or_esdf_last_stats(self._h, out)
        keys = ("marked", "with_sites", "to_clear", "clear_candidates", "cleared", "swept",
                "face_passes", "rings")
        return dict(zip(keys, list(out)))

    def tsdf_block_indices(self):
        n = lib().or_tsdf_num_blocks(self._h)
        out = np.zeros((max(n, 1), 3), dtype=np.int32)
        lib().or_tsdf_block_indices(self._h, _ip(out), n)
        return out[:n].copy()

    def esdf_block_indices(self):
        n = lib().or_esdf_num_blocks(self._h)
        out = np.zeros((max(n, 1), 3), dtype=np.int32)
        lib().or_esdf_block_indices(self._h, _ip(out), n)
        return out[:n].copy()

    def tsdf_block(self, idx):
        k = np.asarray(idx, dtype=np.int32)
        out = np.zeros((8, 8, 8), dtype=TSDF_VOXEL_DTYPE)
        ok = lib().or_tsdf_get_block(self._h, _ip(k), out.ctypes.data)
        return out if ok else None

    def esdf_block(self, idx):
        k = np.asarray(idx, dtype=np.int32)
        out = np.zeros((8, 8, 8), dtype=ESDF_VOXEL_DTYPE)
        ok = lib().or_esdf_get_block(self._h, _ip(k), out.ctypes.data)
        return out if ok else None

    def set_tsdf_block(self, idx, voxels):
        k = np.asarray(idx, dtype=np.int32)
        v = np.ascontiguousarray(voxels, dtype=TSDF_VOXEL_DTYPE).reshape(8, 8, 8)
        lib().or_tsdf_set_block(self._h, _ip(k), v.ctypes.data)

    def cache_last_viewpoint(self, enable):
        """ViewCalculator::cache_last_viewpoint of the projective integrator (default on, like the reference)."""
        lib().or_map_cache_last_viewpoint(self._h, 1 if enable else 0)

    def set_esdf_block(self, idx, voxels):
        """Test hook: place an EsdfBlock with the given voxels (ESDF_VOXEL_DTYPE, (8, 8, 8))."""
        k = np.asarray(idx, dtype=np.int32)
        v = np.ascontiguousarray(voxels, dtype=ESDF_VOXEL_DTYPE).reshape(8, 8, 8)
        lib().or_esdf_set_block(self._h, _ip(k), v.ctypes.data)

    def set_freespace_block(self, idx, voxels):
        k = np.asarray(idx, dtype=np.int32)
        v = np.ascontiguousarray(voxels, dtype=FREESPACE_VOXEL_DTYPE).reshape(8, 8, 8)
        lib().or_freespace_set_block(self._h, _ip(k), v.ctypes.data)

    def tsdf_layer(self):
        """{(x,y,z): (8,8,8) structured array} for every allocated TSDF block."""
        return {tuple(int(c) for c in k): self.tsdf_block(k) for k in self.tsdf_block_indices()}

    def esdf_layer(self):
        return {tuple(int(c) for c in k): self.esdf_block(k) for k in self.esdf_block_indices()}


def planar_column_bounds(block_size, plane, above_plane_m, thickness_m, block_xy, voxel_xy):
    """PlanarSliceColumnBoundsGetter::getColumnBounds -> (min block z, min voxel z, max block z, max voxel z)."""
    out = np.zeros(4, np.int32)
    pl = np.ascontiguousarray(plane, dtype=np.float32).reshape(4)
    lib().or_planar_column_bounds(float(block_size), _fp(pl), float(above_plane_m), float(thickness_m), int(block_xy[0]),
                                  int(block_xy[1]), int(voxel_xy[0]), int(voxel_xy[1]), _ip(out))
    return tuple(int(v) for v in out)


def planar_num_blocks_in_column(block_size, thickness_m):
    return int(lib().or_planar_num_blocks_in_column(float(block_size), float(thickness_m)))


def block_and_voxel_from_1d(block_size, p):
    out = np.zeros(2, np.int32)
    lib().or_block_and_voxel_from_1d(float(block_size), float(p), _ip(out))
    return int(out[0]), int(out[1])


def camera_project(cam, p_C):
    p = np.asarray(p_C, dtype=np.float32)
    uv = np.zeros(2, np.float32)
    ok = lib().or_camera_project(C.byref(cam), _fp(p), _fp(uv))
    return (uv if ok else None)


def camera_vector_from_image_plane(cam, u, v):
    out = np.zeros(3, np.float32)
    lib().or_camera_vector_from_image_plane(C.byref(cam), float(u), float(v), _fp(out))
    return out


def num_threads():
    return lib().or_num_threads()


def set_num_threads(n):
    lib().or_set_num_threads(int(n))
