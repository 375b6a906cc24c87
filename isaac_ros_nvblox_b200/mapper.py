"""Host-side mirror of nvblox::Mapper for the depth-integration path.

Same names and argument meaning as the reference's C++ classes, restricted to this
path (nvblox/include/nvblox/mapper/mapper.h:107-836):
    Mapper(voxel_size_m)                       mapper.h:119-124
    Mapper.integrate_depth(depth, T_L_C, cam)  mapper.h:167-172  (integrateDepth)
    Mapper.update_esdf()                       mapper.h:326      (updateEsdf)
    Mapper.tsdf_layer() / esdf_layer()         mapper.h:372,393
    Mapper.tsdf_integrator() / esdf_integrator()  parameter setters
    Mapper(voxel_size_m, projective_layer_type=ProjectiveLayerType.kOccupancy)  mapper.h:52-53,119-124
    Mapper.occupancy_layer() / occupancy_integrator()  mapper.h:374,456
    ViewCalculator.get_blocks_in_image_view_raycast  view_calculator.h:75-80
Everything forwards to the C-ABI in libnvblox_b200.so through ctypes; numpy arrays
are host buffers, integers are raw device pointers.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import (NvbCamera, NvbDecayExclusion, NvbEsdfParams, NvbEsdfSliceParams, NvbFreespaceParams, NvbMapperOptions, NvbOccupancyDecayParams,
                   NvbOccupancyParams, NvbTsdfDecayParams, NvbTsdfParams, check)

TSDF_VOXEL_DTYPE = np.dtype([("distance", "<f4"), ("weight", "<f4")])
ESDF_VOXEL_DTYPE = np.dtype(
    [("squared_distance_vox", "<f4"), ("parent_direction", "<i4", (3,)),
     ("is_inside", "u1"), ("observed", "u1"), ("is_site", "u1"), ("pad", "u1")])

OCCUPANCY_VOXEL_DTYPE = np.dtype([("log_odds", "<f4")])  # map/voxels.h:51-53
COLOR_VOXEL_DTYPE = np.dtype([("color", "u1", (3,)), ("pad", "u1"), ("weight", "<f4")])  # ColorVoxel (map/voxels.h:77-83)
FREESPACE_VOXEL_DTYPE = np.dtype([("last_occupied_timestamp_ms", "<i8"), ("consecutive_occupancy_duration_ms", "<i8"),
                                  ("is_high_confidence_freespace", "u1"), ("pad", "u1", (7,))])  # map/voxels.h:38-52


class ProjectiveLayerType:
    """mapper/mapper.h:52-53."""
    kTsdf = _lib.NVB_PROJECTIVE_TSDF
    kOccupancy = _lib.NVB_PROJECTIVE_OCCUPANCY
    kTsdfWithFreespace = _lib.NVB_PROJECTIVE_TSDF_WITH_FREESPACE


STAGE_NAMES = ("view_calculator/raycast", "tsdf/integrate/allocate_blocks", "tsdf/integrate/update_blocks",
               "esdf/integrate/mark_sites", "esdf/integrate/clear", "esdf/integrate/compute")


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def colmajor(T):
    """4x4 transform -> 16 float32 in Eigen::Isometry3f::data() (column-major) order."""
    return np.ascontiguousarray(np.asarray(T, dtype=np.float32).T).reshape(16)


class Camera:
    """nvblox::Camera(fu, fv, cu, cv, width, height, distortion_params) (sensors/camera.h:33-203).
    radial = (k1..k6), tangential = (p1, p2) = RadialTangentialDistortionParams; None = std::nullopt."""

    def __init__(self, fu, fv, cu, cv, width, height, radial=None, tangential=None):
        has = radial is not None or tangential is not None
        k = [float(v) for v in (radial or (0, 0, 0, 0, 0, 0))]
        p = [float(v) for v in (tangential or (0, 0))]
        self.c = NvbCamera(float(fu), float(fv), float(cu), float(cv), int(width), int(height), 1 if has else 0,
                           *k, *p)

    fu = property(lambda s: s.c.fu)
    fv = property(lambda s: s.c.fv)
    cu = property(lambda s: s.c.cu)
    cv = property(lambda s: s.c.cv)
    width = property(lambda s: s.c.width)
    height = property(lambda s: s.c.height)


class _Layer:
    """BlockLayer queries (map/layer.h:76-311) answered from the device-resident map."""

    def __init__(self, mapper, layer_id, dtype):
        self._m, self._id, self._dtype = mapper, layer_id, dtype

    def num_blocks(self):
        n = C.c_int32(0)
        check(self._m._L.nvb_layer_num_blocks(self._m._h, self._id, C.byref(n)))
        return n.value

    def get_all_block_indices(self):
        n = self.num_blocks()
        out = np.zeros((max(n, 1), 3), dtype=np.int32)
        cnt = C.c_int32(0)
        check(self._m._L.nvb_layer_block_indices(self._m._h, self._id, _ip(out), n, C.byref(cnt)))
        return out[:n].copy()

    def get_blocks(self, indices):
        """(n,3) int32 -> ((n,8,8,8) voxel array, (n,) found mask)."""
        idx = np.ascontiguousarray(indices, dtype=np.int32).reshape(-1, 3)
        n = idx.shape[0]
        out = np.zeros((max(n, 1), 8, 8, 8), dtype=self._dtype)
        found = np.zeros(max(n, 1), dtype=np.uint8)
        check(self._m._L.nvb_layer_get_blocks(self._m._h, self._id, _ip(idx), n, out.ctypes.data,
                                              found.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out[:n], found[:n].astype(bool)

    def get_block_at_index(self, index):
        v, f = self.get_blocks(np.asarray(index, dtype=np.int32).reshape(1, 3))
        return v[0] if f[0] else None

    def is_block_allocated(self, index):
        return self.get_block_at_index(index) is not None

    def set_blocks(self, indices, voxels):
        idx = np.ascontiguousarray(indices, dtype=np.int32).reshape(-1, 3)
        v = np.ascontiguousarray(voxels, dtype=self._dtype).reshape(idx.shape[0], 8, 8, 8)
        check(self._m._L.nvb_layer_set_blocks(self._m._h, self._id, _ip(idx), idx.shape[0], v.ctypes.data))

    def block_device_ptr(self, index):
        k = np.asarray(index, dtype=np.int32)
        p = C.c_void_p(0)
        check(self._m._L.nvb_layer_block_device_ptr(self._m._h, self._id, _ip(k), C.byref(p)))
        return p.value or 0

    def as_dict(self):
        idx = self.get_all_block_indices()
        if len(idx) == 0:
            return {}
        v, _ = self.get_blocks(idx)
        return {tuple(int(c) for c in k): v[i] for i, k in enumerate(idx)}


class _TsdfIntegrator:
    """ProjectiveTsdfIntegrator parameter surface (projective_tsdf_integrator.h:59-121,
    projective_integrator.h:56-85, view_calculator.h:88-145)."""

    def __init__(self, mapper):
        self._m = mapper

    def _get(self):
        p = NvbTsdfParams()
        check(self._m._L.nvb_mapper_get_tsdf_params(self._m._h, C.byref(p)))
        return p

    def _set(self, **kw):
        p = self._get()
        for k, v in kw.items():
            if k in ("workspace_min", "workspace_max"):
                setattr(p, k, (C.c_float * 3)(*v))
            else:
                setattr(p, k, v)
        check(self._m._L.nvb_mapper_set_tsdf_params(self._m._h, C.byref(p)))

    def params(self, **kw):
        if kw:
            self._set(**kw)
        return self._get()

    def truncation_distance_vox(self, v=None):
        if v is not None:
            self._set(truncation_distance_vox=float(v))
        return self._get().truncation_distance_vox

    def max_integration_distance_m(self, v=None):
        if v is not None:
            self._set(max_integration_distance_m=float(v))
        return self._get().max_integration_distance_m

    def max_weight(self, v=None):
        if v is not None:
            self._set(max_weight=float(v))
        return self._get().max_weight

    def invalid_depth_decay_factor(self, v=None):
        if v is not None:
            self._set(invalid_depth_decay_factor=float(v))
        return self._get().invalid_depth_decay_factor

    def weighting_function_type(self, v=None):
        if v is not None:
            self._set(weighting_type=int(v))
        return self._get().weighting_type

    def cache_last_viewpoint(self, v=None):
        """ViewCalculator::cache_last_viewpoint (view_calculator.h:196): on by default, like the reference."""
        if v is not None:
            check(self._m._L.nvb_mapper_set_cache_last_viewpoint(self._m._h, 1 if v else 0))
        return bool(self._m._L.nvb_mapper_get_cache_last_viewpoint(self._m._h))

    def raycast_subsampling_factor(self, v=None):
        if v is not None:
            self._set(raycast_subsampling=int(v))
        return self._get().raycast_subsampling


class _OccupancyIntegrator(_TsdfIntegrator):
    """ProjectiveOccupancyIntegrator parameter surface (projective_occupancy_integrator.h:57-109) on top of
    the shared ProjectiveIntegrator / ViewCalculator one."""

    def _get_occ(self):
        p = NvbOccupancyParams()
        check(self._m._L.nvb_mapper_get_occupancy_params(self._m._h, C.byref(p)))
        return p

    def occupancy_params(self, **kw):
        p = self._get_occ()
        for k, v in kw.items():
            setattr(p, k, float(v))
        if kw:
            check(self._m._L.nvb_mapper_set_occupancy_params(self._m._h, C.byref(p)))
        return p

    def free_region_occupancy_probability(self, v=None):
        return self.occupancy_params(**({} if v is None else {"free_region_occupancy_probability": v})) \
            .free_region_occupancy_probability

    def occupied_region_occupancy_probability(self, v=None):
        return self.occupancy_params(**({} if v is None else {"occupied_region_occupancy_probability": v})) \
            .occupied_region_occupancy_probability

    def unobserved_region_occupancy_probability(self, v=None):
        return self.occupancy_params(**({} if v is None else {"unobserved_region_occupancy_probability": v})) \
            .unobserved_region_occupancy_probability

    def occupied_region_half_width_m(self, v=None):
        return self.occupancy_params(**({} if v is None else {"occupied_region_half_width_m": v})) \
            .occupied_region_half_width_m


class _ColorIntegrator:
    """ProjectiveColorIntegrator parameter surface (projective_appearance_integrator.h:96-148) and its SphereTracer."""

    def __init__(self, mapper):
        self._m = mapper

    def params(self, **kw):
        p = _lib.NvbColorParams()
        check(self._m._L.nvb_mapper_get_color_params(self._m._h, C.byref(p)))
        for k, v in kw.items():
            setattr(p, k, v)
        if kw:
            check(self._m._L.nvb_mapper_set_color_params(self._m._h, C.byref(p)))
        return p

    def render_depth(self, T_L_C, camera, truncation_distance_m, ray_subsampling_factor=1):
        """SphereTracer::renderImageOnGPU(camera, T_L_C, tsdf_layer, truncation_distance_m, ..., ray_subsampling_factor)
        -> (height / f, width / f) float32, -1 where a ray found no surface."""
        f = int(ray_subsampling_factor)
        out = np.zeros((max(camera.c.height // max(f, 1), 1), max(camera.c.width // max(f, 1), 1)), np.float32)
        T = colmajor(T_L_C)
        check(self._m._L.nvb_sphere_tracer_render_depth(self._m._h, _fp(T), C.byref(camera.c), float(truncation_distance_m), f,
                                                        _fp(out)))
        return out


class _FreespaceIntegrator:
    """FreespaceIntegrator parameter surface (freespace_integrator.h:75-128) + updateFreespaceLayer on a block list."""

    def __init__(self, mapper):
        self._m = mapper

    def params(self, **kw):
        p = NvbFreespaceParams()
        check(self._m._L.nvb_mapper_get_freespace_params(self._m._h, C.byref(p)))
        for k, v in kw.items():
            setattr(p, k, v)
        if kw:
            check(self._m._L.nvb_mapper_set_freespace_params(self._m._h, C.byref(p)))
        return p

    def update_freespace_layer(self, block_indices, update_time_ms, depth=None, T_L_C=None, camera=None,
                               max_view_distance_m=0.0, truncation_distance_m=0.0):
        idx = np.ascontiguousarray(block_indices, dtype=np.int32).reshape(-1, 3)
        if depth is not None:
            depth = np.ascontiguousarray(depth, dtype=np.float32)
            T = colmajor(T_L_C)
            check(self._m._L.nvb_freespace_update_blocks(self._m._h, _ip(idx), idx.shape[0], int(update_time_ms),
                                                         depth.ctypes.data, _lib.NVB_MEM_HOST, depth.shape[0], depth.shape[1],
                                                         _fp(T), C.byref(camera.c), float(max_view_distance_m),
                                                         float(truncation_distance_m)))
        else:
            check(self._m._L.nvb_freespace_update_blocks(self._m._h, _ip(idx), idx.shape[0], int(update_time_ms), None, 0, 0, 0,
                                                         None, None, 0.0, 0.0))


class _DecayIntegrator:
    """TsdfDecayIntegrator / OccupancyDecayIntegrator parameter surface (tsdf_decay_integrator.h:73-101,
    occupancy_decay_integrator.h:72-101, internal/decay_integrator_base.h:50-58)."""

    def __init__(self, mapper, occupancy):
        self._m, self._occ = mapper, occupancy

    def params(self, **kw):
        L, h = self._m._L, self._m._h
        p = NvbOccupancyDecayParams() if self._occ else NvbTsdfDecayParams()
        get = L.nvb_mapper_get_occupancy_decay_params if self._occ else L.nvb_mapper_get_tsdf_decay_params
        put = L.nvb_mapper_set_occupancy_decay_params if self._occ else L.nvb_mapper_set_tsdf_decay_params
        check(get(h, C.byref(p)))
        for k, v in kw.items():
            setattr(p, k, v)
        if kw:
            check(put(h, C.byref(p)))
        return p

    def deallocate_decayed_blocks(self, v=None):
        return bool(self.params(**({} if v is None else {"deallocate_decayed_blocks": 1 if v else 0})).deallocate_decayed_blocks)

    def decay_factor(self, v=None):
        return self.params(**({} if v is None else {"decay_factor": float(v)})).decay_factor

    def decay_to_free(self, v):
        """OccupancyDecayIntegrator::decay_to_free (src/integrators/occupancy_decay_integrator.cu:59-72)."""
        return self.params(decay_to_probability=0.49 if v else 0.5).decay_to_probability


class _EsdfIntegrator:
    """EsdfIntegrator parameter surface (esdf_integrator.h:178-283) + integrateBlocks."""

    def __init__(self, mapper):
        self._m = mapper

    def _get(self):
        p = NvbEsdfParams()
        check(self._m._L.nvb_mapper_get_esdf_params(self._m._h, C.byref(p)))
        return p

    def params(self, **kw):
        p = self._get()
        for k, v in kw.items():
            setattr(p, k, v)
        if kw:
            check(self._m._L.nvb_mapper_set_esdf_params(self._m._h, C.byref(p)))
        return p

    def max_esdf_distance_m(self, v=None):
        return self.params(**({} if v is None else {"max_esdf_distance_m": float(v)})).max_esdf_distance_m

    def max_site_distance_vox(self, v=None):
        return self.params(**({} if v is None else {"max_site_distance_vox": float(v)})).max_site_distance_vox

    def min_weight(self, v=None):
        return self.params(**({} if v is None else {"min_weight": float(v)})).min_weight

    def occupied_threshold(self, v=None):
        return self.params(**({} if v is None else {"occupied_threshold": float(v)})).occupied_threshold

    def slice_params(self, **kw):
        """esdf_slice_min_height / esdf_slice_max_height / esdf_slice_height (esdf_integrator.h:216-256)."""
        p = NvbEsdfSliceParams()
        check(self._m._L.nvb_mapper_get_esdf_slice_params(self._m._h, C.byref(p)))
        for k, v in kw.items():
            setattr(p, k, float(v))
        if kw:
            check(self._m._L.nvb_mapper_set_esdf_slice_params(self._m._h, C.byref(p)))
        return p

    def integrate_slice(self, block_indices, ground_plane=None):
        """EsdfIntegrator::integrateSlice(layer, block_indices[, ground_plane], esdf_layer): constant-z slice, or the planar
        one when ground_plane = (nx, ny, nz, d) (unit normal, n . p + d = 0) is given."""
        idx = np.ascontiguousarray(block_indices, dtype=np.int32).reshape(-1, 3)
        if ground_plane is None:
            check(self._m._L.nvb_esdf_integrate_slice_blocks(self._m._h, _ip(idx), idx.shape[0]))
        else:
            pl = np.ascontiguousarray(ground_plane, dtype=np.float32).reshape(4)
            check(self._m._L.nvb_esdf_integrate_slice_planar_blocks(self._m._h, _fp(pl), _ip(idx), idx.shape[0]))

    def integrate_blocks(self, block_indices):
        """EsdfIntegrator::integrateBlocks(tsdf_layer | occupancy_layer, block_indices, esdf_layer)."""
        idx = np.ascontiguousarray(block_indices, dtype=np.int32).reshape(-1, 3)
        check(self._m._L.nvb_esdf_integrate_blocks(self._m._h, _ip(idx), idx.shape[0]))

    def last_stats(self):
        out = (C.c_int64 * 8)()
        check(self._m._L.nvb_mapper_last_esdf_stats(self._m._h, out))
        keys = ("marked", "with_sites", "to_clear", "clear_candidates", "cleared", "swept", "face_passes", "rings")
        return dict(zip(keys, list(out)))

    def clear_blocks_read(self):
        """Blocks the last clear pass read (<= clear_candidates: candidates whose parents cannot lie in a to-clear block are skipped)."""
        out = C.c_int64(0)
        check(self._m._L.nvb_mapper_esdf_clear_blocks_read(self._m._h, C.byref(out)))
        return int(out.value)


class _MeshIntegrator:
    """MeshIntegrator parameters (mesh/mesh_integrator.h:83-87,126-133)."""

    def __init__(self, mapper):
        self._m = mapper

    def params(self, **kw):
        p = _lib.NvbMeshParams()
        check(self._m._L.nvb_mapper_get_mesh_params(self._m._h, C.byref(p)))
        if kw:
            for k, v in kw.items():
                setattr(p, k, v)
            check(self._m._L.nvb_mapper_set_mesh_params(self._m._h, C.byref(p)))
        return p

    def min_weight(self, v=None):
        return self.params(**({} if v is None else {"min_weight": float(v)})).min_weight

    def weld_vertices(self, v=None):
        return bool(self.params(**({} if v is None else {"weld_vertices": 1 if v else 0})).weld_vertices)

    def integrate_blocks(self, blocks, update_color=False):
        """MeshIntegrator::integrateBlocksGPU(tsdf_layer, block_indices, mesh_layer) [+ updateAppearance]."""
        b = np.ascontiguousarray(blocks, dtype=np.int32).reshape(-1, 3)
        check(self._m._L.nvb_mesh_integrate_blocks(self._m._h, _ip(b), b.shape[0], 1 if update_color else 0))

    def integrate_mesh_from_distance_field(self, update_color=False):
        """MeshIntegrator::integrateMeshFromDistanceField: every block of the TSDF layer."""
        self.integrate_blocks(self._m.tsdf_layer().get_all_block_indices(), update_color)

    def update_color(self, blocks=None):
        """MeshIntegrator::updateAppearance (blocks=None: every mesh block)."""
        if blocks is None:
            blocks = self._m.mesh_layer().get_all_block_indices()
        b = np.ascontiguousarray(blocks, dtype=np.int32).reshape(-1, 3)
        check(self._m._L.nvb_mesh_update_color(self._m._h, _ip(b), b.shape[0]))


class _MeshLayer:
    """MeshBlockLayer queries (mesh/mesh_block.h:32-83; map/layer.h) answered from the device arena."""

    def __init__(self, mapper):
        self._m = mapper

    def num_blocks(self):
        n = C.c_int32(0)
        rc = self._m._L.nvb_layer_num_blocks(self._m._h, _lib.NVB_LAYER_MESH, C.byref(n))
        return n.value if rc == 0 else 0  # no mesh update yet: an empty layer

    def get_all_block_indices(self):
        n = self.num_blocks()
        out = np.zeros((max(n, 1), 3), dtype=np.int32)
        if n:
            cnt = C.c_int32(0)
            check(self._m._L.nvb_layer_block_indices(self._m._h, _lib.NVB_LAYER_MESH, _ip(out), n, C.byref(cnt)))
        return out[:n].copy()

    def block_sizes(self, indices):
        """(n,3) -> (n,3) int32 {vertices, triangle indices, colours}; -1 where there is no mesh block."""
        idx = np.ascontiguousarray(indices, dtype=np.int32).reshape(-1, 3)
        out = np.full((max(idx.shape[0], 1), 3), -1, dtype=np.int32)
        check(self._m._L.nvb_mesh_block_sizes(self._m._h, _ip(idx), idx.shape[0], _ip(out)))
        return out[:idx.shape[0]]

    def get_blocks(self, indices):
        """(n,3) -> list of {"vertices" (v,3) f32, "normals" (v,3) f32, "triangles" (t,) i32, "colors" (c,4) u8} or None."""
        idx = np.ascontiguousarray(indices, dtype=np.int32).reshape(-1, 3)
        sz = self.block_sizes(idx)
        live = np.maximum(sz, 0).astype(np.int64)
        tv, tt, tc = (int(x) for x in live.sum(axis=0)) if idx.shape[0] else (0, 0, 0)
        V, N = np.zeros((max(tv, 1), 3), np.float32), np.zeros((max(tv, 1), 3), np.float32)
        T, Cc = np.zeros(max(tt, 1), np.int32), np.zeros((max(tc, 1), 4), np.uint8)
        caps = (C.c_int64 * 3)(tv, tt, tc)
        check(self._m._L.nvb_mesh_get_blocks(self._m._h, _ip(idx), idx.shape[0], V.ctypes.data, N.ctypes.data, T.ctypes.data,
                                             Cc.ctypes.data, caps))
        out, ov, ot, oc = [], 0, 0, 0
        for i in range(idx.shape[0]):
            if sz[i, 0] < 0:
                out.append(None)
                continue
            nv, nt, nc = (int(x) for x in sz[i])
            out.append({"vertices": V[ov:ov + nv].copy(), "normals": N[ov:ov + nv].copy(), "triangles": T[ot:ot + nt].copy(),
                        "colors": Cc[oc:oc + nc].copy()})
            ov, ot, oc = ov + nv, ot + nt, oc + nc
        return out

    def get_block_at_index(self, index):
        return self.get_blocks(np.asarray(index, dtype=np.int32).reshape(1, 3))[0]

    def is_block_allocated(self, index):
        return bool(self.block_sizes(np.asarray(index, dtype=np.int32).reshape(1, 3))[0, 0] >= 0)

    def as_dict(self):
        idx = self.get_all_block_indices()
        return {tuple(int(c) for c in k): b for k, b in zip(idx, self.get_blocks(idx))}

    def arena_stats(self):
        out = (C.c_int64 * 4)()
        check(self._m._L.nvb_mesh_arena_stats(self._m._h, out))
        return {"capacity": out[0], "used": out[1], "last_update_vertices": out[2]}


class EsdfSlicer:
    """EsdfSlicer (integrators/esdf_slicer.h:36-138): distance-map image and occupancy grid of an ESDF slice."""

    def __init__(self, mapper):
        self._m = mapper

    def slice_layer_to_distance_image(self, slice_height, unobserved_value=1000.0, with_occupancy_grid=False):
        """-> (aabb [min xyz, max xyz], (rows, cols) float32 image[, int8 grid]); rows follow y, columns x."""
        L, h = self._m._L, self._m._h
        aabb = np.zeros(6, np.float32)
        r, c = C.c_int32(0), C.c_int32(0)
        check(L.nvb_esdf_slice_distance_image(h, float(slice_height), float(unobserved_value), _fp(aabb), None, None, 0,
                                              C.byref(r), C.byref(c)))
        img = np.zeros((max(r.value, 1), max(c.value, 1)), np.float32)
        grid = np.zeros((max(r.value, 1), max(c.value, 1)), np.int8)
        if r.value * c.value > 0:
            check(L.nvb_esdf_slice_distance_image(h, float(slice_height), float(unobserved_value), _fp(aabb), _fp(img),
                                                  grid.ctypes.data_as(C.POINTER(C.c_int8)) if with_occupancy_grid else None,
                                                  r.value * c.value, C.byref(r), C.byref(c)))
        img = img[:r.value, :c.value]
        return (aabb, img, grid[:r.value, :c.value]) if with_occupancy_grid else (aabb, img)

    def get_aabb_of_layer_at_height(self, slice_height):
        """EsdfSlicer::getAabbOfLayerAtHeight (esdf_slicer.h:29-35) -> (6,) float32, or None for an empty box."""
        aabb = np.zeros(6, np.float32)
        empty = C.c_int32(1)
        check(self._m._L.nvb_esdf_slice_aabb(self._m._h, float(slice_height), _fp(aabb), C.byref(empty)))
        return None if empty.value else aabb

    def slice_layer_to_distance_image_in_aabb(self, slice_height, aabb, unobserved_value=1000.0):
        """EsdfSlicer::sliceLayerToDistanceImage(layer, slice_height, unobserved_value, aabb, image): a given box."""
        L, h = self._m._L, self._m._h
        box = np.ascontiguousarray(aabb, np.float32).reshape(6)
        r, c = C.c_int32(0), C.c_int32(0)
        check(L.nvb_esdf_slice_distance_image_in_aabb(h, float(slice_height), float(unobserved_value), _fp(box), None, None, 0,
                                                      C.byref(r), C.byref(c)))
        img = np.zeros((max(r.value, 1), max(c.value, 1)), np.float32)
        if r.value * c.value > 0:
            check(L.nvb_esdf_slice_distance_image_in_aabb(h, float(slice_height), float(unobserved_value), _fp(box), _fp(img), None,
                                                          r.value * c.value, C.byref(r), C.byref(c)))
        return img[:r.value, :c.value]

    def slice_layers_to_combined_distance_image(self, other_mapper, slice_height_1, slice_height_2, unobserved_value=1000.0,
                                                with_occupancy_grid=False):
        """EsdfSlicer::sliceLayersToCombinedDistanceImage (esdf_slicer.h:78-118, src/integrators/esdf_slicer.cu:201-240): this
        mapper's ESDF layer and another mapper's (e.g. MultiMapper's static and dynamic maps) sliced on the merged box of their
        slices, element-wise minimum. -> (aabb, image[, grid]); (None, None[, None]) if neither layer has a block there."""
        other = EsdfSlicer(other_mapper)
        boxes = [b for b in (self.get_aabb_of_layer_at_height(slice_height_1), other.get_aabb_of_layer_at_height(slice_height_2))
                 if b is not None]
        if not boxes:
            return (None, None, None) if with_occupancy_grid else (None, None)
        aabb = np.concatenate([np.min([b[:3] for b in boxes], axis=0), np.max([b[3:] for b in boxes], axis=0)]).astype(np.float32)
        img = np.minimum(self.slice_layer_to_distance_image_in_aabb(slice_height_1, aabb, unobserved_value),
                         other.slice_layer_to_distance_image_in_aabb(slice_height_2, aabb, unobserved_value))
        if not with_occupancy_grid:
            return aabb, img
        # occupancyGridFromSliceImageKernel (src/integrators/esdf_slicer.cu:78-110)
        grid = np.where(img < np.float32(1e-2), 100, 0).astype(np.int8)
        grid[np.abs(img - np.float32(unobserved_value)) < np.float32(1e-2)] = -1
        return aabb, img, grid


class Mapper:
    """nvblox::Mapper(voxel_size_m, projective_layer_type) with a projective (TSDF or occupancy) and an ESDF layer."""

    def esdf_time_split(self):
        out = (C.c_int64 * 4)()
        check(self._L.nvb_mapper_esdf_time_split(self._h, out))
        return {"barrier_wait_ns_cta0": out[0], "axis_ns_cta0": out[1], "slowest_cta_work_ns": out[2], "barriers": out[3]}

    def __init__(self, voxel_size_m, device=0, tsdf_capacity_blocks=0, esdf_capacity_blocks=0,
                 esdf_persistent=3, projective_layer_type=ProjectiveLayerType.kTsdf, keep_last_view=False):
        self._L = _lib.load()
        o = NvbMapperOptions()
        self._L.nvb_default_mapper_options(C.byref(o))
        o.voxel_size_m = float(voxel_size_m)
        o.device = int(device)
        if tsdf_capacity_blocks:
            o.tsdf_capacity_blocks = int(tsdf_capacity_blocks)
        if esdf_capacity_blocks:
            o.esdf_capacity_blocks = int(esdf_capacity_blocks)
        o.esdf_persistent = 1 if esdf_persistent is True else int(esdf_persistent)  # (True: the four-phase wavefront, as in round 1) 0 host loop, 1 four-phase wavefront, 2 gather-replay wavefront, 3 exchange-slab wavefront
        o.projective_layer_type = int(projective_layer_type)
        o.keep_last_view = 1 if keep_last_view else 0
        self._projective_layer_type = int(projective_layer_type)
        h = C.c_void_p(0)
        check(self._L.nvb_mapper_create(C.byref(o), C.byref(h)))
        self._h = h
        self._tsdf = _Layer(self, _lib.NVB_LAYER_TSDF, TSDF_VOXEL_DTYPE)
        self._occupancy = _Layer(self, _lib.NVB_LAYER_OCCUPANCY, OCCUPANCY_VOXEL_DTYPE)
        self._freespace = _Layer(self, _lib.NVB_LAYER_FREESPACE, FREESPACE_VOXEL_DTYPE)
        self._esdf = _Layer(self, _lib.NVB_LAYER_ESDF, ESDF_VOXEL_DTYPE)
        self._color = _Layer(self, _lib.NVB_LAYER_COLOR, COLOR_VOXEL_DTYPE)
        self._keep = []  # host buffers of in-flight async frames

    def close(self):
        if getattr(self, "_h", None):
            self._L.nvb_mapper_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- accessors -------------------------------------------------------
    def voxel_size(self):
        return self._L.nvb_mapper_voxel_size(self._h)

    def block_size(self):
        return self._L.nvb_mapper_block_size(self._h)

    def tsdf_layer(self):
        return self._tsdf

    def occupancy_layer(self):
        return self._occupancy

    def freespace_layer(self):
        return self._freespace

    def mark_unobserved_tsdf_free_inside_radius(self, center, radius):
        """Mapper::markUnobservedTsdfFreeInsideRadius(center, radius) (mapper.h:352-356) -> the blocks inside the radius."""
        c = np.ascontiguousarray(center, dtype=np.float32).reshape(3)
        n = C.c_int32(0)
        cap = 1 << 16
        out = np.empty((cap, 3), dtype=np.int32)
        check(self._L.nvb_mapper_mark_unobserved_free_inside_radius(self._h, _fp(c), float(radius), _ip(out), cap, C.byref(n)))
        if n.value > cap:
            raise RuntimeError("more than %d blocks inside the radius" % cap)
        return out[:n.value].copy()

    def color_layer(self):
        return self._color

    def mesh_layer(self):
        """Mapper::color_mesh_layer()."""
        return _MeshLayer(self)

    def mesh_integrator(self):
        """Mapper::color_mesh_integrator()."""
        return _MeshIntegrator(self)

    def save_color_mesh_as_ply(self, filename):
        """Mapper::saveColorMeshAsPly (src/mapper/mapper.cpp:694-696)."""
        from . import io as _io
        return _io.output_color_mesh_layer_to_ply(self.mesh_layer(), filename)

    def update_mesh(self, update_full_layer=False):
        """Mapper::updateColorMesh(UpdateFullLayer) (mapper.h; src/mapper/mapper.cpp:371-406)."""
        check(self._L.nvb_mapper_update_mesh(self._h, 1 if update_full_layer else 0))

    def color_integrator(self):
        return _ColorIntegrator(self)

    def integrate_color(self, color, T_L_C, camera, mask=None, mask_mode=0, return_blocks=True):
        """Mapper::integrateColor(color_frame[, mask], T_L_C, camera) (mapper.h:202-207). color: (rows, cols, 3) uint8 RGB.
        Returns updated_blocks (n, 3) (unordered)."""
        c = np.ascontiguousarray(color, dtype=np.uint8)
        if c.ndim != 3 or c.shape[2] != 3:
            raise ValueError("color must be (rows, cols, 3) uint8")
        mk = None
        if mask is not None:
            mk = np.ascontiguousarray(mask, dtype=np.uint8)
            if mk.shape != c.shape[:2]:
                raise ValueError("mask must have the colour image's size")
        T = colmajor(T_L_C)
        cap = 1 << 14 if return_blocks else 0
        n = C.c_int32(0)
        out = np.empty((max(cap, 1), 3), dtype=np.int32)
        check(self._L.nvb_mapper_integrate_color(self._h, c.ctypes.data, None if mk is None else mk.ctypes.data, mask_mode,
                                                 _lib.NVB_MEM_HOST, c.shape[0], c.shape[1], _fp(T), C.byref(camera.c),
                                                 _ip(out) if return_blocks else None, cap, C.byref(n)))
        if not return_blocks:
            return None
        if n.value > cap:
            raise RuntimeError("more than %d colour blocks in one frame" % cap)
        return out[:n.value].copy()

    def integrate_color_device(self, color_ptr, rows, cols, T_L_C, camera, mask_ptr=0, mask_mode=0):
        """Same, for an RGB frame already resident in HBM (raw device pointers); enqueued without synchronising."""
        T = colmajor(T_L_C)
        check(self._L.nvb_mapper_integrate_color(self._h, color_ptr, mask_ptr or None, mask_mode, _lib.NVB_MEM_DEVICE, rows, cols,
                                                 _fp(T), C.byref(camera.c), None, 0, None))

    def last_color_blocks(self):
        """updated_blocks of the last colour frame (after synchronize())."""
        n = C.c_int32(0)
        check(self._L.nvb_mapper_last_color_blocks(self._h, None, 0, C.byref(n)))
        out = np.empty((max(n.value, 1), 3), dtype=np.int32)
        check(self._L.nvb_mapper_last_color_blocks(self._h, _ip(out), n.value, C.byref(n)))
        return out[:n.value].copy()

    def freespace_integrator(self):
        return _FreespaceIntegrator(self)

    def update_freespace(self, update_time_ms, depth=None, T_L_C=None, camera=None, update_full_layer=False):
        """Mapper::updateFreespace(update_time_ms, T_L_C, sensor, depth_frame, update_full_layer) (mapper.h:196-214)."""
        if depth is not None:
            depth = np.ascontiguousarray(depth, dtype=np.float32)
            T = colmajor(T_L_C)
            check(self._L.nvb_mapper_update_freespace(self._h, int(update_time_ms), depth.ctypes.data, _lib.NVB_MEM_HOST,
                                                      depth.shape[0], depth.shape[1], _fp(T), C.byref(camera.c),
                                                      1 if update_full_layer else 0))
        else:
            check(self._L.nvb_mapper_update_freespace(self._h, int(update_time_ms), None, 0, 0, 0, None, None,
                                                      1 if update_full_layer else 0))

    def esdf_layer(self):
        return self._esdf

    def projective_layer_type(self):
        return self._projective_layer_type

    def tsdf_integrator(self):
        return _TsdfIntegrator(self)

    def occupancy_integrator(self):
        return _OccupancyIntegrator(self)

    def esdf_integrator(self):
        return _EsdfIntegrator(self)

    def tsdf_decay_integrator(self):
        return _DecayIntegrator(self, False)

    def occupancy_decay_integrator(self):
        return _DecayIntegrator(self, True)

    def decay(self, depth=None, T_L_C=None, camera=None, excluded_blocks=None, exclusion_center=None,
              exclusion_radius_m=None):
        """Mapper::decayTsdf / decayOccupancy on the mapper's projective layer (mapper.h:268-292). depth=None: every
        voxel decays (decay*AllVoxels); with a view, voxels that have a depth measurement in it are spared
        (decay*ExcludeLastView -- pass the last integrated frame). Returns the (n,3) indices of the deallocated
        blocks (gone from the projective and the ESDF layer)."""
        x = NvbDecayExclusion()
        keep = None
        if excluded_blocks is not None and len(excluded_blocks):
            keep = np.ascontiguousarray(excluded_blocks, dtype=np.int32).reshape(-1, 3)
            x.excluded_blocks_xyz_host, x.num_excluded_blocks = _ip(keep), keep.shape[0]
        if exclusion_center is not None and exclusion_radius_m is not None:
            x.has_exclusion_sphere = 1
            x.exclusion_center = (C.c_float * 3)(*[float(v) for v in exclusion_center])
            x.exclusion_radius_m = float(exclusion_radius_m)
        n = C.c_int32(0)
        cap = max(self._occupancy.num_blocks() if self._projective_layer_type == 1 else self._tsdf.num_blocks(), 1)
        out = np.zeros((cap, 3), dtype=np.int32)
        if depth is not None:
            depth = np.ascontiguousarray(depth, dtype=np.float32)
            T = colmajor(T_L_C)
            check(self._L.nvb_mapper_decay(self._h, C.byref(x), depth.ctypes.data, _lib.NVB_MEM_HOST, depth.shape[0],
                                           depth.shape[1], _fp(T), C.byref(camera.c), _ip(out), cap, C.byref(n)))
        else:
            check(self._L.nvb_mapper_decay(self._h, C.byref(x), None, 0, 0, 0, None, None, _ip(out), cap, C.byref(n)))
        return out[:n.value].copy()

    def decay_exclude_last_view(self):
        """Mapper::decayTsdfExcludeLastView / decayOccupancyExcludeLastView with the view kept by the mapper
        (Mapper(..., keep_last_view=True))."""
        n = C.c_int32(0)
        cap = max(self._occupancy.num_blocks() if self._projective_layer_type == 1 else self._tsdf.num_blocks(), 1)
        out = np.zeros((cap, 3), dtype=np.int32)
        check(self._L.nvb_mapper_decay_exclude_last_view(self._h, None, _ip(out), cap, C.byref(n)))
        return out[:n.value].copy()

    def decay_tsdf(self, **kw):
        assert self._projective_layer_type != ProjectiveLayerType.kOccupancy
        return self.decay(**kw)

    def decay_occupancy(self, **kw):
        assert self._projective_layer_type == ProjectiveLayerType.kOccupancy
        return self.decay(**kw)

    def cuda_stream(self):
        return self._L.nvb_mapper_stream(self._h)

    def esdf_reserved_sms(self, v=None):
        """SMs the ESDF wavefront leaves to concurrently running kernels (default 2; 4 on a multi-GPU rank that merges)."""
        if v is not None:
            check(self._L.nvb_mapper_set_esdf_reserved_sms(self._h, int(v)))
        return int(self._L.nvb_mapper_get_esdf_reserved_sms(self._h))

    def do_depth_preprocessing(self, v=None):
        """Mapper::do_depth_preprocessing (mapper.h; mapper_params.h:33-37, default off)."""
        en, n = C.c_int32(0), C.c_int32(0)
        check(self._L.nvb_mapper_get_depth_preprocessing(self._h, C.byref(en), C.byref(n)))
        if v is not None:
            check(self._L.nvb_mapper_set_depth_preprocessing(self._h, 1 if v else 0, n.value))
            return bool(v)
        return bool(en.value)

    def depth_preprocessing_num_dilations(self, v=None):
        """Mapper::depth_preprocessing_num_dilations (mapper_params.h:39-42, default 4)."""
        en, n = C.c_int32(0), C.c_int32(0)
        check(self._L.nvb_mapper_get_depth_preprocessing(self._h, C.byref(en), C.byref(n)))
        if v is not None:
            check(self._L.nvb_mapper_set_depth_preprocessing(self._h, en.value, int(v)))
            return int(v)
        return int(n.value)

    def dilate_invalid_regions_device(self, depth_ptr, out_ptr, rows, cols, num_dilations, invalid_depth_threshold=1e-2,
                                      invalid_depth_value=0.0):
        """DepthPreprocessor::dilateInvalidRegionsAsync (sensors/depth_preprocessing.h:33-38) on device images, enqueued on
        cuda_stream(); the output must not alias the input."""
        check(self._L.nvb_depth_dilate_invalid(self._h, depth_ptr, out_ptr, rows, cols, int(num_dilations),
                                               float(invalid_depth_threshold), float(invalid_depth_value)))

    def clear(self):
        check(self._L.nvb_mapper_clear(self._h))

    # --- the hot path ----------------------------------------------------
    @staticmethod
    def _frame_args(depth, mask):
        if isinstance(depth, (int, np.integer)):
            raise TypeError("pass device frames through integrate_depth_device")
        d = np.ascontiguousarray(depth, dtype=np.float32)
        mk = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        return d, mk

    def integrate_depth(self, depth, T_L_C, camera, mask=None, mask_mode=0, return_blocks=True):
        """Mapper::integrateDepth. depth: (rows, cols) float32 host array. Returns updated_blocks (n,3)."""
        d, mk = self._frame_args(depth, mask)
        T = colmajor(T_L_C)
        cap = 0
        out = None
        if return_blocks:
            cap = 1 << 14
            out = np.empty((cap, 3), dtype=np.int32)
        n = C.c_int32(0)
        check(self._L.nvb_mapper_integrate_depth(
            self._h, d.ctypes.data, None if mk is None else mk.ctypes.data, mask_mode, _lib.NVB_MEM_HOST,
            d.shape[0], d.shape[1], _fp(T), C.byref(camera.c), None if out is None else _ip(out), cap,
            C.byref(n)))
        if out is not None and n.value > cap:
            # list longer than the buffer: the frame IS integrated; read the full list back
            cap = n.value
            out = np.empty((cap, 3), dtype=np.int32)
            check(self._L.nvb_mapper_last_frame_blocks(self._h, _ip(out), cap, C.byref(n)))
        return None if out is None else out[:n.value].copy()

    def integrate_depth_device(self, depth_ptr, rows, cols, T_L_C, camera, mask_ptr=0, mask_mode=0, sync=False):
        """Same, for a frame already resident in HBM (raw device pointers). Asynchronous unless sync."""
        T = colmajor(T_L_C)
        check(self._L.nvb_mapper_integrate_depth_async(self._h, depth_ptr, mask_ptr or None, mask_mode,
                                                       _lib.NVB_MEM_DEVICE, rows, cols, _fp(T), C.byref(camera.c)))
        if sync:
            self.synchronize()

    def integrate_depth_async(self, depth, T_L_C, camera, mask=None, mask_mode=0):
        """Host frame, enqueued without synchronising (buffers are kept alive until synchronize())."""
        d, mk = self._frame_args(depth, mask)
        T = colmajor(T_L_C)
        self._keep.append((d, mk))
        check(self._L.nvb_mapper_integrate_depth_async(self._h, d.ctypes.data, None if mk is None else mk.ctypes.data,
                                                       mask_mode, _lib.NVB_MEM_HOST, d.shape[0], d.shape[1], _fp(T),
                                                       C.byref(camera.c)))

    def integrate_depth_host_ptr_async(self, depth_host_ptr, rows, cols, T_L_C, camera):
        """Host frame given as a raw (ideally pinned) pointer the caller keeps alive."""
        T = colmajor(T_L_C)
        check(self._L.nvb_mapper_integrate_depth_async(self._h, depth_host_ptr, None, 0, _lib.NVB_MEM_HOST, rows, cols,
                                                       _fp(T), C.byref(camera.c)))

    def update_esdf(self, update_full_layer=False, sync=True):
        """Mapper::updateEsdf."""
        if sync:
            check(self._L.nvb_mapper_update_esdf(self._h, 1 if update_full_layer else 0))
            self._keep.clear()
        else:
            check(self._L.nvb_mapper_update_esdf_async(self._h, 1 if update_full_layer else 0))

    def update_esdf_slice(self, update_full_layer=False, ground_plane=None):
        """Mapper::updateEsdfSlice(update_full_layer, ground_plane) (mapper.h:331-343): the 2-D ESDF on the slice layer."""
        if ground_plane is None:
            check(self._L.nvb_mapper_update_esdf_slice(self._h, 1 if update_full_layer else 0))
        else:
            pl = np.ascontiguousarray(ground_plane, dtype=np.float32).reshape(4)
            check(self._L.nvb_mapper_update_esdf_slice_planar(self._h, _fp(pl), 1 if update_full_layer else 0))

    def synchronize(self):
        check(self._L.nvb_mapper_synchronize(self._h))
        self._keep.clear()

    def join_streams(self):
        """Make later work on cuda_stream() wait for the ESDF side stream (device-side, no host sync)."""
        check(self._L.nvb_mapper_join_streams(self._h))

    def last_frame_block_count(self):
        n = C.c_int32(0)
        check(self._L.nvb_mapper_last_frame_block_count(self._h, C.byref(n)))
        return n.value

    # --- instrumentation -------------------------------------------------
    def enable_profiling(self, on=True):
        check(self._L.nvb_mapper_enable_profiling(self._h, 1 if on else 0))

    def stage_times(self, reset=False):
        ms = (C.c_double * 6)()
        calls = (C.c_int64 * 6)()
        check(self._L.nvb_mapper_stage_times(self._h, ms, calls, 1 if reset else 0))
        return {STAGE_NAMES[i]: (ms[i], calls[i]) for i in range(6)}

    def kernel_launches(self):
        return int(self._L.nvb_mapper_kernel_launches(self._h))


class ViewCalculator:
    """ViewCalculator::getBlocksInImageViewRaycast (view_calculator.h:75-80)."""

    def __init__(self, mapper):
        self._m = mapper

    def get_blocks_in_image_view_raycast(self, depth, T_L_C, camera, block_size,
                                         max_integration_distance_behind_surface_m, max_integration_distance_m):
        d = np.ascontiguousarray(depth, dtype=np.float32)
        T = colmajor(T_L_C)
        cap = 1 << 16
        while True:
            out = np.zeros((cap, 3), dtype=np.int32)
            n = C.c_int32(0)
            check(self._m._L.nvb_view_raycast(self._m._h, d.ctypes.data, _lib.NVB_MEM_HOST, d.shape[0], d.shape[1],
                                              _fp(T), C.byref(camera.c), float(block_size),
                                              float(max_integration_distance_behind_surface_m),
                                              float(max_integration_distance_m), _ip(out), cap, C.byref(n)))
            if n.value <= cap:
                return out[:n.value].copy()
            cap = n.value
