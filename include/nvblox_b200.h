/*
 * nvblox_b200.h -- C-ABI of the B200-native depth-integration hot path.
 *
 * This is the drop-in boundary for ONE path of nvblox_core:
 *   ViewCalculator::getBlocksInImageViewRaycast  ->
 *   ProjectiveTsdfIntegrator::integrateFrame     ->
 *   EsdfIntegrator::integrateBlocks
 * as driven by nvblox::Mapper::integrateDepth / Mapper::updateEsdf.
 *
 * The reference exposes this path as a C++ ABI (libnvblox_lib.so + Eigen-typed
 * headers). The library below is the C core a source-compatible `nvblox/...`
 * header set forwards to (see INTEGRATION.md and include/nvblox/): plain
 * pointers and sizes only, no Eigen / torch / STL types in any signature.
 *
 * Citations: paths are relative to
 *   /root/reference/nvblox_ros/nvblox_core/nvblox/   ("C/" in SURVEY.md)
 *
 * Conventions
 *   - Every function returns NVB_OK (0) or a negative NvbStatus; the message of
 *     the last failure on the calling thread is nvb_last_error(). (The reference
 *     aborts the process through glog CHECK on the same conditions,
 *     C/include/nvblox/core/internal/error_check.h:28-65; the C++ mirror in
 *     include/nvblox/ turns a non-zero status back into an abort.)
 *   - Transforms are 16 floats, 4x4 column-major = Eigen::Isometry3f::data()
 *     (C/include/nvblox/core/types.h:141-153).
 *   - Block indices are int32 triples (Index3D = Eigen::Vector3i).
 *   - Voxel layouts in device memory are the reference's
 *     (C/include/nvblox/map/voxels.h:28-74, blox.h:28-67): a block is
 *     voxels[8][8][8], z fastest; TsdfVoxel 8 B, EsdfVoxel 20 B.
 *   - Functions without the _async suffix are synchronous at return, like the
 *     reference (projective_integrator_impl.cuh:305, esdf_integrator.cu:258).
 *   - One NvbMapper is driven from one host thread (same as nvblox::Mapper).
 */
#ifndef NVBLOX_B200_H_
#define NVBLOX_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define NVB_API
#else
#define NVB_API __attribute__((visibility("default")))
#endif

typedef enum {
  NVB_OK = 0,
  NVB_ERR_INVALID_ARGUMENT = -1,
  NVB_ERR_CUDA = -2,
  NVB_ERR_CAPACITY = -3,      /* a slab / hash could not be grown */
  NVB_ERR_INDEX_RANGE = -4,   /* a block index does not fit the 21-bit hash key */
  NVB_ERR_NO_DEVICE = -5
} NvbStatus;

/* Where a caller buffer lives. */
typedef enum { NVB_MEM_HOST = 0, NVB_MEM_DEVICE = 1 } NvbMemory;

/* Layers of the map (C/include/nvblox/map/common_names.h TsdfLayer / EsdfLayer). */
typedef enum { NVB_LAYER_TSDF = 0, NVB_LAYER_ESDF = 1, NVB_LAYER_OCCUPANCY = 2, NVB_LAYER_FREESPACE = 3, NVB_LAYER_COLOR = 4, NVB_LAYER_MESH = 5 } NvbLayer;

/* ProjectiveLayerType of a Mapper (C/include/nvblox/mapper/mapper.h:40-48): which layer integrateDepth feeds. */
typedef enum {
  NVB_PROJECTIVE_TSDF = 0,
  NVB_PROJECTIVE_OCCUPANCY = 1,
  NVB_PROJECTIVE_TSDF_WITH_FREESPACE = 2 /* TSDF + FreespaceLayer (dynablox), used by the ESDF to ignore free voxels */
} NvbProjectiveLayerType;

/* nvblox::Camera (C/include/nvblox/sensors/camera.h:193-203) with its
 * std::optional<RadialTangentialDistortionParams> (C/include/nvblox/sensors/distortion.h:24-62):
 * has_distortion = 0 is std::nullopt. */
typedef struct {
  float fu, fv, cu, cv;
  int32_t width, height;
  int32_t has_distortion;
  float k1, k2, k3, k4, k5, k6; /* radial: numerator k1..k3, denominator k4..k6 */
  float p1, p2;                 /* tangential */
} NvbCamera;

/* WeightingFunctionType (C/include/nvblox/integrators/weighting_function.h:11-18). */
typedef enum {
  NVB_WEIGHT_CONSTANT = 0,
  NVB_WEIGHT_CONSTANT_DROPOFF = 1,
  NVB_WEIGHT_INVERSE_SQUARE = 2,
  NVB_WEIGHT_INVERSE_SQUARE_DROPOFF = 3,
  NVB_WEIGHT_INVERSE_SQUARE_TSDF_DISTANCE_PENALTY = 4,
  NVB_WEIGHT_LINEAR_WITH_MAX = 5
} NvbWeighting;

/* WorkspaceBoundsType (C/include/nvblox/geometry/workspace_bounds.h:24). */
typedef enum { NVB_WS_UNBOUNDED = 0, NVB_WS_HEIGHT_BOUNDS = 1, NVB_WS_BOUNDING_BOX = 2 } NvbWorkspaceBounds;

/* MaskMode (C/include/nvblox/sensors/image.h:383). */
typedef enum { NVB_MASK_NON_INVERTED = 0, NVB_MASK_INVERTED = 1 } NvbMaskMode;

/* ProjectiveIntegrator / ProjectiveTsdfIntegrator / ViewCalculator parameters
 * (C/include/nvblox/integrators/projective_integrator_params.h:24-63,
 *  view_calculator_params.h:22-60). */
typedef struct {
  float truncation_distance_vox;    /* 4   */
  float max_integration_distance_m; /* 7   */
  float max_weight;                 /* 5   */
  float invalid_depth_decay_factor; /* -1 (off) */
  int32_t weighting_type;           /* NVB_WEIGHT_INVERSE_SQUARE */
  int32_t raycast_subsampling;      /* 4   */
  int32_t workspace_bounds_type;    /* NVB_WS_UNBOUNDED */
  float workspace_min[3];
  float workspace_max[3];
} NvbTsdfParams;

/* EsdfIntegrator parameters (C/include/nvblox/integrators/esdf_integrator_params.h:22-31). */
typedef struct {
  float max_esdf_distance_m;   /* 2    */
  float max_site_distance_vox; /* 1    */
  float min_weight;            /* 1e-4 */
  float occupied_threshold;    /* 0.5: probability above which an occupancy voxel is inside an obstacle
                                  (EsdfIntegrator::occupied_threshold, esdf_integrator.h:198,219,375) */
} NvbEsdfParams;

/* ProjectiveOccupancyIntegrator's inverse sensor model
 * (C/include/nvblox/integrators/occupancy_integrator_params.h:21-40). */
typedef struct {
  float free_region_occupancy_probability;       /* 0.3 */
  float occupied_region_occupancy_probability;   /* 0.7 */
  float unobserved_region_occupancy_probability; /* 0.5 */
  float occupied_region_half_width_m;            /* 0.1 */
} NvbOccupancyParams;

/* TsdfVoxel / EsdfVoxel as stored in HBM (C/include/nvblox/map/voxels.h:28-34,55-74). */
typedef struct {
  float distance;
  float weight;
} NvbTsdfVoxel;

typedef struct {
  float squared_distance_vox;
  int32_t parent_direction[3];
  uint8_t is_inside, observed, is_site, pad_;
} NvbEsdfVoxel;

/* FreespaceVoxel (C/include/nvblox/map/voxels.h:38-52): two nvblox::Time (int64 milliseconds) and a bool. */
typedef struct {
  int64_t last_occupied_timestamp_ms;
  int64_t consecutive_occupancy_duration_ms;
  uint8_t is_high_confidence_freespace, pad_[7];
} NvbFreespaceVoxel;

/* ColorVoxel (C/include/nvblox/map/voxels.h:77-83): Color = 3 bytes RGB (core/color.h:28-68), one byte of padding, float weight.
 * A freshly allocated block holds ColorVoxel() = Color::Gray() (127, 127, 127), weight 0. */
typedef struct {
  uint8_t r, g, b, pad_;
  float weight;
} NvbColorVoxel;

/* Construction options. capacity = number of 8x8x8 blocks each layer's slab is
 * sized for up front (it grows by doubling, which is a synchronising event;
 * BlockMemoryPool does the same, C/include/nvblox/map/internal/impl/
 * block_memory_pool_impl.h:30-73). 0 selects the default (65536). */
typedef struct {
  float voxel_size_m;
  int32_t device;                /* CUDA device ordinal */
  int32_t tsdf_capacity_blocks;
  int32_t esdf_capacity_blocks;
  int32_t esdf_persistent;       /* how computeEsdf runs. 3 (default): exchange-slab wavefront, whole update in one cooperative
                                    launch, one grid barrier per ring (adds two ESDF-sized slabs + the candidate records);
                                    1: four-phase wavefront, one cooperative launch, four barriers per ring;
                                    2: gather-replay wavefront, two barriers per ring;
                                    0: one launch per ring phase with a host-read counter, like the reference */
  int32_t projective_layer_type; /* NvbProjectiveLayerType: TSDF (default) or occupancy */
  int32_t keep_last_view;        /* 1: every integrated frame leaves a device copy of its depth image, pose and camera
                                    behind for nvb_mapper_decay_exclude_last_view, like Mapper::integrateDepth does
                                    (mapper_impl.h:70-78); 0 (default): no copy, pass the view to nvb_mapper_decay */
} NvbMapperOptions;

/* = nvblox::Mapper restricted to {TsdfLayer, EsdfLayer, ProjectiveTsdfIntegrator,
 * EsdfIntegrator, BlocksToUpdateTracker(kEsdf)} (C/include/nvblox/mapper/mapper.h:107-836). */
typedef struct NvbMapper NvbMapper;

NVB_API const char* nvb_last_error(void);
NVB_API const char* nvb_version(void);
/* Number of CUDA devices visible (0 if none / no driver). */
NVB_API int32_t nvb_device_count(void);

NVB_API void nvb_default_mapper_options(NvbMapperOptions* opts);
NVB_API void nvb_default_tsdf_params(NvbTsdfParams* p);
NVB_API void nvb_default_esdf_params(NvbEsdfParams* p);

/* Mapper::Mapper(voxel_size_m, ...) (mapper.h:119-124). */
NVB_API int32_t nvb_mapper_create(const NvbMapperOptions* opts, NvbMapper** out);
NVB_API void nvb_mapper_destroy(NvbMapper* m);
/* LayerCake::clear + tracker reset. */
NVB_API int32_t nvb_mapper_clear(NvbMapper* m);
/* Mapper::tsdf_integrator().<setters> (projective_tsdf_integrator.h:59-121,
 * projective_integrator.h:56-85, view_calculator.h:88-145). */
NVB_API int32_t nvb_mapper_set_tsdf_params(NvbMapper* m, const NvbTsdfParams* p);
NVB_API int32_t nvb_mapper_get_tsdf_params(const NvbMapper* m, NvbTsdfParams* p);
/* Mapper::esdf_integrator().<setters> (esdf_integrator.h:178-283). */
NVB_API int32_t nvb_mapper_set_esdf_params(NvbMapper* m, const NvbEsdfParams* p);
NVB_API int32_t nvb_mapper_get_esdf_params(const NvbMapper* m, NvbEsdfParams* p);
/* Mapper::occupancy_integrator().<setters> (C/include/nvblox/integrators/projective_occupancy_integrator.h:57-88).
 * With an occupancy mapper, nvb_mapper_integrate_depth runs ProjectiveOccupancyIntegrator::integrateFrame
 * (same raycast + block list, UpdateOccupancyVoxelFunctor, projective_occupancy_integrator_impl.cuh:27-73) and
 * nvb_mapper_update_esdf runs EsdfIntegrator::integrateBlocks(OccupancyLayer, ...) (esdf_integrator.h:72-80). */
NVB_API void nvb_default_occupancy_params(NvbOccupancyParams* p);
NVB_API int32_t nvb_mapper_set_occupancy_params(NvbMapper* m, const NvbOccupancyParams* p);
NVB_API int32_t nvb_mapper_get_occupancy_params(const NvbMapper* m, NvbOccupancyParams* p);

/* TsdfDecayIntegrator / OccupancyDecayIntegrator parameters (C/include/nvblox/integrators/tsdf_decay_integrator_params.h:21-48,
 * occupancy_decay_integrator_params.h:21-43, internal/decay_integrator_base_params.h:22-29; setters
 * tsdf_decay_integrator.h:73-101, occupancy_decay_integrator.h:72-101). */
typedef struct NvbTsdfDecayParams {
  float decay_factor;                   /* 0.95: weight *= decay_factor                                  */
  float decayed_weight_threshold;       /* 1e-3: weights never decay below it; below = "fully decayed" */
  int32_t set_free_distance_on_decayed; /* 0                                                            */
  float free_distance_vox;              /* 4                                                            */
  int32_t deallocate_decayed_blocks;    /* 1: blocks whose voxels are all fully decayed leave the map   */
} NvbTsdfDecayParams;
typedef struct NvbOccupancyDecayParams {
  float free_region_decay_probability;     /* 0.55, in [0.5, 1]  */
  float occupied_region_decay_probability; /* 0.4, in [0, 0.5)   */
  float decay_to_probability;              /* 0.5 (decay_to_free(true): 0.49, occupancy_decay_integrator.h:35-36) */
  int32_t deallocate_decayed_blocks;       /* 1                  */
} NvbOccupancyDecayParams;
/* DecayBlockExclusionOptions (C/include/nvblox/integrators/internal/decayer.h:31-44): blocks that are spared. */
typedef struct NvbDecayExclusion {
  const int32_t* excluded_blocks_xyz_host; /* may be NULL */
  int32_t num_excluded_blocks;
  int32_t has_exclusion_sphere;            /* blocks whose origin is within the sphere are spared */
  float exclusion_center[3];
  float exclusion_radius_m;
} NvbDecayExclusion;
NVB_API void nvb_default_tsdf_decay_params(NvbTsdfDecayParams* p);
NVB_API int32_t nvb_mapper_set_tsdf_decay_params(NvbMapper* m, const NvbTsdfDecayParams* p);
NVB_API int32_t nvb_mapper_get_tsdf_decay_params(const NvbMapper* m, NvbTsdfDecayParams* p);
NVB_API void nvb_default_occupancy_decay_params(NvbOccupancyDecayParams* p);
NVB_API int32_t nvb_mapper_set_occupancy_decay_params(NvbMapper* m, const NvbOccupancyDecayParams* p);
NVB_API int32_t nvb_mapper_get_occupancy_decay_params(const NvbMapper* m, NvbOccupancyDecayParams* p);
/* The constant-z slice of the 2-D ESDF (esdf_slice_min_height / esdf_slice_max_height / esdf_slice_height,
 * C/include/nvblox/integrators/esdf_integrator_params.h:33-43; setters esdf_integrator.h:216-256). */
typedef struct NvbEsdfSliceParams {
  float slice_min_height_m; /* 0 */
  float slice_max_height_m; /* 1 */
  float slice_height_m;     /* 1: z of the output slice */
  /* planar slices (slice_height_above_plane_m / slice_height_thickness_m, esdf_integrator_params.h:45-52) */
  float slice_height_above_plane_m; /* 0   */
  float slice_height_thickness_m;   /* 0.1 */
} NvbEsdfSliceParams;
NVB_API void nvb_default_esdf_slice_params(NvbEsdfSliceParams* p);
NVB_API int32_t nvb_mapper_set_esdf_slice_params(NvbMapper* m, const NvbEsdfSliceParams* p);
NVB_API int32_t nvb_mapper_get_esdf_slice_params(const NvbMapper* m, NvbEsdfSliceParams* p);

/* FreespaceIntegrator parameters (C/include/nvblox/integrators/freespace_integrator_params.h:22-58; setters
 * freespace_integrator.h:75-128). */
typedef struct NvbFreespaceParams {
  float max_tsdf_distance_for_occupancy_m;                 /* 0.15 */
  int64_t max_unobserved_to_keep_consecutive_occupancy_ms; /* 200  */
  int64_t min_duration_since_occupied_for_freespace_ms;    /* 1000 */
  int64_t min_consecutive_occupancy_duration_for_reset_ms; /* 2000 */
  int32_t check_neighborhood;                              /* 1    */
  int32_t initialize_to_high_confidence_freespace;         /* 0    */
} NvbFreespaceParams;
NVB_API void nvb_default_freespace_params(NvbFreespaceParams* p);
NVB_API int32_t nvb_mapper_set_freespace_params(NvbMapper* m, const NvbFreespaceParams* p);
NVB_API int32_t nvb_mapper_get_freespace_params(const NvbMapper* m, NvbFreespaceParams* p);
NVB_API float nvb_mapper_voxel_size(const NvbMapper* m);
NVB_API float nvb_mapper_block_size(const NvbMapper* m);

/* ViewCalculator::getBlocksInImageViewRaycast<Camera> (view_calculator.h:75-80,
 * view_calculator_impl.cuh:117-198). Does not touch the map. Writes up to
 * cap triples to out_xyz_host (x-fastest order inside the view AABB) and the
 * full count to *out_count. */
NVB_API int32_t nvb_view_raycast(NvbMapper* m, const float* depth, int32_t depth_memory,
                                 int32_t rows, int32_t cols, const float* T_L_C,
                                 const NvbCamera* cam, float block_size,
                                 float max_integration_distance_behind_surface_m,
                                 float max_integration_distance_m, int32_t* out_xyz_host,
                                 int32_t cap, int32_t* out_count);

/* Mapper::markUnobservedTsdfFreeInsideRadius(center, radius) (C/include/nvblox/mapper/mapper.h:352-356, src/mapper/mapper.cpp:494-507)
 * = Projective{Tsdf,Occupancy}Integrator::markUnobservedFreeInsideRadius (projective_integrator_impl.cuh:408-462): every block
 * whose box is closer than `radius` to `center` is allocated in the projective layer and its unobserved voxels are set to
 * slightly observed free space (TSDF: truncation distance, weight 0.1 where weight < 1e-3; occupancy: log odds -2e-4 where
 * |log odds| < 1e-4); the blocks join the tracker, so the next ESDF update covers them. updated_xyz_host may be NULL; it receives up to cap triples (unordered). */
NVB_API int32_t nvb_mapper_mark_unobserved_free_inside_radius(NvbMapper* m, const float center[3], float radius,
                                                              int32_t* updated_xyz_host, int32_t cap, int32_t* out_count);

/* ProjectiveColorIntegrator's parameters (C/include/nvblox/integrators/projective_appearance_integrator.h:96-175,
 * projective_integrator_params.h:24-75), with those of its SphereTracer (rays/sphere_tracer.h:204-218) and of its own
 * ViewCalculator's workspace bounds. sphere_tracer_maximum_ray_length_m is a separate field because the reference
 * copies max_integration_distance_m into the tracer in the constructor only
 * (src/integrators/projective_appearance_integrator.cu:61). */
typedef struct {
  float max_integration_distance_m;                 /* 7.0 */
  float truncation_distance_vox;                    /* 4.0 */
  float max_weight;                                 /* 5.0 */
  float measurement_weight;                         /* 0.8, in (0, 1] */
  int32_t sphere_tracing_ray_subsampling_factor;    /* 4; must divide the image size */
  int32_t sphere_tracer_maximum_steps;              /* 100 */
  float sphere_tracer_maximum_ray_length_m;         /* 7.0 */
  float sphere_tracer_surface_distance_epsilon_vox; /* 0.1 */
  int32_t workspace_bounds_type;                    /* NVB_WS_UNBOUNDED */
  float workspace_min[3], workspace_max[3];
} NvbColorParams;
NVB_API void nvb_default_color_params(NvbColorParams* p);
NVB_API int32_t nvb_mapper_set_color_params(NvbMapper* m, const NvbColorParams* p);
NVB_API int32_t nvb_mapper_get_color_params(const NvbMapper* m, NvbColorParams* p);

/* Mapper::integrateColor(MaskedColorImageConstView, T_L_C, Camera) (mapper.h:202-207, mapper_impl.h:104-130) =
 * ProjectiveColorIntegrator::integrateFrame (src/integrators/projective_appearance_integrator.cu:68-165): the TSDF blocks
 * in the camera's view (ViewCalculator::getBlocksInImageViewProjection) that touch the truncation band get a colour block;
 * a sphere-traced synthetic depth image of the TSDF layer (SphereTracer::renderImageOnGPU) decides occlusion; visible voxels
 * blend the bilinearly interpolated pixel in (UpdateAppearanceVoxelFunctor). color = rows * cols * 3 bytes, RGB, row-major;
 * mask (rows * cols, may be NULL) and mask_mode as for depth. With an occupancy mapper the call does nothing, as in the
 * reference. updated_xyz_host may be NULL; otherwise it receives up to cap triples of `updated_blocks` (unordered) and
 * *out_count the full count. */
NVB_API int32_t nvb_mapper_integrate_color(NvbMapper* m, const uint8_t* color, const uint8_t* mask, int32_t mask_mode,
                                           int32_t memory, int32_t rows, int32_t cols, const float* T_L_C,
                                           const NvbCamera* cam, int32_t* updated_xyz_host, int32_t cap, int32_t* out_count);
/* A call with device-resident images (memory = NVB_MEM_DEVICE) and no output pointers is enqueued without synchronising;
 * after nvb_mapper_synchronize, this returns the `updated_blocks` of the last colour frame. */
NVB_API int32_t nvb_mapper_last_color_blocks(NvbMapper* m, int32_t* out_xyz_host, int32_t cap, int32_t* out_count);
/* SphereTracer::renderImageOnGPU(camera, T_L_C, tsdf_layer, truncation_distance_m, &depth, ..., ray_subsampling_factor)
 * (C/src/rays/sphere_tracer.cu:389-485) with the colour integrator's tracer settings: out_depth_host receives
 * (height / f) * (width / f) floats, -1 where a ray found no surface. */
NVB_API int32_t nvb_sphere_tracer_render_depth(NvbMapper* m, const float* T_L_C, const NvbCamera* cam,
                                               float truncation_distance_m, int32_t ray_subsampling_factor,
                                               float* out_depth_host);

/* Mapper::integrateDepth(MaskedDepthImageConstView, T_L_C, Camera)
 * (mapper.h:167-172, mapper_impl.h:28-81) =
 * ProjectiveTsdfIntegrator::integrateFrame (projective_tsdf_integrator.h:48-52)
 * + BlocksToUpdateTracker::addBlocksToUpdate. mask may be NULL
 * (kMaskActiveEverywhere). updated_xyz_host may be NULL; otherwise it receives
 * up to cap triples of `updated_blocks` and *out_count the full count. */
NVB_API int32_t nvb_mapper_integrate_depth(NvbMapper* m, const float* depth, const uint8_t* mask,
                                           int32_t mask_mode, int32_t memory, int32_t rows,
                                           int32_t cols, const float* T_L_C, const NvbCamera* cam,
                                           int32_t* updated_xyz_host, int32_t cap,
                                           int32_t* out_count);

/* Same work, enqueued on the mapper's stream; returns without synchronising.
 * Host depth/mask buffers must stay valid (and should be pinned) until
 * nvb_mapper_synchronize. The per-frame block count can be read afterwards with
 * nvb_mapper_last_frame_block_count. */
NVB_API int32_t nvb_mapper_integrate_depth_async(NvbMapper* m, const float* depth,
                                                 const uint8_t* mask, int32_t mask_mode,
                                                 int32_t memory, int32_t rows, int32_t cols,
                                                 const float* T_L_C, const NvbCamera* cam);

/* Mapper::updateEsdf(UpdateFullLayer) (mapper.h:326, src/mapper/mapper.cpp:408-430):
 * EsdfIntegrator::integrateBlocks over the blocks touched since the last call
 * (all TSDF blocks on the first call or when update_full_layer != 0). */
NVB_API int32_t nvb_mapper_update_esdf(NvbMapper* m, int32_t update_full_layer);
NVB_API int32_t nvb_mapper_update_esdf_async(NvbMapper* m, int32_t update_full_layer);

/* Mapper::decayTsdf / decayOccupancy (mapper.h:268-292; mapper_impl.h:190-265; VoxelDecayer::decay,
 * C/include/nvblox/integrators/internal/cuda/impl/decayer_impl.cuh:150-262) on the mapper's projective layer.
 * depth == NULL: decay*AllVoxels. depth != NULL: decay*ExcludeLastView -- voxels that have a depth measurement in
 * the given view (doesVoxelHaveDepthMeasurement, projective_integrators_common_impl.cuh:58-101, with the
 * integrator's max integration distance and truncation distance) are spared; the caller passes the view it wants
 * excluded (the reference's Mapper keeps a copy of the last frame for this). `exclusion` (may be NULL) spares whole
 * blocks. Fully decayed blocks are deallocated when the parameters say so, from the projective AND the ESDF
 * layer (Mapper::clearBlocksInLayers, src/mapper/mapper.cpp:546-575), their indices are written to
 * removed_xyz_host (up to cap; may be NULL) and counted in *out_count; the next nvb_mapper_update_esdf covers all
 * blocks (BlocksToUpdateTracker::addAllBlocksToUpdate). Synchronous. */
NVB_API int32_t nvb_mapper_decay(NvbMapper* m, const NvbDecayExclusion* exclusion, const float* depth,
                                 int32_t depth_memory, int32_t rows, int32_t cols, const float* T_L_C,
                                 const NvbCamera* cam, int32_t* removed_xyz_host, int32_t cap, int32_t* out_count);

/* Mapper::updateFreespace(update_time_ms, T_L_C, sensor, depth_frame, update_full_layer) (mapper.h:196-214,
 * mapper_impl.h:152-188) on a NVB_PROJECTIVE_TSDF_WITH_FREESPACE mapper: FreespaceIntegrator::updateFreespaceLayer
 * (C/include/nvblox/integrators/internal/cuda/impl/freespace_integrator_impl.cuh:99-383) over the TSDF blocks touched
 * since the last call (all on the first call / update_full_layer). depth != NULL: only voxels with a depth measurement
 * in that view are updated (max view distance = the integrator's max integration distance, truncation = 2 x the
 * truncation distance, mapper_impl.h:157-172); depth == NULL: no viewpoint exclusion. Synchronous. The following
 * nvb_mapper_update_esdf treats high-confidence-free voxels as outside (esdf_integrator.cu:401-415). */
NVB_API int32_t nvb_mapper_update_freespace(NvbMapper* m, int64_t update_time_ms, const float* depth, int32_t depth_memory,
                                            int32_t rows, int32_t cols, const float* T_L_C, const NvbCamera* cam,
                                            int32_t update_full_layer);
/* FreespaceIntegrator::updateFreespaceLayer on an explicit block list (freespace_integrator.h:60-66); max_view_distance_m
 * / truncation_distance_m <= 0 mean "unset" (no limit). Does not consult or reset the tracker. */
NVB_API int32_t nvb_freespace_update_blocks(NvbMapper* m, const int32_t* blocks_xyz_host, int32_t num_blocks,
                                            int64_t update_time_ms, const float* depth, int32_t depth_memory, int32_t rows,
                                            int32_t cols, const float* T_L_C, const NvbCamera* cam,
                                            float max_view_distance_m, float truncation_distance_m);

/* Mapper::decayTsdfExcludeLastView / decayOccupancyExcludeLastView (mapper.h:218-230) with the view the mapper kept
 * (NvbMapperOptions.keep_last_view); decays every voxel if no frame was integrated yet, like the reference. */
NVB_API int32_t nvb_mapper_decay_exclude_last_view(NvbMapper* m, const NvbDecayExclusion* exclusion,
                                                   int32_t* removed_xyz_host, int32_t cap, int32_t* out_count);

/* Mapper::updateEsdfSlice (mapper.h:331-343; EsdfMode::k2D) = EsdfIntegrator::integrateSlice with the constant-z
 * slice description (C/src/integrators/esdf_integrator.cu:283-347, markSitesInSlice :754-1055): the band
 * [slice_min_height, slice_max_height] of the projective layer (honouring the freespace layer if the mapper has one) is
 * squashed onto ONE layer of ESDF blocks at slice_height; clear + computeEsdf then run on that layer. A mapper's ESDF
 * layer is either 3-D or 2-D: mixing nvb_mapper_update_esdf and nvb_mapper_update_esdf_slice is an error, like the
 * reference's EsdfMode check. Synchronous. */
NVB_API int32_t nvb_mapper_update_esdf_slice(NvbMapper* m, int32_t update_full_layer);
/* The same with a PlanarSliceDescription (Mapper::updateEsdfSlice(..., ground_plane), mapper.h:343;
 * EsdfIntegrator::integrateSlice(layer, blocks, ground_plane, esdf), esdf_integrator.h:120-150): the band starts
 * slice_height_above_plane_m above the plane n . p + d = 0 and is slice_height_thickness_m thick, per voxel column
 * (PlanarSliceColumnBoundsGetter). plane = {nx, ny, nz, d} = Plane::normal() (unit length) and Plane::offset(); a
 * near-vertical plane (|nz| < 1e-4) falls back to z = 0 like checkForVerticalPlane. */
NVB_API int32_t nvb_mapper_update_esdf_slice_planar(NvbMapper* m, const float plane[4], int32_t update_full_layer);
NVB_API int32_t nvb_esdf_integrate_slice_planar_blocks(NvbMapper* m, const float plane[4], const int32_t* blocks_xyz_host,
                                                       int32_t num_blocks);
/* EsdfIntegrator::integrateSlice(layer, block_indices, esdf_layer) on an explicit block list (esdf_integrator.h:96-118). */
NVB_API int32_t nvb_esdf_integrate_slice_blocks(NvbMapper* m, const int32_t* blocks_xyz_host, int32_t num_blocks);

/* EsdfSlicer::sliceLayerToDistanceImage (C/include/nvblox/integrators/esdf_slicer.h:52-78, C/src/integrators/esdf_slicer.cu:
 * 25-67,112-215) and, if grid_host != NULL, EsdfSlicer::occupancyGridFromSliceImage (:78-110,254-300) of the ESDF layer
 * (3-D or 2-D) at slice_height_m: one pixel per voxel over the AABB of the ESDF blocks at that height, rows along y,
 * columns along x; value = distance in metres (negative inside), `unobserved_value` where nothing is known; grid = 100
 * occupied (distance < 0.01), 0 free, -1 unknown. Writes the AABB (min xyz, max xyz), the image size, and up to
 * cap_pixels pixels into the host buffers (either may be NULL to query the size). rows = cols = 0 if the layer has no
 * block at that height. */
NVB_API int32_t nvb_esdf_slice_distance_image(NvbMapper* m, float slice_height_m, float unobserved_value,
                                              float aabb_out[6], float* image_host, int8_t* grid_host, int32_t cap_pixels,
                                              int32_t* rows_out, int32_t* cols_out);
/* EsdfSlicer::getAabbOfLayerAtHeight (C/src/integrators/esdf_slicer.cu:112-147): the box of the ESDF blocks at the slice's z
 * index; *empty_out = 1 (aabb_out untouched) if there is none. */
NVB_API int32_t nvb_esdf_slice_aabb(NvbMapper* m, float slice_height_m, float aabb_out[6], int32_t* empty_out);
/* EsdfSlicer::sliceLayerToDistanceImage(layer, slice_height, unobserved_value, aabb, image) on a GIVEN box (:169-199): the
 * building block of sliceLayersToCombinedDistanceImage (C/include/nvblox/integrators/esdf_slicer.h:78-118, esdf_slicer.cu:
 * 201-240), which slices two layers (two mappers here) on the merged box of their slices and takes the element-wise minimum. */
NVB_API int32_t nvb_esdf_slice_distance_image_in_aabb(NvbMapper* m, float slice_height_m, float unobserved_value, const float aabb[6],
                                                      float* image_host, int8_t* grid_host, int32_t cap_pixels, int32_t* rows_out,
                                                      int32_t* cols_out);

/* EsdfIntegrator::integrateBlocks(const TsdfLayer&, const std::vector<Index3D>&, EsdfLayer*)
 * (esdf_integrator.h:56-58, src/integrators/esdf_integrator.cu:220-266) on an
 * explicit block list; does not consult or reset the tracker. */
NVB_API int32_t nvb_esdf_integrate_blocks(NvbMapper* m, const int32_t* blocks_xyz_host,
                                          int32_t num_blocks);

/* cudaStreamSynchronize on the mapper's stream + deferred error check. */
NVB_API int32_t nvb_mapper_synchronize(NvbMapper* m);
NVB_API int32_t nvb_mapper_last_frame_block_count(NvbMapper* m, int32_t* out_count);
/* `updated_blocks` of the most recent frame again (e.g. after a too-small buffer). */
NVB_API int32_t nvb_mapper_last_frame_blocks(NvbMapper* m, int32_t* out_xyz_host, int32_t cap,
                                             int32_t* out_count);
/* Multi-GPU merge step (no reference counterpart: the reference is single-GPU): sorted unique
 * union of block-index lists. xyz_dev holds n int32 triples in device memory (e.g. the NCCL all-gather of the
 * ranks' padded lists; a triple whose x is INT32_MIN is padding), aabb_min/max bound the union. Writes up to cap
 * triples to out_xyz_dev in the view calculator's order (x fastest, then y, then z) and the count to the host. */
NVB_API int32_t nvb_blocks_union(NvbMapper* m, const int32_t* xyz_dev, int32_t n, const int32_t aabb_min[3],
                                 const int32_t aabb_max[3], int32_t* out_xyz_dev, int32_t cap,
                                 int32_t* out_count_host);

/* ViewCalculator::cache_last_viewpoint (C/include/nvblox/integrators/view_calculator.h:196; C/src/integrators/view_calculator.cu:82-88),
 * on by default like in the reference: nvb_mapper_integrate_depth* reuses the block list of one of the last two frames whose
 * pose (1 mm, 0.1 degree) and sensor match -- whatever the depth image holds (ViewpointCache, view_calculator.h:211-244). */
NVB_API int32_t nvb_mapper_set_cache_last_viewpoint(NvbMapper* m, int32_t enable);
NVB_API int32_t nvb_mapper_get_cache_last_viewpoint(const NvbMapper* m);

/* Scheduling knob with no counterpart in the reference: the cooperative ESDF wavefront launches num_SMs - reserved_sms CTAs
 * (default 2), leaving room for the kernels that run concurrently with it -- the next frame's raycast / TSDF chain and, on
 * a multi-GPU rank, the NCCL all-gather of the block-list merge (4 there). Results do not depend on it. */
NVB_API int32_t nvb_mapper_set_esdf_reserved_sms(NvbMapper* m, int32_t reserved_sms);
NVB_API int32_t nvb_mapper_get_esdf_reserved_sms(const NvbMapper* m);

/* ---- Mesh (C/include/nvblox/mesh/mesh_integrator.h:39-162, C/src/mesh/mesh_integrator.cu,
 * C/src/mesh/mesh_integrator_appearance.cu). The mesh layer holds, per VoxelBlock that has triangles, vertices (3 floats),
 * flat per-vertex normals (3 floats), triangle indices into the block's vertices (triplets) and -- after a colour update --
 * one RGBA colour per vertex (MeshBlock, C/include/nvblox/mesh/mesh_block.h:32-83). A block's triangles come out in
 * x-major voxel order (the reference's order within a block is an atomicAdd race, marching_cubes_impl.cuh:11-29). */
typedef struct {
  float min_weight;          /* mesh_integrator_min_weight, 1e-4 (C/include/nvblox/mesh/mesh_integrator_params.h:22-24) */
  int32_t weld_vertices;     /* mesh_integrator_weld_vertices, true (mesh_integrator_params.h:25-27) */
  float cutoff_distance_vox; /* MeshIntegrator::cutoff_distance_vox_, 5 (mesh_integrator.h:129) */
} NvbMeshParams;
NVB_API void nvb_default_mesh_params(NvbMeshParams* p);
NVB_API int32_t nvb_mapper_set_mesh_params(NvbMapper* m, const NvbMeshParams* p);
NVB_API int32_t nvb_mapper_get_mesh_params(const NvbMapper* m, NvbMeshParams* p);
/* Mapper::updateColorMesh(UpdateFullLayer) (C/src/mapper/mapper.cpp:371-406): re-meshes the blocks touched since the last
 * call (or every block) and colours them from the colour layer; a no-op for an occupancy mapper. */
NVB_API int32_t nvb_mapper_update_mesh(NvbMapper* m, int32_t update_full_layer);
/* MeshIntegrator::integrateBlocksGPU (mesh_integrator.cu:66-108) on an explicit list of block indices (host, triples; the
 * ones missing from the TSDF layer are skipped), optionally followed by MeshIntegrator::updateAppearance on the same list. */
NVB_API int32_t nvb_mesh_integrate_blocks(NvbMapper* m, const int32_t* blocks_xyz_host, int32_t num_blocks, int32_t update_color);
/* MeshIntegrator::updateAppearance(colour layer, block list, mesh layer) (mesh_integrator_appearance.cu:71-96,281-380). */
NVB_API int32_t nvb_mesh_update_color(NvbMapper* m, const int32_t* blocks_xyz_host, int32_t num_blocks);
/* Sizes of the listed mesh blocks: sizes_out[3 i ..] = {vertices, triangle indices, colours}, or -1s if block i has no mesh
 * block. The block indices of the layer: nvb_layer_block_indices(m, NVB_LAYER_MESH, ...). */
NVB_API int32_t nvb_mesh_block_sizes(NvbMapper* m, const int32_t* blocks_xyz_host, int32_t num_blocks, int32_t* sizes_out);
/* The listed mesh blocks packed back to back in list order into host buffers (any of them may be NULL): 3 floats per
 * vertex / normal, one int32 per triangle index (relative to the block's first vertex), 4 bytes RGBA per colour.
 * caps = capacities of the buffers in {vertices, triangle indices, colours}. */
NVB_API int32_t nvb_mesh_get_blocks(NvbMapper* m, const int32_t* blocks_xyz_host, int32_t num_blocks, float* vertices_out,
                                    float* normals_out, int32_t* triangles_out, uint8_t* colors_out, const int64_t caps[3]);
/* out = {arena capacity, fill level, vertices emitted by the last update, reserved} in vertices. */
NVB_API int32_t nvb_mesh_arena_stats(NvbMapper* m, int64_t out[4]);

/* Mapper::do_depth_preprocessing / depth_preprocessing_num_dilations (C/include/nvblox/mapper/mapper_params.h:33-42, defaults
 * off / 4; Mapper::preprocessDepthImageAsync, C/src/mapper/mapper.cpp:335-352): when enabled, nvb_mapper_integrate_depth*
 * integrates (and keeps as the last view) a copy of the depth image whose invalid regions were dilated. */
NVB_API int32_t nvb_mapper_set_depth_preprocessing(NvbMapper* m, int32_t enable, int32_t num_dilations);
NVB_API int32_t nvb_mapper_get_depth_preprocessing(const NvbMapper* m, int32_t* enable, int32_t* num_dilations);
/* DepthPreprocessor::dilateInvalidRegionsAsync (C/src/sensors/depth_preprocessing.cpp; C/include/nvblox/sensors/
 * depth_preprocessing.h): pixels with depth < invalid_depth_threshold (the class default is 1e-2) are invalid; the invalid
 * mask is dilated num_dilations times with a 3x3 structuring element (replicated border) and every masked pixel is written
 * as invalid_depth_value (default 0). Device pointers, rows*cols floats, out_dev must not alias depth_dev; enqueued on the
 * mapper's stream. */
NVB_API int32_t nvb_depth_dilate_invalid(NvbMapper* m, const float* depth_dev, float* out_dev, int32_t rows, int32_t cols,
                                         int32_t num_dilations, float invalid_depth_threshold, float invalid_depth_value);

/* Device-resident merge of the ranks' updated-block lists (multi-GPU; the reference has no counterpart:
 * it is single-GPU and keeps `updated_blocks` in a host std::vector, C/include/nvblox/mapper/internal/impl/mapper_impl.h:40-60).
 * A SEGMENT is an int32 device array [count, x0, y0, z0, x1, ...] of capacity cap_entries (1 + 3 * cap_entries ints).
 *  - nvb_mapper_append_frame_blocks: appends the block list of the last integrated frame to a segment, enqueued on the mapper's
 *    stream (no host synchronisation; the caller zeroes segment[0] when a batch starts);
 *  - nvb_blocks_union_segments: sorted unique union (x fastest, then y, then z) of num_segments gathered segments laid out every
 *    segment_stride_ints ints -- AABB, bitset marking and ordered compaction all sized on the device -- written to
 *    out_xyz_dev / out_count_dev; enqueued on `stream` (a cudaStream_t; NULL = the mapper's stream), no host synchronisation;
 *  - nvb_blocks_union_status: synchronises and reports whether a union overflowed its bitset (1) -- a test / debug call. */
NVB_API int32_t nvb_mapper_append_frame_blocks(NvbMapper* m, int32_t* segment_dev, int32_t cap_entries);
NVB_API int32_t nvb_blocks_union_segments(NvbMapper* m, const int32_t* segments_dev, int32_t num_segments,
                                          int32_t segment_stride_ints, int32_t cap_entries, int32_t* out_xyz_dev,
                                          int32_t out_cap, int32_t* out_count_dev, void* stream);
NVB_API int32_t nvb_blocks_union_status(NvbMapper* m, int32_t* out_error);

/* Device-side join (no host synchronisation): work enqueued on nvb_mapper_stream() after this
 * call also waits for the ESDF wavefront, which runs on an internal side stream so that it
 * overlaps the next frame's TSDF chain. Needed before recording an event on the mapper's stream. */
NVB_API int32_t nvb_mapper_join_streams(NvbMapper* m);
/* The CUDA stream (cudaStream_t) all of the mapper's work is enqueued on
 * (Mapper's shared CudaStream, src/mapper/mapper.cpp:28-46). */
NVB_API void* nvb_mapper_stream(NvbMapper* m);

/* BlockLayer queries (C/include/nvblox/map/layer.h:76-311). All synchronising. */
NVB_API int32_t nvb_layer_num_blocks(NvbMapper* m, int32_t layer, int32_t* out_count);      /* numBlocks        */
NVB_API int32_t nvb_layer_block_indices(NvbMapper* m, int32_t layer, int32_t* out_xyz_host,
                                        int32_t cap, int32_t* out_count);                    /* getAllBlockIndices */
/* getBlockAtIndex(...)->voxels copied to host: out_host receives n blocks of
 * block_bytes (4096 TSDF / 10240 ESDF); found[i] = 0 for unallocated indices
 * (their output bytes are zero). block_bytes: 4096 TSDF / 10240 ESDF / 2048 occupancy. */
NVB_API int32_t nvb_layer_get_blocks(NvbMapper* m, int32_t layer, const int32_t* xyz_host,
                                     int32_t n, void* out_host, uint8_t* found_host);
/* allocateBlockAtIndex + host->device copy of the voxels (tests, map loading). */
NVB_API int32_t nvb_layer_set_blocks(NvbMapper* m, int32_t layer, const int32_t* xyz_host,
                                     int32_t n, const void* in_host);
/* getBlockAtIndex(index).get(): raw device pointer of a block, NULL if absent.
 * Valid until the layer's slab grows or the map is cleared. */
NVB_API int32_t nvb_layer_block_device_ptr(NvbMapper* m, int32_t layer, const int32_t xyz[3],
                                           void** out_ptr);
NVB_API int32_t nvb_layer_block_bytes(int32_t layer);

/* Counters of the last ESDF update (for the roofline's algorithmic bytes):
 * [0] blocks marked, [1] blocks with sites, [2] blocks to clear, [3] clear-pass
 * candidate blocks, [4] blocks cleared, [5] swept blocks, [6] (block,direction)
 * face passes, [7] rings. Synchronising. */
NVB_API int32_t nvb_mapper_last_esdf_stats(NvbMapper* m, int64_t out[8]);

/* Time split of the last ESDF wavefront launch as seen by CTA 0 (ns): [0] grid barriers,
 * [1] face-propagation phases, [2] scan + sweep phases, [3] number of grid barriers. Synchronising. */
NVB_API int32_t nvb_mapper_esdf_time_split(NvbMapper* m, int64_t out[4]);

/* Blocks the clear pass of the last ESDF update actually read (clearAllInvalidKernel's candidates,
 * nvblox/src/integrators/esdf_integrator.cu:1587-1647, minus those whose parent box holds no to-clear block; the
 * `clear_candidates` statistic keeps counting the reference's candidates). */
NVB_API int32_t nvb_mapper_esdf_clear_blocks_read(NvbMapper* m, int64_t* out);

/* Debug: work time (ns) of the slowest CTA in each barrier-delimited phase of the last wavefront. */
NVB_API int32_t nvb_mapper_debug_phase_max(NvbMapper* m, int64_t* out, int32_t cap);

/* Per-stage device time of the frames since the last reset, measured with CUDA
 * events on the mapper's stream when profiling is enabled (same names as the
 * reference's timers: "tsdf/integrate", "esdf/integrate", ...;
 * projective_integrator_impl.cuh:230-270, esdf_integrator.cu:224-252).
 * stage ids: 0 view raycast, 1 block compaction+allocation, 2 tsdf update,
 * 3 esdf allocate+mark, 4 esdf clear, 5 esdf compute. out_ms / out_calls have 6 entries. */
NVB_API int32_t nvb_mapper_enable_profiling(NvbMapper* m, int32_t enable);
NVB_API int32_t nvb_mapper_stage_times(NvbMapper* m, double* out_ms, int64_t* out_calls,
                                       int32_t reset);
/* Number of kernels this library launched on behalf of the mapper since creation. */
NVB_API int64_t nvb_mapper_kernel_launches(const NvbMapper* m);

#ifdef __cplusplus
}
#endif
#endif /* NVBLOX_B200_H_ */
