/*
 * nvblox_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the nvblox_core depth-integration hot path
 * (view raycast -> projective TSDF integration -> incremental ESDF), written
 * from the reference's behaviour (file:line citations are on every function in
 * nvblox_oracle.c; paths are relative to
 * /root/reference/nvblox_ros/nvblox_core/nvblox/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library. The product path
 * (isaac_ros_nvblox_b200/) never does.
 *
 * Parity status: PARITY UNPINNED against a reference binary. The reference cannot be built here (Eigen/stdgpu/glog/... are
 * un-vendored network fetches and there is no GPU), so this oracle is pinned
 * against the reference's own known-answer tests (tests/test_oracle_kat.py,
 * test_oracle_color_kat.py, test_oracle_esdf_scenes_kat.py) and NOT against a
 * reference binary.
 *
 * Arithmetic: IEEE binary32, one rounding per operation, no FMA contraction
 * (compile with -ffp-contract=off). 3-term sums use Eigen's unrolled
 * reduction order a0 + (a1 + a2). float->int casts follow the CUDA device
 * semantics (NaN -> 0, saturating), because the reference runs these casts on
 * the GPU.
 */
#ifndef NVBLOX_ORACLE_H_
#define NVBLOX_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Pinhole camera with optional radial-tangential distortion (include/nvblox/sensors/camera.h:193-203). */
typedef struct {
  float fu, fv, cu, cv;
  int32_t width, height;
  /* std::optional<RadialTangentialDistortionParams> (sensors/distortion.h:24-62, camera.h:202) */
  int32_t has_distortion;
  float k1, k2, k3, k4, k5, k6; /* radial numerator k1..k3, denominator k4..k6 */
  float p1, p2;                 /* tangential */
} OrCamera;

/* Weighting modes (include/nvblox/integrators/weighting_function.h:11-18). */
enum {
  OR_WEIGHT_CONSTANT = 0,
  OR_WEIGHT_CONSTANT_DROPOFF = 1,
  OR_WEIGHT_INVERSE_SQUARE = 2,
  OR_WEIGHT_INVERSE_SQUARE_DROPOFF = 3,
  OR_WEIGHT_INVERSE_SQUARE_TSDF_DISTANCE_PENALTY = 4,
  OR_WEIGHT_LINEAR_WITH_MAX = 5
};

/* Workspace bounds (include/nvblox/geometry/workspace_bounds.h:24). */
enum { OR_WS_UNBOUNDED = 0, OR_WS_HEIGHT_BOUNDS = 1, OR_WS_BOUNDING_BOX = 2 };

/* Mask modes (include/nvblox/sensors/image.h:383). */
enum { OR_MASK_NON_INVERTED = 0, OR_MASK_INVERTED = 1 };

typedef struct {
  float truncation_distance_vox;    /* default 4   */
  float max_integration_distance_m; /* default 7   */
  float max_weight;                 /* default 5   */
  float invalid_depth_decay_factor; /* default -1  */
  int32_t weighting_type;           /* default OR_WEIGHT_INVERSE_SQUARE */
  int32_t raycast_subsampling;      /* default 4   */
  int32_t workspace_bounds_type;    /* default OR_WS_UNBOUNDED */
  float workspace_min[3];
  float workspace_max[3];
} OrTsdfParams;

typedef struct {
  float max_esdf_distance_m;   /* default 2    */
  float max_site_distance_vox; /* default 1    */
  float min_weight;            /* default 1e-4 */
  float occupied_threshold;    /* default 0.5 (probability; OccupancyLayer input only) */
} OrEsdfParams;

/* ProjectiveOccupancyIntegrator's sensor model (integrators/occupancy_integrator_params.h:21-40). */
typedef struct {
  float free_region_occupancy_probability;       /* 0.3 */
  float occupied_region_occupancy_probability;   /* 0.7 */
  float unobserved_region_occupancy_probability; /* 0.5 */
  float occupied_region_half_width_m;            /* 0.1 */
} OrOccupancyParams;

/* Voxel layouts = the reference's (include/nvblox/map/voxels.h:28-74). */
typedef struct {
  float distance;
  float weight;
} OrTsdfVoxel; /* 8 B */

typedef struct {
  float squared_distance_vox;
  int32_t parent_direction[3];
  uint8_t is_inside, observed, is_site, pad_;
} OrEsdfVoxel; /* 20 B */

/* FreespaceVoxel (map/voxels.h:38-52): two Time (int64 ms) fields and a bool. */
typedef struct {
  int64_t last_occupied_timestamp_ms;
  int64_t consecutive_occupancy_duration_ms;
  uint8_t is_high_confidence_freespace, pad_[7];
} OrFreespaceVoxel; /* 24 B */

/* FreespaceIntegrator parameters (integrators/freespace_integrator_params.h:22-58). */
typedef struct {
  float max_tsdf_distance_for_occupancy_m;                   /* 0.15 */
  int64_t max_unobserved_to_keep_consecutive_occupancy_ms;   /* 200  */
  int64_t min_duration_since_occupied_for_freespace_ms;      /* 1000 */
  int64_t min_consecutive_occupancy_duration_for_reset_ms;   /* 2000 */
  int32_t check_neighborhood;                                /* 1    */
  int32_t initialize_to_high_confidence_freespace;           /* 0    */
} OrFreespaceParams;

typedef struct OrMap OrMap;

void or_default_tsdf_params(OrTsdfParams* p);
void or_default_esdf_params(OrEsdfParams* p);

OrMap* or_map_create(float voxel_size_m);
void or_map_destroy(OrMap* map);
void or_map_clear(OrMap* map);

/* RayCaster (rays/internal/impl/ray_caster_impl.h:26-90). Returns the number
 * of cells, writing up to cap triples to out_xyz. */
int32_t or_raycast_cells(const float origin[3], const float dest[3], float scale,
                         int32_t* out_xyz, int32_t cap);

/* ViewCalculator::getBlocksInImageViewRaycast (view_calculator_impl.cuh:117-198).
 * T_L_C: 4x4 column-major (Eigen Isometry3f::data()). Returns count (x-fastest
 * order inside the view AABB), writes up to cap triples. */
int32_t or_view_raycast(const float* depth, int32_t rows, int32_t cols,
                        const float* T_L_C, const OrCamera* cam, float block_size,
                        float truncation_distance_m, const OrTsdfParams* params,
                        int32_t* out_xyz, int32_t cap);

/* ProjectiveTsdfIntegrator::integrateFrame. mask may be NULL. Returns the
 * number of updated blocks (== raycast blocks), writes up to cap triples. */
int32_t or_tsdf_integrate(OrMap* map, const float* depth, const uint8_t* mask,
                          int32_t mask_mode, int32_t rows, int32_t cols,
                          const float* T_L_C, const OrCamera* cam,
                          const OrTsdfParams* params, int32_t* out_xyz,
                          int32_t cap);

/* Same update but on a caller-given block list (bench split timing). */
void or_tsdf_integrate_blocks(OrMap* map, const float* depth, const uint8_t* mask,
                              int32_t mask_mode, int32_t rows, int32_t cols,
                              const float* T_L_C, const OrCamera* cam,
                              const OrTsdfParams* params, const int32_t* blocks_xyz,
                              int32_t num_blocks);

/* ProjectiveOccupancyIntegrator::integrateFrame (integrators/projective_occupancy_integrator.h:51-55). `params`
 * carries the shared ProjectiveIntegrator settings (truncation, max distance, raycast, workspace); its
 * truncation_distance_vox is raised in place when smaller than the occupied half width
 * (src/integrators/projective_occupancy_integrator.cu:42-65). */
void or_default_occupancy_params(OrOccupancyParams* p);
int32_t or_occupancy_integrate(OrMap* map, const float* depth, const uint8_t* mask, int32_t mask_mode,
                               int32_t rows, int32_t cols, const float* T_L_C, const OrCamera* cam,
                               OrTsdfParams* params, const OrOccupancyParams* occ, int32_t* out_xyz, int32_t cap);
/* EsdfIntegrator::integrateBlocks(OccupancyLayer, blocks, EsdfLayer*) (integrators/esdf_integrator.h:72-80). */
void or_esdf_integrate_occupancy(OrMap* map, const int32_t* blocks_xyz, int32_t num_blocks,
                                 const OrEsdfParams* params);
int32_t or_occupancy_num_blocks(const OrMap* map);
int32_t or_occupancy_block_indices(const OrMap* map, int32_t* out_xyz, int32_t cap);
int32_t or_occupancy_get_block(const OrMap* map, const int32_t xyz[3], float* out_log_odds);

/* Test helper: overwrite / create an occupancy block (512 log-odds, [x][y][z] order). */
void or_occupancy_set_block(OrMap* map, const int32_t xyz[3], const float* in_log_odds);

/* TsdfDecayIntegrator / OccupancyDecayIntegrator (integrators/tsdf_decay_integrator_params.h:21-48,
 * occupancy_decay_integrator_params.h:21-43, internal/decay_integrator_base_params.h:22-29). */
typedef struct {
  float decay_factor;                   /* 0.95 */
  float decayed_weight_threshold;       /* 1e-3 */
  int32_t set_free_distance_on_decayed; /* 0    */
  float free_distance_vox;              /* 4    */
  int32_t deallocate_decayed_blocks;    /* 1    */
} OrTsdfDecayParams;
typedef struct {
  float free_region_decay_probability;     /* 0.55 */
  float occupied_region_decay_probability; /* 0.4  */
  float decay_to_probability;              /* 0.5 (Mapper's occupancy_decay_to_free: 0.49, occupancy_decay_integrator.h:35-36) */
  int32_t deallocate_decayed_blocks;       /* 1    */
} OrOccupancyDecayParams;
/* DecayBlockExclusionOptions (integrators/internal/decayer.h:31-44). */
typedef struct {
  const int32_t* excluded_blocks_xyz;
  int32_t num_excluded_blocks;
  int32_t has_exclusion_sphere;
  float exclusion_center[3];
  float exclusion_radius_m;
} OrDecayExclusion;
void or_default_tsdf_decay_params(OrTsdfDecayParams* p);
void or_default_occupancy_decay_params(OrOccupancyDecayParams* p);
/* VoxelDecayer::decay (integrators/internal/cuda/impl/decayer_impl.cuh:150-262). depth == NULL: every voxel decays;
 * otherwise voxels with a depth measurement in that view are spared (DepthObservationSpace). clear_esdf: 1 = also remove
 * the deallocated blocks from the (3-D) ESDF and the freespace layer, like Mapper::decayTsdfInternal does; 2 = the ESDF is
 * a 2-D slice (heights of the last or_esdf_integrate_slice call): a slice block goes when its column is empty. Returns the number of deallocated
 * blocks, writes up to cap triples. `exclusion` may be NULL. */
int32_t or_tsdf_decay(OrMap* map, const OrTsdfDecayParams* params, const OrDecayExclusion* exclusion, const float* depth,
                      int32_t rows, int32_t cols, const float* T_L_C, const OrCamera* cam, float max_view_distance_m,
                      float truncation_distance_m, int32_t clear_esdf, int32_t* out_xyz, int32_t cap);
int32_t or_occupancy_decay(OrMap* map, const OrOccupancyDecayParams* params, const OrDecayExclusion* exclusion,
                           const float* depth, int32_t rows, int32_t cols, const float* T_L_C, const OrCamera* cam,
                           float max_view_distance_m, float truncation_distance_m, int32_t clear_esdf, int32_t* out_xyz,
                           int32_t cap);

/* FreespaceIntegrator::updateFreespaceLayer (integrators/freespace_integrator.h:60-66): dynablox freespace update of
 * the listed blocks at `update_time_ms`. depth == NULL: no viewpoint exclusion; otherwise only voxels with a depth
 * measurement in that view are updated. The time of the last call is kept in the map. */
void or_default_freespace_params(OrFreespaceParams* p);
void or_freespace_update(OrMap* map, const int32_t* blocks_xyz, int32_t num_blocks, int64_t update_time_ms,
                         const OrFreespaceParams* params, const float* depth, int32_t rows, int32_t cols,
                         const float* T_L_C, const OrCamera* cam, float max_view_distance_m, float truncation_distance_m);
int32_t or_freespace_num_blocks(const OrMap* map);
int32_t or_freespace_block_indices(const OrMap* map, int32_t* out_xyz, int32_t cap);
int32_t or_freespace_get_block(const OrMap* map, const int32_t xyz[3], OrFreespaceVoxel* out);
/* EsdfIntegrator::integrateBlocks(TsdfLayer, FreespaceLayer, blocks, EsdfLayer*) (esdf_integrator.h:64-70). */
void or_esdf_integrate_with_freespace(OrMap* map, const int32_t* blocks_xyz, int32_t num_blocks,
                                      const OrEsdfParams* params);

/* EsdfIntegrator::integrateBlocks(TsdfLayer, blocks, EsdfLayer*). */
void or_esdf_integrate(OrMap* map, const int32_t* blocks_xyz, int32_t num_blocks,
                       const OrEsdfParams* params);

/* EsdfIntegrator::integrateSlice(layer [, freespace], blocks, esdf) with a constant-z slice
 * (integrators/esdf_integrator.h:96-150; esdf_slice_min_height / max_height / slice_height,
 * esdf_integrator_params.h:33-43): 2-D ESDF on one layer of blocks at z_output_m. */
void or_esdf_integrate_slice(OrMap* map, int32_t from_occupancy, int32_t use_freespace, const int32_t* blocks_xyz,
                             int32_t num_blocks, const OrEsdfParams* params, float z_min_m, float z_max_m,
                             float z_output_m);

/* integrateSlice with a PlanarSliceDescription: plane = (nx, ny, nz, d), unit normal, n . p + d = 0
 * (slice_height_above_plane_m 0, slice_height_thickness_m 0.1 by default, esdf_integrator_params.h:45-52). */
/* ColorVoxel (map/voxels.h:77-83): Color (3 bytes) + 1 byte padding + float weight. */
typedef struct {
  uint8_t r, g, b, pad;
  float weight;
} OrColorVoxel;
/* ProjectiveColorIntegrator + its SphereTracer + its ViewCalculator (projective_appearance_integrator.h:150-175) */
typedef struct {
  float max_integration_distance_m;
  float truncation_distance_vox;
  float max_weight;
  float measurement_weight;
  int32_t sphere_tracing_ray_subsampling_factor;
  int32_t sphere_tracer_maximum_steps;
  float sphere_tracer_maximum_ray_length_m;
  float sphere_tracer_surface_distance_epsilon_vox;
  int32_t workspace_bounds_type;
  float workspace_min[3], workspace_max[3];
} OrColorParams;
int32_t or_mark_unobserved_free_inside_radius(OrMap* map, int32_t occupancy, const float center[3], float radius,
                                              float truncation_distance_m, int32_t* out_xyz, int32_t cap);
int32_t or_view_projection_blocks(float block_size, const float* T_L_C, const OrCamera* cam, float max_distance,
                                  int32_t* out_xyz, int32_t cap);
void or_freespace_set_block(OrMap* m, const int32_t xyz[3], const OrFreespaceVoxel* in);
float or_weighting(int32_t type, float measured, float voxel_depth, float trunc);
void or_default_color_params(OrColorParams* p);
float or_round_through_half(float f);
int32_t or_sphere_trace_ray(const OrMap* map, const float origin[3], const float direction[3], float truncation_distance_m,
                            int32_t maximum_steps, float maximum_ray_length_m, float surface_distance_epsilon_m, float* t_out);
void or_sphere_trace_image(const OrMap* map, const float* T_L_C, const OrCamera* cam, float truncation_distance_m,
                           int32_t maximum_steps, float maximum_ray_length_m, float surface_distance_epsilon_m,
                           int32_t ray_subsampling_factor, float* out);
int32_t or_color_integrate(OrMap* map, const uint8_t* color, const uint8_t* mask, int32_t mask_mode, int32_t rows, int32_t cols,
                           const float* T_L_C, const OrCamera* cam, const OrColorParams* P, int32_t* out_xyz, int32_t cap);
int32_t or_color_num_blocks(const OrMap* m);
int32_t or_color_block_indices(const OrMap* m, int32_t* out, int32_t cap);
int32_t or_color_get_block(const OrMap* m, const int32_t xyz[3], OrColorVoxel* out);
void or_planar_column_bounds(float block_size, const float plane[4], float above_plane_m, float thickness_m, int32_t bx, int32_t by,
                             int32_t vx, int32_t vy, int32_t out[4]);
int32_t or_planar_num_blocks_in_column(float block_size, float thickness_m);
void or_block_and_voxel_from_1d(float block_size, float p, int32_t out[2]);
void or_esdf_integrate_slice_planar(OrMap* map, int32_t from_occupancy, int32_t use_freespace, const int32_t* blocks_xyz,
                                    int32_t num_blocks, const OrEsdfParams* params, const float plane[4], float above_plane_m,
                                    float thickness_m, float z_output_m);

/* EsdfSlicer::sliceLayerToDistanceImage + occupancyGridFromSliceImage (integrators/esdf_slicer.h:52-118): the distance map
 * (m, negative inside, `unobserved_value` where nothing is known) of the ESDF layer at `slice_height`, one pixel per voxel
 * over the AABB of the blocks at that height (aabb_out = min xyz, max xyz); rows follow y, columns x. */
int32_t or_esdf_slice_image(const OrMap* map, float slice_height, float unobserved_value, float aabb_out[6],
                            float* image_out, int8_t* grid_out, int32_t cap, int32_t* rows_out, int32_t* cols_out);
/* EsdfSlicer::getAabbOfLayerAtHeight (:112-147); 0 = no block at that height */
int32_t or_esdf_slice_aabb(const OrMap* map, float slice_height, float aabb_out[6]);
/* EsdfSlicer::sliceLayerToDistanceImage on a given AABB (:169-199): what sliceLayersToCombinedDistanceImage (:201-240) runs on
 * each of its two layers with their merged box before the element-wise minimum */
int32_t or_esdf_slice_image_in_aabb(const OrMap* map, float slice_height, float unobserved_value, const float aabb[6],
                                    float* image_out, int8_t* grid_out, int32_t cap, int32_t* rows_out, int32_t* cols_out);

/* Statistics of the last or_esdf_integrate call: [0] blocks marked, [1] blocks
 * with sites, [2] blocks to clear, [3] candidate blocks scanned by the clear
 * pass, [4] blocks cleared, [5] total swept blocks, [6] total (block,direction)
 * face passes, [7] rings. */
void or_esdf_last_stats(const OrMap* map, int64_t out[8]);

/* Read-back. */
int32_t or_tsdf_num_blocks(const OrMap* map);
int32_t or_esdf_num_blocks(const OrMap* map);
int32_t or_tsdf_block_indices(const OrMap* map, int32_t* out_xyz, int32_t cap);
int32_t or_esdf_block_indices(const OrMap* map, int32_t* out_xyz, int32_t cap);
/* Copies the 512 voxels of a block ([x][y][z] order). Returns 0 if absent. */
int32_t or_tsdf_get_block(const OrMap* map, const int32_t xyz[3], OrTsdfVoxel* out);
int32_t or_esdf_get_block(const OrMap* map, const int32_t xyz[3], OrEsdfVoxel* out);
/* Test helper: overwrite / create a TSDF block. */
/* ViewCalculator::cache_last_viewpoint (integrators/view_calculator.h:196; default on) of the map's projective integrator */
void or_map_cache_last_viewpoint(OrMap* map, int32_t enable);
/* DepthPreprocessor::dilateInvalidRegionsAsync (src/sensors/depth_preprocessing.cpp:36-58; the class defaults are
 * threshold 1e-2, value 0): out = depth with the (depth < threshold) regions dilated num_dilations times by 3x3. */
void or_depth_dilate_invalid(const float* depth, int32_t rows, int32_t cols, int32_t num_dilations, float threshold,
                             float value, float* out);
void or_tsdf_set_block(OrMap* map, const int32_t xyz[3], const OrTsdfVoxel* in);
void or_esdf_set_block(OrMap* map, const int32_t xyz[3], const OrEsdfVoxel* in); /* test hook */

/* ---- Mesh integrator (mesh/mesh_integrator.h; src/mesh/mesh_integrator.cu, mesh_integrator_appearance.cu) */
typedef struct {
  float min_weight;          /* mesh_integrator_params.h:22-24, default 1e-4 */
  int32_t weld_vertices;     /* mesh_integrator_params.h:25-27, default true */
  float cutoff_distance_vox; /* MeshIntegrator::cutoff_distance_vox_ (mesh_integrator.h:129), 5 */
} OrMeshParams;
void or_default_mesh_params(OrMeshParams* p);
/* MeshIntegrator::integrateBlocksGPU (mesh_integrator.cu:66-108): re-meshes the listed blocks (those in the TSDF layer).
 * A block's triangles come out in x-major voxel order (the reference's order is an atomicAdd race). */
void or_mesh_integrate_blocks(OrMap* map, const int32_t* blocks_xyz, int32_t num_blocks, const OrMeshParams* params);
/* MeshIntegrator::updateAppearanceGPU for the colour layer (mesh_integrator_appearance.cu:281-380) */
void or_mesh_update_color(OrMap* map, const int32_t* blocks_xyz, int32_t num_blocks);
int32_t or_mesh_num_blocks(const OrMap* map);
int32_t or_mesh_block_indices(const OrMap* map, int32_t* out_xyz, int32_t cap);
/* out = {vertices, triangle indices, colours}; 0 if the block is absent */
int32_t or_mesh_block_sizes(const OrMap* map, const int32_t xyz[3], int32_t out[3]);
/* vertices / normals: 3 floats each; triangles: indices into vertices; colours: rgba bytes. Any pointer may be NULL. */
int32_t or_mesh_get_block(const OrMap* map, const int32_t xyz[3], float* vertices, float* normals, int32_t* triangles,
                          uint8_t* colors);

/* Camera::project / vectorFromImagePlaneCoordinates exposed for the distortion tests. */
int32_t or_camera_project(const OrCamera* cam, const float p_C[3], float uv[2]);
void or_camera_vector_from_image_plane(const OrCamera* cam, float u, float v, float out[3]);

/* Number of OpenMP threads the oracle will use (1 if built without OpenMP). */
int32_t or_num_threads(void);
void or_set_num_threads(int32_t n);

#ifdef __cplusplus
}
#endif
#endif /* NVBLOX_ORACLE_H_ */
