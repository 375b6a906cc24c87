// nvb_api.cu -- the C-ABI (include/nvblox_b200.h) and the host-side orchestration.
//
// Host responsibilities are reduced to what the reference also does on the host
// for this path and cannot be avoided: the view AABB from the camera pose
// (Camera::getViewAABB, nvblox/src/sensors/camera.cpp:31-83; ViewCalculator setup,
// view_calculator_impl.cuh:137-156), T_C_L = T_L_C^-1
// (projective_integrator_impl.cuh:268), and capacity bookkeeping. Block lists,
// allocation, the update tracker and the ESDF wavefront stay on the device.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "nvb_internal.cuh"

using namespace nvb;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

#define NVB_CUDA(expr)                                                                               \
  do {                                                                                               \
    cudaError_t e_ = (expr);                                                                         \
    if (e_ != cudaSuccess)                                                                           \
      return fail(NVB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));                 \
  } while (0)

constexpr int kDefaultCapacity = 1 << 18;
constexpr int kStagingBuffers = 3;
constexpr int kCountRing = 8;
constexpr int kNumStages = 6;

struct StageEvent {
  cudaEvent_t start, stop;
  int stage;
};

}  // namespace

struct NvbMapper {
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;
  // The ESDF wavefront only touches the ESDF layer: it runs on its own stream so that the next
  // frame's raycast / compaction / TSDF update overlaps it.
  cudaStream_t esdf_stream = nullptr;
  cudaEvent_t esdf_ready = nullptr;  // TSDF chain of the update done on `stream`
  cudaEvent_t mark_done = nullptr;   // allocate + mark done on `esdf_stream`
  cudaEvent_t esdf_done = nullptr;   // wavefront done on `esdf_stream`
  bool esdf_in_flight = false;
  float voxel_size = 0.05f, block_size = 0.4f;
  NvbTsdfParams tp;
  NvbEsdfParams ep;
  NvbOccupancyParams op;
  // NVB_PROJECTIVE_TSDF or NVB_PROJECTIVE_OCCUPANCY: which voxel type the projective layer (`tsdf` below) holds
  // (ProjectiveLayerType, mapper/mapper.h:52-53).
  int projective_layer_type = 0;
  int esdf_persistent = 1;
  int esdf_reserved_sms = 2;  // SMs the exchange-slab wavefront leaves to concurrently running kernels (nvb_esdf_wavex.cu)

  DevLayer tsdf{}, esdf{};
  DevLayer freespace{};    // FreespaceLayer of a NVB_PROJECTIVE_TSDF_WITH_FREESPACE mapper
  DevLayer color{};        // ColorLayer, created by the first nvb_mapper_integrate_color
  NvbColorParams cp{};
  int4* color_work = nullptr;  // blocks of the last colour frame {x, y, z, colour slot}
  int color_work_cap = 0;
  float* color_synth = nullptr;  // sphere-traced synthetic depth
  size_t color_synth_cap = 0;
  unsigned char* color_stage = nullptr;  // host colour image / mask staged on the device
  size_t color_stage_cap = 0;
  unsigned char* color_mask_stage = nullptr;
  size_t color_mask_stage_cap = 0;
  NvbFreespaceParams fp;
  NvbEsdfSliceParams sp;
  int esdf_mode = 0;                // EsdfMode: 0 unset, 1 3-D, 2 2-D slice (mapper.h:61, src/mapper/mapper.cpp:408-470)
  unsigned long long* colset = nullptr;
  size_t colset_n = 0;
  int* cols = nullptr;
  int cols_cap = 0;
  long long fs_last_update_ms = 0;  // FreespaceIntegrator::last_update_time_ms_ (freespace_integrator.h:171)
  int* dirty_fs = nullptr;          // the tracker's second consumer (BlocksToUpdateType::kFreespace)
  int* todo_fs_slots = nullptr;
  bool fs_tracker_initialized = false;
  int4* fs_work = nullptr;
  int fs_work_cap = 0;
  int tsdf_count_ub = 0;   // host-side upper bound of *tsdf.count
  int esdf_extra_ub = 0;   // blocks submitted to the ESDF through explicit lists

  // ViewpointCache of the projective integrator's ViewCalculator (C/include/nvblox/integrators/view_calculator.h:196,211-244):
  // up to two (pose, sensor) -> block-list entries, newest first. A list is kept as the view bitset it was compacted from (a
  // few KB on the device): a hit skips the raycast and replays the compaction + allocation, which yields the same list in the
  // same order and re-allocates blocks that were deallocated in between, like allocateBlocksWhereRequired does in the reference.
  BlockTensorMap tsdf_tmap;  // TMA descriptor of the TSDF slab (nvb_tsdf.cu)
  int4* union_list = nullptr;  // nvb_blocks_union's own output list
  int* union_list_count = nullptr;
  int union_list_cap = 0;

  // Mesh layer (nvb_mesh.cu): header slab + one arena for vertices / normals / triangle indices / colours
  DevLayer mesh{};
  float* mesh_v = nullptr;
  float* mesh_n = nullptr;
  int* mesh_t = nullptr;
  unsigned char* mesh_c = nullptr;
  long long mesh_arena_cap = 0;  // entries
  float* mesh_alt_v = nullptr;   // the spare arena a repack moves the live segments into (then the two swap)
  float* mesh_alt_n = nullptr;
  int* mesh_alt_t = nullptr;
  unsigned char* mesh_alt_c = nullptr;
  long long mesh_alt_cap = 0;
  int* mesh_state = nullptr;     // kArena* ints
  int* mesh_counts = nullptr;
  int* mesh_offsets = nullptr;
  int mesh_list_cap = 0;
  int* mesh_xyz_dev = nullptr;
  int mesh_xyz_cap = 0;
  int* dirty_mesh = nullptr;  // the tracker's third consumer (BlocksToUpdateType::kColorMesh)
  int* todo_mesh_slots = nullptr;
  bool mesh_tracker_initialized = false;
  NvbMeshParams mp{1e-4f, 1, 5.0f};
  int cache_last_viewpoint = 1;
  // Mapper::do_depth_preprocessing / depth_preprocessing_num_dilations (mapper_params.h:33-42; mapper.cpp:335-352)
  int do_depth_preprocessing = 0;
  int depth_preprocessing_num_dilations = 4;
  float* pre_depth = nullptr;
  size_t pre_depth_cap = 0;
  int vc_n = 0;
  float vc_T[2][16];
  NvbCamera vc_cam[2];
  ViewGrid vc_grid[2];
  long long vc_cells[2] = {0, 0};
  unsigned int* vc_bits[2] = {nullptr, nullptr};
  size_t vc_bits_cap[2] = {0, 0};

  // per-frame scratch
  unsigned int* bits = nullptr;
  size_t bits_words_cap = 0;
  int4* frame_blocks = nullptr;
  int frame_cap = 0;
  int* frame_count = nullptr;
  unsigned long long* tile_state = nullptr;
  int tile_cap = 0;
  unsigned int* ticket = nullptr;
  unsigned int ticket_base = 0;
  unsigned int epoch = 0;
  int* error_dev = nullptr;

  // depth / mask staging for host inputs
  float* depth_stage[kStagingBuffers] = {nullptr, nullptr, nullptr};
  unsigned char* mask_stage[kStagingBuffers] = {nullptr, nullptr, nullptr};
  size_t stage_pixels = 0;
  cudaEvent_t stage_copied[kStagingBuffers];
  cudaEvent_t stage_consumed[kStagingBuffers];
  bool stage_used[kStagingBuffers] = {false, false, false};
  unsigned long long frame_seq = 0;

  // tracker
  int* dirty = nullptr;
  int* todo_slots = nullptr;
  int* todo_count = nullptr;
  bool tracker_initialized = false;

  // esdf scratch
  int4* work = nullptr;
  int* esdf_ints = nullptr;  // small counters block
  int* upd_list = nullptr;
  int* clr_list = nullptr;
  int* clr_cand = nullptr;  // candidates of the clear pass that survive the pruning (select kernel -> process kernel)
  int* cleared_list = nullptr;
  int* ring_a = nullptr;
  int* ring_b = nullptr;
  int* stamp_a = nullptr;
  int* stamp_b = nullptr;
  int* nbr = nullptr;
  int* nbr27 = nullptr;
  unsigned char* shadow = nullptr;
  int shadow_cap = 0;
  unsigned char* xslab = nullptr;  // exchange-slab wavefront (esdf_persistent == 3): 2 x capacity ESDF blocks
  int* xrec = nullptr;             // ... 2 x CTAs x xseg candidate records of 32 ints ...
  int xseg = 0;
  int* xcounts = nullptr;          // ... and 2 x CTAs count pairs
  int* cand_stamp = nullptr;
  // decay integrators
  NvbTsdfDecayParams tdp;
  NvbOccupancyDecayParams odp;
  int4* dead = nullptr;          // deallocated projective blocks of the last decay call
  int dead_cap = 0;
  int* skip_stamp = nullptr;     // per projective slot: == skip_seq -> block excluded from this decay call
  int skip_cap = 0;
  int skip_seq = 0;
  int* dead_cleared_xyz = nullptr;
  int dead_cleared_cap = 0;
  // last integrated view (Mapper::last_posed_depth_image_, mapper.h:830-833), kept when keep_last_view is set
  int keep_last_view = 0;
  float* last_depth = nullptr;
  size_t last_depth_cap = 0;
  int last_rows = 0, last_cols = 0;
  float last_T_L_C[16];
  NvbCamera last_cam;
  bool has_last_view = false;
  int* cand_a = nullptr;
  int* cand_b = nullptr;
  int ges_switch = 160;
  int* seed_upd = nullptr;
  int* seed_clr = nullptr;
  // device-resident merge of block lists (nvb_blocks_union_segments): own scratch, usable on any stream
  unsigned int* union_bits = nullptr;
  long long union_bits_cap = 0;      // in bits
  int* union_state = nullptr;        // AABB, error flag, words in use
  unsigned int* clr_bits = nullptr;  // to-clear bitmap of the current update (2048 words)
  unsigned int* psum = nullptr;  // per ESDF slot: box of the block offsets its voxels' parents point into (clear-pass pruning)
  bool prune_default = false;    // esdf_persistent == 3 and not switched off (NVB_CLEAR_PRUNE=0)
  bool prune_ok = false;         // the summaries are upper bounds for every block (only the exchange-slab wavefront keeps them)
  int update_seq = 0;
  long long* stats = nullptr;
  unsigned int* barrier = nullptr;
  unsigned long long* phase_max = nullptr;
  int* xyz_upload = nullptr;
  int xyz_upload_cap = 0;

  // pinned host scratch
  int* h_ints = nullptr;  // [0] frame count, [1] error, [2..] misc, [8], [9] prefetched error words (pinned)
  int4* h_list = nullptr;  // pinned: the last frame's block list lands here (readFrameList)
  int last_frame_n = 0;
  int* h_count_ring = nullptr;
  cudaEvent_t count_events[kCountRing];
  bool count_pending[kCountRing];
  long long count_cum_at[kCountRing];  // cells_cum when the read-back was enqueued
  int count_ring_head = 0;
  int tsdf_count_confirmed = 0;        // last *tsdf.count seen by the host ...
  long long confirmed_cum = 0;         // ... and cells_cum at that moment
  long long cells_cum = 0;             // sum of the view-AABB cells of every frame enqueued so far

  long long launches = 0;
  bool profiling = false;
  std::vector<StageEvent> stage_events;
  double stage_ms[kNumStages] = {0, 0, 0, 0, 0, 0};
  long long stage_calls[kNumStages] = {0, 0, 0, 0, 0, 0};
};

namespace {

cudaError_t syncAll(NvbMapper* m) {
  cudaError_t e = cudaStreamSynchronize(m->stream);
  if (e != cudaSuccess) return e;
  if (m->esdf_stream) e = cudaStreamSynchronize(m->esdf_stream);
  m->esdf_in_flight = false;
  return e;
}

// Device-side join: later work on `stream` waits for the wavefront on `esdf_stream`.
cudaError_t joinEsdf(NvbMapper* m) {
  if (!m->esdf_in_flight) return cudaSuccess;
  m->esdf_in_flight = false;
  return cudaStreamWaitEvent(m->stream, m->esdf_done, 0);
}

int nextPow2(long long v) {
  long long p = 1;
  while (p < v) p <<= 1;
  return (int)p;
}

int allocLayer(DevLayer* L, int capacity, int block_bytes, cudaStream_t stream) {
  L->capacity = capacity;
  L->block_bytes = block_bytes;
  NVB_CUDA(cudaMalloc(&L->blocks, (size_t)capacity * block_bytes));
  NVB_CUDA(cudaMemsetAsync(L->blocks, 0, (size_t)capacity * block_bytes, stream));
  NVB_CUDA(cudaMalloc(&L->block_index, (size_t)capacity * 3 * sizeof(int)));
  NVB_CUDA(cudaMalloc(&L->count, sizeof(int)));
  NVB_CUDA(cudaMemsetAsync(L->count, 0, sizeof(int), stream));
  NVB_CUDA(cudaMalloc(&L->free_slots, (size_t)capacity * sizeof(int)));
  NVB_CUDA(cudaMalloc(&L->free_count, sizeof(int)));
  NVB_CUDA(cudaMemsetAsync(L->free_count, 0, sizeof(int), stream));
  const int hcap = nextPow2(2ll * capacity);
  L->hash.mask = (unsigned int)hcap - 1;
  NVB_CUDA(cudaMalloc(&L->hash.keys, (size_t)hcap * sizeof(unsigned long long)));
  NVB_CUDA(cudaMalloc(&L->hash.vals, (size_t)hcap * sizeof(int)));
  launchFillU64(L->hash.keys, kEmptyKey, (size_t)hcap, stream);
  return NVB_OK;
}

void freeLayer(DevLayer* L) {
  cudaFree(L->blocks), cudaFree(L->block_index), cudaFree(L->count), cudaFree(L->hash.keys), cudaFree(L->hash.vals);
  cudaFree(L->free_slots), cudaFree(L->free_count);
  *L = DevLayer{};
}

template <typename T>
int reallocCopy(T** p, size_t old_n, size_t new_n, bool zero_rest, cudaStream_t stream) {
  T* q = nullptr;
  NVB_CUDA(cudaMalloc(&q, new_n * sizeof(T)));
  if (zero_rest) NVB_CUDA(cudaMemsetAsync(q, 0, new_n * sizeof(T), stream));
  if (*p && old_n) NVB_CUDA(cudaMemcpyAsync(q, *p, old_n * sizeof(T), cudaMemcpyDeviceToDevice, stream));
  NVB_CUDA(cudaStreamSynchronize(stream));
  if (*p) cudaFree(*p);
  *p = q;
  return NVB_OK;
}

// Doubling growth of a layer slab (BlockMemoryPool expansion,
// map/internal/impl/block_memory_pool_impl.h:54-73). Synchronising and rare.
int growLayer(NvbMapper* m, DevLayer* L, int new_capacity) {
  NVB_CUDA(syncAll(m));
  int count = 0;
  NVB_CUDA(cudaMemcpy(&count, L->count, sizeof(int), cudaMemcpyDeviceToHost));
  count = std::min(count, L->capacity);
  DevLayer N{};
  N.count = L->count;
  N.free_count = L->free_count;
  NVB_CUDA(cudaMalloc(&N.free_slots, (size_t)new_capacity * sizeof(int)));
  NVB_CUDA(cudaMemcpyAsync(N.free_slots, L->free_slots, (size_t)L->capacity * sizeof(int), cudaMemcpyDeviceToDevice,
                           m->stream));
  N.capacity = new_capacity;
  N.block_bytes = L->block_bytes;
  NVB_CUDA(cudaMalloc(&N.blocks, (size_t)new_capacity * L->block_bytes));
  NVB_CUDA(cudaMemsetAsync(N.blocks, 0, (size_t)new_capacity * L->block_bytes, m->stream));
  NVB_CUDA(cudaMemcpyAsync(N.blocks, L->blocks, (size_t)count * L->block_bytes, cudaMemcpyDeviceToDevice, m->stream));
  NVB_CUDA(cudaMalloc(&N.block_index, (size_t)new_capacity * 3 * sizeof(int)));
  NVB_CUDA(cudaMemcpyAsync(N.block_index, L->block_index, (size_t)count * 3 * sizeof(int), cudaMemcpyDeviceToDevice,
                           m->stream));
  const int hcap = nextPow2(2ll * new_capacity);
  N.hash.mask = (unsigned int)hcap - 1;
  NVB_CUDA(cudaMalloc(&N.hash.keys, (size_t)hcap * sizeof(unsigned long long)));
  NVB_CUDA(cudaMalloc(&N.hash.vals, (size_t)hcap * sizeof(int)));
  launchFillU64(N.hash.keys, kEmptyKey, (size_t)hcap, m->stream);
  launchRehash(N, count, m->stream);
  NVB_CUDA(syncAll(m));
  cudaFree(L->blocks), cudaFree(L->block_index), cudaFree(L->hash.keys), cudaFree(L->hash.vals), cudaFree(L->free_slots);
  *L = N;
  return NVB_OK;
}

int allocEsdfScratch(NvbMapper* m, int old_cap, int cap) {
  int rc;
  if ((rc = reallocCopy(&m->work, 0, (size_t)cap, false, m->stream))) return rc;
  if ((rc = reallocCopy(&m->upd_list, 0, (size_t)cap, false, m->stream))) return rc;
  if ((rc = reallocCopy(&m->clr_list, 0, (size_t)cap, false, m->stream))) return rc;
  if ((rc = reallocCopy(&m->clr_cand, 0, (size_t)cap, false, m->stream))) return rc;
  if ((rc = reallocCopy(&m->cleared_list, (size_t)old_cap, (size_t)cap, true, m->stream))) return rc;
  if ((rc = reallocCopy(&m->ring_a, 0, (size_t)cap, false, m->stream))) return rc;
  if ((rc = reallocCopy(&m->ring_b, 0, (size_t)cap, false, m->stream))) return rc;
  if ((rc = reallocCopy(&m->stamp_a, (size_t)old_cap, (size_t)cap, true, m->stream))) return rc;
  if ((rc = reallocCopy(&m->stamp_b, (size_t)old_cap, (size_t)cap, true, m->stream))) return rc;
  if ((rc = reallocCopy(&m->seed_upd, (size_t)old_cap, (size_t)cap, true, m->stream))) return rc;
  if ((rc = reallocCopy(&m->seed_clr, (size_t)old_cap, (size_t)cap, true, m->stream))) return rc;
  if ((rc = reallocCopy(&m->psum, 2 * (size_t)old_cap, 2 * (size_t)cap, true, m->stream))) return rc;
  // neighbour table: 0xFE bytes = "unknown" (< -1) for slots that were never linked
  {
    int* q = nullptr;
    NVB_CUDA(cudaMalloc(&q, (size_t)cap * 6 * sizeof(int)));
    NVB_CUDA(cudaMemsetAsync(q, 0xFE, (size_t)cap * 6 * sizeof(int), m->stream));
    if (m->nbr && old_cap)
      NVB_CUDA(cudaMemcpyAsync(q, m->nbr, (size_t)old_cap * 6 * sizeof(int), cudaMemcpyDeviceToDevice, m->stream));
    NVB_CUDA(syncAll(m));
    if (m->nbr) cudaFree(m->nbr);
    m->nbr = q;
  }
  {
    int* q = nullptr;
    NVB_CUDA(cudaMalloc(&q, (size_t)cap * 27 * sizeof(int)));
    NVB_CUDA(cudaMemsetAsync(q, 0xFE, (size_t)cap * 27 * sizeof(int), m->stream));
    if (m->nbr27 && old_cap)
      NVB_CUDA(cudaMemcpyAsync(q, m->nbr27, (size_t)old_cap * 27 * sizeof(int), cudaMemcpyDeviceToDevice, m->stream));
    NVB_CUDA(syncAll(m));
    if (m->nbr27) cudaFree(m->nbr27);
    m->nbr27 = q;
  }
  if ((rc = reallocCopy(&m->cand_stamp, (size_t)old_cap, (size_t)cap, true, m->stream))) return rc;
  if ((rc = reallocCopy(&m->cand_a, 0, (size_t)cap, false, m->stream))) return rc;
  if ((rc = reallocCopy(&m->cand_b, 0, (size_t)cap, false, m->stream))) return rc;
  if (m->esdf_persistent == 2) {
    // gather-emulate-sweep wavefront: second ESDF slab (contents only live inside one launch)
    if (m->shadow) cudaFree(m->shadow);
    m->shadow = nullptr;
    NVB_CUDA(cudaMalloc(&m->shadow, (size_t)cap * kEsdfBlockBytes));
    m->shadow_cap = cap;
  }
  if (m->esdf_persistent == 3) {
    // exchange-slab wavefront: two slabs by ring parity + the candidate records (contents only live inside one launch)
    NVB_CUDA(syncAll(m));
    if (m->xslab) cudaFree(m->xslab);
    if (m->xrec) cudaFree(m->xrec);
    m->xslab = nullptr, m->xrec = nullptr;
    NVB_CUDA(cudaMalloc(&m->xslab, 2 * (size_t)cap * kEsdfBlockBytes));
    // A CTA registers at most 6 candidates per candidate it owns, i.e. <= 6 * ceil(cap / CTAs) < cap / 16 + 64 per ring.
    const int ctas = std::min(m->num_sms, esdfWaveXMaxCtas());
    m->xseg = cap / 16 + 64;
    NVB_CUDA(cudaMalloc(&m->xrec, 2 * (size_t)ctas * m->xseg * 32 * sizeof(int)));
    if (!m->xcounts) {
      NVB_CUDA(cudaMalloc(&m->xcounts, esdfWaveXFlagBytes()));
      NVB_CUDA(cudaMemsetAsync(m->xcounts, 0, esdfWaveXFlagBytes(), m->stream));
    }
  }
  return NVB_OK;
}

int allocTsdfSide(NvbMapper* m, int old_cap, int cap) {
  int rc;
  if ((rc = reallocCopy(&m->dirty, (size_t)old_cap, (size_t)cap, true, m->stream))) return rc;
  if ((rc = reallocCopy(&m->todo_slots, (size_t)old_cap, (size_t)cap, true, m->stream))) return rc;
  if (m->projective_layer_type == NVB_PROJECTIVE_TSDF_WITH_FREESPACE) {
    if ((rc = reallocCopy(&m->dirty_fs, (size_t)old_cap, (size_t)cap, true, m->stream))) return rc;
    if ((rc = reallocCopy(&m->todo_fs_slots, (size_t)old_cap, (size_t)cap, true, m->stream))) return rc;
  }
  if (m->dirty_mesh) {
    if ((rc = reallocCopy(&m->dirty_mesh, (size_t)old_cap, (size_t)cap, true, m->stream))) return rc;
    if ((rc = reallocCopy(&m->todo_mesh_slots, (size_t)old_cap, (size_t)cap, true, m->stream))) return rc;
  }
  return NVB_OK;
}

constexpr int kHostListCap = 1 << 15;  // entries of the pinned frame-list buffer (512 KiB)

// esdf_ints layout
enum { kWorkCount = 0, kUpdCount = 1, kClrCount = 2, kClrAabb = 3, kClearedCount = 9, kRingCount = 10, kRingId = 14,
       kTodoCount = 15, kFrameCount = 16, kError = 17, kClearedSeq = 18, kTailState = 20, kDeadCount = 22, kDeadClearedCount = 23, kGesCounts = 24, kTodoFsCount = 28, kFsWorkCount = 29, kColsCount = 30, kColorWorkCount = 31, kXTail = 32, kTodoMeshCount = 36, kClrCandCount = 37, kNumInts = 40 };

float logOddsFromProbability(float p);

EsdfCtx makeEsdfCtx(NvbMapper* m) {
  EsdfCtx c{};
  c.tsdf = m->tsdf, c.esdf = m->esdf;
  c.freespace = m->freespace;
  c.use_freespace = m->projective_layer_type == NVB_PROJECTIVE_TSDF_WITH_FREESPACE ? 1 : 0;
  c.work = m->work;
  c.work_count = m->esdf_ints + kWorkCount;
  c.upd_list = m->upd_list, c.upd_count = m->esdf_ints + kUpdCount;
  c.clr_list = m->clr_list, c.clr_count = m->esdf_ints + kClrCount;
  c.clr_cand = m->clr_cand, c.clr_cand_count = m->esdf_ints + kClrCandCount;
  c.clr_aabb = m->esdf_ints + kClrAabb;
  c.cleared_list = m->cleared_list, c.cleared_count = m->esdf_ints + kClearedCount;
  c.ring_a = m->ring_a, c.ring_b = m->ring_b;
  c.ring_count = m->esdf_ints + kRingCount;
  c.tail_state = m->esdf_ints + kTailState;
  c.stamp_a = m->stamp_a, c.stamp_b = m->stamp_b;
  c.ring_id = m->esdf_ints + kRingId;
  c.nbr = m->nbr, c.seed_upd = m->seed_upd, c.seed_clr = m->seed_clr;
  c.clr_bits = m->clr_bits;
  c.psum = m->psum, c.prune = (m->prune_ok && m->esdf_persistent == 3) ? 1 : 0;
  c.nbr27 = m->nbr27, c.shadow = m->shadow, c.cand_stamp = m->cand_stamp;
  c.xslab = m->xslab, c.xrec = m->xrec, c.xtail = m->esdf_ints + kXTail, c.xseg = m->xseg, c.xcounts = m->xcounts;
  c.ges_counts = m->esdf_ints + kGesCounts;
  c.cand_a = m->cand_a, c.cand_b = m->cand_b, c.ges_switch = m->ges_switch;
  c.colset_keys = m->colset, c.colset_mask = m->colset_n ? (unsigned int)(m->colset_n - 1) : 0u;
  c.cols = m->cols, c.cols_count = m->esdf_ints + kColsCount;
  c.dead_cleared_xyz = m->dead_cleared_xyz;
  c.dead_cleared_count = m->dead_cleared_xyz ? m->esdf_ints + kDeadClearedCount : nullptr;
  c.cleared_seq = m->esdf_ints + kClearedSeq;
  c.update_seq = m->update_seq;
  c.barrier = m->barrier;
  c.phase_max = m->phase_max;
  c.stats = m->stats;
  c.error = m->error_dev;
  // esdf_integrator.cu:693-696, 672-676
  const float max_esdf_distance_vox = m->ep.max_esdf_distance_m / m->voxel_size;
  c.max_sq = max_esdf_distance_vox * max_esdf_distance_vox;
  c.max_esdf_distance_m = m->ep.max_esdf_distance_m;
  c.max_site_distance_m = m->ep.max_site_distance_vox * m->voxel_size;
  c.min_weight = m->ep.min_weight;
  c.block_size = m->block_size;
  // OccupancySiteFunctor (esdf_integrator.cu:71-75, 140-170)
  c.from_occupancy = m->projective_layer_type == NVB_PROJECTIVE_OCCUPANCY ? 1 : 0;
  c.occupied_threshold_log_odds = logOddsFromProbability(m->ep.occupied_threshold);
  return c;
}

// logOddsFromProbability (core/log_odds.h:23-30): clamp to [1e-3, 1 - 1e-3], then log(p / (1 - p)). The reference
// evaluates it on the host too (integrator members and setters), so glibc's logf is the function to match.
float logOddsFromProbability(float p) {
  p = fmaxf(1e-3f, fminf(p, 1.0f - 1e-3f));
  return logf(p / (1.0f - p));
}

Rigid rigidFromColMajor(const float* T) {
  Rigid r;
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) r.r[i][j] = T[j * 4 + i];
    r.t[i] = T[12 + i];
  }
  return r;
}

// Transform::inverse() for an isometry: R' = R^T, t' = -(R^T t).
Rigid invertRigid(const Rigid& T) {
  Rigid o;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) o.r[i][j] = T.r[j][i];
  for (int i = 0; i < 3; i++) o.t[i] = -sum3(o.r[i][0] * T.t[0], o.r[i][1] * T.t[1], o.r[i][2] * T.t[2]);
  return o;
}

// Camera::getViewAABB (src/sensors/camera.cpp:31-83) + applyWorkspaceBounds
// (src/geometry/workspace_bounds.cpp:20-61) + the block-index AABB of
// getBlocksInImageViewRaycast (view_calculator_impl.cuh:137-156).
// Returns false when the workspace-clipped AABB is empty.
bool computeViewGrid(const NvbCamera& cam, const Rigid& T_L_C, float block_size, float max_dist,
                     const NvbTsdfParams& P, ViewGrid* g, long long* cells) {
  const float w = (float)cam.width, h = (float)cam.height;
  const float ux[4] = {0.0f, w, w, 0.0f};
  const float vy[4] = {0.0f, 0.0f, h, h};
  Vec3 ray[4];
  for (int k = 0; k < 4; k++) {
    // Camera::vectorFromImagePlaneCoordinates (sensors/internal/impl/camera_impl.h:89-104)
    float nx = (ux[k] - cam.cu) / cam.fu, ny = (vy[k] - cam.cv) / cam.fv;
    if (cam.has_distortion) removeDistortion(cam, nx, ny);
    ray[k] = Vec3{nx, ny, 1.0f};
  }
  const int order[4] = {2, 1, 0, 3};
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int k = 0; k < 8; k++) {
    const float d = (k < 4) ? 0.0f : max_dist;
    const Vec3 r = ray[order[k & 3]];
    const Vec3 c = transformPoint(T_L_C, Vec3{d * r.x, d * r.y, d * r.z});
    const float cl[3] = {c.x, c.y, c.z};
    for (int i = 0; i < 3; i++) {
      lo[i] = std::min(lo[i], cl[i]);
      hi[i] = std::max(hi[i], cl[i]);
    }
  }
  if (P.workspace_bounds_type == NVB_WS_HEIGHT_BOUNDS) {
    lo[2] = std::max(lo[2], P.workspace_min[2]);
    hi[2] = std::min(hi[2], P.workspace_max[2]);
  } else if (P.workspace_bounds_type == NVB_WS_BOUNDING_BOX) {
    for (int i = 0; i < 3; i++) {
      lo[i] = std::max(P.workspace_min[i], lo[i]);
      hi[i] = std::min(P.workspace_max[i], hi[i]);
    }
  }
  if (lo[0] > hi[0] || lo[1] > hi[1] || lo[2] > hi[2]) return false;
  const int3 mn = blockIndexFromPosition(block_size, Vec3{lo[0], lo[1], lo[2]});
  const int3 mx = blockIndexFromPosition(block_size, Vec3{hi[0], hi[1], hi[2]});
  g->min_index = mn;
  g->size = make_int3(mx.x - mn.x + 1, mx.y - mn.y + 1, mx.z - mn.z + 1);
  *cells = (long long)g->size.x * g->size.y * g->size.z;
  if (*cells <= 0 || *cells > 0x7fffffffll) {
    *cells = -1;
    return false;
  }
  g->linear_size = (int)*cells;
  g->num_words = (g->linear_size + 31) / 32;
  return true;
}

void beginStageOn(NvbMapper* m, int stage, cudaStream_t st) {
  if (!m->profiling) return;
  StageEvent ev;
  ev.stage = stage;
  cudaEventCreate(&ev.start), cudaEventCreate(&ev.stop);
  cudaEventRecord(ev.start, st);
  m->stage_events.push_back(ev);
}
void endStageOn(NvbMapper* m, cudaStream_t st) {
  if (!m->profiling) return;
  cudaEventRecord(m->stage_events.back().stop, st);
}
void beginStage(NvbMapper* m, int stage) { beginStageOn(m, stage, m->stream); }
void endStage(NvbMapper* m) { endStageOn(m, m->stream); }
void collectStages(NvbMapper* m) {
  for (auto& ev : m->stage_events) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ev.start, ev.stop) == cudaSuccess) {
      m->stage_ms[ev.stage] += ms;
      m->stage_calls[ev.stage]++;
    }
    cudaEventDestroy(ev.start), cudaEventDestroy(ev.stop);
  }
  m->stage_events.clear();
}

// Poll the asynchronous read-backs of *tsdf.count to tighten the host-side bound.
// count <= confirmed + (cells enqueued since the confirmed read-back).
void pollCounts(NvbMapper* m) {
  for (int k = 0; k < kCountRing; k++) {
    if (m->count_pending[k] && cudaEventQuery(m->count_events[k]) == cudaSuccess) {
      m->count_pending[k] = false;
      if (m->count_cum_at[k] > m->confirmed_cum) {
        m->confirmed_cum = m->count_cum_at[k];
        m->tsdf_count_confirmed = m->h_count_ring[k];
      }
    }
  }
  const long long ub = (long long)m->tsdf_count_confirmed + (m->cells_cum - m->confirmed_cum);
  m->tsdf_count_ub = (int)std::min<long long>(ub, 0x7fffffff);
}

int ensureTsdfCapacity(NvbMapper* m, long long new_cells) {
  pollCounts(m);
  if ((long long)m->tsdf_count_ub + new_cells <= m->tsdf.capacity) return NVB_OK;
  // refine the bound with a synchronous read
  NVB_CUDA(syncAll(m));
  int count = 0;
  NVB_CUDA(cudaMemcpy(&count, m->tsdf.count, sizeof(int), cudaMemcpyDeviceToHost));
  for (int k = 0; k < kCountRing; k++) m->count_pending[k] = false;
  m->tsdf_count_confirmed = count;
  m->confirmed_cum = m->cells_cum;
  m->tsdf_count_ub = count;
  if ((long long)count + new_cells <= m->tsdf.capacity) return NVB_OK;
  long long cap = m->tsdf.capacity;
  while (cap < (long long)count + new_cells) cap *= 2;
  if (cap > (1ll << 28)) return fail(NVB_ERR_CAPACITY, "TSDF layer would exceed 2^28 blocks");
  const int old = m->tsdf.capacity;
  int rc = growLayer(m, &m->tsdf, (int)cap);
  if (rc) return rc;
  return allocTsdfSide(m, old, (int)cap);
}

// The explicit-list entry points are synchronous: afterwards the real fill level of the ESDF slab is known and replaces
// the running sum of list lengths (which would otherwise grow the slab without need in a long session).
int tightenEsdfBound(NvbMapper* m) {
  int count = 0;
  NVB_CUDA(cudaMemcpy(&count, m->esdf.count, sizeof(int), cudaMemcpyDeviceToHost));
  const int from_tsdf = std::min(m->tsdf_count_ub, m->tsdf.capacity);
  m->esdf_extra_ub = std::max(0, std::min(count, m->esdf.capacity) - from_tsdf);
  return NVB_OK;
}

int ensureEsdfCapacity(NvbMapper* m, long long needed_total) {
  if (needed_total <= m->esdf.capacity) return NVB_OK;
  long long cap = m->esdf.capacity;
  while (cap < needed_total) cap *= 2;
  if (cap > (1ll << 28)) return fail(NVB_ERR_CAPACITY, "ESDF layer would exceed 2^28 blocks");
  const int old = m->esdf.capacity;
  int rc = growLayer(m, &m->esdf, (int)cap);
  if (rc) return rc;
  return allocEsdfScratch(m, old, (int)cap);
}

int ensureFrameScratch(NvbMapper* m, const ViewGrid& g) {
  if ((size_t)g.num_words > m->bits_words_cap) {
    NVB_CUDA(syncAll(m));
    const size_t cap = (size_t)(1.5 * g.num_words) + 64;  // kBufferExpansionFactor, view_calculator_impl.cuh:159
    if (m->bits) cudaFree(m->bits);
    NVB_CUDA(cudaMalloc(&m->bits, cap * sizeof(unsigned int)));
    NVB_CUDA(cudaMemsetAsync(m->bits, 0, cap * sizeof(unsigned int), m->stream));
    m->bits_words_cap = cap;
  }
  if (g.linear_size > m->frame_cap) {
    NVB_CUDA(syncAll(m));
    const int cap = (int)std::min<long long>((long long)(1.5 * g.linear_size) + 64, 0x7fffffff);
    if (m->frame_blocks) cudaFree(m->frame_blocks);
    NVB_CUDA(cudaMalloc(&m->frame_blocks, (size_t)cap * sizeof(int4)));
    m->frame_cap = cap;
  }
  const int tiles = compactNumTiles(g);
  if (tiles > m->tile_cap) {
    NVB_CUDA(syncAll(m));
    const int cap = tiles * 2;
    if (m->tile_state) cudaFree(m->tile_state);
    NVB_CUDA(cudaMalloc(&m->tile_state, (size_t)cap * sizeof(unsigned long long)));
    NVB_CUDA(cudaMemsetAsync(m->tile_state, 0, (size_t)cap * sizeof(unsigned long long), m->stream));
    m->tile_cap = cap;
  }
  return NVB_OK;
}

int ensureStaging(NvbMapper* m, size_t pixels) {
  if (pixels <= m->stage_pixels) return NVB_OK;
  NVB_CUDA(syncAll(m));
  NVB_CUDA(cudaStreamSynchronize(m->copy_stream));
  for (int k = 0; k < kStagingBuffers; k++) {
    if (m->depth_stage[k]) cudaFree(m->depth_stage[k]);
    if (m->mask_stage[k]) cudaFree(m->mask_stage[k]);
    NVB_CUDA(cudaMalloc(&m->depth_stage[k], pixels * sizeof(float)));
    NVB_CUDA(cudaMalloc(&m->mask_stage[k], pixels));
    m->stage_used[k] = false;
  }
  m->stage_pixels = pixels;
  return NVB_OK;
}

int checkDeviceError(NvbMapper* m) {
  int err = 0;
  NVB_CUDA(cudaMemcpy(&err, m->error_dev, sizeof(int), cudaMemcpyDeviceToHost));
  if (err) {
    cudaMemsetAsync(m->error_dev, 0, sizeof(int), m->stream);
    cudaStreamSynchronize(m->stream);
    if (err & 2) return fail(NVB_ERR_INDEX_RANGE, "a block index does not fit the 21-bit hash key");
    if (err & 4) return fail(NVB_ERR_CAPACITY, "a block-list segment of the multi-GPU merge overflowed (nvb_mapper_append_frame_blocks)");
    // Roll the overflow back: find-or-insert left the keys it could not serve in the hash (value -1) and the fill level
    // above the capacity. Clamp the level and rebuild the hashes from the live slots, so that the same indices can be
    // allocated again once the caller has made room (or after the next growth).
    syncAll(m);
    DevLayer* layers[] = {&m->tsdf, &m->esdf, &m->freespace, &m->color, &m->mesh};
    for (DevLayer* L : layers) {
      if (!L->blocks) continue;
      int count = 0;
      if (cudaMemcpy(&count, L->count, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess || count <= L->capacity) continue;
      cudaMemcpyAsync(L->count, &L->capacity, sizeof(int), cudaMemcpyHostToDevice, m->stream);
      launchFillU64(L->hash.keys, kEmptyKey, (size_t)L->hash.mask + 1, m->stream);
      launchRehash(*L, L->capacity, m->stream);
      cudaStreamSynchronize(m->stream);
    }
    return fail(NVB_ERR_CAPACITY, "a layer slab overflowed on the device");
  }
  return NVB_OK;
}

int validateFrameArgs(const NvbMapper* m, const float* depth, int rows, int cols, const float* T, const NvbCamera* cam) {
  if (!m || !depth || !T || !cam) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  if (rows <= 0 || cols <= 0) return fail(NVB_ERR_INVALID_ARGUMENT, "image must have positive size");
  if (!(cam->fu != 0.0f) || !(cam->fv != 0.0f)) return fail(NVB_ERR_INVALID_ARGUMENT, "camera focal length is zero");
  return NVB_OK;
}

// arePosesClose (C/src/geometry/transforms.cpp:20-36) in binary32, Eigen::AngleAxisf(R).angle() through the quaternion.
bool posesClose(const float* T1, const float* T2, float tol_m, float tol_deg) {
  auto R_ = [](const float* T, int i, int j) { return T[j * 4 + i]; };
  float inv[16] = {0};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) inv[j * 4 + i] = R_(T1, j, i);
  for (int i = 0; i < 3; i++) inv[12 + i] = -(inv[0 * 4 + i] * T1[12] + (inv[1 * 4 + i] * T1[13] + inv[2 * 4 + i] * T1[14]));
  float R[3][3], t[3];
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) R[i][j] = (R_(inv, i, 0) * R_(T2, 0, j) + R_(inv, i, 1) * R_(T2, 1, j)) + R_(inv, i, 2) * R_(T2, 2, j);
    t[i] = ((R_(inv, i, 0) * T2[12] + R_(inv, i, 1) * T2[13]) + R_(inv, i, 2) * T2[14]) + inv[12 + i];
  }
  if (std::sqrt(t[0] * t[0] + (t[1] * t[1] + t[2] * t[2])) > tol_m) return false;
  float w, x, y, z;
  const float tr = R[0][0] + R[1][1] + R[2][2];
  if (tr > 0.0f) {
    float q = std::sqrt(tr + 1.0f);
    w = 0.5f * q;
    q = 0.5f / q;
    x = (R[2][1] - R[1][2]) * q, y = (R[0][2] - R[2][0]) * q, z = (R[1][0] - R[0][1]) * q;
  } else {
    int i = 0;
    if (R[1][1] > R[0][0]) i = 1;
    if (R[2][2] > R[i][i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    float q = std::sqrt(R[i][i] - R[j][j] - R[k][k] + 1.0f);
    float v[3];
    v[i] = 0.5f * q;
    q = 0.5f / q;
    w = (R[k][j] - R[j][k]) * q;
    v[j] = (R[j][i] + R[i][j]) * q;
    v[k] = (R[k][i] + R[i][k]) * q;
    x = v[0], y = v[1], z = v[2];
  }
  const float n = std::sqrt(x * x + (y * y + z * z));
  const float angle = n != 0.0f ? 2.0f * std::atan2(n, std::fabs(w)) : 0.0f;
  const float deg = (float)((double)(angle * 180.0f) / 3.14159265358979323846);
  return !(std::fabs(deg) > tol_deg);
}
// operator==(Camera, Camera) (C/include/nvblox/sensors/internal/impl/camera_impl.h:134-156)
bool camerasEqual(const NvbCamera& a, const NvbCamera& b) {
  bool same = std::fabs((double)(a.fu - b.fu)) <= 0.1 && std::fabs((double)(a.fv - b.fv)) <= 0.1 &&
              std::fabs((double)(a.cu - b.cu)) <= 0.1 && std::fabs((double)(a.cv - b.cv)) <= 0.1 && a.width == b.width &&
              a.height == b.height && (a.has_distortion != 0) == (b.has_distortion != 0);
  if (same && a.has_distortion && b.has_distortion)
    same = a.k1 == b.k1 && a.k2 == b.k2 && a.k3 == b.k3 && a.k4 == b.k4 && a.k5 == b.k5 && a.k6 == b.k6 && a.p1 == b.p1 && a.p2 == b.p2;
  return same;
}

// DepthPreprocessor's invalid_depth_threshold_ / invalid_depth_value_ (include/nvblox/sensors/depth_preprocessing.h)
constexpr float kInvalidDepthThreshold = 1e-2f;
constexpr float kInvalidDepthValue = 0.0f;
constexpr int kMaxDilations = 64;

// The depth-integration chain for one frame, enqueued on m->stream.
int enqueueFrame(NvbMapper* m, const float* depth, const unsigned char* mask, int mask_mode, int memory, int rows,
                 int cols, const float* T_L_C_cm, const NvbCamera* cam, float block_size, float trunc_m,
                 float max_dist, bool integrate) {
  const Rigid T_L_C = rigidFromColMajor(T_L_C_cm);
  ViewGrid grid{};
  long long cells = 0;
  // ViewpointCache::getCachedResult (view_calculator_impl.h:120-155): keyed on the pose and the sensor only
  int cache_hit = -1;
  if (integrate && m->cache_last_viewpoint)
    for (int i = 0; i < m->vc_n && cache_hit < 0; i++)
      if (posesClose(T_L_C_cm, m->vc_T[i], 0.001f, 0.1f) && camerasEqual(*cam, m->vc_cam[i])) cache_hit = i;
  bool visible;
  if (cache_hit >= 0) {
    grid = m->vc_grid[cache_hit], cells = m->vc_cells[cache_hit], visible = true;
  } else {
    visible = computeViewGrid(*cam, T_L_C, block_size, max_dist, m->tp, &grid, &cells);
  }
  if (!visible) {
    if (cells < 0) return fail(NVB_ERR_CAPACITY, "view AABB has more than 2^31 blocks");
    NVB_CUDA(cudaMemsetAsync(m->frame_count, 0, sizeof(int), m->stream));  // empty workspace -> empty list
    return NVB_OK;
  }
  int rc;
  if ((rc = ensureFrameScratch(m, grid))) return rc;
  if (integrate && (rc = ensureTsdfCapacity(m, cells))) return rc;

  // Inputs: device pointers are used in place; host buffers go through a small
  // ring of staging buffers on a copy stream so the upload of frame k+1 overlaps
  // the kernels of frame k.
  const float* depth_dev = depth;
  const unsigned char* mask_dev = mask;
  int stage_slot = -1;
  if (memory == NVB_MEM_HOST) {
    const size_t pixels = (size_t)rows * cols;
    if ((rc = ensureStaging(m, pixels))) return rc;
    stage_slot = (int)(m->frame_seq % kStagingBuffers);
    if (m->stage_used[stage_slot]) NVB_CUDA(cudaStreamWaitEvent(m->copy_stream, m->stage_consumed[stage_slot], 0));
    NVB_CUDA(cudaMemcpyAsync(m->depth_stage[stage_slot], depth, pixels * sizeof(float), cudaMemcpyHostToDevice,
                             m->copy_stream));
    if (mask)
      NVB_CUDA(cudaMemcpyAsync(m->mask_stage[stage_slot], mask, pixels, cudaMemcpyHostToDevice, m->copy_stream));
    NVB_CUDA(cudaEventRecord(m->stage_copied[stage_slot], m->copy_stream));
    NVB_CUDA(cudaStreamWaitEvent(m->stream, m->stage_copied[stage_slot], 0));
    depth_dev = m->depth_stage[stage_slot];
    mask_dev = mask ? m->mask_stage[stage_slot] : nullptr;
  }
  m->frame_seq++;
  if (integrate && m->do_depth_preprocessing) {
    // Mapper::preprocessDepthImageAsync (src/mapper/mapper.cpp:335-352): the integrators and the saved last view
    // both see the dilated copy (mapper_impl.h:38-76)
    // CHECK_GE(rows, 3), CHECK_GE(cols, 3) (src/sensors/depth_preprocessing.cpp:64-65)
    if (rows < 3 || cols < 3) return fail(NVB_ERR_INVALID_ARGUMENT, "depth preprocessing needs an image of at least 3x3");
    const size_t pixels = (size_t)rows * cols;
    if (m->pre_depth_cap < pixels) {
      NVB_CUDA(syncAll(m));
      if (m->pre_depth) cudaFree(m->pre_depth);
      m->pre_depth = nullptr, m->pre_depth_cap = 0;
      NVB_CUDA(cudaMalloc(&m->pre_depth, pixels * sizeof(float)));
      m->pre_depth_cap = pixels;
    }
    launchDilateInvalid(depth_dev, m->pre_depth, rows, cols, m->depth_preprocessing_num_dilations, kInvalidDepthThreshold,
                        kInvalidDepthValue, m->stream);
    m->launches++;
    depth_dev = m->pre_depth;
  }
  if (integrate && m->keep_last_view) {
    const size_t pixels = (size_t)rows * cols;
    if (m->last_depth_cap < pixels) {
      NVB_CUDA(syncAll(m));
      if (m->last_depth) cudaFree(m->last_depth);
      NVB_CUDA(cudaMalloc(&m->last_depth, pixels * sizeof(float)));
      m->last_depth_cap = pixels;
    }
    NVB_CUDA(cudaMemcpyAsync(m->last_depth, depth_dev, pixels * sizeof(float), cudaMemcpyDeviceToDevice, m->stream));
    m->last_rows = rows, m->last_cols = cols, m->last_cam = *cam, m->has_last_view = true;
    memcpy(m->last_T_L_C, T_L_C_cm, sizeof(m->last_T_L_C));
  }

  beginStage(m, 0);
  if (cache_hit >= 0) {
    // the cached view bitset instead of a raycast
    NVB_CUDA(cudaMemcpyAsync(m->bits, m->vc_bits[cache_hit], (size_t)grid.num_words * sizeof(unsigned int), cudaMemcpyDeviceToDevice,
                             m->stream));
  } else {
    launchViewRaycast(depth_dev, rows, cols, T_L_C, *cam, block_size, trunc_m, max_dist, m->tp.raycast_subsampling,
                      grid, m->bits, m->stream);
    m->launches++;
    if (integrate && m->cache_last_viewpoint) {
      // ViewpointCache::storeResultInCache (view_calculator_impl.h:157-174): newest first, the oldest of two is dropped
      if (m->vc_n == 2) m->vc_n = 1;
      if (m->vc_n == 1) {
        std::swap(m->vc_bits[0], m->vc_bits[1]), std::swap(m->vc_bits_cap[0], m->vc_bits_cap[1]);
        memcpy(m->vc_T[1], m->vc_T[0], sizeof(m->vc_T[0]));
        m->vc_cam[1] = m->vc_cam[0], m->vc_grid[1] = m->vc_grid[0], m->vc_cells[1] = m->vc_cells[0];
      }
      if (m->vc_bits_cap[0] < (size_t)grid.num_words) {
        NVB_CUDA(syncAll(m));
        if (m->vc_bits[0]) cudaFree(m->vc_bits[0]);
        m->vc_bits_cap[0] = (size_t)(1.5 * grid.num_words) + 64;
        NVB_CUDA(cudaMalloc(&m->vc_bits[0], m->vc_bits_cap[0] * sizeof(unsigned int)));
      }
      NVB_CUDA(cudaMemcpyAsync(m->vc_bits[0], m->bits, (size_t)grid.num_words * sizeof(unsigned int), cudaMemcpyDeviceToDevice,
                               m->stream));
      memcpy(m->vc_T[0], T_L_C_cm, sizeof(m->vc_T[0]));
      m->vc_cam[0] = *cam, m->vc_grid[0] = grid, m->vc_cells[0] = cells;
      m->vc_n++;
    }
  }
  endStage(m);

  beginStage(m, 1);
  CompactArgs ca{};
  ca.bits = m->bits;
  ca.grid = grid;
  ca.frame_blocks = m->frame_blocks;
  ca.frame_count = m->frame_count;
  ca.tile_state = m->tile_state;
  ca.ticket = m->ticket;
  ca.ticket_base = m->ticket_base;
  ca.epoch = ++m->epoch;
  ca.allocate = integrate ? 1 : 0;
  ca.layer = m->tsdf;
  ca.error = m->error_dev;
  ca.dirty = (integrate && m->tracker_initialized) ? m->dirty : nullptr;
  ca.todo_slots = m->todo_slots;
  ca.todo_count = m->todo_count;
  ca.dirty2 = (integrate && m->dirty_fs && m->fs_tracker_initialized) ? m->dirty_fs : nullptr;
  ca.todo2_slots = m->todo_fs_slots;
  ca.todo2_count = m->esdf_ints + kTodoFsCount;
  ca.dirty3 = (integrate && m->dirty_mesh && m->mesh_tracker_initialized) ? m->dirty_mesh : nullptr;
  ca.todo3_slots = m->todo_mesh_slots;
  ca.todo3_count = m->esdf_ints + kTodoMeshCount;
  launchCompactAllocate(ca, m->stream);
  if (compactUsesTickets(grid)) m->ticket_base += (unsigned int)compactNumTiles(grid);
  endStage(m);
  m->launches++;

  if (!integrate) {
    launchClearBits(m->bits, grid.num_words, m->stream);
    m->launches++;
  }
  if (integrate) {
    beginStage(m, 2);
    TsdfKernelParams p;
    p.block_size = block_size;
    p.voxel_size = block_size * (1.0f / kVps);       // blockSizeToVoxelSize, indexing_impl.h:26-29
    p.half_voxel_size = block_size * (0.5f / kVps);  // indexing_impl.h:75-77
    p.truncation_distance_m = trunc_m;
    p.max_integration_distance_m = max_dist;
    p.max_weight = m->tp.max_weight;
    p.invalid_depth_decay_factor = m->tp.invalid_depth_decay_factor;
    p.weighting_type = m->tp.weighting_type;
    const Rigid T_C_L = invertRigid(T_L_C);
    if (m->projective_layer_type == NVB_PROJECTIVE_OCCUPANCY) {
      // ProjectiveOccupancyIntegrator::setFunctorParameters (src/integrators/projective_occupancy_integrator.cu:41-49)
      OccKernelParams o;
      o.free_log_odds = logOddsFromProbability(m->op.free_region_occupancy_probability);
      o.occupied_log_odds = logOddsFromProbability(m->op.occupied_region_occupancy_probability);
      o.unobserved_log_odds = logOddsFromProbability(m->op.unobserved_region_occupancy_probability);
      o.occupied_half_width_m = m->op.occupied_region_half_width_m;
      o.min_log_odds = logOddsFromProbability(0.01f);
      o.max_log_odds = logOddsFromProbability(0.99f);
      launchOccupancyIntegrate(m->frame_blocks, m->frame_count, m->tsdf.blocks, depth_dev, mask_dev, mask_mode, rows,
                               cols, T_C_L, *cam, p, o, m->num_sms, m->bits, grid.num_words, m->stream);
    } else {
      // the slab's tensor descriptor follows reallocations (growLayer) by being re-encoded when base or capacity changed
      if (tsdfUseTma() && (m->tsdf_tmap.base != m->tsdf.blocks || m->tsdf_tmap.capacity != m->tsdf.capacity)) {
        if (encodeBlockTensorMap(&m->tsdf_tmap, m->tsdf.blocks, m->tsdf.capacity, kTsdfBlockBytes))
          return fail(NVB_ERR_CUDA, "cuTensorMapEncodeTiled failed for the TSDF slab");
      }
      launchTsdfIntegrate(m->frame_blocks, m->frame_count, m->tsdf.blocks, depth_dev, mask_dev, mask_mode, rows, cols,
                          T_C_L, *cam, p, m->num_sms, m->bits, grid.num_words, tsdfUseTma() ? &m->tsdf_tmap : nullptr, m->stream);
    }
    endStage(m);
    m->launches++;
    // asynchronous read-back of the slab fill level for the host-side capacity bound
    m->cells_cum += cells;
    const int k = m->count_ring_head;
    if (!m->count_pending[k]) {
      m->count_ring_head = (k + 1) % kCountRing;
      NVB_CUDA(cudaMemcpyAsync(m->h_count_ring + k, m->tsdf.count, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
      NVB_CUDA(cudaEventRecord(m->count_events[k], m->stream));
      m->count_pending[k] = true;
      m->count_cum_at[k] = m->cells_cum;
    }
    m->tsdf_count_ub = (int)std::min<long long>((long long)m->tsdf_count_ub + cells, 0x7fffffff);
  }
  if (stage_slot >= 0) {
    NVB_CUDA(cudaEventRecord(m->stage_consumed[stage_slot], m->stream));
    m->stage_used[stage_slot] = true;
  }
  return NVB_OK;
}

// The error word, copied to pinned host memory behind everything that is enqueued so far on both streams: after the next
// syncAll the host knows whether any kernel raised an error without a blocking copy of its own (a 4-byte cudaMemcpy is a
// full host round trip, ~10 us, and the synchronous API paid two of them per frame).
constexpr int kHostErrMain = 8, kHostErrEsdf = 9;
int enqueueErrorCopies(NvbMapper* m) {
  m->h_ints[kHostErrMain] = 0, m->h_ints[kHostErrEsdf] = 0;
  NVB_CUDA(cudaMemcpyAsync(m->h_ints + kHostErrMain, m->error_dev, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
  if (m->esdf_stream)
    NVB_CUDA(cudaMemcpyAsync(m->h_ints + kHostErrEsdf, m->error_dev, sizeof(int), cudaMemcpyDeviceToHost, m->esdf_stream));
  return NVB_OK;
}
// After enqueueErrorCopies + syncAll.
int checkPrefetchedError(NvbMapper* m) {
  if ((m->h_ints[kHostErrMain] | m->h_ints[kHostErrEsdf]) == 0) return NVB_OK;
  return checkDeviceError(m);  // slow path: re-reads the word, rolls back, reports
}

// Count + list of the last frame's blocks with ONE synchronisation: the count, a speculative prefix of the list (sized from
// the previous frame's count) and the error word travel to pinned host memory behind the frame's kernels.
int readFrameList(NvbMapper* m, int32_t* out_xyz, int32_t cap, int32_t* out_count) {
  int want = 0;
  if (out_xyz && cap > 0 && m->h_list) {
    want = std::min<long long>(std::min<long long>(cap, kHostListCap), (long long)m->last_frame_n * 3 / 2 + 512);
    want = std::min(want, m->frame_cap);
  }
  NVB_CUDA(cudaMemcpyAsync(m->h_ints, m->frame_count, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
  if (want > 0)
    NVB_CUDA(cudaMemcpyAsync(m->h_list, m->frame_blocks, (size_t)want * sizeof(int4), cudaMemcpyDeviceToHost, m->stream));
  int rc = enqueueErrorCopies(m);
  if (rc) return rc;
  NVB_CUDA(syncAll(m));
  const int n = m->h_ints[0];
  m->last_frame_n = n;
  if (out_count) *out_count = n;
  if (out_xyz && cap > 0 && n > 0) {
    const int k = std::min(n, cap);
    const int4* src = m->h_list;
    std::vector<int4> tmp;
    if (k > want) {  // the frame has more blocks than the speculative prefix: one more copy
      tmp.resize((size_t)k);
      NVB_CUDA(cudaMemcpy(tmp.data(), m->frame_blocks, (size_t)k * sizeof(int4), cudaMemcpyDeviceToHost));
      src = tmp.data();
    }
    for (int i = 0; i < k; i++) out_xyz[3 * i] = src[i].x, out_xyz[3 * i + 1] = src[i].y, out_xyz[3 * i + 2] = src[i].z;
  }
  return NVB_OK;
}

int enqueueEsdf(NvbMapper* m, const int* in_xyz_dev, int n_explicit, bool from_tracker, bool slice = false,
                const float* plane = nullptr) {
  int rc;
  int upper;
  // EsdfMode: a mapper's ESDF layer is 3-D or a 2-D slice, never both (src/mapper/mapper.cpp:410-415,436-441)
  const int want_mode = slice ? 2 : 1;
  if (m->esdf_mode != 0 && m->esdf_mode != want_mode)
    return fail(NVB_ERR_INVALID_ARGUMENT, "the ESDF layer of this mapper is already in the other mode (3-D vs 2-D slice)");
  m->esdf_mode = want_mode;
  if (from_tracker) {
    pollCounts(m);
    upper = std::min(m->tsdf_count_ub, m->tsdf.capacity);
  } else {
    upper = n_explicit;
    m->esdf_extra_ub += n_explicit;
  }
  if (upper <= 0) upper = 1;
  if ((rc = ensureEsdfCapacity(m, (long long)std::min(m->tsdf_count_ub, m->tsdf.capacity) + m->esdf_extra_ub))) return rc;
  if (slice) {
    // column set + column list sized to the projective layer
    const size_t want = (size_t)nextPow2(2ll * std::max(m->tsdf.capacity, upper));
    if (m->colset_n < want || m->cols_cap < std::max(m->tsdf.capacity, upper)) {
      NVB_CUDA(syncAll(m));
      if (m->colset) cudaFree(m->colset);
      if (m->cols) cudaFree(m->cols);
      NVB_CUDA(cudaMalloc(&m->colset, want * sizeof(unsigned long long)));
      m->colset_n = want;
      m->cols_cap = std::max(m->tsdf.capacity, upper);
      NVB_CUDA(cudaMalloc(&m->cols, (size_t)m->cols_cap * 2 * sizeof(int)));
    }
  }
  m->update_seq++;
  EsdfCtx c = makeEsdfCtx(m);
  c.tracker_dirty = from_tracker ? m->dirty : nullptr;
  c.tracker_todo_count = from_tracker ? m->todo_count : nullptr;
  if (slice) {
    // getBlockAndVoxelIndexFrom1DPositionInLayer (core/internal/impl/indexing_impl.h:105-115), on the host like the
    // reference (ConstantZColumnBoundsGetter's constructor and markSitesInSlice, esdf_integrator.cu:959-962)
    auto split = [&](float p, int* b, int* v) {
      const float inv = (float)(1.0 / (double)(m->block_size * (1.0f / kVps)));
      *b = (int)std::floor(p / m->block_size);
      *v = std::min((int)((p - m->block_size * (float)*b) * inv), kVps - 1);
    };
    split(m->sp.slice_min_height_m, &c.slice_min_bz, &c.slice_min_vz);
    split(m->sp.slice_max_height_m, &c.slice_max_bz, &c.slice_max_vz);
    split(m->sp.slice_height_m, &c.slice_out_bz, &c.slice_out_vz);
    if (c.slice_max_bz < c.slice_min_bz) return fail(NVB_ERR_INVALID_ARGUMENT, "slice_max_height below slice_min_height");
    if (plane) {
      c.slice_planar = 1;
      c.plane_nx = plane[0], c.plane_ny = plane[1], c.plane_nz = plane[2], c.plane_d = plane[3];
      if (std::fabs(c.plane_nz) < 1e-4f) c.plane_nx = 0.0f, c.plane_ny = 0.0f, c.plane_nz = 1.0f, c.plane_d = 0.0f;  // checkForVerticalPlane
      c.slice_above_plane_m = m->sp.slice_height_above_plane_m;
      c.slice_thickness_m = m->sp.slice_height_thickness_m;
    }
  }
  int launches = 0;
  cudaError_t e;
  auto allocAndMark = [&](cudaStream_t st) {
    if (slice) {
      launchEsdfSliceAllocateAndMark(c, from_tracker ? nullptr : in_xyz_dev, from_tracker ? m->todo_slots : nullptr,
                                     from_tracker ? m->todo_count : nullptr, from_tracker ? upper : n_explicit, m->num_sms, st);
      m->launches += 3;
    } else {
      if (from_tracker) launchEsdfAllocate(c, nullptr, m->todo_slots, m->todo_count, upper, st);
      else launchEsdfAllocate(c, in_xyz_dev, nullptr, nullptr, n_explicit, st);
      launchEsdfMark(c, upper, m->num_sms, st);
      m->launches += 2;
    }
  };
  if (m->esdf_persistent) {
    // The whole ESDF chain (allocate, mark, clear, wavefront) runs back to back on the side stream; the frame's
    // critical path has no cross-stream hand-over. `stream` only waits for the mark kernel: after it nothing on
    // the side stream reads the projective layer or the tracker, so the next frame's raycast / compaction /
    // TSDF update overlaps the clear pass and the wavefront. (The previous wavefront is ordered before this
    // chain by the side stream itself.)
    cudaStream_t es = m->esdf_stream;
    NVB_CUDA(cudaEventRecord(m->esdf_ready, m->stream));  // projective layer + tracker of this update are final
    NVB_CUDA(cudaStreamWaitEvent(es, m->esdf_ready, 0));
    beginStageOn(m, 3, es);
    allocAndMark(es);
    endStageOn(m, es);
    NVB_CUDA(cudaEventRecord(m->mark_done, es));
    NVB_CUDA(cudaStreamWaitEvent(m->stream, m->mark_done, 0));
    beginStageOn(m, 4, es);
    m->launches += launchEsdfClear(c, m->esdf.capacity, m->num_sms, es);
    endStageOn(m, es);
    beginStageOn(m, 5, es);
    e = m->esdf_persistent == 3   ? launchEsdfComputeX(c, m->num_sms, m->esdf_reserved_sms, es, &launches)
        : m->esdf_persistent == 2 ? launchEsdfComputeGes(c, m->num_sms, es, &launches)
                                  : launchEsdfComputePersistent(c, m->num_sms, es, &launches);
    endStageOn(m, es);
    if (e == cudaSuccess) {
      NVB_CUDA(cudaEventRecord(m->esdf_done, es));
      m->esdf_in_flight = true;
    }
  } else {
    NVB_CUDA(joinEsdf(m));
    beginStage(m, 3);
    allocAndMark(m->stream);
    endStage(m);
    beginStage(m, 4);
    m->launches += launchEsdfClear(c, m->esdf.capacity, m->num_sms, m->stream);
    endStage(m);
    beginStage(m, 5);
    e = runEsdfComputeHostLoop(c, m->num_sms, m->stream, &launches);
    endStage(m);
  }
  m->launches += launches;
  if (e != cudaSuccess) return fail(NVB_ERR_CUDA, std::string("ESDF compute launch: ") + cudaGetErrorString(e));
  return NVB_OK;
}

}  // namespace

// ---------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------
extern "C" {

const char* nvb_last_error(void) { return g_last_error.c_str(); }
const char* nvb_version(void) { return "nvblox_b200 0.1 (sm_100a)"; }

int32_t nvb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

void nvb_default_mapper_options(NvbMapperOptions* o) {
  if (!o) return;
  o->voxel_size_m = 0.05f;
  o->device = 0;
  o->tsdf_capacity_blocks = kDefaultCapacity;
  o->esdf_capacity_blocks = kDefaultCapacity;
  o->esdf_persistent = 3;
  o->projective_layer_type = NVB_PROJECTIVE_TSDF;
  o->keep_last_view = 0;
}
void nvb_default_tsdf_params(NvbTsdfParams* p) {
  if (!p) return;
  memset(p, 0, sizeof(*p));
  p->truncation_distance_vox = 4.0f;
  p->max_integration_distance_m = 7.0f;
  p->max_weight = 5.0f;
  p->invalid_depth_decay_factor = -1.0f;
  p->weighting_type = NVB_WEIGHT_INVERSE_SQUARE;
  p->raycast_subsampling = 4;
  p->workspace_bounds_type = NVB_WS_UNBOUNDED;
}
void nvb_default_esdf_params(NvbEsdfParams* p) {
  if (!p) return;
  p->max_esdf_distance_m = 2.0f;
  p->max_site_distance_vox = 1.0f;
  p->min_weight = 1e-4f;
  p->occupied_threshold = 0.5f;
}
void nvb_default_occupancy_params(NvbOccupancyParams* p) {
  if (!p) return;
  // integrators/occupancy_integrator_params.h:21-40
  p->free_region_occupancy_probability = 0.3f;
  p->occupied_region_occupancy_probability = 0.7f;
  p->unobserved_region_occupancy_probability = 0.5f;
  p->occupied_region_half_width_m = 0.1f;
}

// Everything nvb_mapper_create allocates, on an already constructed object (so that a failure half way can be undone
// by nvb_mapper_destroy).
static int createMapperResources(const NvbMapperOptions* opts, NvbMapper* m) {
  m->device = opts->device;
  cudaDeviceProp prop;
  NVB_CUDA(cudaGetDeviceProperties(&prop, opts->device));
  m->num_sms = prop.multiProcessorCount;
  m->voxel_size = opts->voxel_size_m;
  m->block_size = opts->voxel_size_m * (float)kVps;  // voxelSizeToBlockSize (indexing_impl.h:22-24)
  nvb_default_tsdf_params(&m->tp);
  nvb_default_esdf_params(&m->ep);
  nvb_default_occupancy_params(&m->op);
  nvb_default_tsdf_decay_params(&m->tdp);
  nvb_default_occupancy_decay_params(&m->odp);
  nvb_default_freespace_params(&m->fp);
  nvb_default_esdf_slice_params(&m->sp);
  nvb_default_color_params(&m->cp);
  m->projective_layer_type = opts->projective_layer_type;
  m->keep_last_view = opts->keep_last_view ? 1 : 0;
  m->esdf_persistent = opts->esdf_persistent;
  // A/B switch for measurements: 0 host loop, 1 four-phase wavefront, 2 gather-emulate-sweep wavefront
  if (const char* e = getenv("NVB_ESDF_MODE")) m->esdf_persistent = atoi(e);
  if (const char* e = getenv("NVB_WAVEX_RESERVED_SMS")) m->esdf_reserved_sms = std::max(0, std::min(64, atoi(e)));
  if (const char* e = getenv("NVB_GES_SWITCH")) m->ges_switch = atoi(e);
  {
    const char* e = getenv("NVB_CLEAR_PRUNE");
    m->prune_default = m->esdf_persistent == 3 && !(e && atoi(e) == 0);
    m->prune_ok = m->prune_default;
  }
  NVB_CUDA(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
  NVB_CUDA(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
  {
    // The ESDF chain is the frame's critical path; the next frame's raycast / compaction / TSDF update only has to finish
    // before the next mark kernel. NVB_ESDF_STREAM_PRIORITY=0 switches the preference off (A/B).
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    const char* pe = getenv("NVB_ESDF_STREAM_PRIORITY");
    const int prio = (pe && atoi(pe) == 0) ? lo : hi;
    NVB_CUDA(cudaStreamCreateWithPriority(&m->esdf_stream, cudaStreamNonBlocking, prio));
  }
  NVB_CUDA(cudaEventCreateWithFlags(&m->esdf_ready, cudaEventDisableTiming));
  NVB_CUDA(cudaEventCreateWithFlags(&m->esdf_done, cudaEventDisableTiming));
  NVB_CUDA(cudaEventCreateWithFlags(&m->mark_done, cudaEventDisableTiming));
  const int tcap = opts->tsdf_capacity_blocks > 0 ? opts->tsdf_capacity_blocks : kDefaultCapacity;
  const int ecap = std::max(opts->esdf_capacity_blocks > 0 ? opts->esdf_capacity_blocks : kDefaultCapacity, tcap);
  int rc;
  if ((rc = allocLayer(&m->tsdf, tcap, m->projective_layer_type == NVB_PROJECTIVE_OCCUPANCY ? kOccBlockBytes : kTsdfBlockBytes,
                       m->stream))) return rc;
  if ((rc = allocLayer(&m->esdf, ecap, kEsdfBlockBytes, m->stream))) return rc;
  if (m->projective_layer_type == NVB_PROJECTIVE_TSDF_WITH_FREESPACE &&
      (rc = allocLayer(&m->freespace, tcap, kFreespaceBlockBytes, m->stream)))
    return rc;
  if ((rc = allocTsdfSide(m, 0, tcap))) return rc;
  if ((rc = allocEsdfScratch(m, 0, ecap))) return rc;
  NVB_CUDA(cudaMalloc(&m->esdf_ints, kNumInts * sizeof(int)));
  NVB_CUDA(cudaMemsetAsync(m->esdf_ints, 0, kNumInts * sizeof(int), m->stream));
  const int one = 1;
  NVB_CUDA(cudaMemcpyAsync(m->esdf_ints + kRingId, &one, sizeof(int), cudaMemcpyHostToDevice, m->stream));
  m->todo_count = m->esdf_ints + kTodoCount;
  m->frame_count = m->esdf_ints + kFrameCount;
  m->error_dev = m->esdf_ints + kError;
  NVB_CUDA(cudaMalloc(&m->clr_bits, 2048 * sizeof(unsigned int)));
  NVB_CUDA(cudaMalloc(&m->stats, 16 * sizeof(long long)));
  NVB_CUDA(cudaMemsetAsync(m->stats, 0, 16 * sizeof(long long), m->stream));
  NVB_CUDA(cudaMalloc(&m->phase_max, 4000 * sizeof(unsigned long long)));
  NVB_CUDA(cudaMemsetAsync(m->phase_max, 0, 4000 * sizeof(unsigned long long), m->stream));
  NVB_CUDA(cudaMalloc(&m->barrier, 64));
  NVB_CUDA(cudaMemsetAsync(m->barrier, 0, 64, m->stream));
  NVB_CUDA(cudaMalloc(&m->ticket, 64));
  NVB_CUDA(cudaMemsetAsync(m->ticket, 0, 64, m->stream));
  NVB_CUDA(cudaMallocHost(&m->h_ints, 64 * sizeof(int)));
  memset(m->h_ints, 0, 64 * sizeof(int));
  NVB_CUDA(cudaMallocHost(&m->h_list, (size_t)kHostListCap * sizeof(int4)));
  NVB_CUDA(cudaMallocHost(&m->h_count_ring, kCountRing * sizeof(int)));
  for (int k = 0; k < kCountRing; k++) {
    NVB_CUDA(cudaEventCreateWithFlags(&m->count_events[k], cudaEventDisableTiming));
    m->count_pending[k] = false;
    m->count_cum_at[k] = 0;
  }
  for (int k = 0; k < kStagingBuffers; k++) {
    NVB_CUDA(cudaEventCreateWithFlags(&m->stage_copied[k], cudaEventDisableTiming));
    NVB_CUDA(cudaEventCreateWithFlags(&m->stage_consumed[k], cudaEventDisableTiming));
  }
  NVB_CUDA(syncAll(m));
  return NVB_OK;
}

int32_t nvb_mapper_create(const NvbMapperOptions* opts, NvbMapper** out) {
  if (!opts || !out) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  if (!(opts->voxel_size_m > 0.0f)) return fail(NVB_ERR_INVALID_ARGUMENT, "voxel_size_m must be > 0");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(NVB_ERR_NO_DEVICE, "no CUDA device: the depth-integration path has no CPU fallback");
  }
  if (opts->device < 0 || opts->device >= ndev) return fail(NVB_ERR_INVALID_ARGUMENT, "bad device ordinal");
  if (opts->projective_layer_type != NVB_PROJECTIVE_TSDF && opts->projective_layer_type != NVB_PROJECTIVE_OCCUPANCY &&
      opts->projective_layer_type != NVB_PROJECTIVE_TSDF_WITH_FREESPACE)
    return fail(NVB_ERR_INVALID_ARGUMENT, "unknown projective_layer_type");
  NVB_CUDA(cudaSetDevice(opts->device));
  NvbMapper* m = new NvbMapper();
  const int rc = createMapperResources(opts, m);
  if (rc != NVB_OK) {
    // no leak on a failed create: the partial object goes through the normal destructor (null handles are skipped by
    // the CUDA runtime with an error code that is cleared here; the message of the original failure is kept)
    const std::string why = g_last_error;
    nvb_mapper_destroy(m);
    cudaGetLastError();
    g_last_error = why;
    return rc;
  }
  *out = m;
  return NVB_OK;
}

void nvb_mapper_destroy(NvbMapper* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  syncAll(m);
  cudaStreamSynchronize(m->copy_stream);
  collectStages(m);
  cudaEventDestroy(m->esdf_ready), cudaEventDestroy(m->esdf_done), cudaEventDestroy(m->mark_done);
  cudaStreamDestroy(m->esdf_stream);
  freeLayer(&m->tsdf), freeLayer(&m->esdf);
  if (m->freespace.blocks) freeLayer(&m->freespace);
  if (m->color.blocks) freeLayer(&m->color);
  cudaFree(m->color_work), cudaFree(m->color_synth), cudaFree(m->color_stage), cudaFree(m->color_mask_stage);
  cudaFree(m->dirty_fs), cudaFree(m->todo_fs_slots), cudaFree(m->fs_work), cudaFree(m->colset), cudaFree(m->cols);
  cudaFree(m->bits), cudaFree(m->frame_blocks), cudaFree(m->tile_state), cudaFree(m->ticket);
  for (int k = 0; k < kStagingBuffers; k++) {
    cudaFree(m->depth_stage[k]), cudaFree(m->mask_stage[k]);
    cudaEventDestroy(m->stage_copied[k]), cudaEventDestroy(m->stage_consumed[k]);
  }
  cudaFree(m->dirty), cudaFree(m->todo_slots);
  cudaFree(m->work), cudaFree(m->esdf_ints), cudaFree(m->upd_list), cudaFree(m->clr_list), cudaFree(m->clr_cand), cudaFree(m->cleared_list);
  cudaFree(m->ring_a), cudaFree(m->ring_b), cudaFree(m->stamp_a), cudaFree(m->stamp_b);
  cudaFree(m->nbr), cudaFree(m->seed_upd), cudaFree(m->seed_clr), cudaFree(m->psum);
  cudaFree(m->nbr27), cudaFree(m->shadow), cudaFree(m->cand_stamp), cudaFree(m->cand_a), cudaFree(m->cand_b);
  cudaFree(m->xslab), cudaFree(m->xrec), cudaFree(m->xcounts);
  cudaFree(m->dead), cudaFree(m->skip_stamp), cudaFree(m->dead_cleared_xyz), cudaFree(m->last_depth);
  cudaFree(m->pre_depth);
  cudaFree(m->union_list), cudaFree(m->union_list_count);
  if (m->mesh.blocks) freeLayer(&m->mesh);
  cudaFree(m->mesh_v), cudaFree(m->mesh_n), cudaFree(m->mesh_t), cudaFree(m->mesh_c), cudaFree(m->mesh_state);
  cudaFree(m->mesh_alt_v), cudaFree(m->mesh_alt_n), cudaFree(m->mesh_alt_t), cudaFree(m->mesh_alt_c);
  cudaFree(m->mesh_counts), cudaFree(m->mesh_offsets), cudaFree(m->mesh_xyz_dev), cudaFree(m->dirty_mesh), cudaFree(m->todo_mesh_slots);
  cudaFree(m->clr_bits), cudaFree(m->union_bits), cudaFree(m->union_state);
  cudaFree(m->vc_bits[0]), cudaFree(m->vc_bits[1]);
  cudaFree(m->stats), cudaFree(m->barrier), cudaFree(m->phase_max), cudaFree(m->xyz_upload);
  cudaFreeHost(m->h_ints), cudaFreeHost(m->h_count_ring), cudaFreeHost(m->h_list);
  for (int k = 0; k < kCountRing; k++) cudaEventDestroy(m->count_events[k]);
  cudaStreamDestroy(m->stream), cudaStreamDestroy(m->copy_stream);
  delete m;
}

int32_t nvb_mapper_clear(NvbMapper* m) {
  if (!m) return fail(NVB_ERR_INVALID_ARGUMENT, "null mapper");
  NVB_CUDA(cudaSetDevice(m->device));
  NVB_CUDA(syncAll(m));
  std::vector<DevLayer*> layers = {&m->tsdf, &m->esdf};
  if (m->freespace.blocks) layers.push_back(&m->freespace);
  if (m->color.blocks) layers.push_back(&m->color);
  for (DevLayer* L : layers) {
    int count = 0;
    NVB_CUDA(cudaMemcpy(&count, L->count, sizeof(int), cudaMemcpyDeviceToHost));
    count = std::min(count, L->capacity);
    // re-zero only the slots that were handed out: the slab invariant is "free slots are zero"
    NVB_CUDA(cudaMemsetAsync(L->blocks, 0, (size_t)count * L->block_bytes, m->stream));
    NVB_CUDA(cudaMemsetAsync(L->count, 0, sizeof(int), m->stream));
    NVB_CUDA(cudaMemsetAsync(L->free_count, 0, sizeof(int), m->stream));
    launchFillU64(L->hash.keys, kEmptyKey, (size_t)L->hash.mask + 1, m->stream);
  }
  NVB_CUDA(cudaMemsetAsync(m->dirty, 0, (size_t)m->tsdf.capacity * sizeof(int), m->stream));
  NVB_CUDA(cudaMemsetAsync(m->todo_count, 0, sizeof(int), m->stream));
  NVB_CUDA(cudaMemsetAsync(m->esdf_ints + kClearedCount, 0, sizeof(int), m->stream));
  NVB_CUDA(cudaMemsetAsync(m->esdf_ints + kClearedSeq, 0, sizeof(int), m->stream));
  NVB_CUDA(cudaMemsetAsync(m->esdf_ints + kDeadClearedCount, 0, sizeof(int), m->stream));
  NVB_CUDA(cudaMemsetAsync(m->seed_upd, 0, (size_t)m->esdf.capacity * sizeof(int), m->stream));
  NVB_CUDA(cudaMemsetAsync(m->seed_clr, 0, (size_t)m->esdf.capacity * sizeof(int), m->stream));
  NVB_CUDA(cudaMemsetAsync(m->psum, 0, 2 * (size_t)m->esdf.capacity * sizeof(int), m->stream));
  m->prune_ok = m->prune_default;  // an empty layer: every summary is exact again
  NVB_CUDA(cudaMemsetAsync(m->nbr, 0xFE, (size_t)m->esdf.capacity * 6 * sizeof(int), m->stream));
  NVB_CUDA(cudaMemsetAsync(m->nbr27, 0xFE, (size_t)m->esdf.capacity * 27 * sizeof(int), m->stream));
  NVB_CUDA(cudaMemsetAsync(m->error_dev, 0, sizeof(int), m->stream));
  m->tracker_initialized = false;
  m->fs_tracker_initialized = false;
  m->mesh_tracker_initialized = false;
  if (m->mesh.blocks) {
    NVB_CUDA(cudaMemsetAsync(m->mesh.blocks, 0, (size_t)m->mesh.capacity * m->mesh.block_bytes, m->stream));
    NVB_CUDA(cudaMemsetAsync(m->mesh.count, 0, sizeof(int), m->stream));
    NVB_CUDA(cudaMemsetAsync(m->mesh.free_count, 0, sizeof(int), m->stream));
    launchFillU64(m->mesh.hash.keys, kEmptyKey, (size_t)m->mesh.hash.mask + 1, m->stream);
    NVB_CUDA(cudaMemsetAsync(m->mesh_state, 0, kArenaInts * sizeof(int), m->stream));
    NVB_CUDA(cudaMemsetAsync(m->esdf_ints + kTodoMeshCount, 0, sizeof(int), m->stream));
    NVB_CUDA(cudaMemsetAsync(m->dirty_mesh, 0, (size_t)m->tsdf.capacity * sizeof(int), m->stream));
  }
  m->esdf_mode = 0;
  m->fs_last_update_ms = 0;
  NVB_CUDA(cudaMemsetAsync(m->esdf_ints + kTodoFsCount, 0, sizeof(int), m->stream));
  if (m->dirty_fs) NVB_CUDA(cudaMemsetAsync(m->dirty_fs, 0, (size_t)m->tsdf.capacity * sizeof(int), m->stream));
  m->has_last_view = false;
  m->tsdf_count_ub = 0, m->tsdf_count_confirmed = 0, m->esdf_extra_ub = 0;
  m->cells_cum = 0, m->confirmed_cum = 0;
  for (int k = 0; k < kCountRing; k++) m->count_pending[k] = false;
  NVB_CUDA(syncAll(m));
  return NVB_OK;
}

int32_t nvb_mapper_set_tsdf_params(NvbMapper* m, const NvbTsdfParams* p) {
  if (!m || !p) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  // CHECK_GT(max_weight, 0) etc. (src/integrators/projective_tsdf_integrator.cu:61-64)
  if (!(p->max_weight > 0.0f)) return fail(NVB_ERR_INVALID_ARGUMENT, "max_weight must be > 0");
  if (!(p->truncation_distance_vox > 0.0f)) return fail(NVB_ERR_INVALID_ARGUMENT, "truncation_distance_vox must be > 0");
  if (p->raycast_subsampling < 1) return fail(NVB_ERR_INVALID_ARGUMENT, "raycast_subsampling must be >= 1");
  if (p->weighting_type < 0 || p->weighting_type > NVB_WEIGHT_LINEAR_WITH_MAX)
    return fail(NVB_ERR_INVALID_ARGUMENT, "unknown weighting_type");
  if (p->workspace_bounds_type < 0 || p->workspace_bounds_type > NVB_WS_BOUNDING_BOX)
    return fail(NVB_ERR_INVALID_ARGUMENT, "unknown workspace_bounds_type");
  m->tp = *p;
  return NVB_OK;
}
int32_t nvb_mapper_get_tsdf_params(const NvbMapper* m, NvbTsdfParams* p) {
  if (!m || !p) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  *p = m->tp;
  return NVB_OK;
}
int32_t nvb_mapper_set_esdf_params(NvbMapper* m, const NvbEsdfParams* p) {
  if (!m || !p) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  // CHECK_GT in the setters (src/integrators/esdf_integrator.cu:55-68)
  if (!(p->max_esdf_distance_m > 0.0f) || !(p->max_site_distance_vox > 0.0f) || !(p->min_weight > 0.0f))
    return fail(NVB_ERR_INVALID_ARGUMENT, "ESDF parameters must be > 0");
  // occupied_threshold: CHECK_GE(0) / CHECK_LE(1) (esdf_integrator.cu:71-75)
  if (!(p->occupied_threshold >= 0.0f && p->occupied_threshold <= 1.0f))
    return fail(NVB_ERR_INVALID_ARGUMENT, "occupied_threshold must be a probability");
  m->ep = *p;
  return NVB_OK;
}
int32_t nvb_mapper_get_esdf_params(const NvbMapper* m, NvbEsdfParams* p) {
  if (!m || !p) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  *p = m->ep;
  return NVB_OK;
}
int32_t nvb_mapper_set_occupancy_params(NvbMapper* m, const NvbOccupancyParams* p) {
  if (!m || !p) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  // CHECK(value >= 0 && value <= 1) in the probability setters (src/integrators/projective_occupancy_integrator.cu:71-99);
  // the half width has no check there
  const float probs[3] = {p->free_region_occupancy_probability, p->occupied_region_occupancy_probability,
                          p->unobserved_region_occupancy_probability};
  for (float q : probs)
    if (!(q >= 0.0f && q <= 1.0f)) return fail(NVB_ERR_INVALID_ARGUMENT, "occupancy probabilities must be in [0, 1]");
  m->op = *p;
  return NVB_OK;
}
int32_t nvb_mapper_get_occupancy_params(const NvbMapper* m, NvbOccupancyParams* p) {
  if (!m || !p) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  *p = m->op;
  return NVB_OK;
}
void nvb_default_tsdf_decay_params(NvbTsdfDecayParams* p) {
  if (!p) return;
  // integrators/tsdf_decay_integrator_params.h:21-48, internal/decay_integrator_base_params.h:22-29
  p->decay_factor = 0.95f;
  p->decayed_weight_threshold = 1e-3f;
  p->set_free_distance_on_decayed = 0;
  p->free_distance_vox = 4.0f;
  p->deallocate_decayed_blocks = 1;
}
void nvb_default_occupancy_decay_params(NvbOccupancyDecayParams* p) {
  if (!p) return;
  // integrators/occupancy_decay_integrator_params.h:21-43
  p->free_region_decay_probability = 0.55f;
  p->occupied_region_decay_probability = 0.4f;
  p->decay_to_probability = 0.5f;
  p->deallocate_decayed_blocks = 1;
}
int32_t nvb_mapper_set_tsdf_decay_params(NvbMapper* m, const NvbTsdfDecayParams* p) {
  if (!m || !p) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  // CHECK_GT(decay_factor, 0) / CHECK_LT(decay_factor, 1) (src/integrators/tsdf_decay_integrator.cu:30-37)
  if (!(p->decay_factor > 0.0f && p->decay_factor < 1.0f)) return fail(NVB_ERR_INVALID_ARGUMENT, "decay_factor must be in (0, 1)");
  m->tdp = *p;
  return NVB_OK;
}
int32_t nvb_mapper_get_tsdf_decay_params(const NvbMapper* m, NvbTsdfDecayParams* p) {
  if (!m || !p) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  *p = m->tdp;
  return NVB_OK;
}
int32_t nvb_mapper_set_occupancy_decay_params(NvbMapper* m, const NvbOccupancyDecayParams* p) {
  if (!m || !p) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  // CHECKs of the setters (src/integrators/occupancy_decay_integrator.cu:24-57)
  if (!(p->free_region_decay_probability >= 0.5f && p->free_region_decay_probability <= 1.0f))
    return fail(NVB_ERR_INVALID_ARGUMENT, "free_region_decay_probability must be in [0.5, 1]");
  if (!(p->occupied_region_decay_probability >= 0.0f && p->occupied_region_decay_probability < 0.5f))
    return fail(NVB_ERR_INVALID_ARGUMENT, "occupied_region_decay_probability must be in [0, 0.5)");
  if (!(p->decay_to_probability >= 0.0f && p->decay_to_probability <= 1.0f))
    return fail(NVB_ERR_INVALID_ARGUMENT, "decay_to_probability must be a probability");
  m->odp = *p;
  return NVB_OK;
}
int32_t nvb_mapper_get_occupancy_decay_params(const NvbMapper* m, NvbOccupancyDecayParams* p) {
  if (!m || !p) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  *p = m->odp;
  return NVB_OK;
}

int32_t nvb_mapper_decay(NvbMapper* m, const NvbDecayExclusion* exclusion, const float* depth, int32_t depth_memory,
                         int32_t rows, int32_t cols, const float* T_L_C, const NvbCamera* cam, int32_t* removed_xyz_host,
                         int32_t cap, int32_t* out_count) {
  if (!m) return fail(NVB_ERR_INVALID_ARGUMENT, "null mapper");
  if (out_count) *out_count = 0;
  if (depth) {
    int rc = validateFrameArgs(m, depth, rows, cols, T_L_C, cam);
    if (rc) return rc;
  }
  if (exclusion && exclusion->num_excluded_blocks > 0 && !exclusion->excluded_blocks_xyz_host)
    return fail(NVB_ERR_INVALID_ARGUMENT, "exclusion list is null");
  NVB_CUDA(cudaSetDevice(m->device));
  NVB_CUDA(syncAll(m));  // the decay touches both layers: nothing of the ESDF chain may be in flight
  const bool occupancy = m->projective_layer_type == NVB_PROJECTIVE_OCCUPANCY;
  DevLayer& P = m->tsdf;
  // scratch sized to the layer
  if (m->dead_cap < P.capacity) {
    if (m->dead) cudaFree(m->dead);
    NVB_CUDA(cudaMalloc(&m->dead, (size_t)P.capacity * sizeof(int4)));
    m->dead_cap = P.capacity;
  }
  if (m->skip_cap < P.capacity) {
    if (m->skip_stamp) cudaFree(m->skip_stamp);
    NVB_CUDA(cudaMalloc(&m->skip_stamp, (size_t)P.capacity * sizeof(int)));
    NVB_CUDA(cudaMemsetAsync(m->skip_stamp, 0, (size_t)P.capacity * sizeof(int), m->stream));
    m->skip_cap = P.capacity;
  }
  if (m->dead_cleared_cap < m->esdf.capacity) {
    int* q = nullptr;
    NVB_CUDA(cudaMalloc(&q, (size_t)m->esdf.capacity * 3 * sizeof(int)));
    if (m->dead_cleared_xyz) {
      NVB_CUDA(cudaMemcpy(q, m->dead_cleared_xyz, (size_t)m->dead_cleared_cap * 3 * sizeof(int), cudaMemcpyDeviceToDevice));
      cudaFree(m->dead_cleared_xyz);
    }
    m->dead_cleared_xyz = q;
    m->dead_cleared_cap = m->esdf.capacity;
  }
  DecayArgs a{};
  a.layer = P;
  a.occupancy = occupancy ? 1 : 0;
  a.p.block_size = m->block_size;
  a.p.voxel_size = m->block_size * (1.0f / kVps);
  a.p.half_voxel_size = m->block_size * (0.5f / kVps);
  // DepthObservationSpace of Mapper::decay*ExcludeLastView (mapper_impl.h:190-203,227-243)
  a.p.max_integration_distance_m = m->tp.max_integration_distance_m;
  a.p.truncation_distance_m = m->tp.truncation_distance_vox * m->voxel_size;
  if (occupancy) {
    a.free_log_odds = logOddsFromProbability(m->odp.free_region_decay_probability);
    a.occupied_log_odds = logOddsFromProbability(m->odp.occupied_region_decay_probability);
    a.to_log_odds = logOddsFromProbability(m->odp.decay_to_probability);
    a.deallocate = m->odp.deallocate_decayed_blocks ? 1 : 0;
  } else {
    a.decay_factor = m->tdp.decay_factor;
    a.weight_threshold = m->tdp.decayed_weight_threshold;
    a.set_free_distance = m->tdp.set_free_distance_on_decayed ? 1 : 0;
    a.free_distance_m = m->tdp.free_distance_vox * m->voxel_size;  // tsdf_decay_integrator_impl.cuh:85
    a.deallocate = m->tdp.deallocate_decayed_blocks ? 1 : 0;
  }
  // block exclusion
  std::vector<int> excl;
  int* excl_dev = nullptr;
  if (exclusion && exclusion->num_excluded_blocks > 0) {
    const int ne = exclusion->num_excluded_blocks;
    NVB_CUDA(cudaMalloc(&excl_dev, (size_t)ne * 3 * sizeof(int)));
    NVB_CUDA(cudaMemcpyAsync(excl_dev, exclusion->excluded_blocks_xyz_host, (size_t)ne * 3 * sizeof(int), cudaMemcpyHostToDevice,
                             m->stream));
    m->skip_seq++;
    launchMarkSkipped(P, excl_dev, ne, m->skip_stamp, m->skip_seq, m->stream);
    a.skip_stamp = m->skip_stamp;
    a.skip_seq = m->skip_seq;
  }
  if (exclusion && exclusion->has_exclusion_sphere && exclusion->exclusion_radius_m * exclusion->exclusion_radius_m > 0.0f) {
    a.has_sphere = 1;
    a.cx = exclusion->exclusion_center[0], a.cy = exclusion->exclusion_center[1], a.cz = exclusion->exclusion_center[2];
    a.r2 = exclusion->exclusion_radius_m * exclusion->exclusion_radius_m;
  }
  // view exclusion
  float* depth_tmp = nullptr;
  if (depth) {
    const float* depth_dev = depth;
    if (depth_memory == NVB_MEM_HOST) {
      NVB_CUDA(cudaMalloc(&depth_tmp, (size_t)rows * cols * sizeof(float)));
      NVB_CUDA(cudaMemcpyAsync(depth_tmp, depth, (size_t)rows * cols * sizeof(float), cudaMemcpyHostToDevice, m->stream));
      depth_dev = depth_tmp;
    }
    a.depth = depth_dev, a.rows = rows, a.cols = cols;
    a.T_C_L = invertRigid(rigidFromColMajor(T_L_C));
    a.cam = *cam;
  }
  a.dead = m->dead;
  a.dead_count = m->esdf_ints + kDeadCount;
  a.tracker_dirty = m->dirty;
  NVB_CUDA(cudaMemsetAsync(a.dead_count, 0, sizeof(int), m->stream));
  launchDecay(a, m->num_sms, m->stream);
  m->launches++;
  int n_dead = 0;
  NVB_CUDA(cudaMemcpyAsync(&n_dead, a.dead_count, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
  NVB_CUDA(cudaStreamSynchronize(m->stream));
  if (depth_tmp) cudaFree(depth_tmp);
  if (excl_dev) cudaFree(excl_dev);
  if (n_dead > 0) {
    // Mapper::clearBlocksInLayers: the same blocks leave the ESDF layer; then both hashes are rebuilt without them
    EsdfCtx c = makeEsdfCtx(m);
    if (m->esdf_mode == 2) {
      c.slice_mode = 1;
      // getBlockIndexFromPositionInLayer of the slice heights (src/mapper/mapper.cpp:574-590)
      c.slice_min_bz = (int)std::floor(m->sp.slice_min_height_m / m->block_size);
      c.slice_max_bz = (int)std::floor(m->sp.slice_max_height_m / m->block_size);
      c.slice_out_bz = (int)std::floor(m->sp.slice_height_m / m->block_size);
    }
    launchEsdfRemoveBlocks(c, m->dead, a.dead_count, n_dead, m->stream);
    // Voxels of other blocks may keep parents inside the removed blocks: the reference clears them when they happen to be
    // candidates of a later clear pass, which the per-block parent boxes cannot tell. No pruning from here on.
    m->prune_ok = false;
    std::vector<DevLayer*> touched = {&m->tsdf, &m->esdf};
    if (m->freespace.blocks) {
      launchRemoveBlocks(m->freespace, m->dead, a.dead_count, n_dead, m->stream);
      touched.push_back(&m->freespace);
    }
    if (m->color.blocks) {
      launchRemoveBlocks(m->color, m->dead, a.dead_count, n_dead, m->stream);
      touched.push_back(&m->color);
    }
    // ColorMeshLayer::clearBlocksAsync (Mapper::clearBlocksInLayers, src/mapper/mapper.cpp:552-557): the arena segments
    // of the removed headers are no longer referenced and are dropped by the next arena repack
    if (m->mesh.blocks) {
      launchRemoveBlocks(m->mesh, m->dead, a.dead_count, n_dead, m->stream);
      touched.push_back(&m->mesh);
    }
    for (DevLayer* L : touched) {
      int hw = 0;
      NVB_CUDA(cudaMemcpyAsync(&hw, L->count, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
      NVB_CUDA(cudaStreamSynchronize(m->stream));
      hw = std::min(hw, L->capacity);
      launchFillU64(L->hash.keys, kEmptyKey, (size_t)L->hash.mask + 1, m->stream);
      launchRehash(*L, hw, m->stream);
    }
    m->launches += 6;
    if (m->dirty_fs) NVB_CUDA(cudaMemsetAsync(m->dirty_fs, 0, (size_t)m->tsdf.capacity * sizeof(int), m->stream));
    if (m->dirty_mesh) NVB_CUDA(cudaMemsetAsync(m->dirty_mesh, 0, (size_t)m->tsdf.capacity * sizeof(int), m->stream));
    if (out_count) *out_count = n_dead;
    if (removed_xyz_host && cap > 0) {
      const int k = std::min(n_dead, (int)cap);
      std::vector<int4> tmp((size_t)k);
      NVB_CUDA(cudaMemcpyAsync(tmp.data(), m->dead, (size_t)k * sizeof(int4), cudaMemcpyDeviceToHost, m->stream));
      NVB_CUDA(cudaStreamSynchronize(m->stream));
      for (int i = 0; i < k; i++)
        removed_xyz_host[3 * i] = tmp[i].y, removed_xyz_host[3 * i + 1] = tmp[i].z, removed_xyz_host[3 * i + 2] = tmp[i].w;
    }
  }
  // BlocksToUpdateTracker::addAllBlocksToUpdate (mapper_impl.h:208-211): the next ESDF update covers every block
  m->tracker_initialized = false;
  m->fs_tracker_initialized = false;
  m->mesh_tracker_initialized = false;
  NVB_CUDA(cudaMemsetAsync(m->todo_count, 0, sizeof(int), m->stream));
  NVB_CUDA(cudaMemsetAsync(m->esdf_ints + kTodoFsCount, 0, sizeof(int), m->stream));
  NVB_CUDA(cudaMemsetAsync(m->esdf_ints + kTodoMeshCount, 0, sizeof(int), m->stream));
  NVB_CUDA(syncAll(m));
  return checkDeviceError(m);
}

void nvb_default_freespace_params(NvbFreespaceParams* p) {
  if (!p) return;
  // integrators/freespace_integrator_params.h:22-58
  p->max_tsdf_distance_for_occupancy_m = 0.15f;
  p->max_unobserved_to_keep_consecutive_occupancy_ms = 200;
  p->min_duration_since_occupied_for_freespace_ms = 1000;
  p->min_consecutive_occupancy_duration_for_reset_ms = 2000;
  p->check_neighborhood = 1;
  p->initialize_to_high_confidence_freespace = 0;
}
int32_t nvb_mapper_set_freespace_params(NvbMapper* m, const NvbFreespaceParams* p) {
  if (!m || !p) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  m->fp = *p;  // the reference's setters do not check (src/integrators/freespace_integrator.cu:35-83)
  return NVB_OK;
}
int32_t nvb_mapper_get_freespace_params(const NvbMapper* m, NvbFreespaceParams* p) {
  if (!m || !p) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  *p = m->fp;
  return NVB_OK;
}

namespace {
// FreespaceIntegrator::updateFreespaceLayer on the tracker's list (in_xyz_dev == nullptr) or on an explicit one.
int freespaceUpdateImpl(NvbMapper* m, const int* in_xyz_dev, int n_explicit, long long now_ms, const float* depth, int memory,
                        int rows, int cols, const float* T_L_C, const NvbCamera* cam, float max_view_distance_m,
                        float truncation_distance_m) {
  // the freespace slab follows the TSDF slab's capacity
  if (m->freespace.capacity < m->tsdf.capacity) {
    int rc = growLayer(m, &m->freespace, m->tsdf.capacity);
    if (rc) return rc;
  }
  int upper = in_xyz_dev ? n_explicit : std::min(m->tsdf_count_ub, m->tsdf.capacity);
  if (!in_xyz_dev) pollCounts(m), upper = std::min(m->tsdf_count_ub, m->tsdf.capacity);
  if (upper < 1) upper = 1;
  if (m->fs_work_cap < upper) {
    NVB_CUDA(cudaStreamSynchronize(m->stream));
    if (m->fs_work) cudaFree(m->fs_work);
    NVB_CUDA(cudaMalloc(&m->fs_work, (size_t)upper * 2 * sizeof(int4)));
    m->fs_work_cap = upper * 2;
  }
  FreespaceArgs a{};
  a.tsdf = m->tsdf, a.fs = m->freespace;
  if (in_xyz_dev) {
    a.in_xyz = in_xyz_dev, a.n_explicit = n_explicit;
  } else {
    a.todo_slots = m->todo_fs_slots, a.todo_count = m->esdf_ints + kTodoFsCount, a.tracker_dirty = m->dirty_fs;
  }
  a.work = m->fs_work, a.work_count = m->esdf_ints + kFsWorkCount, a.error = m->error_dev;
  a.max_tsdf_distance_for_occupancy_m = m->fp.max_tsdf_distance_for_occupancy_m;
  a.max_unobserved_ms = m->fp.max_unobserved_to_keep_consecutive_occupancy_ms;
  a.min_free_ms = m->fp.min_duration_since_occupied_for_freespace_ms;
  a.min_reset_ms = m->fp.min_consecutive_occupancy_duration_for_reset_ms;
  a.check_neighborhood = m->fp.check_neighborhood ? 1 : 0;
  a.init_high_confidence = m->fp.initialize_to_high_confidence_freespace ? 1 : 0;
  a.last_update_ms = m->fs_last_update_ms, a.now_ms = now_ms;
  a.p.block_size = m->block_size;
  a.p.voxel_size = m->block_size * (1.0f / kVps);
  a.p.half_voxel_size = m->block_size * (0.5f / kVps);
  a.p.max_integration_distance_m = max_view_distance_m > 0.0f ? max_view_distance_m : FLT_MAX;
  a.p.truncation_distance_m = truncation_distance_m > 0.0f ? truncation_distance_m : FLT_MAX;
  float* depth_tmp = nullptr;
  if (depth) {
    const float* depth_dev = depth;
    if (memory == NVB_MEM_HOST) {
      NVB_CUDA(cudaMalloc(&depth_tmp, (size_t)rows * cols * sizeof(float)));
      NVB_CUDA(cudaMemcpyAsync(depth_tmp, depth, (size_t)rows * cols * sizeof(float), cudaMemcpyHostToDevice, m->stream));
      depth_dev = depth_tmp;
    }
    a.depth = depth_dev, a.rows = rows, a.cols = cols;
    a.T_C_L = invertRigid(rigidFromColMajor(T_L_C));
    a.cam = *cam;
  }
  launchFreespaceUpdate(a, upper, m->num_sms, m->stream);
  m->launches += 2;
  m->fs_last_update_ms = now_ms;
  NVB_CUDA(cudaStreamSynchronize(m->stream));
  if (depth_tmp) cudaFree(depth_tmp);
  return checkDeviceError(m);
}
}  // namespace

int32_t nvb_mapper_update_freespace(NvbMapper* m, int64_t update_time_ms, const float* depth, int32_t depth_memory,
                                    int32_t rows, int32_t cols, const float* T_L_C, const NvbCamera* cam,
                                    int32_t update_full_layer) {
  if (!m) return fail(NVB_ERR_INVALID_ARGUMENT, "null mapper");
  if (m->projective_layer_type != NVB_PROJECTIVE_TSDF_WITH_FREESPACE)
    return fail(NVB_ERR_INVALID_ARGUMENT, "the mapper has no freespace layer");  // CHECK(hasFreespaceLayer(...)), mapper_impl.h:180-181
  if (depth) {
    int rc = validateFrameArgs(m, depth, rows, cols, T_L_C, cam);
    if (rc) return rc;
  }
  NVB_CUDA(cudaSetDevice(m->device));
  if (!m->fs_tracker_initialized || update_full_layer) {
    launchTodoAll(m->tsdf, m->dirty_fs, m->todo_fs_slots, m->esdf_ints + kTodoFsCount, m->stream);
    m->launches++;
    m->fs_tracker_initialized = true;
  }
  // kTruncationDistanceMultiplier = 2 (mapper_impl.h:157-172)
  return freespaceUpdateImpl(m, nullptr, 0, update_time_ms, depth, depth_memory, rows, cols, T_L_C, cam,
                             m->tp.max_integration_distance_m, 2.0f * (m->tp.truncation_distance_vox * m->voxel_size));
}

int32_t nvb_freespace_update_blocks(NvbMapper* m, const int32_t* blocks_xyz_host, int32_t num_blocks, int64_t update_time_ms,
                                    const float* depth, int32_t depth_memory, int32_t rows, int32_t cols, const float* T_L_C,
                                    const NvbCamera* cam, float max_view_distance_m, float truncation_distance_m) {
  if (!m) return fail(NVB_ERR_INVALID_ARGUMENT, "null mapper");
  if (m->projective_layer_type != NVB_PROJECTIVE_TSDF_WITH_FREESPACE)
    return fail(NVB_ERR_INVALID_ARGUMENT, "the mapper has no freespace layer");
  if (num_blocks < 0 || (num_blocks > 0 && !blocks_xyz_host)) return fail(NVB_ERR_INVALID_ARGUMENT, "bad block list");
  if (num_blocks == 0) return NVB_OK;  // early return (:337-339)
  if (depth) {
    int rc = validateFrameArgs(m, depth, rows, cols, T_L_C, cam);
    if (rc) return rc;
  }
  NVB_CUDA(cudaSetDevice(m->device));
  for (int i = 0; i < num_blocks; i++)
    if (!indexInRange(blocks_xyz_host[3 * i], blocks_xyz_host[3 * i + 1], blocks_xyz_host[3 * i + 2]))
      return fail(NVB_ERR_INDEX_RANGE, "block index outside +-2^20");
  // a set, like every caller's list in the reference (find-or-insert needs unique keys per launch)
  struct K3 {
    int x, y, z;
  };
  std::vector<K3> v((size_t)num_blocks);
  memcpy(v.data(), blocks_xyz_host, (size_t)num_blocks * sizeof(K3));
  std::sort(v.begin(), v.end(), [](const K3& a, const K3& b) { return a.x != b.x ? a.x < b.x : (a.y != b.y ? a.y < b.y : a.z < b.z); });
  v.erase(std::unique(v.begin(), v.end(), [](const K3& a, const K3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }), v.end());
  num_blocks = (int)v.size();
  int* xyz_dev = nullptr;
  NVB_CUDA(cudaMalloc(&xyz_dev, (size_t)num_blocks * 3 * sizeof(int)));
  NVB_CUDA(cudaMemcpyAsync(xyz_dev, v.data(), (size_t)num_blocks * 3 * sizeof(int), cudaMemcpyHostToDevice, m->stream));
  NVB_CUDA(cudaStreamSynchronize(m->stream));
  const int rc = freespaceUpdateImpl(m, xyz_dev, num_blocks, update_time_ms, depth, depth_memory, rows, cols, T_L_C, cam,
                                     max_view_distance_m, truncation_distance_m);
  cudaFree(xyz_dev);
  return rc;
}

int32_t nvb_mapper_decay_exclude_last_view(NvbMapper* m, const NvbDecayExclusion* exclusion, int32_t* removed_xyz_host,
                                           int32_t cap, int32_t* out_count) {
  if (!m) return fail(NVB_ERR_INVALID_ARGUMENT, "null mapper");
  if (!m->keep_last_view) return fail(NVB_ERR_INVALID_ARGUMENT, "the mapper was created without keep_last_view");
  if (!m->has_last_view)  // "Last view not set for sensor type. Decaying all voxels" (mapper_impl.h:200-203)
    return nvb_mapper_decay(m, exclusion, nullptr, 0, 0, 0, nullptr, nullptr, removed_xyz_host, cap, out_count);
  return nvb_mapper_decay(m, exclusion, m->last_depth, NVB_MEM_DEVICE, m->last_rows, m->last_cols, m->last_T_L_C,
                          &m->last_cam, removed_xyz_host, cap, out_count);
}

float nvb_mapper_voxel_size(const NvbMapper* m) { return m ? m->voxel_size : 0.0f; }
float nvb_mapper_block_size(const NvbMapper* m) { return m ? m->block_size : 0.0f; }

int32_t nvb_view_raycast(NvbMapper* m, const float* depth, int32_t depth_memory, int32_t rows, int32_t cols,
                         const float* T_L_C, const NvbCamera* cam, float block_size,
                         float max_integration_distance_behind_surface_m, float max_integration_distance_m,
                         int32_t* out_xyz_host, int32_t cap, int32_t* out_count) {
  int rc = validateFrameArgs(m, depth, rows, cols, T_L_C, cam);
  if (rc) return rc;
  if (!(block_size > 0.0f)) return fail(NVB_ERR_INVALID_ARGUMENT, "block_size must be > 0");
  NVB_CUDA(cudaSetDevice(m->device));
  if ((rc = enqueueFrame(m, depth, nullptr, 0, depth_memory, rows, cols, T_L_C, cam, block_size,
                         max_integration_distance_behind_surface_m, max_integration_distance_m, false)))
    return rc;
  return readFrameList(m, out_xyz_host, cap, out_count);
}

int32_t nvb_mapper_integrate_depth_async(NvbMapper* m, const float* depth, const uint8_t* mask, int32_t mask_mode,
                                         int32_t memory, int32_t rows, int32_t cols, const float* T_L_C,
                                         const NvbCamera* cam) {
  int rc = validateFrameArgs(m, depth, rows, cols, T_L_C, cam);
  if (rc) return rc;
  NVB_CUDA(cudaSetDevice(m->device));
  // max_integration_distance_behind_surface_m = truncation_distance_vox * voxel_size
  // (projective_integrator_impl.cuh:234-235)
  if (m->projective_layer_type == NVB_PROJECTIVE_OCCUPANCY &&
      m->tp.truncation_distance_vox * m->voxel_size < m->op.occupied_region_half_width_m) {
    // "truncation distance must be >= occupied region half width": the integrator raises it through its own
    // setter, so the new value persists (src/integrators/projective_occupancy_integrator.cu:51-64)
    m->tp.truncation_distance_vox = m->op.occupied_region_half_width_m / m->voxel_size;
  }
  const float trunc_m = m->tp.truncation_distance_vox * m->voxel_size;
  return enqueueFrame(m, depth, mask, mask_mode, memory, rows, cols, T_L_C, cam, m->block_size, trunc_m,
                      m->tp.max_integration_distance_m, true);
}

int32_t nvb_mapper_mark_unobserved_free_inside_radius(NvbMapper* m, const float center[3], float radius, int32_t* updated_xyz_host,
                                                      int32_t cap, int32_t* out_count) {
  if (!m || !center) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  if (!(radius > 0.0f)) return fail(NVB_ERR_INVALID_ARGUMENT, "radius must be positive");  // CHECK_GT(radius, 0.0f)
  if (out_count) *out_count = 0;
  NVB_CUDA(cudaSetDevice(m->device));
  MarkFreeArgs a{};
  const Vec3 mn{center[0] - radius, center[1] - radius, center[2] - radius};
  const Vec3 mx{center[0] + radius, center[1] + radius, center[2] + radius};
  a.lo = blockIndexFromPosition(m->block_size, mn);
  const int3 hi = blockIndexFromPosition(m->block_size, mx);
  a.size = make_int3(hi.x - a.lo.x + 1, hi.y - a.lo.y + 1, hi.z - a.lo.z + 1);
  const long long cells = (long long)a.size.x * a.size.y * a.size.z;
  if (cells <= 0 || cells > (1ll << 26)) return fail(NVB_ERR_CAPACITY, "the sphere covers more than 2^26 blocks");
  if (!indexInRange(a.lo.x, a.lo.y, a.lo.z) || !indexInRange(hi.x, hi.y, hi.z))
    return fail(NVB_ERR_INDEX_RANGE, "block index outside +-2^20");
  a.cells = (int)cells;
  int rc;
  if ((rc = ensureTsdfCapacity(m, cells))) return rc;
  m->cells_cum += cells;
  m->tsdf_count_ub += (int)cells;
  NVB_CUDA(syncAll(m));  // rare, synchronous call: the ESDF side stream may still be reading the projective layer
  a.layer = m->tsdf;
  a.occupancy = m->projective_layer_type == NVB_PROJECTIVE_OCCUPANCY ? 1 : 0;
  a.cx = center[0], a.cy = center[1], a.cz = center[2];
  a.radius = radius;
  a.block_size = m->block_size;
  a.trunc_m = m->tp.truncation_distance_vox * m->voxel_size;  // get_truncation_distance_m(layer->voxel_size())
  a.error = m->error_dev;
  if (m->tracker_initialized) a.dirty = m->dirty, a.todo_slots = m->todo_slots, a.todo_count = m->todo_count;
  if (m->dirty_fs && m->fs_tracker_initialized)
    a.dirty2 = m->dirty_fs, a.todo2_slots = m->todo_fs_slots, a.todo2_count = m->esdf_ints + kTodoFsCount;
  if (m->dirty_mesh && m->mesh_tracker_initialized)
    a.dirty3 = m->dirty_mesh, a.todo3_slots = m->todo_mesh_slots, a.todo3_count = m->esdf_ints + kTodoMeshCount;
  int4* out_dev = nullptr;
  NVB_CUDA(cudaMalloc(&out_dev, ((size_t)cells + 1) * sizeof(int4)));
  a.out = out_dev + 1;
  a.out_count = reinterpret_cast<int*>(out_dev);
  NVB_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(int4), m->stream));
  launchMarkFreeSphere(a, m->num_sms, m->stream);
  m->launches++;
  int n = 0;
  NVB_CUDA(cudaMemcpyAsync(&n, a.out_count, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
  NVB_CUDA(cudaStreamSynchronize(m->stream));
  if (out_count) *out_count = n;
  if (updated_xyz_host && cap > 0 && n > 0) {
    const int k = std::min(n, (int)cap);
    std::vector<int4> tmp((size_t)k);
    NVB_CUDA(cudaMemcpyAsync(tmp.data(), a.out, (size_t)k * sizeof(int4), cudaMemcpyDeviceToHost, m->stream));
    NVB_CUDA(cudaStreamSynchronize(m->stream));
    for (int i = 0; i < k; i++)
      updated_xyz_host[3 * i] = tmp[i].x, updated_xyz_host[3 * i + 1] = tmp[i].y, updated_xyz_host[3 * i + 2] = tmp[i].z;
  }
  cudaFree(out_dev);
  return checkDeviceError(m);
}

// ---------------------------------------------------------------------------
// Colour integration (nvb_color.cu)
// ---------------------------------------------------------------------------
void nvb_default_color_params(NvbColorParams* p) {
  if (!p) return;
  memset(p, 0, sizeof(*p));
  // integrators/projective_integrator_params.h:24-75, projective_appearance_integrator.h:164, rays/sphere_tracer.h:216-218
  p->max_integration_distance_m = 7.0f;
  p->truncation_distance_vox = 4.0f;
  p->max_weight = 5.0f;
  p->measurement_weight = 0.8f;
  p->sphere_tracing_ray_subsampling_factor = 4;
  p->sphere_tracer_maximum_steps = 100;
  p->sphere_tracer_maximum_ray_length_m = 7.0f;
  p->sphere_tracer_surface_distance_epsilon_vox = 0.1f;
  p->workspace_bounds_type = NVB_WS_UNBOUNDED;
}
int32_t nvb_mapper_set_color_params(NvbMapper* m, const NvbColorParams* p) {
  if (!m || !p) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  // the reference's setters CHECK these (projective_appearance_integrator.cu:168-208, projective_integrator.cpp setters)
  if (!(p->max_integration_distance_m > 0.0f) || !(p->truncation_distance_vox > 0.0f) || !(p->max_weight > 0.0f) ||
      !(p->measurement_weight > 0.0f) || !(p->measurement_weight <= 1.0f) || p->sphere_tracing_ray_subsampling_factor <= 0 ||
      p->sphere_tracer_maximum_steps <= 0 || !(p->sphere_tracer_maximum_ray_length_m > 0.0f) ||
      !(p->sphere_tracer_surface_distance_epsilon_vox > 0.0f))
    return fail(NVB_ERR_INVALID_ARGUMENT, "colour integrator parameter out of range");
  m->cp = *p;
  return NVB_OK;
}
int32_t nvb_mapper_get_color_params(const NvbMapper* m, NvbColorParams* p) {
  if (!m || !p) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  *p = m->cp;
  return NVB_OK;
}

namespace {
// binary32 -> binary16 -> binary32, round to nearest even (__float2half followed by the implicit __half -> float)
float roundThroughHalf(float f) {
  uint32_t x;
  memcpy(&x, &f, 4);
  const uint32_t sign = x & 0x80000000u, ax = x & 0x7fffffffu;
  uint32_t out;
  if (ax >= 0x7f800000u) {
    out = ax;
  } else if (ax >= 0x477ff000u) {
    out = 0x7f800000u;
  } else if (ax < 0x38800000u) {
    const float r = std::nearbyint(std::fabs(f) * 16777216.0f) * (1.0f / 16777216.0f);
    memcpy(&out, &r, 4);
  } else {
    out = (ax + 0x0fffu + ((ax >> 13) & 1u)) & 0xffffe000u;
  }
  out |= sign;
  float r;
  memcpy(&r, &out, 4);
  return r;
}

int ensureColorLayer(NvbMapper* m) {
  int rc;
  if (!m->color.blocks) {
    if ((rc = allocLayer(&m->color, m->tsdf.capacity, kColorBlockBytes, m->stream))) return rc;
  } else if (m->color.capacity < m->tsdf.capacity) {
    if ((rc = growLayer(m, &m->color, m->tsdf.capacity))) return rc;
  }
  if (m->color_work_cap < m->tsdf.capacity) {
    NVB_CUDA(syncAll(m));
    if (m->color_work) cudaFree(m->color_work);
    NVB_CUDA(cudaMalloc(&m->color_work, (size_t)m->tsdf.capacity * sizeof(int4)));
    m->color_work_cap = m->tsdf.capacity;
  }
  return NVB_OK;
}

// The tracer's half of ColorArgs + the launch. synth must hold (height / f) * (width / f) floats.
int fillTracerArgs(NvbMapper* m, ColorArgs* a, const float* T_L_C_cm, const NvbCamera* cam, float trunc_m, int f) {
  if (f <= 0 || cam->width % f != 0 || cam->height % f != 0)
    return fail(NVB_ERR_INVALID_ARGUMENT, "the ray subsampling factor must divide the image size");  // CHECK_EQ, sphere_tracer.cu:432-433
  a->tsdf = m->tsdf;
  a->T_L_C = rigidFromColMajor(T_L_C_cm);
  a->T_C_L = invertRigid(a->T_L_C);
  a->cam = *cam;
  a->block_size = m->block_size;
  a->voxel_size = m->block_size * (1.0f / kVps);
  a->half_voxel_size = m->block_size * (0.5f / kVps);
  a->voxel_size_inv = (float)(1.0 / (double)(m->block_size * (1.0f / kVps)));  // indexing_impl.h:41
  a->trunc_m = trunc_m;
  a->subsample = f;
  a->drows = cam->height / f, a->dcols = cam->width / f;  // getSubsampledImageSize (sphere_tracer.cu:335-339)
  a->max_steps = m->cp.sphere_tracer_maximum_steps;
  a->max_ray_len = m->cp.sphere_tracer_maximum_ray_length_m;
  a->eps_m = m->cp.sphere_tracer_surface_distance_epsilon_vox * m->voxel_size;
  const size_t need = (size_t)a->drows * a->dcols;
  if (m->color_synth_cap < need) {
    NVB_CUDA(syncAll(m));
    if (m->color_synth) cudaFree(m->color_synth);
    NVB_CUDA(cudaMalloc(&m->color_synth, need * sizeof(float)));
    m->color_synth_cap = need;
  }
  a->synth = m->color_synth;
  return NVB_OK;
}

int stageBytes(NvbMapper* m, unsigned char** buf, size_t* cap, const unsigned char* host, size_t bytes) {
  if (*cap < bytes) {
    NVB_CUDA(syncAll(m));
    if (*buf) cudaFree(*buf);
    NVB_CUDA(cudaMalloc(buf, bytes));
    *cap = bytes;
  }
  NVB_CUDA(cudaMemcpyAsync(*buf, host, bytes, cudaMemcpyHostToDevice, m->stream));
  return NVB_OK;
}
}  // namespace

int32_t nvb_sphere_tracer_render_depth(NvbMapper* m, const float* T_L_C, const NvbCamera* cam, float truncation_distance_m,
                                       int32_t ray_subsampling_factor, float* out_depth_host) {
  if (!m || !T_L_C || !cam || !out_depth_host) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  if (m->projective_layer_type == NVB_PROJECTIVE_OCCUPANCY)
    return fail(NVB_ERR_INVALID_ARGUMENT, "the sphere tracer needs a TSDF layer");
  NVB_CUDA(cudaSetDevice(m->device));
  ColorArgs a{};
  int rc = fillTracerArgs(m, &a, T_L_C, cam, truncation_distance_m, ray_subsampling_factor);
  if (rc) return rc;
  launchSphereTrace(a, m->stream);
  m->launches++;
  NVB_CUDA(cudaMemcpyAsync(out_depth_host, a.synth, (size_t)a.drows * a.dcols * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
  NVB_CUDA(cudaStreamSynchronize(m->stream));
  return checkDeviceError(m);
}

int32_t nvb_mapper_integrate_color(NvbMapper* m, const uint8_t* color, const uint8_t* mask, int32_t mask_mode, int32_t memory,
                                   int32_t rows, int32_t cols, const float* T_L_C, const NvbCamera* cam,
                                   int32_t* updated_xyz_host, int32_t cap, int32_t* out_count) {
  if (!m || !color || !T_L_C || !cam) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  if (rows <= 0 || cols <= 0) return fail(NVB_ERR_INVALID_ARGUMENT, "image must have positive size");
  if (!(cam->fu != 0.0f) || !(cam->fv != 0.0f)) return fail(NVB_ERR_INVALID_ARGUMENT, "camera focal length is zero");
  if (out_count) *out_count = 0;
  // "Color is only integrated for Tsdf layers (not for occupancy)" (mapper_impl.h:118-119)
  if (m->projective_layer_type == NVB_PROJECTIVE_OCCUPANCY) return NVB_OK;
  NVB_CUDA(cudaSetDevice(m->device));
  int rc;
  if ((rc = ensureColorLayer(m))) return rc;
  const float trunc_m = m->cp.truncation_distance_vox * m->voxel_size;
  ColorArgs a{};
  if ((rc = fillTracerArgs(m, &a, T_L_C, cam, trunc_m, m->cp.sphere_tracing_ray_subsampling_factor))) return rc;
  a.color = m->color;
  // Camera::getViewAABB(T_L_C, 1e-6, max_integration_distance + truncation) + the integrator's own workspace bounds
  // (view_calculator_impl.h:47-58)
  {
    const float max_distance = m->cp.max_integration_distance_m + trunc_m;
    const float w = (float)cam->width, h = (float)cam->height;
    const float ux[4] = {0.0f, w, w, 0.0f}, vy[4] = {0.0f, 0.0f, h, h};
    Vec3 ray[4];
    for (int k = 0; k < 4; k++) {
      float nx = (ux[k] - cam->cu) / cam->fu, ny = (vy[k] - cam->cv) / cam->fv;
      if (cam->has_distortion) removeDistortion(*cam, nx, ny);
      ray[k] = Vec3{nx, ny, 1.0f};
    }
    const int order[4] = {2, 1, 0, 3};
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int k = 0; k < 8; k++) {
      const float d = (k < 4) ? 1e-6f : max_distance;
      const Vec3 r = ray[order[k & 3]];
      const Vec3 c = transformPoint(a.T_L_C, Vec3{d * r.x, d * r.y, d * r.z});
      const float cl[3] = {c.x, c.y, c.z};
      for (int i = 0; i < 3; i++) lo[i] = std::min(lo[i], cl[i]), hi[i] = std::max(hi[i], cl[i]);
    }
    if (m->cp.workspace_bounds_type == NVB_WS_HEIGHT_BOUNDS) {
      lo[2] = std::max(lo[2], m->cp.workspace_min[2]);
      hi[2] = std::min(hi[2], m->cp.workspace_max[2]);
    } else if (m->cp.workspace_bounds_type == NVB_WS_BOUNDING_BOX) {
      for (int i = 0; i < 3; i++) lo[i] = std::max(m->cp.workspace_min[i], lo[i]), hi[i] = std::min(hi[i], m->cp.workspace_max[i]);
    }
    if (lo[0] > hi[0] || lo[1] > hi[1] || lo[2] > hi[2]) return NVB_OK;  // empty workspace intersection: nothing in view
    a.aabb_lo = blockIndexFromPosition(m->block_size, Vec3{lo[0], lo[1], lo[2]});
    a.aabb_hi = blockIndexFromPosition(m->block_size, Vec3{hi[0], hi[1], hi[2]});
    // Camera::getNormalizedViewport(getViewportMargin(height)) (src/sensors/camera.cpp:85-96, view_calculator_impl.h:81-83)
    const float margin = (float)cam->height / 20.0f;
    float x0 = (-margin - cam->cu) / cam->fu, y0 = (-margin - cam->cv) / cam->fv;
    float x1 = (((float)cam->width + margin) - cam->cu) / cam->fu, y1 = (((float)cam->height + margin) - cam->cv) / cam->fv;
    if (cam->has_distortion) removeDistortion(*cam, x0, y0), removeDistortion(*cam, x1, y1);
    a.vmin_x = x0, a.vmin_y = y0, a.vmax_x = x1, a.vmax_y = y1;
  }
  a.max_integration_distance_m = m->cp.max_integration_distance_m;
  a.max_weight = m->cp.max_weight;
  a.measurement_weight = m->cp.measurement_weight;
  {  // blendTwoArrays (projective_appearance_integrator.cu:287-306)
    float w_old = 1.0f - m->cp.measurement_weight, w_new = m->cp.measurement_weight;
    const float total = w_old + w_new;
    w_old /= total, w_new /= total;
    a.w_old_h = roundThroughHalf(w_old), a.w_new_h = roundThroughHalf(w_new);
  }
  a.work = m->color_work;
  a.work_count = m->esdf_ints + kColorWorkCount;
  a.error = m->error_dev;
  a.rows = rows, a.cols = cols;
  a.depth_subsample = rows / a.drows;  // projective_integrator_impl.cuh:320
  if (a.depth_subsample <= 0) return fail(NVB_ERR_INVALID_ARGUMENT, "the colour image is smaller than the synthetic depth image");
  a.mask_mode = mask_mode;
  if (memory == NVB_MEM_HOST) {
    if ((rc = stageBytes(m, &m->color_stage, &m->color_stage_cap, color, (size_t)rows * cols * 3))) return rc;
    a.color_image = m->color_stage;
    a.mask = nullptr;
    if (mask) {
      if ((rc = stageBytes(m, &m->color_mask_stage, &m->color_mask_stage_cap, mask, (size_t)rows * cols))) return rc;
      a.mask = m->color_mask_stage;
    }
  } else {
    a.color_image = color, a.mask = mask;
  }
  NVB_CUDA(cudaMemsetAsync(a.work_count, 0, sizeof(int), m->stream));
  launchColorSelect(a, m->num_sms, m->stream);
  launchSphereTrace(a, m->stream);
  launchColorIntegrate(a, m->num_sms, m->stream);
  m->launches += 3;
  // Device-resident frames with no output requested stay asynchronous (read the list later with
  // nvb_mapper_last_color_blocks); host buffers must be released and outputs filled, so those calls synchronise.
  if (memory == NVB_MEM_DEVICE && !updated_xyz_host && !out_count) return NVB_OK;
  NVB_CUDA(cudaStreamSynchronize(m->stream));
  int rc2 = nvb_mapper_last_color_blocks(m, updated_xyz_host, cap, out_count);
  if (rc2) return rc2;
  return checkDeviceError(m);
}

int32_t nvb_mapper_last_color_blocks(NvbMapper* m, int32_t* out_xyz_host, int32_t cap, int32_t* out_count) {
  if (!m) return fail(NVB_ERR_INVALID_ARGUMENT, "null mapper");
  if (out_count) *out_count = 0;
  if (!m->color.blocks || !m->color_work) return NVB_OK;
  NVB_CUDA(cudaSetDevice(m->device));
  int n = 0;
  NVB_CUDA(cudaMemcpyAsync(&n, m->esdf_ints + kColorWorkCount, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
  NVB_CUDA(cudaStreamSynchronize(m->stream));
  if (out_count) *out_count = n;
  if (out_xyz_host && cap > 0 && n > 0) {
    const int k = std::min(n, (int)cap);
    std::vector<int4> tmp((size_t)k);
    NVB_CUDA(cudaMemcpyAsync(tmp.data(), m->color_work, (size_t)k * sizeof(int4), cudaMemcpyDeviceToHost, m->stream));
    NVB_CUDA(cudaStreamSynchronize(m->stream));
    for (int i = 0; i < k; i++)
      out_xyz_host[3 * i] = tmp[i].x, out_xyz_host[3 * i + 1] = tmp[i].y, out_xyz_host[3 * i + 2] = tmp[i].z;
  }
  return NVB_OK;
}

int32_t nvb_mapper_integrate_depth(NvbMapper* m, const float* depth, const uint8_t* mask, int32_t mask_mode,
                                   int32_t memory, int32_t rows, int32_t cols, const float* T_L_C,
                                   const NvbCamera* cam, int32_t* updated_xyz_host, int32_t cap, int32_t* out_count) {
  int rc = nvb_mapper_integrate_depth_async(m, depth, mask, mask_mode, memory, rows, cols, T_L_C, cam);
  if (rc) return rc;
  if ((rc = readFrameList(m, updated_xyz_host, cap, out_count))) return rc;
  return checkPrefetchedError(m);
}

int32_t nvb_mapper_update_esdf_async(NvbMapper* m, int32_t update_full_layer) {
  if (!m) return fail(NVB_ERR_INVALID_ARGUMENT, "null mapper");
  NVB_CUDA(cudaSetDevice(m->device));
  if (!m->tracker_initialized || update_full_layer) {
    // First query of the tracker, or UpdateFullLayer::kYes: every TSDF block
    // (map/blocks_to_update_tracker.cpp:107-124, src/mapper/mapper.cpp:523-537).
    launchTodoAll(m->tsdf, m->dirty, m->todo_slots, m->todo_count, m->stream);
    m->launches++;
    m->tracker_initialized = true;
  }
  return enqueueEsdf(m, nullptr, 0, true);
}

int32_t nvb_mapper_update_esdf(NvbMapper* m, int32_t update_full_layer) {
  int rc = nvb_mapper_update_esdf_async(m, update_full_layer);
  if (rc) return rc;
  return nvb_mapper_synchronize(m);
}

int32_t nvb_esdf_integrate_blocks(NvbMapper* m, const int32_t* blocks_xyz_host, int32_t num_blocks) {
  if (!m) return fail(NVB_ERR_INVALID_ARGUMENT, "null mapper");
  if (num_blocks < 0 || (num_blocks > 0 && !blocks_xyz_host)) return fail(NVB_ERR_INVALID_ARGUMENT, "bad block list");
  if (num_blocks == 0) return NVB_OK;  // early return, esdf_integrator.cu:226-228
  NVB_CUDA(cudaSetDevice(m->device));
  // The list is a set for every caller of the reference (Mapper::getBlocksToUpdate); make it one.
  struct K {
    int x, y, z;
  };
  std::vector<K> v((size_t)num_blocks);
  memcpy(v.data(), blocks_xyz_host, (size_t)num_blocks * sizeof(K));
  std::sort(v.begin(), v.end(), [](const K& a, const K& b) {
    if (a.x != b.x) return a.x < b.x;
    if (a.y != b.y) return a.y < b.y;
    return a.z < b.z;
  });
  v.erase(std::unique(v.begin(), v.end(), [](const K& a, const K& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }),
          v.end());
  for (const K& k : v)
    if (!indexInRange(k.x, k.y, k.z)) return fail(NVB_ERR_INDEX_RANGE, "block index outside +-2^20");
  const int n = (int)v.size();
  if (n > m->xyz_upload_cap) {
    NVB_CUDA(syncAll(m));
    if (m->xyz_upload) cudaFree(m->xyz_upload);
    NVB_CUDA(cudaMalloc(&m->xyz_upload, (size_t)n * 2 * 3 * sizeof(int)));
    m->xyz_upload_cap = n * 2;
  }
  // Stream-ordered upload: a blocking cudaMemcpy from pageable memory may return before the DMA has landed,
  // and the mapper's stream is non-blocking (not ordered against the legacy default stream).
  NVB_CUDA(cudaMemcpyAsync(m->xyz_upload, v.data(), (size_t)n * 3 * sizeof(int), cudaMemcpyHostToDevice, m->stream));
  NVB_CUDA(cudaStreamSynchronize(m->stream));  // v is pageable and about to go out of scope
  int rc = enqueueEsdf(m, m->xyz_upload, n, false);
  if (rc) return rc;
  if ((rc = nvb_mapper_synchronize(m))) return rc;
  return tightenEsdfBound(m);
}

void nvb_default_esdf_slice_params(NvbEsdfSliceParams* p) {
  if (!p) return;
  // integrators/esdf_integrator_params.h:33-43
  p->slice_min_height_m = 0.0f;
  p->slice_max_height_m = 1.0f;
  p->slice_height_m = 1.0f;
  p->slice_height_above_plane_m = 0.0f;
  p->slice_height_thickness_m = 0.1f;
}
int32_t nvb_mapper_set_esdf_slice_params(NvbMapper* m, const NvbEsdfSliceParams* p) {
  if (!m || !p) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  if (!(p->slice_max_height_m >= p->slice_min_height_m)) return fail(NVB_ERR_INVALID_ARGUMENT, "slice_max_height below slice_min_height");
  // CHECK_GE(slice_height_above_plane_m, 0) / CHECK_GT(slice_height_thickness_m, 0) (esdf_integrator.cu:928-929)
  if (!(p->slice_height_above_plane_m >= 0.0f) || !(p->slice_height_thickness_m > 0.0f))
    return fail(NVB_ERR_INVALID_ARGUMENT, "planar slice: height above plane must be >= 0 and thickness > 0");
  m->sp = *p;
  return NVB_OK;
}
int32_t nvb_mapper_get_esdf_slice_params(const NvbMapper* m, NvbEsdfSliceParams* p) {
  if (!m || !p) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  *p = m->sp;
  return NVB_OK;
}

static int32_t updateEsdfSliceImpl(NvbMapper* m, const float* plane, int32_t update_full_layer);
int32_t nvb_mapper_update_esdf_slice(NvbMapper* m, int32_t update_full_layer) { return updateEsdfSliceImpl(m, nullptr, update_full_layer); }
int32_t nvb_mapper_update_esdf_slice_planar(NvbMapper* m, const float plane[4], int32_t update_full_layer) {
  if (!plane) return fail(NVB_ERR_INVALID_ARGUMENT, "null plane");
  return updateEsdfSliceImpl(m, plane, update_full_layer);
}
static int32_t updateEsdfSliceImpl(NvbMapper* m, const float* plane, int32_t update_full_layer) {
  if (!m) return fail(NVB_ERR_INVALID_ARGUMENT, "null mapper");
  NVB_CUDA(cudaSetDevice(m->device));
  if (!m->tracker_initialized || update_full_layer) {
    launchTodoAll(m->tsdf, m->dirty, m->todo_slots, m->todo_count, m->stream);
    m->launches++;
    m->tracker_initialized = true;
  }
  int rc = enqueueEsdf(m, nullptr, 0, true, true, plane);
  if (rc) return rc;
  return nvb_mapper_synchronize(m);
}

static int32_t integrateSliceBlocksImpl(NvbMapper* m, const float* plane, const int32_t* blocks_xyz_host, int32_t num_blocks);
int32_t nvb_esdf_integrate_slice_blocks(NvbMapper* m, const int32_t* blocks_xyz_host, int32_t num_blocks) {
  return integrateSliceBlocksImpl(m, nullptr, blocks_xyz_host, num_blocks);
}
int32_t nvb_esdf_integrate_slice_planar_blocks(NvbMapper* m, const float plane[4], const int32_t* blocks_xyz_host,
                                               int32_t num_blocks) {
  if (!plane) return fail(NVB_ERR_INVALID_ARGUMENT, "null plane");
  return integrateSliceBlocksImpl(m, plane, blocks_xyz_host, num_blocks);
}
static int32_t integrateSliceBlocksImpl(NvbMapper* m, const float* plane, const int32_t* blocks_xyz_host, int32_t num_blocks) {
  if (!m) return fail(NVB_ERR_INVALID_ARGUMENT, "null mapper");
  if (num_blocks < 0 || (num_blocks > 0 && !blocks_xyz_host)) return fail(NVB_ERR_INVALID_ARGUMENT, "bad block list");
  if (num_blocks == 0) return NVB_OK;  // early return (:289-291)
  NVB_CUDA(cudaSetDevice(m->device));
  for (int i = 0; i < num_blocks; i++)
    if (!indexInRange(blocks_xyz_host[3 * i], blocks_xyz_host[3 * i + 1], blocks_xyz_host[3 * i + 2]))
      return fail(NVB_ERR_INDEX_RANGE, "block index outside +-2^20");
  if (num_blocks > m->xyz_upload_cap) {
    NVB_CUDA(syncAll(m));
    if (m->xyz_upload) cudaFree(m->xyz_upload);
    NVB_CUDA(cudaMalloc(&m->xyz_upload, (size_t)num_blocks * 2 * 3 * sizeof(int)));
    m->xyz_upload_cap = num_blocks * 2;
  }
  NVB_CUDA(cudaMemcpyAsync(m->xyz_upload, blocks_xyz_host, (size_t)num_blocks * 3 * sizeof(int), cudaMemcpyHostToDevice, m->stream));
  NVB_CUDA(cudaStreamSynchronize(m->stream));
  int rc = enqueueEsdf(m, m->xyz_upload, num_blocks, false, true, plane);
  if (rc) return rc;
  if ((rc = nvb_mapper_synchronize(m))) return rc;
  return tightenEsdfBound(m);
}

int32_t nvb_esdf_slice_aabb(NvbMapper* m, float slice_height_m, float aabb_out[6], int32_t* empty_out) {
  if (!m || !aabb_out || !empty_out) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  *empty_out = 1;
  NVB_CUDA(cudaSetDevice(m->device));
  NVB_CUDA(syncAll(m));
  const int zb = (int)std::floor(slice_height_m / m->block_size);
  int* box_dev = m->esdf_ints + kGesCounts;  // four scratch ints (the wavefront is idle)
  const int init[4] = {INT32_MAX, INT32_MAX, INT32_MIN, INT32_MIN};
  NVB_CUDA(cudaMemcpyAsync(box_dev, init, sizeof(init), cudaMemcpyHostToDevice, m->stream));
  launchSliceAabb(m->esdf, zb, box_dev, m->stream);
  int box[4];
  NVB_CUDA(cudaMemcpyAsync(box, box_dev, sizeof(box), cudaMemcpyDeviceToHost, m->stream));
  NVB_CUDA(cudaStreamSynchronize(m->stream));
  NVB_CUDA(cudaMemsetAsync(box_dev, 0, sizeof(box), m->stream));
  m->launches++;
  if (box[0] > box[2]) return NVB_OK;  // no block at that height: empty AABB (:166-168)
  // getAABBOfBlock: [index * block_size, (index + 1) * block_size]
  const float bs = m->block_size;
  aabb_out[0] = (float)box[0] * bs, aabb_out[1] = (float)box[1] * bs, aabb_out[2] = (float)zb * bs;
  aabb_out[3] = ((float)box[2] + 1.0f) * bs, aabb_out[4] = ((float)box[3] + 1.0f) * bs, aabb_out[5] = ((float)zb + 1.0f) * bs;
  *empty_out = 0;
  return NVB_OK;
}

int32_t nvb_esdf_slice_distance_image_in_aabb(NvbMapper* m, float slice_height_m, float unobserved_value, const float aabb[6],
                                              float* image_host, int8_t* grid_host, int32_t cap_pixels, int32_t* rows_out,
                                              int32_t* cols_out) {
  if (!m || !aabb || !rows_out || !cols_out) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  *rows_out = *cols_out = 0;
  NVB_CUDA(cudaSetDevice(m->device));
  NVB_CUDA(syncAll(m));
  const float bs = m->block_size;
  const float voxel_size = bs / (float)kVps;
  const float sx = aabb[3] - aabb[0], sy = aabb[4] - aabb[1];
  if (!(sx > 0.0f) || !(sy > 0.0f)) return NVB_OK;  // aabb.isEmpty() (:175-177)
  const int cols = (int)std::ceil(sx / voxel_size), rows = (int)std::ceil(sy / voxel_size);
  *rows_out = rows, *cols_out = cols;
  const long long npix = (long long)rows * cols;
  if (npix <= 0 || (!image_host && !grid_host) || cap_pixels <= 0) return NVB_OK;
  float* img_dev = nullptr;
  signed char* grid_dev = nullptr;
  if (image_host) NVB_CUDA(cudaMalloc(&img_dev, (size_t)npix * sizeof(float)));
  if (grid_host) NVB_CUDA(cudaMalloc(&grid_dev, (size_t)npix));
  launchSliceImage(m->esdf, bs, aabb[0], aabb[1], slice_height_m, unobserved_value, rows, cols, img_dev, grid_dev, m->stream);
  m->launches++;
  const size_t k = (size_t)std::min<long long>(npix, cap_pixels);
  if (image_host) NVB_CUDA(cudaMemcpyAsync(image_host, img_dev, k * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
  if (grid_host) NVB_CUDA(cudaMemcpyAsync(grid_host, grid_dev, k, cudaMemcpyDeviceToHost, m->stream));
  NVB_CUDA(cudaStreamSynchronize(m->stream));
  cudaFree(img_dev), cudaFree(grid_dev);
  return NVB_OK;
}

int32_t nvb_esdf_slice_distance_image(NvbMapper* m, float slice_height_m, float unobserved_value, float aabb_out[6],
                                      float* image_host, int8_t* grid_host, int32_t cap_pixels, int32_t* rows_out,
                                      int32_t* cols_out) {
  if (!m || !rows_out || !cols_out) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  *rows_out = *cols_out = 0;
  float box[6];
  int32_t empty = 1;
  int rc = nvb_esdf_slice_aabb(m, slice_height_m, box, &empty);
  if (rc || empty) return rc;
  if (aabb_out)
    for (int a = 0; a < 6; a++) aabb_out[a] = box[a];
  return nvb_esdf_slice_distance_image_in_aabb(m, slice_height_m, unobserved_value, box, image_host, grid_host, cap_pixels, rows_out,
                                               cols_out);
}

int32_t nvb_mapper_synchronize(NvbMapper* m) {
  if (!m) return fail(NVB_ERR_INVALID_ARGUMENT, "null mapper");
  NVB_CUDA(cudaSetDevice(m->device));
  int rc = enqueueErrorCopies(m);
  if (rc) return rc;
  NVB_CUDA(syncAll(m));
  NVB_CUDA(cudaGetLastError());
  collectStages(m);
  return checkPrefetchedError(m);
}

int32_t nvb_mapper_last_frame_block_count(NvbMapper* m, int32_t* out_count) {
  if (!m || !out_count) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  NVB_CUDA(cudaSetDevice(m->device));
  return readFrameList(m, nullptr, 0, out_count);
}

int32_t nvb_mapper_last_frame_blocks(NvbMapper* m, int32_t* out_xyz_host, int32_t cap, int32_t* out_count) {
  if (!m) return fail(NVB_ERR_INVALID_ARGUMENT, "null mapper");
  NVB_CUDA(cudaSetDevice(m->device));
  return readFrameList(m, out_xyz_host, cap, out_count);
}

int32_t nvb_blocks_union(NvbMapper* m, const int32_t* xyz_dev, int32_t n, const int32_t aabb_min[3],
                         const int32_t aabb_max[3], int32_t* out_xyz_dev, int32_t cap, int32_t* out_count_host) {
  if (!m || !aabb_min || !aabb_max || (n > 0 && !xyz_dev)) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  NVB_CUDA(cudaSetDevice(m->device));
  ViewGrid g{};
  g.min_index = make_int3(aabb_min[0], aabb_min[1], aabb_min[2]);
  const long long sx = (long long)aabb_max[0] - aabb_min[0] + 1, sy = (long long)aabb_max[1] - aabb_min[1] + 1,
                  sz = (long long)aabb_max[2] - aabb_min[2] + 1;
  if (n <= 0 || sx <= 0 || sy <= 0 || sz <= 0) {
    if (out_count_host) *out_count_host = 0;
    return NVB_OK;
  }
  if (sx * sy * sz > 0x7fffffffll) return fail(NVB_ERR_CAPACITY, "union AABB has more than 2^31 cells");
  g.size = make_int3((int)sx, (int)sy, (int)sz);
  g.linear_size = (int)(sx * sy * sz);
  g.num_words = (g.linear_size + 31) / 32;
  int rc;
  // the union has its own list: the last frame's block list (nvb_mapper_last_frame_blocks) survives a merge
  ViewGrid scratch = g;
  scratch.linear_size = 0;  // bitset + tile state only
  if ((rc = ensureFrameScratch(m, scratch))) return rc;
  const int need = std::min(n, g.linear_size);
  if (need > m->union_list_cap) {
    NVB_CUDA(cudaStreamSynchronize(m->stream));
    if (m->union_list) cudaFree(m->union_list);
    m->union_list = nullptr, m->union_list_cap = 0;
    const int cap2 = (int)std::min<long long>((long long)(1.5 * need) + 64, 0x7fffffff);
    NVB_CUDA(cudaMalloc(&m->union_list, (size_t)cap2 * sizeof(int4)));
    m->union_list_cap = cap2;
  }
  if (!m->union_list_count) NVB_CUDA(cudaMalloc(&m->union_list_count, sizeof(int)));
  launchMarkList(xyz_dev, n, g, m->bits, m->stream);
  CompactArgs ca{};
  ca.bits = m->bits;
  ca.grid = g;
  ca.frame_blocks = m->union_list;
  ca.frame_count = m->union_list_count;
  ca.tile_state = m->tile_state;
  ca.ticket = m->ticket;
  ca.ticket_base = m->ticket_base;
  ca.epoch = ++m->epoch;
  ca.allocate = 0;
  ca.layer = m->tsdf;
  ca.error = m->error_dev;
  launchCompactAllocate(ca, m->stream);
  if (compactUsesTickets(g)) m->ticket_base += (unsigned int)compactNumTiles(g);
  launchClearBits(m->bits, g.num_words, m->stream);
  if (out_xyz_dev && cap > 0) launchUnpackList(m->union_list, m->union_list_count, out_xyz_dev, cap, m->stream);
  m->launches += 4;
  if (out_count_host) {
    NVB_CUDA(cudaMemcpyAsync(m->h_ints, m->union_list_count, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    NVB_CUDA(cudaStreamSynchronize(m->stream));
    *out_count_host = m->h_ints[0];
  }
  return NVB_OK;
}

int32_t nvb_mapper_append_frame_blocks(NvbMapper* m, int32_t* segment_dev, int32_t cap_entries) {
  if (!m || !segment_dev || cap_entries <= 0) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  NVB_CUDA(cudaSetDevice(m->device));
  if (!m->frame_blocks) return NVB_OK;  // no frame integrated yet
  launchAppendFrame(m->frame_blocks, m->frame_count, segment_dev, cap_entries, m->error_dev, m->stream);
  m->launches++;
  return NVB_OK;
}

int32_t nvb_blocks_union_segments(NvbMapper* m, const int32_t* segments_dev, int32_t num_segments, int32_t segment_stride_ints,
                                  int32_t cap_entries, int32_t* out_xyz_dev, int32_t out_cap, int32_t* out_count_dev, void* stream) {
  if (!m || !segments_dev || !out_xyz_dev || !out_count_dev || num_segments <= 0 || cap_entries <= 0 ||
      segment_stride_ints < 1 + 3 * cap_entries)
    return fail(NVB_ERR_INVALID_ARGUMENT, "bad argument");
  NVB_CUDA(cudaSetDevice(m->device));
  if (!m->union_bits) {
    // 2^28 cells = 32 MiB of bits: a union AABB of e.g. 1024 x 1024 x 256 blocks (410 m x 410 m x 102 m at 5 cm voxels)
    m->union_bits_cap = 1ll << 28;
    NVB_CUDA(cudaMalloc(&m->union_bits, (size_t)(m->union_bits_cap / 8)));
    NVB_CUDA(cudaMemset(m->union_bits, 0, (size_t)(m->union_bits_cap / 8)));
    NVB_CUDA(cudaMalloc(&m->union_state, 8 * sizeof(int)));
    NVB_CUDA(cudaMemset(m->union_state, 0, 8 * sizeof(int)));
  }
  launchUnionSegments(segments_dev, num_segments, segment_stride_ints, cap_entries, m->union_state, m->union_bits, m->union_bits_cap,
                      out_xyz_dev, out_cap, out_count_dev, stream ? (cudaStream_t)stream : m->stream);
  m->launches += 5;
  return NVB_OK;
}

int32_t nvb_blocks_union_status(NvbMapper* m, int32_t* out_error) {
  if (!m || !out_error) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  NVB_CUDA(cudaSetDevice(m->device));
  *out_error = 0;
  if (m->union_state) {
    NVB_CUDA(cudaDeviceSynchronize());
    int st[8];
    NVB_CUDA(cudaMemcpy(st, m->union_state, sizeof(st), cudaMemcpyDeviceToHost));
    *out_error = st[6];
  }
  return NVB_OK;
}

int32_t nvb_mapper_set_cache_last_viewpoint(NvbMapper* m, int32_t enable) {
  if (!m) return fail(NVB_ERR_INVALID_ARGUMENT, "null mapper");
  m->cache_last_viewpoint = enable ? 1 : 0;
  if (!enable) m->vc_n = 0;
  return NVB_OK;
}
int32_t nvb_mapper_get_cache_last_viewpoint(const NvbMapper* m) { return m ? m->cache_last_viewpoint : 0; }

int32_t nvb_mapper_set_esdf_reserved_sms(NvbMapper* m, int32_t reserved_sms) {
  if (!m) return fail(NVB_ERR_INVALID_ARGUMENT, "null mapper");
  if (reserved_sms < 0 || reserved_sms > 64) return fail(NVB_ERR_INVALID_ARGUMENT, "reserved_sms must be in [0, 64]");
  m->esdf_reserved_sms = reserved_sms;
  return NVB_OK;
}
int32_t nvb_mapper_get_esdf_reserved_sms(const NvbMapper* m) { return m ? m->esdf_reserved_sms : 0; }

int32_t nvb_mapper_set_depth_preprocessing(NvbMapper* m, int32_t enable, int32_t num_dilations) {
  if (!m) return fail(NVB_ERR_INVALID_ARGUMENT, "null mapper");
  if (num_dilations < 0 || num_dilations > kMaxDilations)
    return fail(NVB_ERR_INVALID_ARGUMENT, "depth_preprocessing_num_dilations must be in [0, 64]");
  m->do_depth_preprocessing = enable ? 1 : 0;
  m->depth_preprocessing_num_dilations = num_dilations;
  return NVB_OK;
}
int32_t nvb_mapper_get_depth_preprocessing(const NvbMapper* m, int32_t* enable, int32_t* num_dilations) {
  if (!m) return fail(NVB_ERR_INVALID_ARGUMENT, "null mapper");
  if (enable) *enable = m->do_depth_preprocessing;
  if (num_dilations) *num_dilations = m->depth_preprocessing_num_dilations;
  return NVB_OK;
}
int32_t nvb_depth_dilate_invalid(NvbMapper* m, const float* depth_dev, float* out_dev, int32_t rows, int32_t cols,
                                 int32_t num_dilations, float invalid_depth_threshold, float invalid_depth_value) {
  if (!m || !depth_dev || !out_dev) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  if (rows < 3 || cols < 3) return fail(NVB_ERR_INVALID_ARGUMENT, "rows and cols must be >= 3");
  if (depth_dev == out_dev) return fail(NVB_ERR_INVALID_ARGUMENT, "the output image must not alias the input");
  if (num_dilations < 0 || num_dilations > kMaxDilations)
    return fail(NVB_ERR_INVALID_ARGUMENT, "num_dilations must be in [0, 64]");
  NVB_CUDA(cudaSetDevice(m->device));
  launchDilateInvalid(depth_dev, out_dev, rows, cols, num_dilations, invalid_depth_threshold, invalid_depth_value, m->stream);
  NVB_CUDA(cudaGetLastError());
  return NVB_OK;
}

int32_t nvb_mapper_join_streams(NvbMapper* m) {
  if (!m) return fail(NVB_ERR_INVALID_ARGUMENT, "null mapper");
  NVB_CUDA(cudaSetDevice(m->device));
  NVB_CUDA(joinEsdf(m));
  return NVB_OK;
}

void* nvb_mapper_stream(NvbMapper* m) { return m ? (void*)m->stream : nullptr; }

// The projective layer is a TsdfLayer or an OccupancyLayer, never both (Mapper allocates the one its
// ProjectiveLayerType names, src/mapper/mapper.cpp:32-52): asking for the other one is an unknown layer.
static DevLayer* layerOf(NvbMapper* m, int layer) {
  if (layer == NVB_LAYER_TSDF) return m->projective_layer_type != NVB_PROJECTIVE_OCCUPANCY ? &m->tsdf : nullptr;
  if (layer == NVB_LAYER_FREESPACE) return m->freespace.blocks ? &m->freespace : nullptr;
  if (layer == NVB_LAYER_COLOR) {
    // created on first use (integration or query); an occupancy mapper never has one (mapper_impl.h:118-119)
    if (!m->color.blocks && m->projective_layer_type != NVB_PROJECTIVE_OCCUPANCY && ensureColorLayer(m)) return nullptr;
    return m->color.blocks ? &m->color : nullptr;
  }
  if (layer == NVB_LAYER_OCCUPANCY) return m->projective_layer_type == NVB_PROJECTIVE_OCCUPANCY ? &m->tsdf : nullptr;
  if (layer == NVB_LAYER_ESDF) return &m->esdf;
  if (layer == NVB_LAYER_MESH) return m->mesh.blocks ? &m->mesh : nullptr;  // the headers; the geometry: nvb_mesh_get_blocks
  return nullptr;
}

int32_t nvb_layer_block_bytes(int32_t layer) {
  if (layer == NVB_LAYER_TSDF) return kTsdfBlockBytes;
  if (layer == NVB_LAYER_ESDF) return kEsdfBlockBytes;
  if (layer == NVB_LAYER_OCCUPANCY) return kOccBlockBytes;
  if (layer == NVB_LAYER_FREESPACE) return kFreespaceBlockBytes;
  if (layer == NVB_LAYER_COLOR) return kColorBlockBytes;
  return 0;
}

int32_t nvb_layer_num_blocks(NvbMapper* m, int32_t layer, int32_t* out_count) {
  if (!m || !out_count) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  DevLayer* L = layerOf(m, layer);
  if (!L) return fail(NVB_ERR_INVALID_ARGUMENT, "unknown layer");
  NVB_CUDA(cudaSetDevice(m->device));
  NVB_CUDA(syncAll(m));
  int count = 0, nfree = 0;
  NVB_CUDA(cudaMemcpy(&count, L->count, sizeof(int), cudaMemcpyDeviceToHost));
  NVB_CUDA(cudaMemcpy(&nfree, L->free_count, sizeof(int), cudaMemcpyDeviceToHost));
  *out_count = std::min(count, L->capacity) - nfree;  // slots handed out minus deallocated ones
  return NVB_OK;
}

int32_t nvb_layer_block_indices(NvbMapper* m, int32_t layer, int32_t* out_xyz_host, int32_t cap, int32_t* out_count) {
  int n = 0;
  int rc = nvb_layer_num_blocks(m, layer, &n);
  if (rc) return rc;
  if (out_count) *out_count = n;
  DevLayer* L = layerOf(m, layer);
  if (out_xyz_host && cap > 0 && n > 0) {
    int hw = 0;
    NVB_CUDA(cudaMemcpy(&hw, L->count, sizeof(int), cudaMemcpyDeviceToHost));
    hw = std::min(hw, L->capacity);
    if (hw == n) {  // no deallocated slots below the high-water mark
      NVB_CUDA(cudaMemcpy(out_xyz_host, L->block_index, (size_t)std::min(n, cap) * 3 * sizeof(int), cudaMemcpyDeviceToHost));
    } else {
      std::vector<int> all((size_t)hw * 3);
      NVB_CUDA(cudaMemcpy(all.data(), L->block_index, all.size() * sizeof(int), cudaMemcpyDeviceToHost));
      int k = 0;
      for (int sl = 0; sl < hw && k < cap; sl++) {
        if (all[3 * sl] == kDeadSlotX) continue;
        out_xyz_host[3 * k] = all[3 * sl], out_xyz_host[3 * k + 1] = all[3 * sl + 1], out_xyz_host[3 * k + 2] = all[3 * sl + 2];
        k++;
      }
    }
  }
  return NVB_OK;
}

int32_t nvb_layer_get_blocks(NvbMapper* m, int32_t layer, const int32_t* xyz_host, int32_t n, void* out_host,
                             uint8_t* found_host) {
  if (!m || (n > 0 && (!xyz_host || !out_host))) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  DevLayer* L = layerOf(m, layer);
  if (!L) return fail(NVB_ERR_INVALID_ARGUMENT, "unknown layer");
  if (n <= 0) return NVB_OK;
  NVB_CUDA(cudaSetDevice(m->device));
  NVB_CUDA(syncAll(m));
  int* xyz_dev = nullptr;
  unsigned char *out_dev = nullptr, *found_dev = nullptr;
  NVB_CUDA(cudaMalloc(&xyz_dev, (size_t)n * 3 * sizeof(int)));
  NVB_CUDA(cudaMalloc(&out_dev, (size_t)n * L->block_bytes));
  NVB_CUDA(cudaMalloc(&found_dev, (size_t)n));
  NVB_CUDA(cudaMemcpyAsync(xyz_dev, xyz_host, (size_t)n * 3 * sizeof(int), cudaMemcpyHostToDevice, m->stream));
  launchGatherBlocks(*L, xyz_dev, n, out_dev, found_dev, m->stream);
  m->launches++;
  NVB_CUDA(cudaMemcpyAsync(out_host, out_dev, (size_t)n * L->block_bytes, cudaMemcpyDeviceToHost, m->stream));
  if (found_host) NVB_CUDA(cudaMemcpyAsync(found_host, found_dev, (size_t)n, cudaMemcpyDeviceToHost, m->stream));
  NVB_CUDA(syncAll(m));
  cudaFree(xyz_dev), cudaFree(out_dev), cudaFree(found_dev);
  return NVB_OK;
}

int32_t nvb_layer_set_blocks(NvbMapper* m, int32_t layer, const int32_t* xyz_host, int32_t n, const void* in_host) {
  if (!m || (n > 0 && (!xyz_host || !in_host))) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  DevLayer* L = layerOf(m, layer);
  if (!L) return fail(NVB_ERR_INVALID_ARGUMENT, "unknown layer");
  if (n <= 0) return NVB_OK;
  NVB_CUDA(cudaSetDevice(m->device));
  for (int i = 0; i < n; i++)
    if (!indexInRange(xyz_host[3 * i], xyz_host[3 * i + 1], xyz_host[3 * i + 2]))
      return fail(NVB_ERR_INDEX_RANGE, "block index outside +-2^20");
  int rc;
  if (layer == NVB_LAYER_TSDF) {
    if ((rc = ensureTsdfCapacity(m, n))) return rc;
    m->cells_cum += n;
    m->tsdf_count_ub += n;
  } else {
    m->esdf_extra_ub += n;
    if ((rc = ensureEsdfCapacity(m, (long long)std::min(m->tsdf_count_ub, m->tsdf.capacity) + m->esdf_extra_ub))) return rc;
  }
  L = layerOf(m, layer);
  NVB_CUDA(syncAll(m));
  int* xyz_dev = nullptr;
  unsigned char* in_dev = nullptr;
  NVB_CUDA(cudaMalloc(&xyz_dev, (size_t)n * 3 * sizeof(int)));
  NVB_CUDA(cudaMalloc(&in_dev, (size_t)n * L->block_bytes));
  NVB_CUDA(cudaMemcpyAsync(xyz_dev, xyz_host, (size_t)n * 3 * sizeof(int), cudaMemcpyHostToDevice, m->stream));
  NVB_CUDA(cudaMemcpyAsync(in_dev, in_host, (size_t)n * L->block_bytes, cudaMemcpyHostToDevice, m->stream));
  launchScatterBlocks(*L, xyz_dev, n, in_dev, m->error_dev, m->stream);
  m->launches++;
  if (layer == NVB_LAYER_ESDF) m->prune_ok = false;  // voxels written from outside: the parent boxes are no longer bounds
  NVB_CUDA(syncAll(m));
  cudaFree(xyz_dev), cudaFree(in_dev);
  if (layer == NVB_LAYER_TSDF) {
    // the tracker is told like after an integration: a later updateEsdf must see these blocks
    m->tracker_initialized = false;
  } else {
    // blocks created outside the ESDF update path are not linked: forget the neighbour table,
    // it is re-resolved lazily through the hash
    NVB_CUDA(cudaMemsetAsync(m->nbr, 0xFE, (size_t)m->esdf.capacity * 6 * sizeof(int), m->stream));
    NVB_CUDA(cudaStreamSynchronize(m->stream));
  }
  return checkDeviceError(m);
}

int32_t nvb_layer_block_device_ptr(NvbMapper* m, int32_t layer, const int32_t xyz[3], void** out_ptr) {
  if (!m || !xyz || !out_ptr) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  DevLayer* L = layerOf(m, layer);
  if (!L) return fail(NVB_ERR_INVALID_ARGUMENT, "unknown layer");
  NVB_CUDA(cudaSetDevice(m->device));
  NVB_CUDA(syncAll(m));
  // host-side probe of the device hash (small, synchronous): copy the probe window
  *out_ptr = nullptr;
  if (!indexInRange(xyz[0], xyz[1], xyz[2])) return NVB_OK;
  const unsigned long long key = packIndex(xyz[0], xyz[1], xyz[2]);
  unsigned int p = hashKey(key) & L->hash.mask;
  for (unsigned int probes = 0; probes <= L->hash.mask; probes++) {
    unsigned long long k = 0;
    NVB_CUDA(cudaMemcpy(&k, L->hash.keys + p, sizeof(k), cudaMemcpyDeviceToHost));
    if (k == key) {
      int slot = -1;
      NVB_CUDA(cudaMemcpy(&slot, L->hash.vals + p, sizeof(int), cudaMemcpyDeviceToHost));
      if (slot >= 0) *out_ptr = L->blocks + (size_t)slot * L->block_bytes;
      return NVB_OK;
    }
    if (k == kEmptyKey) return NVB_OK;
    p = (p + 1) & L->hash.mask;
  }
  return NVB_OK;
}

int32_t nvb_mapper_last_esdf_stats(NvbMapper* m, int64_t out[8]) {
  if (!m || !out) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  NVB_CUDA(cudaSetDevice(m->device));
  NVB_CUDA(syncAll(m));
  long long tmp[8];
  NVB_CUDA(cudaMemcpy(tmp, m->stats, sizeof(tmp), cudaMemcpyDeviceToHost));
  for (int i = 0; i < 8; i++) out[i] = tmp[i];
  return NVB_OK;
}

int32_t nvb_mapper_esdf_time_split(NvbMapper* m, int64_t out[4]) {
  if (!m || !out) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  NVB_CUDA(cudaSetDevice(m->device));
  NVB_CUDA(syncAll(m));
  long long tmp[5];
  NVB_CUDA(cudaMemcpy(tmp, m->stats + 8, sizeof(tmp), cudaMemcpyDeviceToHost));
  for (int i = 0; i < 4; i++) out[i] = tmp[i];
  out[0] = tmp[0];
  // [1]+[2] are CTA 0's own work; tmp[4] is the sum over phases of the slowest CTA's work: report it in [1]
  // of a second call convention: keep the API at 4 entries, fold it in as out[2] = slowest-CTA work total.
  out[2] = tmp[4];
  return NVB_OK;
}

int32_t nvb_mapper_esdf_clear_blocks_read(NvbMapper* m, int64_t* out) {
  if (!m || !out) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  NVB_CUDA(cudaSetDevice(m->device));
  NVB_CUDA(syncAll(m));
  long long v = 0;
  NVB_CUDA(cudaMemcpy(&v, m->stats + 13, sizeof(v), cudaMemcpyDeviceToHost));
  *out = v;
  return NVB_OK;
}

int32_t nvb_mapper_debug_phase_max(NvbMapper* m, int64_t* out, int32_t cap) {
  if (!m || !out) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  NVB_CUDA(cudaSetDevice(m->device));
  NVB_CUDA(syncAll(m));
  if (cap > 4000) cap = 4000;
  NVB_CUDA(cudaMemcpy(out, m->phase_max, (size_t)cap * sizeof(long long), cudaMemcpyDeviceToHost));
  return NVB_OK;
}

int32_t nvb_mapper_enable_profiling(NvbMapper* m, int32_t enable) {
  if (!m) return fail(NVB_ERR_INVALID_ARGUMENT, "null mapper");
  NVB_CUDA(cudaSetDevice(m->device));
  NVB_CUDA(syncAll(m));
  collectStages(m);
  m->profiling = enable != 0;
  return NVB_OK;
}

int32_t nvb_mapper_stage_times(NvbMapper* m, double* out_ms, int64_t* out_calls, int32_t reset) {
  if (!m) return fail(NVB_ERR_INVALID_ARGUMENT, "null mapper");
  NVB_CUDA(cudaSetDevice(m->device));
  NVB_CUDA(syncAll(m));
  collectStages(m);
  for (int i = 0; i < kNumStages; i++) {
    if (out_ms) out_ms[i] = m->stage_ms[i];
    if (out_calls) out_calls[i] = m->stage_calls[i];
    if (reset) m->stage_ms[i] = 0, m->stage_calls[i] = 0;
  }
  return NVB_OK;
}

int64_t nvb_mapper_kernel_launches(const NvbMapper* m) { return m ? m->launches : 0; }

}  // extern "C"

// ---------------------------------------------------------------------------
// Mesh (nvb_mesh.cu)
// ---------------------------------------------------------------------------
namespace {

int ensureMeshLayer(NvbMapper* m) {
  int rc;
  if (!m->mesh.blocks) {
    if ((rc = allocLayer(&m->mesh, m->tsdf.capacity, kMeshHeaderBytes, m->stream))) return rc;
    NVB_CUDA(cudaMalloc(&m->mesh_state, kArenaInts * sizeof(int)));
    NVB_CUDA(cudaMemsetAsync(m->mesh_state, 0, kArenaInts * sizeof(int), m->stream));
    NVB_CUDA(cudaMalloc(&m->dirty_mesh, (size_t)m->tsdf.capacity * sizeof(int)));
    NVB_CUDA(cudaMemsetAsync(m->dirty_mesh, 0, (size_t)m->tsdf.capacity * sizeof(int), m->stream));
    NVB_CUDA(cudaMalloc(&m->todo_mesh_slots, (size_t)m->tsdf.capacity * sizeof(int)));
  } else if (m->mesh.capacity < m->tsdf.capacity) {
    if ((rc = growLayer(m, &m->mesh, m->tsdf.capacity))) return rc;
  }
  return NVB_OK;
}

MeshCtx makeMeshCtx(NvbMapper* m) {
  MeshCtx c{};
  c.tsdf = m->tsdf, c.color = m->color, c.mesh = m->mesh;
  c.vertices = m->mesh_v, c.normals = m->mesh_n, c.triangles = m->mesh_t, c.colors_raw = m->mesh_c;
  c.colors = reinterpret_cast<uchar4*>(m->mesh_c);
  c.arena_state = m->mesh_state;
  c.counts = m->mesh_counts, c.offsets = m->mesh_offsets;
  c.block_size = m->block_size, c.voxel_size = m->voxel_size;
  c.min_weight = m->mp.min_weight, c.cutoff_distance_m = m->mp.cutoff_distance_vox * m->voxel_size;
  c.weld = m->mp.weld_vertices ? 1 : 0;
  c.error = m->error_dev;
  return c;
}

// Moves the live segments into a fresh arena of at least `need` free entries behind them (growth and garbage collection
// are the same operation: a segment is live while a header points at it).
int repackMeshArena(NvbMapper* m, long long need) {
  NVB_CUDA(cudaStreamSynchronize(m->stream));
  int nslots = 0;
  NVB_CUDA(cudaMemcpy(&nslots, m->mesh.count, sizeof(int), cudaMemcpyDeviceToHost));
  nslots = std::min(nslots, m->mesh.capacity);
  int* sizes = nullptr;
  int* new_off = nullptr;
  NVB_CUDA(cudaMalloc(&sizes, ((size_t)nslots + 1) * sizeof(int)));
  NVB_CUDA(cudaMalloc(&new_off, ((size_t)nslots + 1) * sizeof(int)));
  MeshCtx c = makeMeshCtx(m);
  launchMeshCompactSizes(c, nslots, sizes, m->stream);
  std::vector<int> h((size_t)nslots + 1, 0), o((size_t)nslots + 1, 0);
  NVB_CUDA(cudaMemcpyAsync(h.data(), sizes, (size_t)nslots * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
  NVB_CUDA(cudaStreamSynchronize(m->stream));
  long long live = 0;
  for (int i = 0; i < nslots; i++) o[i] = (int)live, live += h[i];
  // at least half of the arena is free behind the live data after a repack: with an update re-emitting ~1/5 of the live
  // vertices, that is several updates between repacks
  long long cap = std::max<long long>(m->mesh_arena_cap, 1 << 20);
  while (cap < 2 * (live + need)) cap *= 2;
  if (cap > 0x7fffffffll) {
    cudaFree(sizes), cudaFree(new_off);
    return fail(NVB_ERR_CAPACITY, "mesh arena beyond 2^31 vertices");
  }
  // the spare arena of the previous repack is reused when it has the right size (no cudaMalloc in the steady state)
  if (m->mesh_alt_cap != cap) {
    cudaFree(m->mesh_alt_v), cudaFree(m->mesh_alt_n), cudaFree(m->mesh_alt_t), cudaFree(m->mesh_alt_c);
    m->mesh_alt_v = m->mesh_alt_n = nullptr, m->mesh_alt_t = nullptr, m->mesh_alt_c = nullptr, m->mesh_alt_cap = 0;
    NVB_CUDA(cudaMalloc(&m->mesh_alt_v, (size_t)cap * 12));
    NVB_CUDA(cudaMalloc(&m->mesh_alt_n, (size_t)cap * 12));
    NVB_CUDA(cudaMalloc(&m->mesh_alt_t, (size_t)cap * 4));
    NVB_CUDA(cudaMalloc(&m->mesh_alt_c, (size_t)cap * 4));
    m->mesh_alt_cap = cap;
  }
  NVB_CUDA(cudaMemcpyAsync(new_off, o.data(), (size_t)nslots * sizeof(int), cudaMemcpyHostToDevice, m->stream));
  if (m->mesh_v) launchMeshCompactMove(c, nslots, new_off, m->mesh_alt_v, m->mesh_alt_n, m->mesh_alt_t, m->mesh_alt_c, m->num_sms, m->stream);
  const int state[kArenaInts] = {(int)live, 0, (int)live, 0};
  NVB_CUDA(cudaMemcpyAsync(m->mesh_state, state, sizeof(state), cudaMemcpyHostToDevice, m->stream));
  NVB_CUDA(cudaStreamSynchronize(m->stream));
  cudaFree(sizes), cudaFree(new_off);
  std::swap(m->mesh_v, m->mesh_alt_v), std::swap(m->mesh_n, m->mesh_alt_n), std::swap(m->mesh_t, m->mesh_alt_t);
  std::swap(m->mesh_c, m->mesh_alt_c), std::swap(m->mesh_arena_cap, m->mesh_alt_cap);
  m->launches += 2;
  return NVB_OK;
}

// One mesh update over a device list (explicit indices, or TSDF slots from the tracker).
int meshUpdateImpl(NvbMapper* m, const int* xyz_dev, const int* slots_dev, const int* count_dev, int upper, bool color) {
  int rc;
  if (upper <= 0) return NVB_OK;
  if (m->mesh_list_cap < upper) {
    NVB_CUDA(cudaStreamSynchronize(m->stream));
    cudaFree(m->mesh_counts), cudaFree(m->mesh_offsets);
    m->mesh_counts = m->mesh_offsets = nullptr;
    const int cap = std::max(upper, 2 * m->mesh_list_cap);
    NVB_CUDA(cudaMalloc(&m->mesh_counts, (size_t)cap * sizeof(int)));
    NVB_CUDA(cudaMalloc(&m->mesh_offsets, (size_t)cap * sizeof(int)));
    m->mesh_list_cap = cap;
  }
  MeshCtx c = makeMeshCtx(m);
  c.in_xyz = xyz_dev, c.in_slots = slots_dev, c.in_count_dev = count_dev, c.in_count_host = upper;
  c.tracker_dirty = slots_dev ? m->dirty_mesh : nullptr;
  launchMeshCount(c, upper, m->num_sms, m->stream);
  m->launches += 2;
  // the one number the host needs: does the update fit behind the arena's fill level?
  int state[kArenaInts];
  NVB_CUDA(cudaMemcpyAsync(state, m->mesh_state, sizeof(state), cudaMemcpyDeviceToHost, m->stream));
  NVB_CUDA(cudaStreamSynchronize(m->stream));
  if ((long long)state[kArenaLastBase] + state[kArenaLastTotal] > m->mesh_arena_cap) {
    if ((rc = repackMeshArena(m, state[kArenaLastTotal]))) return rc;
    c = makeMeshCtx(m);
    c.in_xyz = xyz_dev, c.in_slots = slots_dev, c.in_count_dev = count_dev, c.in_count_host = upper;
    launchMeshScan(c, m->stream);  // the offsets move with the fill level
    m->launches++;
  }
  launchMeshEmit(c, upper, m->num_sms, m->stream);
  m->launches += c.weld ? 2 : 1;
  if (color) {
    launchMeshColor(c, upper, m->num_sms, m->stream);
    m->launches++;
  }
  return NVB_OK;
}

int uploadMeshList(NvbMapper* m, const int32_t* xyz_host, int n, int* out_unique) {
  for (int i = 0; i < n; i++)
    if (!indexInRange(xyz_host[3 * i], xyz_host[3 * i + 1], xyz_host[3 * i + 2]))
      return fail(NVB_ERR_INDEX_RANGE, "block index outside +-2^20");
  // a set, like the tracker's list in the reference (the mesh layer's find-or-insert needs unique keys per launch)
  struct K3 {
    int x, y, z;
  };
  std::vector<K3> v((size_t)n);
  memcpy(v.data(), xyz_host, (size_t)n * sizeof(K3));
  std::sort(v.begin(), v.end(), [](const K3& a, const K3& b) { return a.x != b.x ? a.x < b.x : (a.y != b.y ? a.y < b.y : a.z < b.z); });
  v.erase(std::unique(v.begin(), v.end(), [](const K3& a, const K3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }), v.end());
  const int u = (int)v.size();
  if (m->mesh_xyz_cap < u) {
    NVB_CUDA(cudaStreamSynchronize(m->stream));
    cudaFree(m->mesh_xyz_dev);
    m->mesh_xyz_dev = nullptr;
    const int cap = std::max(u, 2 * m->mesh_xyz_cap);
    NVB_CUDA(cudaMalloc(&m->mesh_xyz_dev, (size_t)cap * 3 * sizeof(int)));
    m->mesh_xyz_cap = cap;
  }
  NVB_CUDA(cudaMemcpyAsync(m->mesh_xyz_dev, v.data(), (size_t)u * sizeof(K3), cudaMemcpyHostToDevice, m->stream));
  NVB_CUDA(cudaStreamSynchronize(m->stream));  // v dies with this scope
  *out_unique = u;
  return NVB_OK;
}

}  // namespace

extern "C" {

void nvb_default_mesh_params(NvbMeshParams* p) {
  if (!p) return;
  p->min_weight = 1e-4f;          // mesh/mesh_integrator_params.h:22-24
  p->weld_vertices = 1;           // mesh/mesh_integrator_params.h:25-27
  p->cutoff_distance_vox = 5.0f;  // MeshIntegrator::cutoff_distance_vox_, mesh/mesh_integrator.h:129
}
int32_t nvb_mapper_set_mesh_params(NvbMapper* m, const NvbMeshParams* p) {
  if (!m || !p) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  m->mp = *p;
  return NVB_OK;
}
int32_t nvb_mapper_get_mesh_params(const NvbMapper* m, NvbMeshParams* p) {
  if (!m || !p) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  *p = m->mp;
  return NVB_OK;
}

int32_t nvb_mapper_update_mesh(NvbMapper* m, int32_t update_full_layer) {
  if (!m) return fail(NVB_ERR_INVALID_ARGUMENT, "null mapper");
  // "Mesh is only updated for Tsdf layers (not for occupancy)" (src/mapper/mapper.cpp:380-383)
  if (m->projective_layer_type == NVB_PROJECTIVE_OCCUPANCY) return NVB_OK;
  NVB_CUDA(cudaSetDevice(m->device));
  int rc;
  if ((rc = ensureMeshLayer(m))) return rc;
  pollCounts(m);
  const int upper = std::min(m->tsdf_count_ub, m->tsdf.capacity);
  if (!m->mesh_tracker_initialized || update_full_layer) {
    launchTodoAll(m->tsdf, m->dirty_mesh, m->todo_mesh_slots, m->esdf_ints + kTodoMeshCount, m->stream);
    m->launches++;
    m->mesh_tracker_initialized = true;
  }
  // integrateBlocksGPU + updateAppearance over the tracker's blocks, then markBlocksAsUpdated (mapper.cpp:385-395)
  if ((rc = meshUpdateImpl(m, nullptr, m->todo_mesh_slots, m->esdf_ints + kTodoMeshCount, upper, true))) return rc;
  NVB_CUDA(cudaMemsetAsync(m->esdf_ints + kTodoMeshCount, 0, sizeof(int), m->stream));
  NVB_CUDA(cudaStreamSynchronize(m->stream));
  return checkDeviceError(m);
}

int32_t nvb_mesh_integrate_blocks(NvbMapper* m, const int32_t* blocks_xyz_host, int32_t num_blocks, int32_t update_color) {
  if (!m) return fail(NVB_ERR_INVALID_ARGUMENT, "null mapper");
  if (m->projective_layer_type == NVB_PROJECTIVE_OCCUPANCY) return fail(NVB_ERR_INVALID_ARGUMENT, "the mapper has no TSDF layer");
  if (num_blocks < 0 || (num_blocks > 0 && !blocks_xyz_host)) return fail(NVB_ERR_INVALID_ARGUMENT, "bad block list");
  NVB_CUDA(cudaSetDevice(m->device));
  int rc;
  if ((rc = ensureMeshLayer(m))) return rc;
  if (num_blocks == 0) return NVB_OK;  // (:73-75)
  int u = 0;
  if ((rc = uploadMeshList(m, blocks_xyz_host, num_blocks, &u))) return rc;
  if ((rc = meshUpdateImpl(m, m->mesh_xyz_dev, nullptr, nullptr, u, update_color != 0))) return rc;
  NVB_CUDA(cudaStreamSynchronize(m->stream));
  return checkDeviceError(m);
}

int32_t nvb_mesh_update_color(NvbMapper* m, const int32_t* blocks_xyz_host, int32_t num_blocks) {
  if (!m) return fail(NVB_ERR_INVALID_ARGUMENT, "null mapper");
  if (num_blocks < 0 || (num_blocks > 0 && !blocks_xyz_host)) return fail(NVB_ERR_INVALID_ARGUMENT, "bad block list");
  NVB_CUDA(cudaSetDevice(m->device));
  int rc;
  if ((rc = ensureMeshLayer(m))) return rc;
  if (num_blocks == 0 || !m->mesh_v) return NVB_OK;
  int u = 0;
  if ((rc = uploadMeshList(m, blocks_xyz_host, num_blocks, &u))) return rc;
  MeshCtx c = makeMeshCtx(m);
  c.in_xyz = m->mesh_xyz_dev, c.in_count_host = u;
  launchMeshColor(c, u, m->num_sms, m->stream);
  m->launches++;
  NVB_CUDA(cudaStreamSynchronize(m->stream));
  return NVB_OK;
}

int32_t nvb_mesh_block_sizes(NvbMapper* m, const int32_t* blocks_xyz_host, int32_t num_blocks, int32_t* sizes_out) {
  if (!m || (num_blocks > 0 && (!blocks_xyz_host || !sizes_out))) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  if (num_blocks <= 0) return NVB_OK;
  NVB_CUDA(cudaSetDevice(m->device));
  for (int i = 0; i < 3 * num_blocks; i++) sizes_out[i] = -1;
  if (!m->mesh.blocks) return NVB_OK;
  NVB_CUDA(syncAll(m));
  int* dev = nullptr;
  NVB_CUDA(cudaMalloc(&dev, (size_t)num_blocks * 7 * sizeof(int)));  // headers[4n] (int4-aligned) | xyz[3n]
  int* xyz_dev = dev + 4 * (size_t)num_blocks;
  NVB_CUDA(cudaMemcpy(xyz_dev, blocks_xyz_host, (size_t)num_blocks * 3 * sizeof(int), cudaMemcpyHostToDevice));
  launchMeshHeaders(makeMeshCtx(m), xyz_dev, num_blocks, dev, m->stream);
  std::vector<int> h((size_t)num_blocks * 4);
  NVB_CUDA(cudaMemcpyAsync(h.data(), dev, h.size() * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
  NVB_CUDA(cudaStreamSynchronize(m->stream));
  cudaFree(dev);
  for (int i = 0; i < num_blocks; i++)
    if (h[4 * i] >= 0) sizes_out[3 * i] = h[4 * i + 1], sizes_out[3 * i + 1] = h[4 * i + 2], sizes_out[3 * i + 2] = h[4 * i + 3];
  return NVB_OK;
}

int32_t nvb_mesh_get_blocks(NvbMapper* m, const int32_t* blocks_xyz_host, int32_t num_blocks, float* vertices_out,
                            float* normals_out, int32_t* triangles_out, uint8_t* colors_out, const int64_t caps[3]) {
  if (!m || !caps || (num_blocks > 0 && !blocks_xyz_host)) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  if (num_blocks <= 0 || !m->mesh.blocks) return NVB_OK;
  NVB_CUDA(cudaSetDevice(m->device));
  NVB_CUDA(syncAll(m));
  const size_t n = (size_t)num_blocks;
  int* dev = nullptr;  // headers[4n] (int4-aligned) | xyz[3n] | dst[3n]
  NVB_CUDA(cudaMalloc(&dev, n * 10 * sizeof(int)));
  NVB_CUDA(cudaMemcpy(dev + 4 * n, blocks_xyz_host, n * 3 * sizeof(int), cudaMemcpyHostToDevice));
  MeshCtx c = makeMeshCtx(m);
  launchMeshHeaders(c, dev + 4 * n, num_blocks, dev, m->stream);
  std::vector<int> h(n * 4), dst(n * 3);
  NVB_CUDA(cudaMemcpyAsync(h.data(), dev, h.size() * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
  NVB_CUDA(cudaStreamSynchronize(m->stream));
  long long tv = 0, tt = 0, tc = 0;
  for (size_t i = 0; i < n; i++) {
    dst[3 * i] = (int)tv, dst[3 * i + 1] = (int)tt, dst[3 * i + 2] = (int)tc;
    if (h[4 * i] >= 0) tv += h[4 * i + 1], tt += h[4 * i + 2], tc += h[4 * i + 3];
  }
  if (tv > caps[0] || tt > caps[1] || tc > caps[2]) {
    cudaFree(dev);
    return fail(NVB_ERR_CAPACITY, "output buffers smaller than the listed mesh blocks");
  }
  float *pv = nullptr, *pn = nullptr;
  int* pt = nullptr;
  unsigned char* pc = nullptr;
  NVB_CUDA(cudaMalloc(&pv, std::max<size_t>(1, (size_t)tv * 12)));
  NVB_CUDA(cudaMalloc(&pn, std::max<size_t>(1, (size_t)tv * 12)));
  NVB_CUDA(cudaMalloc(&pt, std::max<size_t>(1, (size_t)tt * 4)));
  NVB_CUDA(cudaMalloc(&pc, std::max<size_t>(1, (size_t)tc * 4)));
  NVB_CUDA(cudaMemcpyAsync(dev + 7 * n, dst.data(), n * 3 * sizeof(int), cudaMemcpyHostToDevice, m->stream));
  launchMeshPack(c, dev, dev + 7 * n, num_blocks, pv, pn, pt, pc, m->num_sms, m->stream);
  if (vertices_out && tv) NVB_CUDA(cudaMemcpyAsync(vertices_out, pv, (size_t)tv * 12, cudaMemcpyDeviceToHost, m->stream));
  if (normals_out && tv) NVB_CUDA(cudaMemcpyAsync(normals_out, pn, (size_t)tv * 12, cudaMemcpyDeviceToHost, m->stream));
  if (triangles_out && tt) NVB_CUDA(cudaMemcpyAsync(triangles_out, pt, (size_t)tt * 4, cudaMemcpyDeviceToHost, m->stream));
  if (colors_out && tc) NVB_CUDA(cudaMemcpyAsync(colors_out, pc, (size_t)tc * 4, cudaMemcpyDeviceToHost, m->stream));
  NVB_CUDA(cudaStreamSynchronize(m->stream));
  cudaFree(dev), cudaFree(pv), cudaFree(pn), cudaFree(pt), cudaFree(pc);
  return NVB_OK;
}

int32_t nvb_mesh_arena_stats(NvbMapper* m, int64_t out[4]) {
  if (!m || !out) return fail(NVB_ERR_INVALID_ARGUMENT, "null argument");
  out[0] = m->mesh_arena_cap, out[1] = out[2] = out[3] = 0;
  if (!m->mesh.blocks) return NVB_OK;
  NVB_CUDA(cudaSetDevice(m->device));
  NVB_CUDA(syncAll(m));
  int state[kArenaInts];
  NVB_CUDA(cudaMemcpy(state, m->mesh_state, sizeof(state), cudaMemcpyDeviceToHost));
  out[1] = state[kArenaUsed], out[2] = state[kArenaLastTotal], out[3] = state[kArenaGarbage];
  return NVB_OK;
}

}  // extern "C"
