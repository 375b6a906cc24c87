// nvb_esdf_wavex.cu -- the ESDF wavefront (computeEsdf, nvblox/src/integrators/esdf_integrator.cu:1465-1496) with ONE
// grid barrier per ring and no contended atomics: "exchange-slab" wavefront, esdf_persistent = 3.
//
// The four-phase wavefront (nvb_esdf_wave.cu) pays four grid barriers per ring whatever the ring's size (1.25 us each on 148
// SMs, ~18 rings per frame) plus the dependent L2 round trips between them. What a ring costs is its dependent chain, so this
// kernel shortens the chain:
//
//   * A ring's six face passes (+x,-x,+y,-y,+z,-z, each seeing the previous ones, :1323-1386) move information across block
//     boundaries by one voxel. Their effect on a block B is a function of B, of the one-voxel halo around it (10x10x10 voxels
//     out of the 3x3x3 block neighbourhood) as it was at the START of the ring, and of which of those 27 blocks are members
//     (sources) of the ring. The owner of a CANDIDATE block (a face neighbour of a member) gathers that region into shared
//     memory, replays the passes there in the reference's order, and -- if B changed, i.e. B is a member of ring+1 -- sweeps
//     it at once (sweepBlockBandKernel, :1390-1431). No communication inside a ring.
//   * Sources are always MEMBERS. Every member's state at the start of the ring is published in an exchange slab indexed by
//     slot, X[ring & 1]: the owner of a block that changes in ring r writes the new block to the layer (read by the block's
//     next owner only) and to X[(r+1) & 1] (read by its neighbours' owners during ring r+1) while ring r's readers read
//     X[r & 1]. Two slabs by ring parity, one barrier per ring, no version words, no copy-back phase.
//   * Only what can matter is fetched and replayed. A pass-p pair (source voxel -> destination voxel) matters iff its source
//     block is a member and its destination block is B or a member that can still influence B through the LATER passes: the
//     blocks whose state after pass p matters are D6 = {B}, D5 = D6 + {B+z}, D4 = D5 + {B-z}, D3 = D4 + (D4 + y),
//     D2 = D3 + (D3 - y), D1 = D2 + (D2 + x) (backwards through -z,+z,-y,+y,-x,+x). With one to three member neighbours --
//     the usual case -- one or two of the six passes are live and a few per cent of the 1 200 pair slots; halo voxels of
//     blocks that are in no live pair are not fetched (they would miss L2: nobody wrote them lately).
//   * Region layout in shared memory: a plane of 16-byte cells {squared distance, parent} and a plane of flag words, voxel
//     index rx*110 + ry*11 + rz. Every line of the three sweeps and every pair of the replay is ONE conflict-free 128-bit
//     access per voxel (the 20-byte array-of-structures layout of the layer costs five 32-bit accesses with up to 8-way bank
//     conflicts on z lines; with eight groups per SM the kernel is bound by shared-memory instructions, not by latency).
//   * The sweep keeps, per line, the running site as (offset along the line, squared perpendicular offset, the two
//     perpendicular components): the loop-carried chain per voxel is one subtract, one multiply-add, one compare, one select.
//   * No hot words. Candidates of ring r+1 are registered by the owners of the blocks that changed in ring r (unique by an
//     atomicExch on a per-slot stamp) as RECORDS {slot, its 27 neighbour slots} in a per-CTA segment of the ring's list,
//     positions from a shared-memory counter; each CTA publishes its two counts (registrations, changed blocks) with a plain
//     store before the barrier, and after it every CTA reads the 148 pairs, scans them and deals the ring's candidates
//     round-robin (CTA, then 64-thread group). A same-address global atomic costs ~27 cycles per arrival (serialised in L2):
//     600 registrations per ring on one counter were 8 us.
//   * Rings with at most one candidate per 64-thread group of ONE CTA are run by CTA 0 alone with block-level barriers (half
//     of a frame's rings); everybody else waits at one grid barrier.
//   * 512-thread CTAs = 8 groups per SM x 148 SMs: 1 184 candidates in flight at once.
#include "nvb_esdf_wave_common.cuh"

#include <cstdlib>

namespace nvb {

namespace {

#ifndef NVB_WAVEX_THREADS
#define NVB_WAVEX_THREADS 512
#endif
#ifndef NVB_WAVEX_MAXREG
#define NVB_WAVEX_MAXREG 96  // 512 x 96 = 3/4 of the register file: the next frame's raycast / compaction / TSDF CTAs co-reside (measured: 2 980 vs 2 895 frames/s with 128)
#endif
#ifndef NVB_WAVEX_TAIL
#define NVB_WAVEX_TAIL 1
#endif
#ifndef NVB_WAVEX_PROF
#define NVB_WAVEX_PROF 0  // 1: per-stage cycle counters of CTA 0 / group 0 (profiles/wavex_split.py); costs 16 registers
#endif
#if NVB_WAVEX_PROF
#define X_PROF_BEGIN() long long tq = clock64();
#define X_PROF(i) prof[i] += clock64() - tq, tq = clock64();
#define X_PROF_COUNT(i) prof[i]++;
#else
#define X_PROF_BEGIN()
#define X_PROF(i)
#define X_PROF_COUNT(i)
#endif
constexpr int kXT = NVB_WAVEX_THREADS;
constexpr int kXG = kXT / 64;
constexpr int kRecInts = 32;      // candidate record: [0] slot, [1..27] its 3x3x3 neighbour slots, [28..31] unused (128-byte records)
constexpr int kMaxCtas = 192;     // per-CTA flags
constexpr int kRX = 110, kRY = 11;  // region voxel index = rx * 110 + ry * 11 + rz, rx, ry, rz in 0..9 (block voxel + 1)
constexpr int kRegionVox = 1100;
constexpr int kFlagBase = 4 * kRegionVox;        // word offset of the flag plane
constexpr int kXRegionWords = 5 * kRegionVox;    // 5 500 words = 22 000 bytes per group
constexpr size_t kXSmemBytes = (size_t)kXG * kXRegionWords * sizeof(unsigned int);

struct XTables {
  unsigned int halo[8][64];     // halo voxel copies: dst voxel | src voxel in its block << 11 | d27 << 20 | valid << 25
  unsigned int pair[6][4][64];  // boundary pairs per pass: src voxel | dst voxel << 11 | d27 of the source block << 22 | inner << 27 | valid << 28
};

struct XShared {
  int rec[kXG][kRecInts];  // the record of the candidate each group is working on
  int nb[kXG][8];          // face neighbours (+x,-x,+y,-y,+z,-z) of the block being registered from
  int pos[kXG][8];         // position won for that neighbour in this CTA's segment, -1: somebody else registered it
  unsigned int mask[kXG];
  int changed[kXG];
  unsigned int live[kXG][8];  // per pass: source blocks whose pairs are replayed (liveMasks)
  int seg[kXG][2];         // where the group's entry lives: registering CTA, index in its segment
  int pre[kMaxCtas + 1];   // exclusive scan of the per-CTA registration counts of the current ring
  int warp_tot[2][8];
  int ncand, nchanged;     // this CTA's registrations / changed blocks in the ring it is processing
  int next;                // next unclaimed entry of this CTA's share of the ring (groups pull work: a changed candidate costs
                           // twice an unchanged one, a static deal left groups with two changed ones on the critical path)
  int cur[kXG];
  int bcast[4];
};

// What the per-candidate functions need of the kernel parameter, in shared memory: they are real calls (one copy of the code for
// the grid rings and the single-CTA tail), and a reference to a kernel parameter would be copied to the stack at every call.
struct XCtx {
  unsigned char* blocks;  // the ESDF layer's slab
  int* block_index;
  DevHash hash;
  int* nbr;
  int* nbr27;
  int* cand_stamp;
  unsigned int* psum;     // per slot: box of the block offsets the voxels' parents point into (clear-pass pruning)
  int* stamp[2];          // member stamps by ring parity
  unsigned char* X[2];    // exchange slabs by ring parity
  int* recs[2];           // candidate records by ring parity: one segment of `seg` records per CTA
  int seg;
  float max_sq;
  __device__ __forceinline__ int* segment(int p, int cta) const { return recs[p] + (size_t)cta * seg * kRecInts; }
};

__device__ __forceinline__ int rvox(int rx, int ry, int rz) { return rx * kRX + ry * kRY + rz; }
__device__ __forceinline__ int faceEntry(int f) {  // +x,-x,+y,-y,+z,-z -> index into a 3x3x3 row
  return f == 0 ? 22 : (f == 1 ? 4 : (f == 2 ? 16 : (f == 3 ? 10 : (f == 4 ? 14 : 12))));
}
__device__ __forceinline__ int boundaryOff(int r) { return r == 0 ? -1 : (r == 9 ? 1 : 0); }   // region coordinate -> block offset
__device__ __forceinline__ int boundaryLoc(int r) { return r == 0 ? 7 : (r == 9 ? 0 : r - 1); }  // ... and voxel coordinate in that block

// nbr27 entry of `slot` towards offset d (0..26), resolving "never linked" entries (blocks created outside the ESDF
// update path) through the hash once.
__device__ __forceinline__ int neighbor27(const XCtx& c, int slot, int d) {
  int v = __ldcg(c.nbr27 + 27 * slot + d);
  if (v < -1) {
    const int* bi = c.block_index + 3 * slot;
    v = (d == 13) ? slot : hashFind(c.hash, bi[0] + d / 9 - 1, bi[1] + (d / 3) % 3 - 1, bi[2] + d % 3 - 1);
    c.nbr27[27 * slot + d] = v;
  }
  return v;
}
// face neighbour of `slot` (+x,-x,+y,-y,+z,-z), same resolution rule
__device__ __forceinline__ int neighbor6(const XCtx& c, int slot, int dir) {
  int v = __ldcg(c.nbr + 6 * slot + dir);
  if (v < -1) {
    const int* bi = c.block_index + 3 * slot;
    const int d = (dir & 1) ? -1 : 1;
    v = hashFind(c.hash, bi[0] + ((dir >> 1) == 0 ? d : 0), bi[1] + ((dir >> 1) == 1 ? d : 0), bi[2] + ((dir >> 1) == 2 ? d : 0));
    c.nbr[6 * slot + dir] = v;
  }
  return v;
}

// Per-launch tables (the same for every candidate).
__device__ __forceinline__ void initTables(XTables& tab, int tid) {
  // ---- halo copies. Batches 0..5: the six faces (+x,-x,+y,-y,+z,-z), one voxel per lane, so a batch is live or dead for the
  // whole group; lanes follow the source block's memory order where its face is contiguous. Batches 6, 7: edges and corners.
  for (int i = tid; i < 512; i += kXT) {
    const int k = i >> 6, lane = i & 63;
    unsigned int e = 0;
    int ra = 1, r1 = 1, r2 = 1, axis = 0;  // region coordinates: along `axis`, and the two others in axis order
    bool valid = true;
    if (k < 6) {
      axis = k >> 1;
      ra = (k & 1) ? 0 : 9, r1 = (lane >> 3) + 1, r2 = (lane & 7) + 1;
    } else {
      const int n = i - 384;
      if (n < 96) {  // 12 edges x 8 voxels
        const int edge = n >> 3, cn = edge & 3;
        axis = edge >> 2;
        ra = (n & 7) + 1, r1 = (cn & 1) ? 9 : 0, r2 = (cn & 2) ? 9 : 0;
      } else if (n < 104) {
        const int cn = n - 96;
        ra = (cn & 1) ? 9 : 0, r1 = (cn & 2) ? 9 : 0, r2 = (cn & 4) ? 9 : 0;
      } else {
        valid = false;
      }
    }
    if (valid) {
      const int rx = axis == 0 ? ra : r1, ry = axis == 0 ? r1 : (axis == 1 ? ra : r2), rz = axis == 2 ? ra : r2;
      const int d = (boundaryOff(rx) + 1) * 9 + (boundaryOff(ry) + 1) * 3 + (boundaryOff(rz) + 1);
      e = (unsigned)rvox(rx, ry, rz) | ((unsigned)(boundaryLoc(rx) * 64 + boundaryLoc(ry) * 8 + boundaryLoc(rz)) << 11) |
          ((unsigned)d << 20) | (1u << 25);
    }
    tab.halo[k][lane] = e;
  }
  // ---- boundary pairs. Slot 0 / 1: the 8x8 interior of the two boundary planes of the pass (one source block each, so the
  // slot is live or dead for the whole group); slots 2, 3: the 2 x 36 border pairs (sources in edge / corner blocks).
  for (int i = tid; i < 6 * 4 * 64; i += kXT) {
    const int pass = i >> 8, j = (i >> 6) & 3, lane = i & 63;
    const int axis = pass >> 1, dir = (pass & 1) ? -1 : 1;
    int plane, u, w;
    bool valid = true;
    if (j < 2) {
      plane = j, u = (lane >> 3) + 1, w = (lane & 7) + 1;
    } else {
      const int n = (j - 2) * 64 + lane;
      plane = n / 36;
      const int q = n % 36;
      if (q < 10) u = 0, w = q;
      else if (q < 20) u = 9, w = q - 10;
      else if (q < 28) u = q - 20 + 1, w = 0;
      else u = q - 28 + 1, w = 9;
      if (n >= 72) valid = false, plane = 0, u = 1, w = 1;
    }
    unsigned int e = 0;
    if (valid) {
      const int sa = dir > 0 ? (plane ? 8 : 0) : (plane ? 9 : 1);  // source coordinate along the axis
      const int so = dir > 0 ? (plane ? 0 : -1) : (plane ? 1 : 0);  // block offset of the source along the axis
      const int da = sa + dir;
      const int A = axis == 0 ? 9 : (axis == 1 ? 3 : 1), U = axis == 0 ? 3 : 9, W = axis == 2 ? 3 : 1;
      const int d = 13 + so * A + boundaryOff(u) * U + boundaryOff(w) * W;
      const int inner = da >= 1 && da <= 8 && u >= 1 && u <= 8 && w >= 1 && w <= 8;
      const int sv = axis == 0 ? rvox(sa, u, w) : (axis == 1 ? rvox(u, sa, w) : rvox(u, w, sa));
      const int dv = axis == 0 ? rvox(da, u, w) : (axis == 1 ? rvox(u, da, w) : rvox(u, w, da));
      e = (unsigned)sv | ((unsigned)dv << 11) | ((unsigned)d << 22) | ((unsigned)inner << 27) | (1u << 28);
    }
    tab.pair[pass][j][lane] = e;
  }
}

// Which source blocks are live in each pass, and which neighbours have to be fetched at all (see the header): bit d of
// live[p] <=> the pairs of pass p whose source block is d are replayed.
struct LiveMasks {
  unsigned int* live;   // [6], in shared memory (indexed by a run-time pass number)
  unsigned int needed;  // members (without B) that take part in a live pair
};
__device__ __forceinline__ LiveMasks liveMasks(unsigned int mask, unsigned int* live_smem, bool writer) {
  LiveMasks L;
  L.live = live_smem;
  const unsigned int ok = mask | (1u << 13);  // destinations that matter at all: B or a member
  // D(p+1): blocks whose state after pass p can still reach B; bit d = (dx+1)*9 + (dy+1)*3 + (dz+1)
  const unsigned int D[6] = {0x7FFFE00u, 0x3FE00u, 0x3F000u, 0x7000u, 0x6000u, 0x2000u};
  unsigned int acc = 0;
#pragma unroll
  for (int p = 0; p < 6; p++) {
    const int axis = p >> 1;
    const int A = axis == 0 ? 9 : (axis == 1 ? 3 : 1);
    const unsigned int lo = axis == 0 ? 0x000001ffu : (axis == 1 ? 0x001c0e07u : 0x01249249u);  // blocks at offset -1 along the axis
    const unsigned int dst_ok = ok & D[p];
    unsigned int lv, dst;
    if ((p & 1) == 0) {  // +dir: sources at offset -1, 0; destination = source + A
      lv = mask & (lo | (lo << A)) & (dst_ok >> A);
      dst = lv << A;
    } else {  // -dir: sources at offset 0, +1; destination = source - A
      lv = mask & ((lo << A) | (lo << (2 * A))) & (dst_ok << A);
      dst = lv >> A;
    }
    if (writer) live_smem[p] = lv;
    acc |= lv | dst;
  }
  L.needed = acc & mask & ~(1u << 13);
  return L;
}

// ---- the candidate's own block: one z-row (8 voxels, 160 contiguous bytes) per lane, through registers into the two planes
struct OwnRegs {
  uint4 q[10];
};
__device__ __forceinline__ OwnRegs ownLoad(const unsigned char* blk, int lane64) {
  OwnRegs o;
  const uint4* src = reinterpret_cast<const uint4*>(blk) + lane64 * 10;
#pragma unroll
  for (int i = 0; i < 10; i++) o.q[i] = __ldcg(src + i);
  return o;
}
__device__ __forceinline__ unsigned int ownWord(const OwnRegs& o, int w) {  // w is a compile-time constant after unrolling
  const uint4& v = o.q[w >> 2];
  return (w & 3) == 0 ? v.x : ((w & 3) == 1 ? v.y : ((w & 3) == 2 ? v.z : v.w));
}
__device__ __forceinline__ void ownToShared(unsigned int* R, const OwnRegs& o, int lane64) {
  uint4* A = reinterpret_cast<uint4*>(R);
  const int v0 = rvox((lane64 >> 3) + 1, (lane64 & 7) + 1, 1);
#pragma unroll
  for (int z = 0; z < 8; z++) {
    A[v0 + z] = make_uint4(ownWord(o, 5 * z), ownWord(o, 5 * z + 1), ownWord(o, 5 * z + 2), ownWord(o, 5 * z + 3));
    R[kFlagBase + v0 + z] = ownWord(o, 5 * z + 4);
  }
}
// inner 8x8x8 of the region -> the block in the layer (if `to_layer`) and its copy in an exchange slab; half a z-row (four
// voxels = 80 bytes = five 16-byte words) at a time. On the way the box of the BLOCK OFFSETS the voxels' parents point into is
// collected and published in c.psum (EsdfCtx::psum; one word per warp of the group, so no barrier is needed): the clear pass of
// later updates reads a candidate block only if that box contains a to-clear block.
__device__ __forceinline__ void ownStore(unsigned char* layer_blk, bool to_layer, unsigned char* x_blk, const unsigned int* R,
                                         unsigned int* psum_slot, int lane64) {
  const uint4* A = reinterpret_cast<const uint4*>(R);
  const int x = lane64 >> 3, y = lane64 & 7;
  const int v0 = rvox(x + 1, y + 1, 1);
  uint4* dl = reinterpret_cast<uint4*>(layer_blk) + lane64 * 10;
  uint4* dx = reinterpret_cast<uint4*>(x_blk) + lane64 * 10;
  int lo0 = 99, lo1 = 99, lo2 = 99, hi0 = -99, hi1 = -99, hi2 = -99;
#pragma unroll
  for (int h = 0; h < 2; h++) {
    unsigned int w[20];
#pragma unroll
    for (int z = 0; z < 4; z++) {
      const uint4 a = A[v0 + 4 * h + z];
      w[5 * z] = a.x, w[5 * z + 1] = a.y, w[5 * z + 2] = a.z, w[5 * z + 3] = a.w;
      w[5 * z + 4] = R[kFlagBase + v0 + 4 * h + z];
      if ((a.y | a.z | a.w) != 0u) {
        const int b0 = (x + (int)a.y) >> 3, b1 = (y + (int)a.z) >> 3, b2 = (4 * h + z + (int)a.w) >> 3;  // floor: arithmetic shift
        lo0 = min(lo0, b0), hi0 = max(hi0, b0), lo1 = min(lo1, b1), hi1 = max(hi1, b1), lo2 = min(lo2, b2), hi2 = max(hi2, b2);
      }
    }
#pragma unroll
    for (int i = 0; i < 5; i++) {
      const uint4 v = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
      if (to_layer) __stcg(dl + 5 * h + i, v);
      __stcg(dx + 5 * h + i, v);
    }
  }
  if (psum_slot) {
    lo0 = __reduce_min_sync(0xffffffffu, lo0), lo1 = __reduce_min_sync(0xffffffffu, lo1), lo2 = __reduce_min_sync(0xffffffffu, lo2);
    hi0 = __reduce_max_sync(0xffffffffu, hi0), hi1 = __reduce_max_sync(0xffffffffu, hi1), hi2 = __reduce_max_sync(0xffffffffu, hi2);
    if ((lane64 & 31) == 0) {
      unsigned int v = 0u;
      if (lo0 <= hi0) {
        const bool fits = lo0 >= -16 && lo1 >= -16 && lo2 >= -16 && hi0 <= 15 && hi1 <= 15 && hi2 <= 15;
        v = fits ? ((1u << 31) | (unsigned)(lo0 + 16) | ((unsigned)(hi0 + 16) << 5) | ((unsigned)(lo1 + 16) << 10) |
                    ((unsigned)(hi1 + 16) << 15) | ((unsigned)(lo2 + 16) << 20) | ((unsigned)(hi2 + 16) << 25))
                 : 0xffffffffu;
      }
      psum_slot[lane64 >> 5] = v;
    }
  }
}

// ---- halo voxels of the neighbours in `needed`, from the exchange slab of this ring. The eight batches of the table (six
// faces, edges + corners) are walked four live ones at a time: a candidate with up to four live batches -- nearly all of them
// -- pays one round trip.
__device__ __forceinline__ void haloGather(const XTables& tab, unsigned int* R, const int* row, unsigned int needed,
                                           const unsigned char* X, int lane64) {
  uint4* A = reinterpret_cast<uint4*>(R);
  // faces in batch order: +x,-x,+y,-y,+z,-z = blocks 22, 4, 16, 10, 14, 12; everything else is an edge or corner block
  unsigned int bl = ((needed >> 22) & 1u) | (((needed >> 4) & 1u) << 1) | (((needed >> 16) & 1u) << 2) | (((needed >> 10) & 1u) << 3) |
                    (((needed >> 14) & 1u) << 4) | (((needed >> 12) & 1u) << 5);
  if (needed & ~((1u << 22) | (1u << 4) | (1u << 16) | (1u << 10) | (1u << 14) | (1u << 12))) bl |= 0xC0u;
  while (bl) {  // group-uniform
    unsigned int hw[4][5];
    unsigned int ent[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      unsigned int e = 0;
      if (bl) {
        const int k = __ffs(bl) - 1;
        bl &= bl - 1;
        e = tab.halo[k][lane64];
        if (!((e >> 25) & 1u) || !((needed >> ((e >> 20) & 31u)) & 1u)) e = 0;
      }
      ent[j] = e;
      if (e) {
        const int slot = row[(e >> 20) & 31u];
        const unsigned int* src =
            reinterpret_cast<const unsigned int*>(X + (size_t)slot * kEsdfBlockBytes) + ((e >> 11) & 511u) * kEsdfVoxelWords;
#pragma unroll
        for (int w = 0; w < 5; w++) hw[j][w] = __ldcg(src + w);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (ent[j]) {
        const int v = ent[j] & 2047u;
        A[v] = make_uint4(hw[j][0], hw[j][1], hw[j][2], hw[j][3]);
        R[kFlagBase + v] = hw[j][4];
      }
    }
  }
}

// ---- the six passes of updateLocalNeighborBands (:1323-1386) restricted to the pairs that matter, updateSingleNeighbor
// (:602-633) per pair. Inside one pass sources and destinations are disjoint planes, so its pairs are independent.
// The loops are NOT unrolled and the pass is a run-time value: a candidate executes this code once, and straight-line code
// that is executed once is fetched from L2 (the first version, six unrolled passes, spent a quarter of its samples waiting for
// instructions).
__device__ __forceinline__ bool replayX(const XTables& tab, unsigned int* R, const LiveMasks& L, int lane64, int group,
                                        float max_sq) {
  uint4* A = reinterpret_cast<uint4*>(R);
  const unsigned int* live_p = L.live;
  bool changed = false;
#pragma unroll 1
  for (int pass = 0; pass < 6; pass++) {
    const unsigned int live = live_p[pass];
    if (!live) continue;  // group-uniform
    const int axis = pass >> 1, dir = (pass & 1) ? -1 : 1;
    const int s0 = axis == 0 ? dir : 0, s1 = axis == 1 ? dir : 0, s2 = axis == 2 ? dir : 0;
#pragma unroll 1
    for (int j = 0; j < 4; j++) {
      unsigned int e = tab.pair[pass][j][lane64];
      if (!((e >> 28) & 1u) || !((live >> ((e >> 22) & 31u)) & 1u)) e = 0;
      if (__any_sync(0xffffffffu, e != 0u)) {  // slots 0 and 1 are live or dead for the whole group
        const int sv = e & 2047u, dv = (e >> 11) & 2047u;  // (voxel 0 for a dead lane: any valid address)
        const uint4 es = A[sv], ns = A[dv];
        const unsigned int ef = R[kFlagBase + sv], nf = R[kFlagBase + dv];
        const bool ok = e != 0u && flagObserved(ef) && flagObserved(nf) && !flagSite(nf) && !(__uint_as_float(es.x) >= max_sq);
        const int d0 = (int)es.y - s0, d1 = (int)es.z - s1, d2 = (int)es.w - s2;
        const float pdist = (float)(d0 * d0 + (d1 * d1 + d2 * d2));
        if (ok && __uint_as_float(ns.x) > pdist) {
          A[dv] = make_uint4(__float_as_uint(pdist), (unsigned)d0, (unsigned)d1, (unsigned)d2);
          changed = changed || ((e >> 27) & 1u);
        }
      }
    }
    groupSync(group);
  }
  return changed;
}

// ---- sweepSingleBand (:542-600) for one line of the block, on registers. AXIS is the line's direction; (a, b) are the two
// other voxel coordinates in axis order. Written for few instructions on a short loop-carried chain (a block's sweep is
// 3 axes x 16 sequential steps on two warps: what it costs is instructions x issue latency):
//   * the running site of the scan is kept RELATIVE TO THE LINE: la = its offset along the line, (lo1, lo2) = its two
//     perpendicular components (constant along the line), lp2 = lo1^2 + lo2^2; its squared distance to the voxel at `pos`
//     is (la - pos)^2 + lp2 and that voxel's new parent (la - pos, lo1, lo2);
//   * squared distances are exact integers (or max_sq), so "sq > n" is the integer test ceil(sq) > n;
//   * "no site seen yet" is lp2 = 2^28 (never closer than anything), "voxel cannot be improved" (unobserved or a site) is
//     ceil(sq) := INT_MIN, so the reference's test `found && observed && !site && sq > d` is ONE compare;
//   * a voxel that offers a site to the scan -- a site itself (offer = its own position: its parent registers are zeroed) or an
//     observed voxel with a valid distance (offer = its parent) -- has its bit in `tk`; the reference's four cases become
//     `improve` and `take = tk && !improve`.
// Chain per voxel: subtract, multiply-add, compare, predicate, select. Forward pass, then backward over the updated registers.
// Measured alone on one SM (profiles/microbench_sweep.cu): 3 084 cycles per block (930 per axis, ~380 instructions per warp);
// a single-copy version (run-time axis, rolled passes on a reversed register image) takes 5 633, the 20-byte-voxel version
// with (l0, l1, l2) sites of round 1 ~4 750; with all eight groups of an SM sweeping at once: 5 570.
template <int AXIS>
__device__ __forceinline__ bool sweepLineX(unsigned int* R, int a, int b, float max_sq) {
  uint4* A = reinterpret_cast<uint4*>(R);
  const int v0 = AXIS == 0 ? rvox(1, a + 1, b + 1) : (AXIS == 1 ? rvox(a + 1, 1, b + 1) : rvox(a + 1, b + 1, 1));
  constexpr int stride = AXIS == 0 ? kRX : (AXIS == 1 ? kRY : 1);
  constexpr int kNone = 1 << 28, kCap = 1 << 27;
  int T[kVps], pa[kVps], po1[kVps], po2[kVps], pp2[kVps];
  unsigned int tk = 0, dirty = 0;
#pragma unroll
  for (int i = 0; i < kVps; i++) {
    const uint4 q = A[v0 + i * stride];
    const unsigned int fl = R[kFlagBase + v0 + i * stride];
    const float sq = __uint_as_float(q.x);
    const bool o = flagObserved(fl), st = flagSite(fl);
    const int t = min(__float2int_ru(sq), kCap);
    T[i] = (o && !st) ? t : INT_MIN;
    const int qa = (int)(AXIS == 0 ? q.y : (AXIS == 1 ? q.z : q.w));
    const int q1 = (int)(AXIS == 0 ? q.z : q.y);
    const int q2 = (int)(AXIS == 2 ? q.z : q.w);
    pa[i] = st ? 0 : qa, po1[i] = st ? 0 : q1, po2[i] = st ? 0 : q2;
    pp2[i] = po1[i] * po1[i] + po2[i] * po2[i];
    if (o && (st || sq < max_sq)) tk |= 1u << i;
  }
#pragma unroll
  for (int pass = 0; pass < 2; pass++) {
    int la = 0, lo1 = 0, lo2 = 0, lp2 = kNone;
#pragma unroll
    for (int kk = 0; kk < kVps; kk++) {
      const int k = pass ? (kVps - 1 - kk) : kk;  // line position
      const int t = la - k;
      const int pd = t * t + lp2;
      const bool improve = T[k] > pd;  // the running site is closer than the voxel's value
      const bool take = ((tk >> k) & 1u) && !improve;
      const int offer_a = pa[k] + k;
      if (improve) {
        T[k] = pd, pa[k] = t, po1[k] = lo1, po2[k] = lo2, pp2[k] = lp2;
        dirty |= 1u << k;  // (pd < old sq <= max_sq: the voxel now has a valid distance ...
      }
      la = take ? offer_a : la;
      lo1 = take ? po1[k] : lo1;
      lo2 = take ? po2[k] : lo2;
      lp2 = take ? pp2[k] : lp2;
    }
    tk |= dirty;  // ... and offers its parent to the backward scan)
  }
#pragma unroll
  for (int i = 0; i < kVps; i++) {
    if ((dirty >> i) & 1u) {
      const unsigned int x = (unsigned)(AXIS == 0 ? pa[i] : po1[i]);
      const unsigned int y = (unsigned)(AXIS == 0 ? po1[i] : (AXIS == 1 ? pa[i] : po2[i]));
      const unsigned int z = (unsigned)(AXIS == 2 ? pa[i] : po2[i]);
      A[v0 + i * stride] = make_uint4(__float_as_uint((float)T[i]), x, y, z);
    }
  }
  return dirty != 0;
}
// sweepBlockBandKernel (:1390-1431) for the block in the region: x lines, y lines, z lines.
__device__ __forceinline__ bool sweepBlockX(unsigned int* R, int group, int lane64, float max_sq) {
  const int a = lane64 >> 3, b = lane64 & 7;
  bool ch = sweepLineX<0>(R, a, b, max_sq);
  groupSync(group);
  ch |= sweepLineX<1>(R, a, b, max_sq);
  groupSync(group);
  ch |= sweepLineX<2>(R, a, b, max_sq);
  return ch;
}

// Registration of the face neighbours of a block that is a member of ring `target` as candidates of that ring, in three
// steps so that the atomic and the neighbours' rows travel while the block is swept (nothing consumes their results before
// registerFinish):
//   begin : atomicExch on the candidates' stamps (unique registration) + prefetch of their 3x3x3 rows,
//   claim : positions in this CTA's segment for the registrations this group won (shared-memory counter),
//   finish: write the records {slot, row}.
// xs.nb[group][0..5] holds the six face neighbours (slot or < 0).
struct RegState {
  int old;    // lanes 0..5: previous stamp of the neighbour
  int pos;    // lanes 0..5: position won in the CTA's segment, -1: somebody else registered it
  int v[3];   // row entries (w, d) = ((lane + 64 k) / 27, (lane + 64 k) % 27) of the six neighbours
};
__device__ __forceinline__ RegState registerBegin(const XCtx& c, XShared& xs, int group, int lane64, int target) {
  RegState s;
  s.old = target, s.pos = -1;
  if (lane64 < 6) {
    const int nb = xs.nb[group][lane64];
    if (nb >= 0) s.old = atomicExch(c.cand_stamp + nb, target);
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const int idx = lane64 + 64 * k;
    s.v[k] = -1;
    if (idx < 162) {
      const int nb = xs.nb[group][idx / 27];
      if (nb >= 0) s.v[k] = __ldcg(c.nbr27 + 27 * nb + idx % 27);  // consumed in registerFinish: stays in flight
    }
  }
  return s;
}
__device__ __forceinline__ void registerClaim(XShared& xs, RegState& s, int group, int lane64, int target) {
  if (lane64 < 32) {  // the group's first warp
    const bool win = lane64 < 6 && xs.nb[group][lane64] >= 0 && s.old != target;
    const unsigned int ballot = __ballot_sync(0xffffffffu, win);
    if (ballot) {
      int base = 0;
      if (lane64 == 0) base = atomicAdd(&xs.ncand, __popc(ballot));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (win) s.pos = base + __popc(ballot & ((1u << lane64) - 1u));
    }
  }
}
__device__ __forceinline__ void registerFinish(const XCtx& c, XShared& xs, const RegState& s, int group, int lane64,
                                               int* segment) {
  if (lane64 < 6) xs.pos[group][lane64] = s.pos;
  groupSync(group);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const int idx = lane64 + 64 * k;
    if (idx < 162) {
      const int w = idx / 27, pos = xs.pos[group][w];
      if (pos >= 0) {
        int v = s.v[k];
        if (v < -1) v = neighbor27(c, xs.nb[group][w], idx % 27);  // never linked: resolve through the hash once
        __stcg(segment + (size_t)pos * kRecInts + 1 + idx % 27, v);
      }
    }
  }
  if (lane64 < 6 && s.pos >= 0) __stcg(segment + (size_t)s.pos * kRecInts, xs.nb[group][lane64]);
}

struct XState {
  int ring, ci;
};
// One candidate of ring `st.ring`: gather, replay, and if it changed: sweep, publish, register its neighbours for ring+1.
__device__ __noinline__ void processCandidate(const XCtx& c, const XTables& tab, XShared& xs, unsigned int* R, int ring,
                                              int seg_cta, int seg_idx, int cta, int group, int lane64, long long* prof) {
  const int ci = ring & 1, ni = ci ^ 1;
  X_PROF_BEGIN()
  if (lane64 < 28) xs.rec[group][lane64] = __ldcg(c.segment(ci, seg_cta) + (size_t)seg_idx * kRecInts + lane64);
  groupSync(group);
  X_PROF(0)
  const int slot = xs.rec[group][0];
  const int* row = &xs.rec[group][1];
  // membership of the 27 blocks in this ring (sources of the passes); the loads travel with the loads of the own block
  int sv = ring - 1;
  if (lane64 < 27 && row[lane64] >= 0) sv = __ldcg(c.stamp[ci] + row[lane64]);
  OwnRegs own = ownLoad(c.blocks + (size_t)slot * kEsdfBlockBytes, lane64);
  const unsigned int m = __ballot_sync(0xffffffffu, lane64 < 27 && sv == ring);
  if (lane64 == 0) xs.mask[group] = m, xs.changed[group] = 0;
  ownToShared(R, own, lane64);
  groupSync(group);
  const LiveMasks L = liveMasks(xs.mask[group], xs.live[group], lane64 == 0);
  X_PROF(1)
  haloGather(tab, R, row, L.needed, c.X[ci], lane64);
  groupSync(group);
  X_PROF(2)
  const bool ch = replayX(tab, R, L, lane64, group, c.max_sq);
  if (ch) xs.changed[group] = 1;
  groupSync(group);
  X_PROF(3)
  X_PROF_COUNT(6)
  if (xs.changed[group]) {
    // B is a member of ring+1
    if (lane64 < 6) xs.nb[group][lane64] = row[faceEntry(lane64)];
    if (lane64 == 0) {
      __stcg(c.stamp[ni] + slot, ring + 1);
      atomicAdd(&xs.nchanged, 1);
    }
    groupSync(group);
    RegState rs = registerBegin(c, xs, group, lane64, ring + 1);
    sweepBlockX(R, group, lane64, c.max_sq);
    registerClaim(xs, rs, group, lane64, ring + 1);
    groupSync(group);
    X_PROF(4)
    ownStore(c.blocks + (size_t)slot * kEsdfBlockBytes, true, c.X[ni] + (size_t)slot * kEsdfBlockBytes, R, c.psum + 2 * (size_t)slot, lane64);
    registerFinish(c, xs, rs, group, lane64, c.segment(ni, cta));
    X_PROF_COUNT(7)
  }
  groupSync(group);
  X_PROF(5)
}

// A member of the initial list of a computeEsdf call (ring `st.ring`): sweep in place, publish, register its neighbours as
// the candidates of this ring.
__device__ __noinline__ void processSeed(const XCtx& c, XShared& xs, unsigned int* R, int ring, int slot, int cta, int group,
                                         int lane64) {
  const int ci = ring & 1;
  unsigned char* blk = c.blocks + (size_t)slot * kEsdfBlockBytes;
  OwnRegs own = ownLoad(blk, lane64);
  if (lane64 < 6) xs.nb[group][lane64] = neighbor6(c, slot, lane64);
  if (lane64 == 0) {
    __stcg(c.stamp[ci] + slot, ring);
    xs.changed[group] = 0;
  }
  groupSync(group);
  RegState rs = registerBegin(c, xs, group, lane64, ring);
  ownToShared(R, own, lane64);
  groupSync(group);
  const bool ch = sweepBlockX(R, group, lane64, c.max_sq);
  registerClaim(xs, rs, group, lane64, ring);
  if (ch) xs.changed[group] = 1;
  groupSync(group);
  // (an unchanged block keeps its parent box)
  ownStore(blk, xs.changed[group] != 0, c.X[ci] + (size_t)slot * kEsdfBlockBytes, R, xs.changed[group] ? c.psum + 2 * (size_t)slot : nullptr, lane64);
  registerFinish(c, xs, rs, group, lane64, c.segment(ci, cta));
  groupSync(group);
}

// Grid barrier + all-gather of the per-CTA counts {registrations, changed blocks}; leaves the exclusive scan of the
// registrations in xs.pre and returns the totals. Two implementations, measured on 148 x 512 threads
// (profiles/microbench_xbarrier.cu, cycles per barrier incl. the scan, for 1 / 8 / 32 / 64 / 148 participating CTAs):
//   flags  : 2 688 / 2 772 / 3 454 / 3 795 / 5 764 -- every CTA owns one 64-bit flag {generation, registrations, changed} per
//            barrier parity (one per 128-byte line); arrival = fence + plain store, waiting = thread t polls the flag of CTA t
//            and then holds its counts. No read-modify-write on a shared word, but 148 x 148 polls. (Kept in the microbenchmark.)
//   counter: 3 573 / 3 687 / 3 712 / 3 723 / 3 884 -- counts stored to per-CTA slots (two arrays by barrier parity), fence,
//            atomic arrival counter polled by thread 0, fence, then one read of the 148 slots. Used here.
__device__ __forceinline__ void counterBarrierScan(XShared& xs, unsigned int* bar, unsigned int& generation, int2* counts, int nctas,
                                                   int cta, int tid, int* K, int* M) {
  __syncthreads();
  if (tid == 0) {
    __stcg(counts + cta, make_int2(xs.ncand, xs.nchanged));
    xs.ncand = 0, xs.nchanged = 0;
  }
  gridBarrier(bar, generation, nctas);
  int2 v = make_int2(0, 0);
  if (tid < nctas) v = __ldcg(counts + tid);
  const int lane = tid & 31, warp = tid >> 5;
  int inc = v.x, chg = v.y;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) chg += __shfl_xor_sync(0xffffffffu, chg, o);
  if (warp < 8 && lane == 31) xs.warp_tot[0][warp] = inc;
  if (warp < 8 && lane == 0) xs.warp_tot[1][warp] = chg;
  __syncthreads();
  int base = 0, mtot = 0;
#pragma unroll
  for (int w = 0; w < 8; w++) {
    if (w < warp) base += xs.warp_tot[0][w];
    if (w * 32 < nctas) mtot += xs.warp_tot[1][w];
  }
  if (tid < nctas) xs.pre[tid + 1] = base + inc;
  if (tid == 0) xs.pre[0] = 0, xs.next = 0;
  __syncthreads();
  *K = xs.pre[nctas];
  *M = mtot;
}

__global__ void __maxnreg__(NVB_WAVEX_MAXREG) esdfWaveXKernel(EsdfCtx c) {
  extern __shared__ __align__(16) unsigned int smem[];
  __shared__ XTables tab;
  __shared__ XShared xs;
  __shared__ XCtx xc;
  const int cta = blockIdx.x, nctas = gridDim.x;
  const int tid = threadIdx.x, group = tid >> 6, lane64 = tid & 63;
  // Empty block list: integrateBlocksTemplate returns before touching anything (:226-228).
  if (*(volatile int*)c.work_count == 0) return;
  initTables(tab, tid);
  if (tid == 0) {
    xs.ncand = 0, xs.nchanged = 0, xs.next = 0;
    xc.blocks = c.esdf.blocks, xc.block_index = c.esdf.block_index, xc.hash = c.esdf.hash;
    xc.nbr = c.nbr, xc.nbr27 = c.nbr27, xc.cand_stamp = c.cand_stamp, xc.psum = c.psum;
    xc.stamp[0] = c.stamp_a, xc.stamp[1] = c.stamp_b;
    xc.X[0] = c.xslab, xc.X[1] = c.xslab + (size_t)c.esdf.capacity * kEsdfBlockBytes;
    xc.recs[0] = c.xrec, xc.recs[1] = c.xrec + (size_t)nctas * c.xseg * kRecInts;
    xc.seg = c.xseg, xc.max_sq = c.max_sq;
  }
  __syncthreads();
  XState st;
  st.ring = *(volatile int*)c.ring_id;
  int swept = 0, faces = 0, rings = 0, n_bar = 0, n_tail = 0;
#if NVB_WAVEX_PROF
  long long t_bar = 0, t_work = 0, t0 = globalTimerNs(), t1;
#endif
  unsigned int* R = smem + group * kXRegionWords;
#if NVB_WAVEX_PROF
  long long prof_store[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long* prof = prof_store;
#else
  long long* prof = nullptr;
#endif
  // prof: this group, cycles: record fetch, stamps + own block, halo, replay, sweep, stores + records; candidates, changed
#if NVB_WAVEX_PROF
  int dbgK = 0, dbgM = 0;
#define X_DBG(k, m) dbgK = (k), dbgM = (m);
#else
#define X_DBG(k, m)
#endif
#if NVB_WAVEX_PROF
#define X_TIME_WORK()                                                                                               \
  t1 = globalTimerNs(), t_work += t1 - t0;                                                                          \
  if (tid == 0 && n_bar < 1000) atomicMax((unsigned long long*)c.phase_max + n_bar, (unsigned long long)(t1 - t0)); \
  if (cta == 0 && tid == 0 && n_bar < 1000) c.phase_max[1000 + n_bar] = dbgK, c.phase_max[2000 + n_bar] = dbgM, c.phase_max[3000 + n_bar] = t1 - t0; \
  t0 = t1;
#define X_TIME_BARRIER() t1 = globalTimerNs(), t_bar += t1 - t0, t0 = t1;
#else
#define X_TIME_WORK()
#define X_TIME_BARRIER()
#endif
  // Barrier number n_bar of this launch: publishes what this CTA registered / changed in the phase, returns the totals and
  // leaves the exclusive scan of the registrations in xs.pre.
  int2* const count_base = reinterpret_cast<int2*>(c.xcounts);
  unsigned int generation = 0;
#define X_BARRIER(Kout, Mout)                                                                                                    \
  X_TIME_WORK()                                                                                                                  \
  counterBarrierScan(xs, c.barrier, generation, count_base + (n_bar & 1) * kMaxCtas, nctas, cta, tid, &(Kout), &(Mout));             \
  X_TIME_BARRIER()                                                                                                               \
  n_bar++;
  for (int pass = 0; pass < 2; pass++) {
    // pass 0: blocks with sites; pass 1: the persistent cleared list (:254-257)
    const int* src = pass ? c.cleared_list : c.upd_list;
    const int n0 = pass ? *(volatile int*)c.cleared_count : *(volatile int*)c.upd_count;
    if (n0 == 0) continue;
    st.ci = st.ring & 1;
    // ---- seeds: the call's block list is ring `ring`
    X_DBG(-1, n0)
    for (;;) {
      if (lane64 == 0) xs.cur[group] = atomicAdd(&xs.next, 1);
      groupSync(group);
      const long long e = (long long)cta + (long long)xs.cur[group] * nctas;
      if (e >= n0) break;
      processSeed(xc, xs, R, st.ring, __ldcg(src + e), cta, group, lane64);
    }
    int M = n0, K, unused;
    X_BARRIER(K, unused)  // K: candidates of ring `ring`
    swept += n0;
    while (true) {
      X_DBG(K, M)
      if (NVB_WAVEX_TAIL && K <= kXG) {
        // ---- tail: the ring fits one CTA. CTA 0 runs whole rings with block-level barriers until the wavefront dies out
        // or outgrows it (its registrations all land in its own segment); everybody else waits at ONE grid barrier.
        if (cta == 0) {
          int t = 0;
          if (group < K) {  // where the (at most 8) entries of the current ring live
            for (int s = lane64; s < nctas; s += 64)
              if (xs.pre[s] <= group && group < xs.pre[s + 1]) xs.seg[group][0] = s, xs.seg[group][1] = group - xs.pre[s];
          }
          __syncthreads();
          while (true) {
            if (group < K) processCandidate(xc, tab, xs, R, st.ring, xs.seg[group][0], xs.seg[group][1], 0, group, lane64, prof);
            __syncthreads();
            const int K2 = xs.ncand, M2 = xs.nchanged;
            __syncthreads();
            if (tid == 0) xs.ncand = 0, xs.nchanged = 0;
            if (tid < kXG) xs.seg[tid][0] = 0, xs.seg[tid][1] = tid;  // the next ring's entries: segment 0, in order
            faces += 6 * M, rings++, swept += M2, t++;
            st.ring++, st.ci ^= 1;
            M = M2, K = K2;
            __syncthreads();
            if (M == 0 || K > kXG) break;
          }
          n_tail += t;
          if (tid == 0) c.xtail[0] = t, c.xtail[1] = K, c.xtail[2] = M;
        }
        {
          int k_unused, m_unused;  // (the tail's counts travel through xtail: CTA 0 alone registered)
          X_BARRIER(k_unused, m_unused)
        }
        if (cta != 0) {
          if (tid == 0) xs.bcast[0] = __ldcg(c.xtail + 0), xs.bcast[1] = __ldcg(c.xtail + 1), xs.bcast[2] = __ldcg(c.xtail + 2);
          __syncthreads();
          const int t = xs.bcast[0];
          K = xs.bcast[1], M = xs.bcast[2];
          __syncthreads();
          st.ring += t, st.ci ^= (t & 1);
        }
        if (M == 0) break;
        // the K candidates of the current ring were all registered by CTA 0
        for (int s = tid; s <= nctas; s += kXT) xs.pre[s] = s == 0 ? 0 : K;
        __syncthreads();
        continue;
      }
      // ---- grid ring: candidates dealt round-robin over CTAs, then over the CTA's groups
      for (;;) {
        if (lane64 == 0) xs.cur[group] = atomicAdd(&xs.next, 1);
        groupSync(group);
        const long long e = (long long)cta + (long long)xs.cur[group] * nctas;
        if (e >= K) break;
        for (int s = lane64; s < nctas; s += 64)
          if (xs.pre[s] <= e && e < xs.pre[s + 1]) xs.seg[group][0] = s, xs.seg[group][1] = (int)e - xs.pre[s];
        groupSync(group);
        processCandidate(xc, tab, xs, R, st.ring, xs.seg[group][0], xs.seg[group][1], cta, group, lane64, prof);
      }
      int K2, M2;
      X_BARRIER(K2, M2)
      faces += 6 * M, rings++, swept += M2;
      st.ring++, st.ci ^= 1;
      M = M2, K = K2;
      if (M == 0) break;
    }
    st.ring++;  // the next computeEsdf call's stamps must not alias this one's
  }
#undef X_BARRIER
#undef X_TIME_WORK
#undef X_TIME_BARRIER
#undef X_DBG
  if (cta == 0 && tid == 0) {
    *c.ring_id = st.ring + 1;
    c.stats[4] = *(volatile int*)c.cleared_count;
    c.stats[5] = swept, c.stats[6] = faces, c.stats[7] = rings;
    c.stats[10] = n_tail, c.stats[11] = n_bar;
#if NVB_WAVEX_PROF
    c.stats[8] = t_bar, c.stats[9] = t_work;
    long long sum_max = 0;
    for (int q = 0; q < n_bar && q < 1000; q++) sum_max += (long long)c.phase_max[q];
    c.stats[12] = sum_max;
#endif
#if NVB_WAVEX_PROF
    for (int q = 0; q < 8; q++) c.phase_max[3990 + q] = prof[q];
#endif
  }
}

}  // namespace

int esdfWaveXMaxCtas() { return kMaxCtas; }
size_t esdfWaveXFlagBytes() { return 2 * (size_t)kMaxCtas * sizeof(int2); }

cudaError_t launchEsdfComputeX(const EsdfCtx& c, int num_sms, int reserved_sms, cudaStream_t stream, int* launches) {
  static int per_sm = -1;
  if (per_sm < 0) {
    cudaFuncSetAttribute(esdfWaveXKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kXSmemBytes);
    int v = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, esdfWaveXKernel, kXT, kXSmemBytes) != cudaSuccess) v = 0;
    per_sm = v;
  }
  if (per_sm <= 0) return cudaErrorLaunchOutOfResources;
  EsdfCtx cc = c;
  void* args[] = {&cc};
  (*launches)++;
  // `reserved_sms` SMs are left to the other resident kernels: the raycast / compaction / TSDF kernels of the next frame then
  // run there instead of stealing issue slots from ring-critical CTAs (measured on the 80-frame C2 bench,
  // profiles/r2_run9.sh: 0 / 2 / 4 / 8 / 16 reserved -> 3069 / 3213 / 3194 / 3198 / 3151 frames/s), and a multi-GPU
  // rank's NCCL all-gather can start -- and wait for its peers -- while a wavefront is in flight (a cooperative grid
  // that fills every SM serialises the two: 0.30 ms per merge with 0 reserved, 0.10 ms with 4, profiles/r2_run8.sh).
  int grid = num_sms < kMaxCtas ? num_sms : kMaxCtas;
  if (reserved_sms > 0 && grid - reserved_sms >= 8) grid -= reserved_sms;
  return cudaLaunchCooperativeKernel((const void*)esdfWaveXKernel, dim3(grid), dim3(kXT), args, kXSmemBytes, stream);
}

}  // namespace nvb
