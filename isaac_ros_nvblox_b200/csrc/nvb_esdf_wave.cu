// nvb_esdf_wave.cu -- the ESDF wavefront (computeEsdf, nvblox/src/integrators/esdf_integrator.cu:1465-1496)
// as ONE cooperative persistent launch.
//
// computeEsdf(list): sweep(list); while (list not empty) { list = updateNeighborBands(list); sweep(list); }
// runs twice per update: for the blocks with sites and for the persistent "cleared" list (:254-257).
// The result depends on the order of the passes (x,y,z in-block sweeps; +x,-x,+y,-y,+z,-z face propagation,
// each seeing the previous ones), so that order is kept: per ring three axis phases and one sweep phase,
// separated by grid barriers.
//
// What a phase costs is the SLOWEST CTA's dependent chain plus a 1.25 us barrier (measured,
// profiles/README.md), ~18 rings x 4 phases per frame. The design therefore minimises round trips and
// imbalance per phase, not bytes:
//   * Ring membership is a per-slot stamp (stamp[r & 1][slot] == r): the face operations test "is my
//     neighbour a member of this ring" with one load that travels with the voxel loads.
//   * Ring r+1's member list is built unique with atomicExch on that stamp. Updated blocks are first recorded
//     in shared memory during the three axis phases and appended once per ring (two dependent L2 atomics per
//     ring instead of per phase); a warp reserves its list range with one atomicAdd. No sort/unique launch.
//   * Members are dealt round-robin from the list (CTA c takes entries c, c+G, ...): every CTA gets
//     ceil(n/G) blocks. (Static slot ownership was tried: max/mean 2.5 and the slowest CTA sets the pace.)
//     A CTA keeps its members and their six neighbour slots (nbr table, built at allocation) in shared
//     memory for the three axis phases: an axis phase is ONE L2 round trip (face voxels + neighbour stamp).
//   * The +dir/-dir passes of an axis are fused per block interface (exact, see axisMembers): 3 phases, not 6;
//     one warp per interface, two face voxels per lane, no block-level synchronisation inside the phase.
//   * A block's three-axis sweep is a ~2 000-instruction dependent chain on two warps; it runs on registers
//     with an integer, branch-free formulation (sweepLineRegs) -- 10 us -> ~2.5 us per block.
//   * 256-thread CTAs with up to 128 registers (half an SM's register file) so that the next frame's
//     raycast / compaction / TSDF kernels co-reside while the wavefront runs on its side stream.
#include "nvb_esdf_common.cuh"

namespace nvb {

namespace {

#ifndef NVB_WAVE_INLINE
#define NVB_WAVE_INLINE 1
#endif
#if NVB_WAVE_INLINE
#define NVB_WAVE_FN __forceinline__
#else
#define NVB_WAVE_FN __noinline__
#endif
#ifndef NVB_WAVE_THREADS
#define NVB_WAVE_THREADS 256
#endif
constexpr int kWT = NVB_WAVE_THREADS;   // threads per CTA. Measured (profiles/wave_variants.sh): 256 and 512 threads are equally fast
                                        // (the phase time is one block's dependent chain), 1024 threads / 64 registers spills and is 50 % slower;
                                        // 256 x 128 registers = half of an SM's register file, so the next frame's kernels co-reside.
constexpr int kWG = kWT / 64;           // groups of 64 threads
constexpr int kWaveMaxMembers = 1024;   // owned candidates scanned per round
constexpr int kPendMax = 1024;         // pending "updated block" records per CTA per ring before a forced flush
constexpr int kNbrCache = 128;          // members whose neighbour slots are cached in smem
// Shared-memory image of an ESDF block for the sweeps: the 20-byte AoS voxels with ONE pad word after
// every row of 8 voxels: word(v, f) = 5 v + f + (v >> 3). With this pitch the y- and z-line accesses of a
// warp (32 lines) hit 32 distinct banks and x-lines are 2-way (the unpadded copy is 4-way / 8-way), which
// matters because all 16 groups of the CTA share one shared-memory pipe.
constexpr int kPadBlockWords = kBlockWords + kVpb / kVps;  // 2560 + 64
#ifndef NVB_WAVE_PAD
#define NVB_WAVE_PAD 0  // measured: the padded image is 30 % slower (scalar smem stores); see profiles/README.md
#endif
constexpr int kSweepBlockWords = NVB_WAVE_PAD ? kPadBlockWords : kBlockWords;
constexpr size_t kWaveSmemBytes = (size_t)kWG * kSweepBlockWords * sizeof(unsigned int);  // sweep buffers

// HBM -> padded smem image. Global side: 128-bit coalesced loads (640 chunks per block). Shared side:
// the pad makes chunk destinations unaligned, so each chunk is stored as four 32-bit words; lane groups
// of 8 rotate which word they store so that the 32 lanes of one store instruction hit 32 banks.
__device__ __forceinline__ unsigned int pick(const uint4& q, int j) {
  return j == 0 ? q.x : (j == 1 ? q.y : (j == 2 ? q.z : q.w));
}
__device__ __forceinline__ void loadBlockPadded(unsigned int* sm, const unsigned int* g, int lane64) {
  const uint4* src = reinterpret_cast<const uint4*>(g);
  const int rot = (lane64 >> 3) & 3;
  constexpr int kBatch = 5;  // 5 chunks (20 registers) in flight per thread, twice
#pragma unroll 1
  for (int k0 = 0; k0 < kBlockWords / 4 / 64; k0 += kBatch) {
    uint4 q[kBatch];
#pragma unroll
    for (int k = 0; k < kBatch; k++) q[k] = __ldcg(src + lane64 + (k0 + k) * 64);
#pragma unroll
    for (int k = 0; k < kBatch; k++) {
      const int c = lane64 + (k0 + k) * 64;
      unsigned int* dst = sm + 4 * c + c / 10;  // row = (4c) / 40
#pragma unroll
      for (int t = 0; t < 4; t++) {
        const int j = (t + rot) & 3;
        dst[j] = pick(q[k], j);
      }
    }
  }
}
__device__ __forceinline__ void storeBlockPadded(unsigned int* g, const unsigned int* sm, int lane64) {
  uint4* dst = reinterpret_cast<uint4*>(g);
  const int rot = (lane64 >> 3) & 3;
#pragma unroll
  for (int k = 0; k < kBlockWords / 4 / 64; k++) {
    const int c = lane64 + k * 64;
    const unsigned int* src = sm + 4 * c + c / 10;
    unsigned int w0 = 0, w1 = 0, w2 = 0, w3 = 0;
#pragma unroll
    for (int t = 0; t < 4; t++) {
      const int j = (t + rot) & 3;
      const unsigned int val = src[j];
      w0 = (j == 0) ? val : w0, w1 = (j == 1) ? val : w1, w2 = (j == 2) ? val : w2, w3 = (j == 3) ? val : w3;
    }
    __stcg(dst + c, make_uint4(w0, w1, w2, w3));
  }
}

__device__ __forceinline__ int resolveNeighbor(const EsdfCtx& c, int slot, int dir) {
  int v = __ldcg(c.nbr + 6 * slot + dir);
  if (v < -1) {  // unknown (block created outside the ESDF update path): resolve through the hash once
    const int* bi = c.esdf.block_index + 3 * slot;
    int x = bi[0], y = bi[1], z = bi[2];
    const int d = (dir & 1) ? -1 : 1;
    if ((dir >> 1) == 0) x += d;
    else if ((dir >> 1) == 1) y += d;
    else z += d;
    v = hashFind(c.esdf.hash, x, y, z);
    c.nbr[6 * slot + dir] = v;
  }
  return v;
}

struct WaveShared {
  int members[kWaveMaxMembers];
  int nbr[kNbrCache * 6];
  int scan[kWT / 32];
  int count;
  int changed[kWG];
  int slot[kWG * 2];
  int upd[kWG * 2];
  int npend;
  int pend[kPendMax];  // blocks updated by this CTA's face operations in the current ring (with duplicates)
};

// Members of a ring are dealt round-robin over the CTAs from the ring's global list: CTA c takes entries
// c, c+G, c+2G, ... so every CTA gets ceil(n/G) or floor(n/G) blocks (the static slot-ownership scheme this
// replaces had max/mean of 2.5, and the slowest CTA is what a phase costs). Entries [first, first+cap) of this
// CTA's share are cached in shared memory. Optionally stamps them (initial list of a computeEsdf call).
__device__ NVB_WAVE_FN int loadMembers(WaveShared& sh, const int* list, int n, int cta, int nctas, int first,
                                       int* stamp_out, int stamp_value) {
  const int tid = threadIdx.x;
  const int mine = (n > cta) ? (n - cta + nctas - 1) / nctas : 0;  // entries of this CTA
  int k = mine - first;
  k = k < 0 ? 0 : (k > kWaveMaxMembers ? kWaveMaxMembers : k);
  for (int j = tid; j < k; j += kWT) {
    const int slot = __ldcg(list + cta + (first + j) * nctas);
    sh.members[j] = slot;
    if (stamp_out) stamp_out[slot] = stamp_value;
  }
  __syncthreads();
  return k;
}

// Neighbour slots of the first kNbrCache members -> shared memory (one thread per (member, dir)).
__device__ __forceinline__ void prefetchNeighbors(const EsdfCtx& c, WaveShared& sh, int k) {
  const int tid = threadIdx.x;
  const int kc = k < kNbrCache ? k : kNbrCache;
  for (int q = tid; q < kc * 6; q += kWT) sh.nbr[q] = resolveNeighbor(c, sh.members[q / 6], q % 6);
}

// sweepSingleBand (:542-600) on registers, written for a SHORT DEPENDENT CHAIN: one block's sweep is
// 3 axes x 16 sequential steps on two warps, so its latency is (instructions on the chain) x (ALU
// latency), not throughput. Per step the only loop-carried state is the candidate site (l0,l1,l2) and
// `found`; everything that does not depend on it is hoisted:
//   * squared distances are exact integers (or max_sq), so "sq > |d|^2" is evaluated as the integer test
//     ceil(sq) > |d|^2 -- no int->float conversion on the chain; "sq < max_sq" becomes a bit mask;
//   * each voxel's own parent position (parent + voxel) is precomputed;
//   * the four cases of the reference (unobserved / site / first valid voxel / candidate vs own) are
//     folded into branch-free selects.
// The line's 8 voxels are loaded once, walked forward, the register image is reversed and walked again
// (= the backward pass); changed voxels are written back. `axis` and `pass` are run-time values so that
// ONE copy of the 8-step body serves all six passes.
// `sm` is the block image in shared memory, v0 the line's first voxel, `stride` the voxel stride along the
// line; (c0,c1,c2) are the voxel coordinates at position 0.
__device__ __forceinline__ bool sweepLineRegs(unsigned int* sm, int v0, int stride, int c0, int c1, int c2, int axis,
                                              float max_sq) {
  int T[kVps];  // ceil(squared distance): sq > n  <=>  T > n for every integer n
  int p0[kVps], p1[kVps], p2[kVps];
  unsigned int obs = 0, site = 0, valid = 0, dirty = 0;
#pragma unroll
  for (int i = 0; i < kVps; i++) {
    const int v = v0 + i * stride;
    const unsigned int* e = sm + v * kEsdfVoxelWords + (NVB_WAVE_PAD ? (v >> 3) : 0);
    const float sq = __uint_as_float(e[0]);
    T[i] = __float2int_ru(sq);
    p0[i] = (int)e[1], p1[i] = (int)e[2], p2[i] = (int)e[3];
    const unsigned int fl = e[4];
    if (flagObserved(fl)) obs |= 1u << i;
    if (flagSite(fl)) site |= 1u << i;
    if (sq < max_sq) valid |= 1u << i;
  }
  const int a0 = (axis == 0), a1 = (axis == 1), a2 = (axis == 2);
#pragma unroll 1
  for (int pass = 0; pass < 2; pass++) {
    int l0 = 0, l1 = 0, l2 = 0;
    bool found = false;
#pragma unroll
    for (int k = 0; k < kVps; k++) {
      // register slot k holds line position pos = k (pass 0) or 7 - k (pass 1, image reversed)
      const int pos = pass ? (kVps - 1 - k) : k;
      const int x0 = c0 + a0 * pos, x1 = c1 + a1 * pos, x2 = c2 + a2 * pos;  // voxel coordinates
      const bool o = (obs >> k) & 1u, st = (site >> k) & 1u, vl = (valid >> k) & 1u;
      const int own0 = p0[k] + x0, own1 = p1[k] + x1, own2 = p2[k] + x2;    // off the chain
      const int d0 = l0 - x0, d1 = l1 - x1, d2 = l2 - x2;
      const int pd = d0 * d0 + (d1 * d1 + d2 * d2);
      const bool better = found && o && !st && (T[k] > pd);  // candidate site is closer than the voxel's value
      const bool take_site = o && st;
      const bool take_own = o && !st && !better && vl;       // voxel's own parent becomes the running site
      if (better) {
        p0[k] = d0, p1[k] = d1, p2[k] = d2, T[k] = pd;
        dirty |= 1u << k;
        valid |= 1u << k;  // pd < old sq <= max_sq
      }
      l0 = take_site ? x0 : (take_own ? own0 : l0);
      l1 = take_site ? x1 : (take_own ? own1 : l1);
      l2 = take_site ? x2 : (take_own ? own2 : l2);
      found = found || take_site || take_own;
    }
    // reverse the register image (and the bit masks) for the other direction / back to line order
#pragma unroll
    for (int k = 0; k < kVps / 2; k++) {
      const int r = kVps - 1 - k;
      int ti = T[k]; T[k] = T[r]; T[r] = ti;
      ti = p0[k]; p0[k] = p0[r]; p0[r] = ti;
      ti = p1[k]; p1[k] = p1[r]; p1[r] = ti;
      ti = p2[k]; p2[k] = p2[r]; p2[r] = ti;
    }
    obs = __brev(obs) >> 24, site = __brev(site) >> 24, dirty = __brev(dirty) >> 24, valid = __brev(valid) >> 24;
  }
  // after two reversals slot i is line position i again
#pragma unroll
  for (int i = 0; i < kVps; i++) {
    if ((dirty >> i) & 1u) {
      const int v = v0 + i * stride;
      unsigned int* e = sm + v * kEsdfVoxelWords + (NVB_WAVE_PAD ? (v >> 3) : 0);
      e[0] = __float_as_uint((float)T[i]);
      e[1] = (unsigned)p0[i], e[2] = (unsigned)p1[i], e[3] = (unsigned)p2[i];
    }
  }
  return dirty != 0;
}

// sweepBlockBandKernel (:1390-1431) for the cached members, kWG blocks at a time.
__device__ NVB_WAVE_FN void sweepMembers(const EsdfCtx& c, WaveShared& sh, int k, unsigned int* smem,
                                         bool prefetch_nbr) {
  const int tid = threadIdx.x, group = tid >> 6, lane64 = tid & 63;
  unsigned int* sm = smem + group * kSweepBlockWords;
  const int a = lane64 >> 3, b = lane64 & 7;
  for (int base = 0; base < k; base += kWG) {
    const int item = base + group;
    const int slot = item < k ? sh.members[item] : -1;
    if (lane64 == 0) sh.changed[group] = 0;
    if (slot >= 0) {
      if (NVB_WAVE_PAD) loadBlockPadded(sm, esdfBlockPtr(c.esdf, slot), lane64);
      else loadBlockGroup(sm, esdfBlockPtr(c.esdf, slot), lane64);
    }
    // neighbour slots for the coming axis phases: issued behind the block loads, not in front of them
    if (prefetch_nbr && base == 0) prefetchNeighbors(c, sh, k);
    __syncthreads();
    bool ch = false;
#pragma unroll 1
    for (int axis = 0; axis < 3; axis++) {
      // x lines: (x, a, b); y lines: (a, y, b); z lines: (a, b, z)
      const int v0 = (axis == 0) ? (a * 8 + b) : ((axis == 1) ? (a * 64 + b) : (a * 64 + b * 8));
      const int stride = (axis == 0) ? 64 : ((axis == 1) ? 8 : 1);
      const int c0 = (axis == 0) ? 0 : a;
      const int c1 = (axis == 0) ? a : ((axis == 1) ? 0 : b);
      const int c2 = (axis == 2) ? 0 : b;
      if (slot >= 0) ch |= sweepLineRegs(sm, v0, stride, c0, c1, c2, axis, c.max_sq);
      __syncthreads();
    }
    if (ch) sh.changed[group] = 1;
    __syncthreads();
    if (slot >= 0 && sh.changed[group]) {
      if (NVB_WAVE_PAD) storeBlockPadded(esdfBlockPtr(c.esdf, slot), sm, lane64);
      else storeBlockGroup(esdfBlockPtr(c.esdf, slot), sm, lane64);
    }
    __syncthreads();
  }
}

// The +dir and -dir passes of one axis of updateNeighborBands (:1323-1386), fused per block
// interface (see nvb_esdf.cu phaseNeighbors for the ownership rule and why it is exact):
//   group side 0 ("hi"): interface (b, b+d): P = b -> b+d, then Q = b+d -> b if b+d is a member;
//   group side 1 ("lo"): interface (b-d, b) only when b-d is NOT a member: Q = b -> b-d.
// Destination blocks are stamped for ring+1 with a plain store.
__device__ NVB_WAVE_FN void axisMembers(const EsdfCtx& c, WaveShared& sh, int axis, int k, const int* stamp_cur, int ring,
                                        int* stamp_nxt, int* list_nxt, int* count_nxt) {
  // One warp per (member, side) interface, two face voxels per lane: kWT/32 interfaces = kWT/64 members per
  // iteration, so the handful of members a CTA owns are all in flight at once (one L2 round trip per phase).
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int kWarps = kWT / 32;
  const int entry_in_cta = warp >> 1, side = warp & 1;
  const int strideA = (axis == 0) ? 64 : ((axis == 1) ? 8 : 1);
  for (int base = 0; base < k; base += kWarps / 2) {
    const int item = base + entry_in_cta;
    int mine = -1, other = -1;
    if (item < k) {
      mine = sh.members[item];
      other = item < kNbrCache ? sh.nbr[item * 6 + axis * 2 + side] : resolveNeighbor(c, mine, axis * 2 + side);
    }
    bool updA = false, updB = false;  // A = low block's hi face, B = high block's lo face
    if (mine >= 0 && other >= 0) {
      // membership of the neighbour and the face voxels are fetched in the same round trip
      const int other_stamp = __ldcg(stamp_cur + other);
      unsigned int* blkA = esdfBlockPtr(c.esdf, side == 0 ? mine : other);
      unsigned int* blkB = esdfBlockPtr(c.esdf, side == 0 ? other : mine);
      VoxelRegs A[2], B[2];
      unsigned int *gHi[2], *gLo[2];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int f = lane + 32 * h, u = f >> 3, w = f & 7;
        const int faceBase = (axis == 0) ? (u * 8 + w) : ((axis == 1) ? (u * 64 + w) : (u * 64 + w * 8));
        gHi[h] = blkA + (faceBase + (kVps - 1) * strideA) * kEsdfVoxelWords;
        gLo[h] = blkB + faceBase * kEsdfVoxelWords;
        A[h] = loadVoxel(gHi[h]);
        B[h] = loadVoxel(gLo[h]);
      }
#pragma unroll
      for (int h = 0; h < 2; h++) {
        if (side == 0) {
          updB |= updateSingleNeighbor(A[h], B[h], gLo[h], axis, +1, c.max_sq);  // P: mine -> mine + d
          if (other_stamp == ring) updA |= updateSingleNeighbor(B[h], A[h], gHi[h], axis, -1, c.max_sq);  // Q
        } else if (other_stamp != ring) {
          updA |= updateSingleNeighbor(B[h], A[h], gHi[h], axis, -1, c.max_sq);  // Q: mine -> mine - d
        }
      }
    }
    updA = __any_sync(0xffffffffu, updA);
    updB = __any_sync(0xffffffffu, updB);
    if (lane == 0 && mine >= 0 && other >= 0) {
      const int slotA = side == 0 ? mine : other, slotB = side == 0 ? other : mine;
      // Record the updated blocks; they are appended to ring+1 once per ring (flushPending), so the two
      // dependent L2 atomics of the unique append are paid once instead of in each of the three axis phases.
      if (updA) {
        const int q = atomicAdd(&sh.npend, 1);
        if (q < kPendMax) sh.pend[q] = slotA;
        else if (atomicExch(stamp_nxt + slotA, ring + 1) != ring + 1) list_nxt[atomicAdd(count_nxt, 1)] = slotA;
      }
      if (updB) {
        const int q = atomicAdd(&sh.npend, 1);
        if (q < kPendMax) sh.pend[q] = slotB;
        else if (atomicExch(stamp_nxt + slotB, ring + 1) != ring + 1) list_nxt[atomicAdd(count_nxt, 1)] = slotB;
      }
    }
  }
}

// 96 registers x 512 threads = 3/4 of the register file: the wavefront runs on a side stream and must
// leave room for the next frame's raycast / compaction / TSDF CTAs on the same SM.
// Unique append of the recorded blocks to ring+1: the stamp (atomicExch) dedupes across CTAs and doubles as the
// membership flag of ring+1; each warp reserves its range of the list with one atomicAdd.
__device__ NVB_WAVE_FN void flushPending(WaveShared& sh, int* stamp_nxt, int ring, int* list_nxt, int* count_nxt) {
  __syncthreads();
  const int tid = threadIdx.x, lane = tid & 31;
  const int np = sh.npend < kPendMax ? sh.npend : kPendMax;
  for (int base = 0; base < np; base += kWT) {
    const int q = base + tid;
    int slot = -1;
    bool fresh = false;
    if (q < np) {
      slot = sh.pend[q];
      fresh = atomicExch(stamp_nxt + slot, ring + 1) != ring + 1;
    }
    const unsigned int ballot = __ballot_sync(0xffffffffu, fresh);
    if (ballot) {
      int basepos = 0;
      if (lane == 0) basepos = atomicAdd(count_nxt, __popc(ballot));
      basepos = __shfl_sync(0xffffffffu, basepos, 0);
      if (fresh) list_nxt[basepos + __popc(ballot & ((1u << lane) - 1u))] = slot;
    }
  }
  __syncthreads();
  if (tid == 0) sh.npend = 0;
  __syncthreads();
}

#ifndef NVB_WAVE_MAXREG
#define NVB_WAVE_MAXREG 128
#endif
__global__ void __maxnreg__(NVB_WAVE_MAXREG) esdfWaveKernel(EsdfCtx c) {
  extern __shared__ __align__(16) unsigned int smem[];
  __shared__ WaveShared sh;
  const int cta = blockIdx.x, nctas = gridDim.x;
  // Empty block list: integrateBlocksTemplate returns before touching anything (:226-228).
  if (*(volatile int*)c.work_count == 0) return;
  if (threadIdx.x == 0) sh.npend = 0;
  __syncthreads();
  unsigned int generation = 0;
  int ring = *(volatile int*)c.ring_id;
  int* stamp[2] = {c.stamp_a, c.stamp_b};
  int* list[2] = {c.ring_a, c.ring_b};
  long long swept = 0, faces = 0, rings = 0;
  // CTA 0 keeps a coarse time split (ns): barriers (incl. waiting for the slowest CTA), axis phases,
  // sweep phases
  long long t_bar = 0, t_axis = 0, t_sweep = 0, n_bar = 0, t0 = globalTimerNs(), t1;
#define NVB_TICK(acc)   \
  t1 = globalTimerNs(); \
  acc += t1 - t0;       \
  t0 = t1;
  // every CTA: duration of its own work in the phase that ends at barrier #n_bar -> max over CTAs
  long long tw0 = globalTimerNs();
#define NVB_PHASE_MAX()                                                                                   \
  if (threadIdx.x == 0 && n_bar < 1000) {                                                                 \
    atomicMax((unsigned long long*)c.phase_max + n_bar, (unsigned long long)(globalTimerNs() - tw0));     \
  }
#define NVB_PHASE_BEGIN() tw0 = globalTimerNs();
  for (int pass = 0; pass < 2; pass++) {
    // pass 0: blocks with sites; pass 1: the persistent cleared list (:254-257)
    const int* src = pass ? c.cleared_list : c.upd_list;
    int n = pass ? *(volatile int*)c.cleared_count : *(volatile int*)c.upd_count;
    if (n == 0) continue;
    int ci = ring & 1;
    const int* cur = src;  // the first ring's members are read straight from the source list
    auto share = [&](int count) { return (count > cta) ? (count - cta + nctas - 1) / nctas : 0; };
    auto roundsOf = [&](int count) { return ((count + nctas - 1) / nctas + kWaveMaxMembers - 1) / kWaveMaxMembers; };
    // Initial sweep of the source list; its members are stamped as ring `ring`.
    {
      const int rounds = roundsOf(n);
      for (int r = 0; r < rounds; r++) {
        const int k = loadMembers(sh, cur, n, cta, nctas, r * kWaveMaxMembers, stamp[ci], ring);
        sweepMembers(c, sh, k, smem, rounds == 1);
      }
    }
    if (cta == 0 && threadIdx.x == 0) c.ring_count[ci ^ 1] = 0;
    NVB_TICK(t_sweep)
    NVB_PHASE_MAX()
    gridBarrier(c.barrier, generation, nctas);
    NVB_TICK(t_bar)
    n_bar++;
    NVB_PHASE_BEGIN()
    swept += n;
    while (n > 0) {
      const int ni = ci ^ 1;
      const int rounds = roundsOf(n);
#pragma unroll 1
      for (int axis = 0; axis < 3; axis++) {
        for (int r = 0; r < rounds; r++) {
          int k = share(n);
          if (rounds > 1) {
            k = loadMembers(sh, cur, n, cta, nctas, r * kWaveMaxMembers, nullptr, 0);
            prefetchNeighbors(c, sh, k);
            __syncthreads();
          }
          axisMembers(c, sh, axis, k, stamp[ci], ring, stamp[ni], list[ni], c.ring_count + ni);
          if (rounds > 1) flushPending(sh, stamp[ni], ring, list[ni], c.ring_count + ni);
        }
        if (axis == 2 && rounds == 1) flushPending(sh, stamp[ni], ring, list[ni], c.ring_count + ni);
        NVB_TICK(t_axis)
        NVB_PHASE_MAX()
        gridBarrier(c.barrier, generation, nctas);
        NVB_TICK(t_bar)
        n_bar++;
        NVB_PHASE_BEGIN()
      }
      faces += 6ll * n;
      // ring+1 = the blocks appended during the three axis phases
      const int n_next = *(volatile int*)(c.ring_count + ni);
      {
        const int rounds_next = roundsOf(n_next);
        for (int r = 0; r < rounds_next; r++) {
          const int k = loadMembers(sh, list[ni], n_next, cta, nctas, r * kWaveMaxMembers, nullptr, 0);
          sweepMembers(c, sh, k, smem, rounds_next == 1);
        }
      }
      if (cta == 0 && threadIdx.x == 0) c.ring_count[ci] = 0;  // becomes the append counter of ring+2
      NVB_TICK(t_sweep)
      NVB_PHASE_MAX()
      gridBarrier(c.barrier, generation, nctas);
      NVB_TICK(t_bar)
      n_bar++;
      NVB_PHASE_BEGIN()
      swept += n_next;
      rings++;
      ring++;
      ci = ni;
      cur = list[ni];
      n = n_next;
    }
    ring++;
    if (cta == 0 && threadIdx.x == 0) c.ring_count[0] = c.ring_count[1] = 0;
    NVB_PHASE_MAX()
    gridBarrier(c.barrier, generation, nctas);
    NVB_TICK(t_bar)
    n_bar++;
    NVB_PHASE_BEGIN()
  }
#undef NVB_TICK
  if (cta == 0 && threadIdx.x == 0) {
    *c.ring_id = ring + 1;
    c.stats[4] = *(volatile int*)c.cleared_count;
    c.stats[5] = swept, c.stats[6] = faces, c.stats[7] = rings;
    c.stats[8] = t_bar, c.stats[9] = t_axis, c.stats[10] = t_sweep, c.stats[11] = n_bar;
    long long sum_max = 0;
    for (int q = 0; q < n_bar && q < 1000; q++) sum_max += (long long)c.phase_max[q];
    c.stats[12] = sum_max;  // sum over phases of the slowest CTA's work time
  }
}

}  // namespace

int esdfPersistentMaxCtas(int num_sms) {
  static int per_sm = -1;
  if (per_sm < 0) {
    cudaFuncSetAttribute(esdfWaveKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWaveSmemBytes);
    int v = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, esdfWaveKernel, kWT, kWaveSmemBytes) != cudaSuccess) v = 0;
    per_sm = v;
  }
  return per_sm * num_sms;
}

cudaError_t launchEsdfComputePersistent(const EsdfCtx& c, int num_sms, cudaStream_t stream, int* launches) {
  const int max_ctas = esdfPersistentMaxCtas(num_sms);
  if (max_ctas <= 0) return cudaErrorLaunchOutOfResources;
  // One CTA per SM: the wavefront is latency-bound, more CTAs only make the barrier slower.
  int grid = num_sms < max_ctas ? num_sms : max_ctas;
  EsdfCtx cc = c;
  void* args[] = {&cc};
  (*launches)++;
  return cudaLaunchCooperativeKernel((const void*)esdfWaveKernel, dim3(grid), dim3(kWT), args, kWaveSmemBytes, stream);
}

}  // namespace nvb
